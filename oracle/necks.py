"""Oracle: Sam3DualViTDetNeck (SimpleFPN, dual SAM3 / SAM2 branches) -- TEST INFRASTRUCTURE ONLY.
Restates sam3/sam3/model/necks.py:13-125 on a reference-keyed state_dict (prefix `convs.` / `sam2_convs.`)."""
from __future__ import annotations

import torch.nn.functional as F


def branch(sd, p, x, scale):
    if scale == 4.0:
        x = F.conv_transpose2d(x, sd[p + "dconv_2x2_0.weight"], sd[p + "dconv_2x2_0.bias"], stride=2)
        x = F.gelu(x)
        x = F.conv_transpose2d(x, sd[p + "dconv_2x2_1.weight"], sd[p + "dconv_2x2_1.bias"], stride=2)
    elif scale == 2.0:
        x = F.conv_transpose2d(x, sd[p + "dconv_2x2.weight"], sd[p + "dconv_2x2.bias"], stride=2)
    elif scale == 0.5:
        x = F.max_pool2d(x, 2, 2)
    x = F.conv2d(x, sd[p + "conv_1x1.weight"], sd[p + "conv_1x1.bias"])
    return F.conv2d(x, sd[p + "conv_3x3.weight"], sd[p + "conv_3x3.bias"], padding=1)


def neck(sd, x, scale_factors=(4.0, 2.0, 1.0, 0.5), prefix="convs."):
    return [branch(sd, f"{prefix}{i}.", x, s) for i, s in enumerate(scale_factors)]
