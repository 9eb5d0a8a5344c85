"""Oracle: RepViT backbone as used by the stage-1 student (TEST INFRASTRUCTURE ONLY).
Restates sam3/sam3/backbones/repvit.py: Conv2d_BN :27-49, Residual :51-81, RepVGGDW :84-122, RepViTBlock :125-161,
RepViT.features :219-246, configs :253-384; SqueezeExcite = timm.layers.SqueezeExcite(inp, 0.25) (un-vendored;
mean_HW -> fc1 (1x1, bias) -> ReLU -> fc2 (1x1, bias) -> sigmoid gate); RepViTAdapter stage1/model.py:287-296.
Un-fused (exactly what the reference module executes); eval-mode BatchNorm by default, batch statistics inside
`oracle.efficientvit.bn_batch_stats()` (the train-mode forward whose autograd is the oracle of the RepViT backward)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

def _stage_cfgs(widths, depths):
    """[k, t, c, SE, HS, s] rows of repvit.py:280-520: SE alternates inside a stage (stage 1 starts with SE, later stages
    with the stride-2 block that has none), never on a stage's last block; HS flag on the last two stages."""
    rows = []
    for si, (c, n) in enumerate(zip(widths, depths)):
        for i in range(n):
            se = ((i % 2 == 0) if si == 0 else (i % 2 == 1)) and i != n - 1
            rows.append([3, 2, c, int(se), int(si >= 2), 2 if (si > 0 and i == 0) else 1])
    return rows


CFGS = {
    "repvit_m0_9": _stage_cfgs((48, 96, 192, 384), (3, 4, 16, 3)),      # repvit.py:280-313
    "repvit_m1_1": _stage_cfgs((64, 128, 256, 512), (3, 4, 14, 3)),     # repvit.py:353-384
    "repvit_m2_3": _stage_cfgs((80, 160, 320, 640), (7, 8, 36, 3)),     # repvit.py:442-505
}


def conv_bn(sd, p, x, stride=1, pad=0, groups=1):
    from .efficientvit import _bn
    x = F.conv2d(x, sd[p + ".c.weight"], None, stride=stride, padding=pad, groups=groups)
    return _bn(x, sd, p + ".bn")


def squeeze_excite(sd, p, x):
    s = x.mean((2, 3), keepdim=True)
    s = F.relu(F.conv2d(s, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
    s = F.conv2d(s, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])
    return x * torch.sigmoid(s)


def block(sd, p, x, stride, use_se):
    c = x.shape[1]
    if stride == 2:
        x = conv_bn(sd, p + ".token_mixer.0", x, stride=2, pad=1, groups=c)
        if use_se:
            x = squeeze_excite(sd, p + ".token_mixer.1", x)
        x = conv_bn(sd, p + ".token_mixer.2", x)
    else:
        q = p + ".token_mixer.0"
        y = conv_bn(sd, q + ".conv", x, pad=1, groups=c) + F.conv2d(x, sd[q + ".conv1.weight"], sd[q + ".conv1.bias"], groups=c) + x
        from .efficientvit import _bn
        x = _bn(y, sd, q + ".bn")
        if use_se:
            x = squeeze_excite(sd, p + ".token_mixer.1", x)
    m = p + ".channel_mixer.m"
    return x + conv_bn(sd, m + ".2", F.gelu(conv_bn(sd, m + ".0", x)))


def backbone(sd, p, x, variant="repvit_m1_1", return_blocks=False):
    """RepViTAdapter.forward: iterate model.features (stage1/model.py:293-296)."""
    f = p + "features."
    x = conv_bn(sd, f + "0.0", x, stride=2, pad=1)
    x = conv_bn(sd, f + "0.2", F.gelu(x), stride=2, pad=1)
    outs = []
    for i, (k, t, c, se, hs, s) in enumerate(CFGS[variant], 1):
        x = block(sd, f"{f}{i}", x, s, bool(se))
        outs.append(x)
    return (x, outs) if return_blocks else x


def image_student_encoder(sd, x, embed_size, variant="repvit_m1_1"):
    from .efficientvit import student_head
    return student_head(sd, backbone(sd, "backbone.model.", x, variant), embed_size)
