"""Oracle: EfficientViT backbone + student head (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

Functional fp32 restatement, driven by a reference-keyed state_dict `sd`:
  ConvLayer        sam3/sam3/backbones/efficientvit/nn/ops.py:39-80
  DSConv / MBConv  ops.py:273-312 / 315-367
  LiteMLA          ops.py:521-671
  EfficientViTBlock / ResidualBlock   ops.py:674-732 / 740-770
  EfficientViTBackbone                efficientvit/backbone.py:32-156, variants :158-196
  ImageStudentEncoder (head + resize) stage1/model.py:188-211
Eval-mode semantics (BatchNorm uses running statistics) by default; inside `with bn_batch_stats():` every BatchNorm2d
behaves as in module.train() (batch statistics, running buffers of `sd` updated in place with momentum 0.1) -- the
train-mode forward whose autograd is the oracle for the student backward (stage1/train_image_encoder_stage1.py:154-268).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

VARIANTS = {  # backbone.py:158-196
    "b0": dict(width_list=[8, 16, 32, 64, 128], depth_list=[1, 2, 2, 2, 2], dim=16),
    "b1": dict(width_list=[16, 32, 64, 128, 256], depth_list=[1, 2, 3, 3, 4], dim=16),
    "b2": dict(width_list=[24, 48, 96, 192, 384], depth_list=[1, 3, 4, 4, 6], dim=32),
    "b3": dict(width_list=[32, 64, 128, 256, 512], depth_list=[1, 4, 6, 6, 9], dim=32),
}
BN_EPS = 1e-5
_BN_TRAIN = False


class bn_batch_stats:
    """Context manager: nn.BatchNorm2d train-mode semantics for every norm in this file."""

    def __enter__(self):
        global _BN_TRAIN
        self.prev, _BN_TRAIN = _BN_TRAIN, True

    def __exit__(self, *exc):
        global _BN_TRAIN
        _BN_TRAIN = self.prev


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=_BN_TRAIN, momentum=0.1, eps=BN_EPS)


def conv_layer(sd, p, x, *, stride=1, groups=1, act=None):
    """conv (same padding) -> BatchNorm2d if `p.norm.*` exists -> activation.  ops.py:39-80."""
    w = sd[p + ".conv.weight"]
    b = sd.get(p + ".conv.bias")
    k = w.shape[-1]
    x = F.conv2d(x, w, b, stride=stride, padding=k // 2, groups=groups)
    if (p + ".norm.weight") in sd:
        x = _bn(x, sd, p + ".norm")
    if act == "hswish":
        x = F.hardswish(x)
    elif act is not None:
        raise ValueError(act)
    return x


def dsconv(sd, p, x, stride):
    c = x.shape[1]
    x = conv_layer(sd, p + ".depth_conv", x, stride=stride, groups=c, act="hswish")
    return conv_layer(sd, p + ".point_conv", x)


def mbconv(sd, p, x, stride):
    x = conv_layer(sd, p + ".inverted_conv", x, act="hswish")
    x = conv_layer(sd, p + ".depth_conv", x, stride=stride, groups=x.shape[1], act="hswish")
    return conv_layer(sd, p + ".point_conv", x)


def lite_mla(sd, p, x, dim, eps=1.0e-15):
    """ops.py:584-671: multi-scale qkv, ReLU kernel, linear (HW > dim) or quadratic attention."""
    qkv = conv_layer(sd, p + ".qkv", x)
    c3 = qkv.shape[1]
    heads = c3 // (3 * dim)
    agg = F.conv2d(qkv, sd[p + ".aggreg.0.0.weight"], sd.get(p + ".aggreg.0.0.bias"), padding=2, groups=c3)
    agg = F.conv2d(agg, sd[p + ".aggreg.0.1.weight"], sd.get(p + ".aggreg.0.1.bias"), groups=3 * heads)
    ms = torch.cat([qkv, agg], dim=1)
    B, _, H, W = ms.shape
    # ops.py:586-589 / 627-628: the attention runs with autocast disabled, fp16 inputs promoted to fp32 (a no-op on the fp32 CPU path)
    with torch.autocast(device_type=ms.device.type, enabled=False):
        t = ms.float() if ms.dtype == torch.float16 else ms
        t = t.reshape(B, -1, 3 * dim, H * W)
        q, k, v = t[:, :, :dim], t[:, :, dim:2 * dim], t[:, :, 2 * dim:]
        q, k = F.relu(q), F.relu(k)
        if H * W > dim:
            v1 = F.pad(v, (0, 0, 0, 1), mode="constant", value=1.0)
            out = torch.matmul(torch.matmul(v1, k.transpose(-1, -2)), q)
            out = out[:, :, :-1] / (out[:, :, -1:] + eps)
        else:
            att = torch.matmul(k.transpose(-1, -2), q)
            att = att / (att.sum(dim=2, keepdim=True) + eps)
            out = torch.matmul(v, att)
    out = out.to(ms.dtype)
    out = out.reshape(B, -1, H, W)
    return conv_layer(sd, p + ".proj", out)


def backbone(sd, p, x, variant="b1", return_stages=False):
    cfg = VARIANTS[variant]
    widths, depths, dim = cfg["width_list"], cfg["depth_list"], cfg["dim"]
    stages = {}
    x = conv_layer(sd, f"{p}input_stem.op_list.0", x, stride=2, act="hswish")
    for i in range(depths[0]):
        x = x + dsconv(sd, f"{p}input_stem.op_list.{1 + i}.main", x, 1)
    stages["stage0"] = x
    sid = 0
    for w, d in zip(widths[1:3], depths[1:3]):
        for i in range(d):
            y = mbconv(sd, f"{p}stages.{sid}.op_list.{i}.main", x, 2 if i == 0 else 1)
            x = y if i == 0 else x + y
        sid += 1
        stages[f"stage{sid}"] = x
    for w, d in zip(widths[3:], depths[3:]):
        x = mbconv(sd, f"{p}stages.{sid}.op_list.0.main", x, 2)
        for i in range(d):
            q = f"{p}stages.{sid}.op_list.{1 + i}"
            x = x + lite_mla(sd, q + ".context_module.main", x, dim)
            x = x + mbconv(sd, q + ".local_module.main", x, 1)
        sid += 1
        stages[f"stage{sid}"] = x
    return (x, stages) if return_stages else x


def student_head(sd, feats, embed_size):
    """stage1/model.py:194-211: Conv1x1(no bias) -> BN -> GELU(erf) -> Conv3x3(pad 1, bias) -> bilinear."""
    x = F.conv2d(feats, sd["head.0.weight"])
    x = _bn(x, sd, "head.1")
    x = F.gelu(x)
    x = F.conv2d(x, sd["head.3.weight"], sd["head.3.bias"], padding=1)
    if x.shape[-1] != embed_size or x.shape[-2] != embed_size:
        x = F.interpolate(x, size=(embed_size, embed_size), mode="bilinear", align_corners=False)
    return x


def image_student_encoder(sd, x, embed_size, variant="b1", return_stages=False):
    """ImageStudentEncoder(EfficientViTAdapter(backbone)).forward  (stage1/model.py:201-211, 327-335)."""
    feats, stages = backbone(sd, "backbone.model.", x, variant, return_stages=True)
    out = student_head(sd, feats, embed_size)
    return (out, stages) if return_stages else out
