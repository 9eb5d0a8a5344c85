"""Oracle: TinyViT backbone as used by the stage-1 student "TV-M" (TEST INFRASTRUCTURE ONLY).
Restates sam3/sam3/backbones/tiny_vit.py: Conv2d_BN :29-53, PatchEmbed :67-84, MBConv :87-125, PatchMerging :128-154,
ConvLayer :157-193, Mlp :196-216, Attention :219-293 (learned relative bias table indexed by |dy|,|dx|),
TinyViTBlock :296-386 (zero padding to a window multiple BEFORE the attention LayerNorm, so padded tokens enter the
softmax as LN(0) = beta), BasicLayer :393-454, TinyViT :460-607, tiny_vit_11m_224 :669-679;
TinyViTAdapter stage1/model.py:299-324.  DropPath = identity (eval mode, or training with drop_path_rate 0: the
stochastic-depth draws are not part of the oracle)."""
from __future__ import annotations

import itertools

import torch
import torch.nn.functional as F

_COMMON = dict(depths=[2, 2, 6, 2], window_sizes=[7, 7, 14, 7], mlp_ratio=4.0, mbconv_expand_ratio=4.0)
VARIANTS = {  # tiny_vit.py:656-692
    "tiny_vit_5m": dict(embed_dims=[64, 128, 160, 320], num_heads=[2, 4, 5, 10], **_COMMON),
    "tiny_vit_11m": dict(embed_dims=[64, 128, 256, 448], num_heads=[2, 4, 8, 14], **_COMMON),
    "tiny_vit_21m": dict(embed_dims=[96, 192, 384, 576], num_heads=[3, 6, 12, 18], **_COMMON),
}


def conv_bn(sd, p, x, stride=1, pad=0, groups=1):
    from .efficientvit import _bn      # eval-mode BN by default, batch statistics inside efficientvit.bn_batch_stats()
    x = F.conv2d(x, sd[p + ".c.weight"], None, stride=stride, padding=pad, groups=groups)
    return _bn(x, sd, p + ".bn")


def mbconv(sd, p, x):
    y = F.gelu(conv_bn(sd, p + ".conv1", x))
    y = F.gelu(conv_bn(sd, p + ".conv2", y, pad=1, groups=y.shape[1]))
    y = conv_bn(sd, p + ".conv3", y)
    return F.gelu(y + x)


def patch_merging(sd, p, x, hw):
    if x.dim() == 3:
        B = x.shape[0]
        x = x.view(B, hw[0], hw[1], -1).permute(0, 3, 1, 2)
    x = F.gelu(conv_bn(sd, p + ".conv1", x))
    x = F.gelu(conv_bn(sd, p + ".conv2", x, stride=2, pad=1, groups=x.shape[1]))
    x = conv_bn(sd, p + ".conv3", x)
    return x.flatten(2).transpose(1, 2)


def bias_index(ws):
    pts = list(itertools.product(range(ws), range(ws)))
    offs, idxs = {}, []
    for p1 in pts:
        for p2 in pts:
            o = (abs(p1[0] - p2[0]), abs(p1[1] - p2[1]))
            if o not in offs:
                offs[o] = len(offs)
            idxs.append(offs[o])
    return torch.tensor(idxs, dtype=torch.long).view(len(pts), len(pts))


def attention(sd, p, x, heads, ws):
    B, N, C = x.shape
    kd = C // heads
    x = F.layer_norm(x, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    qkv = F.linear(x, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).view(B, N, heads, -1)
    q, k, v = qkv.split([kd, kd, kd], dim=3)
    q, k, v = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
    bias = sd[p + ".attention_biases"][:, bias_index(ws)]
    a = (q @ k.transpose(-2, -1)) * kd ** -0.5 + bias
    o = (a.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def block(sd, p, x, hw, heads, ws):
    H, W = hw
    B, L, C = x.shape
    res = x
    if H == ws and W == ws:
        x = attention(sd, p + ".attn", x, heads, ws)
    else:
        x = x.view(B, H, W, C)
        pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
        if pb or pr:
            x = F.pad(x, (0, 0, 0, pr, 0, pb))
        pH, pW = H + pb, W + pr
        nH, nW = pH // ws, pW // ws
        x = x.view(B, nH, ws, nW, ws, C).transpose(2, 3).reshape(B * nH * nW, ws * ws, C)
        x = attention(sd, p + ".attn", x, heads, ws)
        x = x.view(B, nH, nW, ws, ws, C).transpose(2, 3).reshape(B, pH, pW, C)[:, :H, :W].contiguous().view(B, L, C)
    x = res + x
    x = x.transpose(1, 2).reshape(B, C, H, W)
    x = conv_bn(sd, p + ".local_conv", x, pad=1, groups=C).view(B, C, L).transpose(1, 2)
    m = p + ".mlp"
    y = F.layer_norm(x, (C,), sd[m + ".norm.weight"], sd[m + ".norm.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[m + ".fc1.weight"], sd[m + ".fc1.bias"])), sd[m + ".fc2.weight"], sd[m + ".fc2.bias"])
    return x + y


def backbone(sd, p, x, variant="tiny_vit_11m"):
    """TinyViTAdapter.forward: patch_embed, layers, reshape to [B,C,h,w] (stage1/model.py:309-317)."""
    cfg = VARIANTS[variant]
    x = conv_bn(sd, p + "patch_embed.seq.0", x, stride=2, pad=1)
    x = conv_bn(sd, p + "patch_embed.seq.2", F.gelu(x), stride=2, pad=1)
    hw = (x.shape[2], x.shape[3])
    for i in range(cfg["depths"][0]):
        x = mbconv(sd, f"{p}layers.0.blocks.{i}", x)
    x = patch_merging(sd, p + "layers.0.downsample", x, hw)
    hw = ((hw[0] - 1) // 2 + 1, (hw[1] - 1) // 2 + 1)
    for li in range(1, len(cfg["depths"])):
        for i in range(cfg["depths"][li]):
            x = block(sd, f"{p}layers.{li}.blocks.{i}", x, hw, cfg["num_heads"][li], cfg["window_sizes"][li])
        if li < len(cfg["depths"]) - 1:
            x = patch_merging(sd, f"{p}layers.{li}.downsample", x, hw)
            hw = ((hw[0] - 1) // 2 + 1, (hw[1] - 1) // 2 + 1)
    B, N, C = x.shape
    return x.view(B, hw[0], hw[1], C).permute(0, 3, 1, 2).contiguous()


def image_student_encoder(sd, x, embed_size, variant="tiny_vit_11m"):
    from .efficientvit import student_head
    return student_head(sd, backbone(sd, "backbone.model.", x, variant), embed_size)
