"""Oracle: SAM3 ViT trunk (teacher encoder) -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Functional fp32 restatement driven by a reference-keyed state_dict:
  compute_axial_cis / apply_rotary_enc   sam3/sam3/model/vitdet.py:41-57, 68-90
  window_partition / unpartition         vitdet.py:93-139
  get_abs_pos (tiling branch)            vitdet.py:175-236
  PatchEmbed                             vitdet.py:299-336
  Attention.forward                      vitdet.py:466-515 (rope setup :421-457)
  Block.forward                          vitdet.py:597-613
  ViT.forward                            vitdet.py:813-859; SAM3 config model_builder.py:70-97
  timm Mlp                               fc1 -> GELU(erf) -> fc2 (vitdet.py:585-590)
Only the configuration family SAM3 uses is restated: no cls token retained, tiled abs-pos with a cls row
in the table, ln_pre, no ln_post, rope with interpolation (scale = window / input size), no rel-pos.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SAM3_VIT = dict(img_size=1008, pretrain_img_size=336, patch_size=14, embed_dim=1024, depth=32, num_heads=16,
                mlp_ratio=4.625, window_size=24, global_att_blocks=(7, 15, 23, 31))


def axial_cis(head_dim, end_x, end_y, theta=10000.0, scale_pos=1.0):
    """Complex rotation table [end_x*end_y, head_dim/2]: first half indexed by x, second half by y."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 4)[: head_dim // 4].float() / head_dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx = (t % end_x).float() * scale_pos
    ty = torch.div(t, end_x, rounding_mode="floor").float() * scale_pos
    fx, fy = torch.outer(tx, freqs), torch.outer(ty, freqs)
    return torch.cat([torch.polar(torch.ones_like(fx), fx), torch.polar(torch.ones_like(fy), fy)], dim=-1)


def rope(x, cis):
    """x: [B, heads, L, D] real -> rotated, pairs (2i, 2i+1) form complex numbers."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(xc * cis.view(1, 1, *cis.shape)).flatten(3).type_as(x)     # vitdet.py:86-89: .type_as(xq)


def attention(sd, p, x, num_heads, cis):
    """x: [B', H', W', C] (a window or the full map)."""
    Bp, H, W, C = x.shape
    L = H * W
    qkv = F.linear(x, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(Bp, L, 3, num_heads, -1)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    q, k = rope(q, cis), rope(k, cis)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.view(Bp, num_heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(Bp, H, W, C)
    return F.linear(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def block(sd, p, x, num_heads, window, cis):
    B, H, W, C = x.shape
    y = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-5)
    if window > 0:
        assert H % window == 0 and W % window == 0, "oracle restates the no-padding case SAM3 uses (72 = 3 x 24)"
        y = y.view(B, H // window, window, W // window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, C)
    y = attention(sd, p + ".attn", y, num_heads, cis)
    if window > 0:
        y = y.view(B, H // window, W // window, window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    x = x + y
    y = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps=1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])), sd[p + ".mlp.fc2.weight"],
                 sd[p + ".mlp.fc2.bias"])
    return x + y


def abs_pos_tiled(pos_embed, h, w):
    """pos_embed [1, 1+s*s, C] (cls row first) -> [1,h,w,C] by tiling the s x s table (vitdet.py:205-214)."""
    tab = pos_embed[:, 1:]
    s = int(math.sqrt(tab.shape[1]))
    if s == h and s == w:
        return tab.reshape(1, h, w, -1)
    t = tab.reshape(1, s, s, -1).permute(0, 3, 1, 2)
    t = t.tile([1, 1, h // s + 1, w // s + 1])[:, :, :h, :w]
    return t.permute(0, 2, 3, 1)


def vit_trunk(sd, p, x, cfg=SAM3_VIT, return_blocks=False):
    """x: [B,3,S,S] fp32 -> [B, C, S/patch, S/patch] (the last global block's output, NCHW)."""
    ps, C, heads = cfg["patch_size"], cfg["embed_dim"], cfg["num_heads"]
    win, glob = cfg["window_size"], tuple(cfg["global_att_blocks"])
    x = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd.get(p + "patch_embed.proj.bias"), stride=ps).permute(0, 2, 3, 1)
    h, w = x.shape[1], x.shape[2]
    x = x + abs_pos_tiled(sd[p + "pos_embed"], h, w)
    x = F.layer_norm(x, (C,), sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], eps=1e-5)
    hd = C // heads
    cis_win = axial_cis(hd, win, win, scale_pos=1.0).to(x.device)    # rope_pt_size == window -> scale 1
    cis_glob = axial_cis(hd, h, w, scale_pos=win / h).to(x.device)   # interpolated: scale = rope_pt / input
    outs = []
    for i in range(cfg["depth"]):
        g = i in glob
        x = block(sd, f"{p}blocks.{i}", x, heads, 0 if g else win, cis_glob if g else cis_win)
        if return_blocks:
            outs.append(x)
    out = x.permute(0, 3, 1, 2)
    return (out, outs) if return_blocks else out
