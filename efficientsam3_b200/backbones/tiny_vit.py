"""TinyViT backbone (stage-1 student "TV-M" = tiny_vit_11m), B200-native.  Module tree / state_dict keys of
sam3/sam3/backbones/tiny_vit.py (Conv2d_BN :29-53, PatchEmbed :67-84, MBConv :87-125, PatchMerging :128-154,
ConvLayer :157-193, Mlp :196-216, Attention :219-293, TinyViTBlock :296-386, BasicLayer :393-454, TinyViT :460-607).

Eval-mode execution (token stream NHWC bf16 = raster tokens x channels):
  patch embed     es3_stem_conv3x3_s2 + es3_conv3x3_s2_narrow_bf16
  MBConv          es3_gemm_bf16_ex (conv1+BN+GELU) -> es3_dwconv_tiled_bf16 (+GELU) -> es3_gemm_bf16_ex (conv3+BN, +x, GELU
                  after the residual)
  PatchMerging    gemm+GELU -> depthwise s2 + GELU -> gemm
  TinyViTBlock    es3_layernorm_bf16 -> qkv gemm -> es3_win_attn_bias_bf16 (in-place window gather, zero-padded windows
                  reproduced with the constant qkv(LN(0)) token) -> proj gemm + residual -> depthwise local_conv ->
                  es3_layernorm_bf16 -> fc1 gemm + GELU -> fc2 gemm + residual
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

from .. import ops
from ..nn_utils import NativePlanMixin, bn_scale_bias, dw_weight, pack_patch_embed, pw_weight


class Conv2d_BN(nn.Sequential):
    def __init__(self, a, b, ks=1, stride=1, pad=0, dilation=1, groups=1, bn_weight_init=1):
        super().__init__()
        self.add_module("c", nn.Conv2d(a, b, ks, stride, pad, dilation, groups, bias=False))
        bn = nn.BatchNorm2d(b)
        nn.init.constant_(bn.weight, bn_weight_init)
        nn.init.constant_(bn.bias, 0)
        self.add_module("bn", bn)


class PatchEmbed(nn.Module):
    def __init__(self, in_chans, embed_dim, resolution, activation):
        super().__init__()
        res = (resolution, resolution) if isinstance(resolution, int) else tuple(resolution)
        self.patches_resolution = (res[0] // 4, res[1] // 4)
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        n = embed_dim
        self.seq = nn.Sequential(Conv2d_BN(in_chans, n // 2, 3, 2, 1), activation(), Conv2d_BN(n // 2, n, 3, 2, 1))


class MBConv(nn.Module):
    def __init__(self, in_chans, out_chans, expand_ratio, activation, drop_path):
        super().__init__()
        self.in_chans, self.hidden_chans, self.out_chans = in_chans, int(in_chans * expand_ratio), out_chans
        self.conv1 = Conv2d_BN(in_chans, self.hidden_chans, ks=1)
        self.act1 = activation()
        self.conv2 = Conv2d_BN(self.hidden_chans, self.hidden_chans, ks=3, stride=1, pad=1, groups=self.hidden_chans)
        self.act2 = activation()
        self.conv3 = Conv2d_BN(self.hidden_chans, out_chans, ks=1, bn_weight_init=0.0)
        self.act3 = activation()
        self.drop_path = nn.Identity()


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, out_dim, activation):
        super().__init__()
        self.input_resolution, self.dim, self.out_dim = input_resolution, dim, out_dim
        self.act = activation()
        self.conv1 = Conv2d_BN(dim, out_dim, 1, 1, 0)
        self.conv2 = Conv2d_BN(out_dim, out_dim, 3, 2, 1, groups=out_dim)
        self.conv3 = Conv2d_BN(out_dim, out_dim, 1, 1, 0)


class ConvLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, activation, drop_path=0.0, downsample=None, use_checkpoint=False,
                 out_dim=None, conv_expand_ratio=4.0):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([MBConv(dim, dim, conv_expand_ratio, activation, 0.0) for _ in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, out_dim=out_dim, activation=activation) if downsample else None


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.norm = nn.LayerNorm(in_features)
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.act = act_layer()
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    def __init__(self, dim, key_dim, num_heads=8, attn_ratio=4, resolution=(14, 14)):
        super().__init__()
        self.num_heads, self.scale, self.key_dim = num_heads, key_dim ** -0.5, key_dim
        self.nh_kd = key_dim * num_heads
        self.d = int(attn_ratio * key_dim)
        self.dh = self.d * num_heads
        self.attn_ratio = attn_ratio
        self.norm = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, self.dh + self.nh_kd * 2)
        self.proj = nn.Linear(self.dh, dim)
        points = list(itertools.product(range(resolution[0]), range(resolution[1])))
        offsets, idxs = {}, []
        for p1 in points:
            for p2 in points:
                off = (abs(p1[0] - p2[0]), abs(p1[1] - p2[1]))
                if off not in offsets:
                    offsets[off] = len(offsets)
                idxs.append(offsets[off])
        self.attention_biases = nn.Parameter(torch.zeros(num_heads, len(offsets)))
        self.register_buffer("attention_bias_idxs", torch.LongTensor(idxs).view(len(points), len(points)), persistent=False)


class TinyViTBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, mlp_ratio=4.0, drop=0.0, drop_path=0.0,
                 local_conv_size=3, activation=nn.GELU):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads, self.window_size = dim, input_resolution, num_heads, window_size
        assert dim % num_heads == 0
        self.drop_path = nn.Identity()
        self.attn = Attention(dim, dim // num_heads, num_heads, attn_ratio=1, resolution=(window_size, window_size))
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=activation, drop=drop)
        self.local_conv = Conv2d_BN(dim, dim, ks=local_conv_size, stride=1, pad=local_conv_size // 2, groups=dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, drop=0.0, drop_path=0.0,
                 downsample=None, use_checkpoint=False, local_conv_size=3, activation=nn.GELU, out_dim=None):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([TinyViTBlock(dim, input_resolution, num_heads, window_size, mlp_ratio, drop, 0.0,
                                                  local_conv_size, activation) for _ in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, out_dim=out_dim, activation=activation) if downsample else None


# --------------------------------------------------------------------------------------------- native plans
def _cb(cb: Conv2d_BN, dev):
    return bn_scale_bias(cb.bn, None, cb.c.out_channels, dev)


class _PW:
    def __init__(self, cb, act, dev):
        # (TinyViT keeps the BN scale in the fp32 epilogue: folding it into the bf16 weights bought 0.4 % and moved tiny_vit_5m from 1.46e-2
        #  to 1.78e-2 of the reference, against a 2e-2 tolerance -- profiles/r2af_tests.log; RepViT / EfficientViT fold it)
        self.w = pw_weight(cb.c)
        self.s, self.b = _cb(cb, dev)
        self.act = act

    def __call__(self, x, residual=None, act_after_res=False):
        B, H, W, C = x.shape
        r = residual.reshape(-1, residual.shape[-1]) if residual is not None else None
        return ops.gemm(x.reshape(-1, C), self.w, scale=self.s, bias=self.b, act=self.act, residual=r,
                        act_after_res=act_after_res).view(B, H, W, -1)


class _DW:
    def __init__(self, cb, act, dev):
        s, b = _cb(cb, dev)
        self.w, self.b, self.act = dw_weight(cb.c, s), b, act
        self.stride = cb.c.stride[0]

    def __call__(self, x):
        return ops.dwconv(x, self.w, self.b, 3, self.stride, self.act)


class _MBConvPlan:
    def __init__(self, m: MBConv, dev):
        self.c1, self.c2, self.c3 = _PW(m.conv1, "gelu", dev), _DW(m.conv2, "gelu", dev), _PW(m.conv3, "gelu", dev)

    def __call__(self, x):
        return self.c3(self.c2(self.c1(x)), residual=x, act_after_res=True)   # act3(conv3(.) + shortcut)


class _MergePlan:
    def __init__(self, m: PatchMerging, dev):
        self.c1, self.c2, self.c3 = _PW(m.conv1, "gelu", dev), _DW(m.conv2, "gelu", dev), _PW(m.conv3, None, dev)

    def __call__(self, x):
        return self.c3(self.c2(self.c1(x)))


def _lin(l: nn.Linear):
    return l.weight.detach().to(torch.bfloat16).contiguous(), l.bias.detach().float().contiguous()


class _BlockPlan:
    def __init__(self, blk: TinyViTBlock, dev):
        at = blk.attn
        if at.key_dim != 32 or blk.window_size not in (7, 14):
            raise NotImplementedError("native TinyViT attention: head_dim 32, window 7 or 14 (tiny_vit_5m/11m/21m)")
        self.ws, self.heads, self.scale = blk.window_size, at.num_heads, at.scale
        self.n1 = (at.norm.weight.detach().float().contiguous(), at.norm.bias.detach().float().contiguous(), at.norm.eps)
        self.qkv, self.proj = _lin(at.qkv), _lin(at.proj)
        self.bias = at.attention_biases.detach().float()[:, at.attention_bias_idxs].contiguous()   # [heads, N, N]
        # the reference zero-pads tokens BEFORE the attention LayerNorm (tiny_vit.py:352-360): a padded token is
        # LN(0) = beta, its qkv row is W beta + b.  One tiny CUDA-core GEMM at plan time.
        beta = self.n1[1].view(1, -1).contiguous()
        self.qkv_pad = ops.gemm_simt(beta, at.qkv.weight.detach().float().contiguous(), bias=self.qkv[1],
                                     out_dtype=torch.bfloat16).reshape(-1).contiguous()
        self.local = _DW(blk.local_conv, None, dev)
        m = blk.mlp
        self.n2 = (m.norm.weight.detach().float().contiguous(), m.norm.bias.detach().float().contiguous(), m.norm.eps)
        self.fc1, self.fc2 = _lin(m.fc1), _lin(m.fc2)

    def __call__(self, x):  # x [B,H,W,C] bf16
        B, H, W, C = x.shape
        x2 = x.view(-1, C)
        y = ops.layernorm_bf16(x2, *self.n1)
        qkv = ops.gemm(y, self.qkv[0], bias=self.qkv[1])
        a = ops.win_attn_bias(qkv, self.qkv_pad, self.bias, B, H, W, C, self.heads, self.ws, self.scale)
        x2 = ops.gemm(a, self.proj[0], bias=self.proj[1], residual=x2)
        x2 = self.local(x2.view(B, H, W, C)).view(-1, C)
        y = ops.layernorm_bf16(x2, *self.n2)
        h = ops.gemm(y, self.fc1[0], bias=self.fc1[1], act="gelu")
        return ops.gemm(h, self.fc2[0], bias=self.fc2[1], residual=x2).view(B, H, W, C)


class TinyViT(nn.Module, NativePlanMixin):
    def __init__(self, img_size=224, in_chans=3, num_classes=1000, embed_dims=[96, 192, 384, 768], depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_sizes=[7, 7, 14, 7], mlp_ratio=4.0, drop_rate=0.0, drop_path_rate=0.1,
                 use_checkpoint=False, mbconv_expand_ratio=4.0, local_conv_size=3, layer_lr_decay=1.0):
        super().__init__()
        self.num_classes, self.depths, self.num_layers, self.mlp_ratio = num_classes, depths, len(depths), mlp_ratio
        activation = nn.GELU
        self.patch_embed = PatchEmbed(in_chans=in_chans, embed_dim=embed_dims[0], resolution=img_size, activation=activation)
        self.patches_resolution = self.patch_embed.patches_resolution
        self.layers = nn.ModuleList()
        cur = self.patches_resolution
        for i in range(self.num_layers):
            kw = dict(dim=embed_dims[i], input_resolution=cur, depth=depths[i],
                      downsample=PatchMerging if i < self.num_layers - 1 else None,
                      out_dim=embed_dims[min(i + 1, len(embed_dims) - 1)], activation=activation)
            if i == 0:
                layer = ConvLayer(conv_expand_ratio=mbconv_expand_ratio, **kw)
            else:
                layer = BasicLayer(num_heads=num_heads[i], window_size=window_sizes[i], mlp_ratio=mlp_ratio, drop=drop_rate,
                                   local_conv_size=local_conv_size, **kw)
            self.layers.append(layer)
            if i < self.num_layers - 1:
                cur = ((cur[0] - 1) // 2 + 1, (cur[1] - 1) // 2 + 1)
        self.norm_head = nn.LayerNorm(embed_dims[-1]) if num_classes > 0 else nn.Identity()
        self.head = nn.Linear(embed_dims[-1], num_classes) if num_classes > 0 else nn.Identity()
        # stochastic-depth rates (tiny_vit.py:491-503: linspace over all blocks); the eval plan ignores them, the training graph
        # (backbones/tinyvit_train.py) applies timm's DropPath with them
        dpr = [float(v) for v in torch.linspace(0, drop_path_rate, sum(depths))]
        k = 0
        for layer in self.layers:
            for blk in layer.blocks:
                blk.drop_path_rate = dpr[k]
                k += 1
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _build_plan(self):
        dev = next(self.parameters()).device
        c0, c1 = self.patch_embed.seq[0], self.patch_embed.seq[2]
        s0, b0 = _cb(c0, dev)
        s1, b1 = _cb(c1, dev)
        w0, b0, w1 = pack_patch_embed(c0.c.weight, s0, b0, c1.c.weight)
        steps = [lambda x: ops.conv3x3_s2_narrow(ops.stem_conv3x3_s2(x, w0, b0, "gelu"), w1, s1, b1, None)]
        for li, layer in enumerate(self.layers):
            for blk in layer.blocks:
                steps.append(_MBConvPlan(blk, dev) if li == 0 else _BlockPlan(blk, dev))
            if layer.downsample is not None:
                steps.append(_MergePlan(layer.downsample, dev))
        return steps

    @torch.no_grad()
    def forward_nhwc(self, x):
        self._require_eval("TinyViT.forward")
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected a CUDA fp32 NCHW image batch [B,3,H,W]; the native path has no CPU fallback")
        for f in self._plan():
            x = f(x)
        return x


def _tiny_vit(defaults, kwargs):
    kw = dict(depths=[2, 2, 6, 2], window_sizes=[7, 7, 14, 7], **defaults)
    kw.update(kwargs)
    return TinyViT(**kw)


def tiny_vit_5m_224(pretrained=False, **kwargs):
    return _tiny_vit(dict(embed_dims=[64, 128, 160, 320], num_heads=[2, 4, 5, 10], drop_path_rate=0.0), kwargs)


def tiny_vit_11m_224(pretrained=False, **kwargs):
    return _tiny_vit(dict(embed_dims=[64, 128, 256, 448], num_heads=[2, 4, 8, 14], drop_path_rate=0.1), kwargs)


def tiny_vit_21m_224(pretrained=False, **kwargs):
    return _tiny_vit(dict(embed_dims=[96, 192, 384, 576], num_heads=[3, 6, 12, 18], drop_path_rate=0.2), kwargs)
