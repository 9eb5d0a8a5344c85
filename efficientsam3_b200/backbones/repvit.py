"""RepViT backbone (stage-1 student "RV-M" = repvit_m1_1), B200-native.  Module tree / state_dict keys of
sam3/sam3/backbones/repvit.py (Conv2d_BN :27-49, Residual :51-81, RepVGGDW :84-122, RepViTBlock :125-161,
RepViT :219-246, configs :253-472); SqueezeExcite parameters follow timm.layers.SqueezeExcite (fc1 / fc2 1x1 convs
with bias).  Eval-mode execution:
  patch embed   es3_stem_conv3x3_s2 (3->C/2, BN, GELU) + es3_conv3x3_s2_narrow_bf16 (mma.sync implicit GEMM)
  RepVGGDW      re-parameterised on the host exactly as the reference's own fuse() (repvit.py:97-122) into one
                depthwise 3x3 + bias -> es3_dwconv_tiled_bf16
  SqueezeExcite es3_channel_mean -> es3_gemm_simt x2 (ReLU, sigmoid) -> es3_scale_channels
  1x1 convs     es3_gemm_bf16 (tcgen05) with folded BN, GELU and the residual add in the epilogue
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..nn_utils import NativePlanMixin, bn_scale_bias, dw_weight, pack_patch_embed, pw_weight, pw_weight_scaled


def _make_divisible(v, divisor, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class Conv2d_BN(nn.Sequential):
    def __init__(self, a, b, ks=1, stride=1, pad=0, dilation=1, groups=1, bn_weight_init=1, resolution=-10000):
        super().__init__()
        self.add_module("c", nn.Conv2d(a, b, ks, stride, pad, dilation, groups, bias=False))
        self.add_module("bn", nn.BatchNorm2d(b))
        nn.init.constant_(self.bn.weight, bn_weight_init)
        nn.init.constant_(self.bn.bias, 0)


class Residual(nn.Module):
    def __init__(self, m, drop=0.0):
        super().__init__()
        self.m = m
        self.drop = drop


class RepVGGDW(nn.Module):
    def __init__(self, ed):
        super().__init__()
        self.conv = Conv2d_BN(ed, ed, 3, 1, 1, groups=ed)
        self.conv1 = nn.Conv2d(ed, ed, 1, 1, 0, groups=ed)
        self.dim = ed
        self.bn = nn.BatchNorm2d(ed)


class SqueezeExcite(nn.Module):
    """Parameter layout of timm.layers.SqueezeExcite(channels, rd_ratio=0.25): rd = make_divisible(C*0.25, 8,
    round_limit=0.)."""

    def __init__(self, channels, rd_ratio=0.25):
        super().__init__()
        rd = max(8, int(channels * rd_ratio + 4) // 8 * 8)
        self.fc1 = nn.Conv2d(channels, rd, kernel_size=1, bias=True)
        self.fc2 = nn.Conv2d(rd, channels, kernel_size=1, bias=True)


class RepViTBlock(nn.Module):
    def __init__(self, inp, hidden_dim, oup, kernel_size, stride, use_se, use_hs):
        super().__init__()
        assert stride in [1, 2]
        self.identity = stride == 1 and inp == oup
        assert hidden_dim == 2 * inp
        self.stride, self.use_se = stride, bool(use_se)
        if stride == 2:
            self.token_mixer = nn.Sequential(
                Conv2d_BN(inp, inp, kernel_size, stride, (kernel_size - 1) // 2, groups=inp),
                SqueezeExcite(inp, 0.25) if use_se else nn.Identity(),
                Conv2d_BN(inp, oup, ks=1, stride=1, pad=0))
            self.channel_mixer = Residual(nn.Sequential(Conv2d_BN(oup, 2 * oup, 1, 1, 0), nn.GELU(),
                                                        Conv2d_BN(2 * oup, oup, 1, 1, 0, bn_weight_init=0)))
        else:
            assert self.identity
            self.token_mixer = nn.Sequential(RepVGGDW(inp), SqueezeExcite(inp, 0.25) if use_se else nn.Identity())
            self.channel_mixer = Residual(nn.Sequential(Conv2d_BN(inp, hidden_dim, 1, 1, 0), nn.GELU(),
                                                        Conv2d_BN(hidden_dim, oup, 1, 1, 0, bn_weight_init=0)))


class Classfier(nn.Module):
    def __init__(self, dim, num_classes, distillation=True):
        super().__init__()
        assert num_classes == 0, "the stage-1 student builds RepViT with num_classes=0 (stage1/model.py:393)"
        self.classifier = nn.Identity()
        self.distillation = distillation


def _fold_cb(cb: Conv2d_BN, dev):
    return bn_scale_bias(cb.bn, None, cb.c.out_channels, dev)


class _SEPlan:
    def __init__(self, se: SqueezeExcite):
        self.w1 = se.fc1.weight.detach().float().reshape(se.fc1.out_channels, -1).contiguous()
        self.b1 = se.fc1.bias.detach().float().contiguous()
        self.w2 = se.fc2.weight.detach().float().reshape(se.fc2.out_channels, -1).contiguous()
        self.b2 = se.fc2.bias.detach().float().contiguous()

    def __call__(self, x):
        m = ops.channel_mean(x)
        h = ops.gemm_simt(m, self.w1, bias=self.b1, act="relu", out_dtype=torch.float32)
        g = ops.gemm_simt(h, self.w2, bias=self.b2, act="sigmoid", out_dtype=torch.float32)
        return ops.scale_channels(x, g)


class _PW:
    def __init__(self, cb: Conv2d_BN, act, dev):
        self.s, self.b = _fold_cb(cb, dev)
        self.w = pw_weight_scaled(cb.c, self.s)               # BN scale folded before the bf16 rounding: bias-only GEMM epilogue
        self.act = act

    def __call__(self, x, residual=None):
        B, H, W, C = x.shape
        r = residual.view(-1, residual.shape[-1]) if residual is not None else None
        return ops.gemm(x.view(-1, C), self.w, bias=self.b, act=self.act, residual=r).view(B, H, W, -1)


class _BlockPlan:
    def __init__(self, blk: RepViTBlock, dev):
        self.stride = blk.stride
        tm = blk.token_mixer
        if blk.stride == 2:
            s, b = _fold_cb(tm[0], dev)
            self.dw_w, self.dw_b = dw_weight(tm[0].c, s), b
            self.se = _SEPlan(tm[1]) if blk.use_se else None
            self.pw = _PW(tm[2], None, dev)
        else:
            rv = tm[0]
            # the reference's own re-parameterisation (RepVGGDW.fuse, repvit.py:97-122), in fp32
            s1, b1 = _fold_cb(rv.conv, dev)
            w = rv.conv.c.weight.detach().float() * s1.view(-1, 1, 1, 1)
            w[:, :, 1, 1] += rv.conv1.weight.detach().float()[:, :, 0, 0] + 1.0
            bsum = b1 + rv.conv1.bias.detach().float()
            s2 = rv.bn.weight.detach().float() / torch.sqrt(rv.bn.running_var.detach().float() + rv.bn.eps)
            w = w * s2.view(-1, 1, 1, 1)
            self.dw_b = (rv.bn.bias.detach().float() + (bsum - rv.bn.running_mean.detach().float()) * s2).contiguous()
            self.dw_w = w.reshape(w.shape[0], 9).t().contiguous()
            self.se = _SEPlan(tm[1]) if blk.use_se else None
            self.pw = None
        cm = blk.channel_mixer.m
        self.m0, self.m2 = _PW(cm[0], "gelu", dev), _PW(cm[2], None, dev)

    def __call__(self, x):
        x = ops.dwconv(x, self.dw_w, self.dw_b, 3, self.stride, None)
        if self.se is not None:
            x = self.se(x)
        if self.pw is not None:
            x = self.pw(x)
        return self.m2(self.m0(x), residual=x)


class RepViT(nn.Module, NativePlanMixin):
    def __init__(self, cfgs, num_classes=1000, distillation=False):
        super().__init__()
        self.cfgs = cfgs
        input_channel = self.cfgs[0][2]
        patch_embed = nn.Sequential(Conv2d_BN(3, input_channel // 2, 3, 2, 1), nn.GELU(),
                                    Conv2d_BN(input_channel // 2, input_channel, 3, 2, 1))
        layers = [patch_embed]
        for k, t, c, use_se, use_hs, s in self.cfgs:
            output_channel = _make_divisible(c, 8)
            exp_size = _make_divisible(input_channel * t, 8)
            layers.append(RepViTBlock(input_channel, exp_size, output_channel, k, s, use_se, use_hs))
            input_channel = output_channel
        self.features = nn.ModuleList(layers)
        self.classifier = Classfier(output_channel, num_classes, distillation)

    def _build_plan(self):
        dev = next(self.parameters()).device
        pe = self.features[0]
        c0, c1 = pe[0], pe[2]
        s0, b0 = _fold_cb(c0, dev)
        s1, b1 = _fold_cb(c1, dev)
        w0, b0, w1 = pack_patch_embed(c0.c.weight, s0, b0, c1.c.weight)
        steps = [lambda x: ops.conv3x3_s2_narrow(ops.stem_conv3x3_s2(x, w0, b0, "gelu"), w1, s1, b1, None)]
        steps += [_BlockPlan(b, dev) for b in list(self.features)[1:]]
        return steps

    @torch.no_grad()
    def forward_nhwc(self, x):
        self._require_eval("RepViT.forward")
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected a CUDA fp32 NCHW image batch [B,3,H,W]; the native path has no CPU fallback")
        for f in self._plan():
            x = f(x)
        return x


def _rv(cfgs, num_classes=1000, distillation=False):
    return RepViT(cfgs, num_classes=num_classes, distillation=distillation)


def _stage_cfgs(widths, depths):
    """RepViT block table rows [k, t, c, SE, HS, s] (repvit.py:280-520): four stages; a stage after the first opens with a
    stride-2 block; SqueezeExcite alternates (first stage starts with SE, later ones with the plain downsample block) and
    the last block of a stage never has it; hard-swish flag (unused by the blocks) is set in the last two stages."""
    rows = []
    for si, (c, n) in enumerate(zip(widths, depths)):
        for i in range(n):
            se = (i % 2 == 0) if si == 0 else (i % 2 == 1)
            if i == n - 1:
                se = False
            rows.append([3, 2, c, int(se), int(si >= 2), 2 if (si > 0 and i == 0) else 1])
    return rows


def repvit_m0_9(pretrained=False, num_classes=1000, distillation=False):
    return _rv(_stage_cfgs((48, 96, 192, 384), (3, 4, 16, 3)), num_classes, distillation)


def repvit_m1_1(pretrained=False, num_classes=1000, distillation=False):
    return _rv(_stage_cfgs((64, 128, 256, 512), (3, 4, 14, 3)), num_classes, distillation)


def repvit_m2_3(pretrained=False, num_classes=1000, distillation=False):
    return _rv(_stage_cfgs((80, 160, 320, 640), (7, 8, 36, 3)), num_classes, distillation)
