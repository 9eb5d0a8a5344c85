"""Sam3DualViTDetNeck (SimpleFPN a la ViTDet, dual SAM3 / SAM2 branches), B200-native.  Module tree / state_dict keys
of sam3/sam3/model/necks.py:13-125 (`convs.{i}.{dconv_2x2_0,dconv_2x2_1,dconv_2x2,conv_1x1,conv_3x3}`, `sam2_convs.*`,
`trunk.*`).  ConvTranspose2d(2,2) -> es3_convt2x2_bf16 (tcgen05 GEMM + depth-to-space epilogue, GELU fused),
1x1 -> es3_gemm_bf16, 3x3 -> es3_conv3x3_bf16 (implicit GEMM), MaxPool -> es3_maxpool2x2_bf16."""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn

from .. import ops
from ..nn_utils import NativePlanMixin, conv3x3_weight, pw_weight


class Sam3DualViTDetNeck(nn.Module, NativePlanMixin):
    def __init__(self, trunk, position_encoding, d_model, scale_factors=(4.0, 2.0, 1.0, 0.5), add_sam2_neck=False):
        super().__init__()
        self.trunk = trunk
        self.position_encoding = position_encoding     # sine PE is not consumed on the point-prompt path
        self.convs = nn.ModuleList()
        self.scale_factors = list(scale_factors)
        dim = self.trunk.channel_list[-1]
        for scale in self.scale_factors:
            cur = nn.Sequential()
            if scale == 4.0:
                cur.add_module("dconv_2x2_0", nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2))
                cur.add_module("gelu", nn.GELU())
                cur.add_module("dconv_2x2_1", nn.ConvTranspose2d(dim // 2, dim // 4, kernel_size=2, stride=2))
                out_dim = dim // 4
            elif scale == 2.0:
                cur.add_module("dconv_2x2", nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2))
                out_dim = dim // 2
            elif scale == 1.0:
                out_dim = dim
            elif scale == 0.5:
                cur.add_module("maxpool_2x2", nn.MaxPool2d(kernel_size=2, stride=2))
                out_dim = dim
            else:
                raise NotImplementedError(f"scale_factor={scale} is not supported yet.")
            cur.add_module("conv_1x1", nn.Conv2d(out_dim, d_model, kernel_size=1, bias=True))
            cur.add_module("conv_3x3", nn.Conv2d(d_model, d_model, kernel_size=3, padding=1, bias=True))
            self.convs.append(cur)
        self.sam2_convs = deepcopy(self.convs) if add_sam2_neck else None

    def _build_plan(self):
        def ct(m):
            return ops.convt2x2_weight(m.weight), m.bias.detach().float().repeat(4).contiguous()

        def branch(seq, scale):
            d = dict(scale=scale, w1=pw_weight(seq.conv_1x1), b1=seq.conv_1x1.bias.detach().float().contiguous(),
                     w3=conv3x3_weight(seq.conv_3x3), b3=seq.conv_3x3.bias.detach().float().contiguous())
            if scale == 4.0:
                d["ct0"], d["ct1"] = ct(seq.dconv_2x2_0), ct(seq.dconv_2x2_1)
            elif scale == 2.0:
                d["ct"] = ct(seq.dconv_2x2)
            return d

        plan = {"sam3": [branch(s, sc) for s, sc in zip(self.convs, self.scale_factors)]}
        if self.sam2_convs is not None:
            plan["sam2"] = [branch(s, sc) for s, sc in zip(self.sam2_convs, self.scale_factors)]
        return plan

    @staticmethod
    def _run_branch_strict(seq, scale, x):  # x: [B,h,w,C] fp32 NHWC; strict precision mode (fp32 kernels, raw parameters)
        f = lambda t: t.detach().float().contiguous()
        if scale == 4.0:
            x = ops.convt2x2_f32(x, f(seq.dconv_2x2_0.weight), f(seq.dconv_2x2_0.bias), act="gelu")
            x = ops.convt2x2_f32(x, f(seq.dconv_2x2_1.weight), f(seq.dconv_2x2_1.bias))
        elif scale == 2.0:
            x = ops.convt2x2_f32(x, f(seq.dconv_2x2.weight), f(seq.dconv_2x2.bias))
        elif scale == 0.5:
            raise NotImplementedError("strict precision mode: the 0.5x FPN level (MaxPool) is dropped by scalp=1 and not built")
        y = ops.conv2d_f32(x, f(seq.conv_1x1.weight), 1, 0, bias=f(seq.conv_1x1.bias))
        return ops.conv2d_f32(y, f(seq.conv_3x3.weight), 1, 1, bias=f(seq.conv_3x3.bias))

    @staticmethod
    def _run_branch(bp, x, out_dtype=torch.bfloat16):  # x: [B,h,w,C] bf16 NHWC
        s = bp["scale"]
        if s == 4.0:
            x = ops.convt2x2(x, bp["ct0"][0], bias4=bp["ct0"][1], act="gelu")
            x = ops.convt2x2(x, bp["ct1"][0], bias4=bp["ct1"][1])
        elif s == 2.0:
            x = ops.convt2x2(x, bp["ct"][0], bias4=bp["ct"][1])
        elif s == 0.5:
            x = ops.maxpool2x2(x)
        B, H, W, C = x.shape
        y = ops.gemm(x.view(-1, C), bp["w1"], bias=bp["b1"]).view(B, H, W, -1)
        return ops.conv3x3(y, bp["w3"], bias=bp["b3"], out_dtype=out_dtype)

    @torch.no_grad()
    def forward_nhwc(self, feats_nhwc, branch="sam2", levels=(0, 1, 2), f32_levels=()):
        """Fast path: trunk feature map [B,h,w,C] bf16 NHWC -> list of NHWC levels of one branch (bf16; fp32 for
        the levels listed in f32_levels)."""
        self._require_eval("Sam3DualViTDetNeck.forward")
        if ops.precision() == "strict":
            seqs = self.convs if branch == "sam3" else self.sam2_convs
            return [self._run_branch_strict(seqs[i], self.scale_factors[i], feats_nhwc.float()) for i in levels]
        plan = self._plan()[branch]
        return [self._run_branch(plan[i], feats_nhwc, torch.float32 if i in f32_levels else torch.bfloat16) for i in levels]

    @torch.no_grad()
    def forward(self, tensor_list):
        """Reference contract (necks.py:100-125): images -> (sam3_out, sam3_pos, sam2_out, sam2_pos), NCHW fp32 maps.
        Positional encodings are returned as None (they feed the detector / tracker memory, not this path)."""
        xs = self.trunk(tensor_list)
        n = len(self.scale_factors)
        if ops.precision() == "strict":      # fp32 NHWC stream; the 0.5x level is not built in this mode (scalp=1 drops it)
            x = xs[-1].float().permute(0, 2, 3, 1).contiguous()
            lv = [i for i, sc in enumerate(self.scale_factors) if sc != 0.5]
            nchw = lambda ts: [t.permute(0, 3, 1, 2).contiguous() for t in ts]
            s3 = nchw(self.forward_nhwc(x, "sam3", lv))
            s2 = nchw(self.forward_nhwc(x, "sam2", lv)) if self.sam2_convs is not None else None
            return s3, [None] * len(lv), s2, ([None] * len(lv) if s2 is not None else None)
        x = ops.nchw_f32_to_nhwc(xs[-1])
        s3 = [ops.nhwc_to_nchw_f32(t) for t in self.forward_nhwc(x, "sam3", range(n))]
        s2 = [ops.nhwc_to_nchw_f32(t) for t in self.forward_nhwc(x, "sam2", range(n))] if self.sam2_convs is not None else None
        return s3, [None] * n, s2, ([None] * n if s2 is not None else None)
