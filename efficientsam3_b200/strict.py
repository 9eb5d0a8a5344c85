"""Strict (fp32-class) precision mode of the student encoders: `with ops.strict_precision(): model(x)`.

The bf16 tensor-core path rounds every operand to bf16 (3e-3 .. 1.5e-2 rel-L2 from the reference's fp32 output); north_star's parity
tolerances (embeddings rtol 1e-4) need fp32 operands.  This module walks the same module tree as the reference does
(efficientvit/backbone.py:150-156, nn/ops.py:39-80, 273-367, 521-732; stage1/model.py:201-211) and runs every layer on the fp32
CUDA-core kernels of csrc/strict_f32.cu: activations NHWC fp32, eval-mode BatchNorm folded to per-channel (scale, bias) applied in
the SGEMM / depthwise epilogue, LiteMLA's attention in fp32 exactly as the reference forces it (ops.py:586-589).
Eval mode only; no CPU fallback (the ops raise on CPU tensors)."""
from __future__ import annotations

import torch

from . import ops
from .nn_utils import bn_scale_bias, dw_weight


def _fold(layer):
    conv = layer.conv
    return bn_scale_bias(layer.norm, conv.bias, conv.out_channels, conv.weight.device)


def conv_layer(layer, x, residual=None, nchw=False):
    """ConvLayer (ops.py:39-80): conv (same padding) -> eval BatchNorm -> activation (+ residual).  x NHWC fp32."""
    conv = layer.conv
    scale, bias = _fold(layer)
    ks, stride = conv.kernel_size[0], conv.stride[0]
    if conv.groups == 1:
        return ops.conv2d_f32(x, conv.weight.detach().float(), stride, ks // 2, scale=scale, bias=bias, act=layer.act, residual=residual,
                              nchw=nchw)
    if conv.groups != conv.in_channels or conv.in_channels != conv.out_channels:
        raise NotImplementedError("strict mode: grouped convolutions other than depthwise are handled by their owners (LiteMLA)")
    y = ops.dwconv_f32(x, dw_weight(conv, None), scale, bias, ks, stride, layer.act)
    if residual is not None:
        y = ops.bias_act_res_f32(y, None, None, residual)
    return y


def dsconv(m, x, residual=None):
    return conv_layer(m.point_conv, conv_layer(m.depth_conv, x), residual=residual)


def mbconv(m, x, residual=None):
    return conv_layer(m.point_conv, conv_layer(m.depth_conv, conv_layer(m.inverted_conv, x)), residual=residual)


def lite_mla(m, x):
    """x + LiteMLA(x) (ops.py:656-671 inside ResidualBlock :740-770), linear-attention branch."""
    B, H, W, C = x.shape
    d = m.dim
    if H * W <= d:
        raise NotImplementedError("LiteMLA quadratic branch (H*W <= dim, ops.py:623-654) is not built natively")
    qkv_conv = m.qkv.conv
    c3 = qkv_conv.out_channels
    ms = torch.empty((B, H, W, 2 * c3), device=x.device, dtype=torch.float32)
    ms2 = ms.view(-1, 2 * c3)
    s, b = _fold(m.qkv)
    ops.sgemm(x.view(-1, C), qkv_conv.weight.detach().float().reshape(c3, C), scale=s, bias=b, act=m.qkv.act, out=ms2[:, :c3])
    dwc, pwc = m.aggreg[0][0], m.aggreg[0][1]
    t = ops.dwconv_f32(ms[..., :c3], dw_weight(dwc, None), None, dwc.bias.detach().float() if dwc.bias is not None else None, 5, 1, None)
    groups = pwc.groups
    gs = c3 // groups                                                            # channels per group of the grouped 1x1
    w_bd = torch.block_diag(*pwc.weight.detach().float().reshape(groups, gs, gs)).contiguous()   # [c3, c3], zero off the groups
    ops.sgemm(t.view(-1, c3), w_bd, bias=pwc.bias.detach().float() if pwc.bias is not None else None, out=ms2[:, c3:])
    att = ops.litemla_attn_f32(ms, 2 * m.heads, d, m.eps)                        # [B,H,W,2*heads*dim]
    return conv_layer(m.proj, att, residual=x)


def efficientvit_backbone(bb, x, return_stages=False):
    """EfficientViTBackbone.forward (backbone.py:150-156) on the NCHW fp32 image -> final stage NHWC fp32 (and every stage)."""
    from .backbones.efficientvit import DSConv, EfficientViTBlock, MBConv, ResidualBlock
    stages = {}
    stem = list(bb.input_stem.op_list)
    x = conv_layer(stem[0], x.contiguous(), nchw=True)
    for blk in stem[1:]:
        x = dsconv(blk.main, x, residual=x if blk.shortcut is not None else None)
    stages["stage0"] = x
    for sid, stage in enumerate(bb.stages, 1):
        for op in stage.op_list:
            if isinstance(op, ResidualBlock):
                main = op.main
                res = x if op.shortcut is not None else None
                x = mbconv(main, x, res) if isinstance(main, MBConv) else dsconv(main, x, res)
            elif isinstance(op, EfficientViTBlock):
                x = lite_mla(op.context_module.main, x)
                x = mbconv(op.local_module.main, x, residual=x)
            else:
                raise TypeError(type(op))
        stages[f"stage{sid}"] = x
    return (x, stages) if return_stages else x


def student_head(enc, feats):
    """ImageStudentEncoder.head + resize (stage1/model.py:194-211) on NHWC fp32 features -> NCHW fp32 [B, D, E, E]."""
    h0, bn, _, h3 = enc.head
    s, b = bn_scale_bias(bn, None, h0.out_channels, h0.weight.device)
    y = ops.conv2d_f32(feats, h0.weight.detach().float(), 1, 0, scale=s, bias=b, act="gelu")
    y = ops.conv2d_f32(y, h3.weight.detach().float(), 1, 1, bias=h3.bias.detach().float())
    return ops.bilinear_nhwc_f32_to_nchw(y, enc.embed_size, enc.embed_size)


def student_forward(enc, x):
    """ImageStudentEncoder.forward in the strict mode.  EfficientViT students (b0 / b1 = EV-M / b2)."""
    from .stage1.model import EfficientViTAdapter
    if not isinstance(enc.backbone, EfficientViTAdapter):
        raise NotImplementedError("strict precision mode is built for the EfficientViT students (EV-M headline); "
                                  f"{type(enc.backbone).__name__} runs in the bf16 mode only")
    if enc.training:
        raise NotImplementedError("strict precision mode is eval-only")
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
        raise ValueError("expected a CUDA fp32 NCHW image batch [B,3,H,W]; the native path has no CPU fallback")
    return student_head(enc, efficientvit_backbone(enc.backbone.model, x))
