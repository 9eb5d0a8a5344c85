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


# ------------------------------------------------------------------------------------------------ RepViT
def _cbn(cb, x, act=None, residual=None, nchw=False):
    """Conv2d_BN (repvit.py:36-60, tiny_vit.py:28-58) in eval mode (+ activation, + residual): dense or depthwise."""
    conv = cb.c
    scale, bias = bn_scale_bias(cb.bn, None, conv.out_channels, conv.weight.device)
    ks, stride = conv.kernel_size[0], conv.stride[0]
    if conv.groups == 1:
        return ops.conv2d_f32(x, conv.weight.detach().float(), stride, conv.padding[0], scale=scale, bias=bias, act=act, residual=residual,
                              nchw=nchw)
    assert conv.groups == conv.in_channels == conv.out_channels and residual is None and conv.padding[0] == ks // 2
    return ops.dwconv_f32(x, dw_weight(conv, None), scale, bias, ks, stride, act)


def squeeze_excite(se, x):
    """timm SqueezeExcite (repvit.py:23, 136): x * sigmoid(fc2(relu(fc1(mean_hw x)))), x NHWC fp32."""
    B, H, W, C = x.shape
    m = torch.zeros((B, C), device=x.device, dtype=torch.float32)
    for b in range(B):
        ops.colsum_f32(x[b].view(H * W, C), m[b])
    m = ops.bias_act_res_f32(m * (1.0 / (H * W)))
    h = ops.sgemm(m, se.fc1.weight.detach().float().view(se.fc1.out_channels, C), bias=se.fc1.bias.detach().float(), act="relu")
    g = ops.sgemm(h, se.fc2.weight.detach().float().view(C, -1), bias=se.fc2.bias.detach().float(), act="sigmoid")
    return ops.scale_channels_f32(x, g)


def repvgg_dw(rv, x):
    """RepVGGDW (repvit.py:84-122) in eval mode: bn(conv_bn(x) + conv1(x) + x) as the single depthwise 3x3 the reference's own
    fuse() produces, folded in fp32."""
    dev = x.device
    s1, b1 = bn_scale_bias(rv.conv.bn, None, rv.dim, dev)
    w = rv.conv.c.weight.detach().float() * s1.view(-1, 1, 1, 1)
    w[:, :, 1, 1] += rv.conv1.weight.detach().float()[:, :, 0, 0] + 1.0
    bsum = b1 + rv.conv1.bias.detach().float()
    s2, b2 = bn_scale_bias(rv.bn, None, rv.dim, dev)
    w = (w * s2.view(-1, 1, 1, 1)).reshape(rv.dim, 9).t().contiguous()
    return ops.dwconv_f32(x, w, None, (b2 + bsum * s2).contiguous(), 3, 1, None)


def repvit_block(blk, x):
    """RepViTBlock (repvit.py:139-197): token mixer, then x + channel_mixer(x)."""
    tm = blk.token_mixer
    if blk.stride == 2:
        x = _cbn(tm[0], x)
        if blk.use_se:
            x = squeeze_excite(tm[1], x)
        x = _cbn(tm[2], x)
    else:
        x = repvgg_dw(tm[0], x)
        if blk.use_se:
            x = squeeze_excite(tm[1], x)
    cm = blk.channel_mixer.m
    return _cbn(cm[2], _cbn(cm[0], x, act="gelu"), residual=x)


def repvit_backbone(rv, x):
    """RepViT.features (repvit.py:236-262) on the NCHW fp32 image -> NHWC fp32 features."""
    pe = rv.features[0]
    x = _cbn(pe[2], _cbn(pe[0], x, act="gelu", nchw=True))
    for blk in list(rv.features)[1:]:
        x = repvit_block(blk, x)
    return x


# ------------------------------------------------------------------------------------------------ TinyViT
def _ln(x2, norm):
    return ops.ln_rows_f32(x2, norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous(), norm.eps)


def _linear(l, x2, act=None, residual=None):
    return ops.sgemm(x2, l.weight.detach().float(), bias=l.bias.detach().float(), act=act, residual=residual)


def tinyvit_mbconv(m, x):
    """MBConv (tiny_vit.py:97-133): act3(conv3(act2(conv2(act1(conv1 x)))) + x)."""
    y = _cbn(m.conv2, _cbn(m.conv1, x, act="gelu"), act="gelu")
    conv = m.conv3.c
    s, b = bn_scale_bias(m.conv3.bn, None, conv.out_channels, conv.weight.device)
    B, H, W, C = y.shape
    out = ops.sgemm(y.view(-1, C), conv.weight.detach().float().view(conv.out_channels, C), scale=s, bias=b, act="gelu",
                    residual=x.view(-1, x.shape[-1]), act_after_res=True)
    return out.view(B, H, W, -1)


def tinyvit_merge(m, x):
    """PatchMerging (tiny_vit.py:136-169)."""
    return _cbn(m.conv3, _cbn(m.conv2, _cbn(m.conv1, x, act="gelu"), act="gelu"))


def tinyvit_block(blk, x):
    """TinyViTBlock (tiny_vit.py:290-388): x + attn(window(x)), local depthwise conv, x + mlp(x); windows zero-padded BEFORE the
    attention's LayerNorm, so a padded token's qkv row is W_qkv beta + b."""
    at, mlp = blk.attn, blk.mlp
    B, H, W, C = x.shape
    x2 = x.view(-1, C)
    qkv = _linear(at.qkv, _ln(x2, at.norm))
    pad = _linear(at.qkv, at.norm.bias.detach().float().view(1, -1).contiguous()).view(-1)
    bias = at.attention_biases.detach().float()[:, at.attention_bias_idxs].contiguous()
    a = ops.attention_f32(qkv, B, H, W, at.num_heads, at.key_dim, blk.window_size, at.scale, layout="per_head", bias=bias, pad_row=pad)
    x2 = _linear(at.proj, a, residual=x2)
    x2 = _cbn(blk.local_conv, x2.view(B, H, W, C)).view(-1, C)
    h = _linear(mlp.fc1, _ln(x2, mlp.norm), act="gelu")
    return _linear(mlp.fc2, h, residual=x2).view(B, H, W, C)


def tinyvit_backbone(tv, x):
    """TinyViT.forward_features up to the last layer (tiny_vit.py:560-574; stage1/model.py:299-324 drops the head)."""
    seq = tv.patch_embed.seq
    x = _cbn(seq[2], _cbn(seq[0], x, act="gelu", nchw=True))
    for li, layer in enumerate(tv.layers):
        for blk in layer.blocks:
            x = tinyvit_mbconv(blk, x) if li == 0 else tinyvit_block(blk, x)
        if layer.downsample is not None:
            x = tinyvit_merge(layer.downsample, x)
    return x


# ------------------------------------------------------------------------------------------------ SAM3 ViT trunk (the teacher)
def vit_tokens(vit, x):
    """ViT.forward (vitdet.py:798-879 for the SAM3 configuration) -> fp32 tokens [B*h*w, C] after the last block, (B, h, w)."""
    B, _, S, _ = x.shape
    P, C, heads = vit.patch_size, vit.embed_dim, vit.num_heads
    h = w = S // P
    tok = ops.conv2d_f32(x, vit.patch_embed.proj.weight.detach().float(), P, 0, nchw=True).view(-1, C)
    tab = vit.pos_embed.detach()[0, 1:] if vit.pretrain_use_cls_token else vit.pos_embed.detach()[0]
    import math
    xs = ops.layernorm(tok, vit.ln_pre.weight.detach().float().contiguous(), vit.ln_pre.bias.detach().float().contiguous(), vit.ln_pre.eps,
                       pos=tab.float().contiguous(), pos_size=int(math.isqrt(tab.shape[0])), H=h, W=w, out_bf16=False, out_f32=True)[1]
    for blk in vit.blocks:
        at = blk.attn
        qkv = _linear(at.qkv, _ln(xs, blk.norm1))
        table = torch.view_as_real(at.freqs_cis.detach().to(torch.complex64)).float().contiguous()
        ops.rope_f32(qkv, table, 2 * C, h, w, blk.window_size)
        a = ops.attention_f32(qkv, B, h, w, heads, C // heads, blk.window_size, at.scale)
        xs = _linear(at.proj, a, residual=xs)
        hdn = _linear(blk.mlp.fc1, _ln(xs, blk.norm2), act="gelu")
        xs = _linear(blk.mlp.fc2, hdn, residual=xs)
    return xs, (B, h, w)


def student_forward(enc, x):
    """ImageStudentEncoder.forward in the strict mode, all nine students (EfficientViT b0/b1/b2, RepViT m0.9/m1.1/m2.3, TinyViT 5m/11m/21m)."""
    from .stage1.model import EfficientViTAdapter, RepViTAdapter, TinyViTAdapter
    if enc.training:
        raise NotImplementedError("strict precision mode is eval-only")
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
        raise ValueError("expected a CUDA fp32 NCHW image batch [B,3,H,W]; the native path has no CPU fallback")
    bb = enc.backbone
    if isinstance(bb, EfficientViTAdapter):
        feats = efficientvit_backbone(bb.model, x)
    elif isinstance(bb, RepViTAdapter):
        feats = repvit_backbone(bb.model, x)
    elif isinstance(bb, TinyViTAdapter):
        feats = tinyvit_backbone(bb.model, x)
    else:
        raise NotImplementedError(f"strict precision mode: unknown backbone adapter {type(bb).__name__}")
    return student_head(enc, feats)
