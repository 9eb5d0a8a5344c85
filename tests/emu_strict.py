"""TEST INFRASTRUCTURE ONLY: torch-CPU statements of the strict-mode ops (efficientsam3_b200/ops.py, csrc/strict_f32.cu), same argument
lists.  CPU tests swap them in to check the strict graphs (efficientsam3_b200/strict.py, the strict branches of the SAM heads) against
the oracle without a GPU; GPU tests compare each kernel with its function here.  The product never imports this file."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _act(x, act):
    if act in (None, "none"):
        return x
    return {"relu": F.relu, "hswish": F.hardswish, "gelu": F.gelu, "relu6": F.relu6, "sigmoid": torch.sigmoid}[act](x)


def sgemm(a, w, *, scale=None, bias=None, act=None, residual=None, out=None, act_after_res=False):
    v = a @ w.t()
    if scale is not None:
        v = v * scale
    if bias is not None:
        v = v + bias
    if not act_after_res:
        v = _act(v, act)
    if residual is not None:
        v = v + residual
    if act_after_res:
        v = _act(v, act)
    if out is None:
        return v
    out.copy_(v)
    return out


def conv2d_f32(x, weight, stride=1, pad=0, *, scale=None, bias=None, act=None, residual=None, nchw=False):
    xi = x if nchw else x.permute(0, 3, 1, 2)
    v = F.conv2d(xi, weight, None, stride=stride, padding=pad)
    if scale is not None:
        v = v * scale.view(1, -1, 1, 1)
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    v = _act(v, act).permute(0, 2, 3, 1)
    if residual is not None:
        v = v + residual
    return v.contiguous()


def dwconv_f32(x, w, scale, bias, ks, stride, act, out=None):
    C = x.shape[-1]
    wt = w.t().reshape(C, 1, ks, ks)
    v = F.conv2d(x.permute(0, 3, 1, 2), wt, None, stride=stride, padding=ks // 2, groups=C)
    if scale is not None:
        v = v * scale.view(1, -1, 1, 1)
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    v = _act(v, act).permute(0, 2, 3, 1).contiguous()
    if out is None:
        return v
    out.copy_(v)
    return out


def litemla_attn_f32(ms, heads, dim, eps):
    B, H, W, ld = ms.shape
    t = ms.reshape(B, H * W, heads, 3 * dim).permute(0, 2, 3, 1)                 # [B, heads, 3 dim, HW]
    q, k, v = F.relu(t[:, :, :dim]), F.relu(t[:, :, dim:2 * dim]), t[:, :, 2 * dim:]
    v1 = F.pad(v, (0, 0, 0, 1), value=1.0)
    o = (v1 @ k.transpose(-1, -2)) @ q
    o = o[:, :, :-1] / (o[:, :, -1:] + eps)                                      # [B, heads, dim, HW]
    return o.permute(0, 3, 1, 2).reshape(B, H, W, heads * dim).contiguous()


def bilinear_nhwc_f32_to_nchw(x, Ho, Wo):
    xi = x.permute(0, 3, 1, 2)
    if xi.shape[-2:] == (Ho, Wo):
        return xi.contiguous()
    return F.interpolate(xi, size=(Ho, Wo), mode="bilinear", align_corners=False).contiguous()


def attn_few_keys_f32(q, k, v, B, heads, scale):
    D = q.shape[1]
    Nq, Tk = q.shape[0] // B, k.shape[1]
    qh = q.view(B, Nq, heads, D // heads).transpose(1, 2)
    kh, vh = (t.view(B, Tk, heads, D // heads).transpose(1, 2) for t in (k, v))
    a = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (a @ vh).transpose(1, 2).reshape(B * Nq, D).contiguous()


def ln_rows_gelu_f32(x, w, b, eps):
    return F.gelu(F.layer_norm(x, (x.shape[-1],), w, b, eps))


def bias_act_res_f32(x, bias=None, act=None, residual=None, act_after_res=False):
    v = x if bias is None else x + bias
    if not act_after_res:
        v = _act(v, act)
    if residual is not None:
        v = v + residual
    if act_after_res:
        v = _act(v, act)
    return v


def convt2x2_f32(x, weight, bias=None, act=None, residual=None, act_after_res=False):
    v = F.conv_transpose2d(x.permute(0, 3, 1, 2), weight, None, stride=2).permute(0, 2, 3, 1).contiguous()
    return bias_act_res_f32(v, bias, act, residual, act_after_res)


PATCHED = ["sgemm", "conv2d_f32", "dwconv_f32", "litemla_attn_f32", "bilinear_nhwc_f32_to_nchw", "attn_few_keys_f32", "ln_rows_gelu_f32",
           "bias_act_res_f32", "convt2x2_f32"]


def install(monkeypatch):
    import sys
    from efficientsam3_b200 import ops
    me = sys.modules[__name__]
    for name in PATCHED:
        monkeypatch.setattr(ops, name, getattr(me, name))
