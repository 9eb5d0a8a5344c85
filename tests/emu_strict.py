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


def layernorm(x, gamma, beta, eps=1e-5, *, pos=None, pos_size=0, H=0, W=0, out_bf16=True, out_f32=False):
    """The fp32-output form of ops.layernorm (the strict graphs never ask for bf16): optional tiled abs-pos add, then LayerNorm."""
    assert out_f32 and not out_bf16
    v = x
    if pos is not None:
        C = x.shape[1]
        hh, ww = torch.arange(H).view(H, 1).expand(H, W) % pos_size, torch.arange(W).view(1, W).expand(H, W) % pos_size
        v = (x.view(-1, H * W, C) + pos[(hh * pos_size + ww).reshape(-1)].to(x.dtype)).view(-1, C)
    return None, F.layer_norm(v, (v.shape[1],), gamma.to(v.dtype), beta.to(v.dtype), eps)


def ln_rows_f32(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def rope_f32(qkv, table, rope_cols, H, W, win):
    M = qkv.shape[0]
    hh, ww = torch.arange(H).view(H, 1).expand(H, W), torch.arange(W).view(1, W).expand(H, W)
    idx = ((hh % win) * win + (ww % win)) if win else (hh * W + ww)
    c = torch.view_as_complex(table.to(qkv.dtype).contiguous())[idx.reshape(-1)]            # [HW, 32]
    t = qkv[:, :rope_cols].reshape(M // (H * W), H * W, rope_cols // 64, 32, 2).contiguous()
    r = torch.view_as_real(torch.view_as_complex(t) * c.view(1, H * W, 1, 32))
    qkv[:, :rope_cols] = r.reshape(M, rope_cols)
    return qkv


def attention_f32(qkv, B, H, W, heads, head_dim, win, scale, *, layout="blocks", bias=None, pad_row=None):
    C = heads * head_dim
    t = qkv.view(B, H, W, 3 * C)
    ws = win if win else max(H, W)
    Hp, Wp = (-(-H // win) * win, -(-W // win) * win) if win else (H, W)
    if (Hp, Wp) != (H, W):
        full = pad_row.to(qkv.dtype).view(1, 1, 1, -1).expand(B, Hp, Wp, 3 * C).clone()
        full[:, :H, :W] = t
        t = full
    if win:
        nh, nw = Hp // win, Wp // win
        t = t.view(B, nh, win, nw, win, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nw, win * win, 3 * C)
    else:
        t = t.reshape(B, H * W, 3 * C)
    if layout == "blocks":
        q, k, v = (t[..., i * C:(i + 1) * C].reshape(t.shape[0], t.shape[1], heads, head_dim).transpose(1, 2) for i in range(3))
    else:
        u = t.reshape(t.shape[0], t.shape[1], heads, 3 * head_dim).transpose(1, 2)
        q, k, v = u[..., :head_dim], u[..., head_dim:2 * head_dim], u[..., 2 * head_dim:]
    a = q @ k.transpose(-1, -2) * scale
    if bias is not None:
        a = a + bias.to(a.dtype)
    o = (torch.softmax(a, dim=-1) @ v).transpose(1, 2).reshape(t.shape[0], t.shape[1], C)
    if win:
        o = o.view(B, nh, nw, win, win, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)[:, :H, :W]
    return o.reshape(B * H * W, C).contiguous()


def scale_channels_f32(x, gate):
    return x * gate.view(gate.shape[0], 1, 1, -1)


def colsum_f32(src, out):
    out += src.sum(0)
    return out


PATCHED = ["sgemm", "conv2d_f32", "dwconv_f32", "litemla_attn_f32", "bilinear_nhwc_f32_to_nchw", "attn_few_keys_f32", "ln_rows_gelu_f32",
           "bias_act_res_f32", "convt2x2_f32", "layernorm", "ln_rows_f32", "rope_f32", "attention_f32", "scale_channels_f32", "colsum_f32"]


def install(monkeypatch):
    import sys
    from efficientsam3_b200 import ops
    me = sys.modules[__name__]
    for name in PATCHED:
        monkeypatch.setattr(ops, name, getattr(me, name))
