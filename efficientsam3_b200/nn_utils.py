"""Host-side helpers shared by the module shells: BN folding, weight packing, plan caching.

The shells (`backbones/*.py`, `stage1/model.py`, ...) keep the reference's class names, constructor
signatures and state_dict keys; torch.nn.Conv2d / BatchNorm2d / Linear objects inside them are used
ONLY as parameter containers (their forward is never called) -- all device work goes through
`efficientsam3_b200.ops` -> libes3.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def bn_scale_bias(norm: nn.BatchNorm2d | None, conv_bias: torch.Tensor | None, cout: int, device):
    """Eval-mode BatchNorm folded to per-channel (scale, bias) in fp32; conv bias merged.
    Returns (scale|None, bias|None)."""
    if norm is None:
        return None, (conv_bias.detach().float().contiguous() if conv_bias is not None else None)
    s = (norm.weight.detach().float() / torch.sqrt(norm.running_var.detach().float() + norm.eps))
    b = norm.bias.detach().float() - norm.running_mean.detach().float() * s
    if conv_bias is not None:
        b = b + conv_bias.detach().float() * s
    return s.contiguous(), b.contiguous()


# Packed / re-laid-out copies of a parameter are cached on the module that owns it and rebuilt when the parameter changes:
# (data_ptr, _version) catches torch-side writes, WEIGHTS_EPOCH the fused AdamW kernel, which moves the parameters through raw
# pointers (stage1.optim.FlatAdamW.step bumps it).  The training graphs re-derive ~150 such tensors per iteration (bf16 casts,
# transposes for the input-gradient GEMMs, rotated depthwise taps): with the cache they are built once per optimiser step.
WEIGHTS_EPOCH = 0


def bump_weights_epoch():
    global WEIGHTS_EPOCH
    WEIGHTS_EPOCH += 1


def cached_pack(owner: nn.Module, name: str, src: torch.Tensor, fn):
    key = (src.data_ptr(), src._version, src.device, WEIGHTS_EPOCH)
    cache = owner.__dict__.setdefault("_es3_pack_cache", {})
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        with torch.no_grad():
            ent = (key, fn())
        cache[name] = ent
    return ent[1]


def _pw_weight(conv):
    w = conv.weight.detach()
    return w.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()


def pw_weight(conv: nn.Conv2d) -> torch.Tensor:
    """1x1 conv weight [N,C,1,1] -> bf16 [N,C] (K-major GEMM B operand)."""
    return cached_pack(conv, "pw", conv.weight, lambda: _pw_weight(conv))


def pw_weight_scaled(conv: nn.Conv2d, scale: torch.Tensor | None) -> torch.Tensor:
    """1x1 conv weight with the folded BatchNorm scale multiplied in BEFORE the bf16 rounding: bf16 [N,C].  The GEMM epilogue of an
    eval-mode conv + BN is then bias (+ act) only: the per-channel scale cost 8 broadcast LDG.128 + 32 FMUL per 32-column chunk of
    every row (profiles/r2ae_gemm_bisect_gelu.txt: 60 -> 78 us for scale + bias at M = 131072, N = 512, K = 256)."""
    w = conv.weight.detach().float().reshape(conv.out_channels, -1)
    if scale is not None:
        w = w * scale.view(-1, 1)
    return w.to(torch.bfloat16).contiguous()


def pw_weight_t(conv: nn.Conv2d) -> torch.Tensor:
    """bf16 [C,N]: the transposed 1x1 weight, B operand of the input-gradient GEMM dx = dz . W."""
    return cached_pack(conv, "pw_t", conv.weight, lambda: _pw_weight(conv).t().contiguous())


def _dw_weight(conv, scale):
    w = conv.weight.detach().float()
    c, _, k, _ = w.shape
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    return w.reshape(c, k * k).t().contiguous()


def dw_weight(conv: nn.Conv2d, scale: torch.Tensor | None) -> torch.Tensor:
    """depthwise weight [C,1,k,k] (x folded BN scale) -> fp32 [k*k, C] tap-major."""
    if scale is not None:
        return _dw_weight(conv, scale)
    return cached_pack(conv, "dw", conv.weight, lambda: _dw_weight(conv, None))


def dw_weight_rot(conv: nn.Conv2d) -> torch.Tensor:
    """fp32 [k*k, C] with the taps rotated by 180 degrees: the stride-1 input gradient is the forward kernel on these."""
    return cached_pack(conv, "dw_rot", conv.weight, lambda: _dw_weight(conv, None).flip(0).contiguous())


def pack_patch_embed(w0: torch.Tensor, s0, b0, w1: torch.Tensor):
    """Two-conv patch embed (3 -> C/2 -> C, both 3x3 stride 2; repvit.py:219-223, tiny_vit.py:67-84) packed for
    es3_stem_conv3x3_s2 + es3_conv3x3_s2_narrow_bf16.  The intermediate width C/2 (24/32/40/48) is zero-padded to the
    kernel widths 32 / 48: padded channels carry weight 0 and bias 0, so they stay act(0) = 0 and meet zero weights again.
    Returns (w27 fp32 [27, Cp] with BN scale folded, bias [Cp], w9 bf16 [9, Cout, Cp])."""
    cmid, cout = w0.shape[0], w1.shape[0]
    cp = 32 if cmid <= 32 else 48
    if cmid > 48:
        raise NotImplementedError(f"patch embed with a {cmid}-channel first conv is not instantiated (<= 48)")
    w0 = w0.detach().float()
    if s0 is not None:
        w0 = w0 * s0.view(-1, 1, 1, 1)
    w27 = torch.zeros(27, cp, device=w0.device, dtype=torch.float32)
    w27[:, :cmid] = w0.reshape(cmid, 27).t()
    bias = torch.zeros(cp, device=w0.device, dtype=torch.float32)
    if b0 is not None:
        bias[:cmid] = b0
    w9 = torch.zeros(9, cout, cp, device=w0.device, dtype=torch.bfloat16)
    w9[:, :, :cmid] = w1.detach().permute(2, 3, 0, 1).reshape(9, cout, cmid).to(torch.bfloat16)
    return w27.contiguous(), bias, w9.contiguous()


def conv3x3_weight(conv: nn.Conv2d) -> torch.Tensor:
    """dense 3x3 weight [N,C,3,3] -> bf16 [N, 9*C] with k = (ky*3+kx)*C + c."""
    w = conv.weight.detach()
    n, c = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(n, 9 * c).to(torch.bfloat16).contiguous()


def params_fingerprint(module: nn.Module):
    """Cheap change detector for cached packed weights: (data_ptr, _version) of every tensor."""
    fp = []
    for t in list(module.parameters()) + list(module.buffers()):
        fp.append((t.data_ptr(), t._version))
    return tuple(fp)


class NativePlanMixin:
    """Caches packed / folded weights; rebuilds when parameters, device or mode change."""

    def _plan(self):
        from . import ops
        key = (params_fingerprint(self), self.training, ops.precision())
        if getattr(self, "_plan_key", None) != key:
            self._plan_cache = self._build_plan()
            self._plan_key = key
        return self._plan_cache

    def _require_eval(self, what: str):
        if self.training:
            raise NotImplementedError(
                f"{what}: this module's native sm_100a path is eval-mode only (train-mode forward / backward exist for the "
                "stage-1 student encoders: ImageStudentEncoder; see DESIGN.md).  Call .eval() first.")
