"""Train-mode forward + backward of the TinyViT students (tiny_vit_5m / 11m / 21m, stage-1 configs es_tv_{s,m,l}) on libes3.so.
Same scheme as efficientvit_train.py / repvit_train.py: units with forward(x) / backward(d, grads), wrapped into ONE autograd node
by stage1.model.StudentTrainFunction.  Tokens [B, L, C] and NHWC maps [B, H, W, C] are the same memory.

  patch embed / MBConv / PatchMerging / local_conv   the Conv2d_BN units of the other students (ConvUnit); MBConv applies its last
                                                      GELU after the shortcut add (tiny_vit.py:112-125): a second es3_affine_act pass
  nn.Linear (+ GELU)                                  es3_gemm_bf16 [+ es3_affine_act]; gradients: es3_wgrad_pw, column sums, es3_gemm_bf16 on W^T
  nn.LayerNorm                                        es3_layernorm_bf16 / es3_layernorm_bwd
  window attention                                    the token map is zero-padded to a window multiple BEFORE the attention LayerNorm,
                                                      exactly as the reference does (tiny_vit.py:352-360), so padded tokens need no special
                                                      case: es3_win_attn_bias_bf16 / es3_win_attn_bias_bwd on the padded map, crop afterwards;
                                                      d attention_biases = scatter-add of the [heads, N, N] bias gradient over attention_bias_idxs
  DropPath (timm, tiny_vit.py:56-64)                  per-sample Bernoulli gate through es3_scale_channels (forward and backward)

GPU parity of the whole training step and of es3_layernorm_bwd / es3_win_attn_bias_bwd: tests/test_zz_train_gpu.py (green on a B200 since round 2); the graph logic is also checked in fp64 on CPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .efficientvit_train import ConvUnit, _grad_of
from .repvit_train import PatchEmbedUnit
from .tiny_vit import MBConv, PatchMerging, TinyViTBlock

__all__ = ["TinyViTTrainGraph"]


def _cu(cb, act, kind):
    return ConvUnit(cb.c, cb.bn, act, kind)


def _colsum(d2, out):
    """out[c] += sum_rows d2[:, c]  (bias gradients) on the verified column-reduction kernel."""
    if out is not None:
        ops.bn_act_bwd(d2, d2, None, None, None, "none", dbeta=out, apply=False)


class DropPath:
    """timm DropPath: x * bernoulli(keep)[b] / keep, one draw per sample; identity for rate 0."""

    def __init__(self, rate: float):
        self.rate = float(rate)
        self.gate = None

    def forward(self, x):                      # x [B, ..., C] bf16
        if self.rate <= 0.0:
            return x
        keep = 1.0 - self.rate
        B, C = x.shape[0], x.shape[-1]
        mask = torch.empty((B, 1), device=x.device, dtype=torch.float32).bernoulli_(keep) / keep
        self.gate = mask.expand(B, C).contiguous()
        return ops.scale_channels(x.reshape(B, 1, -1, C), self.gate).view(x.shape)

    def backward(self, d):
        if self.rate <= 0.0:
            return d
        B, C = d.shape[0], d.shape[-1]
        return ops.scale_channels(d.reshape(B, 1, -1, C).contiguous(), self.gate).view(d.shape)


class LinearUnit:
    """nn.Linear [+ activation] on [M, K] rows."""

    def __init__(self, lin: nn.Linear, act=None):
        self.lin, self.act = lin, act
        self.saved = None

    def forward(self, x2, residual=None):
        lin = self.lin
        self.w = lin.weight.detach().to(torch.bfloat16).contiguous()             # [N, K]
        b = lin.bias.detach().float().contiguous()
        if self.act is None:
            self.saved = (x2, None, b)
            return ops.gemm(x2, self.w, bias=b, residual=residual)
        assert residual is None
        z = ops.gemm(x2, self.w)
        self.saved = (x2, z, b)
        return ops.affine_act(z, None, b, self.act)

    def backward(self, dy2, grads, dx_residual=None):
        x2, z, b = self.saved
        self.saved = None
        lin = self.lin
        gb = _grad_of(grads, lin.bias)
        if z is None:
            dz = dy2.contiguous()
            _colsum(dz, gb)
        else:
            dz = ops.bn_act_bwd(dy2.contiguous(), z, None, b, self.act, "none", dbeta=gb)
        gw = _grad_of(grads, lin.weight)
        if gw is not None:
            ops.wgrad_pw(dz, x2, gw)
        return ops.gemm(dz, self.w.t().contiguous(), residual=dx_residual)


class LayerNormUnit:
    def __init__(self, ln: nn.LayerNorm):
        self.ln = ln
        self.saved = None

    def forward(self, x2):
        ln = self.ln
        self.saved = x2
        return ops.layernorm_bf16(x2, ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous(), ln.eps)

    def backward(self, dy2, grads, dres=None):
        ln, x2 = self.ln, self.saved
        self.saved = None
        return ops.layernorm_bwd(x2, dy2.contiguous(), ln.weight.detach().float().contiguous(), ln.eps, _grad_of(grads, ln.weight),
                                 _grad_of(grads, ln.bias), dres)


class MBConvUnit:
    """act3(drop_path(conv3(act2(conv2(act1(conv1(x)))))) + x)  (tiny_vit.py:87-125)."""

    def __init__(self, m: MBConv):
        self.c1, self.c2, self.c3 = _cu(m.conv1, "gelu", "pw"), _cu(m.conv2, "gelu", "dw"), _cu(m.conv3, None, "pw")
        self.dp = DropPath(getattr(m, "drop_path_rate", 0.0))
        self.saved = None

    def forward(self, x):
        y = self.dp.forward(self.c3.forward(self.c2.forward(self.c1.forward(x))))
        B, H, W, C = x.shape
        t = ops.add_bf16(y.view(-1, C), x.view(-1, C)).view(B, H, W, C)
        self.saved = t
        return ops.affine_act(t, None, None, "gelu")

    def backward(self, dout, grads):
        t = self.saved
        self.saved = None
        dt = ops.bn_act_bwd(dout.contiguous(), t, None, None, "gelu", "none")
        d = self.c2.backward(self.c3.backward(self.dp.backward(dt), grads), grads)
        return self.c1.backward(d, grads, dx_residual=dt)


class PatchMergingUnit:
    """conv1 (1x1) -> GELU -> conv2 (dw 3x3, stride 2) -> GELU -> conv3 (1x1)  (tiny_vit.py:128-154)."""

    def __init__(self, m: PatchMerging):
        self.c1, self.c2, self.c3 = _cu(m.conv1, "gelu", "pw"), _cu(m.conv2, "gelu", "dw"), _cu(m.conv3, None, "pw")

    def forward(self, x):
        return self.c3.forward(self.c2.forward(self.c1.forward(x)))

    def backward(self, d, grads):
        return self.c1.backward(self.c2.backward(self.c3.backward(d, grads), grads), grads)


class BlockUnit:
    """TinyViTBlock (tiny_vit.py:296-386): x + dp(attn(windows(pad(x)))); local_conv; x + dp(mlp(x))."""

    def __init__(self, blk: TinyViTBlock):
        at = blk.attn
        if at.key_dim != 32 or blk.window_size not in (7, 14):
            raise NotImplementedError("native TinyViT attention: head_dim 32, window 7 or 14")
        self.blk = blk
        self.n1, self.qkv, self.proj = LayerNormUnit(at.norm), LinearUnit(at.qkv), LinearUnit(at.proj)
        self.local = _cu(blk.local_conv, None, "dw")
        self.n2, self.fc1, self.fc2 = LayerNormUnit(blk.mlp.norm), LinearUnit(blk.mlp.fc1, "gelu"), LinearUnit(blk.mlp.fc2)
        rate = getattr(blk, "drop_path_rate", 0.0)
        self.dp1, self.dp2 = DropPath(rate), DropPath(rate)
        self.saved = None

    def forward(self, x):
        blk, at = self.blk, self.blk.attn
        B, H, W, C = x.shape
        ws = blk.window_size
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        if (Hp, Wp) != (H, W):                               # F.pad(x, (0, 0, 0, pad_r, 0, pad_b)) with zeros
            xp = torch.zeros((B, Hp, Wp, C), device=x.device, dtype=x.dtype)
            xp[:, :H, :W] = x
        else:
            xp = x
        y = self.n1.forward(xp.view(-1, C))
        qkv = self.qkv.forward(y)
        bias = at.attention_biases.detach().float()[:, at.attention_bias_idxs].contiguous()          # [heads, N, N]
        pad_row = torch.zeros(3 * C, device=x.device, dtype=qkv.dtype)                               # no token lies outside the padded map
        a = ops.win_attn_bias(qkv, pad_row, bias, B, Hp, Wp, C, at.num_heads, ws, at.scale)
        o = self.proj.forward(a).view(B, Hp, Wp, C)
        if (Hp, Wp) != (H, W):
            o = o[:, :H, :W].contiguous()
        x1 = ops.add_bf16(x.view(-1, C), self.dp1.forward(o).view(-1, C)).view(B, H, W, C)
        x2 = self.local.forward(x1)
        h = self.fc1.forward(self.n2.forward(x2.view(-1, C)))
        x3 = ops.add_bf16(x2.view(-1, C), self.dp2.forward(self.fc2.forward(h).view(B, H, W, C)).view(-1, C)).view(B, H, W, C)
        self.saved = (qkv, bias, (B, H, W, C, Hp, Wp))
        return x3

    def backward(self, d3, grads):
        blk, at = self.blk, self.blk.attn
        qkv, bias, (B, H, W, C, Hp, Wp) = self.saved
        self.saved = None
        ws = blk.window_size
        d3 = d3.contiguous()
        # x3 = x2 + dp(fc2(gelu(fc1(LN(x2)))))
        dh = self.fc2.backward(self.dp2.backward(d3).view(-1, C), grads)
        dy2 = self.fc1.backward(dh, grads)
        dx2 = self.n2.backward(dy2, grads, dres=d3.view(-1, C)).view(B, H, W, C)
        dx1 = self.local.backward(dx2, grads)
        # x1 = x + dp(crop(proj(attn(qkv(LN(pad(x)))))))
        do = self.dp1.backward(dx1)
        if (Hp, Wp) != (H, W):
            dop = torch.zeros((B, Hp, Wp, C), device=do.device, dtype=do.dtype)
            dop[:, :H, :W] = do
        else:
            dop = do
        da = self.proj.backward(dop.view(-1, C), grads)
        dqkv, dbias = ops.win_attn_bias_bwd(qkv, da.contiguous(), bias, B, Hp, Wp, C, at.num_heads, ws, at.scale)
        g_ab = _grad_of(grads, at.attention_biases)
        if g_ab is not None:                                 # bias[h, i, j] = attention_biases[h, idxs[i, j]]
            g_ab.index_add_(1, at.attention_bias_idxs.reshape(-1), dbias.reshape(at.num_heads, -1).to(g_ab.dtype))
        dy1 = self.qkv.backward(dqkv, grads)
        dxp = self.n1.backward(dy1, grads).view(B, Hp, Wp, C)
        if (Hp, Wp) != (H, W):
            dxp = dxp[:, :H, :W].contiguous()
        return ops.add_bf16(dxp.view(-1, C), dx1.view(-1, C)).view(B, H, W, C)


class TinyViTTrainGraph:
    """patch_embed -> layers[0] (MBConv blocks) -> downsample -> layers[1..] (TinyViTBlocks [+ downsample])  (tiny_vit.py:460-607)."""

    def __init__(self, model):
        self.patch = PatchEmbedUnit(model.patch_embed.seq)
        self.units = []
        for li, layer in enumerate(model.layers):
            for blk in layer.blocks:
                self.units.append(MBConvUnit(blk) if li == 0 else BlockUnit(blk))
            if layer.downsample is not None:
                self.units.append(PatchMergingUnit(layer.downsample))

    def forward(self, x):
        x = self.patch.forward(x)
        for u in self.units:
            x = u.forward(x)
        return x

    def backward(self, d, grads):
        for u in reversed(self.units):
            d = u.backward(d, grads)
        self.patch.backward(d, grads)
