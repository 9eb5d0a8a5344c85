"""TwoWayTransformer, B200-native.  Module tree / state_dict keys of sam3/sam3/sam/transformer.py
(TwoWayTransformer :16-106, TwoWayAttentionBlock :109-182, Attention :185-264).

Execution split: the 8 prompt/output tokens per image are fp32 on CUDA cores (es3_gemm_simt, es3_attn_few_queries,
es3_layernorm_f32); the 5184 image tokens go through the tcgen05 GEMM (k/v/q projections, out_proj + fp32
residual) with es3_attn_few_keys for image->token attention.  The image stream is fp32 in HBM (+ a bf16 copy as
GEMM operand), like the ViT residual stream.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..nn_utils import NativePlanMixin
from .common import MLPBlock


class Attention(nn.Module):
    def __init__(self, embedding_dim, num_heads, downsample_rate=1, dropout=0.0, kv_in_dim=None, use_fa3=False):
        super().__init__()
        assert dropout == 0.0 and kv_in_dim in (None, embedding_dim) and not use_fa3
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)

    @property
    def scale(self):
        return (self.internal_dim // self.num_heads) ** -0.5


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, embedding_dim, num_heads, mlp_dim=2048, activation=nn.ReLU, attention_downsample_rate=2,
                 skip_first_layer_pe=False):
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe


def _f32(lin: nn.Linear):
    return lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()


def _b16(lin: nn.Linear):
    """Weights of a projection over the 5184 image tokens: bf16 for the tcgen05 GEMM, fp32 in the strict precision mode."""
    if ops.precision() == "strict":
        return _f32(lin)
    return lin.weight.detach().to(torch.bfloat16).contiguous(), lin.bias.detach().float().contiguous()


def _img_gemm(x, wb, **kw):
    """Image-token projection: es3_gemm_bf16 on bf16 operands, es3_sgemm_f32 on fp32 operands (strict mode)."""
    if wb[0].dtype == torch.float32:
        kw.pop("out_dtype", None)
        return ops.sgemm(x, wb[0], bias=wb[1], **kw)
    return ops.gemm(x, wb[0], bias=wb[1], **kw)


def _ln(n: nn.LayerNorm):
    return n.weight.detach().float().contiguous(), n.bias.detach().float().contiguous(), n.eps


class TwoWayTransformer(nn.Module, NativePlanMixin):
    def __init__(self, depth, embedding_dim, num_heads, mlp_dim, activation=nn.ReLU, attention_downsample_rate=2):
        super().__init__()
        self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = depth, embedding_dim, num_heads, mlp_dim
        self.layers = nn.ModuleList([
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, activation, attention_downsample_rate,
                                 skip_first_layer_pe=(i == 0)) for i in range(depth)])
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def _build_plan(self):
        def tok_attn(a):   # all-token attention: fp32 weights for gemm_simt
            return dict(q=_f32(a.q_proj), k=_f32(a.k_proj), v=_f32(a.v_proj), o=_f32(a.out_proj), scale=a.scale)

        def t2i(a):        # token queries (fp32), image keys/values (tcgen05)
            return dict(q=_f32(a.q_proj), k=_b16(a.k_proj), v=_b16(a.v_proj), o=_f32(a.out_proj), scale=a.scale)

        def i2t(a):        # image queries (tcgen05), token keys/values (fp32), image out_proj (tcgen05)
            return dict(q=_b16(a.q_proj), k=_f32(a.k_proj), v=_f32(a.v_proj), o=_b16(a.out_proj), scale=a.scale)

        layers = []
        for blk in self.layers:
            layers.append(dict(sa=tok_attn(blk.self_attn), n1=_ln(blk.norm1), t2i=t2i(blk.cross_attn_token_to_image),
                               n2=_ln(blk.norm2), l1=_f32(blk.mlp.lin1), l2=_f32(blk.mlp.lin2), act=blk.mlp.act_name,
                               n3=_ln(blk.norm3), n4=_ln(blk.norm4), i2t=i2t(blk.cross_attn_image_to_token),
                               skip_pe=blk.skip_first_layer_pe))
        return dict(layers=layers, final=t2i(self.final_attn_token_to_image), nf=_ln(self.norm_final_attn))

    # ---- token-major core: keys_f32/keys_b16 [B*HW, C], key_pe [R, C] fp32 (R = HW or B*HW), tokens [B,T,C] fp32
    def run_tokens(self, keys_f32, keys_b16, key_pe, tokens, B):
        p = self._plan()
        C, H = self.embedding_dim, self.num_heads
        T = tokens.shape[1]
        HW = keys_f32.shape[0] // B
        q_pe = tokens.reshape(B * T, C).contiguous()
        queries = q_pe

        def lin(x, wb, act=None, residual=None):
            return ops.gemm_simt(x, wb[0], bias=wb[1], act=act, residual=residual, out_dtype=torch.float32)

        def ln(x, n):
            return ops.layernorm(x, n[0], n[1], n[2], out_bf16=False, out_f32=True)[1]

        def token_to_image(a, queries, k_in_b16, keys_b16):
            qq = ops.add_rows(queries, q_pe)[1]
            q = lin(qq, a["q"]).view(B, T, -1)
            k = _img_gemm(k_in_b16, a["k"]).view(B, HW, -1)
            v = _img_gemm(keys_b16, a["v"]).view(B, HW, -1)
            o = ops.attn_few_queries(q, k, v, H, a["scale"]).view(B * T, -1)
            return lin(o, a["o"], residual=queries)

        strict = ops.precision() == "strict"       # fp32 image stream: the "b16" operands below ARE the fp32 tensors
        if strict:
            keys_b16 = keys_f32

        def key_in():
            if strict:
                return ops.add_rows(keys_f32, key_pe)[1]
            return ops.add_rows(keys_f32, key_pe, out_bf16=True, out_f32=False)[0]

        for lp in p["layers"]:
            sa = lp["sa"]
            if lp["skip_pe"]:
                q_in = queries
                res = None
            else:
                q_in = ops.add_rows(queries, q_pe)[1]
                res = queries
            q = lin(q_in, sa["q"]).view(B, T, -1)
            k = lin(q_in, sa["k"]).view(B, T, -1)
            v = lin(queries, sa["v"]).view(B, T, -1)
            o = ops.attn_few_queries(q, k, v, H, sa["scale"]).view(B * T, -1)
            queries = ln(lin(o, sa["o"], residual=res), lp["n1"])
            # token -> image
            k_in_b16 = key_in()
            queries = ln(token_to_image(lp["t2i"], queries, k_in_b16, keys_b16), lp["n2"])
            # mlp
            hdn = lin(queries, lp["l1"], act=lp["act"])
            queries = ln(lin(hdn, lp["l2"], residual=queries), lp["n3"])
            # image -> token
            a = lp["i2t"]
            qq = ops.add_rows(queries, q_pe)[1]
            kt = lin(qq, a["k"]).view(B, T, -1)
            vt = lin(queries, a["v"]).view(B, T, -1)
            qi = _img_gemm(k_in_b16, a["q"])
            att = (ops.attn_few_keys_f32 if strict else ops.attn_few_keys)(qi, kt, vt, B, H, a["scale"])
            kx = _img_gemm(att, a["o"], residual=keys_f32, out_dtype=torch.float32)
            keys_b16, keys_f32 = ops.layernorm(kx, *lp["n4"], out_bf16=not strict, out_f32=True)
            if strict:
                keys_b16 = keys_f32
        k_in_b16 = key_in()
        queries = ln(token_to_image(p["final"], queries, k_in_b16, keys_b16), p["nf"])
        return queries.view(B, T, C), keys_f32, keys_b16

    @torch.no_grad()
    def forward(self, image_embedding, image_pe, point_embedding):
        """image_embedding, image_pe: [B,C,h,w] fp32 CUDA; point_embedding [B,N,C] fp32 -> (queries [B,N,C],
        keys [B,h*w,C]) fp32, as transformer.py:62-106."""
        self._require_eval("TwoWayTransformer.forward")
        B, C, h, w = image_embedding.shape
        keys_f32, keys_b16 = ops.nchw_to_tokens(image_embedding.float())
        pe_f32, _ = ops.nchw_to_tokens(image_pe.float().expand(B, -1, -1, -1), out_bf16=False)
        q, kf, _ = self.run_tokens(keys_f32, keys_b16, pe_f32, point_embedding.float().contiguous(), B)
        return q, kf.view(B, h * w, C)
