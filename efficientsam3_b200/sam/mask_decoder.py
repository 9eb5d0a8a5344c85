"""MaskDecoder, B200-native.  Module tree / keys of sam3/sam3/sam/mask_decoder.py (MaskDecoder :10-292, MLP :297-319)
in the configuration Sam3TrackerBase._build_sam_heads builds (sam3_tracker_base.py:179-218).

predict_masks (:165-242): tokens = [obj, iou, 4 mask, prompts]; src = image_embeddings + dense;  TwoWayTransformer;
upscaling gelu(LN2d(convT(src) + feat_s1)) -> gelu(convT(.) + feat_s0) on the tcgen05 convT GEMM (depth-to-space
epilogue, fp32 output); 4 hypernetwork MLPs + IoU / object heads in fp32 (es3_gemm_simt); masks = hyper_in @ upscaled
(es3_hyper_masks)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..nn_utils import NativePlanMixin
from .common import LayerNorm2d


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, sigmoid_output=False):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


def _mlp_plan(m: MLP):
    return [(l.weight.detach().float().contiguous(), l.bias.detach().float().contiguous()) for l in m.layers], m.sigmoid_output


def _run_mlp(plan, x):
    layers, sig = plan
    for i, (w, b) in enumerate(layers):
        last = i == len(layers) - 1
        x = ops.gemm_simt(x, w, bias=b, act=("sigmoid" if (last and sig) else None) if last else "relu", out_dtype=torch.float32)
    return x


class MaskDecoder(nn.Module, NativePlanMixin):
    def __init__(self, *, transformer_dim, transformer, num_multimask_outputs=3, activation=nn.GELU, iou_head_depth=3,
                 iou_head_hidden_dim=256, use_high_res_features=False, iou_prediction_use_sigmoid=False,
                 dynamic_multimask_via_stability=False, dynamic_multimask_stability_delta=0.05,
                 dynamic_multimask_stability_thresh=0.98, pred_obj_scores=False, pred_obj_scores_mlp=False,
                 use_multimask_token_for_obj_ptr=False):
        super().__init__()
        if not (use_high_res_features and pred_obj_scores and pred_obj_scores_mlp) or activation is not nn.GELU \
                or dynamic_multimask_via_stability:
            raise NotImplementedError("native MaskDecoder covers the _build_sam_heads configuration "
                                      "(high-res features, object-score MLP, GELU, no dynamic multimask)")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.pred_obj_scores = True
        self.obj_score_token = nn.Embedding(1, transformer_dim)
        self.use_multimask_token_for_obj_ptr = use_multimask_token_for_obj_ptr
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4), activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2), activation())
        self.use_high_res_features = True
        self.conv_s0 = nn.Conv2d(transformer_dim, transformer_dim // 8, kernel_size=1, stride=1)
        self.conv_s1 = nn.Conv2d(transformer_dim, transformer_dim // 4, kernel_size=1, stride=1)
        self.output_hypernetworks_mlps = nn.ModuleList(
            [MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth,
                                       sigmoid_output=iou_prediction_use_sigmoid)
        self.pred_obj_score_head = MLP(transformer_dim, transformer_dim, 1, 3)
        self.dynamic_multimask_via_stability = False

    def _build_plan(self):
        dc1, ln1, _, dc2, _ = self.output_upscaling
        strict = ops.precision() == "strict"
        wdt = torch.float32 if strict else torch.bfloat16
        return dict(
            strict=strict,
            dc1_raw=(dc1.weight.detach().float().contiguous(), dc1.bias.detach().float().contiguous()),
            dc2_raw=(dc2.weight.detach().float().contiguous(), dc2.bias.detach().float().contiguous()),
            out_tok=torch.cat([self.obj_score_token.weight, self.iou_token.weight, self.mask_tokens.weight], 0).detach().float().contiguous(),
            dc1=(ops.convt2x2_weight(dc1.weight), dc1.bias.detach().float().repeat(4).contiguous()),
            ln1=(ln1.weight.detach().float().contiguous(), ln1.bias.detach().float().contiguous(), ln1.eps),
            dc2=(ops.convt2x2_weight(dc2.weight), dc2.bias.detach().float().repeat(4).contiguous()),
            hyper=[_mlp_plan(m) for m in self.output_hypernetworks_mlps], iou=_mlp_plan(self.iou_prediction_head),
            obj=_mlp_plan(self.pred_obj_score_head),
            s0=(self.conv_s0.weight.detach().reshape(self.conv_s0.out_channels, -1).to(wdt).contiguous(),
                self.conv_s0.bias.detach().float().contiguous()),
            s1=(self.conv_s1.weight.detach().reshape(self.conv_s1.out_channels, -1).to(wdt).contiguous(),
                self.conv_s1.bias.detach().float().contiguous()))

    @torch.no_grad()
    def project_high_res(self, feat_s0_nhwc, feat_s1_nhwc):
        """conv_s0 / conv_s1 (1x1) on NHWC bf16 FPN levels -> fp32 NHWC residuals for the upscaling path
        (applied once per image: sam3_tracker_base.py:445-466)."""
        p = self._plan()
        out = []
        for x, (w, b) in ((feat_s0_nhwc, p["s0"]), (feat_s1_nhwc, p["s1"])):
            B, H, W, C = x.shape
            if p["strict"]:
                out.append(ops.sgemm(x.float().view(-1, C), w, bias=b).view(B, H, W, -1))
            else:
                out.append(ops.gemm(x.view(-1, C), w, bias=b, out_dtype=torch.float32).view(B, H, W, -1))
        return out

    @torch.no_grad()
    def predict_tokens(self, keys_f32, keys_b16, key_pe_tokens, sparse, B, h, w, feat_s0, feat_s1, obj_gate=False,
                       multimask_output=True):
        """Token-major core.  keys: image_embeddings + dense, [B*h*w, C]; feat_s0 [B,4h,4w,C/8], feat_s1 [B,2h,2w,C/4]
        fp32 NHWC.  Returns (masks [B,K,4h,4w] fp32, iou [B,K], mask tokens [B,K',C], obj logits [B,1])."""
        p = self._plan()
        C = self.transformer_dim
        tokens = torch.cat([p["out_tok"].unsqueeze(0).expand(B, -1, -1), sparse.float()], dim=1).contiguous()
        hs, kf, kb = self.transformer.run_tokens(keys_f32, keys_b16, key_pe_tokens, tokens, B)
        iou_tok = hs[:, 1].contiguous()
        mask_toks = hs[:, 2:2 + self.num_mask_tokens]
        if p["strict"]:      # fp32 image stream (kf is what run_tokens returned as kb as well)
            up1 = ops.convt2x2_f32(kf.view(B, h, w, C), p["dc1_raw"][0], p["dc1_raw"][1], residual=feat_s1)
            up1 = ops.ln_rows_gelu_f32(up1.view(-1, C // 4), *p["ln1"]).view(B, 2 * h, 2 * w, C // 4)
            up2 = ops.convt2x2_f32(up1, p["dc2_raw"][0], p["dc2_raw"][1], act="gelu", residual=feat_s0, act_after_res=True)
        else:
            up1 = ops.convt2x2(kb.view(B, h, w, C), p["dc1"][0], bias4=p["dc1"][1], residual=feat_s1, out_dtype=torch.float32)
            up1 = ops.ln_rows_gelu(up1.view(-1, C // 4), *p["ln1"]).view(B, 2 * h, 2 * w, C // 4)
            up2 = ops.convt2x2(up1, p["dc2"][0], bias4=p["dc2"][1], act="gelu", residual=feat_s0, out_dtype=torch.float32,
                               act_after_res=True)
        hyper = torch.stack([_run_mlp(p["hyper"][i], mask_toks[:, i].contiguous()) for i in range(self.num_mask_tokens)], 1)
        iou = _run_mlp(p["iou"], iou_tok)
        obj = _run_mlp(p["obj"], hs[:, 0].contiguous())
        if multimask_output:
            K, off = self.num_mask_tokens - 1, 1
        else:
            K, off = 1, 0
        masks = ops.hyper_masks(up2.view(B, 16 * h * w, C // 8), hyper.contiguous(), obj.reshape(-1) if obj_gate else None,
                                -1024.0, K, off).view(B, K, 4 * h, 4 * w)
        iou = iou[:, off:off + K]
        toks = mask_toks[:, 1:] if (multimask_output and self.use_multimask_token_for_obj_ptr) else mask_toks[:, 0:1]
        return masks, iou, toks, obj

    @torch.no_grad()
    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output,
                repeat_image, high_res_features: Optional[List[torch.Tensor]] = None):
        """Reference signature (mask_decoder.py:107-163); NCHW fp32 CUDA tensors in, NCHW fp32 out."""
        self._require_eval("MaskDecoder.forward")
        assert image_pe.size(0) == 1, "image_pe should have size 1 in batch dim (from `get_dense_pe()`)"
        B = sparse_prompt_embeddings.shape[0]
        _, C, h, w = image_embeddings.shape
        if repeat_image:      # one image, B prompts (mask_decoder.py:185-188): the broadcast add below repeats the image
            assert image_embeddings.shape[0] == 1, "repeat_image expects a single image embedding"
        else:
            assert image_embeddings.shape[0] == B
        src = image_embeddings.float() + dense_prompt_embeddings.float()   # torch broadcast add: plumbing at the API edge
        if src.shape[0] != B:
            src = src.expand(B, -1, -1, -1)
        keys_f32, keys_b16 = ops.nchw_to_tokens(src.contiguous())
        pe_tok, _ = ops.nchw_to_tokens(image_pe.float(), out_bf16=False)
        f0, f1 = high_res_features
        feat_s0 = f0.float().expand(B, -1, -1, -1).permute(0, 2, 3, 1).contiguous()
        feat_s1 = f1.float().expand(B, -1, -1, -1).permute(0, 2, 3, 1).contiguous()
        return self.predict_tokens(keys_f32, keys_b16, pe_tok, sparse_prompt_embeddings, B, h, w, feat_s0, feat_s1,
                                   obj_gate=False, multimask_output=multimask_output)
