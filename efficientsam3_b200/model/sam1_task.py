"""Point-prompt segmentation pipeline (BASELINE config 3): SAM3 ViT trunk -> SAM2-branch FPN -> SAM heads, batched.

Native counterpart of the reference call chain
  SAM3InteractiveImagePredictor.set_image_batch   sam3/sam3/model/sam1_task_predictor.py:121-166
  Sam3TrackerBase.forward_image (conv_s0/conv_s1)  sam3/sam3/model/sam3_tracker_base.py:445-466
  SAM3InteractiveImagePredictor._predict           sam1_task_predictor.py:329-430
  Sam3TrackerBase._forward_sam_heads               sam3_tracker_base.py:220-389
with the reference's per-image Python loop (sam1_task_predictor.py:168-228) replaced by one batched pass
(B images x 1 prompt each, `repeat_image=False` semantics).  Parameter names follow the reference tracker
(`sam_prompt_encoder.*`, `sam_mask_decoder.*`, `no_mem_embed`, `backbone.vision_backbone.{trunk,convs,sam2_convs}.*`).
Not built: connected-component hole filling (sam1_utils.py:84-105), box / mask prompts, tracker memory (obj_ptr).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..sam import MaskDecoder, PromptEncoder, TwoWayTransformer
from .necks import Sam3DualViTDetNeck
from .vitdet import create_sam3_vit_backbone

NO_OBJ_SCORE = -1024.0


class _Holder(nn.Module):
    pass


class Sam3PointPromptSegmenter(nn.Module):
    def __init__(self, image_size=1008, backbone_stride=14, hidden_dim=256, vit_overrides=None):
        super().__init__()
        self.image_size, self.hidden_dim = image_size, hidden_dim
        e = image_size // backbone_stride
        trunk = create_sam3_vit_backbone(**(vit_overrides or {}))
        self.backbone = _Holder()
        self.backbone.vision_backbone = Sam3DualViTDetNeck(trunk=trunk, position_encoding=None, d_model=hidden_dim,
                                                           scale_factors=[4.0, 2.0, 1.0, 0.5], add_sam2_neck=True)
        self.no_mem_embed = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        nn.init.trunc_normal_(self.no_mem_embed, std=0.02)
        self.sam_prompt_encoder = PromptEncoder(embed_dim=hidden_dim, image_embedding_size=(e, e),
                                                input_image_size=(image_size, image_size), mask_in_chans=16)
        self.sam_mask_decoder = MaskDecoder(
            num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=hidden_dim, mlp_dim=2048, num_heads=8),
            transformer_dim=hidden_dim, iou_head_depth=3, iou_head_hidden_dim=256, use_high_res_features=True,
            iou_prediction_use_sigmoid=True, pred_obj_scores=True, pred_obj_scores_mlp=True,
            use_multimask_token_for_obj_ptr=True)
        self._features = None
        self.eval()

    @torch.no_grad()
    def set_image_batch(self, images: torch.Tensor):
        """images [B,3,S,S] fp32 CUDA, already resized / normalised (mean 0.5, std 0.5 as Sam3Processor does)."""
        neck = self.backbone.vision_backbone
        tok, (B, h, w) = neck.trunk.forward_tokens(images)
        feats = ops.add_rows(tok, None, out_bf16=True, out_f32=False)[0].view(B, h, w, -1)
        l288, l144, l72 = neck.forward_nhwc(feats, "sam2", (0, 1, 2), f32_levels=(2,))
        md = self.sam_mask_decoder
        feat_s0, feat_s1 = md.project_high_res(l288, l144)
        # image_embed = 72^2 level + no_mem_embed (sam1_task_predictor.py:157) + no_mask dense embedding (mask_decoder.py:189)
        addc = (self.no_mem_embed.detach().reshape(-1) + self.sam_prompt_encoder.no_mask_embed.weight.detach().reshape(-1)).float().contiguous()
        C = l72.shape[-1]
        keys_b16, keys_f32 = ops.add_rows(l72.view(-1, C), addc.view(1, C), out_bf16=True, out_f32=True)
        self._features = dict(B=B, h=h, w=w, keys_f32=keys_f32, keys_b16=keys_b16, feat_s0=feat_s0, feat_s1=feat_s1,
                              pe=self.sam_prompt_encoder.pe_layer.tokens((h, w)))
        return self

    @torch.no_grad()
    def predict_batch(self, point_coords, point_labels, multimask_output=True, return_logits=False):
        """point_coords [B,P,2] (x,y in input-image pixels), point_labels [B,P] -> dict with low-res logits
        [B,K,4h,4w], high-res logits or bool masks [B,K,S,S], ious [B,K], object logits [B,1], best index [B]."""
        f = self._features
        assert f is not None, "call set_image_batch first"
        sparse, _ = self.sam_prompt_encoder(points=(point_coords, point_labels), boxes=None, masks=None)
        low, iou, toks, obj = self.sam_mask_decoder.predict_tokens(
            f["keys_f32"], f["keys_b16"], f["pe"], sparse, f["B"], f["h"], f["w"], f["feat_s0"], f["feat_s1"],
            obj_gate=True, multimask_output=multimask_output)
        S = self.image_size
        high, binm = ops.bilinear_nchw(low, S, S, binarize_thr=None if return_logits else 0.0, want_float=return_logits)
        best = torch.argmax(iou, dim=-1)
        return dict(low_res_multimasks=low, high_res=high if return_logits else binm.bool(), ious=iou,
                    object_score_logits=obj, best=best, sam_tokens=toks)
