"""Oracle: SAM-style prompt encoder + TwoWayTransformer + MaskDecoder (TEST INFRASTRUCTURE ONLY).

Functional fp32 restatement driven by reference-keyed state_dicts:
  PositionEmbeddingRandom / PromptEncoder   sam3/sam3/sam/prompt_encoder.py:200-243, 12-197 (points + no-mask path)
  Attention / TwoWayAttentionBlock / TwoWayTransformer   sam3/sam3/sam/transformer.py:185-264, 109-182, 16-106
  MaskDecoder.predict_masks / forward / MLP / LayerNorm2d sam3/sam3/sam/mask_decoder.py:165-242, 107-163, 297-319;
                                                          sam/common.py:27-39
  head glue (obj-score gating, upsample, best-IoU pick)   sam3/sam3/model/sam3_tracker_base.py:220-389
Configuration as built by Sam3TrackerBase._build_sam_heads (:179-218): depth 2, dim 256, 8 heads, mlp 2048,
ReLU MLP, downsample 2, 3 multimask outputs, high-res features, sigmoid IoU, obj-score MLP.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NO_OBJ_SCORE = -1024.0


# ----------------------------------------------------------------------------- prompt encoder
def pe_encoding(gauss, coords01):
    c = 2 * coords01 - 1
    c = c @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd, p, h, w):
    g = sd[p + "pe_layer.positional_encoding_gaussian_matrix"]
    ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
    grid = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=-1)
    return pe_encoding(g, grid).permute(2, 0, 1).unsqueeze(0)       # [1, C, h, w]


def prompt_encoder_points(sd, p, coords, labels, image_size, embed_hw):
    """coords [B,P,2] (x,y) pixels, labels [B,P] in {-1,0,1,2,3}; boxes=None -> a padding point is appended."""
    B = coords.shape[0]
    pts = coords + 0.5
    pts = torch.cat([pts, torch.zeros(B, 1, 2)], dim=1)
    lab = torch.cat([labels, -torch.ones(B, 1, dtype=labels.dtype)], dim=1)
    c01 = pts.clone()
    c01[:, :, 0] = c01[:, :, 0] / image_size[1]
    c01[:, :, 1] = c01[:, :, 1] / image_size[0]
    e = pe_encoding(sd[p + "pe_layer.positional_encoding_gaussian_matrix"], c01.float())
    e = torch.where((lab == -1).unsqueeze(-1), torch.zeros_like(e) + sd[p + "not_a_point_embed.weight"], e)
    for i in range(4):
        e = torch.where((lab == i).unsqueeze(-1), e + sd[p + f"point_embeddings.{i}.weight"], e)
    dense = sd[p + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, embed_hw[0], embed_hw[1])
    return e, dense


def _embed_coords(sd, p, pts, lab, image_size):
    c01 = pts.clone()
    c01[:, :, 0] = c01[:, :, 0] / image_size[1]
    c01[:, :, 1] = c01[:, :, 1] / image_size[0]
    e = pe_encoding(sd[p + "pe_layer.positional_encoding_gaussian_matrix"], c01.float())
    e = torch.where((lab == -1).unsqueeze(-1), torch.zeros_like(e) + sd[p + "not_a_point_embed.weight"], e)
    for i in range(4):
        e = torch.where((lab == i).unsqueeze(-1), e + sd[p + f"point_embeddings.{i}.weight"], e)
    return e


def mask_downscaling(sd, p, masks):
    """PromptEncoder._embed_masks (prompt_encoder.py:45-63, 133-136): conv2x2s2 -> LN2d -> GELU -> conv2x2s2 -> LN2d -> GELU -> 1x1."""
    q = p + "mask_downscaling."
    x = F.conv2d(masks, sd[q + "0.weight"], sd[q + "0.bias"], stride=2)
    x = F.gelu(layernorm2d(sd, q + "1", x))
    x = F.conv2d(x, sd[q + "3.weight"], sd[q + "3.bias"], stride=2)
    x = F.gelu(layernorm2d(sd, q + "4", x))
    return F.conv2d(x, sd[q + "6.weight"], sd[q + "6.bias"])


def prompt_encoder(sd, p, points, boxes, masks, image_size, embed_hw):
    """PromptEncoder.forward (prompt_encoder.py:152-197): points get a padding point only when no boxes are given;
    box corners are embedded with labels 2 / 3; dense = mask_downscaling(masks) or the broadcast no_mask_embed."""
    if points is not None:
        bs = points[0].shape[0]
    elif boxes is not None:
        bs = boxes.shape[0]
    elif masks is not None:
        bs = masks.shape[0]
    else:
        bs = 1
    C = sd[p + "no_mask_embed.weight"].shape[1]
    sparse = torch.empty(bs, 0, C)
    if points is not None:
        coords, labels = points
        pts = coords + 0.5
        lab = labels
        if boxes is None:
            pts = torch.cat([pts, torch.zeros(bs, 1, 2)], dim=1)
            lab = torch.cat([lab, -torch.ones(bs, 1, dtype=lab.dtype)], dim=1)
        sparse = torch.cat([sparse, _embed_coords(sd, p, pts, lab, image_size)], dim=1)
    if boxes is not None:
        corners = (boxes + 0.5).reshape(-1, 2, 2)
        lab = torch.tensor([[2, 3]]).expand(corners.shape[0], 2)
        sparse = torch.cat([sparse, _embed_coords(sd, p, corners, lab, image_size)], dim=1)
    if masks is not None:
        dense = mask_downscaling(sd, p, masks)
    else:
        dense = sd[p + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bs, -1, embed_hw[0], embed_hw[1])
    return sparse, dense


# ----------------------------------------------------------------------------- transformer
def attn(sd, p, q, k, v, heads):
    q = F.linear(q, sd[p + ".q_proj.weight"], sd[p + ".q_proj.bias"])
    k = F.linear(k, sd[p + ".k_proj.weight"], sd[p + ".k_proj.bias"])
    v = F.linear(v, sd[p + ".v_proj.weight"], sd[p + ".v_proj.bias"])
    sep = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)
    o = F.scaled_dot_product_attention(sep(q), sep(k), sep(v))
    o = o.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def two_way_transformer(sd, p, image_embedding, image_pe, point_embedding, depth=2, heads=8):
    keys = image_embedding.flatten(2).permute(0, 2, 1)
    key_pe = image_pe.flatten(2).permute(0, 2, 1)
    queries, query_pe = point_embedding, point_embedding
    for i in range(depth):
        q_ = f"{p}layers.{i}"
        if i == 0:
            queries = attn(sd, q_ + ".self_attn", queries, queries, queries, heads)
        else:
            qq = queries + query_pe
            queries = queries + attn(sd, q_ + ".self_attn", qq, qq, queries, heads)
        queries = ln(sd, q_ + ".norm1", queries)
        queries = queries + attn(sd, q_ + ".cross_attn_token_to_image", queries + query_pe, keys + key_pe, keys, heads)
        queries = ln(sd, q_ + ".norm2", queries)
        m = F.linear(F.relu(F.linear(queries, sd[q_ + ".mlp.lin1.weight"], sd[q_ + ".mlp.lin1.bias"])),
                     sd[q_ + ".mlp.lin2.weight"], sd[q_ + ".mlp.lin2.bias"])
        queries = ln(sd, q_ + ".norm3", queries + m)
        keys = keys + attn(sd, q_ + ".cross_attn_image_to_token", keys + key_pe, queries + query_pe, queries, heads)
        keys = ln(sd, q_ + ".norm4", keys)
    queries = queries + attn(sd, p + "final_attn_token_to_image", queries + query_pe, keys + key_pe, keys, heads)
    return ln(sd, p + "norm_final_attn", queries), keys


# ----------------------------------------------------------------------------- mask decoder
def mlp(sd, p, x, n, sigmoid=False):
    for i in range(n):
        x = F.linear(x, sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return torch.sigmoid(x) if sigmoid else x


def layernorm2d(sd, p, x, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][:, None, None] * x + sd[p + ".bias"][:, None, None]


def mask_decoder(sd, p, image_embeddings, image_pe, sparse, dense, multimask_output, high_res_features, repeat_image=False):
    """MaskDecoder.forward with use_high_res_features, pred_obj_scores(+mlp), iou sigmoid,
    use_multimask_token_for_obj_ptr.  repeat_image: one image, several prompts (mask_decoder.py:185-188; the high-res
    features broadcast over the prompt batch).  Returns (masks, iou, sam_tokens, obj_logits)."""
    B = sparse.shape[0]
    if repeat_image:
        image_embeddings = torch.repeat_interleave(image_embeddings, B, dim=0)
    out_tok = torch.cat([sd[p + "obj_score_token.weight"], sd[p + "iou_token.weight"], sd[p + "mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tok.unsqueeze(0).expand(B, -1, -1), sparse], dim=1)
    src = image_embeddings + dense
    pos = image_pe.expand(B, -1, -1, -1)
    b, c, h, w = src.shape
    hs, src = two_way_transformer(sd, p + "transformer.", src, pos, tokens)
    iou_tok, mask_toks = hs[:, 1], hs[:, 2:6]
    src = src.transpose(1, 2).reshape(b, c, h, w)
    feat_s0, feat_s1 = high_res_features
    up = F.conv_transpose2d(src, sd[p + "output_upscaling.0.weight"], sd[p + "output_upscaling.0.bias"], stride=2)
    up = F.gelu(layernorm2d(sd, p + "output_upscaling.1", up + feat_s1))
    up = F.gelu(F.conv_transpose2d(up, sd[p + "output_upscaling.3.weight"], sd[p + "output_upscaling.3.bias"], stride=2) + feat_s0)
    hyper = torch.stack([mlp(sd, f"{p}output_hypernetworks_mlps.{i}", mask_toks[:, i], 3) for i in range(4)], dim=1)
    b2, c2, h2, w2 = up.shape
    masks = (hyper @ up.view(b2, c2, h2 * w2)).view(b2, -1, h2, w2)
    iou = mlp(sd, p + "iou_prediction_head", iou_tok, 3, sigmoid=True)
    obj = mlp(sd, p + "pred_obj_score_head", hs[:, 0], 3)
    if multimask_output:
        return masks[:, 1:], iou[:, 1:], mask_toks[:, 1:], obj
    return masks[:, 0:1], iou[:, 0:1], mask_toks[:, 0:1], obj


def fill_holes(masks, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0):
    """SAM2Transforms.postprocess_masks before the resize (sam1_utils.py:77-105).  Connected components are
    scikit-image's `label` with full (8-) connectivity in the reference's CPU backend (perflib/connected_components.py:
    18-29; the CUDA backends cc_torch / Triton use the same connectivity); restated here with scipy.ndimage.label and a
    3x3 structuring element.  Only `labels > 0` and the per-pixel component area are consumed, so numbering is irrelevant."""
    import numpy as np
    from scipy import ndimage

    def areas_of(binary):           # [N,H,W] bool -> per-pixel component size (0 on background)
        out = np.zeros(binary.shape, dtype=np.int64)
        for i, m in enumerate(binary):
            lab, n = ndimage.label(m, structure=np.ones((3, 3), dtype=np.int32))
            cnt = np.bincount(lab.ravel(), minlength=n + 1)
            cnt[0] = 0
            out[i] = cnt[lab]
        return torch.from_numpy(out)

    masks = masks.float()
    flat = masks.flatten(0, 1)
    if max_hole_area > 0:
        bg = (flat <= mask_threshold)
        a = areas_of(bg.numpy())
        is_hole = (bg & (a <= max_hole_area)).reshape(masks.shape)
        masks = torch.where(is_hole, torch.full_like(masks, mask_threshold + 10.0), masks)
    if max_sprinkle_area > 0:
        fg = (flat > mask_threshold)           # NB: the reference thresholds the ORIGINAL scores here (mask_flat)
        a = areas_of(fg.numpy())
        is_spr = (fg & (a <= max_sprinkle_area)).reshape(masks.shape)
        masks = torch.where(is_spr, torch.full_like(masks, mask_threshold - 10.0), masks)
    return masks


def predict(sd_pe, sd_md, image_embed, high_res_features, point_coords, point_labels, boxes, mask_input, image_size, orig_hw,
            multimask_output=True, return_logits=False, mask_threshold=0.0, max_hole_area=256.0, max_sprinkle_area=0.0):
    """SAM3InteractiveImagePredictor._predict for one image (sam1_task_predictor.py:329-430): boxes are merged in front of
    the points as label-2/3 corner points, the decoder runs with repeat_image when several prompts are given, masks are
    hole-filled at low resolution, resized to orig_hw and thresholded; low-res logits are clamped to +-32.
    image_embed [1,C,h,w] already includes no_mem_embed (set_image, :157)."""
    _, _, h, w = image_embed.shape
    pts = None
    if point_coords is not None:
        pts = (point_coords, point_labels)
    if boxes is not None:
        bc = boxes.reshape(-1, 2, 2)
        bl = torch.tensor([[2, 3]], dtype=torch.int32).repeat(boxes.shape[0], 1)
        pts = (torch.cat([bc, pts[0]], dim=1), torch.cat([bl, pts[1].to(torch.int32)], dim=1)) if pts is not None else (bc, bl)
    sparse, dense = prompt_encoder(sd_pe, "", pts, None, mask_input, (image_size, image_size), (h, w))
    batched = pts is not None and pts[0].shape[0] > 1
    pe = dense_pe(sd_pe, "", h, w)
    low, iou, _, _ = mask_decoder(sd_md, "", image_embed, pe, sparse, dense, multimask_output, high_res_features, repeat_image=batched)
    masks = fill_holes(low, mask_threshold, max_hole_area, max_sprinkle_area)
    masks = F.interpolate(masks, orig_hw, mode="bilinear", align_corners=False)
    low = torch.clamp(low, -32.0, 32.0)
    if not return_logits:
        masks = masks > mask_threshold
    return masks, iou, low


def high_res_from_fpn(sd, p, feat_288, feat_144):
    """conv_s0 / conv_s1 1x1 projections applied once per image (sam3_tracker_base.py:445-466)."""
    return (F.conv2d(feat_288, sd[p + "conv_s0.weight"], sd[p + "conv_s0.bias"]),
            F.conv2d(feat_144, sd[p + "conv_s1.weight"], sd[p + "conv_s1.bias"]))


def forward_sam_heads(sd_pe, sd_md, backbone_features, high_res_features, coords, labels, image_size,
                      multimask_output=True):
    """sam3_tracker_base.py:220-389 without the tracker-memory outputs (obj_ptr)."""
    B, _, h, w = backbone_features.shape
    sparse, dense = prompt_encoder_points(sd_pe, "", coords, labels, (image_size, image_size), (h, w))
    pe = dense_pe(sd_pe, "", h, w)
    low, iou, toks, obj = mask_decoder(sd_md, "", backbone_features, pe, sparse, dense, multimask_output, high_res_features)
    raw = low.float()
    low = torch.where((obj > 0)[:, None, None], low, torch.full_like(low, NO_OBJ_SCORE)).float()
    high = F.interpolate(low, size=(image_size, image_size), mode="bilinear", align_corners=False)
    best = torch.argmax(iou, dim=-1)
    idx = torch.arange(B)
    return dict(low_res_multimasks=low, high_res_multimasks=high, ious=iou, low_res_masks=low[idx, best].unsqueeze(1),
                high_res_masks=high[idx, best].unsqueeze(1), object_score_logits=obj, best=best, low_res_ungated=raw,
                high_res_ungated=F.interpolate(raw, size=(image_size, image_size), mode="bilinear", align_corners=False))
