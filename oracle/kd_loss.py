"""Oracle: stage-1 distillation loss, forward (TEST INFRASTRUCTURE ONLY).
Restates build_valid_mask / masked_mse / masked_cosine_loss of stage1/train_image_encoder_stage1.py:271-307 and
the combination at :205-210 (loss = mse + DISTILL.COSINE * cosine)."""
import torch
import torch.nn.functional as F


def build_valid_mask(img_size, img_size_before_pad, target_hw, device=None):
    B = len(img_size_before_pad)
    valid = torch.zeros(B, 1, img_size, img_size, device=device)
    for i in range(B):
        h, w = img_size_before_pad[i][1:]
        valid[i, :, :h, :w] = 1
    valid = F.interpolate(valid, size=target_hw, mode="bilinear", align_corners=False)
    return (valid > 0.5).float()


def masked_mse(preds, teacher, mask):
    diff = (preds - teacher) * mask
    denom = mask.sum(dim=(1, 2, 3)).clamp(min=1.0)
    return (diff.square().sum(dim=(1, 2, 3)) / denom).mean()


def masked_cosine_loss(preds, teacher, mask):
    loss = (1.0 - F.cosine_similarity(preds, teacher, dim=1)) * mask.squeeze(1)
    denom = mask.squeeze(1).sum(dim=(1, 2)).clamp(min=1.0)
    return (loss.sum(dim=(1, 2)) / denom).mean()


def kd_loss(preds, teacher, img_size, img_size_before_pad, cosine_weight=1.0):
    mask = build_valid_mask(img_size, img_size_before_pad, preds.shape[-2:], device=preds.device)
    mse, cos = masked_mse(preds, teacher, mask), masked_cosine_loss(preds, teacher, mask)
    return mse + cosine_weight * cos, mse, cos
