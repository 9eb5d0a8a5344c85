"""Stage-1 KD loss (forward) on the native path: masked MSE + masked cosine between student and teacher embeddings,
stage1/train_image_encoder_stage1.py:205-210, 271-307.  One streaming kernel + a fixed-order final reduction."""
from __future__ import annotations

import torch

from .. import ops


@torch.no_grad()
def kd_loss(preds: torch.Tensor, teacher: torch.Tensor, img_size: int, img_size_before_pad, cosine_weight: float = 1.0):
    """preds, teacher: [B,C,E,E] fp32 CUDA (NCHW, the modules' outputs); img_size_before_pad: sequence of (3, h, w)
    as the reference loader yields it.  Returns (loss, mse, cosine) 0-dim fp32 CUDA tensors."""
    sizes = torch.tensor([[int(s[1]), int(s[2])] for s in img_size_before_pad], dtype=torch.int32, device=preds.device)
    out, _ = ops.kd_loss_fwd(preds, teacher, sizes, img_size, cosine_weight)
    return out[0], out[1], out[2]


@torch.no_grad()
def kd_eval_step(student, teacher_model, images, img_size_before_pad, cosine_weight: float = 1.0):
    """Online teacher -> student -> loss (north-star wording of the stage-1 step, forward only).
    The teacher embedding is rounded through fp16 as the reference's stored targets are
    (save_embedding_image_stage1.py:89-92: outputs.half())."""
    t = teacher_model(images).half().float()
    s = student(images)
    return kd_loss(s, t, images.shape[-1], img_size_before_pad, cosine_weight)
