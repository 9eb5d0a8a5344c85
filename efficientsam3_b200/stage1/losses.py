"""Stage-1 KD loss and KD steps on the native path: masked MSE + masked cosine between student and teacher embeddings,
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


def kd_train_step(student, optimizer, images, teacher_embeddings, img_size_before_pad, cosine_weight: float = 1.0,
                  clip_grad: float = 5.0, lr: float | None = None, group=None, accumulation_steps: int = 1,
                  update: bool = True):
    """One iteration of the reference's train_one_epoch (stage1/train_image_encoder_stage1.py:185-230), everything on the
    device and on libes3 kernels:

        preds = model(samples)                 train-mode student forward (one autograd node, batch-statistics or frozen BN)
        loss  = masked_mse + COSINE * cosine   es3_kd_loss_fwd            (targets: stored / online teacher embeddings, fp32)
        loss_scaler(loss, optimizer, ...)      es3_kd_loss_bwd -> native student backward -> ONE all-reduce of the flat gradient
                                               arena (data parallel) -> global-norm clip + fused AdamW

    `optimizer` is a stage1.optim.FlatAdamW built on `student` (parameters and gradients live in its flat arenas, so the
    backward accumulates straight into the buffer that is all-reduced).  Gradient accumulation as in the reference
    (TRAIN.ACCUMULATION_STEPS, :212-226): the loss is divided by `accumulation_steps`, gradients add up in the arena, and only
    calls with `update=True` exchange, step and clear it.  Returns the detached (divided) loss, a device scalar -- no host sync.
    """
    from .optim import KDLossFunction
    if not student.training:
        raise RuntimeError("kd_train_step expects student.train() (freeze BN with set_bn_state-style .eval() on the BN modules)")
    sizes = torch.tensor([[int(s[1]), int(s[2])] for s in img_size_before_pad], dtype=torch.int32, device=images.device)
    preds = student(images)
    loss = KDLossFunction.apply(preds, teacher_embeddings, sizes, images.shape[-1], cosine_weight)
    if accumulation_steps != 1:
        loss = loss / accumulation_steps
    if hasattr(optimizer, "begin_backward"):
        optimizer.begin_backward(update, group)              # finished arena ranges may go to NCCL from inside the backward
    (loss * optimizer.loss_scale_tensor[0]).backward()       # GradScaler.scale(loss).backward(): the scale stays on the device
    if update:
        world = optimizer.all_reduce_grads(group)
        optimizer.step(lr=lr, max_norm=clip_grad, world_size=world)
        optimizer.zero_grad()                                # as the reference: clear right after the update (:224-225)
    return loss.detach()


def kd_train_step_online(student, teacher_model, optimizer, images, img_size_before_pad, cosine_weight: float = 1.0,
                         clip_grad: float = 5.0, lr: float | None = None, group=None, teacher_chunk: int = 8,
                         accumulation_steps: int = 1, update: bool = True):
    """The north-star wording of the stage-1 iteration: the frozen SAM3 teacher embeds the batch on the
    fly (no embedding store), the student is trained against it.  Targets are rounded through fp16 exactly as the stored
    embeddings are (save_embedding_image_stage1.py:89-92).  The teacher runs in chunks to bound its activation memory."""
    with torch.no_grad():
        t = torch.cat([teacher_model(images[i:i + teacher_chunk]) for i in range(0, images.shape[0], teacher_chunk)], 0)
        t = t.half().float()
    return kd_train_step(student, optimizer, images, t, img_size_before_pad, cosine_weight, clip_grad, lr, group,
                         accumulation_steps=accumulation_steps, update=update)
