"""Deterministic re-randomisation of a state_dict (TEST INFRASTRUCTURE ONLY).

The reference ships no checkpoints and its default inits leave branches numerically dead (BN weight 0 on
residual tails, zero attention biases: SURVEY.md section 8c), so parity fixtures overwrite EVERY floating
tensor -- parameters and BN running statistics -- with values drawn from one seeded CPU generator, in
sorted-key order.  Test code and tests/golden/gen_golden.py share this function, so weights never
need to be committed.
"""
from __future__ import annotations

import math

import torch


def fill_state_dict(sd: dict, seed: int) -> dict:
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for key in sorted(sd.keys()):
        t = sd[key]
        if not torch.is_floating_point(t):
            out[key] = t.clone()
            continue
        shape = tuple(t.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "running_var":
            v = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            v = torch.randn(shape, generator=g) * 0.1
        elif t.dim() <= 1 and leaf == "weight":          # norm scale (BN / LN / LayerNorm2d)
            v = torch.rand(shape, generator=g) + 0.5
            if ".channel_mixer.m.2.bn." in key:          # RepViT residual tails (reference init: 0): keep the
                v = v * 0.25                             # 24-block residual stream O(1) instead of O(2^24)
            if ".token_mixer.0.bn." in key:              # RepVGGDW output BN sums three branches (3x the variance)
                v = v * 0.55
        elif t.dim() <= 1:                                # biases, LayerScale-like vectors
            v = torch.randn(shape, generator=g) * 0.1
        elif "positional_encoding_gaussian_matrix" in key:   # PositionEmbeddingRandom buffer: unit normal
            v = torch.randn(shape, generator=g)
        elif "pos_embed" in key or "attention_biases" in key or "embed" in leaf or "freqs" in key:
            v = torch.randn(shape, generator=g) * 0.5
        else:                                             # conv / linear weights
            fan_in = max(1, t.numel() // shape[0])
            v = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        out[key] = v.to(t.dtype)
    return out
