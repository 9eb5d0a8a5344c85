"""TEST INFRASTRUCTURE ONLY (not collected by pytest): where does the batch-statistics-BN gradient distance come from?

    python tests/study_bn_precision.py            # CPU, ~1-2 min

The GPU parity test (tests/test_zz_train_gpu.py::test_student_training_step_matches_oracle_autograd[bn_train=True]) measures an
all-gradient rel-L2 of ~0.2 between the bf16-storage training path and autograd of the fp32 oracle.  This script runs the SAME fixture
(efficientvit_b1, 4 x 320^2, embed 20, seeds 1 / 2) through the torch emulation of the training ops (tests/emu_ops.py) in four
storage variants and prints the distance of each from the fp32 oracle:

    A  activations stored in bf16 (what the kernels do)
    B  as A, but every BatchNorm's batch statistics are taken from the UNROUNDED fp32 conv output (= "statistics from the producing
       GEMM's fp32 accumulators")
    C  as B, and the normalisation itself also reads the unrounded conv output (only post-activation tensors are rounded)
    D  nothing rounded (fp32 storage): the emulation's own distance from the oracle

Result (committed in profiles/r2_bn_precision_study.txt): B == A to two digits -- the statistics average M >= 400 samples, their
rounding noise is ~2^-9 / sqrt(M); the distance is the elementwise rounding of the stored activations in a chaotic random-weight
network, which statistics from fp32 accumulators do not touch."""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace as NS

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import emu_ops  # noqa: E402
from oracle import efficientvit as O  # noqa: E402
from oracle.kd_loss import kd_loss  # noqa: E402
from oracle.weights import fill_state_dict  # noqa: E402


class _Patch:
    """monkeypatch.setattr stand-in (emu_ops.install wants one)."""

    def __init__(self):
        self.undo = []

    def setattr(self, obj, name, value):
        self.undo.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def restore(self):
        for obj, name, old in reversed(self.undo):
            setattr(obj, name, old)


def run(variant):
    from efficientsam3_b200 import ops
    from efficientsam3_b200.stage1.model import build_image_student_model
    mp = _Patch()
    emu_ops.install(mp)
    stash = {}
    if variant == "D":
        mp.setattr(emu_ops, "BF", torch.float32)
        mp.setattr(ops, "ACT_DTYPE", torch.float32)
    if variant in ("B", "C"):
        # producers of pre-BN tensors: keep the unrounded fp32 result next to the rounded one they return
        def keep(fn):
            def wrapped(*a, **kw):
                bf = emu_ops.BF
                emu_ops.BF = torch.float32
                try:
                    full = fn(*a, **kw)
                finally:
                    emu_ops.BF = bf
                if not torch.is_tensor(full) or full.dtype != torch.float32:
                    return full
                r = full.to(bf)
                stash[id(r)] = (r, full)
                return r
            return wrapped
        for name in ("gemm", "conv3x3", "dwconv", "stem_conv3x3_s2", "litemla_aggreg_dwpw"):
            if hasattr(ops, name):
                mp.setattr(ops, name, keep(getattr(ops, name)))
        stats0 = ops.bn_stats

        def bn_stats(z, *a, **kw):
            return stats0(stash[id(z)][1] if id(z) in stash else z, *a, **kw)
        mp.setattr(ops, "bn_stats", bn_stats)
        if variant == "C":
            aff0 = ops.affine_act

            def affine_act(z, *a, **kw):
                return aff0(stash[id(z)][1] if id(z) in stash else z, *a, **kw)
            mp.setattr(ops, "affine_act", affine_act)
    try:
        img, embed, B = 320, 20, 4
        cfg = NS(MODEL=NS(BACKBONE="efficientvit_b1"), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))
        m = build_image_student_model(cfg)
        m.load_state_dict(fill_state_dict(m.state_dict(), 11))
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        x = torch.randn(B, 3, img, img, generator=torch.Generator().manual_seed(1))
        teacher = torch.randn(B, 1024, embed, embed, generator=torch.Generator().manual_seed(2))
        sizes = [(3, img, img * 3 // 4) if i % 2 == 0 else (3, img * 2 // 3, img) for i in range(B)]
        torch.Tensor.is_cuda_real = None
        m.train()
        out = m(x)
        loss, _, _ = kd_loss(out.float(), teacher, img, sizes, 1.0)
        loss.backward()
        used = sum(1 for _ in stash)
        grads = {k: p.grad.double().clone() for k, p in m.named_parameters()}
        return sd0, x, teacher, sizes, out.detach().double(), grads, used
    finally:
        mp.restore()


def oracle(sd0, x, teacher, sizes, dtype):
    sd = {k: ((v.to(dtype) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point() and "running" not in k))
          for k, v in sd0.items()}
    with O.bn_batch_stats():
        out = O.image_student_encoder(sd, x.to(dtype), 20, "b1")
    loss, _, _ = kd_loss(out, teacher.to(dtype), 320, sizes, 1.0)
    loss.backward()
    return out.detach().double(), {k: v.grad.double() for k, v in sd.items() if v.requires_grad and v.grad is not None}


def dist(g, ref):
    num = sum((g[k] - ref[k]).pow(2).sum().item() for k in g)
    den = sum(ref[k].pow(2).sum().item() for k in g)
    return (num / den) ** 0.5


def main():
    torch.manual_seed(0)
    res = {}
    for v in "ABCD":
        res[v] = run(v)
        print(f"variant {v} done ({res[v][6]} stashed fp32 conv outputs)", flush=True)
    sd0, x, teacher, sizes = res["A"][:4]
    o32, g32 = oracle(sd0, x, teacher, sizes, torch.float32)
    o64, g64 = oracle(sd0, x, teacher, sizes, torch.float64)
    print(f"fp32 oracle vs fp64 oracle: output rel-L2 {((o32 - o64).norm() / o64.norm()).item():.3e}, all-gradient rel-L2 {dist(g32, g64):.3e}")
    names = dict(A="bf16 storage (the kernels)", B="A + statistics from unrounded fp32 conv outputs",
                 C="B + normalisation reads the unrounded conv output", D="fp32 storage")
    for v in "ABCD":
        out, grads = res[v][4], res[v][5]
        print(f"{v}  {names[v]:58s} output rel-L2 {((out - o32).norm() / o32.norm()).item():.3e}   all-gradient rel-L2 vs fp32 oracle "
              f"{dist(grads, g32):.3e}   vs fp64 oracle {dist(grads, g64):.3e}")


if __name__ == "__main__":
    main()
