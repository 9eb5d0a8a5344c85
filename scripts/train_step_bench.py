"""Stage-1 KD training step of the EV-M student on one GPU (config 2 of BASELINE.json without the frozen teacher, whose
embeddings the reference stage-1 loop reads from the store): train-mode forward (batch-statistics BN) -> KD loss -> native
backward -> fused AdamW.  Prints one JSON line and, with --table, the per-kernel-family table of one step.

    python scripts/train_step_bench.py [--batch 32] [--img 1024] [--embed 64] [--steps 5] [--table out.md] [--frozen-bn]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from types import SimpleNamespace as NS  # noqa: E402

from efficientsam3_b200 import ops  # noqa: E402
from efficientsam3_b200.stage1.model import build_image_student_model  # noqa: E402
from efficientsam3_b200.stage1.optim import FlatAdamW, KDLossFunction  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--img", type=int, default=1024)
    ap.add_argument("--embed", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--backbone", default="efficientvit_b1")
    ap.add_argument("--frozen-bn", action="store_true")
    ap.add_argument("--table", default=None)
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the same iteration of the CPU oracle (PyTorch fp32 eager autograd, "
                                                                "all host threads) on a bounded sample: batch 2")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, S, E = a.batch, a.img, a.embed
    torch.manual_seed(0)
    cfg = NS(MODEL=NS(BACKBONE=a.backbone), DATA=NS(IMG_SIZE=S), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=E))
    m = build_image_student_model(cfg).to(dev).train()
    if a.frozen_bn:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eval()
    opt = FlatAdamW(m, lr=1e-4, weight_decay=0.01)
    xs = [torch.randn(B, 3, S, S, device=dev) for _ in range(2)]          # 2 x 403 MB at the default size: > L2
    teacher = torch.randn(B, 1024, E, E, device=dev).half().float()
    sz = torch.tensor([[S, S * 3 // 4] if i % 2 == 0 else [S * 2 // 3, S] for i in range(B)], dtype=torch.int32, device=dev)
    state = {"i": 0, "loss": None}

    def step():
        opt.zero_grad()
        loss = KDLossFunction.apply(m(xs[state["i"] % 2]), teacher, sz, S, 1.0)
        loss.backward()
        opt.step(max_norm=5.0)
        state["i"] += 1
        state["loss"] = loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    n0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (ops.launch_count - n0) // a.steps
    # forward only (train mode), for the fwd / bwd split
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        m(xs[0])
        f0.record()
        for _ in range(3):
            m(xs[1])
        f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / 3
    prof = ops.Profiler()
    ops.set_profiler(prof)
    step()
    ops.set_profiler(None)
    agg = prof.summary()
    rows = sorted(((k, v["ms"], v["calls"], v["bytes"], v["flops"]) for k, v in agg.items()), key=lambda r: -r[1])
    line = {"workload": f"stage-1 KD training step, {a.backbone} student, batch {B} x 3x{S}x{S}, embed {E}, "
                        f"{'frozen' if a.frozen_bn else 'batch-statistics'} BN, stored-teacher targets, AdamW",
            "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1), "train_forward_ms": round(fwd_ms, 3),
            "es3_launches_per_step": int(launches), "loss": float(state["loss"].item()),
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2),
            "profiled_kernel_ms": round(sum(r[1] for r in rows), 3),
            "top": [{"kernel": k, "ms": round(t, 3), "calls": c} for k, t, c, _, _ in rows[:8]]}
    if a.cpu_baseline:
        line["cpu_baseline"] = cpu_oracle_train_step(S, E, a.backbone)
    print(json.dumps(line), flush=True)
    if a.table:
        with open(a.table, "w") as f:
            f.write(f"KD training step: {line['workload']}: {ms:.2f} ms/step ({line['images_per_s']} img/s), train-mode forward "
                    f"{fwd_ms:.2f} ms, {launches} es3 launches/step, peak memory {line['peak_mem_GB']} GB\n\n")
            f.write("| kernel family | launches/step | ms/step | alg GB/s | alg TFLOP/s |\n|---|---|---|---|---|\n")
            for k, t, c, kb, kf in rows:
                f.write(f"| `{k}` | {c} | {t:.4f} | {kb/1e9/(t/1e3):.0f} | {kf/1e12/(t/1e3):.1f} |\n")


def cpu_oracle_train_step(S, E, backbone, batch=2):
    """The same KD iteration (train-mode forward, KD loss, backward; no optimiser) on the CPU oracle -- a reported baseline only."""
    import time
    from oracle import efficientvit as O
    from oracle.kd_loss import kd_loss
    from oracle.weights import fill_state_dict
    assert backbone.startswith("efficientvit_"), "the CPU baseline leg is wired for the EfficientViT oracle"
    variant = backbone.split("_")[-1]
    cfg = NS(MODEL=NS(BACKBONE=backbone), DATA=NS(IMG_SIZE=S), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=E))
    sd0 = fill_state_dict(build_image_student_model(cfg).state_dict(), 7)
    x = torch.randn(batch, 3, S, S, generator=torch.Generator().manual_seed(1))
    teacher = torch.randn(batch, 1024, E, E, generator=torch.Generator().manual_seed(2))
    sizes = [(3, S, S)] * batch
    ts = []
    for _ in range(2):
        sd = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v.clone()) for k, v in sd0.items()}
        t0 = time.perf_counter()
        with O.bn_batch_stats():
            out = O.image_student_encoder(sd, x, E, variant)
        loss, _, _ = kd_loss(out, teacher, S, sizes, 1.0)
        loss.backward()
        ts.append(time.perf_counter() - t0)
    return {"images_per_s": round(batch / min(ts), 3), "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"best of 2 iterations of batch {batch} x 3x{S}x{S}, PyTorch-CPU fp32 eager autograd of the oracle port"}


if __name__ == "__main__":
    main()
