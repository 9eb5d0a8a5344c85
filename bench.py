#!/usr/bin/env python
"""bench.py -- encoder images/sec at 1024^2 for the EV-M student (BASELINE.json metric), B200-native path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B] [--img S]

One "step" = one forward pass of the hot path over one synthetic batch (per GPU):
  ImageStudentEncoder(efficientvit_b1) : [B,3,S,S] fp32 NCHW -> [B,1024,E,E] fp32   (stage1/model.py:188-211)
N > 1 (torchrun): images are sharded across ranks, no data-path collective (inference is embarrassingly
parallel) -> "scaling": "weak"; timing = max over ranks of device time.

Keys beyond the base contract: `roofline` (dominant kernel family, measured live with CUDA events on the
launch stream), `cpu_baseline` (the oracle port timed on this box's host cores, rank 0, N=1), `e2e`
(pinned host batch -> H2D -> module forward -> D2H of the step's scalar metric, every step).
`--impl reference` times the CPU oracle port (the reference is Python/PyTorch; /root/reference does not
exist on the GPU box, and the oracle is pinned to it by tests/golden) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "encoder images/sec @1024^2 (EV-M student forward)"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_student(img, embed, device, backbone="efficientvit_b1"):
    from efficientsam3_b200.stage1.model import build_image_student_model
    cfg = NS(MODEL=NS(BACKBONE=backbone), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))
    torch.manual_seed(0)
    m = build_image_student_model(cfg)
    with torch.no_grad():  # random-init weights of the named architecture + non-trivial BN statistics
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_var.uniform_(0.5, 1.5)
                mod.running_mean.normal_(0, 0.1)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.1)
    return m.to(device).eval()


def run_native(args):
    from efficientsam3_b200 import ops
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, S, E = args.batch, args.img, args.embed
    model = build_student(S, E, dev)

    g = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(B, 3, S, S, generator=g).pin_memory() for _ in range(2)]
    x_dev = [h.to(dev) for h in host]  # 2 x 403 MB at B=32,S=1024: each larger than the 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput
    for i in range(args.warmup):
        out = model(x_dev[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = model(x_dev[i % 2])
    e1.record()
    barrier()
    launches = ops.launch_count - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ---------------------------------------------------------------- end-to-end (host buffers)
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(x_dev[0]) for _ in range(2)]
    metric_host = torch.zeros(1).pin_memory()

    def e2e_pass(nsteps):
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        with torch.cuda.stream(copy_stream):
            stage[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(nsteps):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(freed[nxt])
                    stage[nxt].copy_(host[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            torch.cuda.current_stream().wait_event(ready[cur])
            y = model(stage[cur])
            freed[cur].record()
            metric_host.copy_(y[:, :, ::8, ::8].abs().mean().reshape(1), non_blocking=True)  # step metric, 4 bytes D2H
            torch.cuda.current_stream().synchronize()  # the caller consumes the metric every step

    e2e_pass(2)
    barrier()
    t0 = time.perf_counter()
    e2e_pass(args.steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (t.item() / 1e3)

    # the same end-to-end loop with bf16 images in pinned host memory (half the PCIe bytes; the student's first kernel rounds the
    # image to bf16 operands anyway) -- reported beside the headline e2e, never instead of it
    e2e_bf16 = None
    if rank == 0 and world == 1 and not args.no_also:
        keep_host, keep_stage = host, stage
        try:
            host16 = [h.to(torch.bfloat16).pin_memory() for h in keep_host]
            host, stage = host16, [torch.empty_like(x_dev[0], dtype=torch.bfloat16) for _ in range(2)]
            e2e_pass(2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_pass(args.steps)
            torch.cuda.synchronize()
            e2e_bf16 = {"value": round(B * args.steps / (time.perf_counter() - t0), 1), "unit": UNIT,
                        "h2d_bytes_per_step": host16[0].numel() * 2, "input": "bf16 NCHW in pinned host memory"}
        except Exception as e:
            e2e_bf16 = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            host, stage = keep_host, keep_stage

    # ---------------------------------------------------------------- roofline of the dominant kernel
    roofline, kernel_table = None, None
    if rank == 0:
        prof = ops.Profiler()
        ops.set_profiler(prof)
        for i in range(3):
            model(x_dev[i % 2])
        ops.set_profiler(None)
        agg = prof.summary()
        kernel_table = sorted(((k, v["ms"] / 3, v["calls"] // 3, v["bytes"] / 3, v["flops"] / 3) for k, v in agg.items()),
                              key=lambda r: -r[1])
        name, kms, calls, kbytes, kflops = kernel_table[0]
        pk = _peaks()
        gbs = kbytes / 1e9 / (kms / 1e3)
        tfs = kflops / 1e12 / (kms / 1e3)
        frac_hbm, frac_tc = gbs / pk["hbm"], tfs / pk["tf_sustained"]
        if frac_tc > frac_hbm:
            roofline = {"bound": "tensor", "achieved": round(tfs, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                        "frac": round(frac_tc, 4), "traffic": None}
        else:
            roofline = {"bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(frac_hbm, 4), "traffic": None}
        # DRAM bytes per launch of that kernel from an `ncu --set full` capture (profiles/r1_traffic.json; null if not captured)
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get(name)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]     # of the captured launch below, not the family average
                roofline["traffic_launch"] = tr.get("captured_launch")
        roofline.update(kernel=name, launches_per_step=calls, ms_per_step=round(kms, 4),
                        share_of_step=round(kms / sum(r[1] for r in kernel_table), 4), peak_source=pk["src"])
        # whole-step figures against SURVEY section 8d's per-image algorithmic bytes / flops
        alg_bytes = 110e6 * (S / 1008.0) ** 2 * B
        alg_flops = 40.2e9 * (S / 1008.0) ** 2 * B
        roofline["step"] = {"alg_GB_per_step": round(alg_bytes / 1e9, 3), "hbm_frac": round(alg_bytes / 1e9 / (ms_step / 1e3) / pk["hbm"], 4),
                            "alg_TFLOP_per_step": round(alg_flops / 1e12, 3),
                            "tensor_frac": round(alg_flops / 1e12 / (ms_step / 1e3) / pk["tf_sustained"], 4)}

    # ---------------------------------------------------------------- the other encoders of the north-star (brief)
    also = None
    if rank == 0 and world == 1 and not args.no_also:
        also = other_encoders(dev, S, E)

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_oracle_throughput(S, E, batch=min(B, 2), steps=3, warmup=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"EV-M student encoder forward (efficientvit_b1 + 1024-ch head, eval), batch {B}/GPU x "
                                   f"3x{S}x{S} fp32 NCHW -> 1024x{E}x{E} fp32; random-init weights",
                       "batch_per_gpu": B, "global_batch": B * world, "img": S, "embed": E, "parallelism": f"dp{world}",
                       "l2_policy": f"inputs alternate between two {x_dev[0].numel()*4/1e6:.0f} MB buffers (> 126 MB L2)"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": host[0].numel() * 4,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "also": also,
        }
        if e2e_bf16 is not None and also is not None:
            also["e2e_bf16_host_input"] = e2e_bf16
        if kernel_table is not None and args.table:
            with open(args.table, "w") as f:
                f.write("| kernel family | launches/step | ms/step | alg GB/s | alg TFLOP/s |\n|---|---|---|---|---|\n")
                for k, kms, c, kb, kf in kernel_table:
                    f.write(f"| `{k}` | {c} | {kms:.4f} | {kb/1e9/(kms/1e3):.0f} | {kf/1e12/(kms/1e3):.1f} |\n")
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _time_steps(fn, warm, steps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def other_encoders(dev, S, E):
    """Device-resident images/s of the other paths the north-star names, same run, short loops (reported beside the
    headline; each has its own parity tests).  Teacher / config 3 run at their native 1008 px."""
    from efficientsam3_b200.model.sam1_task import Sam3PointPromptSegmenter
    from efficientsam3_b200.stage1.model import SAM3ImageTeacherEncoder
    out = {}
    torch.manual_seed(0)
    rv = build_student(S, E, dev, "repvit_m1_1")
    x = torch.randn(32, 3, S, S, device=dev)
    ms = _time_steps(lambda: rv(x), 2, 5)
    out["rvm_student_forward"] = {"images_per_s": round(32 / ms * 1e3, 1), "ms_per_step": round(ms, 3), "batch": 32, "img": S}
    del rv
    tv = build_student(S, E, dev, "tiny_vit_11m")
    ms = _time_steps(lambda: tv(x), 2, 5)
    out["tvm_student_forward"] = {"images_per_s": round(32 / ms * 1e3, 1), "ms_per_step": round(ms, 3), "batch": 32, "img": S}
    del tv, x
    t = SAM3ImageTeacherEncoder(embed_size=72).to(dev)
    x = torch.randn(8, 3, 1008, 1008, device=dev)
    ms = _time_steps(lambda: t(x), 1, 3)
    out["teacher_vit_forward"] = {"images_per_s": round(8 / ms * 1e3, 2), "ms_per_step": round(ms, 2), "batch": 8, "img": 1008,
                                  "alg_TFLOP_per_s": round(5.4 * 8 / ms, 1)}
    del t
    seg = Sam3PointPromptSegmenter().to(dev)
    coords = torch.rand(8, 1, 2, device=dev) * 1008
    labels = torch.ones(8, 1, dtype=torch.int32, device=dev)
    ms = _time_steps(lambda: seg.set_image_batch(x).predict_batch(coords, labels, multimask_output=True), 1, 3)
    out["config3_vit_fpn_mask_decoder"] = {"images_per_s": round(8 / ms * 1e3, 2), "ms_per_step": round(ms, 2), "batch": 8,
                                           "img": 1008, "prompt": "1 point / image, multimask"}
    del seg, x
    # optimiser side of the KD step (A20): grad norm + fused AdamW over an EV-M-sized flat arena; HBM streams
    from efficientsam3_b200.stage1.optim import FlatAdamW
    ev = build_student(S, E, dev, "efficientvit_b1")
    for p in ev.parameters():
        p.requires_grad_(True)
    opt = FlatAdamW(ev, lr=1e-3, weight_decay=0.01)
    opt.flat_grad.normal_(0, 1e-3)
    ms = _time_steps(lambda: opt.step(max_norm=5.0), 3, 20)
    gbs = opt.numel * 32 / ms / 1e6     # 4 B (norm) + 28 B (AdamW: read p,g,m,v, write p,m,v) per parameter
    out["kd_optimizer_step_evm"] = {"ms_per_step": round(ms, 4), "params": int(opt.numel), "alg_GB_per_s": round(gbs, 1),
                                    "frac_of_hbm_peak": round(gbs / _peaks()["hbm"], 3)}
    del ev, opt
    torch.cuda.empty_cache()
    # the whole stage-1 KD training iteration (config 2 minus the frozen teacher, whose embeddings the reference loop reads
    # from the store): train-mode student forward (batch-statistics BN) -> KD loss -> native backward -> fused AdamW
    try:
        out["kd_train_step_evm"] = kd_train_step_leg(dev, S, E, batch=32, steps=3, warm=2)
    except Exception as e:  # never let a side measurement take the headline line down
        out["kd_train_step_evm"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    # the same iteration for the RepViT-M1.1 student (config 4, per-GPU share: 32 of the global 256).  Its training graph is a host
    # composition of kernels that each have GPU parity tests; the whole step was first run on a GPU by this measurement.
    try:
        torch.cuda.empty_cache()
        out["kd_train_step_rvm"] = kd_train_step_leg(dev, S, E, batch=32, steps=3, warm=2, backbone="repvit_m1_1")
    except Exception as e:
        out["kd_train_step_rvm"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    # config 2 as BASELINE.json words it: EV-M student + frozen ViT teacher in the loop, batch 32, at the teacher's native 1008 px
    try:
        torch.cuda.empty_cache()
        out["config2_online_kd_step"] = online_kd_step_leg(dev, batch=32, steps=2, warm=1)
    except Exception as e:
        out["config2_online_kd_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    # EfficientSAM3 as deployed for the SAM-1 task: EV-M student encoder + SAM2-branch FPN + mask decoder, 1 point / image
    try:
        torch.cuda.empty_cache()
        out["efficientsam3_evm_point_prompt"] = efficientsam3_point_leg(dev, batch=8)
    except Exception as e:
        out["efficientsam3_evm_point_prompt"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    try:
        torch.cuda.empty_cache()
    except Exception:
        pass
    return out


def efficientsam3_point_leg(dev, batch):
    from efficientsam3_b200.model_builder import build_efficientsam3_point_segmenter
    seg = build_efficientsam3_point_segmenter("efficientvit", "b1").to(dev)
    x = torch.randn(batch, 3, 1008, 1008, device=dev)
    coords = torch.rand(batch, 1, 2, device=dev) * 1008
    labels = torch.ones(batch, 1, dtype=torch.int32, device=dev)
    ms = _time_steps(lambda: seg.set_image_batch(x).predict_batch(coords, labels, multimask_output=True), 2, 5)
    return {"images_per_s": round(batch / ms * 1e3, 1), "ms_per_step": round(ms, 2), "batch": batch, "img": 1008,
            "what": "EV-M student encoder -> 1024x72x72 -> SAM2-branch FPN -> prompt encoder + TwoWay mask decoder -> 1008^2 masks"}


def online_kd_step_leg(dev, batch, steps, warm):
    from efficientsam3_b200.stage1.losses import kd_train_step_online
    from efficientsam3_b200.stage1.model import SAM3ImageTeacherEncoder
    from efficientsam3_b200.stage1.optim import FlatAdamW
    S, E = 1008, 72
    teacher = SAM3ImageTeacherEncoder(embed_size=E).to(dev)
    m = build_student(S, E, dev).train()
    opt = FlatAdamW(m, lr=1e-4, weight_decay=0.01)
    x = torch.randn(batch, 3, S, S, device=dev)
    sizes = [(3, S, S * 3 // 4) if i % 2 == 0 else (3, S * 2 // 3, S) for i in range(batch)]
    state = {}

    def step():
        state["loss"] = kd_train_step_online(m, teacher, opt, x, sizes, 1.0, 5.0)

    ms = _time_steps(step, warm, steps)
    return {"ms_per_step": round(ms, 1), "images_per_s": round(batch / ms * 1e3, 1), "batch": batch, "img": S, "embed": E,
            "what": "frozen ViT teacher forward (chunks of 8) -> fp16-rounded targets -> EV-M train-mode forward -> KD loss -> backward -> AdamW",
            "loss": round(float(state["loss"].item()), 3), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}


def kd_train_step_leg(dev, S, E, batch, steps, warm, dist=None, backbone="efficientvit_b1"):
    """ms per stage-1 KD training iteration on this rank's GPU (data parallel when `dist` is given: one all-reduce of the flat
    gradient arena per step, stage1/optim.FlatAdamW.all_reduce_grads); max over ranks is taken by the caller."""
    from efficientsam3_b200 import ops
    from efficientsam3_b200.stage1.losses import kd_train_step
    from efficientsam3_b200.stage1.optim import FlatAdamW
    m = build_student(S, E, dev, backbone).train()
    opt = FlatAdamW(m, lr=1e-4, weight_decay=0.01)
    xs = [torch.randn(batch, 3, S, S, device=dev) for _ in range(2)]
    teacher = torch.randn(batch, 1024, E, E, device=dev).half().float()
    sizes = [(3, S, S * 3 // 4) if i % 2 == 0 else (3, S * 2 // 3, S) for i in range(batch)]
    state = {"i": 0}

    def step():
        state["loss"] = kd_train_step(m, opt, xs[state["i"] % 2], teacher, sizes, 1.0, 5.0)
        state["i"] += 1

    n0 = ops.launch_count
    ms = _time_steps(step, warm, steps)
    world = dist.get_world_size() if dist is not None else 1
    return {"student": backbone, "ms_per_step": round(ms, 2), "images_per_s_per_gpu": round(batch / ms * 1e3, 1), "batch_per_gpu": batch, "img": S,
            "bn": "batch statistics", "loss": round(float(state["loss"].item()), 3), "world": world,
            "es3_launches_per_step": (ops.launch_count - n0) // (warm + steps), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}


def cpu_oracle_throughput(S, E, batch, steps, warmup):
    """Times the CPU oracle port (oracle/efficientvit.py; pinned to the reference by tests/golden) on all host
    threads.  The only place bench.py touches oracle/ -- as a baseline, never as the product path."""
    from oracle import efficientvit as O
    from oracle.weights import fill_state_dict
    from efficientsam3_b200.stage1.model import build_image_student_model
    cfg = NS(MODEL=NS(BACKBONE="efficientvit_b1"), DATA=NS(IMG_SIZE=S), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=E))
    sd = fill_state_dict(build_image_student_model(cfg).state_dict(), 7)
    x = torch.randn(batch, 3, S, S, generator=torch.Generator().manual_seed(1))
    cores = torch.get_num_threads()
    with torch.no_grad():
        for _ in range(warmup):
            O.image_student_encoder(sd, x, E, "b1")
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            O.image_student_encoder(sd, x, E, "b1")
            ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return {"value": round(batch / sec, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} timed passes of batch {batch} x 3x{S}x{S} (median), PyTorch-CPU fp32 eager oracle port, "
                      f"torch threads={cores}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    S, E = args.img, args.embed
    batch = min(args.batch, 2)
    t0 = time.perf_counter()
    cpu = cpu_oracle_throughput(S, E, batch=batch, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(batch / cpu["value"] * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"EV-M student encoder forward on host CPU (oracle port of the reference modules), "
                               f"bounded sample: batch {batch} x 3x{S}x{S}", "img": S, "embed": E},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--img", type=int, default=1024)
    ap.add_argument("--embed", type=int, default=64, help="stage1/config.py:21,50 pair IMG_SIZE 1024 with EMBED_SIZE 64")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the brief RV-M / teacher / config-3 measurements")
    ap.add_argument("--table", default=None, help="write the per-kernel-family table (markdown) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
