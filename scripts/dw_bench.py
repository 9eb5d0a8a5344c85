"""Depthwise-conv kernel timings at the EV-M / RV-M training shapes (batch 32): python scripts/dw_bench.py"""
import sys
import torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops

dev = torch.device("cuda", 0)
SHAPES = [(3, 1, 256, 128), (3, 1, 128, 256), (3, 1, 64, 512), (3, 1, 32, 1024), (3, 2, 256, 128), (5, 1, 64, 384), (5, 1, 32, 768),
          (3, 1, 256, 48), (3, 1, 128, 96)]
impl = sys.argv[1] if len(sys.argv) > 1 else None
print(f"impl = {impl or 'default routing'}")
print("| ks | stride | HxW | C | us | alg GB/s |\n|---|---|---|---|---|---|")
for ks, st, hw, C in SHAPES:
    x = torch.randn(32, hw, hw, C, device=dev).to(torch.bfloat16)
    w = torch.randn(ks * ks, C, device=dev)
    b = torch.randn(C, device=dev)
    for _ in range(3):
        y = ops.dwconv(x, w, b, ks, st, None, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.dwconv(x, w, b, ks, st, None, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"| {ks} | {st} | {hw} | {C} | {us:.1f} | {(x.numel() + y.numel()) * 2 / us / 1e3:.0f} |")
