"""LiteMLA aggregation kernel on the EV-M stage-3 / stage-4 shapes: python scripts/aggreg_probe.py [3|4]"""
import sys, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
st = sys.argv[1] if len(sys.argv) > 1 else "3"
B, H, C = (32, 64, 128) if st == "3" else (32, 32, 256)
C3 = 3 * C
g = torch.Generator().manual_seed(0)
ms = torch.randn(B, H, H, 2 * C3, generator=g).bfloat16().cuda()
wd, wp = ops.litemla_dwpw_weights(torch.randn(25, C3, generator=g).cuda() / 5, torch.randn(C3, 16, generator=g).cuda() / 4)
for _ in range(3):
    ops.litemla_aggreg_dwpw(ms, wd, wp, C3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.litemla_aggreg_dwpw(ms, wd, wp, C3)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
nb = 2 * B * H * H * C3 * 2
print(f"aggreg_dwpw stage {st}: {t:.4f} ms  {nb / t / 1e6:.0f} GB/s algorithmic ({nb / 1e6:.0f} MB)")
