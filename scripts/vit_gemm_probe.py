"""The four ViT-trunk GEMMs at the teacher's shape (M = 8 x 72 x 72 tokens): python scripts/vit_gemm_probe.py [qkv|proj|fc1|fc2]"""
import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
which = sys.argv[1:] or ["qkv", "proj", "fc1", "fc2"]
M, C = 8 * 72 * 72, 1024
g = torch.Generator().manual_seed(0)
mk = lambda *s: (torch.randn(*s, generator=g) / math.sqrt(s[-1])).bfloat16().cuda()
x = torch.randn(M, C, generator=g).bfloat16().cuda()
xf = torch.randn(M, C, generator=g).cuda()
rope = torch.randn(576, 32, 2, generator=g).cuda()
BN = 0
cases = {
    "qkv": lambda: ops.gemm(x, wqkv, bias=bq, rope=(rope, 2 * C, 72, 72, 24), bn_hint=BN),
    "proj": lambda: ops.gemm(x, wproj, bias=bp, residual=xf, out_dtype=torch.float32, bn_hint=BN),
    "fc1": lambda: ops.gemm(x, wfc1, bias=b1, act="gelu", bn_hint=BN),
    "fc2": lambda: ops.gemm(h, wfc2, bias=bp, residual=xf, out_dtype=torch.float32, bn_hint=BN),
}
wqkv, bq = mk(3 * C, C), torch.zeros(3 * C).cuda()
wproj, bp = mk(C, C), torch.zeros(C).cuda()
wfc1, b1 = mk(4736, C), torch.zeros(4736).cuda()
wfc2 = mk(C, 4736)
h = torch.randn(M, 4736, generator=g).bfloat16().cuda()
fl = {"qkv": 2 * M * 3 * C * C, "proj": 2 * M * C * C, "fc1": 2 * M * 4736 * C, "fc2": 2 * M * 4736 * C}
import itertools
for n, BN in itertools.product(which, (0, 128, 64)):
    f = cases[n]
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{n} bn={BN}: {ms:.3f} ms  {fl[n] / ms / 1e9:.0f} TFLOP/s")
