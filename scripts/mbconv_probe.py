"""Times the fused MBConv kernels on the EV-M shapes: python scripts/mbconv_probe.py [a|b|c] [tc|mma]"""
import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
cfgs = {"a": (32, 128, 32, 1, True, 256), "b": (64, 256, 64, 1, True, 128), "c": (16, 64, 32, 2, False, 512),
        "d": (32, 128, 64, 2, False, 256), "e": (64, 256, 128, 2, False, 128)}
cin, mid, cout, stride, res, S = cfgs[sys.argv[1] if len(sys.argv) > 1 else "a"]
impls = sys.argv[2:] or ["mma", "tc"]
B = 32
g = torch.Generator().manual_seed(0)
x = torch.randn(B, S, S, cin, generator=g).bfloat16().cuda()
w1 = (torch.randn(mid, cin, generator=g) / math.sqrt(cin)).bfloat16().cuda()
w3 = (torch.randn(cout, mid, generator=g) / math.sqrt(mid)).bfloat16().cuda()
f = lambda n: torch.rand(n, generator=g).cuda()
wdw = (torch.randn(9, mid, generator=g) / 3).cuda()
args = (x, w1, f(mid) + 0.5, f(mid), wdw, f(mid), w3, f(cout) + 0.5, f(cout), stride, res, "hswish")
outs = {}
for impl in impls:
    y = ops.mbconv_fused(*args, impl=impl)
    if y is None:
        print(impl, "not instantiated"); continue
    for _ in range(3):
        y = ops.mbconv_fused(*args, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = ops.mbconv_fused(*args, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nb = x.numel() * 2 + y.numel() * 2
    print(f"{impl}: {tuple(y.shape)} {ms:.4f} ms  {nb / ms / 1e6:.0f} GB/s algorithmic")
    outs[impl] = y.float()
if len(outs) == 2:
    a, b = outs["mma"], outs["tc"]
    print("tc vs mma: max abs diff %.4g (scale %.3g)" % ((a - b).abs().max().item(), a.abs().max().item()))
