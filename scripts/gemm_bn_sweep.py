"""Tile-width sweep of the tcgen05 GEMM on the EV-M pointwise shapes (M, K, N, act, residual): which BN does pick_bn want?"""
import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
shapes = [
    (131072, 128, 512, "hswish", False), (131072, 128, 384, None, False), (131072, 256, 128, None, True),
    (32768, 256, 1024, "hswish", False), (32768, 256, 768, None, False), (32768, 512, 256, None, True),
    (32768, 256, 1024, "gelu", False), (32768, 128, 512, "hswish", False),
]
g = torch.Generator().manual_seed(0)
for M, K, N, act, res in shapes:
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    s = torch.ones(N).cuda(); b = torch.zeros(N).cuda()
    r = torch.randn(M, N, generator=g).bfloat16().cuda() if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    line = f"M={M} K={K} N={N} act={act} res={res}:"
    for bn in (0, 64, 128, 256):
        if bn > N and bn != 0:
            continue
        f = lambda: ops.gemm(a, w, scale=s, bias=b, act=act, residual=r, out=out, bn_hint=bn)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        line += f"  bn{bn}={e0.elapsed_time(e1) / 10 * 1e3:.1f}us"
    print(line)
