"""Times es3_gemm_bf16 on the pointwise-conv shapes of EV-M (device events) -- used for ncu captures."""
import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
shapes = [  # M, N, K, act
    (131072, 512, 128, "hswish"),
    (8388608, 64, 16, "hswish"),
    (32768, 1024, 256, "gelu"),
    (131072, 384, 128, None),
    (524288, 256, 64, "hswish"),
]
sel = [int(a) for a in sys.argv[1:]] or range(len(shapes))
for i in sel:
    M, N, K, act = shapes[i]
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    s = torch.ones(N, device="cuda"); b = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2): ops.gemm(a, w, scale=s, bias=b, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.gemm(a, w, scale=s, bias=b, act=act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (M * K + M * N) * 2 / 1e9
    print(f"M={M} N={N} K={K} act={act}: {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
