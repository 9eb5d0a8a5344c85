"""Bottleneck bisection of gemm_tc on the short-K shapes (ES3_GEMM_DBG switches in csrc/gemm_tc.cu): which of stores / TMEM loads /
epilogue math / operand loads bounds a tile.  Numbers are for diagnosis only (the switched-off variants compute garbage)."""
import math, os, sys, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
shapes = [(131072, 512, 256, "gelu", True), (131072, 512, 256, None, True), (131072, 512, 256, None, False), (2097152, 128, 64, "gelu", True),
          (2097152, 128, 64, None, True), (41472, 4736, 1024, "gelu", True), (41472, 4736, 1024, None, True), (131072, 512, 128, "hswish", True), (131072, 384, 128, None, False), (32768, 1024, 256, None, True), (2097152, 128, 64, None, True),
          (131072, 256, 512, None, True), (41472, 1024, 1024, None, True)]
DBG = [(0, "full"), (1, "no stores"), (4, "epilogue = hand-back only")]
for M, N, K, act, sb in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    s = torch.ones(N, device="cuda") if sb else None
    b = torch.zeros(N, device="cuda") if sb else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    print(f"--- M={M} N={N} K={K} act={act} scale/bias={sb}: {(M*K+M*N)*2/1e6:.0f} MB algorithmic")
    for hint in (0, 64, 128, 256):
        if hint > N:
            continue
        row = []
        for dbg, name in DBG:
            os.environ["ES3_GEMM_DBG"] = str(dbg)
            for _ in range(2):
                ops.gemm(a, w, scale=s, bias=b, act=act, out=out, bn_hint=hint)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gemm(a, w, scale=s, bias=b, act=act, out=out, bn_hint=hint)
            e1.record(); torch.cuda.synchronize()
            row.append(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3:.1f}")
        print(f"  bn_hint={hint}: " + " | ".join(row) + "  (us)")
os.environ["ES3_GEMM_DBG"] = "0"
