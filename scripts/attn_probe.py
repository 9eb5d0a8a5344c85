"""tcgen05 flash attention on the teacher's shapes: python scripts/attn_probe.py [win|global]"""
import sys, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "win"
B, H, C, heads = 8, 72, 1024, 16
win = 24 if mode == "win" else 0
qkv = (torch.randn(B * H * H, 3 * C, generator=torch.Generator().manual_seed(0)) * 0.5).bfloat16().cuda()
for _ in range(3):
    o = ops.attention(qkv, B, H, H, C, heads, win, 0.125, impl="tc")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    o = ops.attention(qkv, B, H, H, C, heads, win, 0.125, impl="tc")
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
L = win * win if win else H * H
fl = 4 * B * H * H * L * C
print(f"attention_tc {mode}: L={L} {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s")
