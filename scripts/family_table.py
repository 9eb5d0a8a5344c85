"""Per-kernel-family table (CUDA events around every C-ABI call, ops.Profiler) of one model forward:
    python scripts/family_table.py [efficientvit_b1|repvit_m1_1|tiny_vit_11m|teacher] [batch] [img] [out.md]"""
import sys
import torch
sys.path.insert(0, "/root/repo")
import bench
from efficientsam3_b200 import ops

name = sys.argv[1] if len(sys.argv) > 1 else "efficientvit_b1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
out = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda", 0)
if name == "teacher":
    from efficientsam3_b200.stage1.model import SAM3ImageTeacherEncoder
    torch.manual_seed(0)
    m = SAM3ImageTeacherEncoder(embed_size=72).to(dev)
    S = 1008
else:
    m = bench.build_student(S, S // 16, dev, name)
x = torch.randn(B, 3, S, S, device=dev)
for _ in range(3):
    m(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m(x)
e1.record()
torch.cuda.synchronize()
step = e0.elapsed_time(e1) / 5
prof = ops.Profiler()
ops.set_profiler(prof)
for _ in range(3):
    m(x)
ops.set_profiler(None)
agg = prof.summary()
rows = sorted(((k, v["ms"] / 3, v["calls"] // 3, v["bytes"] / 3, v["flops"] / 3) for k, v in agg.items()), key=lambda r: -r[1])
lines = [f"{name} B={B} S={S}: {step:.3f} ms/step un-profiled ({B / step * 1e3:.1f} img/s); sum of families {sum(r[1] for r in rows):.3f} ms", "",
         "| kernel family | launches/step | ms/step | alg GB/s | alg TFLOP/s |", "|---|---|---|---|---|"]
for k, ms, c, kb, kf in rows:
    lines.append(f"| `{k}` | {c} | {ms:.4f} | {kb / 1e9 / (ms / 1e3):.0f} | {kf / 1e12 / (ms / 1e3):.1f} |")
text = "\n".join(lines)
print(text)
if out:
    open(out, "w").write(text + "\n")
