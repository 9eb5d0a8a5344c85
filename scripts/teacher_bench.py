import sys, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
from efficientsam3_b200.stage1.model import SAM3ImageTeacherEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
t = SAM3ImageTeacherEncoder(embed_size=72).cuda()
x = torch.randn(B, 3, 1008, 1008, device="cuda")
for _ in range(2): y = t(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): y = t(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"teacher B={B}: {ms:.2f} ms/step  {B/ms*1000:.2f} img/s  ({5.4*B/ms:.1f} TFLOP/s algorithmic)")
prof = ops.Profiler(); ops.set_profiler(prof); t(x); ops.set_profiler(None)
agg = prof.summary()
tot = sum(v["ms"] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:28s} calls={v['calls']:3d} {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%  {v['bytes']/1e9/(v['ms']/1e3):7.0f} GB/s {v['flops']/1e12/(v['ms']/1e3):7.1f} TF/s")
