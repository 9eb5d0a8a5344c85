"""BASELINE config 3: SAM3 ViT trunk + SAM2-branch FPN + TwoWayTransformer mask decoder, 1 point prompt, batch 8 x 1008^2."""
import sys, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
from efficientsam3_b200.model.sam1_task import Sam3PointPromptSegmenter
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
seg = Sam3PointPromptSegmenter().cuda()
x = torch.randn(B, 3, 1008, 1008, device="cuda")
coords = torch.rand(B, 1, 2, device="cuda") * 1008
labels = torch.ones(B, 1, dtype=torch.int32, device="cuda")
def step():
    return seg.set_image_batch(x).predict_batch(coords, labels, multimask_output=True)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"config3 B={B}: {ms:.2f} ms/step  {B/ms*1000:.2f} img/s; masks {tuple(out['high_res'].shape)} {out['high_res'].dtype}")
prof = ops.Profiler(); ops.set_profiler(prof); step(); ops.set_profiler(None)
agg = prof.summary(); tot = sum(v["ms"] for v in agg.values())
grp = {"trunk": 0.0, "neck": 0.0, "heads": 0.0}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:18]:
    print(f"{k:30s} calls={v['calls']:3d} {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%  {v['bytes']/1e9/(v['ms']/1e3):7.0f} GB/s {v['flops']/1e12/(v['ms']/1e3):7.1f} TF/s")
print("sum of kernels", round(tot, 2), "ms")
