import sys, time, torch
sys.path.insert(0, "/root/repo")
from types import SimpleNamespace as NS
from efficientsam3_b200.stage1.model import build_image_student_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = NS(MODEL=NS(BACKBONE="efficientvit_b1"), DATA=NS(IMG_SIZE=S), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=72))
torch.manual_seed(0)
m = build_image_student_model(cfg).cuda().eval()
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.running_var.uniform_(0.5, 1.5); mod.running_mean.normal_(0, 0.1); mod.weight.data.uniform_(0.5, 1.5)
x = torch.randn(B, 3, S, S, device="cuda")
for _ in range(3): y = m(x)
torch.cuda.synchronize()
# per-step timing
steps, marks = m.backbone.model._plan()
evs = []
t = x
for i, f in enumerate(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t = f(t); e1.record(); evs.append((type(f).__name__, tuple(t.shape), e0, e1))
torch.cuda.synchronize()
tot = 0
for n, shp, e0, e1 in evs:
    ms = e0.elapsed_time(e1); tot += ms
    print(f"{n:16s} {str(shp):28s} {ms:8.3f} ms")
print("backbone total", tot)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = m(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"B={B} S={S}: {ms:.3f} ms/step  {B/ms*1000:.1f} img/s")
