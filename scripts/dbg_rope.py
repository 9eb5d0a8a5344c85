import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
from efficientsam3_b200.model.vitdet import compute_axial_cis
H=W=8; win=4; heads=2; C=128; B=2
g = torch.Generator().manual_seed(1)
a = torch.randn(B*H*W, C, generator=g).bfloat16().cuda()
w = (torch.randn(3*C, C, generator=g)/math.sqrt(C)).bfloat16().cuda()
cis = compute_axial_cis(64, win, win)
tab = torch.view_as_real(cis).float().contiguous().cuda()
print("tab row1 first pairs", tab[1,:2], tab[1,16:18])
out = ops.gemm(a, w, rope=(tab, 2*C, H, W, win), out_dtype=torch.float32)
lin = ops.gemm(a, w, out_dtype=torch.float32)
# infer applied angle for pair 0 of head 0 (q) per row
x = torch.view_as_complex(lin[:, :2].contiguous()); y = torch.view_as_complex(out[:, :2].contiguous())
ang = torch.angle(y / x)
print("applied angle pair0 rows 0..15:", [round(v, 3) for v in ang[:16].tolist()])
x = torch.view_as_complex(lin[:, 32:34].contiguous()); y = torch.view_as_complex(out[:, 32:34].contiguous())
print("applied angle pair16 rows 0..15:", [round(v, 3) for v in torch.angle(y / x)[:16].tolist()])
hh = torch.arange(H).view(H,1).expand(H,W); ww = torch.arange(W).view(1,W).expand(H,W)
idx = ((hh%win)*win + (ww%win)).reshape(-1)
print("expected angle pair0:", [round(v,3) for v in torch.angle(cis[idx[:16], 0]).tolist()])
print("expected angle pair16:", [round(v,3) for v in torch.angle(cis[idx[:16], 16]).tolist()])
