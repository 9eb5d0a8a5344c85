import sys, math, torch
sys.path.insert(0, "/root/repo")
from efficientsam3_b200 import ops
cfgs = {"a": (32, 128, 32, 1, True, 256), "b": (64, 256, 64, 1, True, 128), "c": (16, 64, 32, 2, False, 512)}
cin, mid, cout, stride, res, S = cfgs[sys.argv[1] if len(sys.argv) > 1 else "a"]
B = 32
g = torch.Generator().manual_seed(0)
x = torch.randn(B, S, S, cin, generator=g).bfloat16().cuda()
w1 = (torch.randn(mid, cin, generator=g) / math.sqrt(cin)).bfloat16().cuda()
w3 = (torch.randn(cout, mid, generator=g) / math.sqrt(mid)).bfloat16().cuda()
f = lambda n: torch.rand(n, generator=g).cuda()
wdw = (torch.randn(9, mid, generator=g) / 3).cuda()
for _ in range(3):
    y = ops.mbconv_fused(x, w1, f(mid) + 0.5, f(mid), wdw, f(mid), w3, f(cout) + 0.5, f(cout), stride, res, "hswish")
torch.cuda.synchronize()
print(y.shape)
