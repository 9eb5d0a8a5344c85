"""Three warm-up forwards and one more of the EV-M student at the bench shape (for ncu: skip the first 3 x 62 es3 launches)."""
import sys
import torch
sys.path.insert(0, "/root/repo")
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
name = sys.argv[2] if len(sys.argv) > 2 else "efficientvit_b1"
dev = torch.device("cuda", 0)
m = bench.build_student(1024, 64, dev, name)
x = torch.randn(B, 3, 1024, 1024, device=dev)
for _ in range(4):
    y = m(x)
torch.cuda.synchronize()
print(float(y.abs().mean()))
