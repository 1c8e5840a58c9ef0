"""Development probe (GPU): N forwards of MLICPP_L at 1920x1088 (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
from oracle import weights
net = bench.seeded_model("MLICPP_L", "cuda:0").set_precision(sys.argv[3] if len(sys.argv) > 3 else "bf16")
x = weights.synthetic_image(B, 1088, 1920, seed=2024, kind="rand").cuda()
for _ in range(n):
    out = net(x)
torch.cuda.synchronize()
print("ok", net.last_launch_count, float(out["x_hat"].mean()))
