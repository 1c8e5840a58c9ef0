"""Development probe (GPU): tensor-core vs CUDA-core linear-attention context kernel on a full forward (bf16 mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import weights
net = bench.seeded_model("MLICPP_L", "cuda:0").set_precision("bf16")
x = weights.synthetic_image(2, 256, 384, seed=2024).cuda()
a = net(x); ca = net.compress(x)
os.environ["MLIC_LIN_SIMT"] = "1"
b = net(x); cb = net.compress(x)
del os.environ["MLIC_LIN_SIMT"]
d = (a["x_hat"] - b["x_hat"]).abs().max().item()
l = (a["likelihoods"]["y_likelihoods"] - b["likelihoods"]["y_likelihoods"]).abs().max().item()
s = (ca["symbols"] != cb["symbols"]).float().mean().item()
print(f"mma vs simt lin_ctx: x_hat maxdiff {d:.3e}, y_lik maxdiff {l:.3e}, symbol mismatch rate {s:.2e}")
