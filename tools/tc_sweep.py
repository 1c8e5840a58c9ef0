import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
def run(name, B, H, W, Cin, N, ks, sh, act):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5
    b = torch.randn(N) * 0.1
    out, ms = ops.conv2d_nhwc(x, w, b, 1, ks // 2, act, sh, None, True, 20)
    flops = 2.0 * B * H * W * N * Cin * ks * ks
    byts = x.numel() * 2 + out.numel() * 2
    print(f"{name:40s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s(alg)", flush=True)
for act in (None, "gelu"):
    for N in (64, 128, 192, 256):
        run(f"pw 192->{N} @544x960 act={act}", 1, 544, 960, 192, N, 1, False, act)
for K in (64, 128, 384, 768):
    run(f"pw {K}->192 @544x960 act=None", 1, 544, 960, K, 192, 1, False, None)
