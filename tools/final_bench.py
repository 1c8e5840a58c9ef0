"""Development probe (GPU): the final 192 -> 3x2x2 subpel conv, implicit-GEMM vs shift-sum form."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.randn(B, 544, 960, 192, device="cuda").to(torch.bfloat16)
w = torch.randn(12, 192, 3, 3) / 40; b = torch.randn(12) * 0.1
for impl in (0, 1):
    out, ms = ops.final_subpel(x, w, b, impl=impl, iters=20)
    print(f"final subpel B={B} impl={impl}: {ms*1e3:8.1f} us  ({x.numel()*2/ms/1e6:.0f} GB/s of input)")
