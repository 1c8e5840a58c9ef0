"""Development probe (GPU): one two-SM 3x3 192 -> 768 sub-pixel conv at 272 x 480 (g_s stage 5 of one 1920x1088 image) for an
ncu --set full capture (-k regex:conv3_pair)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
B = int(os.environ.get("PB", "1"))
x = torch.randn(B, 272, 480, 192, device="cuda").to(torch.bfloat16)
w = torch.randn(768, 192, 3, 3) / 41.6
b = torch.randn(768) * 0.1
out, ms = ops.conv2d_nhwc(x, w, b, 1, 1, "gelu", True, None, 2, int(os.environ.get("PIT", "3")))
print(f"conv3_pair b{B}: {ms*1e3:.1f} us {2.0*B*272*480*768*1728/max(ms,1e-9)/1e9:.0f} TFLOP/s")
