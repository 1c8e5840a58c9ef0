import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from mlic_b200 import ops
torch.manual_seed(0)
for (B, H, W, C, s) in [(1, 544, 960, 192, 1), (1, 272, 480, 192, 1), (1, 544, 960, 192, 2), (1, 68, 120, 640, 1), (2, 68, 120, 352, 1), (1, 37, 53, 104, 2)]:
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, 1, 3, 3) * 0.3; b = torch.randn(C) * 0.1
    out, ms = ops.dwconv3x3_nhwc(x, w, b, s, "gelu" if C == 640 else None, 20)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), stride=s, padding=1, groups=C)
    if C == 640: ref = F.gelu(ref)
    d = (out.float() - ref.permute(0, 2, 3, 1)).abs().max().item()
    byts = x.numel() * 2 + out.numel() * 2
    print(f"dw {B}x{H}x{W}x{C} s{s}: {ms*1e3:8.1f} us {byts/ms/1e6:8.1f} GB/s maxdiff {d:.3e}", flush=True)
# small-Cin pointwise
for act in (None, "gelu"):
    x = torch.rand(1, 1088, 1920, 3, device="cuda").to(torch.bfloat16)
    w = torch.randn(192, 3, 1, 1); b = torch.randn(192) * 0.1
    out, ms = ops.conv2d_nhwc(x, w, b, 2, 0, act, False, None, True, 20)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), stride=2)
    if act: ref = F.gelu(ref)
    print(f"pw 3->192 s2 act={act}: {ms*1e3:8.1f} us maxdiff {(out.float()-ref.permute(0,2,3,1)).abs().max().item():.3e}")
