"""Development probe (GPU): the CUDA-core kernels of the bf16 path alone (g_a head, local attention, stand-alone depthwise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
which = sys.argv[1:] or ["head", "attn", "dw"]
if "head" in which:
    B = 4
    x = torch.rand(B, 3, 1088, 1920, device="cuda")
    t, s, ms = ops.ga_head(x, torch.randn(3, 1, 3, 3) * 0.3, torch.randn(3) * 0.1, torch.randn(192, 3, 1, 1), torch.randn(192) * 0.1,
                           torch.randn(192, 3, 1, 1), torch.randn(192) * 0.1, 10)
    byts = x.numel() * 4 + 2 * t.numel() * 2
    print(f"ga_head {B} images: {ms*1e3:8.1f} us  {byts/ms/1e6:8.1f} GB/s", flush=True)
if "attn" in which:
    B = 8
    F = torch.randn(B, 68, 120, 96, device="cuda").to(torch.bfloat16)
    O, ms = ops.local_attn(F, torch.randn(2, 25, 25) * 0.1, 2, 10)
    print(f"local_attn_mma {B} images: {ms*1e3:8.1f} us  in {F.numel()*2/1e6:.0f} MB out {O.numel()*2/1e6:.0f} MB  {(F.numel()+O.numel())*2/ms/1e6:8.1f} GB/s", flush=True)
if "dw" in which:
    for (B, H, W, C, s) in [(1, 544, 960, 192, 1), (8, 68, 120, 512, 1)]:
        x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
        out, ms = ops.dwconv3x3_nhwc(x, torch.randn(C, 1, 3, 3) * 0.3, torch.randn(C) * 0.1, s, None, 10)
        print(f"dw {B}x{H}x{W}x{C} s{s}: {ms*1e3:8.1f} us {(x.numel()+out.numel())*2/ms/1e6:8.1f} GB/s", flush=True)
