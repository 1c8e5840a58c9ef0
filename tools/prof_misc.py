"""Development probe (GPU): one launch each of the non-GEMM heavy kernels at 4 images (for ncu --set full captures):
g_a head, final sub-pixel conv (shift-sum), local attention, and timing lines when run without ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
B = int(os.environ.get("PB", "4"))
it = int(os.environ.get("PIT", "5"))
which = sys.argv[1:] or ["head", "final", "lattn"]
if "head" in which:
    x = torch.rand(B, 3, 1088, 1920, device="cuda")
    dw = torch.randn(3, 1, 3, 3) / 3; db = torch.randn(3) * 0.1
    pw = torch.randn(192, 3, 1, 1); pb = torch.randn(192) * 0.1
    sw = torch.randn(192, 3, 1, 1); sb = torch.randn(192) * 0.1
    t, s, ms = ops.ga_head(x, dw, db, pw, pb, sw, sb, it)
    print(f"ga_head b{B}: {ms*1e3:.1f} us  {(t.numel()*4 + x.numel()*4)/max(ms,1e-9)/1e6:.0f} GB/s", flush=True)
if "final" in which:
    x = torch.randn(B, 544, 960, 192, device="cuda").to(torch.bfloat16)
    w = torch.randn(12, 192, 3, 3) / 40; b = torch.randn(12) * 0.1
    o, ms = ops.final_subpel(x, w, b, 1, it)
    print(f"final_subpel ss b{B}: {ms*1e3:.1f} us  {(x.numel()*2 + o.numel()*4)/max(ms,1e-9)/1e6:.0f} GB/s", flush=True)
if "lattn" in which:
    F = torch.randn(B * 8, 68, 120, 96, device="cuda").to(torch.bfloat16)
    rb = torch.randn(2 * 625) * 0.1
    O, ms = ops.local_attn(F, rb, 2, it)
    print(f"local_attn b{B*8}: {ms*1e3:.1f} us  {(F.numel()*2 + O.numel()*2)/max(ms,1e-9)/1e6:.0f} GB/s", flush=True)
