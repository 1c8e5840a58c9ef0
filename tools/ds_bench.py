"""Development probe (GPU): micro-benchmark + check of the fused DepthWiseConv (dw3x3 producer + tcgen05 1x1 GEMM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mlic_b200 import ops
SHAPES = [("ds192 @544x960 b4", 4, 544, 960, 192, 192, True, "gelu"), ("ds192 @544x960 b4 nores", 4, 544, 960, 192, 192, False, "gelu"),
          ("ds192 @544x960 b4 nores noact", 4, 544, 960, 192, 192, False, None), ("ds192 @544x960 b4 res noact", 4, 544, 960, 192, 192, True, None),
          ("ds192 @272x480 b4", 4, 272, 480, 192, 192, True, "gelu"), ("ds192 @136x240 b4", 4, 136, 240, 192, 192, True, "gelu"),
          ("ds 128->128 @68x120 b4", 4, 68, 120, 128, 128, False, None),
          ("lrp2 224->128 @68x120 b32", 32, 68, 120, 224, 128, False, "gelu"), ("lrp4 128->32 @68x120 b32", 32, 68, 120, 128, 32, False, None),
          ("cc2 192->128 @68x120 b32", 32, 68, 120, 192, 128, False, "gelu")]
only = [a for a in sys.argv[1:] if not a.startswith("--")]
torch.manual_seed(0)
for name, B, H, W, Cin, N, res, act in SHAPES:
    if only and not any(o in name for o in only):
        continue
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    dw = torch.randn(Cin, 1, 3, 3) / 3; db = torch.randn(Cin) * 0.1
    pw = (torch.randn(N, Cin, 1, 1) / Cin ** 0.5).to(torch.bfloat16).float(); pb = torch.randn(N) * 0.1
    r = torch.randn(B, H, W, N, device="cuda").to(torch.bfloat16) if res else None
    out, ms = ops.dsconv_nhwc(x, dw, db, pw, pb, 1, act, r, True, 20)
    byts = x.numel() * 2 + out.numel() * 2 * (2 if res else 1)
    msg = f"{name:28s} fused {ms*1e3:8.1f} us {byts/ms/1e6:7.0f} GB/s(alg)"
    if "--check" in sys.argv:
        out2, ms2 = ops.dsconv_nhwc(x, dw, db, pw, pb, 1, act, r, False, 3)
        y = F.conv2d(x.float().permute(0, 3, 1, 2), dw.cuda(), db.cuda(), padding=1, groups=Cin)
        y = F.conv2d(y, pw.cuda(), pb.cuda())
        if act == "gelu": y = F.gelu(y)
        y = y.permute(0, 2, 3, 1)
        if r is not None: y = y + r.float()
        msg += f" | unfused {ms2*1e3:8.1f} us | maxdiff vs torch fp32: fused {(out.float()-y).abs().max():.3e} unfused {(out2.float()-y).abs().max():.3e}"
        fl = (y.to(torch.bfloat16).float() - y)
        for nm, o in (("fused", out), ("unfused", out2)):
            er = o.float() - y
            msg += f"\n      {nm}: rms err {er.pow(2).mean().sqrt():.3e} (bf16 rounding floor {fl.pow(2).mean().sqrt():.3e}) mean err {er.mean():+.3e} (floor {fl.mean():+.3e})"
    print(msg, flush=True)
