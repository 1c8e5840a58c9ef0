"""Development probe (GPU): the two-SM (cta_group::2) DepthWiseConv / (I)GDN-tail kernel (ds_pair.cu) against a torch fp32
statement and against the single-SM kernels: correctness on ragged shapes, timing at the transform resolutions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mlic_b200 import ops

torch.manual_seed(0)
check = "--check" in sys.argv
only = [a for a in sys.argv[1:] if not a.startswith("--")]


def mk(B, H, W, C):
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    dw = torch.randn(C, 1, 3, 3) / 3; db = torch.randn(C) * 0.1
    pw = (torch.randn(C, C, 1, 1) / C ** 0.5).to(torch.bfloat16).float(); pb = torch.randn(C) * 0.1
    return x, dw, db, pw, pb


def ref_ds(x, dw, db, pw, pb, act, r):
    C = x.shape[-1]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), dw.cuda(), db.cuda(), padding=1, groups=C)
    y = y.to(torch.bfloat16).float()                     # the A operand is bf16
    y = F.conv2d(y, pw.cuda(), pb.cuda())
    if act == "gelu": y = F.gelu(y)
    y = y.permute(0, 2, 3, 1)
    if r is not None: y = y + r.float()
    return y


DS = [("ds192 @544x960 b4 gelu res", 4, 544, 960, 192, "gelu", True), ("ds192 @544x960 b4 gelu", 4, 544, 960, 192, "gelu", False),
      ("ds192 @272x480 b4 gelu res", 4, 272, 480, 192, "gelu", True), ("ds192 @37x53 b3 ragged", 3, 37, 53, 192, "gelu", True),
      ("ds192 @9x17 b1 none", 1, 9, 17, 192, None, False), ("ds128 @68x120 b2", 2, 68, 120, 128, "gelu", True)]
for name, B, H, W, C, act, res in DS:
    if only and not any(o in name for o in only): continue
    x, dw, db, pw, pb = mk(B, H, W, C)
    r = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16) if res else None
    out, ms = ops.dsconv_nhwc(x, dw, db, pw, pb, 1, act, r, 2, 20)
    out1, ms1 = ops.dsconv_nhwc(x, dw, db, pw, pb, 1, act, r, 1, 20)
    byts = x.numel() * 2 * (3 if res else 2)
    msg = f"{name:30s} pair {ms*1e3:8.1f} us {byts/ms/1e6:6.0f} GB/s | single-SM {ms1*1e3:8.1f} us {byts/ms1/1e6:6.0f} GB/s"
    if check:
        y = ref_ds(x, dw, db, pw, pb, act, r)
        msg += f" | maxdiff vs torch: pair {(out.float()-y).abs().max():.3e} single {(out1.float()-y).abs().max():.3e} pair-vs-single {(out.float()-out1.float()).abs().max():.3e}"
    print(msg, flush=True)

TAIL = [("tail192 gdn @544x960 b4", 4, 544, 960, 192, False, True), ("tail192 igdn @544x960 b4", 4, 544, 960, 192, True, True),
        ("tail192 gdn @272x480 b4", 4, 272, 480, 192, False, True), ("tail192 gdn @37x53 b3 ragged", 3, 37, 53, 192, False, True),
        ("tail192 igdn @9x17 b1 nores", 1, 9, 17, 192, True, False), ("tail128 gdn @68x120 b2", 2, 68, 120, 128, False, True)]
for name, B, H, W, C, inv, res in TAIL:
    if only and not any(o in name for o in only): continue
    x, dw, db, pw, pb = mk(B, H, W, C)
    gamma = (torch.rand(C, C) * 0.02 + 0.1 * torch.eye(C)).to(torch.bfloat16).float(); beta = torch.rand(C) + 0.5
    r = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16) if res else None
    out, ms = ops.ds_gdn_nhwc(x, dw, db, pw, pb, gamma, beta, inv, r, 2, 20)
    out1, ms1 = ops.ds_gdn_nhwc(x, dw, db, pw, pb, gamma, beta, inv, r, 1, 20)
    byts = x.numel() * 2 * (3 if res else 2)
    msg = f"{name:30s} pair {ms*1e3:8.1f} us {byts/ms/1e6:6.0f} GB/s | two kernels {ms1*1e3:8.1f} us"
    if check:
        v = ref_ds(x, dw, db, pw, pb, None, None)                       # fp32 v
        nrm = torch.einsum("bhwc,nc->bhwn", (v * v).to(torch.bfloat16).float(), gamma.cuda()) + beta.cuda()
        y = v * (torch.sqrt(nrm) if inv else torch.rsqrt(nrm))
        if r is not None: y = y + r.float()
        msg += f" | maxdiff vs torch: pair {(out.float()-y).abs().max():.3e} two-kernel {(out1.float()-y).abs().max():.3e} (out scale {y.abs().mean():.2f})"
    print(msg, flush=True)
