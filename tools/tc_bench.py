"""Development probe (GPU): micro-benchmark + correctness of the conv kernels on the MLICPP_L GEMM shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mlic_b200 import ops

SHAPES = [  # name, B, H, W, Cin, N, ks, shuffle
    ("pw192 @544x960", 1, 544, 960, 192, 192, 1, False),
    ("pw192 @272x480", 1, 272, 480, 192, 192, 1, False),
    ("subpel 192->768 @272x480", 1, 272, 480, 192, 768, 3, True),
    ("subpel 192->768 @136x240", 1, 136, 240, 192, 768, 3, True),
    ("subpel 320->768 @68x120", 1, 68, 120, 320, 768, 3, True),
    ("final 192->12 @544x960", 1, 544, 960, 192, 12, 3, True),
    ("EP 960->320 @68x120", 1, 68, 120, 960, 320, 1, False),
    ("EP 128->64 @68x120", 1, 68, 120, 128, 64, 1, False),
    ("reproj5x5 288->96 @68x120", 1, 68, 120, 288, 96, 5, False),
    ("fusion 800->64 @68x120", 1, 68, 120, 800, 64, 1, False),
    ("lrp 352->224 @68x120", 1, 68, 120, 352, 224, 1, False),
    ("hs subpel 480->1920 @34x60", 1, 34, 60, 480, 1920, 3, True),
    ("ragged 100->72 @13x21", 2, 13, 21, 104, 72, 3, False),
    ("pw 320->320 @68x120 b2", 2, 68, 120, 320, 320, 1, False),
    ("subpel 192->768 @272x480 b4", 4, 272, 480, 192, 768, 3, True),
    ("ragged3x3 128->512 @13x21 b3", 3, 13, 21, 128, 512, 3, True),
    ("plain3x3 64->256 @20x37 b1", 1, 20, 37, 64, 256, 3, False),
    ("reproj5x5 288->96 @68x120 b32", 32, 68, 120, 288, 96, 5, False),
    ("reproj5x5 160->96 @68x120 b32", 32, 68, 120, 160, 96, 5, False),
    ("reproj5x5 32->64 @68x120 b32", 32, 68, 120, 32, 64, 5, False),
    ("reproj5x5 96->96 ragged @13x21 b3", 3, 13, 21, 96, 96, 5, False),
    ("wide1x1 lrp0 640->224 @68x120 b32", 32, 68, 120, 640, 224, 1, False),
    ("wide1x1 lrp0 352->224 @68x120 b32", 32, 68, 120, 352, 224, 1, False),
    ("wide1x1 ep0 960->320 @68x60 b32", 32, 68, 60, 960, 320, 1, False),
    ("wide1x1 ep2 320->256 @68x60 b32", 32, 68, 60, 320, 256, 1, False),
    ("wide1x1 qkv 288->864 @68x120 b32", 32, 68, 120, 288, 864, 1, False),
    ("wide1x1 pw 320->320 @68x120 b32", 32, 68, 120, 320, 320, 1, False),
    ("wide1x1 ragged 264->200 @37x53 b48", 48, 37, 53, 264, 200, 1, False),
    ("narrow mlp0 96->128 @68x120 b32", 32, 68, 120, 96, 128, 1, False),
    ("narrow mlp4 128->64 @68x120 b32", 32, 68, 120, 128, 64, 1, False),
    ("narrow q 32->32 @68x120 b32", 32, 68, 120, 32, 32, 1, False),
    ("narrow ep4 256->128 @68x60 b32", 32, 68, 60, 256, 128, 1, False),
    ("narrow fusion 800->64 @68x60 b32", 32, 68, 60, 800, 64, 1, False),
]
check = "--check" in sys.argv
only = [a for a in sys.argv[1:] if not a.startswith("--")]
torch.manual_seed(0)
for name, B, H, W, Cin, N, ks, sh in SHAPES:
    if only and not any(o in name for o in only):
        continue
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5
    b = torch.randn(N) * 0.1
    res = None
    if not sh:
        res = torch.randn(B, H, W, N, device="cuda").to(torch.bfloat16)
    out, ms = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", sh, res, True, 20)
    flops = 2.0 * B * H * W * N * Cin * ks * ks
    byts = x.numel() * 2 + out.numel() * 2
    msg = f"{name:32s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s(alg)"
    if ks == 3 and N % 256 == 0 and Cin % 64 == 0 and (not sh or (N // 4) % 64 == 0):
        outp, msp = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", sh, None, 2, 20)
        ref1, _ = (out, ms) if res is None else ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", sh, None, 1, 2)
        msg += f" | two-SM {msp*1e3:9.1f} us {flops/msp/1e9:8.1f} TFLOP/s maxdiff vs one-SM {(outp.float()-ref1.float()).abs().max().item():.3e}"
    if ks == 1 and N >= 192 and Cin >= 256 and B * H * W >= 4 * 148 * 128:
        outp, msp = ops.conv2d_nhwc(x, w, b, 1, 0, "gelu", False, None, 2, 20)
        ref1, ms1 = ops.conv2d_nhwc(x, w, b, 1, 0, "gelu", False, None, 1, 20)
        msg += f" | two-SM {msp*1e3:9.1f} us vs one-SM {ms1*1e3:9.1f} us (no res) maxdiff {(outp.float()-ref1.float()).abs().max().item():.3e}"
    if ks == 5 and N <= 128:
        outp, msp = ops.conv2d_nhwc(x, w, b, 1, 2, None, False, None, 2, 20)
        ref1, ms1 = ops.conv2d_nhwc(x, w, b, 1, 2, None, False, None, 1, 20)
        outt, mst = ops.conv2d_nhwc(x, w, b, 1, 2, None, False, None, 3, 20)
        msg += f" | swapped {mst*1e3:9.1f} us  halo-patch {msp*1e3:9.1f} us  plain {ms1*1e3:9.1f} us (no act / res) maxdiff {(outp.float()-ref1.float()).abs().max().item():.3e} swapped vs halo {(outt.float()-outp.float()).abs().max().item():.3e} differ {float((outt != outp).float().mean()):.2e}"
        if check:
            y = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float().cuda(), b.cuda(), padding=2).permute(0, 2, 3, 1)
            msg += f" vs torch {(outp.float()-y).abs().max().item():.3e} / {(outt.float()-y).abs().max().item():.3e}"
    if check:
        ref, ms2 = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", sh, res, False, 2)
        d = (out.float() - ref.float()).abs().max().item()
        msg += f" | simt {ms2*1e3:9.1f} us maxdiff {d:.3e}"
    print(msg, flush=True)
