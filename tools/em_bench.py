"""Development probe (GPU): the entropy model's GEMM shapes at 8 images per launch (M = 65280 / 32640 rows).
MLIC_TC_DEBUG=96 prints the per-role wait clocks of every launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops

SHAPES = [  # name, B, H, W, Cin, N, ks
    ("lrp0 512->224", 8, 68, 120, 512, 224, 1),
    ("lrp2 224->128", 8, 68, 120, 224, 128, 1),
    ("lrp4 128->32", 8, 68, 120, 128, 32, 1),
    ("ep0 960->320 (half rows)", 4, 68, 120, 960, 320, 1),
    ("ep2 320->256 (half rows)", 4, 68, 120, 320, 256, 1),
    ("ep6 128->64 (half rows)", 4, 68, 120, 128, 64, 1),
    ("qkv 160->480", 8, 68, 120, 160, 480, 1),
    ("reproj5x5 160->96", 8, 68, 120, 160, 96, 5),
    ("reproj5x5 32->64", 8, 68, 120, 32, 64, 5),
    ("mlp 96->128", 8, 68, 120, 96, 128, 1),
    ("mlp 128->64", 8, 68, 120, 128, 64, 1),
    ("q 32->32", 8, 68, 120, 32, 32, 1),
]
only = [a for a in sys.argv[1:] if not a.startswith("--")]
BM = int(os.environ.get("EM_BATCH_MULT", "1"))        # 4: the bench batch (32 images per launch)
iters = 2 if os.environ.get("MLIC_TC_DEBUG") else 20
torch.manual_seed(0)
for name, B, H, W, Cin, N, ks in SHAPES:
    if only and not any(o in name for o in only):
        continue
    B = B * BM
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5
    b = torch.randn(N) * 0.1
    out, ms = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", False, None, True, iters)
    flops = 2.0 * B * H * W * N * Cin * ks * ks
    byts = x.numel() * 2 + out.numel() * 2
    print(f"{name:28s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {byts/ms/1e6:7.1f} GB/s(alg)  in {x.numel()*2/1e6:.0f} MB out {out.numel()*2/1e6:.0f} MB", flush=True)
    sys.stderr.flush()
