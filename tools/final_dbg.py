import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from mlic_b200 import ops
for (B, H, W, C) in [(1, 67, 29, 192), (1, 67, 32, 192), (1, 64, 29, 192), (1, 67, 29, 64), (1, 80, 8, 64), (1, 16, 40, 64)]:
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, C, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(12, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(torch.bfloat16).float()
    b = torch.randn(12, generator=g) * 0.1
    ref = F.pixel_shuffle(F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), padding=1), 2)
    for impl in (0, 1):
        out, _ = ops.final_subpel(x, w, b, impl=impl)
        d = (out - ref).abs()
        bad = (d > 2e-4).nonzero()
        msg = f"shape {(B,H,W,C)} impl {impl}: max err {d.max().item():.3e} bad {bad.shape[0]}"
        if bad.shape[0]:
            hs = sorted(set((bad[:, 2] // 2).tolist())); ws_ = sorted(set((bad[:, 3] // 2).tolist()))
            msg += f" | conv rows {hs[:12]}{'...' if len(hs) > 12 else ''} cols {ws_[:12]}{'...' if len(ws_) > 12 else ''}"
        print(msg, flush=True)
