"""Development probe (GPU): the stand-alone TMA-fed depthwise 3x3 kernel on the shapes of the 32-image step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mlic_b200 import ops
SHAPES = [("lrp0 640ch @68x120 b32", 32, 68, 120, 640, 1, None), ("qkv 864ch @68x120 b32", 32, 68, 120, 864, 1, None),
          ("mlp 128ch @68x120 b32 gelu", 32, 68, 120, 128, 1, "gelu"), ("g_a.2 192ch s2 @544x960 b4", 4, 544, 960, 192, 2, "gelu"),
          ("352ch @68x120 b32", 32, 68, 120, 352, 1, None), ("ragged 104ch @37x53 b3", 3, 37, 53, 104, 1, "gelu"), ("ragged 72ch s2 @37x53 b3", 3, 37, 53, 72, 2, None)]
only = [a for a in sys.argv[1:] if not a.startswith("--")]
torch.manual_seed(0)
for name, B, H, W, C, S, act in SHAPES:
    if only and not any(o in name for o in only):
        continue
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    dw = torch.randn(C, 1, 3, 3) / 3; db = torch.randn(C) * 0.1
    out, ms = ops.dwconv3x3_nhwc(x, dw, db, S, act, 20)
    byts = x.numel() * 2 + out.numel() * 2
    msg = f"{name:30s} {ms*1e3:8.1f} us {byts/ms/1e6:7.0f} GB/s(alg)"
    if "--check" in sys.argv:
        y = F.conv2d(x.float().permute(0, 3, 1, 2), dw.cuda(), db.cuda(), padding=1, stride=S, groups=C)
        if act == "gelu": y = F.gelu(y)
        msg += f" | maxdiff vs torch fp32 {(out.float() - y.permute(0, 2, 3, 1)).abs().max():.3e}"
    print(msg, flush=True)
