"""Development probe (GPU): timings of BASELINE.json configs[2] (MLICPP_M_SMALL_DEC decode-side walk at 1920x1088) and configs[3]
(MLICPP_L_VBR compress-path symbol / CDF-index generation over the 6 rate levels)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mlic_b200 import _lib
from oracle import weights


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


H, W = 1088, 1920
mp = H * W / 1e6
out = {}
sd = bench.seeded_model("MLICPP_M_SMALL_DEC", "cuda:0").set_precision("bf16")
for B in (1, 8, 32):
    x = torch.zeros(B, 3, H, W, device="cuda")
    ms = timed(lambda: sd.net_decoder_forward(x))
    out[f"sd_decoder_walk_b{B}"] = {"ms": ms, "mp_per_s": B * mp / ms * 1e3}
    ms = timed(lambda: sd(x))
    out[f"sd_forward_b{B}"] = {"ms": ms, "mp_per_s": B * mp / ms * 1e3}
del sd
vbr = bench.seeded_model("MLICPP_L_VBR", "cuda:0").set_precision("bf16")
x = weights.synthetic_image(8, H, W, seed=3, kind="rand").cuda()
for s in range(6):
    ms = timed(lambda: vbr._run(_lib.MODE_COMPRESS, x, 8, H, W, vbr._scale(s, 0, True)))
    out[f"vbr_compress_walk_level{s}_b8"] = {"ms": ms, "mp_per_s": 8 * mp / ms * 1e3}
print(json.dumps(out))
