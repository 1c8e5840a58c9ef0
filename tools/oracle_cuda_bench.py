"""Development probe (GPU), context number only ("reference GPU path"): the oracle -- the reference's
algorithm as plain PyTorch ops -- run eagerly ON the B200 (stock ATen / cuDNN / cuBLAS kernels, fp32 and TF32), one
1920x1088 image, next to this repo's engine on the same weights and image.  The oracle is the checker, never the product:
this script only times it for the README / DESIGN table and compares symbols once more at full size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import mlic_oracle as mo
from oracle import weights

DEV = "cuda" if torch.cuda.is_available() else "cpu"
H, W = (1088, 1920) if DEV == "cuda" else (64, 128)
name = "MLICPP_L"


def timed(fn, n=3):
    fn()
    if DEV == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    import time
    t = time.time()
    for _ in range(n):
        fn()
    return (time.time() - t) * 1e3 / n


net = bench.seeded_model(name, DEV if DEV == "cuda" else None)
orc = mo.Oracle(name, net.state_dict())
x = weights.synthetic_image(1, H, W, seed=2024).to(DEV)
torch.set_default_device(DEV)                       # the oracle's index / mask constructors follow the default device
orc.w = {k: v.to(DEV) for k, v in orc.w.items()}
orc.table = orc.table.to(DEV)
mp = H * W / 1e6
out = {"image": f"{W}x{H}", "device": torch.cuda.get_device_name(0) if DEV == "cuda" else "cpu"}
with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ms = timed(lambda: orc.forward(x))
        out["oracle_eager_" + ("tf32" if tf32 else "fp32")] = {"ms": ms, "mp_per_s": mp / ms * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if DEV == "cuda":
        ref = orc.compress_symbols(x)
        for prec in ("fp32", "bf16"):
            net.set_precision(prec)
            ms = timed(lambda: net(x))
            c = net.compress(x)
            out["engine_" + prec] = {"ms": ms, "mp_per_s": mp / ms * 1e3,
                                     "symbols_equal_frac": float((c["symbols"] == ref["symbols"].to(c["symbols"].device)).float().mean()),
                                     "indexes_equal_frac": float((c["indexes"] == ref["indexes"].to(c["indexes"].device)).float().mean())}
print(json.dumps(out))
