"""Development probe (GPU): compress() -> strings -> decompress() of one 1920x1088 image, the host coder timed apart
(north_star: "the serial range/ANS coder stays on the host ... and is timed separately")."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mlic_b200
from mlic_b200 import _lib, coder
from oracle import weights

name = sys.argv[1] if len(sys.argv) > 1 else "MLICPP_L"
net = mlic_b200.get_model(name)
net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0))
net.update(force=True)
net = net.to("cuda").set_precision("bf16")
x = weights.synthetic_image(1, 1088, 1920, seed=31, kind="rand").cuda()


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


walk_ms, o = timed(lambda: net._run(_lib.MODE_COMPRESS, x, 1, 1088, 1920))
comp_ms, c = timed(lambda: net.compress(x))
gc = net.gaussian_conditional
tabs = (gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy())
sym, idx = c["symbols"].cpu().numpy(), c["indexes"].cpu().numpy()
enc_ms, s = timed(lambda: coder.encode_with_indexes(sym, idx, *tabs))
dec_only_ms, _ = timed(lambda: coder.RansDecoder().decode_with_indexes(s, idx, *tabs))
dec_ms, d = timed(lambda: net.decompress(c["strings"], c["shape"]))
fwd_ms, _ = timed(lambda: net(x))
print(json.dumps({"model": name, "image": "1920x1088", "symbols": int(sym.size), "nonzero_symbols": int((sym != 0).sum()),
                  "y_bytes": len(c["strings"][0][0]), "z_bytes": len(c["strings"][1][0]),
                  "bpp": 8 * (len(c["strings"][0][0]) + len(c["strings"][1][0])) / (1088 * 1920),
                  "forward_ms": fwd_ms, "compress_network_walk_ms": walk_ms, "compress_total_ms": comp_ms, "rans_encode_ms": enc_ms,
                  "rans_decode_ms": dec_only_ms, "decompress_total_ms": dec_ms,
                  "decoder_equals_encoder": bool(torch.equal(d["x_hat"], c["x_hat"]))}))
