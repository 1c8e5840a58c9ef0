import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, mlic_b200
from mlic_b200 import coder
from oracle import weights
net = mlic_b200.get_model("MLICPP_L_VBR")
net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0))
net.update(force=True)
net = net.to("cuda").set_precision("fp32")
x = weights.synthetic_image(1, 64, 128, seed=22).cuda()
for s in (0, 5):
    c = net.compress(x, stage=2, s=s, taps=("y_hat",))
    d = net.decompress(c["strings"], c["shape"], stage=2, s=s, taps=("y_hat",))
    gc = net.gaussian_conditional
    tabs = (gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy())
    back = coder.RansDecoder().decode_with_indexes(c["strings"][0][0], c["indexes"].cpu().numpy(), *tabs)
    print("s", s, "host decode equal:", np.array_equal(back, c["symbols"].cpu().numpy()), "sym range", int(c["symbols"].min()), int(c["symbols"].max()))
    diff = (d["y_hat"] - c["y_hat"]).abs()
    print("  y_hat max diff", float(diff.max()), "n diff", int((diff > 0).sum()), "of", diff.numel())
    nz = (diff > 0).nonzero()
    if len(nz):
        ch = nz[:, 1]; hh = nz[:, 2]; ww = nz[:, 3]
        print("  first slices with diffs:", sorted(set((ch // 32).tolist()))[:5], "parities:", sorted(set(((hh + ww) & 1).tolist())))
        i = nz[0]; print("  example", i.tolist(), float(d["y_hat"][tuple(i)]), float(c["y_hat"][tuple(i)]))
