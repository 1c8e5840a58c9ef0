import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mlic_b200
from oracle import mlic_oracle as mo, weights
for name in ("MLICPP_S", "MLICPP_L"):
    B, H, W = 1, 256, 384
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234)); net.update(force=True)
    x = weights.synthetic_image(B, H, W, seed=2024)
    orc = mo.Oracle(name, net.state_dict())
    ref = orc.forward(x, trace=True)
    net = net.cuda()
    for prec, tc in (("fp32", False), ("bf16", False), ("bf16", True)):
        net.set_precision(prec); net.tensor_cores = tc
        out = net(x.cuda(), taps=("y", "y_hat"))
        npx = B * H * W
        by = lambda t: float(-torch.log2(t.double()).sum() / npx)
        yl, zl = out["likelihoods"]["y_likelihoods"].cpu(), out["likelihoods"]["z_likelihoods"].cpu()
        ry, rz = ref["likelihoods"]["y_likelihoods"], ref["likelihoods"]["z_likelihoods"]
        print(name, prec, tc, f"bpp_y {by(yl):.6f} ref {by(ry):.6f} | bpp_z {by(zl):.6f} ref {by(rz):.6f} | y maxdiff {(out['y'].cpu()-ref['trace']['y']).abs().max():.2e} |y|max {ref['trace']['y'].abs().max():.2f}"
              f" | yhat mismatch frac {((out['y_hat'].cpu()-ref['trace']['y_hat']).abs()>0.25).double().mean():.2e} | z_lik mismatch {(zl!=rz).double().mean():.2e}")
