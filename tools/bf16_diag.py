"""Development probe (GPU): bf16-mode deviations from the golden fixtures (PSNR delta, bpp delta, symbol agreement)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_parity_gpu import load_case, build_model, CASES, _psnr
from oracle import mlic_oracle as mo
for name, B, H, W in CASES[:3]:
    for tc in (False, True):
        g, sd, x = load_case(name, B, H, W)
        net = build_model(name, sd, "cuda").set_precision("bf16")
        net.tensor_cores = tc
        out = net(x.cuda())
        ref = {"x_hat": torch.from_numpy(g["x_hat"]), "likelihoods": {"y": torch.from_numpy(g["y_likelihoods"]), "z": torch.from_numpy(g["z_likelihoods"])}}
        ours = {"x_hat": out["x_hat"].cpu(), "likelihoods": {"y": out["likelihoods"]["y_likelihoods"].cpu(), "z": out["likelihoods"]["z_likelihoods"].cpu()}}
        bpp_ref, bpp = mo.rd_stats(ref, x)[0], mo.rd_stats(ours, x)[0]
        c = net.compress(x.cuda())
        print(f"{name:20s} tc={tc}: psnr(ours,ref) {_psnr(ours['x_hat'], ref['x_hat']):.2f} dB | dPSNR vs x {(_psnr(ours['x_hat'], x) - _psnr(ref['x_hat'], x)):+.5f} dB"
              f" | bpp rel {abs(bpp-bpp_ref)/bpp_ref:.2e} | sym {(c['symbols'].cpu().numpy() == g['symbols']).mean():.5f} idx {(c['indexes'].cpu().numpy() == g['indexes']).mean():.5f}")
