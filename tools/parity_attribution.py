"""Development probe (GPU): where the bf16 fast mode loses symbols.  MLICPP_L at BASELINE size (1920x1088), stress weights
(y_gain 8: |y - mu| of several quantisation steps, every symbol non-trivial) and the benchmark's natural weights, against the
CPU oracle: symbol / index agreement, bpp and PSNR(x, x_hat) differences, error of y, and time per forward for every assignment
of precisions to the three stages (g_a, entropy model, g_s).  -> gpurun_out/r02_parity_attribution.json"""
import sys, os, json, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mlic_b200
from oracle import mlic_oracle as mo
from oracle import weights

H, W = 1088, 1920
name = "MLICPP_L"
out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_parity_attribution.json"


def psnr(a, b):
    return 10 * math.log10(1.0 / max(float(((a.double() - b.double()) ** 2).mean()), 1e-30))


res = []
for label, kw in (("stress y_gain=8 sigma_spread=3", dict(y_gain=8.0, sigma_spread=3.0)), ("natural (bench weights)", {})):
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, **kw))
    net.update(force=True)
    x = weights.synthetic_image(1, H, W, seed=2024, kind="rand")
    torch.set_num_threads(os.cpu_count())
    orc = mo.Oracle(name, net.state_dict())
    t0 = time.time()
    ref = orc.forward(x, trace=True)
    sref = orc.compress_symbols(x)
    t_or = time.time() - t0
    bpp_ref, _, psnr_ref = mo.rd_stats(ref, x)
    nz = float((sref["symbols"] != 0).double().mean())
    net = net.cuda()
    xc = x.cuda()
    for prec in ("bf16", ("fp32", "bf16", "bf16"), ("bf16", "fp32", "bf16"), ("fp32", "fp32", "bf16"), "fp32"):
        net.set_precision(prec)
        o = net(xc, taps=("y",))
        c = net.compress(xc)
        torch.cuda.synchronize()
        t1 = time.time()
        for _ in range(3):
            net(xc)
        torch.cuda.synchronize()
        ms = (time.time() - t1) / 3 * 1e3
        ours = {"x_hat": o["x_hat"].cpu(), "likelihoods": {"y": o["likelihoods"]["y_likelihoods"].cpu(), "z": o["likelihoods"]["z_likelihoods"].cpu()}}
        bpp, _, ps = mo.rd_stats(ours, x)
        ey = (o["y"].cpu() - ref["trace"]["y"]).abs()
        row = {"weights": label, "precision": prec if isinstance(prec, str) else "/".join(prec), "nonzero_symbol_share": nz,
               "symbols_equal": float((c["symbols"].cpu() == sref["symbols"]).double().mean()),
               "indexes_equal": float((c["indexes"].cpu() == sref["indexes"]).double().mean()),
               "z_symbols_equal": float((c["z_symbols"].cpu() == sref["z_symbols"]).double().mean()),
               "bpp_ref": bpp_ref, "bpp_rel_diff": (bpp - bpp_ref) / bpp_ref, "psnr_ref": psnr_ref, "psnr_diff_db": ps - psnr_ref,
               "psnr_xhat_vs_ref_xhat": psnr(ours["x_hat"], ref["x_hat"]), "y_abs_err_mean": float(ey.mean()), "y_abs_err_max": float(ey.max()),
               "y_abs_mean": float(ref["trace"]["y"].abs().mean()), "forward_ms": ms, "oracle_s": t_or}
        res.append(row)
        print(json.dumps(row), flush=True)
json.dump(res, open(out_path, "w"), indent=1)
