"""Development probe (GPU): engine vs oracle on the golden cases, printing per-output differences."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mlic_b200
from oracle import mlic_oracle as mo, weights

def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * np.log10(1.0 / max(mse, 1e-30))

def run(name, B, H, W, modes):
    g = np.load(f"tests/golden/{name}_b{B}_{H}x{W}.npz")
    net = mlic_b200.get_model(name)
    sd = weights.seeded_state_dict(net.state_dict(), int(g["meta"][3]), y_gain=float(g["y_gain"]), sigma_spread=float(g["sigma_spread"]))
    net.load_state_dict(sd); net.update(force=True); net = net.cuda()
    x = weights.synthetic_image(B, H, W, seed=2024)
    orc = mo.Oracle(name, net.state_dict())
    vbr = "VBR" in name
    kw = dict(s=int(g["fwd_level"])) if vbr else {}
    ref = orc.forward(x, **kw)
    for prec, tc in modes:
        net.set_precision(prec); net.tensor_cores = tc
        t = time.time()
        try:
            out = net(x.cuda(), taps=("y", "y_hat"), **({"stage": 2, **kw} if vbr else {}))
            torch.cuda.synchronize()
        except Exception as e:
            print(f"{name} {prec} tc={tc}: FAILED {e}"); continue
        xh = out["x_hat"].cpu(); yl = out["likelihoods"]["y_likelihoods"].cpu(); zl = out["likelihoods"]["z_likelihoods"].cpu()
        bpp_ref = mo.rd_stats(ref, x)[0]; bpp = mo.rd_stats({"x_hat": xh, "likelihoods": {"y": yl, "z": zl}}, x)[0]
        print(f"{name} {prec} tc={tc} launches={net.last_launch_count} t={time.time()-t:.2f}s | y maxdiff {np.abs(out['y'].cpu().numpy()-g['y']).max():.3e} "
              f"| x_hat maxdiff {(xh-ref['x_hat']).abs().max():.3e} psnr(ref,ours) {psnr(xh, ref['x_hat']):.1f} dB "
              f"| ylik maxdiff {(yl-ref['likelihoods']['y_likelihoods']).abs().max():.3e} zlik {(zl-ref['likelihoods']['z_likelihoods']).abs().max():.3e} "
              f"| bpp ref {bpp_ref:.5f} ours {bpp:.5f} rel {abs(bpp-bpp_ref)/bpp_ref:.2e}")
        try:
            if vbr:
                for lv in (0, 5):
                    c = net.compress(x.cuda(), stage=2, s=lv)
                    s_, i_ = c["symbols"].cpu().numpy(), c["indexes"].cpu().numpy()
                    print(f"   compress s={lv}: sym mismatch {(s_ != g[f'symbols_s{lv}']).sum()}/{s_.size} idx mismatch {(i_ != g[f'indexes_s{lv}']).sum()}")
            else:
                c = net.compress(x.cuda())
                s_, i_ = c["symbols"].cpu().numpy(), c["indexes"].cpu().numpy()
                print(f"   compress: sym mismatch {(s_ != g['symbols']).sum()}/{s_.size} idx mismatch {(i_ != g['indexes']).sum()} zsym mismatch {(c['z_symbols'].cpu().numpy() != np.round(g['z'] - sd['entropy_bottleneck.quantiles'][:,0,1].numpy().reshape(1,-1,1,1))).sum()}")
            d = net.net_decoder_forward(x.cuda()).cpu()
            print(f"   decoder: maxdiff {np.abs(d.numpy() - g['decoder_x_hat']).max():.3e}")
        except Exception as e:
            print(f"   compress/decoder FAILED: {e}")

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), mlic_b200._lib.lib().mlic_version())
    modes = [("fp32", False), ("bf16", False), ("bf16", True)]
    run("MLICPP_S", 2, 64, 128, modes)
    run("MLICPP_L", 1, 64, 128, modes)
    run("MLICPP_M_SMALL_DEC", 1, 64, 128, modes)
    run("MLICPP_S_VBR", 1, 64, 128, modes)
