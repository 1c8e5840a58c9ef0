"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
Weights and inputs are regenerated from seeds on the consumer side (oracle/weights.py), so
the fixtures hold outputs only: x_hat, likelihoods, the rANS-bound symbol / index lists
captured from the reference's compress(), net_decoder_forward's output, and a few
intermediate tensors (y, z) that localise a mismatch.
"""
import os

import numpy as np
import torch

from oracle import ref_loader, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, B, H, W, seed, y_gain, sigma_spread, vbr levels for compress, vbr level for forward
CASES = [
    ("MLICPP_S", 2, 64, 128, 1234, 16.0, 6.0, None, None),
    ("MLICPP_L", 1, 64, 128, 1234, 16.0, 6.0, None, None),
    ("MLICPP_M_SMALL_DEC", 1, 64, 128, 1234, 16.0, 6.0, None, None),
    ("MLICPP_S_VBR", 1, 64, 128, 1234, 16.0, 6.0, range(6), 2),
    ("MLICPP_L_VBR", 1, 64, 64, 1234, 16.0, 6.0, range(6), 1),
    ("MLICPP_M_SMALL_DEC_VBR", 1, 64, 128, 1234, 16.0, 6.0, range(5), 3),      # mlicpp_sd_vbr.py: 5 gain levels
    ("MLICPP_M", 1, 64, 128, 1234, 16.0, 6.0, None, None),                     # 8 slices of 32
    ("MLICPP_S2", 1, 128, 128, 1234, 16.0, 6.0, None, None),                   # 2 slices of 64 (head_dim 32 everywhere)
    # vr_entbttlnck=True (mlicpp_vbr.py:103-117): variable-rate hyper prior; written as <name>_VRZ_...; z symbols and steps recorded per level
    ("MLICPP_S_VBR", 1, 64, 128, 1234, 16.0, 6.0, range(6), 2, True),
]


def case_file(name, B, H, W, vr=False):
    return os.path.join(OUT, f"{name}{'_VRZ' if vr else ''}_b{B}_{H}x{W}.npz")


def main(only=None):
    """only: optional list of model names -- regenerate just those fixtures."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for case in CASES:
        name, B, H, W, seed, yg, ss, levels, fwd_level = case[:9]
        vr = len(case) > 9 and case[9]
        if only and (name + ("_VRZ" if vr else "")) not in only:
            continue
        net = ref_loader.get_reference_model(name, vr_entbttlnck=vr)
        sd = weights.seeded_state_dict(net.state_dict(), seed, y_gain=yg, sigma_spread=ss)
        net.load_state_dict(sd)
        net.update(force=True)
        x = weights.synthetic_image(B, H, W, seed=2024)
        rec = {"meta": np.array([B, H, W, seed], dtype=np.int64), "y_gain": np.float32(yg),
               "sigma_spread": np.float32(ss)}
        with torch.no_grad():
            out = net(x) if levels is None else net(x, stage=2, s=fwd_level)
            rec["x_hat"] = out["x_hat"].numpy()
            rec["y_likelihoods"] = out["likelihoods"]["y_likelihoods"].numpy()
            rec["z_likelihoods"] = out["likelihoods"]["z_likelihoods"].numpy()
            rec["y"] = net.g_a(x).numpy()
            rec["z"] = net.h_a(net.g_a(x)).numpy()
            if levels is None:
                net.compress(x)
                s_, i_ = ref_loader.recorded_symbols()
                rec["symbols"] = np.asarray(s_, dtype=np.int32)
                rec["indexes"] = np.asarray(i_, dtype=np.int32)
            else:
                rec["fwd_level"] = np.int64(fwd_level)
                for lv in levels:
                    c = net.compress(x, stage=2, s=lv)
                    s_, i_ = ref_loader.recorded_symbols()
                    rec[f"symbols_s{lv}"] = np.asarray(s_, dtype=np.int32)
                    rec[f"indexes_s{lv}"] = np.asarray(i_, dtype=np.int32)
                    if vr:              # the shim's EntropyBottleneckVbr.compress returns (z symbols, qs)
                        rec[f"z_symbols_s{lv}"] = c["strings"][1][0].numpy().astype(np.int32)
                        rec[f"z_qstep_s{lv}"] = np.float32(float(c["strings"][1][1]))
                        o = net(x, stage=2, s=lv)
                        rec[f"z_likelihoods_s{lv}"] = o["likelihoods"]["z_likelihoods"].numpy()
            rec["decoder_x_hat"] = net.net_decoder_forward(x).numpy()
        path = case_file(name, B, H, W, vr)
        np.savez_compressed(path, **rec)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB",
              {k: (v.shape if hasattr(v, 'shape') else v) for k, v in rec.items() if k.startswith(('sym', 'ind'))})


if __name__ == "__main__":
    import sys
    main(sys.argv[1:] or None)
