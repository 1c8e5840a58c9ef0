"""TEST INFRASTRUCTURE ONLY -- dumps the state_dict TEMPLATE of the UNMODIFIED reference models into oracle/shapes/<name>.npz.

Run in the build container only (needs /root/reference):  python -m oracle.make_shapes
A template holds, per state_dict key, either the shape of a float parameter that oracle/weights.py regenerates from a seed, or
the value of a small buffer it leaves alone (relative_position_index, Gain, ...).  Large derived buffers (the quantised CDF
tables filled by update()) are not needed by the forward path and are left out.  With a template, bench.py's reference arm
and CPU baseline build the seeded reference weights without importing the product package.
"""
import json
import os

import numpy as np
import torch

from oracle import ref_loader, weights

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shapes")
NAMES = ["MLICPP_L", "MLICPP_M_SMALL_DEC", "MLICPP_L_VBR", "MLICPP_S"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in NAMES:
        net = ref_loader.get_reference_model(name)
        sd = net.state_dict()
        shapes = {k: tuple(v.shape) for k, v in sd.items()}
        meta, vals = {}, {}
        for k, v in sd.items():
            gen = torch.is_floating_point(v) and weights.fill_value(k, v.shape, 0, shapes) is not None
            if gen:
                meta[k] = {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")}
            elif v.numel() <= 4096:
                meta[k] = {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", ""), "stored": True}
                vals[k] = v.detach().cpu().numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **vals)
        print(name, len(meta), "entries,", len(vals), "stored buffers")


if __name__ == "__main__":
    main()
