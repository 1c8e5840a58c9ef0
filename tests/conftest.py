import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# name, B, H, W  (fixtures produced by oracle/make_golden.py from the unmodified reference)
CASES = [("MLICPP_S", 2, 64, 128), ("MLICPP_L", 1, 64, 128), ("MLICPP_M_SMALL_DEC", 1, 64, 128),
         ("MLICPP_S_VBR", 1, 64, 128), ("MLICPP_L_VBR", 1, 64, 64), ("MLICPP_M_SMALL_DEC_VBR", 1, 64, 128),
         ("MLICPP_M", 1, 64, 128), ("MLICPP_S2", 1, 128, 128)]


def vbr_levels(g):
    """Gain levels a VBR fixture holds symbols for (6 for mlicpp_vbr.py:86-91, 5 for mlicpp_sd_vbr.py:95-100)."""
    return [lv for lv in range(8) if f"symbols_s{lv}" in g]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_case(name, B, H, W):
    """-> (golden npz, seeded state_dict incl. the scale table, input image) for one fixture."""
    import mlic_b200
    from oracle import weights
    g = np.load(os.path.join(GOLDEN, f"{name}_b{B}_{H}x{W}.npz"))
    net = mlic_b200.get_model(name)
    sd = weights.seeded_state_dict(net.state_dict(), int(g["meta"][3]), y_gain=float(g["y_gain"]),
                                   sigma_spread=float(g["sigma_spread"]))
    sd["gaussian_conditional.scale_table"] = mlic_b200.get_scale_table()
    x = weights.synthetic_image(B, H, W, seed=2024)
    return g, sd, x


VR_CASE = ("MLICPP_S_VBR", 1, 64, 128)          # the vr_entbttlnck=True fixture (tests/golden/MLICPP_S_VBR_VRZ_*.npz)


def vr_model():
    """MLICPlusPlusVbr(config, vr_entbttlnck=True): model_loader.get_model never passes the flag (mlicpp_vbr.py:15)."""
    from mlic_b200 import models
    return models.MLICPlusPlusVbr(models.model_config(VR_CASE[0]), name=VR_CASE[0], vr_entbttlnck=True)


def load_vr_case():
    import mlic_b200
    from oracle import weights
    name, B, H, W = VR_CASE
    g = np.load(os.path.join(GOLDEN, f"{name}_VRZ_b{B}_{H}x{W}.npz"))
    sd = weights.seeded_state_dict(vr_model().state_dict(), int(g["meta"][3]), y_gain=float(g["y_gain"]), sigma_spread=float(g["sigma_spread"]))
    sd["gaussian_conditional.scale_table"] = mlic_b200.get_scale_table()
    return g, sd, weights.synthetic_image(B, H, W, seed=2024)


def build_model(name, sd, device=None):
    import mlic_b200
    net = mlic_b200.get_model(name)
    net.load_state_dict(sd)
    net.update(force=True)
    return net.to(device) if device else net


@pytest.fixture(scope="session")
def lib_built():
    from mlic_b200 import build
    return build.build()
