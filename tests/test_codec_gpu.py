"""compress() -> byte strings -> decompress() on the CUDA engine with the host range coder in the slice loop
(models/mlicpp.py:199-378; C ABI mlic_decompress).  The decoder must rebuild exactly the y_hat / x_hat the encoder-side
walk produced, from the bytes alone, and the byte count must agree with the entropy model's own estimate."""
import math

import numpy as np
import pytest
import torch

import mlic_b200
from mlic_b200 import _lib, coder
from oracle import weights

pytestmark = pytest.mark.gpu


def _net(name, precision, y_gain):
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=y_gain, sigma_spread=6.0))
    net.update(force=True)
    return net.to("cuda").set_precision(precision)


@pytest.mark.parametrize("name,precision,B,H,W", [("MLICPP_S", "fp32", 2, 64, 128), ("MLICPP_L", "fp32", 1, 128, 192),
                                                  ("MLICPP_L", "bf16", 2, 256, 384), ("MLICPP_M_SMALL_DEC", "bf16", 1, 128, 128)])
def test_compress_decompress_round_trip(name, precision, B, H, W):
    net = _net(name, precision, 16.0)
    x = weights.synthetic_image(B, H, W, seed=21).cuda()
    c = net.compress(x, taps=("y_hat",))
    (y_strings, z_strings), shape = c["strings"], c["shape"]
    assert len(y_strings) == 1 and len(z_strings) == B and shape == (H // 64, W // 64)
    assert all(isinstance(s, bytes) and len(s) % 4 == 0 for s in y_strings + z_strings)
    # the y string decodes, on the host alone, to the symbol list the engine produced
    gc = net.gaussian_conditional
    tabs = (gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy())
    back = coder.RansDecoder().decode_with_indexes(y_strings[0], c["indexes"].cpu().numpy(), *tabs)
    assert np.array_equal(back, c["symbols"].cpu().numpy())
    d = net.decompress(c["strings"], shape, taps=("y_hat",))
    assert torch.equal(d["y_hat"], c["y_hat"])
    assert torch.equal(d["x_hat"], c["x_hat"])
    assert d["cost_time"] > 0 and net.last_launch_count > 100


@pytest.mark.parametrize("name,levels", [("MLICPP_L_VBR", (0, 3, 5)), ("MLICPP_M_SMALL_DEC_VBR", (0, 2, 4))])
def test_vbr_levels_round_trip(name, levels):
    net = _net(name, "fp32", 16.0)
    x = weights.synthetic_image(1, 64, 128, seed=22).cuda()
    sizes = []
    for s in levels:
        c = net.compress(x, stage=2, s=s, taps=("y_hat",))
        d = net.decompress(c["strings"], c["shape"], stage=2, s=s, taps=("y_hat",))
        assert torch.equal(d["y_hat"], c["y_hat"]) and torch.equal(d["x_hat"], c["x_hat"]), s
        sizes.append(len(c["strings"][0][0]))
    assert sizes[0] > 0


def test_byte_count_matches_the_tables_and_tracks_the_estimate():
    """512x768, 1.5 M symbols: the y string is as long as the quantised tables say it should be (sum of -log2 p over the
    coded symbols, escapes with their bypass nibbles), and stays near forward()'s estimate -sum log2(y_likelihoods).  The two
    differ by construction (64-level sigma grid and 16-bit tables vs continuous sigmas; 1e-9 likelihood floor vs bypass codes
    for far-out symbols), more so with these random-init weights whose sigmas are not calibrated: 25 % is the bar there."""
    net = _net("MLICPP_L", "bf16", 16.0)
    x = weights.synthetic_image(1, 512, 768, seed=23).cuda()
    est_bits = float(-torch.log2(net(x)["likelihoods"]["y_likelihoods"].double()).sum())
    c = net.compress(x)
    got_bits = 8 * len(c["strings"][0][0])
    gc = net.gaussian_conditional
    cdf, ln, off = gc._quantized_cdf.cpu().numpy().astype(np.int64), gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy()
    sym, idx = c["symbols"].cpu().numpy().astype(np.int64), c["indexes"].cpu().numpy()
    mx = ln[idx] - 2
    v = sym - off[idx]
    esc = (v < 0) | (v >= mx)
    raw = np.where(v < 0, -2 * v - 1, 2 * (v - mx))[esc]
    vv = np.where(esc, mx, v)
    ideal = float(-np.log2((cdf[idx, vv + 1] - cdf[idx, vv]) / 65536.0).sum())
    nib = np.maximum(np.ceil(np.log2(raw + 1) / 4), 0).astype(np.int64)
    ideal += float((4 * (nib + nib // 15 + 1)).sum())
    assert sym.size == 2 * 10 * 32 * 32 * 24 and est_bits > 1e5
    assert ideal - 1 <= got_bits <= ideal * 1.0005 + 128, (got_bits, ideal)
    assert abs(got_bits - est_bits) / est_bits < 0.25, (got_bits, est_bits)


def test_decompress_reports_corrupt_input():
    net = _net("MLICPP_S", "fp32", 16.0)
    x = weights.synthetic_image(1, 64, 64, seed=24).cuda()
    c = net.compress(x)
    with pytest.raises(_lib.MlicError, match="range-coder stream"):
        net.decompress([[b"abc"], c["strings"][1]], c["shape"])
    fresh = mlic_b200.get_model("MLICPP_S").to("cuda")
    with pytest.raises(RuntimeError):
        fresh.decompress(c["strings"], c["shape"])            # update() not called: no tables
