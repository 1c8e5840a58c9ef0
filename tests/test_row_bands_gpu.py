"""Row-band sharding on the CUDA engine (C-ABI option "stages", mlic_b200/dist.py): the banded decomposition, evaluated
rank after rank on one GPU, must reproduce the plain forward sample for sample."""
import json
import os
import subprocess
import sys

import pytest
import torch

import mlic_b200
from mlic_b200 import _lib
from mlic_b200.dist import EngineStages, forward_row_bands_emulated
from oracle import weights

pytestmark = pytest.mark.gpu


def _net(name, precision, y_gain=16.0):
    net = mlic_b200.get_model(name)
    sd = weights.seeded_state_dict(net.state_dict(), 1234, y_gain=y_gain, sigma_spread=6.0)
    net.load_state_dict(sd)
    net.update(force=True)
    return net.to("cuda").set_precision(precision)


@pytest.mark.parametrize("name,world", [("MLICPP_L", 2), ("MLICPP_L", 3), ("MLICPP_M_SMALL_DEC", 2)])
def test_fp32_bands_are_bit_exact(name, world):
    net = _net(name, "fp32")
    x = weights.synthetic_image(1, 512, 128, seed=11).cuda()           # 32 latent rows
    ref = net(x, taps=("y", "y_hat"))
    got = forward_row_bands_emulated(EngineStages(net), x, world)
    assert torch.equal(got["y"], ref["y"])
    assert torch.equal(got["y_hat"], ref["y_hat"])
    for k in ("y_likelihoods", "z_likelihoods"):
        assert torch.equal(got["likelihoods"][k], ref["likelihoods"][k])
    assert torch.equal(got["x_hat"], ref["x_hat"])


@pytest.mark.parametrize("world", [2, 4])
def test_bf16_bands_match_plain_forward(world):
    net = _net("MLICPP_L", "bf16", y_gain=1.0)
    x = weights.synthetic_image(2, 512, 256, seed=12, kind="rand").cuda()
    ref = net(x, taps=("y",))
    got = forward_row_bands_emulated(EngineStages(net), x, world)
    # same kernels, same per-sample arithmetic: only tile positions move
    assert float((got["y"] - ref["y"]).abs().max()) <= 1e-6
    assert float((got["x_hat"] - ref["x_hat"]).abs().max()) <= 1e-6
    assert float((got["likelihoods"]["y_likelihoods"] - ref["likelihoods"]["y_likelihoods"]).abs().max()) <= 1e-6


def test_stage_calls_check_their_inputs():
    net = _net("MLICPP_L", "bf16")
    with pytest.raises(ValueError):
        net.entropy_from_y(torch.zeros(1, 320, 6, 8, device="cuda"))          # 96 image rows: not a multiple of 64
    with pytest.raises(ValueError):
        net._run(_lib.MODE_FORWARD, None, 1, 64, 128, stages=4, y_hat=torch.zeros(1, 320, 5, 8, device="cuda"))
    with pytest.raises(_lib.MlicError):
        net._run(_lib.MODE_FORWARD, None, 1, 64, 128, stages=4)                # g_s alone without y_hat
    # a 48-row band is legal for g_a / g_s alone
    y = net.analysis_band(torch.rand(1, 3, 48, 128, device="cuda"))
    assert tuple(y.shape) == (1, 320, 3, 8)
    xh = net.synthesis_band(y)
    assert tuple(xh.shape) == (1, 3, 48, 128)
    # and the next plain call is whole again
    out = net(torch.rand(1, 3, 64, 128, device="cuda"))
    assert tuple(out["x_hat"].shape) == (1, 3, 64, 128)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_real_nccl_ranks():
    """forward_row_bands over REAL ranks (one process per GPU, NCCL P2P halos + all-gather of y): every rank's x_hat rows,
    the likelihoods and y_hat equal the whole-image forward -- exactly in fp32, to 1e-6 in bf16 (same kernels, same arithmetic)."""
    world = 2
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "nccl_row_bands_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "NCCL_DEBUG": "WARN"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world
    assert line["max_abs_err"]["fp32"] == [0.0, 0.0, 0.0]
    assert max(line["max_abs_err"]["bf16"]) <= 1e-6
