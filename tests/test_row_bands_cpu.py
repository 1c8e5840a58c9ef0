"""Row-band sharding of one image (mlic_b200/dist.py): band plans, and a world_size-2 gloo run of the
whole decomposition (halo exchange -> g_a band -> all-gather y -> entropy model -> g_s band) with the ORACLE as the stage
runner, against the oracle's plain forward.  The CUDA engine behind the same driver is tested in test_row_bands_gpu.py."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mlic_b200.dist import (GA_HALO_PX, GS_HALO_ROWS, band_plan, forward_row_bands, forward_row_bands_emulated,
                            shard_range)


class OracleStages:
    """Stage runner with the oracle's transforms (test infrastructure only)."""

    def __init__(self, orc):
        self.o = orc

    def analysis(self, x):
        return self.o.g_a(x)

    def entropy(self, y):
        o = self.o
        z_hat, z_lik = o.entropy_bottleneck(o.h_a(y))
        y_hat, y_lik, _, _ = o._entropy_loop(y, o.h_s(z_hat), "forward", o._gain(1, 0), None)
        return {"y_likelihoods": y_lik, "z_likelihoods": z_lik}, y_hat

    def synthesis(self, y_hat):
        return self.o.g_s(y_hat)


def _oracle(name="MLICPP_S"):
    import mlic_b200
    from oracle import mlic_oracle, weights
    net = mlic_b200.get_model(name)
    sd = weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0)
    sd["gaussian_conditional.scale_table"] = mlic_b200.get_scale_table()
    return mlic_oracle.Oracle(name, sd)


def test_band_plans_cover_the_latent_and_bound_the_halos():
    for h_lat, world in [(136, 8), (136, 4), (68, 2), (32, 3), (16, 2), (9, 1)]:
        rows = []
        for r in range(world):
            pl = band_plan(h_lat, r, world)
            rows += list(range(pl["lo"], pl["hi"]))
            assert pl["ga_top"] == (0 if r == 0 else GA_HALO_PX) and pl["ga_bot"] == (0 if r == world - 1 else GA_HALO_PX)
            assert pl["gs_top"] == min(GS_HALO_ROWS, pl["lo"]) and pl["gs_bot"] == min(GS_HALO_ROWS, h_lat - pl["hi"])
        assert rows == list(range(h_lat))
    # 3840x2176 over 8 GPUs: 17 latent rows per band
    assert [shard_range(136, r, 8) for r in (0, 7)] == [(0, 17), (119, 136)]
    with pytest.raises(ValueError):
        band_plan(12, 0, 4)                 # 3 latent rows per band < the g_a halo


@torch.no_grad()
def test_emulated_bands_equal_plain_forward_oracle():
    from oracle import weights
    orc = _oracle()
    x = weights.synthetic_image(1, 256, 128, seed=5)
    ref = orc.forward(x)
    got = forward_row_bands_emulated(OracleStages(orc), x, 3)
    assert torch.allclose(got["x_hat"], ref["x_hat"], atol=2e-5)
    for k in ("y_likelihoods", "z_likelihoods"):
        assert torch.allclose(got["likelihoods"][k], ref["likelihoods"][k], atol=2e-5)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import weights
    with torch.no_grad():
        orc = _oracle()
        x = weights.synthetic_image(1, 256, 128, seed=5)
        lo, hi = shard_range(16, rank, world)
        out = forward_row_bands(OracleStages(orc), x[:, :, 16 * lo:16 * hi].contiguous(), 16)
    q.put((rank, out["rows"], out["x_hat_band"].numpy(), out["likelihoods"]["y_likelihoods"].numpy(),
           out["likelihoods"]["z_likelihoods"].numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_row_bands_equal_plain_forward_oracle():
    from oracle import weights
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=240) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    with torch.no_grad():
        ref = _oracle().forward(weights.synthetic_image(1, 256, 128, seed=5))
    x_hat = torch.cat([torch.from_numpy(g[2]) for g in got], dim=2)
    assert [g[1] for g in got] == [(0, 8), (8, 16)]
    assert torch.allclose(x_hat, ref["x_hat"], atol=2e-5)
    for g in got:                            # the entropy model is replicated: both ranks hold the whole likelihoods
        assert torch.allclose(torch.from_numpy(g[3]), ref["likelihoods"]["y_likelihoods"], atol=2e-5)
        assert torch.allclose(torch.from_numpy(g[4]), ref["likelihoods"]["z_likelihoods"], atol=2e-5)
