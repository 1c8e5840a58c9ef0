"""world_size-2 gloo test of the multi-GPU plumbing: batch sharding + the aggregate rate/distortion all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mlic_b200.dist import aggregate_rd, max_over_ranks, rd_sums, shard_range


def _fake_outputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 16, 32, generator=g)
    out = {"x_hat": (x + 0.05 * torch.rand(n, 3, 16, 32, generator=g)).clamp(0, 1),
           "likelihoods": {"y": torch.rand(n, 8, 1, 2, generator=g).clamp_min(1e-9), "z": torch.rand(n, 4, 1, 1, generator=g).clamp_min(1e-9)}}
    return x, out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, out = _fake_outputs(6, 7)
    lo, hi = shard_range(6, rank, world)
    part = {"x_hat": out["x_hat"][lo:hi], "likelihoods": {k: v[lo:hi] for k, v in out["likelihoods"].items()}}
    res = aggregate_rd(rd_sums(part, x[lo:hi]))
    t = max_over_ranks(10.0 + rank)
    q.put((rank, res, t))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_aggregate_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    x, out = _fake_outputs(6, 7)
    ref = aggregate_rd(rd_sums(out, x))
    for rank, res, t in got:
        assert res == pytest.approx(ref, rel=1e-12)
        assert t == 11.0
