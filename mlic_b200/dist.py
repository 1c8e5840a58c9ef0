"""Multi-GPU plumbing of the forward path: one process per GPU, images sharded by batch (independent images, no
data-path collective), one all-reduce for the aggregate rate / distortion sums.  Works with the
`nccl` backend on GPUs and `gloo` on CPU (tests)."""
import math

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rd_sums(out, x):
    """Per-rank sums of loss/rd_loss.py:37-48: [sum log2(likelihoods), sum squared error, pixels (unpadded N*H*W)]."""
    n, _, h, w = x.shape
    ll = sum(torch.log2(v.double()).sum() for v in out["likelihoods"].values())
    se = ((out["x_hat"].double() - x.double()) ** 2).sum()
    return torch.stack([ll, se, torch.tensor(float(n * h * w), dtype=torch.float64, device=ll.device)])


def aggregate_rd(sums, group=None):
    """All-reduces the per-rank sums and returns (bpp, mse, psnr) of the whole job."""
    t = sums.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    ll, se, npx = (float(v) for v in t)
    bpp = -ll / npx
    mse = se / (3.0 * npx)
    psnr = 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")
    return bpp, mse, psnr


def max_over_ranks(value, device="cpu", group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
