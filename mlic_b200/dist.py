"""Multi-GPU plumbing of the forward path: one process per GPU.  Default: images sharded by batch (independent images, no
data-path collective), one all-reduce for the aggregate rate / distortion sums.  Large single images: row bands with a
halo exchange and an all-gather of the latent (second half of this file).  SURVEY.md 8e.  Works with the `nccl` backend
on GPUs and `gloo` on CPU (tests)."""
import math

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bind_to_gpu_cpus(device_index):
    """Pins the calling process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root), so
    that the pinned host buffers of the host-buffer path are first-touched next to the GPU that reads them.  Multi-rank jobs on
    a two-socket box otherwise put half of the ranks' staging memory across the socket link.  Returns the core list or None."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


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


# ------------------------------------------------------------------------------------------------------------------
# Row-band sharding of ONE large image (BASELINE configs[4]).
#
# The transforms are local: g_a sees 57 input pixels around a latent sample (3 x [conv3x3 s2, conv3x3, RB] + conv3x3 s2,
# modules/transform/analysis.py:6-22), g_s 6.5 latent rows (modules/transform/synthesis.py:56-73).  The entropy model is not
# (softmax over ALL positions in the linear global contexts, modules/transform/context.py:180-181,235-236), and it is the
# small part of a 4K image, so:
#   1. every rank holds a band of image rows (a multiple of 16) and receives GA_HALO_PX rows from each neighbour
#      (one grouped send/recv pair per neighbour: NCCL P2P over NVLink), runs g_a on band + halo, keeps its own y rows;
#   2. the y bands are all-gathered (20 MB at 4K); every rank runs h_a / EntropyBottleneck / h_s / the slice loop on the
#      whole latent (replicated, bit-identical on all ranks: same kernels, same inputs);
#   3. every rank runs g_s on its latent rows + GS_HALO_ROWS from its own copy of y_hat and keeps its x_hat rows.
# Zero padding at a band's artificial edge only reaches samples inside the halo, which are dropped, so the banded result
# equals the single-GPU result sample for sample (bit-exact in fp32 validation mode; tests/test_row_bands_*.py).

GA_HALO_PX = 64      # 57 rounded up to the 16x sampling phase
GS_HALO_ROWS = 7     # latent rows


def band_plan(h_lat, rank, world):
    """Latent rows [lo, hi) of `rank`, and the halo rows it needs: dict(lo, hi, ga_top, ga_bot (image rows), gs_top, gs_bot (latent rows))."""
    lo, hi = shard_range(h_lat, rank, world)
    if world > 1 and (hi - lo) * 16 < GA_HALO_PX:
        raise ValueError(f"{h_lat} latent rows over {world} ranks: a band must hold at least {GA_HALO_PX // 16} latent rows")
    return dict(lo=lo, hi=hi,
                ga_top=min(GA_HALO_PX, 16 * lo), ga_bot=min(GA_HALO_PX, 16 * (h_lat - hi)),
                gs_top=min(GS_HALO_ROWS, lo), gs_bot=min(GS_HALO_ROWS, h_lat - hi))


def exchange_row_halos(x_band, top, bot, rank, world, group=None):
    """x_band [B,C,rows,W] -> [B,C,top+rows+bot,W]: `top` rows received from rank-1, `bot` rows from rank+1 (every interior
    band edge moves GA_HALO_PX rows each way; one grouped batch of sends / receives)."""
    B, Cc, _, Ww = x_band.shape

    def peer(r):
        return dist.get_global_rank(group, r) if group is not None else r

    ops, up, down = [], None, None
    if rank > 0:
        up = x_band.new_empty(B, Cc, top, Ww)
        ops += [dist.P2POp(dist.isend, x_band[:, :, :GA_HALO_PX].contiguous(), peer(rank - 1), group),
                dist.P2POp(dist.irecv, up, peer(rank - 1), group)]
    if rank < world - 1:
        down = x_band.new_empty(B, Cc, bot, Ww)
        ops += [dist.P2POp(dist.isend, x_band[:, :, -GA_HALO_PX:].contiguous(), peer(rank + 1), group),
                dist.P2POp(dist.irecv, down, peer(rank + 1), group)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    parts = [t for t in (up, x_band, down) if t is not None]
    return torch.cat(parts, dim=2) if len(parts) > 1 else x_band


def gather_row_bands(band, h_lat, world, group=None):
    """All-gathers [B,C,rows_r,W] bands (rows may differ by one between ranks) into the whole [B,C,h_lat,W] tensor."""
    if world == 1:
        return band
    B, Cc, rows, Ww = band.shape
    mx = max(shard_range(h_lat, r, world)[1] - shard_range(h_lat, r, world)[0] for r in range(world))
    send = band if rows == mx else torch.cat([band, band.new_zeros(B, Cc, mx - rows, Ww)], dim=2)
    recv = band.new_empty(world * B, Cc, mx, Ww)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view(world, B, Cc, mx, Ww)
    return torch.cat([recv[r, :, :, :shard_range(h_lat, r, world)[1] - shard_range(h_lat, r, world)[0]] for r in range(world)], dim=2)


class EngineStages:
    """The three stage calls of a `mlic_b200` model (libmlic_b200.so, option "stages")."""

    def __init__(self, net):
        self.net = net

    def analysis(self, x):
        return self.net.analysis_band(x)

    def entropy(self, y):
        return self.net.entropy_from_y(y)

    def synthesis(self, y_hat):
        return self.net.synthesis_band(y_hat)


def forward_row_bands(stages, x_band, h_lat, rank=None, world=None, group=None):
    """forward() of one image (batch) whose rows are sharded over the ranks of `group`.

    x_band: this rank's image rows [B,3,16*(hi-lo),W] (band_plan).  Returns {"x_hat_band", "likelihoods", "y_hat", "rows": (lo, hi)};
    the likelihoods and y_hat are whole and identical on every rank, x_hat_band holds image rows [16*lo, 16*hi)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    pl = band_plan(h_lat, rank, world)
    rows = pl["hi"] - pl["lo"]
    if x_band.shape[2] != 16 * rows:
        raise ValueError(f"rank {rank} expects {16 * rows} image rows, got {x_band.shape[2]}")
    x_ext = exchange_row_halos(x_band, pl["ga_top"], pl["ga_bot"], rank, world, group) if world > 1 else x_band
    y_ext = stages.analysis(x_ext)
    y_band = y_ext[:, :, pl["ga_top"] // 16: pl["ga_top"] // 16 + rows]
    y = gather_row_bands(y_band.contiguous(), h_lat, world, group)
    lik, y_hat = stages.entropy(y)
    x_hat_ext = stages.synthesis(y_hat[:, :, pl["lo"] - pl["gs_top"]: pl["hi"] + pl["gs_bot"]].contiguous())
    x_hat_band = x_hat_ext[:, :, 16 * pl["gs_top"]: 16 * (pl["gs_top"] + rows)]
    return {"x_hat_band": x_hat_band, "likelihoods": lik, "y_hat": y_hat, "rows": (pl["lo"], pl["hi"])}


def forward_row_bands_emulated(stages, x, world):
    """The same decomposition evaluated rank after rank in ONE process (no collectives): what `world` ranks would compute.
    Used by the single-GPU parity test and to check a band plan before a multi-GPU run."""
    h_lat = x.shape[2] // 16
    ys = []
    for r in range(world):
        pl = band_plan(h_lat, r, world)
        x_ext = x[:, :, 16 * pl["lo"] - pl["ga_top"]: 16 * pl["hi"] + pl["ga_bot"]].contiguous()
        y_ext = stages.analysis(x_ext)
        ys.append(y_ext[:, :, pl["ga_top"] // 16: pl["ga_top"] // 16 + pl["hi"] - pl["lo"]])
    y = torch.cat(ys, dim=2)
    lik, y_hat = stages.entropy(y)
    xs = []
    for r in range(world):
        pl = band_plan(h_lat, r, world)
        xe = stages.synthesis(y_hat[:, :, pl["lo"] - pl["gs_top"]: pl["hi"] + pl["gs_bot"]].contiguous())
        xs.append(xe[:, :, 16 * pl["gs_top"]: 16 * (pl["gs_top"] + pl["hi"] - pl["lo"])])
    return {"x_hat": torch.cat(xs, dim=2), "likelihoods": lik, "y_hat": y_hat, "y": y}
