"""Row-band forward of ONE 3840x2176 image (BASELINE configs[4]) over the ranks of a torchrun job (NCCL), checked against the
same image run whole on rank 0 and timed (CUDA events, max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/row_bands_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from mlic_b200.dist import EngineStages, forward_row_bands, shard_range
from oracle import weights

H, W = (int(os.environ.get("RB_H", 2176)), int(os.environ.get("RB_W", 3840)))
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
net = bench.seeded_model("MLICPP_L", dev).set_precision(os.environ.get("RB_PREC", "bf16"))
x = weights.synthetic_image(1, H, W, seed=77, kind="rand")
lo, hi = shard_range(H // 16, rank, world)
xb = x[:, :, 16 * lo:16 * hi].contiguous().to(dev)
st = EngineStages(net)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    out = forward_row_bands(st, xb, H // 16, rank, world)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    out = forward_row_bands(st, xb, H // 16, rank, world)
e1.record()
sync()
t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# whole image on one GPU for the comparison (every rank does it; rank 0 reports)
xf = x.to(dev)
for _ in range(2):
    ref = net(xf)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    ref = net(xf)
e1.record()
torch.cuda.synchronize()
whole_ms = e0.elapsed_time(e1) / 5
err = float((out["x_hat_band"] - ref["x_hat"][:, :, 16 * lo:16 * hi]).abs().max())
lerr = float((out["likelihoods"]["y_likelihoods"] - ref["likelihoods"]["y_likelihoods"]).abs().max())
e = torch.tensor([err, lerr], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"MLICPP_L forward, one {W}x{H} image, row bands over {world} GPU(s)", "precision": net.precision,
                      "banded_ms": float(t), "whole_image_one_gpu_ms": whole_ms, "speedup": whole_ms / float(t),
                      "mp_per_s": H * W / 1e6 / (float(t) * 1e-3),
                      "max_abs_err_x_hat_vs_whole": float(e[0]), "max_abs_err_y_lik_vs_whole": float(e[1])}), flush=True)
if world > 1:
    dist.destroy_process_group()
