"""Worker of tests/test_row_bands_gpu.py::test_real_nccl_ranks (one process per GPU, launched by torch.distributed.run):
forward_row_bands over real NCCL ranks against the same image run whole on every rank.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mlic_b200
from mlic_b200.dist import EngineStages, forward_row_bands, shard_range
from oracle import weights

rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for precision, H, W in (("fp32", 512, 128), ("bf16", 768, 256)):
    net = mlic_b200.get_model("MLICPP_L")
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0))
    net.update(force=True)
    net = net.to(dev).set_precision(precision)
    x = weights.synthetic_image(1, H, W, seed=21)
    lo, hi = shard_range(H // 16, rank, world)
    out = forward_row_bands(EngineStages(net), x[:, :, 16 * lo:16 * hi].contiguous().to(dev), H // 16, rank, world)
    ref = net(x.to(dev), taps=("y_hat",))
    errs = torch.tensor([float((out["x_hat_band"] - ref["x_hat"][:, :, 16 * lo:16 * hi]).abs().max()),
                         float((out["likelihoods"]["y_likelihoods"] - ref["likelihoods"]["y_likelihoods"]).abs().max()),
                         float((out["y_hat"] - ref["y_hat"]).abs().max())], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    res[precision] = [float(e) for e in errs]
if rank == 0:
    print(json.dumps({"world": world, "max_abs_err": res}), flush=True)
dist.destroy_process_group()
