"""Development probe (GPU, torchrun): the host <-> device copy ceiling of the end-to-end bench step with N ranks on one box.
Every rank moves what one `bench.py` step moves through `mlic_run_host` -- 32 images of 1920x1088: 0.80 GB of fp32 x up, 0.80 GB of fp32
x_hat + 0.35 GB of fp32 likelihoods down, from / to pinned memory, uploads and downloads on two streams -- with no kernel in
between.  Prints the aggregate GB/s and the MP/s that copy rate alone would allow.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/hostbw.py"""
import os, time, json
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, H, W = 32, 1088, 1920
up = torch.empty(B * 3 * H * W, dtype=torch.float32).pin_memory()
down = torch.empty(B * 3 * H * W + B * (320 * 68 * 120 + 192 * 17 * 30), dtype=torch.float32).pin_memory()
d_up, d_down = torch.empty_like(up, device="cuda"), torch.empty_like(down, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s_in):
        d_up.copy_(up, non_blocking=True)
    with torch.cuda.stream(s_out):
        down.copy_(d_down, non_blocking=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
steps = 10
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
if rank == 0:
    byts = (up.numel() + down.numel()) * 4
    print(json.dumps({"ranks": world, "bytes_per_rank_step": byts, "s_per_step": dt / steps, "aggregate_GBps": world * byts * steps / dt / 1e9,
                      "copy_only_MPps": world * B * 1920 * 1088 * steps / dt / 1e6}))
if world > 1:
    dist.destroy_process_group()
