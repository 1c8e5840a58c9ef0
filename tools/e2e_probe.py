"""Development probe (GPU): host-buffer call (mlic_run_host) vs device call, pipelined or not, plus raw pinned copies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = bench.seeded_model("MLICPP_L", "cuda:0").set_precision("bf16")
xh = weights.synthetic_image(B, 1088, 1920, seed=2024, kind="rand").pin_memory()
x = xh.cuda()
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print(f"B={B} device call {timed(lambda: net(x)):.2f} ms")
for pipe in ("1", "0"):
    os.environ["MLIC_HOST_PIPE"] = pipe
    print(f"B={B} host call MLIC_HOST_PIPE={pipe}: {timed(lambda: net(xh)):.2f} ms")
d = torch.empty_like(x); ho = torch.empty(B * 36 * 1000 * 1000 // 4, dtype=torch.float32).pin_memory(); do = torch.empty(ho.shape, device="cuda")
print(f"H2D {xh.numel()*4/1e6:.0f} MB: {timed(lambda: d.copy_(xh, non_blocking=True)):.2f} ms; D2H {ho.numel()*4/1e6:.0f} MB: {timed(lambda: ho.copy_(do, non_blocking=True)):.2f} ms")
t0 = time.perf_counter(); a = torch.empty(B * 3 * 1088 * 1920, dtype=torch.float32, pin_memory=True); t1 = time.perf_counter()
del a; b_ = torch.empty(B * 3 * 1088 * 1920, dtype=torch.float32, pin_memory=True); t2 = time.perf_counter()
print(f"pinned alloc 100 MB: first {1e3*(t1-t0):.2f} ms, again {1e3*(t2-t1):.2f} ms")
