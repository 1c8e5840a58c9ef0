"""Development probe (GPU): host-buffer forward vs the images-per-group of the upload / g_a / g_s / download pipeline.
Each setting runs in its own process (the group size is read once)."""
import os, subprocess, sys
code = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from oracle import weights
B = int(sys.argv[1])
net = bench.seeded_model("MLICPP_L", "cuda:0").set_precision("bf16")
xh = weights.synthetic_image(B, 1088, 1920, seed=2024, kind="rand").pin_memory()
x = xh.cuda()
def timed(fn, n=4):
    keep = [fn() for _ in range(2)]; del keep
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
mp = B * 1920 * 1088 / 1e6
d, h = timed(lambda: net(x)), timed(lambda: net(xh))
print(f"B={B} group={os.environ.get('MLIC_PIPE_GROUP','default')}: device {d:.1f} ms ({mp/d*1e3:.0f} MP/s)  host {h:.1f} ms ({mp/h*1e3:.0f} MP/s)", flush=True)
'''
for B in (8, 32):
    for g in ("1", "2", "4", "8"):
        env = dict(os.environ, MLIC_PIPE_GROUP=g)
        subprocess.run([sys.executable, "-c", code, str(B)], env=env)
