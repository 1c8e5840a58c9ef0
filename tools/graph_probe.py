"""Development probe (GPU): one forward replayed from a captured CUDA graph vs launched from the host."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import weights

net = bench.seeded_model("MLICPP_L", "cuda:0").set_precision("bf16")
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    x = weights.synthetic_image(B, 1088, 1920, seed=2024, kind="rand").cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            out = net(x)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            out = net(x)
        e1.record(s)
        s.synchronize()
        plain = e0.elapsed_time(e1) / 5
        ref = {k: v.clone() for k, v in (("x_hat", out["x_hat"]), ("lik", out["likelihoods"]["y_likelihoods"]))}
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = net(x)
        for _ in range(2):
            g.replay()
        s.synchronize()
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
        graph = e0.elapsed_time(e1) / 5
    same = torch.equal(out["x_hat"], ref["x_hat"]) and torch.equal(out["likelihoods"]["y_likelihoods"], ref["lik"])
    mp = B * 1920 * 1088 / 1e6
    print(f"B={B}: host-launched {plain:.2f} ms ({mp/plain*1e3:.0f} MP/s)  graph replay {graph:.2f} ms ({mp/graph*1e3:.0f} MP/s)  identical={same}", flush=True)
    del g
