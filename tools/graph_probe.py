"""Development probe (GPU): fixed-shape calls replayed from a captured CUDA graph (mlic_b200.models.GraphedCall) vs launched
from the host -- MLICPP_L forward and the MLICPP_M_SMALL_DEC decoder-side walk at 1920x1088."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import weights


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


H, W = 1088, 1920
res = {}
for name, what in (("MLICPP_L", "forward"), ("MLICPP_M_SMALL_DEC", "net_decoder_forward")):
    net = bench.seeded_model(name, "cuda:0").set_precision("bf16")
    fn = net if what == "forward" else net.net_decoder_forward
    for B in [int(a) for a in sys.argv[1:]] or [1, 4]:
        x = weights.synthetic_image(B, H, W, seed=2024, kind="rand").cuda()
        plain = timed(lambda: fn(x))
        ref = fn(x)
        ref = (ref["x_hat"] if isinstance(ref, dict) else ref).clone()
        g = net.graphed(B, H, W, fn=fn)
        graph = timed(lambda: g(x))
        out = g(x)
        out = out["x_hat"] if isinstance(out, dict) else out
        mp = B * H * W / 1e6
        res[f"{name}.{what}.b{B}"] = {"host_launched_ms": plain, "graph_replay_ms": graph, "mp_per_s_host": mp / plain * 1e3,
                                      "mp_per_s_graph": mp / graph * 1e3, "identical": bool(torch.equal(out, ref)), "launches": g.launches}
        del g
    del net
print(json.dumps(res))
