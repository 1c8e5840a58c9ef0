import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mlic_b200
from oracle import weights
for name, H, W in (("MLICPP_S", 128, 192), ("MLICPP_L", 256, 384), ("MLICPP_L", 1088, 1920)):
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=8.0, sigma_spread=3.0)); net.update(force=True)
    net = net.cuda().set_precision("bf16")
    x = weights.synthetic_image(1, H, W, seed=3).cuda()
    outs = {}
    for f in (False, True):
        net.fuse = f
        o = net(x, taps=("y", "y_hat"))
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(3): net(x)
        torch.cuda.synchronize()
        outs[f] = o
        print(name, H, W, "fuse", f, "launches", net.last_launch_count, f"{(time.time()-t)/3*1e3:.2f} ms")
    a, b = outs[False], outs[True]
    mse = float(((a["x_hat"] - b["x_hat"]) ** 2).mean())
    print("   y maxdiff", float((a["y"] - b["y"]).abs().max()), "x_hat maxdiff", float((a["x_hat"] - b["x_hat"]).abs().max()), "x_hat mse", mse,
          "y_hat mismatch", float(((a["y_hat"] - b["y_hat"]).abs() > 0.25).float().mean()))
