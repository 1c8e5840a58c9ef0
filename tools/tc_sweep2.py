import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
def run(name, B, H, W, Cin, N, ks, sh, act):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5
    b = torch.randn(N) * 0.1
    out, ms = ops.conv2d_nhwc(x, w, b, 1, ks // 2, act, sh, None, True, 20)
    print(f"{os.environ.get('MLIC_TC_DEBUG')} {name:40s} {ms*1e3:9.1f} us", flush=True)
run("pw 192->192 @544x960", 1, 544, 960, 192, 192, 1, False, None)
run("subpel 192->768 @272x480", 1, 272, 480, 192, 768, 3, True, None)
