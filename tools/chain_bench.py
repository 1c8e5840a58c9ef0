"""Development probe (GPU): the one-launch three-layer chains (chain3.cu) at the bench shapes -- EntropyParameters layers 1..3 and the
LocalContext tail on the 130 560 squeezed rows of 32 images -- against the layer-by-layer GEMMs (ops.conv2d_nhwc) they replace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops

torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 130560
for name, K1, N1, ln in (("EntropyParameters 320-256-128-64", 320, 256, False), ("LocalContext tail 800-64-128-64", 800, 64, True)):
    x = torch.randn(M, K1, device="cuda").to(torch.bfloat16)
    w1 = torch.randn(N1, K1) / K1 ** 0.5; w2 = torch.randn(128, N1) / N1 ** 0.5; w3 = torch.randn(64, 128) / 128 ** 0.5
    b1, b2, b3 = torch.randn(N1) * 0.1, torch.randn(128) * 0.1, torch.randn(64) * 0.1
    g = (torch.ones(N1), torch.zeros(N1)) if ln else None
    out, ms = ops.chain3(x, w1, b1, w2, b2, w3, b3, ln=g, iters=20)
    xs = x.view(1, 1, M, K1)
    t = 0.0
    h, m1 = ops.conv2d_nhwc(xs, w1.view(N1, K1, 1, 1), b1, 1, 0, None if ln else "gelu", False, None, 2, 20); t += m1
    h, m2 = ops.conv2d_nhwc(h, w2.view(128, N1, 1, 1), b2, 1, 0, "gelu", False, None, 2, 20); t += m2
    h, m3 = ops.conv2d_nhwc(h, w3.view(64, 128, 1, 1), b3, 1, 0, None, False, None, 2, 20); t += m3
    byts = M * (K1 * 2 + 64 * (2 if ln else 4))
    print(f"{name:36s} M={M}: chain {ms*1e3:7.1f} us ({byts/ms/1e6:6.0f} GB/s alg) | three GEMM launches {t*1e3:7.1f} us ({m1*1e3:.1f} + {m2*1e3:.1f} + {m3*1e3:.1f}; LayerNorm / residual not included)", flush=True)
