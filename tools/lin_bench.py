"""Development probe (GPU): the linear global attention kernels at the inter-context shapes of the 32-image step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlic_b200 import ops
torch.manual_seed(0)
for D, heads in ((288, 9), (160, 5), (32, 1)):
    qkv = torch.randn(32, 68, 120, 3 * D, device="cuda").to(torch.bfloat16)
    out, ms = ops.lin_attn(qkv, heads, 0, 0, 10)
    print(f"inter D={D:3d} b32: {ms*1e3:7.1f} us  {(qkv.numel()*2*4/3 + out.numel()*2)/ms/1e6:6.0f} GB/s (K read twice)", flush=True)
qkv = torch.randn(32, 68, 120, 96, device="cuda").to(torch.bfloat16)
out, ms = ops.lin_attn(qkv, 2, 1, 2, 10)
print(f"intra D= 32 b32: {ms*1e3:7.1f} us")
