"""Kernel-level parity: the convolution kernels (CUDA-core and tcgen05) and the fused GaussianConditional kernel
against plain PyTorch fp32 references of the same op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mlic_b200 import ops
from oracle import mlic_oracle as mo

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False           # the torch reference must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False

CONV_SHAPES = [  # B, H, W, Cin, N, ks, shuffle
    (1, 24, 40, 192, 192, 1, False), (2, 17, 30, 320, 128, 1, False), (1, 16, 24, 192, 768, 3, True),
    (1, 20, 28, 288, 96, 5, False), (2, 13, 21, 104, 72, 3, False), (1, 16, 16, 64, 12, 3, True),
    (1, 9, 33, 800, 64, 1, False), (1, 34, 60, 480, 1920, 3, True),
]


def _torch_conv(x, w, b, ks, act, shuffle, res):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), padding=ks // 2)
    if act == "gelu":
        y = F.gelu(y)
    if shuffle:
        y = F.pixel_shuffle(y, 2)
    y = y.permute(0, 2, 3, 1)
    return y + res.float() if res is not None else y


@pytest.mark.parametrize("B,H,W,Cin,N,ks,shuffle", CONV_SHAPES)
def test_conv_fp32_cuda_core_kernel(B, H, W, Cin, N, ks, shuffle):
    torch.manual_seed(1)
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5
    b = torch.randn(N) * 0.1
    res = None if shuffle else torch.randn(B, H, W, N, device="cuda")
    out, _ = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", shuffle, res, tensor_cores=False)
    ref = _torch_conv(x, w, b, ks, "gelu", shuffle, res)
    torch.testing.assert_close(out, ref, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("B,H,W,Cin,N,ks,shuffle", CONV_SHAPES)
def test_conv_bf16_tcgen05_kernel(B, H, W, Cin, N, ks, shuffle):
    """bf16 operands, fp32 accumulation: against torch fp32 on the SAME bf16-rounded operands the only differences are
    summation order and the final bf16 rounding of the output (<= 2^-8 relative) plus the approximate-erf GELU."""
    torch.manual_seed(2)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Cin, ks, ks) / (Cin * ks * ks) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(N) * 0.1
    res = None if shuffle else torch.randn(B, H, W, N, device="cuda").to(torch.bfloat16)
    out, _ = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", shuffle, res, tensor_cores=True)
    ref = _torch_conv(x, w, b, ks, "gelu", shuffle, res)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)
    simt, _ = ops.conv2d_nhwc(x, w, b, 1, ks // 2, "gelu", shuffle, res, tensor_cores=False)
    torch.testing.assert_close(out.float(), simt.float(), atol=2e-2, rtol=1e-2)


def test_gaussian_conditional_kernel_bit_exact_indexes_and_symbols():
    g = torch.Generator().manual_seed(3)
    n = 1 << 18
    y = (torch.randn(n, generator=g) * 6).cuda()
    mu = (torch.randn(n, generator=g) * 2).cuda()
    sc = torch.exp(torch.rand(n, generator=g) * 9 - 3).cuda()          # 0.05 .. 400: both clamps of the table are hit
    tab = mo.scale_table()
    sc[:64] = tab.cuda()                                                # exact table entries (the <= comparison)
    y_hat, lik, sym, idx = ops.gaussian_conditional(y, sc, mu, tab)
    assert torch.equal(sym.cpu(), torch.round(y.cpu() - mu.cpu()).to(torch.int32))
    assert torch.equal(y_hat.cpu(), torch.round(y.cpu() - mu.cpu()) + mu.cpu())
    assert torch.equal(idx.cpu(), mo.cdf_indexes(sc.cpu(), tab))
    ref = mo.gaussian_likelihood(y_hat.cpu(), sc.cpu(), mu.cpu())
    np.testing.assert_allclose(lik.cpu().numpy(), ref.numpy(), atol=3e-7, rtol=2e-5)
    # empty input is a no-op
    e = torch.empty(0, device="cuda")
    assert ops.gaussian_conditional(e, e, e)[0].numel() == 0
