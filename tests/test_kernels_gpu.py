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
    # large shapes: 957 M tiles (an odd count: the last CTA pair of the two-SM kernels holds a single tile) and Cin = 320
    (1, 264, 464, 192, 768, 3, True), (1, 136, 240, 320, 256, 3, False),
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


PAIR_CONV_SHAPES = [  # B, H, W, Cin, N, shuffle, act: ragged edges, an odd tile count (the last pair holds one tile), Cin = 320, no shuffle
    (1, 16, 24, 192, 768, True, "gelu"), (3, 13, 21, 128, 512, True, None), (1, 20, 37, 64, 256, False, "gelu"),
    (2, 17, 30, 320, 768, True, "gelu"), (1, 264, 464, 192, 768, True, None), (1, 8, 16, 192, 1024, True, None),
]


@pytest.mark.parametrize("B,H,W,Cin,N,shuffle,act", PAIR_CONV_SHAPES)
def test_conv3x3_two_sm_kernel(B, H, W, Cin, N, shuffle, act):
    """conv3_pair.cu (tcgen05 cta_group::2, M = 256 over a CTA pair, each CTA staging half of the weight tile): same operands and
    accumulation as the one-SM kernel, so the two agree to the last bf16 bit up to fp32 summation order; both against torch fp32
    on the bf16-rounded operands."""
    torch.manual_seed(5)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Cin, 3, 3) / (Cin * 9) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(N) * 0.1
    out, _ = ops.conv2d_nhwc(x, w, b, 1, 1, act, shuffle, None, tensor_cores=2)
    one, _ = ops.conv2d_nhwc(x, w, b, 1, 1, act, shuffle, None, tensor_cores=1)
    ref = _torch_conv(x, w, b, 3, act, shuffle, None)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)
    assert (out.float() - one.float()).abs().max() <= 2.0 ** -6 * max(1.0, float(ref.abs().max()))
    assert float((out != one).float().mean()) < 0.02


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("B,H,W,Cin,N", [(1, 68, 120, 288, 96), (3, 13, 21, 96, 96), (2, 17, 30, 32, 64), (1, 9, 40, 160, 96), (2, 33, 70, 104, 72)])
def test_conv5x5_column_shifted_patch_kernel(B, H, W, Cin, N, variant):
    """conv_halo.cu (the 5x5 re-projections of the global contexts: activations staged once per (chunk, kx), the row taps are aligned UMMA
    descriptors into the patch; variant 2 = pixels as the M operand, 3 = roles swapped: weights as M, 256 pixels as N, direct NHWC stores):
    against torch fp32 on the bf16-rounded operands and against the plain implicit-GEMM kernel."""
    torch.manual_seed(6)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Cin, 5, 5) / (Cin * 25) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(N) * 0.1
    out, _ = ops.conv2d_nhwc(x, w, b, 1, 2, None, False, None, tensor_cores=variant)
    one, _ = ops.conv2d_nhwc(x, w, b, 1, 2, None, False, None, tensor_cores=1)
    ref = _torch_conv(x, w, b, 5, None, False, None)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)
    assert (out.float() - one.float()).abs().max() <= 2.0 ** -6 * max(1.0, float(ref.abs().max()))


def _lin_attn_torch(qkv, heads, par_kv=0, par_q=0):
    """context.py:169-193,226-245 on an NHWC [B,H,W,3D] tensor: per head softmax of K over positions, of Q over channels."""
    B, H, W, D3 = qkv.shape
    D = D3 // 3
    hd = D // heads
    q, k, v = (t.reshape(B, H * W, heads, hd).double() for t in qkv.float().split(D, dim=-1))
    hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    anchor = (((hh + ww) % 2) == 1).reshape(-1).to(qkv.device)

    def keep(par):
        return torch.ones_like(anchor) if par == 0 else (anchor if par == 1 else ~anchor)
    kk, kq = keep(par_kv), keep(par_q)
    k = k.masked_fill(~kk[None, :, None, None], float("-inf"))
    kh = torch.softmax(k, dim=1)
    qh = torch.softmax(q, dim=-1)
    ctx = torch.einsum("bphc,bphd->bhcd", kh, v * kk[None, :, None, None])
    o = torch.einsum("bhcd,bphc->bphd", ctx, qh) * kq[None, :, None, None]
    return o.reshape(B, H, W, D).float()


@pytest.mark.parametrize("B,H,W,D,heads,par_kv,par_q", [(2, 17, 30, 96, 3, 0, 0), (1, 68, 120, 288, 9, 0, 0), (2, 16, 24, 32, 2, 1, 2), (1, 13, 21, 32, 1, 0, 0)])
def test_linear_attention_kernels(B, H, W, D, heads, par_kv, par_q):
    """LinearGlobalInterContext / LinearGlobalIntraContext attention: the fp32 kernels against a float64 torch
    statement, the bf16 kernels (mma.sync context matrix and, for head dim 32 without parity, the mma.sync output kernel) against the
    same statement on the bf16-rounded input."""
    torch.manual_seed(7)
    qkv = torch.randn(B, H, W, 3 * D, device="cuda")
    ref = _lin_attn_torch(qkv, heads, par_kv, par_q)
    out, _ = ops.lin_attn(qkv, heads, par_kv, par_q)
    torch.testing.assert_close(out, ref, atol=2e-5, rtol=1e-4)
    qb = qkv.to(torch.bfloat16)
    refb = _lin_attn_torch(qb, heads, par_kv, par_q)
    outb, _ = ops.lin_attn(qb, heads, par_kv, par_q)
    torch.testing.assert_close(outb.float(), refb, atol=3e-3 * float(refb.abs().max()) + 1e-4, rtol=2e-2)


@pytest.mark.parametrize("B,H,W,Cin,N", [(12, 68, 120, 352, 224), (24, 68, 60, 960, 320), (48, 37, 53, 264, 200)])
def test_wide_1x1_gemm_two_sm_kernel(B, H, W, Cin, N):
    """conv3_pair.cu with ks = 1: the wide 1x1 GEMMs of the entropy model (EntropyParameters / LRP first layers) on a CTA pair, each CTA
    staging half of the weight rows; one column tile (N = 224: a 32-column tail group), two (N = 320: 192 + 128) and a ragged N / Cin."""
    torch.manual_seed(8)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Cin, 1, 1) / Cin ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(N) * 0.1
    out, _ = ops.conv2d_nhwc(x, w, b, 1, 0, "gelu", False, None, tensor_cores=2)
    one, _ = ops.conv2d_nhwc(x, w, b, 1, 0, "gelu", False, None, tensor_cores=1)
    ref = _torch_conv(x, w, b, 1, "gelu", False, None)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)
    assert torch.equal(out, one)                       # same operands, same accumulation order


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


def _local_attn_torch(Fq, rel_bias):
    """Plain PyTorch fp32 statement of LocalContext's windowed attention (context.py:80-107):
    Fq [B,H,W,96] (q|k|v, channel c = d*2 + head) -> O [B,H,W,25,32] (channel = head*16 + d)."""
    B, H, W, _ = Fq.shape
    Fp = F.pad(Fq.permute(0, 3, 1, 2), (2, 2, 2, 2))                                 # zeros of q/k/v outside the image
    win = F.unfold(Fp, 5).reshape(B, 96, 25, H, W).permute(0, 3, 4, 2, 1)             # [B,H,W,25,96]
    q, k, v = [win[..., u * 32:(u + 1) * 32].reshape(B, H, W, 25, 16, 2).permute(0, 1, 2, 5, 3, 4) for u in range(3)]  # [B,H,W,head,25,16]
    hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    anch = torch.zeros(H + 4, W + 4, dtype=torch.bool)
    anch[2:-2, 2:-2] = ((hh + ww) % 2 == 1)
    aw = anch.unfold(0, 5, 1).unfold(1, 5, 1).reshape(H, W, 25).to(Fq.device)         # tap is an in-image anchor
    mask = torch.where(aw[:, :, :, None] & aw[:, :, None, :], 0.0, -100.0)            # [H,W,25,25]
    A = (q * 0.25) @ k.transpose(-1, -2) + rel_bias.to(Fq.device).reshape(1, 1, 1, 2, 25, 25) + mask[None, :, :, None]
    o = torch.softmax(A, -1) @ v                                                       # [B,H,W,head,25,16]
    return o.permute(0, 1, 2, 4, 3, 5).reshape(B, H, W, 25, 32)


@pytest.mark.parametrize("B,H,W", [(1, 8, 16), (2, 13, 22), (1, 68, 120)])
def test_local_attention_kernels(B, H, W):
    g = torch.Generator().manual_seed(7)
    Fq = torch.randn(B, H, W, 96, generator=g).cuda()
    rb = (torch.randn(2, 25, 25, generator=g) * 0.5).cuda()
    O0, _ = ops.local_attn(Fq, rb, impl=0)
    ref = _local_attn_torch(Fq, rb)
    np.testing.assert_allclose(O0.cpu().numpy(), ref.cpu().numpy(), atol=2e-5, rtol=0)
    # tensor-core kernel: bf16 inputs in head-major channel order, non-anchor pixels only, squeezed
    Fb = Fq.to(torch.bfloat16)
    ref_b = _local_attn_torch(Fb.float(), rb)
    Fhm = Fb.reshape(B, H, W, 3, 16, 2).permute(0, 1, 2, 3, 5, 4).reshape(B, H, W, 96).contiguous()
    O2, _ = ops.local_attn(Fhm, rb, impl=2)
    hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W // 2), indexing="ij")
    wfull = 2 * ww + (hh % 2)
    want = ref_b[:, hh, wfull]                                                          # [B,H,W/2,25,32]
    err = (O2.float() - want).abs().max().item()
    assert err < 3e-2, err
    assert (O2.float() - want).abs().mean().item() < 3e-3


def test_ga_head_kernel():
    g = torch.Generator().manual_seed(3)
    B, H, W, N = 2, 36, 260, 192
    x = torch.rand(B, 3, H, W, generator=g).cuda()
    dw, db = torch.randn(3, 1, 3, 3, generator=g) / 3, torch.randn(3, generator=g) * 0.1
    pw, pb = torch.randn(N, 3, 1, 1, generator=g), torch.randn(N, generator=g) * 0.1
    sw, sb = torch.randn(N, 3, 1, 1, generator=g), torch.randn(N, generator=g) * 0.1
    t, s, _ = ops.ga_head(x, dw, db, pw, pb, sw, sb)
    d = F.conv2d(x, dw.cuda(), db.cuda(), stride=2, padding=1, groups=3)
    t_ref = F.gelu(F.conv2d(d, pw.cuda(), pb.cuda())).permute(0, 2, 3, 1)
    s_ref = F.conv2d(x, sw.cuda(), sb.cuda(), stride=2).permute(0, 2, 3, 1)
    for ours, ref in ((t, t_ref), (s, s_ref)):
        tol = 2.0 ** -8 * ref.abs() + 2e-3                  # bf16 rounding of the stored value + fast GELU
        assert bool(((ours.float() - ref).abs() <= tol).all()), (ours.float() - ref).abs().max().item()


@pytest.mark.parametrize("B,H,W,C,stride,act", [(1, 20, 36, 192, 1, None), (2, 17, 30, 352, 1, "gelu"), (1, 37, 53, 104, 2, None),
                                                (1, 24, 40, 192, 2, None), (1, 9, 7, 8, 1, None)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depthwise_kernels(B, H, W, C, stride, act, dtype):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, H, W, C, generator=g).cuda().to(dtype)
    w, b = torch.randn(C, 1, 3, 3, generator=g) * 0.3, torch.randn(C, generator=g) * 0.1
    out, _ = ops.dwconv3x3_nhwc(x, w, b, stride, act)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), stride=stride, padding=1, groups=C)
    if act == "gelu":
        ref = F.gelu(ref)
    ref = ref.permute(0, 2, 3, 1)
    if dtype == torch.float32:
        np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=2e-5, rtol=1e-5)
    else:
        tol = 2.0 ** -8 * ref.abs() + 2e-3
        assert bool(((out.float() - ref).abs() <= tol).all()), (out.float() - ref).abs().max().item()


@pytest.mark.parametrize("B,H,W,Cin,N,act,res", [(1, 24, 40, 192, 192, "gelu", True), (2, 17, 30, 224, 128, "gelu", False),
                                                 (1, 20, 28, 160, 192, "gelu", False), (2, 17, 30, 128, 32, "half_tanh", True),
                                                 (1, 13, 21, 104, 72, None, True)])
@pytest.mark.parametrize("fuse", [True, False])
def test_depthwise_separable_conv_bf16(B, H, W, Cin, N, act, res, fuse):
    """DepthWiseConv (conv.py:46-63) as one tcgen05 kernel with the depthwise producer (fuse) or as dw kernel + GEMM, incl. a
    channel tail in the last 64-channel chunk and the LRP tail (0.5 tanh + residual, quantization.py:30-45)."""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    dw, db = torch.randn(Cin, 1, 3, 3, generator=g) / 3, torch.randn(Cin, generator=g) * 0.1
    pw = (torch.randn(N, Cin, 1, 1, generator=g) / Cin ** 0.5).to(torch.bfloat16).float()
    pb = torch.randn(N, generator=g) * 0.1
    r = torch.randn(B, H, W, N, generator=g).cuda().to(torch.bfloat16) if res else None
    out, _ = ops.dsconv_nhwc(x, dw, db, pw, pb, 1, act, r, fuse)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), dw.cuda(), db.cuda(), padding=1, groups=Cin)
    y = F.conv2d(y, pw.cuda(), pb.cuda())
    y = F.gelu(y) if act == "gelu" else (0.5 * torch.tanh(y) if act == "half_tanh" else y)
    y = y.permute(0, 2, 3, 1)
    if r is not None:
        y = y + r.float()
    # the depthwise result is rounded to bf16 before the GEMM (2^-9 relative on ~sqrt(Cin) terms of size ~1/sqrt(Cin))
    tol = 2.0 ** -7 * y.abs() + 1.5e-2
    assert bool(((out.float() - y).abs() <= tol).all()), (out.float() - y).abs().max().item()
    assert float((out.float() - y).pow(2).mean().sqrt()) < 4e-3


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("B,H,W,C", [(1, 16, 8, 64), (2, 30, 50, 192), (1, 67, 29, 192)])
def test_final_subpel_conv(impl, B, H, W, C):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, C, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(12, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(torch.bfloat16).float()
    b = torch.randn(12, generator=g) * 0.1
    out, _ = ops.final_subpel(x, w, b, impl=impl)
    ref = F.pixel_shuffle(F.conv2d(x.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), padding=1), 2)
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("M,K1,N1,ln", [(1000, 320, 256, False), (128 * 149 + 37, 320, 256, False), (777, 128, 128, False), (1000, 800, 64, True), (128 * 150 + 5, 800, 64, True)])
def test_chained_three_layer_kernel(M, K1, N1, ln):
    """chain3.cu (EntropyParameters layers 1..3, entropy.py:13-17; the LocalContext tail, context.py:108-110): the intermediates are
    bf16 A operands in tensor memory, so the statement it is compared with rounds them to bf16 at the same places (operands bf16,
    accumulation fp32, exact erf GELU: the kernel's tanh-form GELU is within 3e-4); ragged last tile and more tiles than SMs included."""
    torch.manual_seed(8)
    x = torch.randn(M, K1, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(N1, K1) / K1 ** 0.5).to(torch.bfloat16).float()
    w2 = (torch.randn(128, N1) / N1 ** 0.5).to(torch.bfloat16).float()
    w3 = (torch.randn(64, 128) / 128 ** 0.5).to(torch.bfloat16).float()
    b1, b2, b3 = torch.randn(N1) * 0.2, torch.randn(128) * 0.2, torch.randn(64) * 0.2
    bfr = lambda t: t.to(torch.bfloat16).float()
    xf = x.float().cpu()
    if not ln:
        out, _ = ops.chain3(x, w1, b1, w2, b2, w3, b3)
        h1 = bfr(torch.nn.functional.gelu(xf @ w1.T + b1))
        h2 = bfr(torch.nn.functional.gelu(h1 @ w2.T + b2))
        ref = h2 @ w3.T + b3
        assert out.dtype == torch.float32
        torch.testing.assert_close(out.cpu(), ref, atol=2e-2, rtol=1e-2)
    else:
        g, bt = 1.0 + 0.1 * torch.randn(N1), 0.1 * torch.randn(N1)
        out, _ = ops.chain3(x, w1, b1, w2, b2, w3, b3, ln=(g, bt))
        p = xf @ w1.T + b1
        h1 = bfr(torch.nn.functional.layer_norm(p, (N1,), g, bt, 1e-5))
        h2 = bfr(torch.nn.functional.gelu(h1 @ w2.T + b2))
        ref = p + h2 @ w3.T + b3
        assert out.dtype == torch.bfloat16
        torch.testing.assert_close(out.float().cpu(), ref, atol=4e-2, rtol=2e-2)
