"""Kernel-level entry points (tests and micro-benchmarks): thin wrappers over the C ABI."""
import ctypes as C

import torch

from . import _lib

ACT = {None: 0, "none": 0, "gelu": 1, "half_tanh": 2}


def conv2d_nhwc(x, weight, bias=None, stride=1, pad=0, act=None, shuffle=False, residual=None, tensor_cores=True, iters=1):
    """x: CUDA NHWC tensor [B,H,W,Cin], float32 (validation mode) or bfloat16 (fast mode); weight: [N,Cin,ks,ks] (any
    device, fp32); returns (out NHWC of x.dtype, avg ms of launches 2..iters).  tensor_cores: 0 CUDA cores, 1 tcgen05 one-SM kernel,
    2 as 1 with the two-SM kernel (conv3_pair.cu) where it applies (3x3, pad 1, Cin % 64 == 0, N % 256 == 0)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    B, H, W, Cin = x.shape
    N, ks = weight.shape[0], weight.shape[2]
    w = weight.detach().to("cpu", torch.float32).contiguous()
    b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
    Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
    oshape = (B, 2 * Ho, 2 * Wo, N // 4) if shuffle else (B, Ho, Wo, N)
    out = torch.empty(oshape, dtype=x.dtype, device=x.device)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == x.dtype and residual.is_contiguous()
    ms = C.c_float(0)
    prec = _lib.PREC_BF16 if x.dtype == torch.bfloat16 else _lib.PREC_FP32
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_conv2d_nhwc(prec, int(tensor_cores), C.c_void_p(x.data_ptr()), B, H, W, Cin,
                                               C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()) if b is not None else None,
                                               N, ks, stride, pad, ACT[act], 1 if shuffle else 0,
                                               C.c_void_p(residual.data_ptr()) if residual is not None else None,
                                               C.c_void_p(out.data_ptr()), iters, C.byref(ms), C.c_void_p(st)))
    return out, float(ms.value)


def dwconv3x3_nhwc(x, weight, bias, stride=1, act=None, iters=1):
    """Depthwise 3x3 (pad 1) on a CUDA NHWC tensor (fp32 or bf16); weight [C,1,3,3], bias [C] -> (out, avg ms)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    B, H, W, Cc = x.shape
    w = weight.detach().to("cpu", torch.float32).contiguous()
    b = bias.detach().to("cpu", torch.float32).contiguous()
    out = torch.empty((B, (H - 1) // stride + 1, (W - 1) // stride + 1, Cc), dtype=x.dtype, device=x.device)
    ms = C.c_float(0)
    prec = _lib.PREC_BF16 if x.dtype == torch.bfloat16 else _lib.PREC_FP32
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_dwconv3x3_nhwc(prec, C.c_void_p(x.data_ptr()), B, H, W, Cc, C.c_void_p(w.data_ptr()),
                                                  C.c_void_p(b.data_ptr()), stride, ACT[act], C.c_void_p(out.data_ptr()),
                                                  iters, C.byref(ms), C.c_void_p(st)))
    return out, float(ms.value)


def dsconv_nhwc(x, dw_weight, dw_bias, pw_weight, pw_bias, stride=1, act=None, residual=None, fuse=True, iters=1):
    """DepthWiseConv (dw3x3 pad 1 + bias -> 1x1 + bias -> act -> + residual) on a CUDA NHWC tensor -> (out, avg ms)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    B, H, W, Cin = x.shape
    N = pw_weight.shape[0]
    dw = dw_weight.detach().to("cpu", torch.float32).contiguous()
    db = dw_bias.detach().to("cpu", torch.float32).contiguous()
    pw = pw_weight.detach().to("cpu", torch.float32).contiguous()
    pb = pw_bias.detach().to("cpu", torch.float32).contiguous()
    out = torch.empty((B, (H - 1) // stride + 1, (W - 1) // stride + 1, N), dtype=x.dtype, device=x.device)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == x.dtype and residual.is_contiguous()
    ms = C.c_float(0)
    prec = _lib.PREC_BF16 if x.dtype == torch.bfloat16 else _lib.PREC_FP32
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_dsconv_nhwc(prec, int(fuse), C.c_void_p(x.data_ptr()), B, H, W, Cin,
                                               C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr()), C.c_void_p(pw.data_ptr()),
                                               C.c_void_p(pb.data_ptr()), N, stride, ACT[act],
                                               C.c_void_p(residual.data_ptr()) if residual is not None else None,
                                               C.c_void_p(out.data_ptr()), iters, C.byref(ms), C.c_void_p(st)))
    return out, float(ms.value)


def ds_gdn_nhwc(x, dw_weight, dw_bias, pw_weight, pw_bias, gamma, beta, inverse=False, residual=None, fuse=2, iters=1):
    """Tail of ResidualBlockWithStride / ResidualBlockUpsample on a CUDA bf16 NHWC tensor: v = DepthWiseConv(x),
    out = v * rsqrt(gamma v^2 + beta) (IGDN: * sqrt) + residual, with EFFECTIVE gamma [C,C] / beta [C] -> (out, avg ms).
    fuse 2: the two-SM kernel; 1: two fused kernels; 0: unfused."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.bfloat16
    B, H, W, Cc = x.shape
    hs = [t.detach().to("cpu", torch.float32).contiguous() for t in (dw_weight, dw_bias, pw_weight, pw_bias, gamma, beta)]
    out = torch.empty_like(x)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == x.dtype and residual.is_contiguous()
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_ds_gdn_nhwc(int(fuse), C.c_void_p(x.data_ptr()), B, H, W, Cc, *[C.c_void_p(h.data_ptr()) for h in hs],
                                               1 if inverse else 0, C.c_void_p(residual.data_ptr()) if residual is not None else None,
                                               C.c_void_p(out.data_ptr()), iters, C.byref(ms), C.c_void_p(st)))
    return out, float(ms.value)


def gaussian_conditional(y, scales, means, scale_table=None):
    """Fused quantise / likelihood / CDF-index on flat fp32 CUDA tensors -> (y_hat, lik, symbols, indexes)."""
    n = y.numel()
    tab = None if scale_table is None else scale_table.detach().to(y.device, torch.float32).contiguous()
    assert tab is None or tab.numel() == 64
    y_hat, lik = torch.empty_like(y), torch.empty_like(y)
    sym = torch.empty(y.shape, dtype=torch.int32, device=y.device)
    idx = torch.empty(y.shape, dtype=torch.int32, device=y.device)
    with torch.cuda.device(y.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_gaussian_conditional(C.c_void_p(y.data_ptr()), C.c_void_p(scales.data_ptr()),
                                                        C.c_void_p(means.data_ptr()), n,
                                                        C.c_void_p(tab.data_ptr()) if tab is not None else None,
                                                        C.c_void_p(y_hat.data_ptr()),
                                                        C.c_void_p(lik.data_ptr()), C.c_void_p(sym.data_ptr()),
                                                        C.c_void_p(idx.data_ptr()), C.c_void_p(st)))
    return y_hat, lik, sym, idx


def local_attn(F, rel_bias, impl=2, iters=1):
    """LocalContext windowed attention (context.py:80-107) on a CUDA tensor.
    impl 0 / 1: F fp32 [B,H,W,96] in the reference's interleaved head order -> O [B,H,W,25,32] fp32 / bf16 (all pixels);
    impl 2: F bf16 [B,H,W,96] head-major -> O bf16 [B,H,W/2,25,32] (non-anchor pixels, squeezed).  -> (O, avg ms)"""
    assert F.is_cuda and F.is_contiguous() and F.shape[-1] == 96
    B, H, W, _ = F.shape
    assert F.dtype == (torch.bfloat16 if impl == 2 else torch.float32)
    rb = rel_bias.detach().to(F.device, torch.float32).contiguous()
    assert rb.numel() == 2 * 625
    if impl == 2:
        O = torch.empty((B, H, W // 2, 25, 32), dtype=torch.bfloat16, device=F.device)
    else:
        O = torch.empty((B, H, W, 25, 32), dtype=torch.float32 if impl == 0 else torch.bfloat16, device=F.device)
    ms = C.c_float(0)
    with torch.cuda.device(F.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_local_attn(impl, C.c_void_p(F.data_ptr()), B, H, W, C.c_void_p(rb.data_ptr()),
                                              C.c_void_p(O.data_ptr()), iters, C.byref(ms), C.c_void_p(st)))
    return O, float(ms.value)


def lin_attn(qkv, heads, par_kv=0, par_q=0, iters=1):
    """Kernelised global attention (context.py:169-193,226-245) on a CUDA NHWC tensor [B,H,W,3D] (Q | K | V; fp32 or bf16):
    K soft-maxed over positions, Q over the head's channels -> (out [B,H,W,D] of the same dtype, avg ms).  par: 0 none, 1 anchor,
    2 non-anchor positions only."""
    assert qkv.is_cuda and qkv.is_contiguous() and qkv.dtype in (torch.float32, torch.bfloat16) and qkv.shape[-1] % 3 == 0
    B, H, W, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, H, W, D), dtype=qkv.dtype, device=qkv.device)
    ms = C.c_float(0)
    prec = _lib.PREC_BF16 if qkv.dtype == torch.bfloat16 else _lib.PREC_FP32
    with torch.cuda.device(qkv.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_lin_attn(prec, C.c_void_p(qkv.data_ptr()), B, H, W, D, heads, par_kv, par_q,
                                            C.c_void_p(out.data_ptr()), iters, C.byref(ms), C.c_void_p(st)))
    return out, float(ms.value)


def ga_head(x, dw_weight, dw_bias, pw_weight, pw_bias, skip_weight, skip_bias, iters=1):
    """g_a stage-0 head of the bf16 path: x CUDA fp32 NCHW [B,3,H,W] -> (GELU(pw(dw_s2 x)), skip_s2(x)) bf16 NHWC, avg ms."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32 and x.shape[1] == 3
    B, _, H, W = x.shape
    N = pw_weight.shape[0]
    hs = [t.detach().to("cpu", torch.float32).contiguous() for t in (dw_weight, dw_bias, pw_weight, pw_bias, skip_weight, skip_bias)]
    t_out = torch.empty((B, H // 2, W // 2, N), dtype=torch.bfloat16, device=x.device)
    s_out = torch.empty_like(t_out)
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_ga_head(C.c_void_p(x.data_ptr()), B, H, W, *[C.c_void_p(h.data_ptr()) for h in hs], N,
                                           C.c_void_p(t_out.data_ptr()), C.c_void_p(s_out.data_ptr()), iters, C.byref(ms),
                                           C.c_void_p(st)))
    return t_out, s_out, float(ms.value)


def final_subpel(x, weight, bias, impl=1, iters=1):
    """Final g_s layer of the bf16 path: x CUDA bf16 NHWC [B,H,W,C], weight [12,C,3,3], bias [12] -> fp32 NCHW [B,3,2H,2W]
    (impl 0: implicit-GEMM conv, impl 1: shift-sum form), avg ms."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.bfloat16
    B, H, W, Cin = x.shape
    w = weight.detach().to("cpu", torch.float32).contiguous()
    b = bias.detach().to("cpu", torch.float32).contiguous()
    out = torch.empty((B, 3, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_final_subpel(impl, C.c_void_p(x.data_ptr()), B, H, W, Cin, C.c_void_p(w.data_ptr()),
                                                C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), iters, C.byref(ms),
                                                C.c_void_p(st)))
    return out, float(ms.value)


def chain3(x, w1, b1, w2, b2, w3, b3, ln=None, iters=1):
    """Three chained per-row layers in one launch (chain3.cu).  x: CUDA bf16 [M, K1]; w*: [N, K] fp32 (any device).
    ln = None: EntropyParameters tail, out fp32 [M, N3] = W3 GELU(W2 GELU(W1 x + b1) + b2) + b3;
    ln = (gamma, beta): LocalContext tail, p = W1 x + b1, out bf16 = p + W3 GELU(W2 LayerNorm(p) + b2) + b3.  -> (out, avg ms)"""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.bfloat16 and x.dim() == 2
    M, K1 = x.shape
    hs = [t.detach().to("cpu", torch.float32).contiguous() for t in (w1, b1, w2, b2, w3, b3)]
    N1, N2, N3 = hs[0].shape[0], hs[2].shape[0], hs[4].shape[0]
    mode = 0 if ln is None else 1
    g = bt = None
    if ln is not None:
        g, bt = (t.detach().to("cpu", torch.float32).contiguous() for t in ln)
    out = torch.empty((M, N3), dtype=torch.float32 if mode == 0 else torch.bfloat16, device=x.device)
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mlic_chain3(mode, C.c_void_p(x.data_ptr()), M, K1, C.c_void_p(hs[0].data_ptr()), C.c_void_p(hs[1].data_ptr()), N1,
                                          C.c_void_p(hs[2].data_ptr()), C.c_void_p(hs[3].data_ptr()), N2, C.c_void_p(hs[4].data_ptr()),
                                          C.c_void_p(hs[5].data_ptr()), N3, C.c_void_p(g.data_ptr()) if g is not None else None,
                                          C.c_void_p(bt.data_ptr()) if bt is not None else None, C.c_void_p(out.data_ptr()), iters, C.byref(ms),
                                          C.c_void_p(st)))
    return out, float(ms.value)
