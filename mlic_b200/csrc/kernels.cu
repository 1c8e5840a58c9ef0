// CUDA-core kernels of the MLIC++ engine (sm_100a): the generic implicit-GEMM convolution used in
// fp32 validation mode and for the GEMM shapes the tcgen05 kernel does not take, the depthwise 3x3
// stencil, layout conversion, EntropyBottleneck, LocalContext windowed attention, the linear global
// attention and the fused quantise / likelihood / CDF-index kernels.
//
// Reference semantics (paths relative to /root/reference/MLIC++): modules/layers/conv.py:46-63,
// modules/transform/context.py:67-112,169-193,226-245, utils/ckbd.py:35-73,123-144, and the
// CompressAI 1.2.6 GaussianConditional / EntropyBottleneck behaviour restated in SURVEY.md A.7/A.8.
#include "kernels.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

#include <math.h>
#include <algorithm>

namespace mlic {

bool pdl_enabled() {
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    return pdl != 0;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
// Generic implicit-GEMM convolution, fp32 accumulate on CUDA cores.
//   D[m][n] = sum_k A[m][k] * Wt[n][k],  m = (b,oh,ow),  k = tap*Cin + c  (tap = ky*ks + kx)
// Tile 128 x 64 x 16, 256 threads, 8x4 micro-tile, register-prefetched double buffering.
// Every output is one sequential fp32 FMA chain over k in increasing order (deterministic).
// ------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 64, GK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(const T* __restrict__ in, ConvGeom g,
                                                             const float* __restrict__ Wt, Epi e, int vec) {
    pdl_wait();
    __shared__ __align__(16) float As[2][GK][GM + 4];
    __shared__ __align__(16) float Bs[2][GK][GN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long Mtot = (long long)g.B * g.Hout * g.Wout;
    const long long m0 = (long long)blockIdx.x * GM;
    const int n0 = blockIdx.y * GN;
    const int N = e.N;

    // A loader: two rows per thread (lr, lr+64), 4 consecutive k each.
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    int ab[2], aoh[2], aow[2];
    bool aval[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        long long m = m0 + lr + r * 64;
        aval[r] = m < Mtot;
        long long mm = aval[r] ? m : 0;
        int hw = g.Hout * g.Wout;
        ab[r] = (int)(mm / hw);
        int rem = (int)(mm - (long long)ab[r] * hw);
        aoh[r] = rem / g.Wout;
        aow[r] = rem - aoh[r] * g.Wout;
    }
    // B loader: one row n per thread (n = n0 + lr), 4 consecutive k.  (GN = 64 rows x 16 k = 256 x 4)
    const int bn = n0 + lr;
    const bool bval = bn < N;
    const bool avec = vec && ((g.Cin & 3) == 0);
    const bool bvec = (g.Ktot & 3) == 0;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[2][4], rb[4];
    const int nchunk = (g.Ktot + GK - 1) / GK;

    auto gload = [&](int kc) {
        const int k0 = kc * GK + lk;
        // A
        if (avec) {
            int tap = k0 / g.Cin;
            int c = k0 - tap * g.Cin;
            int ky = tap / g.ks, kx = tap - ky * g.ks;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                ra[r][0] = ra[r][1] = ra[r][2] = ra[r][3] = 0.f;
                if (aval[r] && k0 < g.Ktot) {
                    int ih = aoh[r] * g.stride - g.pad + ky;
                    int iw = aow[r] * g.stride - g.pad + kx;
                    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
                        const T* p = in + (((size_t)ab[r] * g.H + ih) * g.W + iw) * g.ld + c;
                        load4(p, ra[r]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = 0.f;
                    int k = k0 + j;
                    if (aval[r] && k < g.Ktot) {
                        int tap = k / g.Cin;
                        int c = k - tap * g.Cin;
                        int ky = tap / g.ks, kx = tap - ky * g.ks;
                        int ih = aoh[r] * g.stride - g.pad + ky;
                        int iw = aow[r] * g.stride - g.pad + kx;
                        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                            v = to_f<T>(in[(((size_t)ab[r] * g.H + ih) * g.W + iw) * g.ld + c]);
                    }
                    ra[r][j] = v;
                }
            }
        }
        // B
        rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
        if (bval) {
            const float* wp = Wt + (size_t)bn * g.Ktot + k0;
            if (bvec && k0 + 3 < g.Ktot) {
                float4 t = *reinterpret_cast<const float4*>(wp);
                rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k0 + j < g.Ktot) rb[j] = wp[j];
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[buf][lk + j][lr] = ra[0][j];
            As[buf][lk + j][lr + 64] = ra[1][j];
            Bs[buf][lk + j][lr] = rb[j];
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kc = 0; kc < nchunk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunk) gload(kc + 1);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nchunk) sstore(buf ^ 1);
        __syncthreads();
    }

    const int n = n0 + tx * 4;
    if (n >= N) return;
    const int hw = g.Hout * g.Wout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= Mtot) continue;
        int b = (int)(m / hw);
        int rem = (int)(m - (long long)b * hw);
        int h = rem / g.Wout;
        int w = rem - h * g.Wout;
        epi_store4<T>(e, b, h, w, n, acc[i], vec != 0);
    }
}

// 1x1 convolution with Cin <= 4 (the two 3 -> N layers that read the image): store-bound, so one thread produces 8
// output columns of one pixel; same k-order fp32 FMA chain as the GEMM kernel.
template <typename T>
__global__ void __launch_bounds__(256) pw_small_cin_kernel(const T* __restrict__ in, ConvGeom g, const float* __restrict__ Wt,
                                                           Epi e, int vec) {
    pdl_wait();
    // thread = (pixel lane, 8-column group); the group's weights stay in registers across the pixel loop
    const int ng = (e.N + 7) / 8;
    const int cgp = threadIdx.x % ng, pl = threadIdx.x / ng, npl = blockDim.x / ng;
    if (pl >= npl) return;
    const int n = cgp * 8;
    float wr[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) wr[j][c] = (n + j < e.N && c < g.Cin) ? Wt[(size_t)(n + j) * g.Cin + c] : 0.f;
    const bool simple = vec && n + 8 <= e.N && !e.res && !e.gdn && !e.premask && !e.pm_w && !e.postmask && !e.shuffle && !e.out2 &&
                        !e.out_f32 && !e.nchw && e.act != ACT_HALF_TANH;
    float br[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = (e.bias && n + j < e.N) ? e.bias[n + j] : 0.f;
    const unsigned npix = (unsigned)g.B * g.Hout * g.Wout;         // < 2^31 (host-checked): 32-bit coordinate arithmetic
    for (unsigned pix = blockIdx.x * npl + pl; pix < npix; pix += gridDim.x * npl) {
        const unsigned q = pix / (unsigned)g.Wout;
        const int w = (int)(pix - q * g.Wout);
        const int b = (int)(q / (unsigned)g.Hout);
        const int h = (int)(q - (unsigned)b * g.Hout);
        const T* ip = in + (((size_t)b * g.H + (size_t)h * g.stride) * g.W + (size_t)w * g.stride) * g.ld;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < g.Cin) x[c] = to_f<T>(ip[c]);
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c < g.Cin) acc = fmaf(x[c], wr[j][c], acc);
            a[j] = acc;
        }
        if (simple) {            // bias + optional GELU, dense NHWC store of 8 columns (16 B in bf16)
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += br[j]; if (e.act == ACT_GELU) a[j] = gelu_erf(a[j]); }
            T* op = reinterpret_cast<T*>(e.out) + (size_t)pix * e.out_ld + n;
            store4(op, a);
            store4(op + 4, a + 4);
        } else {
            epi_store4<T>(e, b, h, w, n, a, vec != 0);
            epi_store4<T>(e, b, h, w, n + 4, a + 4, vec != 0);
        }
    }
}

void launch_conv_gemm_simt(int bf, const void* in, const ConvGeom& g, const float* Wt, const Epi& e, int vec,
                           cudaStream_t s) {
    long long Mtot = (long long)g.B * g.Hout * g.Wout;
    if (Mtot == 0 || e.N == 0) return;
    if (g.ks == 1 && g.pad == 0 && g.Cin <= 4 && e.N <= 2048 && Mtot < (1LL << 31)) {
        const int npl = 256 / ((e.N + 7) / 8);
        int blocks = (int)std::min<long long>(cdiv(Mtot, npl), 148LL * 16);
        if (bf) launch_k(pw_small_cin_kernel<bf16>, dim3(blocks), dim3(256), 0, s, (const bf16*)in, g, Wt, e, vec);
        else launch_k(pw_small_cin_kernel<float>, dim3(blocks), dim3(256), 0, s, (const float*)in, g, Wt, e, vec);
        return;
    }
    dim3 grid(cdiv(Mtot, GM), cdiv(e.N, GN));
    if (bf) launch_k(conv_gemm_simt_kernel<bf16>, dim3(grid), dim3(256), 0, s, (const bf16*)in, g, Wt, e, vec);
    else launch_k(conv_gemm_simt_kernel<float>, dim3(grid), dim3(256), 0, s, (const float*)in, g, Wt, e, vec);
}

// ------------------------------------------------------------------------------------------
// g_a stage-0 head, bf16 fast mode (ResidualBlockWithStride(3 -> N, stride 2), res_blk.py:82-93 with
// DepthWiseConv, conv.py:46-63): straight from the fp32 NCHW image
//     t  = GELU(pw(dw3x3_s2(x)))      [B, H/2, W/2, N] bf16 NHWC
//     sk = skip1x1_s2(x)              [B, H/2, W/2, N] bf16 NHWC
// One block = 64 consecutive output pixels of one output row.  The 3 x 3 x 129 input window is staged in shared
// memory with row-contiguous (coalesced) reads, 192 threads evaluate the 3-channel depthwise conv once per pixel, then
// thread (pixel lane, 8-column group) keeps its 2 x 8 x 3 weights in registers and writes 16 B of each output per
// pixel: a warp store covers whole 128-byte lines.  Store-bound: 2 * N * 2 B per output pixel, no intermediate
// tensor (the NHWC copy of x, the depthwise output) touches HBM.
// ------------------------------------------------------------------------------------------
constexpr int GH_PIX = 64;
constexpr int GH_THREADS = 288;             // warps 0-5: conv1 path (GELU), warps 6-8: skip path (a third of the arithmetic per value)
// One role's share of the block: thread (pixel lane pl, 8-column group g) keeps its 8 x 3 weights and 8 biases in registers and
// walks the pixels pl, pl + npl, ... with one pointer increment per pixel.  (The first version re-read its parameters from the
// constant bank, re-derived the role and the 64-bit address in every iteration and branched around the GELU per channel pair:
// 106 / 71 instructions per 16-byte store against 66 / 26 here; profiles/r02_ncu_head.txt.)
struct GhW { float2 wa[4][3], ba[4]; };
__device__ __forceinline__ void gh_load(GhW& w, int g, const float* __restrict__ wsrc, const float* __restrict__ bsrc) {
    const int n = g * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int c = 0; c < 3; ++c) w.wa[j][c] = make_float2(wsrc[(size_t)(n + 2 * j) * 3 + c], wsrc[(size_t)(n + 2 * j + 1) * 3 + c]);
        w.ba[j] = make_float2(bsrc[n + 2 * j], bsrc[n + 2 * j + 1]);
    }
}
template <bool GELU>
__device__ __forceinline__ void gh_role(const GhW& w, const float4* __restrict__ src, int nvalid, int pl, int npl, bf16* __restrict__ dst, size_t dstep) {
#pragma unroll 2
    for (int p = pl; p < nvalid; p += npl) {
        const float4 a = src[p];
        uint32_t o1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 u = __ffma2_rn(make_float2(a.x, a.x), w.wa[j][0], w.ba[j]);
            u = __ffma2_rn(make_float2(a.y, a.y), w.wa[j][1], u);
            u = __ffma2_rn(make_float2(a.z, a.z), w.wa[j][2], u);
            if (GELU) u = gelu2(u);
            __nv_bfloat162 hu = __floats2bfloat162_rn(u.x, u.y);
            o1[j] = *reinterpret_cast<uint32_t*>(&hu);
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
        dst += dstep;
    }
}
// Persistent: a block keeps its weights in registers and walks row tiles blockIdx.x, blockIdx.x + gridDim.x, ... (with one tile per
// block the 32 strided weight loads of every thread and the two barriers were as long as the tile's 48 stores per pixel: halving the
// loop's instructions alone changed nothing, 547 -> 561 us at 4 images).
__global__ void __launch_bounds__(GH_THREADS, 3) ga_head_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ dw9,
                                                      const float* __restrict__ dwb, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ wsk,
                                                      const float* __restrict__ bsk, int N, bf16* __restrict__ t, int t_ld,
                                                      bf16* __restrict__ sk, int sk_ld, int ntiles) {
    pdl_wait();
    __shared__ float sx[3][3][2 * GH_PIX + 2];
    __shared__ float4 st[2][2][GH_PIX];      // [tile parity][0: depthwise result, 1: the stride-2 sample of x]
    const int Ho = H >> 1, Wo = W >> 1;
    const int tiles_w = (Wo + GH_PIX - 1) / GH_PIX;
    const int ng = N >> 3;
    const bool conv = threadIdx.x < 192;                                       // warp-uniform
    const int tr = conv ? threadIdx.x : threadIdx.x - 192, nthr = conv ? 192 : 96;
    const int g = tr % ng, pl = tr / ng, npl = nthr / ng;
    const bool active = pl < npl;
    GhW w;
    if (active) gh_load(w, g, conv ? w1 : wsk, conv ? b1 : bsk);
    float dwr[9], dwbias = 0.f;
    const int dp = threadIdx.x / 3, dc = threadIdx.x % 3;
    if (threadIdx.x < GH_PIX * 3) {
#pragma unroll
        for (int k = 0; k < 9; ++k) dwr[k] = dw9[k * 3 + dc];
        dwbias = dwb[dc];
    }
    // warp rc = (channel, input row) of the 3 x 3 x 129 window: its 129 floats are fetched one tile AHEAD into registers (5 per lane),
    // so the DRAM latency of the next tile's window runs under this tile's stores
    static_assert(GH_THREADS / 32 == 9 && 2 * GH_PIX + 1 <= 5 * 32, "one warp per (channel, row), five floats per lane");
    const int wr = (threadIdx.x >> 5) % 3, wc = (threadIdx.x >> 5) / 3, lane = threadIdx.x & 31;
    float xr[5];
    auto fetch = [&](int tile) {
        int bid = tile;
        const int tw = bid % tiles_w; bid /= tiles_w;
        const int oh = bid % Ho;
        const int b = bid / Ho;
        const int ih = 2 * oh - 1 + wr;
        const float* row = x + (((size_t)b * 3 + wc) * H + ih) * W;
        const int iw0 = 2 * tw * GH_PIX - 1 + lane;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int iw = iw0 + 32 * i;
            xr[i] = (lane + 32 * i < 2 * GH_PIX + 1 && ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(row + iw) : 0.f;
        }
    };
    if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);
    int par = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, par ^= 1) {
        int bid = tile;
        const int tw = bid % tiles_w; bid /= tiles_w;
        const int oh = bid % Ho;
        const int b = bid / Ho;
        const int ow0 = tw * GH_PIX;
#pragma unroll
        for (int i = 0; i < 5; ++i)
            if (lane + 32 * i < 2 * GH_PIX + 2) sx[wc][wr][lane + 32 * i] = xr[i];
        __syncthreads();
        if (threadIdx.x < GH_PIX * 3) {
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) acc = fmaf(sx[dc][ky][2 * dp + kx], dwr[ky * 3 + kx], acc);
            reinterpret_cast<float*>(&st[par][0][dp])[dc] = acc + dwbias;
            reinterpret_cast<float*>(&st[par][1][dp])[dc] = sx[dc][1][2 * dp + 1];
        }
        if (tile + (int)gridDim.x < ntiles) fetch(tile + (int)gridDim.x);
        __syncthreads();       // st[par] complete; sx may be refilled (st[par ^ 1] of the previous tile is read before the barrier above)
        if (active) {
            const int nvalid = min(GH_PIX, Wo - ow0);
            const size_t pix0 = ((size_t)b * Ho + oh) * Wo + ow0 + pl;
            if (conv) gh_role<true>(w, st[par][0], nvalid, pl, npl, t + pix0 * t_ld + g * 8, (size_t)npl * t_ld);
            else gh_role<false>(w, st[par][1], nvalid, pl, npl, sk + pix0 * sk_ld + g * 8, (size_t)npl * sk_ld);
        }
    }
}
bool ga_head_supported(int H, int W, int N, const Act& t, const Act& sk) {
    return (H % 2) == 0 && (W % 2) == 0 && N >= 8 && (N % 8) == 0 && N <= 768 && (t.ld % 8) == 0 && (sk.ld % 8) == 0 &&
           ((uintptr_t)t.p % 16) == 0 && ((uintptr_t)sk.p % 16) == 0;
}
void launch_ga_head(const float* x, int B, int H, int W, const float* dw9, const float* dwb, const float* w1, const float* b1,
                    const float* wsk, const float* bsk, int N, const Act& t, const Act& sk, cudaStream_t s) {
    const int Ho = H / 2, Wo = W / 2;
    const long long tiles = (long long)B * Ho * ((Wo + GH_PIX - 1) / GH_PIX);
    if (tiles == 0) return;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const long long blocks = std::min<long long>(tiles, (long long)sms * 3);
    launch_k(ga_head_kernel, dim3((unsigned)blocks), dim3(GH_THREADS), 0, s, x, H, W, dw9, dwb, w1, b1, wsk, bsk, N, (bf16*)t.p, t.ld, (bf16*)sk.p, sk.ld, (int)tiles);
}

// ------------------------------------------------------------------------------------------
// Depthwise 3x3 (pad 1, stride 1|2) + bias (+ GELU), NHWC.  HBM-bound: 4 channels per thread,
// channel-fastest thread order so every warp access is one contiguous 128..512 B segment; the
// 9-tap reuse is served by L1/L2.  (modules/layers/conv.py:49-54)
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T* __restrict__ in, int B, int H, int W, int C, int ild,
                                                        T* __restrict__ out, int Ho, int Wo, int old,
                                                        const float* __restrict__ w9, const float* __restrict__ bias,
                                                        int stride, int act) {
    pdl_wait();
    const int cg = (C + VEC - 1) / VEC;
    const long long total = (long long)B * Ho * Wo * cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % cg) * VEC;
        long long p = i / cg;
        int ow = (int)(p % Wo);
        long long q = p / Wo;
        int oh = (int)(q % Ho);
        int b = (int)(q / Ho);
        float a[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int ih = oh * stride - 1 + ky;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int iw = ow * stride - 1 + kx;
                if (iw < 0 || iw >= W) continue;
                const T* ip = in + (((size_t)b * H + ih) * W + iw) * ild + c;
                const float* wp = w9 + (ky * 3 + kx) * C + c;
                if (VEC == 4) {
                    float x[4];
                    load4(ip, x);
                    float4 wv = *reinterpret_cast<const float4*>(wp);
                    a[0] = fmaf(x[0], wv.x, a[0]);
                    a[1 % VEC] = fmaf(x[1 % VEC], wv.y, a[1 % VEC]);
                    a[2 % VEC] = fmaf(x[2 % VEC], wv.z, a[2 % VEC]);
                    a[3 % VEC] = fmaf(x[3 % VEC], wv.w, a[3 % VEC]);
                } else {
                    a[0] = fmaf(to_f<T>(ip[0]), wp[0], a[0]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            a[j] += bias[c + j];
            if (act == ACT_GELU) a[j] = gelu_erf(a[j]);
        }
        T* op = out + (((size_t)b * Ho + oh) * Wo + ow) * old + c;
        if (VEC == 4) store4(op, a);
        else op[0] = from_f<T>(a[0]);
    }
}

// bf16 fast path: one block stages a (TH*S+2) x (TW*S+2) pixel x 64-channel input patch in shared memory with
// 16-byte loads (8 lanes per pixel = one full 128-byte line), then every thread produces 8 channels of one output
// pixel per step from shared memory: the 9-tap reuse never goes back to L2, and every global access is a full line.
template <int S>
struct DwTile { static constexpr int TH = S == 1 ? 8 : 4, TW = S == 1 ? 32 : 16, IH = TH * S + 2, IW = TW * S + 2; };

template <int S>
__global__ void __launch_bounds__(256, 2) dwconv3x3_tiled_kernel(const bf16* __restrict__ in, int H, int W, int C, int ild,
                                                              bf16* __restrict__ out, int Ho, int Wo, int old,
                                                              const float* __restrict__ w9, const float* __restrict__ bias,
                                                              int act, int tilesW, int tilesH) {
    pdl_wait();
    using TT = DwTile<S>;
    __shared__ __align__(16) bf16 sIn[TT::IH * TT::IW * 64];
    const int cb = blockIdx.y * 64;                 // channel block
    const int b = blockIdx.z;
    const int th = blockIdx.x / tilesW, tw = blockIdx.x - th * tilesW;
    const int oh0 = th * TT::TH, ow0 = tw * TT::TW;
    const int ih0 = oh0 * S - 1, iw0 = ow0 * S - 1;
    const int cg = threadIdx.x & 7;                 // 8-channel group inside the 64-channel block
    const int c = cb + cg * 8;
    const bool cok = c < C;                          // C % 8 == 0 (host-checked)
    // stage the input patch (zero outside the image / channel range)
    for (int i = threadIdx.x >> 3; i < TT::IH * TT::IW; i += 32) {
        const int py = i / TT::IW, px = i - py * TT::IW;
        const int ih = ih0 + py, iw = iw0 + px;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (cok && ih >= 0 && ih < H && iw >= 0 && iw < W)
            v = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + ih) * W + iw) * ild + c);
        *reinterpret_cast<uint4*>(sIn + (size_t)i * 64 + cg * 8) = v;
    }
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        if (cok) { w0 = *reinterpret_cast<const float4*>(w9 + t * C + c); w1 = *reinterpret_cast<const float4*>(w9 + t * C + c + 4); }
        wr[t][0] = w0.x; wr[t][1] = w0.y; wr[t][2] = w0.z; wr[t][3] = w0.w;
        wr[t][4] = w1.x; wr[t][5] = w1.y; wr[t][6] = w1.z; wr[t][7] = w1.w;
    }
    {
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (cok) { b0 = *reinterpret_cast<const float4*>(bias + c); b1 = *reinterpret_cast<const float4*>(bias + c + 4); }
        br[0] = b0.x; br[1] = b0.y; br[2] = b0.z; br[3] = b0.w; br[4] = b1.x; br[5] = b1.y; br[6] = b1.z; br[7] = b1.w;
    }
    __syncthreads();
    if (!cok) return;
    for (int p = threadIdx.x >> 3; p < TT::TH * TT::TW; p += 32) {
        const int oy = p / TT::TW, ox = p - oy * TT::TW;
        const int oh = oh0 + oy, ow = ow0 + ox;
        if (oh >= Ho || ow >= Wo) continue;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const uint4 t = *reinterpret_cast<const uint4*>(sIn + (size_t)((oy * S + ky) * TT::IW + ox * S + kx) * 64 + cg * 8);
                const uint32_t wv[4] = {t.x, t.y, t.z, t.w};
                const float* wq = wr[ky * 3 + kx];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a[2 * k] = fmaf(__uint_as_float(wv[k] << 16), wq[2 * k], a[2 * k]);
                    a[2 * k + 1] = fmaf(__uint_as_float(wv[k] & 0xffff0000u), wq[2 * k + 1], a[2 * k + 1]);
                }
            }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v0 = a[2 * k] + br[2 * k], v1 = a[2 * k + 1] + br[2 * k + 1];
            if (act == ACT_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
            o[k] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + oh) * Wo + ow) * old + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// bf16 fast path, second generation (same scheme as the fused depthwise producer of gemm_tc.cu): lane = channel pair
// (one bf16x2 word), so every shared-memory access of a warp is one conflict-free 128-byte pixel row and every global
// store of a warp is one full 128-byte line; warp pw owns output columns 2pw, 2pw+1 of the tile and walks the input rows
// once, feeding packed fp32x2 FMAs against the 9 taps held in 18 registers.  ~64 registers: 4+ blocks per SM.
template <int S>
struct DwLane { static constexpr int TH = S == 1 ? 8 : 4, TW = 16, IH = TH * S + 2, IW = TW * S + 2, NX = 2 * S + 2 - (S - 1); };   // NX: 4 | 5

template <int S>
__global__ void __launch_bounds__(256, 3) dwconv3x3_lane2_kernel(const bf16* __restrict__ in, int H, int W, int C, int ild,
                                                              bf16* __restrict__ out, int Ho, int Wo, int old,
                                                              const float* __restrict__ w9, const float* __restrict__ bias,
                                                              int act, int tilesW) {
    pdl_wait();
    using TT = DwLane<S>;
    __shared__ __align__(16) uint32_t sIn[TT::IH * TT::IW * 32];
    const int cb = blockIdx.y * 64;
    const int b = blockIdx.z;
    const int th = blockIdx.x / tilesW, tw = blockIdx.x - th * tilesW;
    const int oh0 = th * TT::TH, ow0 = tw * TT::TW;
    const int ih0 = oh0 * S - 1, iw0 = ow0 * S - 1;
    {
        const int cg = threadIdx.x & 7;
        const bool cok = cb + cg * 8 < C;                  // C % 8 == 0 (host-checked)
        for (int i = threadIdx.x >> 3; i < TT::IH * TT::IW; i += 32) {
            const int py = i / TT::IW, px = i - py * TT::IW;
            const int ih = ih0 + py, iw = iw0 + px;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cok && ih >= 0 && ih < H && iw >= 0 && iw < W)
                v = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + ih) * W + iw) * ild + cb + cg * 8);
            *reinterpret_cast<uint4*>(sIn + (size_t)i * 32 + cg * 4) = v;
        }
    }
    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = cb + 2 * lane;
    const bool cok = c < C;
    float2 w2[9], b2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) w2[t] = cok ? make_float2(w9[t * C + c], w9[t * C + c + 1]) : make_float2(0.f, 0.f);
    if (cok) b2 = make_float2(bias[c], bias[c + 1]);
    __syncthreads();
    float2 acc[TT::TH][2];
#pragma unroll
    for (int oy = 0; oy < TT::TH; ++oy) { acc[oy][0] = b2; acc[oy][1] = b2; }
    const uint32_t* rp = sIn + (size_t)(2 * pw * S) * 32 + lane;
#pragma unroll
    for (int iy = 0; iy < TT::IH; ++iy) {
        float2 x[TT::NX];
#pragma unroll
        for (int j = 0; j < TT::NX; ++j) x[j] = bf2_to_f2(rp[(size_t)(iy * TT::IW + j) * 32]);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            if ((iy - ky) >= 0 && ((iy - ky) % S) == 0 && (iy - ky) / S < TT::TH) {
                const int oy = (iy - ky) / S;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc[oy][cc] = __ffma2_rn(x[cc * S + kx], w2[ky * 3 + kx], acc[oy][cc]);
            }
        }
        if (iy >= 2 && ((iy - 2) % S) == 0) {
            const int oy = (iy - 2) / S;
            const int oh = oh0 + oy;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int ow = ow0 + 2 * pw + cc;
                float2 v = acc[oy][cc];
                if (act == ACT_GELU) v = gelu2(v);
                if (cok && oh < Ho && ow < Wo) {
                    __nv_bfloat162 hv = __floats2bfloat162_rn(v.x, v.y);
                    *reinterpret_cast<uint32_t*>(out + (((size_t)b * Ho + oh) * Wo + ow) * old + c) = *reinterpret_cast<uint32_t*>(&hv);
                }
            }
        }
    }
}

void launch_dwconv3x3(int bf, const Act& in, const Act& out, const float* w9, const float* bias, int stride, int act,
                      cudaStream_t s) {
    static const int dw_mode = getenv("MLIC_DW") ? atoi(getenv("MLIC_DW")) : 2;      // development: 0 tiled, 1 lane-pair, 2 TMA-fed (default)
    const bool old_dw = dw_mode == 0;
    if (dw_mode == 2 && bf && out.B > 0 && out.H > 0 && out.W > 0 && launch_dwconv3x3_tma(in, out, w9, bias, stride, act, s) == 0) return;
    if (!old_dw && bf && (in.C % 8 == 0) && (in.ld % 8 == 0) && (out.ld % 2 == 0) && (((uintptr_t)in.p) % 16 == 0) &&
        (((uintptr_t)out.p) % 4 == 0) && (stride == 1 || stride == 2) && out.B > 0 && out.H > 0 && out.W > 0) {
        if (stride == 1) {
            const int tw = cdiv(out.W, DwLane<1>::TW), th = cdiv(out.H, DwLane<1>::TH);
            dim3 grid(tw * th, cdiv(in.C, 64), out.B);
            launch_k(dwconv3x3_lane2_kernel<1>, dim3(grid), dim3(256), 0, s, (const bf16*)in.p, in.H, in.W, in.C, in.ld, (bf16*)out.p, out.H, out.W,
                                                           out.ld, w9, bias, act, tw);
        } else {
            const int tw = cdiv(out.W, DwLane<2>::TW), th = cdiv(out.H, DwLane<2>::TH);
            dim3 grid(tw * th, cdiv(in.C, 64), out.B);
            launch_k(dwconv3x3_lane2_kernel<2>, dim3(grid), dim3(256), 0, s, (const bf16*)in.p, in.H, in.W, in.C, in.ld, (bf16*)out.p, out.H, out.W,
                                                           out.ld, w9, bias, act, tw);
        }
        return;
    }
    if (bf && (in.C % 8 == 0) && (in.ld % 8 == 0) && (out.ld % 8 == 0) && (((uintptr_t)in.p) % 16 == 0) &&
        (((uintptr_t)out.p) % 16 == 0) && (stride == 1 || stride == 2) && out.B > 0 && out.H > 0 && out.W > 0) {
        if (stride == 1) {
            const int tw = cdiv(out.W, DwTile<1>::TW), th = cdiv(out.H, DwTile<1>::TH);
            dim3 grid(tw * th, cdiv(in.C, 64), out.B);
            launch_k(dwconv3x3_tiled_kernel<1>, dim3(grid), dim3(256), 0, s, (const bf16*)in.p, in.H, in.W, in.C, in.ld, (bf16*)out.p, out.H, out.W,
                                                           out.ld, w9, bias, act, tw, th);
        } else {
            const int tw = cdiv(out.W, DwTile<2>::TW), th = cdiv(out.H, DwTile<2>::TH);
            dim3 grid(tw * th, cdiv(in.C, 64), out.B);
            launch_k(dwconv3x3_tiled_kernel<2>, dim3(grid), dim3(256), 0, s, (const bf16*)in.p, in.H, in.W, in.C, in.ld, (bf16*)out.p, out.H, out.W,
                                                           out.ld, w9, bias, act, tw, th);
        }
        return;
    }
    const int esz = bf ? 2 : 4;
    bool vec = (in.C % 4 == 0) && (in.ld % 4 == 0) && (out.ld % 4 == 0) &&
               (((uintptr_t)in.p) % (4 * esz) == 0) && (((uintptr_t)out.p) % (4 * esz) == 0);
    long long total = (long long)out.B * out.H * out.W * (vec ? in.C / 4 : in.C);
    if (total == 0) return;
    int blocks = (int)std::min<long long>(cdiv(total, 256), 148LL * 32);
#define DW_LAUNCH(T, V)                                                                                             \
    launch_k(dwconv3x3_kernel<T, V>, dim3(blocks), dim3(256), 0, s, (const T*)in.p, in.B, in.H, in.W, in.C, in.ld, (T*)out.p, out.H, \
                                                  out.W, out.ld, w9, bias, stride, act)
    if (bf) { if (vec) DW_LAUNCH(bf16, 4); else DW_LAUNCH(bf16, 1); }
    else { if (vec) DW_LAUNCH(float, 4); else DW_LAUNCH(float, 1); }
#undef DW_LAUNCH
}

// ------------------------------------------------------------------------------------------
// Layout conversion.  NCHW fp32 <-> NHWC activations through a 32x32 shared-memory transpose so
// both sides are coalesced.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int HW, T* __restrict__ dst, int ld) {
    pdl_wait();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? src[((size_t)b * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < C) dst[((size_t)b * HW + p) * ld + c] = from_f<T>(tile[threadIdx.x][j]);
    }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int ld, int C, int HW, float* __restrict__ dst) {
    pdl_wait();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int p = p0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (p < HW && c < C) ? to_f<T>(src[((size_t)b * HW + p) * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, p = p0 + threadIdx.x;
        if (c < C && p < HW) dst[((size_t)b * C + c) * HW + p] = tile[threadIdx.x][j];
    }
}

void launch_nchw_to_nhwc(int bf, const float* src, const Act& dst, int Csrc, cudaStream_t s) {
    int HW = dst.H * dst.W;
    if (HW == 0 || dst.B == 0) return;
    dim3 grid(cdiv(HW, 32), cdiv(Csrc, 32), dst.B), blk(32, 8);
    if (bf) launch_k(nchw_to_nhwc_kernel<bf16>, dim3(grid), dim3(blk), 0, s, src, Csrc, HW, (bf16*)dst.p, dst.ld);
    else launch_k(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(blk), 0, s, src, Csrc, HW, (float*)dst.p, dst.ld);
}
void launch_nhwc_to_nchw(int bf, const Act& src, float* dst, cudaStream_t s) {
    int HW = src.H * src.W;
    if (HW == 0 || src.B == 0) return;
    dim3 grid(cdiv(HW, 32), cdiv(src.C, 32), src.B), blk(32, 8);
    if (bf) launch_k(nhwc_to_nchw_kernel<bf16>, dim3(grid), dim3(blk), 0, s, (const bf16*)src.p, src.ld, src.C, HW, dst);
    else launch_k(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(blk), 0, s, (const float*)src.p, src.ld, src.C, HW, dst);
}
void launch_nhwc_f32_to_nchw(const float* src, int ld, int B, int H, int W, int C, float* dst, cudaStream_t s) {
    int HW = H * W;
    if (HW == 0 || B == 0) return;
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), B), blk(32, 8);
    launch_k(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(blk), 0, s, src, ld, C, HW, dst);
}

template <typename TS, typename TD>
__global__ void copy_channels_kernel(const TS* __restrict__ src, int sld, TD* __restrict__ dst, int dld, int C,
                                     long long npix) {
    pdl_wait();
    long long total = npix * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i / C;
        int c = (int)(i - p * C);
        dst[p * dld + c] = from_f<TD>(to_f<TS>(src[p * sld + c]));
    }
}
// 8 channels per thread (16-byte accesses), 32-bit index arithmetic: the scalar kernel above ran at 1.5-2.3 TB/s on the two layout copies of
// the walk (y -> bf16 slots, hyper_means -> LRP input prefix)
__device__ __forceinline__ void cc_load8(const float* p, float v[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void cc_load8(const bf16* p, float v[8]) { load8_bf16(p, v); }
__device__ __forceinline__ void cc_store8(float* p, const float v[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void cc_store8(bf16* p, const float v[8]) { *reinterpret_cast<uint4*>(p) = pack8_bf16(v); }
template <typename TS, typename TD>
__global__ void copy_channels_vec8_kernel(const TS* __restrict__ src, int sld, TD* __restrict__ dst, int dld, int C8, int total) {
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int p = i / C8, c = (i - p * C8) * 8;
        float v[8];
        cc_load8(src + (size_t)p * sld + c, v);
        cc_store8(dst + (size_t)p * dld + c, v);
    }
}
template <typename TS, typename TD>
static bool copy_channels_vec8(const TS* src, int sld, TD* dst, int dld, int C, long long npix, cudaStream_t s) {
    if ((C % 8) || (sld % 8) || (dld % 8) || ((uintptr_t)src % 16) || ((uintptr_t)dst % 16) || npix * (C / 8) >= 0x7fffffffLL - 148 * 16 * 256) return false;
    const int total = (int)(npix * (C / 8));
    const int blocks = (int)std::min<long long>(cdiv((long long)total, 256), 148LL * 16);
    launch_k(copy_channels_vec8_kernel<TS, TD>, dim3(blocks), dim3(256), 0, s, src, sld, dst, dld, C / 8, total);
    return true;
}
void launch_copy_channels(int bf, const Act& src, const Act& dst, cudaStream_t s) {
    long long npix = (long long)src.B * src.H * src.W;
    if (npix * src.C == 0) return;
    if (bf ? copy_channels_vec8((const bf16*)src.p, src.ld, (bf16*)dst.p, dst.ld, src.C, npix, s)
           : copy_channels_vec8((const float*)src.p, src.ld, (float*)dst.p, dst.ld, src.C, npix, s)) return;
    int blocks = (int)std::min<long long>(cdiv(npix * src.C, 256), 148LL * 16);
    if (bf) launch_k(copy_channels_kernel<bf16, bf16>, dim3(blocks), dim3(256), 0, s, (const bf16*)src.p, src.ld, (bf16*)dst.p, dst.ld, src.C, npix);
    else launch_k(copy_channels_kernel<float, float>, dim3(blocks), dim3(256), 0, s, (const float*)src.p, src.ld, (float*)dst.p, dst.ld, src.C, npix);
}
void launch_copy_f32_to_act(int bf, const float* src, int ld, const Act& dst, cudaStream_t s) {
    long long npix = (long long)dst.B * dst.H * dst.W;
    if (npix * dst.C == 0) return;
    if (bf ? copy_channels_vec8(src, ld, (bf16*)dst.p, dst.ld, dst.C, npix, s) : copy_channels_vec8(src, ld, (float*)dst.p, dst.ld, dst.C, npix, s)) return;
    int blocks = (int)std::min<long long>(cdiv(npix * dst.C, 256), 148LL * 16);
    if (bf) launch_k(copy_channels_kernel<float, bf16>, dim3(blocks), dim3(256), 0, s, src, ld, (bf16*)dst.p, dst.ld, dst.C, npix);
    else launch_k(copy_channels_kernel<float, float>, dim3(blocks), dim3(256), 0, s, src, ld, (float*)dst.p, dst.ld, dst.C, npix);
}
void launch_fill_zero(void* p, size_t bytes, cudaStream_t s) { if (bytes) cudaMemsetAsync(p, 0, bytes, s); }

// ------------------------------------------------------------------------------------------
// EntropyBottleneck (CompressAI; call site models/mlicpp.py:96-98), eval mode.
//   z_hat = round(z - med) + med ;  lik = sigmoid(L(z_hat+.5)) - sigmoid(L(z_hat-.5)), floored at 1e-9
//   L: 1->3->3->3->3->1 chain, softplus(matrix) and tanh(factor) folded at pack time.
// packed[c] = { M0[3] , M1[9], M2[9], M3[9], M4[3], b0[3], b1[3], b2[3], b3[3], b4[1], f0[3], f1[3], f2[3], f3[3] }
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float eb_logits(const float* P, float x) {
    float v[3], t[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float l = P[j] * x + P[33 + j];
        v[j] = l + P[46 + j] * tanhf(l);
    }
#pragma unroll
    for (int layer = 1; layer < 4; ++layer) {
        const float* Mx = P + 3 + (layer - 1) * 9;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float l = Mx[j * 3 + 0] * v[0];
            l = fmaf(Mx[j * 3 + 1], v[1], l);
            l = fmaf(Mx[j * 3 + 2], v[2], l);
            l += P[33 + layer * 3 + j];
            t[j] = l + P[46 + layer * 3 + j] * tanhf(l);
        }
        v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
    }
    float l = P[30] * v[0];
    l = fmaf(P[31], v[1], l);
    l = fmaf(P[32], v[2], l);
    return l + P[45];
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T>
__global__ void entropy_bottleneck_kernel(const T* __restrict__ z, int zld, T* __restrict__ zh, int zhld, int C,
                                          int HW, long long total, const float* __restrict__ packed,
                                          const float* __restrict__ med, float* __restrict__ lik_nchw,
                                          int32_t* __restrict__ sym_nchw, float qs) {
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i / C;
        int c = (int)(i - p * C);
        int b = (int)(p / HW);
        int hw = (int)(p - (long long)b * HW);
        float m = med[c];
        // qs: quantisation step of the variable-rate hyper prior (EntropyBottleneckVbr, mlicpp_vbr.py:253-259); 1 for the plain bottleneck,
        // where the division, the product and the half step below are exact.  Product and sum rounded separately, as torch evaluates them.
        float q = rintf((to_f<T>(z[p * zld + c]) - m) / qs);
        float v = __fadd_rn(__fmul_rn(q, qs), m);
        if (zh) zh[p * zhld + c] = from_f<T>(v);
        size_t o = ((size_t)b * C + c) * HW + hw;
        if (sym_nchw) sym_nchw[o] = (int32_t)q;
        if (lik_nchw) {
            const float* P = packed + (size_t)c * 58;
            const float half = 0.5f * qs;
            float lo = eb_logits(P, v - half), up = eb_logits(P, v + half);
            float l = sigmoidf_(up) - sigmoidf_(lo);
            lik_nchw[o] = fmaxf(l, 1e-9f);
        }
    }
}
void launch_entropy_bottleneck(int bf, const Act& z, const Act& z_hat, const float* packed, const float* medians,
                               float* z_lik_nchw, int32_t* z_sym_nchw, cudaStream_t s, float qs) {
    long long total = (long long)z.B * z.H * z.W * z.C;
    if (!total) return;
    int blocks = cdiv(total, 128);
    if (bf) launch_k(entropy_bottleneck_kernel<bf16>, dim3(blocks), dim3(128), 0, s, (const bf16*)z.p, z.ld, (bf16*)z_hat.p, z_hat.ld, z.C, z.H * z.W, total, packed, medians, z_lik_nchw, z_sym_nchw, qs);
    else launch_k(entropy_bottleneck_kernel<float>, dim3(blocks), dim3(128), 0, s, (const float*)z.p, z.ld, (float*)z_hat.p, z_hat.ld, z.C, z.H * z.W, total, packed, medians, z_lik_nchw, z_sym_nchw, qs);
}

// EntropyBottleneck.decompress after the range decoder (CompressAI: dequantize(values, medians)): z_hat = sym + median
template <typename T>
__global__ void zsym_to_zhat_kernel(const int32_t* __restrict__ sym_nchw, T* __restrict__ zh, int zhld, int C, int HW,
                                    long long total, const float* __restrict__ med, float qs) {
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i / C;
        int c = (int)(i - p * C);
        int b = (int)(p / HW);
        int hw = (int)(p - (long long)b * HW);
        zh[p * zhld + c] = from_f<T>(__fadd_rn(__fmul_rn((float)sym_nchw[((size_t)b * C + c) * HW + hw], qs), med[c]));
    }
}
void launch_zsym_to_zhat(int bf, const int32_t* z_sym_nchw, const float* medians, const Act& z_hat, cudaStream_t s, float qs) {
    long long total = (long long)z_hat.B * z_hat.H * z_hat.W * z_hat.C;
    if (!total) return;
    int blocks = cdiv(total, 128);
    if (bf) launch_k(zsym_to_zhat_kernel<bf16>, dim3(blocks), dim3(128), 0, s, z_sym_nchw, (bf16*)z_hat.p, z_hat.ld, z_hat.C, z_hat.H * z_hat.W, total, medians, qs);
    else launch_k(zsym_to_zhat_kernel<float>, dim3(blocks), dim3(128), 0, s, z_sym_nchw, (float*)z_hat.p, z_hat.ld, z_hat.C, z_hat.H * z_hat.W, total, medians, qs);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the channel dim of an NHWC view (eps 1e-5, affine); one warp per pixel.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, int xld, int C, long long npix, const float* __restrict__ g,
                                 const float* __restrict__ b, T* __restrict__ out, int old) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = wid; p < npix; p += nw) {
        float v[4];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = lane + j * 32;
            v[j] = c < C ? to_f<T>(x[p * xld + c]) : 0.f;
            sum += v[j];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float mean = sum / C;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = lane + j * 32;
            float d = c < C ? v[j] - mean : 0.f;
            var += d * d;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        float rstd = rsqrtf(var / C + 1e-5f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = lane + j * 32;
            if (c < C) out[p * old + c] = from_f<T>((v[j] - mean) * rstd * g[c] + b[c]);
        }
    }
}
// bf16 fast mode, C = 8 * LP channels (32 | 64): LP lanes per pixel, 16 bytes each, 32 / LP pixels per warp step.  (The generic kernel
// above moves 2 bytes per lane and reduces over the whole warp: 51 us for 261 120 x 32 channels where the traffic takes 5.)
template <int LP>
__global__ void __launch_bounds__(256) layernorm_bf16_vec_kernel(const bf16* __restrict__ x, int xld, long long npix, const float* __restrict__ g,
                                                                 const float* __restrict__ b, bf16* __restrict__ out, int old) {
    pdl_wait();
    constexpr int C = 8 * LP, PPW = 32 / LP;
    const int lane = threadIdx.x & 31, sub = lane % LP, pl = lane / LP;
    float gg[8], bb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { gg[k] = g[sub * 8 + k]; bb[k] = b[sub * 8 + k]; }
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p0 = wid * PPW; p0 < npix; p0 += nw * PPW) {
        const long long p = p0 + pl;
        const bool ok = p < npix;
        float v[8];
        unpack8_bf16(ok ? *reinterpret_cast<const uint4*>(x + p * xld + sub * 8) : make_uint4(0, 0, 0, 0), v);
        float sum = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
        for (int o = LP / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.0f / C);
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] -= mean; var = fmaf(v[k], v[k], var); }
#pragma unroll
        for (int o = LP / 2; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var * (1.0f / C) + 1e-5f);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k] * rstd, gg[k], bb[k]);
        if (ok) *reinterpret_cast<uint4*>(out + p * old + sub * 8) = pack8_bf16(v);
    }
}
void launch_layernorm(int bf, const Act& x, const float* g, const float* b, const Act& out, cudaStream_t s) {
    long long npix = (long long)x.B * x.H * x.W;
    if (!npix) return;
    if (bf && (x.C == 32 || x.C == 64) && (x.ld % 8) == 0 && (out.ld % 8) == 0 && ((uintptr_t)x.p % 16) == 0 && ((uintptr_t)out.p % 16) == 0) {
        const int ppw = x.C == 32 ? 8 : 4;
        const int blocks = (int)std::min<long long>(cdiv(npix, 8LL * ppw), 148LL * 8);
        if (x.C == 32) launch_k(layernorm_bf16_vec_kernel<4>, dim3(blocks), dim3(256), 0, s, (const bf16*)x.p, x.ld, npix, g, b, (bf16*)out.p, out.ld);
        else launch_k(layernorm_bf16_vec_kernel<8>, dim3(blocks), dim3(256), 0, s, (const bf16*)x.p, x.ld, npix, g, b, (bf16*)out.p, out.ld);
        return;
    }
    int blocks = (int)std::min<long long>(cdiv(npix, 8), 148LL * 16);
    if (bf) launch_k(layernorm_kernel<bf16>, dim3(blocks), dim3(256), 0, s, (const bf16*)x.p, x.ld, x.C, npix, g, b, (bf16*)out.p, out.ld);
    else launch_k(layernorm_kernel<float>, dim3(blocks), dim3(256), 0, s, (const float*)x.p, x.ld, x.C, npix, g, b, (float*)out.p, out.ld);
}

// ------------------------------------------------------------------------------------------
// LocalContext windowed attention (modules/transform/context.py:80-107).
//   F[pix][3C] fp32: q = [0,C), k = [C,2C), v = [2C,3C); head split on the way in: channel c = dd*2 + hh.
//   For pixel l, head hh, window taps a,b in 5x5 (zero q/k/v outside the image):
//     A[a][b] = (q_a*scale).k_b + bias[hh][a][b] + mask(a,b);  mask = 0 iff taps a and b are both in-image anchors
//     else -100;  P = softmax_b(A);  o[a][hh*d + dd] = sum_b P[a][b] v_b[dd*2+hh]
//   O[pix][a][c] (activation type) feeds the `fusion` GEMM (K = 25*C).
// One block = 8x8 pixel tile with a 2-pixel halo of F staged in shared memory; one thread per
// (pixel, head, query tap) row.
// ------------------------------------------------------------------------------------------
constexpr int LT = 8;           // tile side
constexpr int LTH = LT + 4;     // with halo

template <typename T, int HD>
__global__ void __launch_bounds__(256) local_attn_kernel(const float* __restrict__ F, int H, int W,
                                                         const float* __restrict__ rel_bias, T* __restrict__ O) {
    pdl_wait();
    constexpr int C = 2 * HD;
    extern __shared__ float sm[];
    float* sF = sm;                          // [LTH*LTH][3C + 1]
    float* sB = sm + LTH * LTH * (3 * C + 1);   // [2][25][25]
    const int b = blockIdx.z;
    const int h0 = blockIdx.y * LT, w0 = blockIdx.x * LT;
    const int FS = 3 * C + 1;
    for (int i = threadIdx.x; i < LTH * LTH * 3 * C; i += blockDim.x) {
        int f = i % (3 * C);
        int pp = i / (3 * C);
        int hh = h0 - 2 + pp / LTH, ww = w0 - 2 + pp % LTH;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = F[(((size_t)b * H + hh) * W + ww) * (3 * C) + f];
        sF[pp * FS + f] = v;
    }
    for (int i = threadIdx.x; i < 2 * 625; i += blockDim.x) sB[i] = rel_bias[i];
    __syncthreads();
    const float scale = rsqrtf((float)HD);
    for (int row = threadIdx.x; row < LT * LT * 2 * 25; row += blockDim.x) {
        int a = row % 25;
        int t = row / 25;
        int hh = t & 1;
        int pl = t >> 1;
        int ph = pl / LT, pw = pl % LT;
        int gh = h0 + ph, gw = w0 + pw;
        if (gh >= H || gw >= W) continue;
        int ay = a / 5, ax = a % 5;
        int qh = gh + ay - 2, qw = gw + ax - 2;
        bool qanch = qh >= 0 && qh < H && qw >= 0 && qw < W && (((qh + qw) & 1) == 1);
        const float* qp = sF + ((ph + ay) * LTH + (pw + ax)) * FS;
        float q[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = qp[d * 2 + hh] * scale;
        float sc[25];
        float mx = -INFINITY;
#pragma unroll
        for (int bb = 0; bb < 25; ++bb) {
            int by = bb / 5, bx = bb % 5;
            const float* kp = sF + ((ph + by) * LTH + (pw + bx)) * FS + C;
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], kp[d * 2 + hh], s);
            int kh = gh + by - 2, kw = gw + bx - 2;
            bool kanch = kh >= 0 && kh < H && kw >= 0 && kw < W && (((kh + kw) & 1) == 1);
            s += sB[(hh * 25 + a) * 25 + bb];
            s += (qanch && kanch) ? 0.0f : -100.0f;
            sc[bb] = s;
            mx = fmaxf(mx, s);
        }
        float den = 0.f;
#pragma unroll
        for (int bb = 0; bb < 25; ++bb) {
            sc[bb] = expf(sc[bb] - mx);
            den += sc[bb];
        }
        float inv = 1.0f / den;
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
        for (int bb = 0; bb < 25; ++bb) {
            int by = bb / 5, bx = bb % 5;
            const float* vp = sF + ((ph + by) * LTH + (pw + bx)) * FS + 2 * C;
            float p = sc[bb] * inv;
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = fmaf(p, vp[d * 2 + hh], o[d]);
        }
        T* op = O + ((((size_t)b * H + gh) * W + gw) * 25 + a) * C + hh * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            float v4[4] = {o[d], o[d + 1], o[d + 2], o[d + 3]};
            store4(op + d, v4);
        }
    }
}
int launch_local_attn(int bf, const float* F, int B, int H, int W, int C, const float* rel_bias, void* O,
                      cudaStream_t s) {
    if (B * H * W == 0) return 0;
    dim3 grid(cdiv(W, LT), cdiv(H, LT), B);
    size_t smem = (size_t)(LTH * LTH * (3 * C + 1) + 2 * 625) * sizeof(float);
#define LA_LAUNCH(T, HD)                                                                                      \
    do {                                                                                                      \
        cudaFuncSetAttribute(local_attn_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        launch_k(local_attn_kernel<T, HD>, dim3(grid), dim3(256), smem, s, F, H, W, rel_bias, (T*)O);                           \
    } while (0)
    if (C == 32) { if (bf) LA_LAUNCH(bf16, 16); else LA_LAUNCH(float, 16); }
    else if (C == 64) { if (bf) LA_LAUNCH(bf16, 32); else LA_LAUNCH(float, 32); }
    else return 1;
#undef LA_LAUNCH
    return 0;
}

// ------------------------------------------------------------------------------------------
// LocalContext windowed attention, bf16 fast mode, on the legacy warp-level tensor path (mma.sync m16n8k16: the
// per-pixel problem is 25 x 25 x 16, far below a tcgen05 tile).  Same definition as local_attn_kernel above, with
//   * F bf16 [pix][3C] in HEAD-MAJOR channel order  q_h0 | q_h1 | k_h0 | k_h1 | v_h0 | v_h1  (16 each; the qkv_proj
//     weight rows are permuted at pack time), so the staged window is directly ldmatrix-addressable;
//   * only NON-ANCHOR pixels are evaluated: the LocalContext output feeds the per-pixel (1x1) non-anchor
//     EntropyParameters stack whose result is multiplied by the non-anchor mask (mlicpp.py:146-150), so the anchor half
//     of the output is never observed;
//   * O is written SQUEEZED, [B][H][W/2][25][C] with w = 2j + (h & 1) (utils/ckbd.py:47-59 order), and the whole
//     fusion/proj/MLP tail of LocalContext runs on that half-size matrix.
// One warp = one pixel: per head S = Q_w K_w^T (32x32 padded, 8 MMAs), + bias + mask, row softmax through quad
// shuffles, O = P V_w (8 MMAs).  Block = 8 x 16 pixel tile (64 non-anchor pixels, 8 per warp) with a 2-pixel halo.
// ------------------------------------------------------------------------------------------
constexpr int LM_TH = 8, LM_TW = 16, LM_HH = LM_TH + 4, LM_HW = LM_TW + 4;
constexpr int LM_PITCH = 208;                        // bytes per staged position: 96 bf16 + 16 B pad (conflict-free rows)
constexpr int LM_F_BYTES = LM_HH * LM_HW * LM_PITCH; // 49 920
constexpr int LM_BIAS_BYTES = 2 * 32 * 32 * 4;
constexpr int LM_OUT_BYTES = 25 * 32 * 2;            // one pixel's [25][C] block
constexpr int LM_SMEM = LM_F_BYTES + 16 + LM_BIAS_BYTES + 8 * LM_OUT_BYTES;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t r[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t r[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr float LM_LOG2E = 1.4426950408889634f;
__global__ void __launch_bounds__(256, 2) local_attn_mma_kernel(const bf16* __restrict__ F, int f_ld, int H, int W,
                                                             const float* __restrict__ rel_bias, bf16* __restrict__ O) {
    pdl_wait();
    extern __shared__ __align__(16) uint8_t lm_smem[];
    uint8_t* sF = lm_smem;
    uint8_t* sZero = lm_smem + LM_F_BYTES;                       // 16 zero bytes: rows of the padded taps 25..31
    float* sB = reinterpret_cast<float*>(lm_smem + LM_F_BYTES + 16);
    uint8_t* sO = lm_smem + LM_F_BYTES + 16 + LM_BIAS_BYTES;
    const int b = blockIdx.z, h0 = blockIdx.y * LM_TH, w0 = blockIdx.x * LM_TW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // stage the halo window: 12 chunks of 16 B per position, zeros outside the image (Unfold pads q/k/v, context.py:80-82)
    for (int i = threadIdx.x; i < LM_HH * LM_HW * 12; i += blockDim.x) {
        const int ch = i % 12, pp = i / 12;
        const int hh = h0 - 2 + pp / LM_HW, ww = w0 - 2 + pp % LM_HW;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = *reinterpret_cast<const uint4*>(F + (((size_t)b * H + hh) * W + ww) * f_ld + ch * 8);
        *reinterpret_cast<uint4*>(sF + pp * LM_PITCH + ch * 16) = v;
    }
    if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(sZero)[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) {
        const int col = i & 31, row = (i >> 5) & 31, hh = i >> 10;
        sB[i] = (col < 25) ? ((row < 25) ? LM_LOG2E * rel_bias[(hh * 25 + row) * 25 + col] : 0.0f) : -30000.0f;      // log2 domain (exp2 below)
    }
    __syncthreads();
    const uint32_t sF_s = (uint32_t)__cvta_generic_to_shared(sF), sZ_s = (uint32_t)__cvta_generic_to_shared(sZero);
    uint8_t* myO = sO + warp * LM_OUT_BYTES;
    const int g = lane >> 2, t = lane & 3;
    for (int it = 0; it < 8; ++it) {
        // non-anchor pixel `it` of this warp's tile row: row ph = warp, column pw = 2 it + ((h0 + warp + w0) & 1)
        const int ph = warp, gh = h0 + ph;
        const int pw = 2 * it + ((gh + w0) & 1), gw = w0 + pw;
        if (gh >= H || gw >= W) continue;                        // warp-uniform
        // window taps that are in-image anchors ((row + col) odd)
        bool anch = false;
        if (lane < 25) {
            const int qh = gh + lane / 5 - 2, qw = gw + lane % 5 - 2;
            anch = qh >= 0 && qh < H && qw >= 0 && qw < W && (((qh + qw) & 1) == 1);
        }
        const uint32_t am = __ballot_sync(0xffffffffu, anch);
        float cmask[4][2];                  // mask of this lane's 8 key columns when the query tap is an anchor: 0 | -100 (x log2 e)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) cmask[nt][e] = ((am >> (nt * 8 + 2 * (lane & 3) + e)) & 1u) ? 0.0f : -100.0f * LM_LOG2E;
        // per-lane ldmatrix row addresses (tap -> staged position), shared by both heads
        auto tap_addr = [&](int tap) -> uint32_t {
            return tap < 25 ? sF_s + (uint32_t)(((ph + tap / 5) * LM_HW + (pw + tap % 5)) * LM_PITCH) : 0u;
        };
        uint32_t qa[2], ka[2], va[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int qrow = 16 * i + (lane & 7) + ((lane >> 3) & 1) * 8;      // A: matrices (r0-7,k0-7) (r8-15,k0-7) (r0-7,k8-15) (r8-15,k8-15)
            uint32_t a0 = tap_addr(qrow);
            qa[i] = a0 ? a0 + (uint32_t)((lane >> 4) * 16) : sZ_s;
            const int krow = 16 * i + (lane & 7) + (lane >> 4) * 8;            // B: (n0-7,k0-7) (n0-7,k8-15) (n8-15,k0-7) (n8-15,k8-15)
            a0 = tap_addr(krow);
            ka[i] = a0 ? a0 + (uint32_t)(64 + ((lane >> 3) & 1) * 16) : sZ_s;
            const int vrow = 16 * i + (lane & 7) + ((lane >> 3) & 1) * 8;      // B^T: (k0-7,n0-7) (k8-15,n0-7) (k0-7,n8-15) (k8-15,n8-15)
            a0 = tap_addr(vrow);
            va[i] = a0 ? a0 + (uint32_t)(128 + (lane >> 4) * 16) : sZ_s;
        }
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
            const uint32_t hoff = (uint32_t)(hh * 32);
            uint32_t qf[2][4], kf[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                ldsm_x4(qa[i] == sZ_s ? sZ_s : qa[i] + hoff, qf[i]);
                ldsm_x4(ka[i] == sZ_s ? sZ_s : ka[i] + hoff, kf[i]);
            }
            float sc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) sc[mt][nt][j] = 0.f;
                    mma_bf16_16816(sc[mt][nt], qf[mt], kf[nt >> 1][(nt & 1) * 2], kf[nt >> 1][(nt & 1) * 2 + 1]);
                }
            // scale, + bias, + mask; row softmax (a row lives in the 4 lanes of a quad).  Everything in the log2 domain (scale and bias
            // pre-multiplied by log2 e: exp2 is one MUFU after one subtraction), the column part of the mask is 8 floats per lane and
            // pixel (cmask, shared by both heads), and P stays un-normalised -- 1 / sum is applied to the 16 x 16 output instead of
            // the 32 x 32 probabilities.
            const float* bh = sB + hh * 1024;
            float inv_row[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = mt * 16 + half * 8 + g;
                    const bool qan = (am >> row) & 1u;
                    float mx = -3.0e38f;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = nt * 8 + 2 * t;
                        const float2 bb = *reinterpret_cast<const float2*>(bh + row * 32 + col);
                        const float s0 = fmaf(sc[mt][nt][half * 2], 0.25f * LM_LOG2E, bb.x) + (qan ? cmask[nt][0] : -100.0f * LM_LOG2E);
                        const float s1 = fmaf(sc[mt][nt][half * 2 + 1], 0.25f * LM_LOG2E, bb.y) + (qan ? cmask[nt][1] : -100.0f * LM_LOG2E);
                        sc[mt][nt][half * 2] = s0; sc[mt][nt][half * 2 + 1] = s1;
                        mx = fmaxf(mx, fmaxf(s0, s1));
                    }
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    float den = 0.f;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        float e0, e1;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(sc[mt][nt][half * 2] - mx));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(sc[mt][nt][half * 2 + 1] - mx));
                        sc[mt][nt][half * 2] = e0; sc[mt][nt][half * 2 + 1] = e1;
                        den += e0 + e1;
                    }
                    den += __shfl_xor_sync(0xffffffffu, den, 1);
                    den += __shfl_xor_sync(0xffffffffu, den, 2);
                    inv_row[mt][half] = 1.0f / den;
                }
            // O = P V: k blocks of 16 key taps; P accumulators re-packed as bf16 A fragments
            uint32_t vf[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) ldsm_x4_trans(va[i] == sZ_s ? sZ_s : va[i] + hoff, vf[i]);
            float oc[2][2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int dn = 0; dn < 2; ++dn) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) oc[mt][dn][j] = 0.f;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        uint32_t pf[4];
                        pf[0] = pack_bf2(sc[mt][2 * kb][0], sc[mt][2 * kb][1]);
                        pf[1] = pack_bf2(sc[mt][2 * kb][2], sc[mt][2 * kb][3]);
                        pf[2] = pack_bf2(sc[mt][2 * kb + 1][0], sc[mt][2 * kb + 1][1]);
                        pf[3] = pack_bf2(sc[mt][2 * kb + 1][2], sc[mt][2 * kb + 1][3]);
                        mma_bf16_16816(oc[mt][dn], pf, vf[kb][dn * 2], vf[kb][dn * 2 + 1]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) oc[mt][dn][j] *= inv_row[mt][j >> 1];
                }
            // rows a < 25 -> staged [25][C] block, channel = hh*16 + d  (context.py:106-107)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = mt * 16 + half * 8 + g;
                    if (row < 25) {
#pragma unroll
                        for (int dn = 0; dn < 2; ++dn)
                            *reinterpret_cast<uint32_t*>(myO + row * 64 + (hh * 16 + dn * 8 + 2 * t) * 2) =
                                pack_bf2(oc[mt][dn][half * 2], oc[mt][dn][half * 2 + 1]);
                    }
                }
        }
        __syncwarp();
        bf16* op = O + ((((size_t)b * H + gh) * (W >> 1)) + (gw >> 1)) * (25 * 32);
        for (int i = lane; i < LM_OUT_BYTES / 16; i += 32)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(op) + i * 16) = *reinterpret_cast<const uint4*>(myO + i * 16);
        __syncwarp();
    }
}
// F: bf16 NHWC [B,H,W,>=96] (head-major q|k|v), O: bf16 [B*H*W/2][25*32]; C must be 32, W even.
int launch_local_attn_mma(const Act& F, const float* rel_bias, void* O, cudaStream_t s) {
    if (F.C != 96 || (F.W & 1) || (F.ld % 8) != 0 || ((uintptr_t)F.p % 16) != 0 || ((uintptr_t)O % 16) != 0) return 1;
    if (F.B * F.H * F.W == 0) return 0;
    static bool attr[64] = {};           // the attribute belongs to the current device, not to the process
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) { cudaFuncSetAttribute(local_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM); attr[dev] = true; }
    dim3 grid(cdiv(F.W, LM_TW), cdiv(F.H, LM_TH), F.B);
    launch_k(local_attn_mma_kernel, dim3(grid), dim3(256), LM_SMEM, s, (const bf16*)F.p, F.ld, F.H, F.W, rel_bias, (bf16*)O);
    return 0;
}

// squeezed non-anchor rows [B][H][W/2][C] -> full NHWC view (non-anchor pixels; anchor pixels are zeroed)
__global__ void unsqueeze_nonanchor_kernel(const bf16* __restrict__ src, int sld, bf16* __restrict__ dst, int dld, int H, int W,
                                           int C8, long long total) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        const long long pix = i / C8;
        const int w = (int)(pix % W);
        const long long bh = pix / W;
        const int h = (int)(bh % H);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (((h + w) & 1) == 0) v = *reinterpret_cast<const uint4*>(src + (bh * (W >> 1) + (w >> 1)) * sld + c * 8);
        *reinterpret_cast<uint4*>(dst + pix * dld + c * 8) = v;
    }
}
void launch_unsqueeze_nonanchor(const Act& src /*[1,1,B*H*W/2,C]*/, const Act& dst /*[B,H,W,C]*/, cudaStream_t s) {
    const long long total = (long long)dst.B * dst.H * dst.W * (dst.C / 8);
    if (total == 0) return;
    int blocks = (int)std::min<long long>(cdiv(total, 256), 148LL * 8);
    launch_k(unsqueeze_nonanchor_kernel, dim3(blocks), dim3(256), 0, s, (const bf16*)src.p, src.ld, (bf16*)dst.p, dst.ld, dst.H, dst.W, dst.C / 8, total);
}

// ------------------------------------------------------------------------------------------
// Linear (kernelised) global attention (context.py:179-188 intra, :234-240 inter).
//   qkv NHWC view: Q = [0,D), K = [D,2D), V = [2D,3D); head g owns channels g*hd..g*hd+hd-1.
//   Khat = softmax over positions (only positions of parity par_kv when set), Qhat = softmax over the hd head
//   channels per position, ctx[g] = Khat V^T (hd x hd), out = ctx^T Qhat (positions of parity par_q, else 0).
// Three deterministic passes over a fixed chunking of the positions:
//   1. per-chunk column max of K                      -> pmax[B][nch][D]
//   2. per-chunk sum exp(K-max), sum exp(K-max) V^T   -> pctx[B][heads][nch][hd*hd + hd]
//   3. reduce chunks (fixed order), normalise, apply to softmax_c(Q)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8_bf16_k(const uint4 t, float v[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
constexpr int LA_CH = 64;   // max chunks
// Position chunks per (image, head).  The count depends on the image size only, so the (fixed-order) reduction of the chunk
// partials is identical whatever the batch size: results are batch-invariant bit for bit.
static inline int lin_chunks(int HW) { int n = (HW + 255) / 256; return n < 1 ? 1 : (n > LA_CH ? LA_CH : n); }

size_t lin_attn_scratch_floats(int B, int heads, int hd, int HW) {
    int nch = lin_chunks(HW);
    size_t D = (size_t)heads * hd;
    return (size_t)B * nch * D + (size_t)B * heads * nch * (hd * hd + hd) + (size_t)B * heads * (hd * hd);
}

template <typename T>
__global__ void __launch_bounds__(256) lin_colmax_kernel(const T* __restrict__ qkv, int ld, int D, int H, int W, int nch,
                                                         int par, float* __restrict__ pmax) {
    pdl_wait();
    // block = (chunk, 32-channel group, b); thread = (position lane 0..7, channel 0..31)
    __shared__ float sm[8][33];
    const int ch = blockIdx.x, b = blockIdx.z;
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    const int HW = H * W;
    const int per = (HW + nch - 1) / nch;
    const int p0 = ch * per, p1 = min(HW, p0 + per);
    if constexpr (sizeof(T) == 2) {
        // bf16: 16-byte loads -- thread = (64 position lanes, 4 channel octets of the 32-channel group); the scalar form below moved 64 bytes
        // per warp instruction and ran at a third of the HBM rate
        if ((ld & 7) == 0 && (D & 31) == 0 && ((uintptr_t)qkv & 15) == 0) {
            __shared__ float sv[64][33];
            const int oc = threadIdx.x & 3, pv = threadIdx.x >> 2;
            const int c8 = blockIdx.y * 32 + oc * 8;
            float mv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) mv[j] = -INFINITY;
            for (int p = p0 + pv; p < p1; p += 64) {
                if (par != PAR_NONE) {
                    int h = p / W, w = p - h * W;
                    if (!parity_keep(par, h, w)) continue;
                }
                float v[8];
                load8_bf16(reinterpret_cast<const bf16*>(qkv) + ((size_t)b * HW + p) * ld + D + c8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) mv[j] = fmaxf(mv[j], v[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[pv][oc * 8 + j] = mv[j];
            __syncthreads();
            if (threadIdx.x < 32) {
                float m = sv[0][threadIdx.x];
                for (int k = 1; k < 64; ++k) m = fmaxf(m, sv[k][threadIdx.x]);
                pmax[((size_t)b * nch + ch) * D + blockIdx.y * 32 + threadIdx.x] = m;
            }
            return;
        }
    }
    float m = -INFINITY;
    if (c < D) {
        for (int p = p0 + pl; p < p1; p += 8) {
            if (par != PAR_NONE) {
                int h = p / W, w = p - h * W;
                if (!parity_keep(par, h, w)) continue;
            }
            m = fmaxf(m, to_f<T>(qkv[((size_t)b * HW + p) * ld + D + c]));
        }
    }
    sm[pl][cl] = m;
    __syncthreads();
    if (pl == 0 && c < D) {
#pragma unroll
        for (int k = 1; k < 8; ++k) m = fmaxf(m, sm[k][cl]);
        pmax[((size_t)b * nch + ch) * D + c] = m;
    }
}

// block = (chunk, head, b); 256 threads.  TP positions are staged per step (bf16: 16-byte loads, all of a step's loads in
// flight together, so a 255-position chunk is two latency rounds instead of eight); thread t then owns PER_T entries
// (c1, c2) of the hd x hd context and sums exp(K[p][c1] - max) * V[p][c2] over the staged positions in position order.
template <typename T, int HD>
__global__ void __launch_bounds__(256) lin_ctx_kernel(const T* __restrict__ qkv, int ld, int D, int H, int W, int nch,
                                                      int par, const float* __restrict__ pmax,
                                                      float* __restrict__ pctx) {
    pdl_wait();
    constexpr int TP = 128;              // positions staged per step
    __shared__ float sE[TP][HD + 1];
    __shared__ float sV[TP][HD + 1];
    __shared__ float sMax[HD];
    const int ch = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int heads = gridDim.y;
    const int HW = H * W;
    const int per = (HW + nch - 1) / nch;
    const int p0 = ch * per, p1 = min(HW, p0 + per);
    if (threadIdx.x < HD) {
        float m = -INFINITY;
        for (int k = 0; k < nch; ++k) m = fmaxf(m, pmax[((size_t)b * nch + k) * D + g * HD + threadIdx.x]);
        sMax[threadIdx.x] = m;
    }
    __syncthreads();
    constexpr int PER_T = (HD * HD + 255) / 256;
    float acc[PER_T];
#pragma unroll
    for (int j = 0; j < PER_T; ++j) acc[j] = 0.f;
    float ssum = 0.f;      // threads < HD accumulate sum exp for channel threadIdx.x
    constexpr int VEC = sizeof(T) == 2 ? 8 : 4;          // elements per 16-byte load
    constexpr int CG = HD / VEC;                         // 16-byte groups per position and array
    const bool vec_ok = (ld % VEC) == 0 && ((D % VEC) == 0) && (((uintptr_t)qkv) % 16 == 0);
    for (int ps = p0; ps < p1; ps += TP) {
        for (int i = threadIdx.x; i < TP * CG; i += blockDim.x) {
            const int pp = i / CG, c = (i - pp * CG) * VEC;
            const int p = ps + pp;
            float e[VEC], v[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) { e[k] = 0.f; v[k] = 0.f; }
            bool keep = p < p1;
            if (keep && par != PAR_NONE) {
                int h = p / W, w = p - h * W;
                keep = parity_keep(par, h, w);
            }
            if (keep) {
                const T* base = qkv + ((size_t)b * HW + p) * ld + g * HD + c;
                if (vec_ok) {
                    if constexpr (sizeof(T) == 2) {
                        unpack8_bf16_k(*reinterpret_cast<const uint4*>(base + D), e);
                        unpack8_bf16_k(*reinterpret_cast<const uint4*>(base + 2 * D), v);
                    } else {
                        load4(reinterpret_cast<const float*>(base + D), e);
                        load4(reinterpret_cast<const float*>(base + 2 * D), v);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { e[k] = to_f<T>(base[D + k]); v[k] = to_f<T>(base[2 * D + k]); }
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) e[k] = expf(e[k] - sMax[c + k]);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) { sE[pp][c + k] = e[k]; sV[pp][c + k] = v[k]; }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER_T; ++j) {
            int idx = threadIdx.x + j * 256;
            if (idx < HD * HD) {
                int c1 = idx / HD, c2 = idx - c1 * HD;
                float a = acc[j];
#pragma unroll 8
                for (int pp = 0; pp < TP; ++pp) a = fmaf(sE[pp][c1], sV[pp][c2], a);
                acc[j] = a;
            }
        }
        if (threadIdx.x < HD) {
            for (int pp = 0; pp < TP; ++pp) ssum += sE[pp][threadIdx.x];
        }
        __syncthreads();
    }
    float* dst = pctx + (((size_t)b * heads + g) * nch + ch) * (HD * HD + HD);
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
        int idx = threadIdx.x + j * 256;
        if (idx < HD * HD) dst[idx] = acc[j];
    }
    if (threadIdx.x < HD) dst[HD * HD + threadIdx.x] = ssum;
}

// bf16 fast mode: the same chunk partial on the warp-level tensor path.  ctx[c1][c2] = sum_p E[p][c1] V[p][c2] is an
// HD x HD x positions GEMM: E = exp(K - max) (rounded to bf16; the normaliser sums the SAME rounded values) and V are staged
// position-major, both fragments come from ldmatrix.trans, accumulation is fp32 in position order.  HD = 32: one 16x8
// output tile per warp; HD = 16: two tiles x four position quarters, summed in a fixed order.
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <int HD>
__global__ void __launch_bounds__(256) lin_ctx_mma_kernel(const bf16* __restrict__ qkv, int ld, int D, int H, int W, int nch, int par,
                                                          const float* __restrict__ pmax, float* __restrict__ pctx) {
    pdl_wait();
    constexpr int TP = 256;                       // positions staged per step
    constexpr int PITCH = HD + 8;                 // bf16 elements per staged row (16-byte aligned, conflict-free ldmatrix rows)
    constexpr int CG = HD / 8;
    constexpr int TILES = (HD / 16) * (HD / 8), KPARTS = 8 / TILES, KLEN = TP / KPARTS;
    __shared__ __align__(16) bf16 sE[TP * PITCH];
    __shared__ __align__(16) bf16 sV[TP * PITCH];
    __shared__ float sMax[HD];
    __shared__ float sPart[KPARTS > 1 ? KPARTS * HD * HD : 1];
    const int ch = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int heads = gridDim.y;
    const int HW = H * W;
    const int per = (HW + nch - 1) / nch;
    const int p0 = ch * per, p1 = min(HW, p0 + per);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < HD) {
        float m = -INFINITY;
        for (int k = 0; k < nch; ++k) m = fmaxf(m, pmax[((size_t)b * nch + k) * D + g * HD + threadIdx.x]);
        sMax[threadIdx.x] = m;
    }
    __syncthreads();
    const int tile = warp % TILES, kpart = warp / TILES;
    const int mt = tile / (HD / 8), nt = tile % (HD / 8);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float ssum = 0.f;
    const uint32_t sE_s = (uint32_t)__cvta_generic_to_shared(sE), sV_s = (uint32_t)__cvta_generic_to_shared(sV);
    for (int ps = p0; ps < p1; ps += TP) {
        for (int i = threadIdx.x; i < TP * CG; i += blockDim.x) {
            const int pp = i / CG, c = (i - pp * CG) * 8;
            const int p = ps + pp;
            uint4 ev = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            bool keep = p < p1;
            if (keep && par != PAR_NONE) {
                int h = p / W, w = p - h * W;
                keep = parity_keep(par, h, w);
            }
            if (keep) {
                const bf16* base = qkv + ((size_t)b * HW + p) * ld + g * HD + c;
                float e[8];
                unpack8_bf16_k(*reinterpret_cast<const uint4*>(base + D), e);
                vv = *reinterpret_cast<const uint4*>(base + 2 * D);
#pragma unroll
                for (int k = 0; k < 8; ++k) e[k] = __expf(e[k] - sMax[c + k]);
                ev = make_uint4(pack_bf2(e[0], e[1]), pack_bf2(e[2], e[3]), pack_bf2(e[4], e[5]), pack_bf2(e[6], e[7]));
            }
            *reinterpret_cast<uint4*>(sE + (size_t)pp * PITCH + c) = ev;
            *reinterpret_cast<uint4*>(sV + (size_t)pp * PITCH + c) = vv;
        }
        __syncthreads();
        if (threadIdx.x < HD) {
            for (int pp = 0; pp < TP; ++pp) ssum += __bfloat162float(sE[(size_t)pp * PITCH + threadIdx.x]);
        }
#pragma unroll 4
        for (int k0 = kpart * KLEN; k0 < (kpart + 1) * KLEN; k0 += 16) {
            uint32_t a[4], b0, b1;
            {   // A = E^T tile: matrices (m 0-7, k 0-7) (m 8-15, k 0-7) (m 0-7, k 8-15) (m 8-15, k 8-15), each stored k-major -> .trans
                const int j = lane >> 3;
                const int pr = k0 + (j >> 1) * 8 + (lane & 7), cc = mt * 16 + (j & 1) * 8;
                ldsm_x4_trans(sE_s + (uint32_t)((pr * PITCH + cc) * 2), a);
            }
            {   // B = V tile (k x 8): matrices (k 0-7), (k 8-15)
                const int pr = k0 + ((lane >> 3) & 1) * 8 + (lane & 7);
                ldsm_x2_trans(sV_s + (uint32_t)((pr * PITCH + nt * 8) * 2), b0, b1);
            }
            mma_bf16_16816(acc, a, b0, b1);
        }
        __syncthreads();
    }
    float* dst = pctx + (((size_t)b * heads + g) * nch + ch) * (HD * HD + HD);
    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = mt * 16 + gq, c0 = nt * 8 + 2 * tq;
    if constexpr (KPARTS == 1) {
        dst[r0 * HD + c0] = acc[0]; dst[r0 * HD + c0 + 1] = acc[1];
        dst[(r0 + 8) * HD + c0] = acc[2]; dst[(r0 + 8) * HD + c0 + 1] = acc[3];
    } else {
        float* sp = sPart + kpart * HD * HD;
        sp[r0 * HD + c0] = acc[0]; sp[r0 * HD + c0 + 1] = acc[1];
        sp[(r0 + 8) * HD + c0] = acc[2]; sp[(r0 + 8) * HD + c0 + 1] = acc[3];
        __syncthreads();
        for (int idx = threadIdx.x; idx < HD * HD; idx += blockDim.x) {
            float a2 = 0.f;
#pragma unroll
            for (int k = 0; k < KPARTS; ++k) a2 += sPart[k * HD * HD + idx];
            dst[idx] = a2;
        }
    }
    if (threadIdx.x < HD) dst[HD * HD + threadIdx.x] = ssum;
}

template <int HD>
__global__ void lin_ctx_reduce_kernel(const float* __restrict__ pctx, int nch, float* __restrict__ ctx) {
    pdl_wait();
    // block = (head, b, quarter of the hd*hd entries); ctx[b][g][c1][c2] = sum_ch pctx / sum_ch S[c1].  The chunk loop is
    // unrolled by 8 with a fixed pairwise order: 8 loads in flight instead of a latency-bound chain, same result every run.
    const size_t bg = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    const float* src = pctx + bg * nch * (HD * HD + HD);
    constexpr int ST = HD * HD + HD;
    const int per = (HD * HD + gridDim.z - 1) / gridDim.z;
    const int i0 = blockIdx.z * per, i1 = min(HD * HD, i0 + per);
    for (int idx = i0 + threadIdx.x; idx < i1; idx += blockDim.x) {
        const int c1 = idx / HD;
        float a = 0.f, sden = 0.f;
        for (int k0 = 0; k0 < nch; k0 += 8) {
            float t[8], u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = k0 + j < nch;
                t[j] = ok ? src[(size_t)(k0 + j) * ST + idx] : 0.f;
                u[j] = ok ? src[(size_t)(k0 + j) * ST + HD * HD + c1] : 0.f;
            }
            a += ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            sden += ((u[0] + u[1]) + (u[2] + u[3])) + ((u[4] + u[5]) + (u[6] + u[7]));
        }
        ctx[bg * HD * HD + idx] = a / sden;
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) lin_out_kernel(const T* __restrict__ qkv, int ld, int H, int W, int par_q,
                                                      const float* __restrict__ ctx, T* __restrict__ out, int old) {
    pdl_wait();
    __shared__ float sC[HD][HD];
    const int g = blockIdx.y, b = blockIdx.z;
    const int heads = gridDim.y;
    const int HW = H * W;
    const float* cp = ctx + ((size_t)b * heads + g) * HD * HD;
    for (int i = threadIdx.x; i < HD * HD; i += blockDim.x) sC[i / HD][i % HD] = cp[i];
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    T* op = out + ((size_t)b * HW + p) * old + g * HD;
    bool keep = true;
    if (par_q != PAR_NONE) {
        int h = p / W, w = p - h * W;
        keep = parity_keep(par_q, h, w);
    }
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    if (keep) {
        const T* qp = qkv + ((size_t)b * HW + p) * ld + g * HD;
        float q[HD];
        float mx = -INFINITY;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            load4(qp + d, q + d);
            mx = fmaxf(fmaxf(mx, q[d]), fmaxf(q[d + 1], fmaxf(q[d + 2], q[d + 3])));
        }
        float den = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = expf(q[d] - mx);
            den += q[d];
        }
        float inv = 1.0f / den;
#pragma unroll
        for (int c1 = 0; c1 < HD; ++c1) {
            float qq = q[c1] * inv;
#pragma unroll
            for (int c2 = 0; c2 < HD; ++c2) o[c2] = fmaf(sC[c1][c2], qq, o[c2]);
        }
    }
#pragma unroll
    for (int d = 0; d < HD; d += 4) store4(op + d, o + d);
}

// bf16 fast mode, head dim 32, no parity filter (the inter context): O[p][:] = softmax_c(Q[p][:]) ctx on the warp-level tensor path.
// The scalar kernel above spends 32 x 32 FMAs and 256 shared loads per pixel-head (265 us at 288 channels, 32 images; the traffic
// floor is ~50 us).  Here a warp takes 16 pixels per step: lane (g, t) loads the 16 bytes of channels 8t .. 8t+7 of rows g and g + 8
// (whole 64-byte rows per quad), the softmax over the 32 channels of a row is a quad reduction, and the un-normalised exponentials
// (rounded to bf16; the normaliser sums the SAME rounded values, as lin_ctx_mma_kernel does) are the A fragments of mma.m16n8k16
// as they lie -- the k index of both operands is permuted so that fragment slot (t, j) is channel 8t + 4 kb + j; the context
// matrix is loaded once per warp into B fragments with the same permutation.  Results leave through a per-warp staging tile so that
// every lane stores 16 contiguous bytes.
constexpr int LO_TILES = 8;                      // 16-pixel tiles per warp
__global__ void __launch_bounds__(128) lin_out_mma32_kernel(const bf16* __restrict__ qkv, int ld, int HW, const float* __restrict__ ctx,
                                                            bf16* __restrict__ out, int old) {
    pdl_wait();
    __shared__ float sC[32][33];
    __shared__ __align__(16) uint8_t sO[4][16 * 80];
    const int g_head = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
    const float* cp = ctx + ((size_t)b * heads + g_head) * 1024;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sC[i >> 5][i & 31] = cp[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // B fragments: n = nt * 8 + g; k rows 2t, 2t+1 -> channels 8t + 4kb + {0, 1}; k rows 2t+8, 2t+9 -> channels 8t + 4kb + {2, 3}
    uint32_t bfr[4][2][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const int n = nt * 8 + g, c = 8 * t + 4 * kb;
            bfr[nt][kb][0] = pack_bf2(sC[c][n], sC[c + 1][n]);
            bfr[nt][kb][1] = pack_bf2(sC[c + 2][n], sC[c + 3][n]);
        }
    uint8_t* myO = sO[warp];
    const int p_base = (blockIdx.x * 4 + warp) * (16 * LO_TILES);
#pragma unroll 1
    for (int tl = 0; tl < LO_TILES; ++tl) {
        const int p0 = p_base + tl * 16;
        if (p0 >= HW) break;                                     // warp-uniform
        float inv[2];
        uint32_t a[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = p0 + g + 8 * h;
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (p < HW) raw = *reinterpret_cast<const uint4*>(qkv + ((size_t)b * HW + p) * ld + g_head * 32 + 8 * t);
            float v[8];
            unpack8_bf16_k(raw, v);
            float mx = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            uint32_t e2[4];
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                e2[j] = pack_bf2(__expf(v[2 * j] - mx), __expf(v[2 * j + 1] - mx));
                den += __uint_as_float(e2[j] << 16) + __uint_as_float(e2[j] & 0xffff0000u);
            }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            inv[h] = 1.0f / den;
            a[0][h] = e2[0]; a[0][2 + h] = e2[1];                // kb 0: slots (2t, 2t+1) and (2t+8, 2t+9) of row g + 8h
            a[1][h] = e2[2]; a[1][2 + h] = e2[3];
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(acc, a[0], bfr[nt][0][0], bfr[nt][0][1]);
            mma_bf16_16816(acc, a[1], bfr[nt][1][0], bfr[nt][1][1]);
            *reinterpret_cast<uint32_t*>(myO + g * 80 + nt * 16 + 4 * t) = pack_bf2(acc[0] * inv[0], acc[1] * inv[0]);
            *reinterpret_cast<uint32_t*>(myO + (g + 8) * 80 + nt * 16 + 4 * t) = pack_bf2(acc[2] * inv[1], acc[3] * inv[1]);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = (lane >> 2) + 8 * h, p = p0 + row;
            if (p < HW) *reinterpret_cast<uint4*>(out + ((size_t)b * HW + p) * old + g_head * 32 + 8 * t) = *reinterpret_cast<const uint4*>(myO + row * 80 + 16 * t);
        }
        __syncwarp();
    }
}

int launch_lin_attn(int bf, const Act& qkv, int D, int heads, int hd, int par_kv, int par_q, float* scratch,
                    const Act& out, cudaStream_t s) {
    const int B = qkv.B, H = qkv.H, W = qkv.W, HW = H * W;
    if (B * HW == 0) return 0;
    const int nch = lin_chunks(HW);
    float* pmax = scratch;
    float* pctx = pmax + (size_t)B * nch * D;
    float* ctx = pctx + (size_t)B * heads * nch * (hd * hd + hd);
#define LIN_LAUNCH(T, HD)                                                                                          \
    do {                                                                                                           \
        launch_k(lin_colmax_kernel<T>, dim3(dim3(nch, (D + 31) / 32, B)), dim3(256), 0, s, (const T*)qkv.p, qkv.ld, D, H, W, nch, par_kv, pmax); \
        if (sizeof(T) == 2 && (qkv.ld % 8) == 0 && (D % 8) == 0 && (((uintptr_t)qkv.p) % 16) == 0 && !getenv("MLIC_LIN_SIMT"))                  \
            launch_k(lin_ctx_mma_kernel<HD>, dim3(dim3(nch, heads, B)), dim3(256), 0, s, (const bf16*)qkv.p, qkv.ld, D, H, W, nch, par_kv, pmax, pctx); \
        else                                                                                                       \
        launch_k(lin_ctx_kernel<T, HD>, dim3(dim3(nch, heads, B)), dim3(256), 0, s, (const T*)qkv.p, qkv.ld, D, H, W, nch, par_kv,   \
                                                                  pmax, pctx);                                     \
        launch_k(lin_ctx_reduce_kernel<HD>, dim3(dim3(heads, B, 4)), dim3(256), 0, s, pctx, nch, ctx);                                  \
        if (sizeof(T) == 2 && HD == 32 && par_q == PAR_NONE && (qkv.ld % 8) == 0 && (out.ld % 8) == 0 && (((uintptr_t)qkv.p) % 16) == 0 &&      \
            (((uintptr_t)out.p) % 16) == 0 && !getenv("MLIC_LIN_SIMT"))                                                                      \
            launch_k(lin_out_mma32_kernel, dim3(dim3(cdiv(HW, 4 * 16 * LO_TILES), heads, B)), dim3(128), 0, s, (const bf16*)qkv.p, qkv.ld, HW, ctx, (bf16*)out.p, out.ld); \
        else                                                                                                       \
        launch_k(lin_out_kernel<T, HD>, dim3(dim3(cdiv(HW, 128), heads, B)), dim3(128), 0, s, (const T*)qkv.p, qkv.ld, H, W, par_q,  \
                                                                            ctx, (T*)out.p, out.ld);               \
    } while (0)
    if (hd == 32) { if (bf) LIN_LAUNCH(bf16, 32); else LIN_LAUNCH(float, 32); }
    else if (hd == 16) { if (bf) LIN_LAUNCH(bf16, 16); else LIN_LAUNCH(float, 16); }
    else return 1;
#undef LIN_LAUNCH
    return 0;
}

// ------------------------------------------------------------------------------------------
// Fused quantise / likelihood / CDF-index (models/mlicpp.py:117,132-135;
// utils/ckbd.py:82-90,123-158).  One thread per (pixel, channel), channel-fastest.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int cdf_index(float sigma, const float* __restrict__ table, int levels) {
    float s = fmaxf(sigma, 0.11f);
    int idx = levels - 1;
    for (int k = 0; k < levels - 1; ++k) idx -= (s <= table[k]) ? 1 : 0;
    return idx;
}
__device__ __forceinline__ float gauss_lik(float yq, float sigma, float mu) {
    float s = fmaxf(sigma, 0.11f);
    float v = fabsf(yq - mu);
    const float c = -0.70710678118654752440f;
    float up = 0.5f * erfcf(c * ((0.5f - v) / s));
    float lo = 0.5f * erfcf(c * ((-0.5f - v) / s));
    return fmaxf(up - lo, 1e-9f);
}

// symbols / gain + means as the reference evaluates it (torch: a product, then a sum: two roundings, no FMA contraction), so
// that the encoder-side and the decoder-side walks rebuild bit-identical y_hat whatever the compiler fuses elsewhere
__device__ __forceinline__ float vbr_deq(float sq, float rgain, float mu) { return __fadd_rn(__fmul_rn(sq, rgain), mu); }

// I: index type -- int when B*H*W*C < 2^31 (the four divisions per element are 64-bit otherwise: a third of the kernel's instructions)
template <typename T, typename I>
__global__ void quant_anchor_kernel(QuantArgs a) {
    pdl_wait();
    // grid = (chunks of a row's W * C elements, H, B): one division (a shift for C = 32) per element instead of four
    const int C = a.C;
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;
    for (int once = (ir < a.W * C) ? 0 : 1; once < 1; ++once) {
        const int h = blockIdx.y, b = blockIdx.z;
        const int w = (C & (C - 1)) ? ir / C : ir >> (31 - __clz(C));
        const int c = ir - w * C;
        const I q = (I)b * a.H + h;
        const I p = q * a.W + w;
        T* slot = reinterpret_cast<T*>(a.slot.p) + (long long)p * a.slot.ld + c;
        if (((h + w) & 1) == 0) { if (a.mode != 3) *slot = from_f<T>(0.f); continue; }
        const long long pe = a.sq ? ((long long)q * (a.W >> 1) + (w >> 1)) : (long long)p;       // row of this pixel in the entropy-parameter buffer
        const float sigma = a.pa[pe * 2 * C + c];
        const float mu = a.pa[pe * 2 * C + C + c];
        float out;
        if (a.mode == 2) {
            out = mu;
        } else if (a.mode >= 3) {       // decompress (utils/ckbd.py:195-211): 3 = index list for the range decoder, 4 = symbols back
            size_t o = (((size_t)b * C + c) * a.H + h) * (a.W / 2) + (w >> 1);
            if (a.mode == 3) { a.idx[o] = cdf_index(a.vbr ? sigma * a.gain : sigma, a.table, a.levels); continue; }
            const float sq = (float)a.sym[o];
            out = a.vbr ? vbr_deq(sq, a.rgain, mu) : sq + mu;
        } else {
            const float y = a.y[(long long)p * a.y_ld + c];
            if (a.mode == 0) {
                out = a.vbr ? vbr_deq(rintf((y - mu) * a.gain), a.rgain, mu) : rintf(y - mu) + mu;
            } else {
                float sq = a.vbr ? rintf((y - mu) - mu) : rintf(y - mu);
                size_t o = (((size_t)b * C + c) * a.H + h) * (a.W / 2) + (w >> 1);
                a.sym[o] = (int32_t)sq;
                a.idx[o] = cdf_index(a.vbr ? sigma * a.gain : sigma, a.table, a.levels);
                out = a.vbr ? vbr_deq(sq, a.rgain, mu) : sq + mu;
            }
        }
        *slot = from_f<T>(out);
    }
}

// I: index type -- int when B*H*W*C < 2^31 (the four divisions per element are 64-bit otherwise: a third of the kernel's instructions)
template <typename T, typename I>
__global__ void quant_nonanchor_kernel(QuantArgs a) {
    pdl_wait();
    // grid = (chunks of a row's W * C elements, H, B): one division (a shift for C = 32) per element instead of four
    const int C = a.C;
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;
    for (int once = (ir < a.W * C) ? 0 : 1; once < 1; ++once) {
        const int h = blockIdx.y, b = blockIdx.z;
        const int w = (C & (C - 1)) ? ir / C : ir >> (31 - __clz(C));
        const int c = ir - w * C;
        const I q = (I)b * a.H + h;
        const I p = q * a.W + w;
        const bool anchor = ((h + w) & 1) == 1;
        T* slot = reinterpret_cast<T*>(a.slot.p) + (long long)p * a.slot.ld + c;
        const float* pp = anchor ? a.pa : a.pn;
        const long long pe = a.sq ? ((long long)q * (a.W >> 1) + (w >> 1)) : (long long)p;
        const float sigma = pp[pe * 2 * C + c];
        const float mu = pp[pe * 2 * C + C + c];
        if (a.mode == 2) {           // decoder walk: both halves take means_anchor (mlicpp.py:405,418)
            if (anchor) *slot = from_f<T>(to_f<T>(*slot) + mu);
            continue;
        }
        if (a.mode >= 3) {           // decompress (utils/ckbd.py:213-229)
            if (anchor) continue;
            size_t o = (((size_t)b * C + c) * a.H + h) * (a.W / 2) + (w >> 1);
            if (a.mode == 3) a.idx[o] = cdf_index(a.vbr ? sigma * a.gain : sigma, a.table, a.levels);
            else { const float sq = (float)a.sym[o]; *slot = from_f<T>(a.vbr ? vbr_deq(sq, a.rgain, mu) : sq + mu); }
            continue;
        }
        const float y = a.y[(long long)p * a.y_ld + c];
        if (a.mode == 0) {
            float lik;
            if (a.vbr) {
                float yg = y * a.gain, sg = sigma * a.gain, mg = mu * a.gain;
                lik = gauss_lik(rintf(yg - mg) + mg, sg, mg);
            } else {
                lik = gauss_lik(rintf(y - mu) + mu, sigma, mu);
            }
            a.lik[(long long)p * a.lik_ld + c] = lik;
            if (!anchor) *slot = from_f<T>(a.vbr ? vbr_deq(rintf((y - mu) * a.gain), a.rgain, mu) : rintf(y - mu) + mu);
        } else if (!anchor) {
            float sq = rintf(y - mu);
            size_t o = (((size_t)b * C + c) * a.H + h) * (a.W / 2) + (w >> 1);
            a.sym[o] = (int32_t)sq;
            a.idx[o] = cdf_index(a.vbr ? sigma * a.gain : sigma, a.table, a.levels);
            *slot = from_f<T>(a.vbr ? vbr_deq(sq, a.rgain, mu) : sq + mu);
        }
    }
}

static bool quant_grid(const QuantArgs& a, dim3& grid, bool& small) {
    const long long total = (long long)a.B * a.H * a.W * a.C;
    if (!total) return false;
    grid = dim3((unsigned)cdiv((long long)a.W * a.C, 256), (unsigned)a.H, (unsigned)a.B);
    small = total < 0x7fffffffLL;
    return true;
}
void launch_quant_anchor(int bf, const QuantArgs& a, cudaStream_t s) {
    dim3 grid; bool small;
    if (!quant_grid(a, grid, small)) return;
    if (bf) { if (small) launch_k(quant_anchor_kernel<bf16, int>, grid, dim3(256), 0, s, a); else launch_k(quant_anchor_kernel<bf16, long long>, grid, dim3(256), 0, s, a); }
    else { if (small) launch_k(quant_anchor_kernel<float, int>, grid, dim3(256), 0, s, a); else launch_k(quant_anchor_kernel<float, long long>, grid, dim3(256), 0, s, a); }
}
void launch_quant_nonanchor(int bf, const QuantArgs& a, cudaStream_t s) {
    dim3 grid; bool small;
    if (!quant_grid(a, grid, small)) return;
    if (bf) { if (small) launch_k(quant_nonanchor_kernel<bf16, int>, grid, dim3(256), 0, s, a); else launch_k(quant_nonanchor_kernel<bf16, long long>, grid, dim3(256), 0, s, a); }
    else { if (small) launch_k(quant_nonanchor_kernel<float, int>, grid, dim3(256), 0, s, a); else launch_k(quant_nonanchor_kernel<float, long long>, grid, dim3(256), 0, s, a); }
}

// Stand-alone flat version (any layout, elementwise): the GaussianConditional boundary of the C ABI.
__global__ void gc_flat_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ mu,
                               size_t n, float* __restrict__ y_hat, float* __restrict__ lik, int32_t* __restrict__ sym,
                               int32_t* __restrict__ idx, const float* __restrict__ table, int levels) {
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float m = mu[i], s = sc[i];
        float q = rintf(y[i] - m);
        float yh = q + m;
        if (y_hat) y_hat[i] = yh;
        if (lik) lik[i] = gauss_lik(yh, s, m);
        if (sym) sym[i] = (int32_t)q;
        if (idx) idx[i] = cdf_index(s, table, levels);
    }
}
void launch_gc_flat(const float* y, const float* sc, const float* mu, size_t n, float* y_hat, float* lik, int32_t* sym,
                    int32_t* idx, const float* table, int levels, cudaStream_t s) {
    if (!n) return;
    int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    launch_k(gc_flat_kernel, dim3(blocks), dim3(256), 0, s, y, sc, mu, n, y_hat, lik, sym, idx, table, levels);
}

// ------------------------------------------------------------------------------------------
// Rate / distortion sums (loss/rd_loss.py:37-48): out[0] += sum log2(lik), out[1] += sum (a-b)^2.
// Deterministic two-level reduction in double: per-block partials then a single-block final pass.
// ------------------------------------------------------------------------------------------
__global__ void reduce_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int mode,
                                      double* __restrict__ partial) {
    pdl_wait();
    __shared__ double sh[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) acc += (double)log2f(a[i]);
        else { double d = (double)a[i] - (double)b[i]; acc += d * d; }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void reduce_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    pdl_wait();
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += sh[0];
}
void launch_reduce(const float* a, const float* b, long long n, int mode, double* partial /*[RD_BLOCKS]*/, double* out,
                   cudaStream_t s) {
    if (n <= 0) return;
    int blocks = (int)std::min<long long>(cdiv(n, 256), RD_BLOCKS);
    launch_k(reduce_partial_kernel, dim3(blocks), dim3(256), 0, s, a, b, n, mode, partial);
    launch_k(reduce_final_kernel, dim3(1), dim3(256), 0, s, partial, blocks, out);
}

}  // namespace mlic
