// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
//   D[m][n] = sum_{tap, c} X[b, oh + ky - pad, ow + kx - pad, c] * Wt[n][tap * Cpad + c]
//
// A tile is 128 output pixels x BN columns.  The 128 pixels are a TH x TW patch of the output grid
// (TH * TW = 128) so that the A operand of every (tap, 64-channel chunk) K step is ONE 4-D TMA box
// {64 ch, TW, TH, 1} of the NHWC activation: the conv halo / zero padding is the TMA's out-of-bounds zero
// fill, and there is no im2col buffer.  The box lands in shared memory as 128 rows of 128 B with the
// 128-byte swizzle, which is exactly the canonical K-major SWIZZLE_128B UMMA layout.  B (weights,
// [N][taps * Cpad] bf16, K-major) is a 2-D TMA box {64, BN}.
//
// Persistent kernel, one CTA per SM, 640 threads: warp 0 = TMA producer (one elected lane), warp 1 = MMA
// issuer (one elected lane: tcgen05.mma cta_group::1 kind::f16, M = 128, N = BN, K = 16), warp 2 = TMEM
// allocator, warps 4..19 = epilogue (tcgen05.ld 32x32b; warp w reads TMEM lanes 32*(w%4).., the four warps of a
// lane quarter split every 64-column block; one output pixel per thread; bias / (I)GDN / GELU / tanh / parity
// mask / residual / pixel-shuffle / x^2 side output).  bf16 outputs are staged in 128B-swizzled shared memory and
// written by TMA stores (full 128-byte lines; one tensor map per pixel-shuffle group); fp32 / ragged outputs are
// written directly.  Shared memory holds a 3..8 stage TMA ring; TMEM holds two accumulator stages so the epilogue
// of tile i overlaps the main loop of tile i+1.  Every mbarrier wait is bounded and traps instead of hanging.
#include "kernels.h"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlic {

static thread_local char g_tc_err[512] = "";
const char* tc_last_error() { return g_tc_err; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int tc_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(err));
        return 1;
    }
    g_encode = (PFN_encodeTiled)fn;
    return 0;
}

void* tc_encode_fn() { return (void*)g_encode; }

// Per-device launch state: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to the CURRENT device,
// not to the process (a model moved to a second GPU of the same process launches there with the first one's attributes otherwise).
constexpr int TC_MAX_DEV = 64;
static int cur_dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < TC_MAX_DEV) ? d : 0; }
static int dev_sms(int dev) {
    static int sms[TC_MAX_DEV] = {};
    if (!sms[dev]) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); sms[dev] = n > 0 ? n : 148; }
    return sms[dev];
}

enum { STORE_DIRECT = 0, STORE_TMA = 1, STORE_NCHW3 = 2, STORE_SS = 3 };

struct TcParams {
    int tilesH, tilesW;     // output patches per image
    int TH, TW;             // patch shape, TH * TW = 128
    int ks, pad;
    int kchunks;            // Cpad / 64
    int Cpad;
    int BN;                 // columns per tile (multiple of 16; of 64 in STORE_TMA mode; <= 256)
    int tilesN;
    int ntiles;             // B * tilesH * tilesW * tilesN
    int stages;
    int acc_stride;         // TMEM columns between the two accumulator stages (power of two >= BN)
    int epi_vec;            // STORE_DIRECT: 16-byte epilogue accesses are legal for every pointer involved
    int ld_vec;             // 16-byte loads of the residual / GDN operand are legal
    int store_mode;
    int bsplit;             // 1: the weight tile of every K step is issued by a second thread (warp 3), the activation patch by warp 0
    int esplit;             // STORE_TMA: epilogue warp groups per 64-column block (narrow GEMMs: 4 or 2 groups share a block, 16 or 32 columns each)
    int nstg;               // STORE_TMA: 16 KB staging buffers behind the stage ring (ring slots x buffers per block)
    int b_resident;         // 1: the whole weight matrix (all K chunks x BN rows, tilesN == 1) is loaded into shared
                            //    memory once per CTA; the ring then carries only the A patches
    int raw_bytes;          // PROD_DW: bytes of one halo patch {64 ch, TW+2, TH+2} (multiple of 128)
    int Cin;                // PROD_DW: channels of the depthwise conv (= K of the GEMM)
    const float* dw_w9;     // PROD_DW: depthwise weights [9][Cin] fp32
    const float* dw_bias;   // PROD_DW: depthwise bias [Cin]
    int ck;                 // checkerboard-squeezed A operand: `a` is a 5-D map (c, j, h & 1, h >> 1, b), box {64, TW, 2, TH/2, 1}
    int tsh, tsw, torg;     // tile origin of tile (th, tw) = (th * tsh + torg, tw * tsw + torg); TH / TW / 0 except in shift-sum mode
    int halo;               // 1 (ks > 1, weights resident): the A stage of a 64-channel chunk is ONE halo patch {64, halo_w, TH+ks-1}
                            //    and the ks*ks taps are UMMA descriptors into it (start + (ky*halo_w + kx) pixels): the input is
                            //    read from L2 once per chunk instead of once per tap
    int halo_w;             // staged patch width in pixels (TW + ks - 1, optionally padded to 16)
    int halo_h;
    int halo_bo;            // descriptor variant: set the base-offset field from the start address
    int acc_split;          // halo + narrow N: ks*ks independent accumulators of BN columns each (taps), summed in the epilogue
    int a_bytes;            // bytes of the A part of a ring stage (TC_A_BYTES, or the halo patch rounded up to 1024)
    int l2_prefetch;        // > 0 (1x1 convs): prefetch the A patch of the tile this many iterations ahead into L2
    int debug;              // development only: bit0 skip epilogue stores, bit1 skip TMEM loads
};

struct TcMaps {
    CUtensorMap a, b;       // operands (PROD_DW: `a` is the halo-patch map, box {64, TW+2, TH+2, 1}, no swizzle)
    CUtensorMap o[4];       // output (one per pixel-shuffle group; o[0] when not shuffled)
    CUtensorMap o2;         // x^2 side output
    CUtensorMap r, gx;      // STORE_TMA: residual / GDN-operand blocks, fetched by TMA into the staging slot (geometry of o[0])
};

constexpr int TC_A_BYTES = 128 * 128;       // 128 rows x 64 bf16
constexpr int TC_SS_PITCH = 116;            // shift-sum: floats per pixel row of the partial-product buffer (464 B: 16-byte stores of 8 consecutive rows hit 8 distinct bank groups)
constexpr int TC_STG_BYTES = 128 * 128;     // one staged 128-pixel x 64-column bf16 block
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 128 + TC_EPI_WARPS * 32;
// Fused A-operand producers (PROD template parameter): the A stage of a 1x1 GEMM is not a TMA copy of the input but
//   PROD_DW: the depthwise 3x3 (+ bias) of the input patch, computed by 8 CUDA-core warps from a TMA-loaded halo patch
//            -> DepthWiseConv (modules/layers/conv.py:46-63) in one kernel, the depthwise output never touches HBM;
//   PROD_SQ: the element-wise square of the input patch (GDN / IGDN: norm = conv1x1(x^2), CompressAI GDN)
//            -> no x^2 side tensor.
enum { PROD_TMA = 0, PROD_DW = 1, PROD_SQ = 2 };
constexpr int TC_PROD_WARPS = 8;                                   // warps 4..11
constexpr int TC_FUSED_EPI_WARPS = 12;                             // warps 12..23 (N <= 192: three 64-column groups)
constexpr int TC_FUSED_THREADS = 128 + (TC_PROD_WARPS + TC_FUSED_EPI_WARPS) * 32;
constexpr int TC_RAW_SLOTS = 2;
constexpr int TC_MAX_N = 2048;
constexpr int TC_MAX_BIAS = TC_MAX_N + 256;

// Epilogue arithmetic for 8 consecutive GEMM columns n..n+7 of one output pixel, in the order of epi_store4
// (common.cuh): premask, + bias, (I)GDN, activation, postmask, + residual.  bf16 activations.
// The epilogue is issue-bound (measured: 72 % issue-slot utilisation with the scalar erf form, ncu profiles/), so the
// arithmetic runs on packed fp32 pairs (FFMA2 / FMUL2 / FADD2: one issue slot per two columns) and GELU uses ONE MUFU:
//   GELU(x) = x Phi(x),  Phi(x) ~= 0.5 + 0.5 tanh(x (a + b t + c t^2)),  t = min(x^2, 64)
// with (a, b, c) the minimax fit to the erf form (|error| <= 2.6e-5 with an exact tanh; tanh.approx adds <= 2^-11
// relative on tanh, i.e. <= 2.5e-4 |x| absolute, below half a bf16 ulp of the stored result for |x| >= 0.13).  Fast
// (bf16) mode only; the fp32 validation path keeps erff.
struct EpiRow {
    const bf16* xp;      // GDN operand row (at column 0) or null
    const bf16* rp;      // residual row (at output channel 0 of this pixel) or null
    bool keep_pre, keep_post;
    int h, w;            // GEMM-space pixel (column-group premasks: Epi::pm_w)
};
// STAGED: the GDN operand / residual chunk of this thread was prefetched into the staging slot (xs / rs, shared memory)
template <int ACT, int GDN, bool RES, bool STAGED>
__device__ __forceinline__ void epi_math8(const Epi& e, const float* __restrict__ sBias, const EpiRow& row, int n, int oc,
                                          float* a, bool ldvec, uint32_t xs = 0, uint32_t rs = 0) {
    const bool full = ldvec && n + 8 <= e.N;
    float2 v[4];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + n), b1 = *reinterpret_cast<const float4*>(sBias + n + 4);
        if (e.pm_w ? !parity_keep(premask_of(e, n), row.h, row.w) : !row.keep_pre) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = 0.0f;
        }
        v[0] = __fadd2_rn(make_float2(a[0], a[1]), make_float2(b0.x, b0.y));
        v[1] = __fadd2_rn(make_float2(a[2], a[3]), make_float2(b0.z, b0.w));
        v[2] = __fadd2_rn(make_float2(a[4], a[5]), make_float2(b1.x, b1.y));
        v[3] = __fadd2_rn(make_float2(a[6], a[7]), make_float2(b1.z, b1.w));
    }
    if constexpr (GDN != GDN_NONE) {
        float2 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = make_float2(0.f, 0.f);
        if constexpr (STAGED) {
            const uint4 t = lds128(xs);
            x[0] = bf2_to_f2(t.x); x[1] = bf2_to_f2(t.y); x[2] = bf2_to_f2(t.z); x[3] = bf2_to_f2(t.w);
        } else if (row.xp) {
            if (full) {
                const uint4 t = *reinterpret_cast<const uint4*>(row.xp + n);
                x[0] = bf2_to_f2(t.x); x[1] = bf2_to_f2(t.y); x[2] = bf2_to_f2(t.z); x[3] = bf2_to_f2(t.w);
            } else {
                float xs1[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xs1[j] = 0.f;
#pragma unroll 1
                for (int j = 0; j < 8; ++j) if (n + j < e.N) xs1[j] = __bfloat162float(row.xp[n + j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = make_float2(xs1[2 * j], xs1[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s0, s1;
            if constexpr (GDN == GDN_FWD) {
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(v[j].x));
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(v[j].y));
            } else {
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(v[j].x));
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(v[j].y));
            }
            v[j] = __fmul2_rn(x[j], make_float2(s0, s1));
        }
    }
    if constexpr (ACT == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = gelu2(v[j]);
    } else if constexpr (ACT == ACT_HALF_TANH) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float2(0.5f * tanh_approx(v[j].x), 0.5f * tanh_approx(v[j].y));      // MUFU.TANH, 2^-11 relative: below the bf16 rounding of the LRP output
    }
    if (!row.keep_post) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float2(0.f, 0.f);
    }
    if constexpr (RES && STAGED) {
        const uint4 t = lds128(rs);
        v[0] = __fadd2_rn(v[0], bf2_to_f2(t.x)); v[1] = __fadd2_rn(v[1], bf2_to_f2(t.y));
        v[2] = __fadd2_rn(v[2], bf2_to_f2(t.z)); v[3] = __fadd2_rn(v[3], bf2_to_f2(t.w));
    } else if constexpr (RES) {
        if (row.rp) {
            if (full) {
                const uint4 t = *reinterpret_cast<const uint4*>(row.rp + oc);
                v[0] = __fadd2_rn(v[0], bf2_to_f2(t.x)); v[1] = __fadd2_rn(v[1], bf2_to_f2(t.y));
                v[2] = __fadd2_rn(v[2], bf2_to_f2(t.z)); v[3] = __fadd2_rn(v[3], bf2_to_f2(t.w));
            } else {
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = 0.f;
#pragma unroll 1
                for (int j = 0; j < 8; ++j) if (n + j < e.N) r[j] = __bfloat162float(row.rp[oc + j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = __fadd2_rn(v[j], make_float2(r[2 * j], r[2 * j + 1]));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[2 * j] = v[j].x; a[2 * j + 1] = v[j].y; }
}

// pixel-shuffle bookkeeping of one GEMM column: group g = n / Cq -> output pixel (2h + (g >> 1), 2w + (g & 1)), channel n % Cq
__device__ __forceinline__ void out_coord(const Epi& e, int h, int w, int n, int& oh, int& ow, int& oc, int& OH, int& OW) {
    oh = h; ow = w; oc = n; OH = e.Hout; OW = e.Wout;
    if (e.shuffle) {
        const int Cq = e.N >> 2;
        const int g = n / Cq;
        oc = n - g * Cq; oh = 2 * h + (g >> 1); ow = 2 * w + (g & 1); OH *= 2; OW *= 2;
    }
}

// STORE_DIRECT: the thread writes its own 8 columns (bias / activation / residual already applied).
__device__ __forceinline__ void direct_store8(const Epi& e, int b, int h, int w, int n, const float* v, bool vec) {
    if (vec && n + 8 <= e.N) {
        int oh, ow, oc, OH, OW;
        out_coord(e, h, w, n, oh, ow, oc, OH, OW);
        const size_t opix = ((size_t)b * OH + oh) * OW + ow;
        if (e.out_f32) {
            float* op = reinterpret_cast<float*>(e.out) + opix * e.out_ld + oc;
            *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.out) + opix * e.out_ld + oc) = pack8_bf16(v);
        }
        if (e.out2) {
            float s8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s8[k] = v[k] * v[k];
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.out2) + opix * e.out2_ld + oc) = pack8_bf16(s8);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {        // ragged N, odd shuffle groups, unaligned views
        const int nj = n + j;
        if (nj >= e.N) break;
        int oh, ow, oc, OH, OW;
        out_coord(e, h, w, nj, oh, ow, oc, OH, OW);
        const size_t opix = ((size_t)b * OH + oh) * OW + ow;
        if (e.out_f32) reinterpret_cast<float*>(e.out)[opix * e.out_ld + oc] = v[j];
        else reinterpret_cast<bf16*>(e.out)[opix * e.out_ld + oc] = __float2bfloat16_rn(v[j]);
        if (e.out2) reinterpret_cast<bf16*>(e.out2)[opix * e.out2_ld + oc] = __float2bfloat16_rn(v[j] * v[j]);
    }
}

// Persistent kernel: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// t = blockIdx.x, blockIdx.x + gridDim.x, ...  (N-tile fastest, so co-running CTAs share the A patch in L2).
template <int ACT, int GDN, bool RES, int PROD>
__global__ void __launch_bounds__(PROD == PROD_TMA ? TC_THREADS : TC_FUSED_THREADS) __maxnreg__(PROD == PROD_TMA ? 96 : 80)
conv_gemm_tc_kernel(const __grid_constant__ TcMaps tm, TcParams p, Epi e, unsigned long long* __restrict__ dbg) {
    constexpr int NTHREADS = PROD == PROD_TMA ? TC_THREADS : TC_FUSED_THREADS;
    constexpr int EW0 = PROD == PROD_TMA ? 4 : 4 + TC_PROD_WARPS;          // first epilogue warp (multiple of 4)
    constexpr int NEPI = PROD == PROD_TMA ? TC_EPI_WARPS : TC_FUSED_EPI_WARPS;
    // dbg (development only, may be null): per CTA {total, producer wait-empty, mma wait-full, mma wait-acc-empty,
    // epilogue wait-acc-full (warp 4), epilogue busy (warp 4)} in SM clocks
    const long long t_start = clock64();
    long long w0c = 0, w1c = 0, w2c = 0, w3c = 0;
#define TIMED_WAIT(acc, bar, par) do { if (dbg) { long long _t = clock64(); mbar_wait(bar, par); acc += clock64() - _t; } else mbar_wait(bar, par); } while (0)
    // single-thread roles (TMA issue, MMA issue): spin where the tensor pipe is the pace (plain GEMMs: wake-up latency counts), park
    // in the fused-producer kernels, which are issue-slot bound (the two polling loops were 20 % of their warp instructions)
    constexpr bool SPIN_PARKS = PROD != PROD_TMA;
#define TIMED_SPIN(acc, bar, par) do { if (dbg) { long long _t = clock64(); mbar_wait<SPIN_PARKS>(bar, par); acc += clock64() - _t; } else mbar_wait<SPIN_PARKS>(bar, par); } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.BN * 128;
    const int stage_bytes = p.a_bytes + (p.b_resident ? 0 : b_bytes);      // multiple of 1024
    const int ksteps_total = p.ks * p.ks * p.kchunks;
    // layout: [resident B: ksteps x b_bytes]? [ring: stages x stage_bytes] [staging]
    uint8_t* bres = base;
    if (p.b_resident) base += (size_t)ksteps_total * b_bytes;
    // fused depthwise producer: [.. ring | staging | raw ring: TC_RAW_SLOTS x raw_bytes | dw weights 9 x Cin + bias Cin (fp32)]
    uint8_t* rawbuf = base + (size_t)p.stages * stage_bytes + (size_t)p.nstg * TC_STG_BYTES;
    float* sDw = reinterpret_cast<float*>(rawbuf + (size_t)TC_RAW_SLOTS * p.raw_bytes);
    (void)rawbuf; (void)sDw;
    if constexpr (PROD == PROD_DW) {
        for (int i = threadIdx.x; i < 9 * p.Cin; i += NTHREADS) sDw[i] = p.dw_w9[i];
        for (int i = threadIdx.x; i < p.Cin; i += NTHREADS) sDw[9 * p.Cin + i] = p.dw_bias[i];
    }
    __shared__ uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], acc_full[2], acc_empty[2], bres_bar;
    __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_RAW_SLOTS];     // fused producers: TMA -> compute warps
    __shared__ uint64_t stg_bar[4][2];                                        // epilogue operand blocks (residual / GDN x) landed
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float sBias[TC_MAX_BIAS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tilesH * p.tilesW;
    const int ksteps = p.ks * p.ks * p.kchunks;
    // tile walk of every role: t = t_first, t_first + t_step, ... < t_end; (nt, mt) = TILE_OF(t), N tile fastest
    const int t_first = (int)blockIdx.x, t_step = (int)gridDim.x, t_end = p.ntiles;
#define TILE_OF(t, nt, mt) const int nt = (t) % p.tilesN; const int mt = (t) / p.tilesN

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // full_bar: one arrival from the TMA producer, or one per compute warp when the A stage is produced on chip
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], PROD == PROD_TMA ? 1 : TC_PROD_WARPS); mbar_init(&empty_bar[s], 1); }
        if (PROD != PROD_TMA) {
            for (int s = 0; s < TC_MAX_STAGES; ++s) mbar_init(&raw_full[s], 1);
            for (int s = 0; s < TC_RAW_SLOTS; ++s) mbar_init(&raw_empty[s], TC_PROD_WARPS);
        }
        // accumulator release: one arrival per epilogue warp that reads it (TMA-store mode: 4 warps per 64-column block)
        const uint32_t nrel = p.store_mode == STORE_TMA ? 4u * (uint32_t)((p.BN + 63) / 64) * (uint32_t)p.esplit : (uint32_t)NEPI;
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], nrel); }
        mbar_init(&bres_bar, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&stg_bar[i >> 1][i & 1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)(2 * p.acc_stride))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TC_MAX_BIAS; i += NTHREADS) sBias[i] = (e.bias && i < e.N) ? e.bias[i] : 0.0f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // Programmatic dependent launch (launch attribute set by launch_conv_tc when enabled; both instructions are no-ops
    // otherwise): everything above -- barrier init, TMEM allocation, bias / depthwise weights (parameters, never written by
    // a kernel of the walk) -- may run while the previous kernel of the stream drains; no activation is touched before the
    // wait.  The trigger lets the NEXT kernel's CTAs do the same as soon as this grid's CTAs leave their SMs.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int rslot = 0;
            uint32_t rphase = 0;
            (void)rslot; (void)rphase;
            if (p.b_resident) {              // weights once per CTA: every K chunk of the (single) N tile
                mbar_expect_tx(&bres_bar, (uint32_t)(ksteps * b_bytes));
                for (int k = 0; k < ksteps; ++k) {
                    const int tap = k / p.kchunks, cc = k - tap * p.kchunks;
                    tma_load_2d(bres + (size_t)k * b_bytes, &tm.b, &bres_bar, tap * p.Cpad + cc * 64, 0);
                }
            }
            for (int t = t_first; t < t_end; t += t_step) {
                TILE_OF(t, nt, mt);
                const int img = mt / tiles_per_img;
                const int trem = mt - img * tiles_per_img;
                const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
                const int h0 = th * p.tsh + p.torg, w0 = tw * p.tsw + p.torg, n0 = nt * p.BN;
                if (p.l2_prefetch && nt == 0) {
                    // HBM-bound pointwise layers: the shared-memory ring holds < 100 KB per SM, not enough bytes in flight
                    // to cover the DRAM latency; pull the A patch of a tile several iterations ahead into L2
                    const int tp = t + p.l2_prefetch * (int)gridDim.x;
                    if (tp < p.ntiles) {
                        const int mp = tp / p.tilesN;
                        const int ip = mp / tiles_per_img;
                        const int rp = mp - ip * tiles_per_img;
                        const int hp = (rp / p.tilesW) * p.tsh + p.torg, wp = (rp % p.tilesW) * p.tsw + p.torg;
                        const int ho = PROD == PROD_DW ? 1 : 0;      // the halo box starts one pixel up / left
                        for (int cc = 0; cc < p.kchunks; ++cc) tma_prefetch_4d(&tm.a, cc * 64, wp - ho, hp - ho, ip);
                    }
                }
                if constexpr (PROD == PROD_DW) {
                    // halo patches {64 ch, TW+2, TH+2} into the raw ring (the compute warps turn them into A stages)
                    for (int k = 0; k < ksteps; ++k) {
                        TIMED_SPIN(w0c, &raw_empty[rslot], rphase ^ 1);
                        mbar_expect_tx(&raw_full[rslot], (uint32_t)p.raw_bytes);
                        tma_load_4d(rawbuf + (size_t)rslot * p.raw_bytes, &tm.a, &raw_full[rslot], k * 64, w0 - 1, h0 - 1, img);
                        if (++rslot == TC_RAW_SLOTS) { rslot = 0; rphase ^= 1; }
                    }
                } else if constexpr (PROD == PROD_SQ) {
                    // the patch lands in the A stage itself; the compute warps square it in place
                    for (int k = 0; k < ksteps; ++k) {
                        TIMED_SPIN(w0c, &empty_bar[stage], phase ^ 1);
                        uint8_t* sa = base + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&raw_full[stage], (uint32_t)TC_A_BYTES);
                        tma_load_4d(sa, &tm.a, &raw_full[stage], k * 64, w0, h0, img);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                } else if (p.halo) {
                    for (int cc = 0; cc < p.kchunks; ++cc) {
                        TIMED_SPIN(w0c, &empty_bar[stage], phase ^ 1);
                        uint8_t* sa = base + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)(128 * p.halo_w * p.halo_h));
                        tma_load_4d(sa, &tm.a, &full_bar[stage], cc * 64, w0 - p.pad, h0 - p.pad, img);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                } else {
                for (int k = 0; k < ksteps; ++k) {
                    TIMED_SPIN(w0c, &empty_bar[stage], phase ^ 1);
                    const int tap = k / p.kchunks, cc = k - tap * p.kchunks;
                    const int ky = tap / p.ks, kx = tap - ky * p.ks;
                    uint8_t* sa = base + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                    if (p.ck) tma_load_5d(sa, &tm.a, &full_bar[stage], cc * 64, w0, 0, h0 >> 1, img);
                    else tma_load_4d(sa, &tm.a, &full_bar[stage], cc * 64, w0 + kx - p.pad, h0 + ky - p.pad, img);
                    if (!p.b_resident && !p.bsplit) tma_load_2d(sa + p.a_bytes, &tm.b, &full_bar[stage], tap * p.Cpad + cc * 64, n0);      // (bsplit: warp 3)
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                }
            }
        }
    } else if (warp == 3) {
        // second TMA issuer (TcParams::bsplit): the weight tile of every K step.  One thread issuing both boxes of a K step was the
        // pace of the big 3x3 convs (busy, not waiting, ~900 clocks per K step against 512 clocks of MMA: tools/tc_bench.py role
        // clocks); a third issuer (the weight tile in halves) adds nothing
        if (lane == 0 && p.bsplit) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_first; t < t_end; t += t_step) {
                const int n0 = ((t) % p.tilesN) * p.BN;
                for (int k = 0; k < ksteps; ++k) {
                    mbar_wait<false>(&empty_bar[stage], phase ^ 1);
                    const int tap = k / p.kchunks, cc = k - tap * p.kchunks;
                    tma_load_2d(base + (size_t)stage * stage_bytes + p.a_bytes, &tm.b, &full_bar[stage], tap * p.Cpad + cc * 64, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            if (p.b_resident) { mbar_wait(&bres_bar, 0); tcgen05_fence_after(); }
            for (int t = t_first; t < t_end; t += t_step, ++it) {
                const int as = it & 1;
                TIMED_SPIN(w2c, &acc_empty[as], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(as * p.acc_stride);
                if (p.halo) {
                    const int taps = p.ks * p.ks;
                    for (int cc = 0; cc < p.kchunks; ++cc) {
                        TIMED_SPIN(w1c, &full_bar[stage], phase);
                        tcgen05_fence_after();
                        const uint32_t sa = smem_u32(base + (size_t)stage * stage_bytes);
                        // Measured on the final 192 -> 12 conv (N = 16): an M = 128, K = 16 MMA with A in shared memory costs ~124 clocks
                        // whatever N is (the A-operand feed, 4 KB per instruction), ~230 when consecutive MMAs switch taps; so the four
                        // K = 16 steps of one tap are issued back to back.  Per-tap accumulators (acc_split) do not help: the chain is
                        // not latency- but A-feed-bound.
                        for (int tap = 0; tap < taps; ++tap) {
                            const int ky = tap / p.ks, kx = tap - ky * p.ks;
                            const uint32_t a0 = sa + (uint32_t)((ky * p.halo_w + kx) * 128);
                            const uint64_t adesc = umma_desc_sw128_ex(a0, (uint32_t)(p.halo_w * 128), p.halo_bo ? ((a0 >> 7) & 7u) : 0u);
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(bres + (size_t)(tap * p.kchunks + cc) * b_bytes));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16(dcol + (uint32_t)(p.acc_split ? tap * p.BN : 0), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                          p.acc_split ? ((cc | kk) ? 1u : 0u) : ((cc | tap | kk) ? 1u : 0u));
                        }
                        tcgen05_commit(&empty_bar[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                } else
                for (int k = 0; k < ksteps; ++k) {
                    TIMED_SPIN(w1c, &full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(base + (size_t)stage * stage_bytes);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(p.b_resident ? smem_u32(bres + (size_t)k * b_bytes) : sa + (uint32_t)p.a_bytes);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)      // 4 x K=16 inside the 128-byte swizzle row: +32 B per step
                        umma_bf16(dcol, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                    tcgen05_commit(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(&acc_full[as]);
            }
        }
    } else if (PROD != PROD_TMA && warp >= 4 && warp < EW0) {
        // ---- fused A-operand producer warps (8 warps, 256 threads)
        const int pt = threadIdx.x - 128;                  // 0..255
        (void)pt;
        int stage = 0, rslot = 0;
        uint32_t phase = 0, rphase = 0;
        (void)rslot; (void)rphase;
        for (int t = t_first; t < t_end; t += t_step) {
            for (int k = 0; k < ksteps; ++k) {
                uint8_t* sa = base + (size_t)stage * stage_bytes;
                if constexpr (PROD == PROD_DW) {
                    // Tile = 8 rows x 16 columns.  Lane l owns channels (2l, 2l+1) of the 64-channel chunk (one bf16x2 word, so
                    // every shared-memory access of the warp is one contiguous 128-byte pixel row: conflict-free), producer
                    // warp pw owns output columns 2pw, 2pw+1.  The 10 x 4 input words of the strip are read once each and
                    // feed up to 9 packed fp32x2 FMAs (FFMA2) against the 9 filter taps held in registers.
                    TIMED_WAIT(w0c, &raw_full[rslot], rphase);         // halo patch landed
                    TIMED_WAIT(w1c, &empty_bar[stage], phase ^ 1);     // A stage free (its MMAs retired)
                    const int pw = warp - 4;
                    const uint32_t rp = smem_u32(rawbuf) + (uint32_t)(rslot * p.raw_bytes + ((2 * pw) * 32 + lane) * 4);
                    const uint32_t wk = smem_u32(sDw) + (uint32_t)((k * 64 + 2 * lane) * 4);         // tap t at wk + t * Cin * 4
                    const uint32_t sa_s = smem_u32(sa);
                    const bool chok = k * 64 + 2 * lane < p.Cin;        // channel tail of the last chunk: taps and bias 0 (the patch is zero-filled there)
                    float2 w2[9];
#pragma unroll
                    for (int tp = 0; tp < 9; ++tp) w2[tp] = chok ? lds_f2(wk + (uint32_t)(tp * p.Cin * 4)) : make_float2(0.f, 0.f);
                    const float2 b2 = chok ? lds_f2(wk + (uint32_t)(9 * p.Cin * 4)) : make_float2(0.f, 0.f);
                    float2 acc[8][2];
#pragma unroll
                    for (int oy = 0; oy < 8; ++oy) { acc[oy][0] = b2; acc[oy][1] = b2; }
                    // the shared-memory accessors are volatile asm (program order is kept), so the loads of input row iy + 1 are
                    // written BEFORE the arithmetic of row iy: their latency hides behind the 18 FFMA2 of the current row
                    uint32_t wv[2][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) wv[0][j] = lds32(rp + (uint32_t)(j * 128));
#pragma unroll
                    for (int iy = 0; iy < 10; ++iy) {
                        if (iy + 1 < 10) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) wv[(iy + 1) & 1][j] = lds32(rp + (uint32_t)(((iy + 1) * 18 + j) * 128));
                        }
                        float2 x[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t wq = wv[iy & 1][j];
                            x[j] = make_float2(__uint_as_float(wq << 16), __uint_as_float(wq & 0xffff0000u));
                        }
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int oy = iy - ky;
                            if (oy >= 0 && oy < 8) {
#pragma unroll
                                for (int c = 0; c < 2; ++c)
#pragma unroll
                                    for (int kx = 0; kx < 3; ++kx) acc[oy][c] = __ffma2_rn(x[c + kx], w2[ky * 3 + kx], acc[oy][c]);
                            }
                        }
                        if (iy >= 2) {
                            const int oy = iy - 2;
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int r = oy * 16 + 2 * pw + c;    // pixel of the tile = A row
                                __nv_bfloat162 hv = __floats2bfloat162_rn(acc[oy][c].x, acc[oy][c].y);
                                sts32(sa_s + (uint32_t)(r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)), *reinterpret_cast<uint32_t*>(&hv));
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> UMMA (async proxy) reads
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&full_bar[stage]); mbar_arrive(&raw_empty[rslot]); }
                    if (++rslot == TC_RAW_SLOTS) { rslot = 0; rphase ^= 1; }
                } else {                                               // PROD_SQ: square the landed A stage in place
                    mbar_wait(&raw_full[stage], phase);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t q4 = smem_u32(sa) + (uint32_t)((pt + 256 * i) * 16);      // layout-agnostic: element-wise
                        float v[8];
                        unpack8_bf16(lds128(q4), v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = v[j] * v[j];
                        sts128(q4, pack8_bf16(v));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= EW0) {
        // epilogue warps: TMEM lane quarter q = warp % 4 (hardware restriction), column quarter cq = (warp - EW0) / 4:
        // four warps per SM sub-partition hide the TMEM-load / shared-memory latencies of the short per-block chain
        const int q = warp & 3, cq = (warp - EW0) >> 2;
        const int r = q * 32 + lane;
        uint8_t* stg = base + (size_t)p.stages * stage_bytes;
        const int per = (e.out2 || GDN != GDN_NONE) ? 2 : 1;      // staging buffers per 64-column block (out/res [+ out2 | GDN operand])
        const uint32_t nring = p.store_mode == STORE_TMA ? (uint32_t)(p.nstg / (((p.BN + 63) / 64) * per)) : 1u;   // slots per group
        uint32_t blk = 0;
        int it = 0;
        // narrow GEMMs (BN <= 128): `esplit` warp groups share one 64-column block (its staging slot, named barrier and store issuer),
        // each taking 64 / esplit columns, so all 16 epilogue warps work instead of 4 or 8 (the epilogue of a 128 x 64 tile was
        // 2.9 k clocks on 4 warps against 1 k clocks of MMA: tools/em_bench.py)
        const int esplit = p.store_mode == STORE_TMA ? p.esplit : 1;
        const int eb = cq / esplit, esub = cq - eb * esplit;                // 64-column block of this group, its share of it
        const bool idle = p.store_mode == STORE_TMA && eb * 64 >= p.BN;     // this warp group owns no columns
        // shift-sum mode: the (interior pixel, output) pairs this thread sums in phase 2 are the same for every tile
        // (packed into one word per pair: py[0:4) px[4:7) n[7:11) ch[11:13) rr[13] s[14] ok[15] pofs[16:32))
        uint32_t ss_pk[2] = {0u, 0u};
        if constexpr (ACT == ACT_NONE && GDN == GDN_NONE && !RES && PROD == PROD_TMA) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int idx = (warp - EW0) * 32 + lane + k * NEPI * 32;
                const int j = idx % 12, rest = idx / 12;
                const int py = 1 + rest % 14, cr = rest / 14;
                const int rr = cr & 1, ch = cr >> 1, px = 1 + (j >> 1), sx = j & 1;
                const int n = (2 * rr + sx) * 3 + ch;
                const int pofs = ((py - 1) * 8 + (px - 1)) * TC_SS_PITCH + n;
                if (p.store_mode == STORE_SS && idx < 84 * 12)
                    ss_pk[k] = (uint32_t)py | ((uint32_t)px << 4) | ((uint32_t)n << 7) | ((uint32_t)ch << 11) | ((uint32_t)rr << 13) |
                               ((uint32_t)sx << 14) | (1u << 15) | ((uint32_t)pofs << 16);
            }
        }
        for (int t = t_first; t < t_end && !idle; t += t_step, ++it) {
            TILE_OF(t, nt, mt);
            const int img = mt / tiles_per_img;
            const int trem = mt - img * tiles_per_img;
            const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
            const int h0 = th * p.tsh + p.torg, w0 = tw * p.tsw + p.torg;
            const int h = h0 + r / p.TW, w = w0 + r % p.TW;
            const int n0 = nt * p.BN;
            const bool valid = h < e.Hout && w < e.Wout;
            const int as = it & 1;
            EpiRow row;
            row.keep_pre = parity_keep(e.premask, h, w);
            row.h = h; row.w = w;
            row.keep_post = parity_keep(e.postmask, h, w);
            row.xp = (GDN != GDN_NONE && valid) ? reinterpret_cast<const bf16*>(e.gdn_x) + (((size_t)img * e.Hout + h) * e.Wout + w) * e.gdn_ld : nullptr;
            if (p.store_mode != STORE_TMA) {         // (TMA-store mode waits after its operand prefetch)
                TIMED_WAIT(w3c, &acc_full[as], (uint32_t)(it >> 1) & 1u);
                tcgen05_fence_after();
            }
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.acc_stride);
            if (p.store_mode == STORE_TMA) {
                // One 64-column block per warp GROUP (4 warps = the 4 TMEM lane quarters): group cq owns columns
                // [64 cq, 64 cq + 64) of every tile, its own swizzled 16 KB staging slot(s), its own named barrier and its
                // own TMA-store issuer, so the blocks of a tile drain concurrently instead of one after the other.
                const int kb = eb * 64;
                if (kb < p.BN) {
                    int g = 0, ocb = n0 + kb;                       // the whole block lies in one pixel-shuffle group
                    if (e.shuffle) { const int Cq = e.N >> 2; g = ocb / Cq; ocb -= g * Cq; }
                    const bool gissuer = (esub == 0 && q == 0 && lane == 0);
                    const int gthreads = 128 * esplit;
                    const int pr0 = esub * (4 / esplit), pr1 = pr0 + 4 / esplit;
                    uint8_t* sb = stg + (size_t)((eb * nring + (blk % nring)) * per) * TC_STG_BYTES;
                    const uint32_t sb_s = smem_u32(sb);
                    long long tq0 = dbg ? clock64() : 0;
                    if (nring == 1) {               // single slot: this group's previous store must have drained it
                        if (gissuer) tma_store_wait_read(0);
                        asm volatile("bar.sync %0, %1;" ::"r"(eb + 1), "r"(gthreads) : "memory");
                    }
                    uint64_t* sbar = &stg_bar[eb][blk % nring];
                    if constexpr (RES || GDN != GDN_NONE) {
                        // Operand prefetch (no pixel shuffle in this mode): the residual / GDN-operand block of this tile is one TMA
                        // box each, landing in the staging slot in the swizzled layout of the output block (out-of-range pixels and
                        // columns arrive as zeros); the copy is in flight while the group waits for the accumulator.  The result
                        // later overwrites the residual in place.  The slot is free here: its last TMA store was drained by the
                        // issuer's wait_group.read before the barrier above (one slot) / before the previous tile's store (two).
                        if (gissuer) {
                            mbar_expect_tx(sbar, (uint32_t)(((RES ? 1 : 0) + (GDN != GDN_NONE ? 1 : 0)) * TC_STG_BYTES));
                            if constexpr (RES) tma_load_4d(sb, &tm.r, sbar, n0 + kb, w0, h0, img);
                            if constexpr (GDN != GDN_NONE) tma_load_4d(sb + TC_STG_BYTES, &tm.gx, sbar, n0 + kb, w0, h0, img);
                        }
                    }
                    if (dbg) { long long _t = clock64(); mbar_wait(&acc_full[as], (uint32_t)(it >> 1) & 1u); w3c += clock64() - _t; }
                    else mbar_wait(&acc_full[as], (uint32_t)(it >> 1) & 1u);
                    tcgen05_fence_after();
                    if constexpr (RES || GDN != GDN_NONE) mbar_wait(sbar, (blk / nring) & 1u);
                    long long tq1 = dbg ? clock64() : 0;
#pragma unroll 1
                    for (int pr = pr0; pr < pr1; ++pr) {
                        // two 8-column groups are fetched from TMEM before any arithmetic
                        const int c = kb + pr * 16;
                        uint32_t raw[2][8];
                        tmem_ld8(trow + (uint32_t)c, raw[0]);
                        tmem_ld8(trow + (uint32_t)(c + 8), raw[1]);
                        tmem_ld_wait();
                        if (pr == pr1 - 1) {        // accumulator fully read by this warp: hand the TMEM stage back
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[as]);
                        }
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[sub][j]);
                            const int jj = pr * 2 + sub;
                            const uint32_t off = (uint32_t)(r * 128 + ((jj ^ (r & 7)) << 4));
                            epi_math8<ACT, GDN, RES, true>(e, sBias, row, n0 + c + sub * 8, ocb + pr * 16 + sub * 8, v, true,
                                                           sb_s + TC_STG_BYTES + off, sb_s + off);
                            sts128(sb_s + off, pack8_bf16(v));
                            if (e.out2) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = v[j] * v[j];
                                sts128(sb_s + TC_STG_BYTES + off, pack8_bf16(v));
                            }
                        }
                    }
                    long long tq2 = dbg ? clock64() : 0;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    // two slots: after this barrier the group's OTHER slot is rewritten next tile; its store must have drained
                    if (nring > 1 && gissuer) tma_store_wait_read(0);
                    asm volatile("bar.sync %0, %1;" ::"r"(eb + 1), "r"(gthreads) : "memory");
                    if (gissuer && !(p.debug & 1)) {
                        tma_store_4d(&tm.o[g], sb, ocb, w0, h0, img);
                        if (e.out2) tma_store_4d(&tm.o2, sb + TC_STG_BYTES, ocb, w0, h0, img);
                        tma_store_commit();
                    }
                    if (dbg) { long long tq3 = clock64(); w0c += tq1 - tq0; w1c += tq2 - tq1; w2c += tq3 - tq2; }
                    ++blk;
                }
            } else if (p.store_mode == STORE_SS) {
                // Shift-sum 3x3 subpel conv with 12 GEMM columns per tap (the final 192 -> 3x2x2 layer).  The accumulator holds the
                // per-tap partial products of the 16 x 8 PATCH pixels (lane = patch pixel, column = tap*12 + n); the output of an
                // interior pixel p is sum_tap P[p + off(tap)][tap].  Phase 1: TMEM -> shared memory (fp32, odd row pitch);
                // phase 2: the 14 x 6 interior pixels x 12 outputs are summed over the 9 taps and written as fp32 NCHW with the
                // pixel shuffle, 12 consecutive floats of one output row per 12 consecutive threads.
                constexpr int SSP = TC_SS_PITCH;
                const uint32_t Ps = smem_u32(stg);      // one buffer (the operand ring needs the space: two tiles of A in flight)
                {
                    // 28 columns per warp group, as seven 16-byte shared stores (explicit st.shared: the generic stores of the first
                    // version went through the global-memory queue, "lg throttle" on every one of them; profiles/r02_ncu_final.txt)
                    const uint32_t prow = Ps + (uint32_t)((r * SSP + cq * 28) * 4);
                    const uint32_t tc0 = trow + (uint32_t)(cq * 28);
                    uint32_t v0[16], v1[8], v2[8];
                    tmem_ld16(tc0, v0);
                    tmem_ld8(tc0 + 16u, v1);
                    tmem_ld8(tc0 + 20u, v2);          // columns 20..27 (overlaps 20..23 of v1: one ld shape less; columns >= 108 are never read)
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts128(prow + (uint32_t)(j * 16), make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]));
                    sts128(prow + 64u, make_uint4(v1[0], v1[1], v1[2], v1[3]));
                    sts128(prow + 80u, make_uint4(v2[0], v2[1], v2[2], v2[3]));
                    sts128(prow + 96u, make_uint4(v2[4], v2[5], v2[6], v2[7]));
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
                asm volatile("bar.sync 9, %0;" ::"r"(NEPI * 32) : "memory");
                if (!(p.debug & 1)) {
                    float* o = reinterpret_cast<float*>(e.out);
                    const int OH = 2 * e.Hout, OW = 2 * e.Wout;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (ss_pk[k] & (1u << 15)) {
                            const int gh = h0 + (int)(ss_pk[k] & 15u), gw = w0 + (int)((ss_pk[k] >> 4) & 7u);
                            if (gh < e.Hout && gw < e.Wout) {
                                const uint32_t pp = Ps + (ss_pk[k] >> 16) * 4u;
                                float t9[9];
#pragma unroll
                                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                                    for (int kx = 0; kx < 3; ++kx) t9[ky * 3 + kx] = __uint_as_float(lds32(pp + (uint32_t)(((ky * 8 + kx) * SSP + (ky * 3 + kx) * 12) * 4)));
                                float acc = sBias[(ss_pk[k] >> 7) & 15u];
#pragma unroll
                                for (int q9 = 0; q9 < 9; ++q9) acc += t9[q9];          // tap order, as the accumulating MMA chain of the plain form
                                o[(((size_t)img * 3 + ((ss_pk[k] >> 11) & 3u)) * OH + 2 * gh + ((ss_pk[k] >> 13) & 1u)) * OW + 2 * gw + ((ss_pk[k] >> 14) & 1u)] = acc;
                            }
                        }
                    }
                }
                asm volatile("bar.sync 9, %0;" ::"r"(NEPI * 32) : "memory");      // P is rewritten by the next tile
                continue;          // (accumulator already released)
            } else if (p.store_mode == STORE_NCHW3) {
                // final subpel conv (N = 12 -> 3 channels): column (2r+s)*3 + ch -> out[b][ch][2h+r][2w+s], fp32 NCHW;
                // consecutive lanes hold consecutive w, so every float2 store instruction writes whole 128-byte lines.
                if (cq == 0) {
                    uint32_t raw[16];
                    tmem_ld16(trow, raw);
                    tmem_ld_wait();
                    if (p.acc_split) {               // sum the per-tap partial accumulators (tap order, fixed)
#pragma unroll 1
                        for (int tap = 1; tap < p.ks * p.ks; ++tap) {
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {         // columns 12..15 are padding: 8 + 4 would do, 8 + 8 keeps one ld shape
                                uint32_t part[8];
                                tmem_ld8(trow + (uint32_t)(tap * p.BN + hf * 8), part);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 8; ++j) raw[hf * 8 + j] = __float_as_uint(__uint_as_float(raw[hf * 8 + j]) + __uint_as_float(part[j]));
                            }
                        }
                    }
                    if (valid && !(p.debug & 1)) {
                        float* o = reinterpret_cast<float*>(e.out);
                        const int OH = 2 * e.Hout, OW = 2 * e.Wout;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr)
                                *reinterpret_cast<float2*>(o + (((size_t)img * 3 + ch) * OH + 2 * h + rr) * OW + 2 * w) =
                                    make_float2(__uint_as_float(raw[(2 * rr) * 3 + ch]) + sBias[(2 * rr) * 3 + ch],
                                                __uint_as_float(raw[(2 * rr + 1) * 3 + ch]) + sBias[(2 * rr + 1) * 3 + ch]);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            } else {
                // STORE_DIRECT: 16-column chunks are dealt to the four column quarters in turn (BN is a multiple of 16)
                for (int c0 = cq * 16; c0 < p.BN; c0 += 16 * (NEPI / 4)) {
#pragma unroll 1
                    for (int sub = 0; sub < 2; ++sub) {
                        const int c = c0 + sub * 8, n = n0 + c;
                        uint32_t raw[8];
                        tmem_ld8(trow + (uint32_t)c, raw);
                        tmem_ld_wait();
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[j]);
                        int oc = n;
                        row.rp = nullptr;
                        if (RES && valid) {
                            int oh, ow, OH, OW;
                            out_coord(e, h, w, n, oh, ow, oc, OH, OW);
                            row.rp = reinterpret_cast<const bf16*>(e.res) + (((size_t)img * OH + oh) * OW + ow) * e.res_ld;
                        }
                        epi_math8<ACT, GDN, RES, false>(e, sBias, row, n, oc, v, p.ld_vec != 0);
                        if (valid && !(p.debug & 1)) direct_store8(e, img, h, w, n, v, p.epi_vec != 0);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            }
        }
        if (p.store_mode == STORE_TMA && q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (dbg && lane == 0) {
        unsigned long long* d = dbg + (size_t)blockIdx.x * 8;
        if (warp == 0) { d[0] = (unsigned long long)(clock64() - t_start); d[1] = (unsigned long long)w0c; }
        if (warp == 1) { d[2] = (unsigned long long)w1c; d[3] = (unsigned long long)w2c; }
        if (PROD != PROD_TMA && warp == 4) { d[6] = (unsigned long long)w0c; d[7] = (unsigned long long)w1c; }
        if (PROD != PROD_TMA && warp == EW0) { d[4] = (unsigned long long)w3c; d[5] = (unsigned long long)(clock64() - t_start);
                         dbg[(size_t)gridDim.x * 8 + blockIdx.x] = (unsigned long long)w2c; dbg[(size_t)gridDim.x * 9 + blockIdx.x] = (unsigned long long)w0c;
                         dbg[(size_t)gridDim.x * 10 + blockIdx.x] = (unsigned long long)w1c; }
        if (PROD == PROD_TMA && warp == EW0) { d[4] = (unsigned long long)w3c; d[5] = (unsigned long long)(clock64() - t_start);
                         d[6] = (unsigned long long)w0c; d[7] = (unsigned long long)w1c; dbg[(size_t)gridDim.x * 8 + blockIdx.x] = (unsigned long long)w2c; }
    }
#undef TIMED_WAIT
#undef TIMED_SPIN
#undef TILE_OF
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.acc_stride))
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Stand-alone depthwise 3x3 (pad 1, stride 1 | 2) + bias (+ GELU), bf16 NHWC, TMA-fed and persistent.
// Same arithmetic scheme as the fused producer above (lane = channel pair, FFMA2 against 9 taps in registers, warp pw owns
// output columns 2pw, 2pw+1), but the halo patch {64 ch, TW*S+2, TH*S+2} of every (tile, 64-channel chunk) item arrives
// through a ring of TMA boxes (zero fill = conv padding and channel tail), so loads of the next items are in flight while
// the 8 compute warps work, and results go straight to global memory (one full 128-byte line per warp store).
// ------------------------------------------------------------------------------------------
// RH = 2 (development switch): 16 compute warps per block, warp = (column pair, half of the tile's rows), 32 compute warps per SM instead
// of 16.  Measured SLOWER (640 channels at 32 x 68 x 120: 183 -> 224 us; stride 2 at 4 x 544 x 960 x 192: 169 -> 205 us): as in the
// two-SM fused kernel, more warps do not help -- the kernel is not short of ready warps (tools/dw_bench.py).
#ifndef MLIC_DW_RH
#define MLIC_DW_RH 1
#endif
template <int S> struct DwT { static constexpr int TH = S == 1 ? 8 : 4, TW = 16, IH = TH * S + 2, IW = TW * S + 2, NX = S == 1 ? 4 : 5,
                                                   RH = MLIC_DW_RH, R = TH / RH, IHW = (R - 1) * S + 3, NCW = 8 * RH, THREADS = (NCW + 1) * 32,
                                                   SLOTS = S == 1 ? (RH == 2 ? 3 : 4) : 2, BYTES = IH * IW * 128; };
template <int S, bool GELU>
__global__ void __launch_bounds__(DwT<S>::THREADS, 2) dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmap, int C, bf16* __restrict__ out, int Ho,
                                                            int Wo, int old, const float* __restrict__ w9, const float* __restrict__ bias,
                                                            int act, int tilesW, int tilesH, int chunks, int nitems) {
    using TT = DwT<S>;
    extern __shared__ uint8_t dw_smem_raw[];
    uint8_t* ring = (uint8_t*)(((uintptr_t)dw_smem_raw + 127) & ~(uintptr_t)127);
    __shared__ uint64_t full_bar[TT::SLOTS], empty_bar[TT::SLOTS];
    __shared__ int4 item_desc[TT::SLOTS];          // (channel chunk, image, first output row, first output column) of the item in a slot
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TT::SLOTS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], TT::NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncthreads();
    pdl_wait();                 // (everything above touches parameters only)
    const int tiles_per_img = tilesW * tilesH;
    if (warp == TT::NCW) {
        if (lane == 0) {
            int n = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
                const int slot = n % TT::SLOTS;
                const uint32_t ph = (uint32_t)(n / TT::SLOTS) & 1u;
                mbar_wait(&empty_bar[slot], ph ^ 1u);
                const int cc = it % chunks, tile = it / chunks;
                const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
                const int th = tr / tilesW, tw = tr - th * tilesW;
                item_desc[slot] = make_int4(cc, b, th * TT::TH, tw * TT::TW);      // published by the arrive below (release) / the wait (acquire)
                mbar_expect_tx(&full_bar[slot], (uint32_t)TT::BYTES);
                tma_load_4d(ring + (size_t)slot * TT::BYTES, &tmap, &full_bar[slot], cc * 64, tw * TT::TW * S - 1, th * TT::TH * S - 1, b);
            }
        }
        return;
    }
    const int pw = warp & 7, rh = warp >> 3;
    const size_t row_stride = (size_t)Wo * old;
    int n = 0;
    // the chunk index advances by gridDim.x (mod chunks) per item: no division in the item loop; the tile coordinates come from the
    // descriptor the TMA thread left in the slot (the three integer divisions per item and warp were ~15 % of the instructions)
    const int cstep = (int)(gridDim.x % (unsigned)chunks);
    int ccur = (int)(blockIdx.x % (unsigned)chunks);
    // the 9 taps + bias of the channel pair come from global memory (L2).  The launcher makes the grid a multiple of the chunk count
    // whenever it can: cstep is then 0, every block stays on ONE 64-channel chunk and loads its taps once (per item, the ten loads with
    // their 64-bit address arithmetic were ~100 of the ~700 warp instructions of an item)
    float2 w2[9], b2;
    auto load_w = [&](int cq) {
        const int cn = cq * 64 + 2 * lane;
        const bool ok = cn < C;
#pragma unroll
        for (int t = 0; t < 9; ++t) w2[t] = ok ? __ldg(reinterpret_cast<const float2*>(w9 + (size_t)t * C + cn)) : make_float2(0.f, 0.f);
        b2 = ok ? __ldg(reinterpret_cast<const float2*>(bias + cn)) : make_float2(0.f, 0.f);
    };
    load_w(ccur);
    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
        const int slot = n % TT::SLOTS;
        const uint32_t ph = (uint32_t)(n / TT::SLOTS) & 1u;
        if (cstep != 0 && n > 0) {
            ccur += cstep;
            if (ccur >= chunks) ccur -= chunks;
            load_w(ccur);
        }
        mbar_wait(&full_bar[slot], ph);
        const int4 ds = item_desc[slot];
        const int cc = ds.x, b = ds.y, oh0 = ds.z + rh * TT::R, ow0 = ds.w;
        const int c = cc * 64 + 2 * lane;
        const bool cok = c < C;
        // output addressing once per item (the per-store 64-bit index arithmetic and bounds tests were a third of this kernel's
        // instructions: profiles/r01_ncu_dwtma.txt): row pointer of the warp's two columns, advanced by one output row per step
        const int ow = ow0 + 2 * pw;
        bf16* orow = out + (((size_t)b * Ho + oh0) * Wo + ow) * old + c;
        const int nrows = cok ? Ho - oh0 : 0;                                   // rows of this warp's share inside the image (<= 0: none / channel tail)
        const bool okq[2] = {ow < Wo, ow + 1 < Wo};
        // explicit shared-state-space loads with immediate offsets (the first version read the ring through a generic pointer: generic LD plus
        // R2UR / IMAD / LEA address arithmetic were 38 % of the kernel's warp instructions, profiles/r02_ncu_dwtma.txt), the words of input row
        // iy + 1 requested before the arithmetic of row iy
        const uint32_t rp = smem_u32(ring) + (uint32_t)(slot * TT::BYTES + (((rh * TT::R * S) * TT::IW + 2 * pw * S) * 32 + lane) * 4);
        float2 acc[TT::R][2];
#pragma unroll
        for (int oy = 0; oy < TT::R; ++oy) { acc[oy][0] = b2; acc[oy][1] = b2; }
        uint32_t wv[2][TT::NX];
#pragma unroll
        for (int j = 0; j < TT::NX; ++j) wv[0][j] = lds32(rp + (uint32_t)(j * 128));
#pragma unroll
        for (int iy = 0; iy < TT::IHW; ++iy) {
            if (iy + 1 < TT::IHW) {
#pragma unroll
                for (int j = 0; j < TT::NX; ++j) wv[(iy + 1) & 1][j] = lds32(rp + (uint32_t)(((iy + 1) * TT::IW + j) * 128));
            }
            float2 x[TT::NX];
#pragma unroll
            for (int j = 0; j < TT::NX; ++j) x[j] = bf2_to_f2(wv[iy & 1][j]);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                if ((iy - ky) >= 0 && ((iy - ky) % S) == 0 && (iy - ky) / S < TT::R) {
                    const int oy = (iy - ky) / S;
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) acc[oy][q] = __ffma2_rn(x[q * S + kx], w2[ky * 3 + kx], acc[oy][q]);
                }
            }
            if (iy >= 2 && ((iy - 2) % S) == 0) {
                const int oy = (iy - 2) / S;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float2 v = acc[oy][q];
                    if (GELU) v = gelu2(v);
                    if (oy < nrows && okq[q]) {
                        __nv_bfloat162 hv = __floats2bfloat162_rn(v.x, v.y);
                        *reinterpret_cast<uint32_t*>(orow + (q ? old : 0)) = *reinterpret_cast<uint32_t*>(&hv);
                    }
                }
                orow += row_stride;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
}

template <int S>
static int launch_dw_tma(const Act& in, const Act& out, const float* w9, const float* bias, int act, cudaStream_t s) {
    using TT = DwT<S>;
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.B};
    cuuint64_t strides[3] = {(cuuint64_t)in.ld * 2, (cuuint64_t)in.W * in.ld * 2, (cuuint64_t)in.H * in.W * in.ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TT::IW, (cuuint32_t)TT::IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(dw) failed: %d", (int)r); return 2; }
    const int tilesW = (out.W + TT::TW - 1) / TT::TW, tilesH = (out.H + TT::TH - 1) / TT::TH, chunks = (in.C + 63) / 64;
    const long long nitems = (long long)out.B * tilesW * tilesH * chunks;
    if (nitems <= 0 || nitems > 0x7fffffffLL) return 3;
    const int smem = TT::SLOTS * TT::BYTES + 128;
    const int dev = cur_dev();
    static bool attr[TC_MAX_DEV] = {};
    if (!attr[dev]) {
        cudaFuncSetAttribute(dwconv3x3_tma_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(dwconv3x3_tma_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr[dev] = true;
    }
    const int num_sms = dev_sms(dev);
    int grid = (int)std::min<long long>(nitems, 2LL * num_sms);
    if (grid >= 2 * chunks) grid = grid / chunks * chunks;        // a block then stays on one channel chunk: taps loaded once (see the kernel)
    if (act == ACT_GELU) launch_k(dwconv3x3_tma_kernel<S, true>, dim3(grid), dim3(TT::THREADS), smem, s, tmap, in.C, (bf16*)out.p, out.H, out.W, out.ld, w9, bias, act, tilesW, tilesH, chunks, (int)nitems);
    else launch_k(dwconv3x3_tma_kernel<S, false>, dim3(grid), dim3(TT::THREADS), smem, s, tmap, in.C, (bf16*)out.p, out.H, out.W, out.ld, w9, bias, act, tilesW, tilesH, chunks, (int)nitems);
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
// bf16 NHWC depthwise 3x3 through the TMA-fed kernel; non-zero: not taken (the caller falls back to the staged kernel)
int launch_dwconv3x3_tma(const Act& in, const Act& out, const float* w9, const float* bias, int stride, int act, cudaStream_t s) {
    if (tc_init()) return 1;
    if ((in.C % 8) != 0 || (in.ld % 8) != 0 || (out.ld % 2) != 0 || ((uintptr_t)in.p % 16) != 0 || ((uintptr_t)out.p % 4) != 0) return 1;
    if (stride == 1) return launch_dw_tma<1>(in, out, w9, bias, act, s);
    if (stride == 2) return launch_dw_tma<2>(in, out, w9, bias, act, s);
    return 1;
}

// ------------------------------------------------------------------------------------------ host side
static int pick_bn(int N, int mult) {
    int nt = (N + 255) / 256;
    int bn = ((N + nt - 1) / nt + mult - 1) / mult * mult;
    return bn < mult ? mult : bn;
}

bool tc_conv_supported(const TcConv& c, const Epi& e) {
    if (!c.in || !c.w) return false;
    if (((uintptr_t)c.in) % 16 != 0 || (c.ld % 8) != 0) return false;      // TMA: 16-byte base / strides
    if (c.sW <= 0 || c.sH <= 0 || (c.sW % 8) != 0 || (c.sH % 8) != 0 || (c.sB % 8) != 0) return false;
    if (c.Cin < 8 || c.Cpad % 64 != 0) return false;
    if (e.N < 8 || e.N > TC_MAX_N) return false;
    if (c.H <= 0 || c.W <= 0 || c.B <= 0) return false;
    if (c.ck && !(c.prod == PROD_TMA && c.ks == 1 && c.pad == 0 && !c.ss && (c.H % 2) == 0 && (c.W % 2) == 0 && e.Hout == c.H && e.Wout == c.W / 2 &&
                  c.sW == c.ld && c.sH == c.W * c.ld)) return false;
    if (c.ss) return c.prod == PROD_TMA && c.ks == 1 && c.pad == 0 && e.N == 108 && e.nchw && e.shuffle && e.out_f32 && !e.res && !e.gdn &&
                     e.act == ACT_NONE && !e.premask && !e.postmask && c.Cpad <= 256;
    if (e.nchw && !(e.shuffle && e.N == 12 && e.out_f32)) return false;
    if (c.prod != PROD_TMA) {
        // fused producers: 1x1 GEMM over a dense-stride-1 view, whole weight matrix resident, <= 3 column groups
        // (depthwise producer: a channel tail is fine, the halo patch arrives zero-filled and the producer zeroes its taps: Cin % 8 for TMA)
        if (c.ks != 1 || c.pad != 0 || (c.Cin % (c.prod == PROD_DW ? 8 : 64)) != 0 || e.N > 192 || (e.N % 8) != 0 || e.shuffle || e.out_f32 || e.out2) return false;
        if ((c.Cpad / 64) > 4 || (size_t)(c.Cpad / 64) * ((e.N + 63) / 64 * 64) * 128 > 80 * 1024) return false;
        if (c.prod == PROD_DW && (!c.dw_w9 || !c.dw_bias)) return false;
        if (c.prod == PROD_DW && e.gdn) return false;
        if (c.prod == PROD_DW && e.act == ACT_HALF_TANH && !e.res) return false;      // (instantiated with the residual only: the LRP tail)
        if (c.prod == PROD_SQ && (!e.gdn || e.act != ACT_NONE)) return false;
        auto ok16 = [](const void* q, int ld) { return q == nullptr || ((((uintptr_t)q) % 16 == 0) && ((ld * 2) % 16 == 0)); };
        if (!e.out || !ok16(e.out, e.out_ld) || !ok16(e.res, e.res_ld) || !ok16(e.gdn_x, e.gdn_ld)) return false;
    }
    return true;
}

static int encode_out_map(CUtensorMap* m, void* basep, int Cdim, int Wd, int Hd, int Bd, size_t sW, size_t sH, size_t sB,
                          int TW, int TH) {
    cuuint64_t dims[4] = {(cuuint64_t)Cdim, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Bd};
    cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(out) failed: %d (C=%d W=%d H=%d B=%d)", (int)r, Cdim, Wd, Hd, Bd);
        return 1;
    }
    return 0;
}

int launch_conv_gemm_tc(const TcConv& c, const Epi& e, int vec, cudaStream_t s) {
    if (tc_init()) return 1;
    TcParams p;
    memset(&p, 0, sizeof p);
    // output patch shape: minimise the number of tiles
    const int cand[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
    long long best = -1;
    for (int i = 0; i < (c.prod == PROD_DW ? 1 : (c.ck ? 4 : 5)); ++i) {       // the depthwise producer is written for the 8 x 16 patch; ck: even TH
        long long t = (long long)((e.Hout + cand[i][0] - 1) / cand[i][0]) * ((e.Wout + cand[i][1] - 1) / cand[i][1]);
        if (best < 0 || t < best) { best = t; p.TH = cand[i][0]; p.TW = cand[i][1]; }
    }
    p.ks = c.ks; p.pad = c.pad; p.Cpad = c.Cpad; p.kchunks = c.Cpad / 64;
    p.a_bytes = TC_A_BYTES;
    if (c.ss) { p.TH = 16; p.TW = 8; }                 // shift-sum mode: 16 x 8 patches stepping by 14 x 6, origin -1
    p.tsh = c.ss ? p.TH - 2 : p.TH; p.tsw = c.ss ? p.TW - 2 : p.TW; p.torg = c.ss ? -1 : 0;
    p.tilesH = (e.Hout + p.tsh - 1) / p.tsh;
    p.tilesW = (e.Wout + p.tsw - 1) / p.tsw;

    auto ok16 = [](const void* q, int ld, int esz) { return q == nullptr || ((((uintptr_t)q) % 16 == 0) && ((ld * esz) % 16 == 0)); };
    const int Cq = e.shuffle ? e.N / 4 : e.N;
    // store mode
    p.store_mode = STORE_DIRECT;
    if (c.ss) p.store_mode = STORE_SS;
    else if (e.nchw) p.store_mode = STORE_NCHW3;
    else if (!e.out_f32 && e.out && ok16(e.out, e.out_ld, 2) && ok16(e.out2, e.out2_ld, 2) && (!e.shuffle || (Cq % 64) == 0) && !(e.shuffle && e.out2) &&
             !(e.out2 && e.gdn) &&
             (!(e.res || e.gdn) || (!e.shuffle && (e.N % 8) == 0 && ok16(e.res, e.res_ld, 2) && ok16(e.gdn_x, e.gdn_ld, 2))))
        p.store_mode = STORE_TMA;
    { const char* d = getenv("MLIC_TC_DEBUG"); p.debug = d ? atoi(d) : 0; }
    if (p.debug & 4) { if (p.store_mode == STORE_TMA) p.store_mode = STORE_DIRECT; }
    p.BN = pick_bn(e.N, p.store_mode == STORE_TMA ? 64 : ((e.shuffle || e.N % 32 == 0) && e.N >= 32 ? 32 : 16));
    // Halo-patch A operand (ks > 1, narrow N with the whole weight matrix resident): 16 x 8 output patch, one TMA box per chunk.
    // MLIC_HALO (development): bit0 enable, bit1 base-offset descriptor variant, bit2 pad the staged patch width to 16 pixels.
    int halo_mode = 1;
    { const char* hv = getenv("MLIC_HALO"); if (hv) halo_mode = atoi(hv); }
    {
        const long long bres_guess = (long long)c.ks * c.ks * p.kchunks * p.BN * 128;
        if ((halo_mode & 1) && c.prod == PROD_TMA && c.ks > 1 && p.BN >= e.N && bres_guess <= 120 * 1024) {
            p.halo = 1; p.TH = 16; p.TW = 8;
            p.halo_w = (halo_mode & 4) ? 16 : p.TW + c.ks - 1;
            p.halo_h = p.TH + c.ks - 1;
            p.halo_bo = (halo_mode & 2) ? 1 : 0;
            p.a_bytes = (128 * p.halo_w * p.halo_h + 1023) / 1024 * 1024;
            p.tsh = p.TH; p.tsw = p.TW;
            p.tilesH = (e.Hout + p.TH - 1) / p.TH;
            p.tilesW = (e.Wout + p.TW - 1) / p.TW;
        }
    }
    const int budget0 = 212 * 1024;     // dynamic shared memory (static: bias + barriers ~ 9.5 KB)
    p.Cin = c.Cin; p.dw_w9 = c.dw_w9; p.dw_bias = c.dw_bias;
    p.raw_bytes = c.prod == PROD_DW ? 128 * (p.TW + 2) * (p.TH + 2) : 0;
    const int extra_bytes = c.prod == PROD_DW ? TC_RAW_SLOTS * p.raw_bytes + 10 * c.Cin * 4
                                              : (c.ss ? (128 * TC_SS_PITCH * 4 + 1023) / 1024 * 1024 : 0);     // shift-sum: the fp32 partial-product buffer
    const int ksteps = p.ks * p.ks * p.kchunks;
    const int bres_bytes = ksteps * p.BN * 128;
    const int per = (e.out2 || e.gdn) ? 2 : 1;
    const int nact = (p.BN + 63) / 64;          // 64-column blocks of a tile (each owned by `esplit` warp groups)
    {
        static const int es_mode = getenv("MLIC_ESPLIT") ? atoi(getenv("MLIC_ESPLIT")) : 1;      // development: 0 = one group per block
        p.esplit = (es_mode && c.prod == PROD_TMA && p.store_mode == STORE_TMA) ? (nact == 1 ? 4 : (nact == 2 ? 2 : 1)) : 1;
    }
    // Shared-memory plan.  Small weight matrices (the 192x192 pointwise / GDN GEMMs) may stay resident (no per-tile B
    // reload); every active warp group owns `ring` staging slots of `per` 16 KB buffers; the rest is the operand ring.
    if (p.halo && !(p.BN >= e.N && bres_bytes <= 120 * 1024)) { snprintf(g_tc_err, sizeof g_tc_err, "halo plan: weights not resident (BN=%d)", p.BN); return 8; }
    const bool want_bres = p.halo || (p.BN >= e.N && bres_bytes <= 80 * 1024 && ksteps <= 4 && (!(p.debug & 8) || c.prod != PROD_TMA));
    int stage_bytes = 0;
    bool planned = false;
    for (int bres = want_bres ? 1 : 0; bres >= ((c.prod != PROD_TMA || p.halo) ? 1 : 0) && !planned; --bres) {
        stage_bytes = p.a_bytes + (bres ? 0 : p.BN * 128);
        const int budget = budget0 - 1024 - (bres ? bres_bytes : 0) - extra_bytes;
        for (int ring = (p.store_mode == STORE_TMA ? 2 : 0); ring >= 0 && !planned; --ring) {
            if (p.store_mode == STORE_TMA && ring == 0) break;
            const int stg_bytes = nact * ring * per * TC_STG_BYTES;
            const int st = (budget - stg_bytes) / stage_bytes;
            const int need = (ring == 2) ? (p.BN > 192 ? 4 : 3) : 2;
            if (st >= need) {
                p.b_resident = bres; p.nstg = nact * ring * per;
                p.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
                planned = true;
            }
        }
    }
    if (!planned) { snprintf(g_tc_err, sizeof g_tc_err, "no shared-memory plan for BN=%d ksteps=%d", p.BN, ksteps); return 8; }
    // L2 prefetch of the A patch a few tiles ahead (MLIC_L2PF_MB = input size in MB from which it is used; default: never).  It paid
    // while the pointwise layers were DRAM-latency-bound; they are epilogue-bound now, and on the slice loop's GEMMs the extra TMA
    // instructions only load the single issuing thread.  Same box, 32 / 8 images per step: always 526 / 501, >= 384 MB 531 / 501,
    // never 533 / 504 MP/s.
    static const double pf_mb = getenv("MLIC_L2PF_MB") ? atof(getenv("MLIC_L2PF_MB")) : 1e12;
    const bool streams = (double)c.B * c.H * c.W * c.Cin * 2.0 >= pf_mb * 1e6;
    p.l2_prefetch = (p.ks == 1 && c.prod != PROD_SQ && !c.ck && !(p.debug & 16) && streams) ? (c.prod == PROD_DW ? 2 : 3) : 0;
    if (c.prod != PROD_TMA && p.store_mode == STORE_DIRECT) p.epi_vec = p.epi_vec && p.ld_vec;
    // (development, MLIC_HALO bit3; measured slower than the single accumulator: 420 us vs 195 us on the final 192 -> 12 conv)
    p.acc_split = ((halo_mode & 8) && p.halo && p.store_mode == STORE_NCHW3 && p.BN == 16 && p.ks * p.ks * p.BN <= 256) ? 1 : 0;
    p.acc_stride = 32;
    while (p.acc_stride < p.BN * (p.acc_split ? p.ks * p.ks : 1)) p.acc_stride <<= 1;
    p.tilesN = (e.N + p.BN - 1) / p.BN;
    p.ntiles = c.B * p.tilesH * p.tilesW * p.tilesN;
    {
        static const int bs_mode = getenv("MLIC_BSPLIT") ? atoi(getenv("MLIC_BSPLIT")) : 1;      // development: 0 = one TMA issuer
        p.bsplit = (bs_mode && c.prod == PROD_TMA && !p.b_resident && !p.halo && !c.ss && !c.ck && ksteps >= 4) ? 1 : 0;
    }
    p.ld_vec = (ok16(e.res, e.res_ld, 2) && ok16(e.gdn_x, e.gdn_ld, 2) && (e.N % 8 == 0) && (!e.shuffle || (Cq % 32) == 0)) ? 1 : 0;
    {
        bool v = vec != 0 && (e.N % 8 == 0) && p.ld_vec;
        v = v && ok16(e.out, e.out_ld, e.out_f32 ? 4 : 2) && ok16(e.out2, e.out2_ld, 2);
        if (e.shuffle) v = v && ((Cq % 32) == 0);
        p.epi_vec = v ? 1 : 0;
    }

    TcMaps tm;
    memset(&tm, 0, sizeof tm);
    p.ck = c.ck;
    if (c.ck) {
        // anchor (ck 1): row h keeps w = 2j + 1 - (h & 1); non-anchor (ck 2): w = 2j + (h & 1)   (anchor = (h + w) odd)
        const size_t ld = (size_t)c.ld;
        const bf16* basep = reinterpret_cast<const bf16*>(c.in) + (c.ck == 1 ? ld : 0);
        const size_t hp_stride = c.ck == 1 ? ((size_t)c.W - 1) * ld : ((size_t)c.W + 1) * ld;
        cuuint64_t dims[5] = {(cuuint64_t)c.Cin, (cuuint64_t)(c.W / 2), 2, (cuuint64_t)(c.H / 2), (cuuint64_t)c.B};
        cuuint64_t strides[4] = {(cuuint64_t)(2 * ld) * 2, (cuuint64_t)hp_stride * 2, (cuuint64_t)(2 * (size_t)c.W * ld) * 2, (cuuint64_t)c.sB * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)p.TW, 2, (cuuint32_t)(p.TH / 2), 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = g_encode(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(basep), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(A, checkerboard) failed: %d (C=%d W=%d H=%d B=%d)", (int)r, c.Cin, c.W, c.H, c.B);
            return 2;
        }
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)c.Cin, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.sW * 2, (cuuint64_t)c.sH * 2, (cuuint64_t)c.sB * 2};
        const bool halo = c.prod == PROD_DW;
        cuuint32_t box[4] = {64, (cuuint32_t)(p.TW + (halo ? 2 : 0)), (cuuint32_t)(p.TH + (halo ? 2 : 0)), 1};
        if (p.halo) { box[1] = (cuuint32_t)p.halo_w; box[2] = (cuuint32_t)p.halo_h; }
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encode(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(c.in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, halo ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d B=%d sW=%d sH=%d)", (int)r,
                     c.Cin, c.W, c.H, c.B, c.sW, c.sH);
            return 2;
        }
    }
    {
        const cuuint64_t Ktot = (cuuint64_t)c.ks * c.ks * c.Cpad;
        cuuint64_t dims[2] = {Ktot, (cuuint64_t)e.N};
        cuuint64_t strides[1] = {Ktot * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&tm.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(c.w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(B) failed: %d (K=%llu N=%d BN=%d)", (int)r,
                     (unsigned long long)Ktot, e.N, p.BN);
            return 3;
        }
    }
    if (p.store_mode == STORE_TMA) {
        bf16* ob = reinterpret_cast<bf16*>(e.out);
        const size_t ld = (size_t)e.out_ld;
        if (!e.shuffle) {
            if (encode_out_map(&tm.o[0], ob, e.N, e.Wout, e.Hout, c.B, ld, (size_t)e.Wout * ld, (size_t)e.Hout * e.Wout * ld, p.TW, p.TH)) return 7;
        } else {
            const size_t OW = 2 * (size_t)e.Wout, OH = 2 * (size_t)e.Hout;
            for (int g = 0; g < 4; ++g) {       // group g = 2r + s -> output pixel (2h + r, 2w + s)
                bf16* bp = ob + ((size_t)(g >> 1) * OW + (size_t)(g & 1)) * ld;
                if (encode_out_map(&tm.o[g], bp, Cq, e.Wout, e.Hout, c.B, 2 * ld, 2 * OW * ld, OH * OW * ld, p.TW, p.TH)) return 7;
            }
        }
        if (e.res && encode_out_map(&tm.r, const_cast<void*>(e.res), e.N, e.Wout, e.Hout, c.B, (size_t)e.res_ld, (size_t)e.Wout * e.res_ld,
                                    (size_t)e.Hout * e.Wout * e.res_ld, p.TW, p.TH)) return 7;
        if (e.gdn && encode_out_map(&tm.gx, const_cast<void*>(e.gdn_x), e.N, e.Wout, e.Hout, c.B, (size_t)e.gdn_ld, (size_t)e.Wout * e.gdn_ld,
                                    (size_t)e.Hout * e.Wout * e.gdn_ld, p.TW, p.TH)) return 7;
        if (e.out2) {
            const size_t ld2 = (size_t)e.out2_ld;
            if (encode_out_map(&tm.o2, e.out2, e.N, e.Wout, e.Hout, c.B, ld2, (size_t)e.Wout * ld2, (size_t)e.Hout * e.Wout * ld2, p.TW, p.TH)) return 7;
        }
    }
    const size_t smem = (size_t)(p.b_resident ? bres_bytes : 0) + (size_t)p.stages * stage_bytes + (size_t)p.nstg * TC_STG_BYTES +
                        (size_t)extra_bytes + 1024;
    if (smem > (size_t)budget0 + 1024) { snprintf(g_tc_err, sizeof g_tc_err, "shared-memory plan too large: %zu", smem); return 8; }
    // one instantiation per (activation, GDN mode, residual[, fused producer]): the epilogue only carries the code its
    // layer needs
    typedef void (*KernelFn)(const TcMaps, TcParams, Epi, unsigned long long*);
    static const KernelFn table[3][3][2] = {
        {{conv_gemm_tc_kernel<0, 0, false, 0>, conv_gemm_tc_kernel<0, 0, true, 0>}, {conv_gemm_tc_kernel<0, 1, false, 0>, conv_gemm_tc_kernel<0, 1, true, 0>},
         {conv_gemm_tc_kernel<0, 2, false, 0>, conv_gemm_tc_kernel<0, 2, true, 0>}},
        {{conv_gemm_tc_kernel<1, 0, false, 0>, conv_gemm_tc_kernel<1, 0, true, 0>}, {conv_gemm_tc_kernel<1, 1, false, 0>, conv_gemm_tc_kernel<1, 1, true, 0>},
         {conv_gemm_tc_kernel<1, 2, false, 0>, conv_gemm_tc_kernel<1, 2, true, 0>}},
        {{conv_gemm_tc_kernel<2, 0, false, 0>, conv_gemm_tc_kernel<2, 0, true, 0>}, {conv_gemm_tc_kernel<2, 1, false, 0>, conv_gemm_tc_kernel<2, 1, true, 0>},
         {conv_gemm_tc_kernel<2, 2, false, 0>, conv_gemm_tc_kernel<2, 2, true, 0>}}};
    static const KernelFn table_dw[3][2] = {{conv_gemm_tc_kernel<0, 0, false, PROD_DW>, conv_gemm_tc_kernel<0, 0, true, PROD_DW>},
                                            {conv_gemm_tc_kernel<1, 0, false, PROD_DW>, conv_gemm_tc_kernel<1, 0, true, PROD_DW>},
                                            {nullptr, conv_gemm_tc_kernel<2, 0, true, PROD_DW>}};
    static const KernelFn table_sq[2][2] = {{conv_gemm_tc_kernel<0, 1, false, PROD_SQ>, conv_gemm_tc_kernel<0, 1, true, PROD_SQ>},
                                            {conv_gemm_tc_kernel<0, 2, false, PROD_SQ>, conv_gemm_tc_kernel<0, 2, true, PROD_SQ>}};
    static bool attr_set[TC_MAX_DEV][3][3][2][3] = {};
    const int dev = cur_dev();
    if (e.act < 0 || e.act > 2 || e.gdn < 0 || e.gdn > 2) { snprintf(g_tc_err, sizeof g_tc_err, "bad epilogue mode"); return 9; }
    const int ri = e.res ? 1 : 0;
    KernelFn fn = table[e.act][e.gdn][ri];
    if (c.prod == PROD_DW) fn = table_dw[e.act][ri];
    else if (c.prod == PROD_SQ) fn = table_sq[e.gdn - 1][ri];
    if (!fn) { snprintf(g_tc_err, sizeof g_tc_err, "no kernel instantiation for act=%d res=%d prod=%d", e.act, ri, c.prod); return 9; }
    if (!attr_set[dev][e.act][e.gdn][ri][c.prod]) {
        cudaError_t er = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, budget0 + 1024);
        if (er != cudaSuccess) {
            snprintf(g_tc_err, sizeof g_tc_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(er));
            return 4;
        }
        attr_set[dev][e.act][e.gdn][ri][c.prod] = true;
    }
    const int num_sms = dev_sms(dev);
    dim3 grid((unsigned)(p.ntiles < num_sms ? p.ntiles : num_sms));
    unsigned long long* dbg = nullptr;
    if (p.debug & 32) {             // development: per-role wait clocks of the first launches, printed to stderr
        static unsigned long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc((void**)&dbuf, 148 * 11 * sizeof(unsigned long long));
        cudaMemsetAsync(dbuf, 0, 148 * 11 * sizeof(unsigned long long), s);
        dbg = dbuf;
    }
    // MLIC_PDL: 1 (default) programmatic stream serialization -- the kernel's prologue overlaps the tail of its predecessor:
    // one 1920x1088 forward 8.61 -> 7.67 ms, 32 images 130.2 -> 129.8 ms (profiles/r01_pdl_probe.txt); 0 plain launch;
    // 2 as 1 with the dynamic shared memory padded above half an SM, so that early CTAs of a small-footprint launch cannot
    // double up on the SMs that free first (measured: no better than 1)
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    const unsigned nthreads = c.prod == PROD_TMA ? TC_THREADS : TC_FUSED_THREADS;
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(nthreads);
        cfg.dynamicSmemBytes = pdl == 2 ? std::max<size_t>(smem, (size_t)117 * 1024) : smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, fn, tm, p, e, dbg);
    } else
    fn<<<grid, nthreads, smem, s>>>(tm, p, e, dbg);
    if (dbg) {
        static int printed = 0;
        unsigned long long h[148 * 11];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        if (printed++ < 3 || (p.debug & 64)) {
            double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (unsigned i = 0; i < grid.x; ++i) { for (int j = 0; j < 8; ++j) a[j] += (double)h[i * 8 + j] / grid.x; a[8] += (double)h[grid.x * 8 + i] / grid.x; }
            if (c.prod) {
                double b9 = 0, b10 = 0;
                for (unsigned i = 0; i < grid.x; ++i) { b9 += (double)h[grid.x * 9 + i] / grid.x; b10 += (double)h[grid.x * 10 + i] / grid.x; }
                fprintf(stderr, "[tc dbg] prod=%d compute warp 4: wait-raw %.0f wait-stage-free %.0f | first epilogue warp: drain+operand wait %.0f math %.0f fence+barrier+store %.0f\n", c.prod, a[6], a[7], b9, b10, a[8]);
            }
            else fprintf(stderr, "[tc dbg] epilogue warp 4: drain-wait %.0f math %.0f fence+barrier+store %.0f\n", a[6], a[7], a[8]);
            fprintf(stderr, "[tc dbg] BN=%d ksteps=%d stages=%d nstg=%d bres=%d tiles/cta=%.1f | clocks: total %.0f prod-wait-empty %.0f mma-wait-full %.0f mma-wait-accempty %.0f epi-wait-accfull %.0f epi-total %.0f\n",
                    p.BN, ksteps, p.stages, p.nstg, p.b_resident, (double)p.ntiles / grid.x, a[0], a[1], a[2], a[3], a[4], a[5]);
        }
    }
    cudaError_t er = cudaGetLastError();
    if (er != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "conv_gemm_tc launch: %s (smem %zu)", cudaGetErrorString(er), smem);
        return 5;
    }
    return 0;
}

}  // namespace mlic
