// Dense k x k convolution (k = 3 | 5, stride 1, pad k / 2) with a NARROW output (N <= 128), bf16 NHWC -- the 5x5 re-projections of the
// global contexts (modules/transform/context.py:160,216: Conv2d(32 i -> 96, 5) and Conv2d(32 -> 64, 5) per slice).
//
// The plain implicit-GEMM path loads one 16 KB activation box per (tap, 64-channel chunk) -- 25 shifted copies of the same pixels,
// 28 KB of operands per K step of ~530 MMA clocks -- and is L2-feed bound: measured 40 B/clock per SM (the chip-wide L2 cap is ~43),
// the MMA thread waiting for operands a third of the time (tools/tc_bench.py, MLIC_TC_DEBUG=32).  Here the activations are staged
// once per (chunk, kx) as a COLUMN-SHIFTED patch {64 ch, 16, 8 + k - 1}: the k row taps of that kx are UMMA descriptors into it whose
// start moves by whole patch rows (ky * 16 pixels = ky * 2 KB: 1024-byte aligned, so the operand feed runs at the rate of a plain
// tile).  A full halo patch with the kx shift in the descriptor as well (start moved by single pixels) was built first and is
// slower than the plain path: 169 clocks per M = 128, K = 16 step against 133, the misaligned 8-row groups straddle two swizzle atoms.
// Activation traffic per tile and chunk: k boxes of 24 KB instead of k*k of 16 KB; the weights stream, one {64, N} box per tap.
//
// 384 threads: warp 0 TMA (patches and weight boxes, in the order the MMA consumes them), warp 1 MMA issue, warp 2 TMEM (two
// accumulator stages), warps 4..11 epilogue (two groups of four: 64 output columns each; bias, bf16, swizzled staging block, TMA store).
#include "kernels.h"
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlic {

constexpr int CH_TH = 8, CH_TW = 16;
constexpr int CH_NPATCH = 4;
constexpr int CH_NSTAGE = 8;                 // at most; fewer when the weight boxes are large (ChParams::nstage)
constexpr int CH_STG_BYTES = 128 * 128;
constexpr int CH_THREADS = 128 + 8 * 32;

struct ChParams {
    int tilesH, tilesW, ntiles;
    int ks, kchunks, BN, N;             // BN = N rounded up to 16 (the MMA's N)
    int patch_h, patch_bytes;           // staged patch: 16 x patch_h pixels, bytes (a multiple of 2 KB)
    int b_bytes;                        // BN * 128
    int nstage;                         // weight boxes in flight
    int Cin;                            // real input channels: a chunk issues ceil(min(64, Cin - 64 cc) / 16) MMAs per tap, not 4
    const float* bias;
};
struct ChMaps { CUtensorMap a, b, o; };

__global__ void __launch_bounds__(CH_THREADS, 1)
conv_halo_kernel(const __grid_constant__ ChMaps tm, const ChParams p, unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t ch_smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)ch_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t p_full[CH_NPATCH], p_empty[CH_NPATCH];
    __shared__ uint64_t full[CH_NSTAGE], empty[CH_NSTAGE];
    __shared__ uint64_t d_full[2], d_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float sBias[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* patches = base;
    uint8_t* ring = base + (size_t)CH_NPATCH * p.patch_bytes;
    uint8_t* stg = ring + (size_t)p.nstage * p.b_bytes;
    const int ngroups = (p.N + 63) / 64;                    // epilogue groups that own columns

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.o) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < CH_NPATCH; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], (uint32_t)(4 * ngroups)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += CH_THREADS) sBias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tiles_per_img = p.tilesH * p.tilesW;
    const int pad = p.ks / 2;
    long long tw0 = 0, tw1 = 0;
    const long long t_start = dbg ? clock64() : 0;
#define CH_TIMED(acc, stmt) do { if (dbg) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)
#define CH_TILE(t)                                                        \
    const int img = (t) / tiles_per_img;                                  \
    const int trem = (t) - img * tiles_per_img;                           \
    const int th = trem / p.tilesW, tw = trem - th * p.tilesW;            \
    const int h0 = th * CH_TH, w0 = tw * CH_TW

    if (warp == 0) {
        if (lane == 0) {
            int pb = 0, s = 0;
            uint32_t pph = 0, ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                CH_TILE(t);
                for (int cc = 0; cc < p.kchunks; ++cc)
                    for (int kx = 0; kx < p.ks; ++kx) {
                        CH_TIMED(tw0, mbar_wait<false>(&p_empty[pb], pph ^ 1));
                        mbar_expect_tx(&p_full[pb], (uint32_t)p.patch_bytes);
                        tma_load_4d(patches + (size_t)pb * p.patch_bytes, &tm.a, &p_full[pb], cc * 64, w0 + kx - pad, h0 - pad, img);
                        if (++pb == CH_NPATCH) { pb = 0; pph ^= 1; }
                        for (int ky = 0; ky < p.ks; ++ky) {
                            CH_TIMED(tw1, mbar_wait<false>(&empty[s], ph ^ 1));
                            mbar_expect_tx(&full[s], (uint32_t)p.b_bytes);
                            tma_load_2d(ring + (size_t)s * p.b_bytes, &tm.b, &full[s], ((ky * p.ks + kx) * p.kchunks + cc) * 64, 0);
                            if (++s == p.nstage) { s = 0; ph ^= 1; }
                        }
                    }
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = (unsigned long long)(clock64() - t_start); dbg[1] = (unsigned long long)tw0; dbg[2] = (unsigned long long)tw1; }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int pb = 0, s = 0, it = 0;
            uint32_t pph = 0, ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                CH_TIMED(tw0, mbar_wait<false>(&d_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u));
                tcgen05_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(acc * 128);
                for (int cc = 0; cc < p.kchunks; ++cc) {
                    const int nk = min(4, (p.Cin - cc * 64 + 15) >> 4);      // the zero-padded tail of the last chunk is not multiplied
                    for (int kx = 0; kx < p.ks; ++kx) {
                        CH_TIMED(tw1, mbar_wait<false>(&p_full[pb], pph));
                        tcgen05_fence_after();
                        const uint32_t pa = smem_u32(patches + (size_t)pb * p.patch_bytes);
                        for (int ky = 0; ky < p.ks; ++ky) {
                            CH_TIMED(tw1, mbar_wait<false>(&full[s], ph));
                            tcgen05_fence_after();
                            // rows r = ty * 16 + tx of the tile are patch pixels (ty + ky) * 16 + tx: a plain 128-row tile starting ky patch rows in
                            const uint64_t adesc = umma_desc_sw128(pa + (uint32_t)(ky * CH_TW * 128));
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + (size_t)s * p.b_bytes));
                            for (int kk = 0; kk < nk; ++kk) umma_bf16(dcol, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (cc | kx | ky | kk) ? 1u : 0u);
                            tcgen05_commit(&empty[s]);
                            if (++s == p.nstage) { s = 0; ph ^= 1; }
                        }
                        tcgen05_commit(&p_empty[pb]);
                        if (++pb == CH_NPATCH) { pb = 0; pph ^= 1; }
                    }
                }
                tcgen05_commit(&d_full[acc]);
            }
            if (dbg && blockIdx.x == 0) { dbg[3] = (unsigned long long)(clock64() - t_start); dbg[4] = (unsigned long long)tw0; dbg[5] = (unsigned long long)tw1; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, eb = (warp - 4) >> 2;
        if (eb < ngroups) {
            const int r = q * 32 + lane;
            const bool gissuer = (q == 0 && lane == 0);
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            uint8_t* sb = stg + (size_t)eb * CH_STG_BYTES;
            const uint32_t sb_s = smem_u32(sb);
            const int ncol = min(64, p.BN - eb * 64);              // accumulator columns of this group (multiple of 16)
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                CH_TILE(t);
                const int acc = it & 1;
                if (gissuer) tma_store_wait_read(0);
                asm volatile("bar.sync %0, 128;" ::"r"(eb + 1) : "memory");
                CH_TIMED(tw0, mbar_wait(&d_full[acc], (uint32_t)(it >> 1) & 1u));
                tcgen05_fence_after();
                const uint32_t trow = lane_base + (uint32_t)(acc * 128 + eb * 64);
#pragma unroll 1
                for (int pr = 0; pr < 4; ++pr) {
                    uint32_t raw[16];
                    if (pr * 16 < ncol) {
                        tmem_ld16(trow + (uint32_t)(pr * 16), raw);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) raw[j] = 0u;
                    }
                    if (pr == 3) {                  // accumulator fully read by this warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_empty[acc]);
                    }
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int jj = pr * 2 + sub;
                        const uint32_t off = (uint32_t)(r * 128 + ((jj ^ (r & 7)) << 4));
                        const float4 ba = *reinterpret_cast<const float4*>(sBias + eb * 64 + jj * 8), bb = *reinterpret_cast<const float4*>(sBias + eb * 64 + jj * 8 + 4);
                        float2 v[4];
                        v[0] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 0]), __uint_as_float(raw[sub * 8 + 1])), make_float2(ba.x, ba.y));
                        v[1] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 2]), __uint_as_float(raw[sub * 8 + 3])), make_float2(ba.z, ba.w));
                        v[2] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 4]), __uint_as_float(raw[sub * 8 + 5])), make_float2(bb.x, bb.y));
                        v[3] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 6]), __uint_as_float(raw[sub * 8 + 7])), make_float2(bb.z, bb.w));
                        uint4 o;
                        { __nv_bfloat162 h0b = __floats2bfloat162_rn(v[0].x, v[0].y), h1b = __floats2bfloat162_rn(v[1].x, v[1].y),
                                         h2b = __floats2bfloat162_rn(v[2].x, v[2].y), h3b = __floats2bfloat162_rn(v[3].x, v[3].y);
                          o.x = *reinterpret_cast<uint32_t*>(&h0b); o.y = *reinterpret_cast<uint32_t*>(&h1b);
                          o.z = *reinterpret_cast<uint32_t*>(&h2b); o.w = *reinterpret_cast<uint32_t*>(&h3b); }
                        sts128(sb_s + off, o);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(eb + 1) : "memory");
                if (gissuer) {
                    tma_store_4d(&tm.o, sb, eb * 64, w0, h0, img);       // columns >= N are clipped by the tensor map
                    tma_store_commit();
                }
            }
            if (gissuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            if (dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { dbg[6] = (unsigned long long)(clock64() - t_start); dbg[7] = (unsigned long long)tw0; }
        }
    }
#undef CH_TILE
#undef CH_TIMED
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}


// ------------------------------------------------------------------------------------------ roles swapped: weights as the M operand
// The kernel above is bound by the A-operand feed: an M = 128, K = 16 MMA costs >= 124 clocks whatever N is, so a conv with
// N = 96 runs at the MMA rate of one with N = 256.  Here the GEMM is transposed: D[n][pixel] = sum_k W[n][k] X[pixel][k] -- the weight box
// {64, N <= 128 rows} is the A operand (M = 128: rows beyond N read whatever follows the stage and land in accumulator lanes nobody reads),
// 256 pixels (TH x TW = 16 x 16 or 8 x 32) are the B operand, again as row-shifted descriptors into a column-shifted patch
// {64 ch, TW, TH + k - 1}: twice the pixels per MMA clock.  The accumulator is [channel lane][pixel column]: an epilogue warp owns 32
// channels x 128 pixels and stores bf16 NHWC directly (32 lanes = 64 contiguous bytes of one pixel; the output is 2 % of the operand traffic).
// 384 threads: warp 0 TMA, warp 1 MMA, warp 2 TMEM (2 x 256 columns), warps 4..11 epilogue.
constexpr int CT_NPATCH = 3;
constexpr int CT_NSTAGE = 8;
constexpr int CT_THREADS = 128 + 8 * 32;

struct CtParams {
    int B, H, W, N, Cin;
    int tilesH, tilesW, ntiles;
    int TH, TW;                         // TH * TW = 256, TW = 16 | 32
    int ks, kchunks, BN;                // BN = N rounded up to 16: rows of a weight box
    int patch_bytes, b_bytes, nstage;
    const float* bias;
    bf16* out; long long out_ld;
};
struct CtMaps { CUtensorMap a, b; };

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_halo_t_kernel(const __grid_constant__ CtMaps tm, const CtParams p, unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t ch_smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)ch_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t p_full[CT_NPATCH], p_empty[CT_NPATCH];
    __shared__ uint64_t full[CT_NSTAGE], empty[CT_NSTAGE];
    __shared__ uint64_t d_full[2], d_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* patches = base;
    uint8_t* ring = base + (size_t)CT_NPATCH * p.patch_bytes;
    const int nq = (p.N + 31) >> 5;                         // lane quarters that hold output channels

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < CT_NPATCH; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], (uint32_t)(2 * nq)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tiles_per_img = p.tilesH * p.tilesW;
    const int pad = p.ks / 2;
    long long tw0 = 0, tw1 = 0;
    const long long t_start = dbg ? clock64() : 0;
#define CT_TIMED(acc, stmt) do { if (dbg) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)
#define CT_TILE(t)                                                        \
    const int img = (t) / tiles_per_img;                                  \
    const int trem = (t) - img * tiles_per_img;                           \
    const int th = trem / p.tilesW, tw = trem - th * p.tilesW;            \
    const int h0 = th * p.TH, w0 = tw * p.TW

    if (warp == 0) {
        if (lane == 0) {
            int pb = 0, s = 0;
            uint32_t pph = 0, ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                CT_TILE(t);
                for (int cc = 0; cc < p.kchunks; ++cc)
                    for (int kx = 0; kx < p.ks; ++kx) {
                        CT_TIMED(tw0, mbar_wait<false>(&p_empty[pb], pph ^ 1));
                        mbar_expect_tx(&p_full[pb], (uint32_t)p.patch_bytes);
                        tma_load_4d(patches + (size_t)pb * p.patch_bytes, &tm.a, &p_full[pb], cc * 64, w0 + kx - pad, h0 - pad, img);
                        if (++pb == CT_NPATCH) { pb = 0; pph ^= 1; }
                        for (int ky = 0; ky < p.ks; ++ky) {
                            CT_TIMED(tw1, mbar_wait<false>(&empty[s], ph ^ 1));
                            mbar_expect_tx(&full[s], (uint32_t)p.b_bytes);
                            tma_load_2d(ring + (size_t)s * p.b_bytes, &tm.b, &full[s], ((ky * p.ks + kx) * p.kchunks + cc) * 64, 0);
                            if (++s == p.nstage) { s = 0; ph ^= 1; }
                        }
                    }
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = (unsigned long long)(clock64() - t_start); dbg[1] = (unsigned long long)tw0; dbg[2] = (unsigned long long)tw1; }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            int pb = 0, s = 0, it = 0;
            uint32_t pph = 0, ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                CT_TIMED(tw0, mbar_wait<false>(&d_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u));
                tcgen05_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(acc * 256);
                for (int cc = 0; cc < p.kchunks; ++cc) {
                    const int nk = min(4, (p.Cin - cc * 64 + 15) >> 4);
                    for (int kx = 0; kx < p.ks; ++kx) {
                        CT_TIMED(tw1, mbar_wait<false>(&p_full[pb], pph));
                        tcgen05_fence_after();
                        const uint32_t pa = smem_u32(patches + (size_t)pb * p.patch_bytes);
                        for (int ky = 0; ky < p.ks; ++ky) {
                            CT_TIMED(tw1, mbar_wait<false>(&full[s], ph));
                            tcgen05_fence_after();
                            const uint64_t adesc = umma_desc_sw128(smem_u32(ring + (size_t)s * p.b_bytes));      // weights: 128 rows from the stage base
                            // pixel r = ty * TW + tx of the tile is patch pixel (ty + ky) * TW + tx: a plain 256-row operand starting ky patch rows in
                            const uint64_t bdesc = umma_desc_sw128(pa + (uint32_t)(ky * p.TW * 128));
                            for (int kk = 0; kk < nk; ++kk) umma_bf16(dcol, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (cc | kx | ky | kk) ? 1u : 0u);
                            tcgen05_commit(&empty[s]);
                            if (++s == p.nstage) { s = 0; ph ^= 1; }
                        }
                        tcgen05_commit(&p_empty[pb]);
                        if (++pb == CT_NPATCH) { pb = 0; pph ^= 1; }
                    }
                }
                tcgen05_commit(&d_full[acc]);
            }
            if (dbg && blockIdx.x == 0) { dbg[3] = (unsigned long long)(clock64() - t_start); dbg[4] = (unsigned long long)tw0; dbg[5] = (unsigned long long)tw1; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        if (q < nq) {
            const int n = q * 32 + lane;
            const bool nok = n < p.N;
            const float bias = (p.bias && nok) ? p.bias[n] : 0.f;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            const int gshift = p.TW == 16 ? 4 : 5;                  // 16 consecutive pixel columns never straddle a tile row
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                CT_TILE(t);
                const int acc = it & 1;
                CT_TIMED(tw0, mbar_wait(&d_full[acc], (uint32_t)(it >> 1) & 1u));
                tcgen05_fence_after();
                const uint32_t trow = lane_base + (uint32_t)(acc * 256 + half * 128);
#pragma unroll 1
                for (int g = 0; g < 8; ++g) {
                    uint32_t raw[16];
                    tmem_ld16(trow + (uint32_t)(g * 16), raw);
                    tmem_ld_wait();
                    if (g == 7) {                   // accumulator fully read by this warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_empty[acc]);
                    }
                    const int r0 = half * 128 + g * 16;
                    const int h = h0 + (r0 >> gshift), wq = w0 + (r0 & (p.TW - 1));
                    if (h < p.H && nok) {
                        bf16* o = p.out + ((size_t)((size_t)img * p.H + h) * p.W + wq) * (size_t)p.out_ld + n;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (wq + j < p.W) o[(size_t)j * p.out_ld] = __float2bfloat16_rn(__uint_as_float(raw[j]) + bias);
                    }
                }
            }
            if (dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { dbg[6] = (unsigned long long)(clock64() - t_start); dbg[7] = (unsigned long long)tw0; }
        }
    }
#undef CT_TILE
#undef CT_TIMED
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled_ch)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local char g_ch_err[256] = "";
const char* conv_halo_last_error() { return g_ch_err; }

bool conv_halo_supported(const ConvHaloArgs& a) {
    if (!a.in || !a.w || !a.out) return false;
    if (a.B <= 0 || a.H <= 0 || a.W <= 0) return false;
    if (a.ks != 3 && a.ks != 5) return false;
    if (a.Cin < 8 || (a.Cin % 8) != 0 || a.Cpad % 64 != 0 || a.Cpad < a.Cin || a.Cpad > 1024) return false;
    if (a.N < 16 || a.N > 128 || (a.N % 8) != 0) return false;
    if (((uintptr_t)a.in % 16) != 0 || (a.ld % 8) != 0 || ((uintptr_t)a.out % 16) != 0 || (a.out_ld % 8) != 0 || ((uintptr_t)a.w % 16) != 0) return false;
    return true;
}

static int launch_conv_halo_t(const ConvHaloArgs& a, cudaStream_t s) {
    PFN_encodeTiled_ch enc = (PFN_encodeTiled_ch)tc_encode_fn();
    CtMaps tm;
    memset(&tm, 0, sizeof tm);
    CtParams p;
    memset(&p, 0, sizeof p);
    // 256-pixel tile: 16 x 16 or 8 x 32, whichever pads the image less (ties: 16 x 16, the smaller halo)
    const long long pad16 = (long long)((a.H + 15) / 16 * 16) * ((a.W + 15) / 16 * 16), pad32 = (long long)((a.H + 7) / 8 * 8) * ((a.W + 31) / 32 * 32);
    static const int force_tw = getenv("MLIC_HALO_TW") ? atoi(getenv("MLIC_HALO_TW")) : 0;
    p.TW = force_tw ? force_tw : (pad32 < pad16 ? 32 : 16);
    p.TH = 256 / p.TW;
    p.B = a.B; p.H = a.H; p.W = a.W; p.N = a.N; p.Cin = a.Cin;
    p.tilesH = (a.H + p.TH - 1) / p.TH; p.tilesW = (a.W + p.TW - 1) / p.TW;
    const long long nt = (long long)a.B * p.tilesH * p.tilesW;
    if (nt <= 0 || nt > 0x3fffffffLL) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: tile count out of range"); return 3; }
    p.ntiles = (int)nt;
    p.ks = a.ks; p.kchunks = a.Cpad / 64; p.BN = (a.N + 15) / 16 * 16;
    const int patch_h = p.TH + a.ks - 1;
    p.patch_bytes = 128 * p.TW * patch_h;
    p.b_bytes = p.BN * 128;
    p.bias = a.bias; p.out = (bf16*)a.out; p.out_ld = a.out_ld;
    const int budget = 232448 - 1024;
    const int tail = (128 - p.BN) * 128;                        // the M = 128 descriptor of the last stage reads this far past it
    p.nstage = (budget - 1024 - CT_NPATCH * p.patch_bytes - tail) / p.b_bytes;
    if (p.nstage > CT_NSTAGE) p.nstage = CT_NSTAGE;
    if (p.nstage < 3) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: shared-memory plan too large"); return 8; }
    const int smem = CT_NPATCH * p.patch_bytes + p.nstage * p.b_bytes + tail + 1024;
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.ld * 2, (cuuint64_t)a.W * a.ld * 2, (cuuint64_t)a.H * a.W * a.ld * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)patch_h, 1};
        CUresult r = enc(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: encode(A) failed: %d", (int)r); return 2; }
    }
    {
        const cuuint64_t Ktot = (cuuint64_t)a.ks * a.ks * a.Cpad;
        cuuint64_t dims[2] = {Ktot, (cuuint64_t)a.N};
        cuuint64_t strides[1] = {Ktot * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: encode(B) failed: %d", (int)r); return 2; }
    }
    static bool attr[64] = {};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    if (!attr[dev]) {
        cudaError_t er = cudaFuncSetAttribute(conv_halo_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
        if (er != cudaSuccess) { snprintf(g_ch_err, sizeof g_ch_err, "cudaFuncSetAttribute(conv_halo_t): %s", cudaGetErrorString(er)); return 4; }
        attr[dev] = true;
    }
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.ntiles < sms ? p.ntiles : sms));
    cfg.blockDim = dim3((unsigned)CT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    static const int dbgmode = getenv("MLIC_TC_DEBUG") ? atoi(getenv("MLIC_TC_DEBUG")) : 0;
    unsigned long long* dbg = nullptr;
    if (dbgmode & 32) {
        static unsigned long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc((void**)&dbuf, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dbuf, 0, 16 * sizeof(unsigned long long), s);
        dbg = dbuf;
    }
    cudaError_t er = cudaLaunchKernelEx(&cfg, conv_halo_t_kernel, tm, p, dbg);
    if (dbg && er == cudaSuccess) {
        static int printed = 0;
        unsigned long long h[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        if (printed++ < 6) {
            const double tiles = (double)((p.ntiles + cfg.gridDim.x - 1) / cfg.gridDim.x);
            fprintf(stderr, "[conv halo-t dbg] tiles/cta %.0f tile %dx%d ks %d Cin %d BN %d stages %d | per tile: tma total %.0f wait-patch-free %.0f wait-stage-free %.0f | mma total %.0f wait-acc-free %.0f wait-operands %.0f | epilogue w4 total %.0f wait-acc %.0f\n",
                    tiles, p.TH, p.TW, p.ks, p.Cin, p.BN, p.nstage, h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles, h[6] / tiles, h[7] / tiles);
        }
    }
    if (er != cudaSuccess) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo_t launch: %s (smem %d)", cudaGetErrorString(er), smem); return 5; }
    return 0;
}

int launch_conv_halo(const ConvHaloArgs& a, cudaStream_t s) {
    if (tc_init()) { snprintf(g_ch_err, sizeof g_ch_err, "%s", tc_last_error()); return 1; }
    if (!conv_halo_supported(a)) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: unsupported layer"); return 1; }
    if (a.swap) return launch_conv_halo_t(a, s);
    PFN_encodeTiled_ch enc = (PFN_encodeTiled_ch)tc_encode_fn();
    ChMaps tm;
    memset(&tm, 0, sizeof tm);
    ChParams p;
    memset(&p, 0, sizeof p);
    p.tilesH = (a.H + CH_TH - 1) / CH_TH; p.tilesW = (a.W + CH_TW - 1) / CH_TW;
    const long long nt = (long long)a.B * p.tilesH * p.tilesW;
    if (nt <= 0 || nt > 0x3fffffffLL) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: tile count out of range"); return 3; }
    p.ntiles = (int)nt;
    p.ks = a.ks; p.kchunks = a.Cpad / 64; p.N = a.N; p.BN = (a.N + 15) / 16 * 16;
    p.patch_h = CH_TH + a.ks - 1;
    p.patch_bytes = 128 * CH_TW * p.patch_h;
    p.b_bytes = p.BN * 128;
    p.bias = a.bias; p.Cin = a.Cin;
    const int budget = 232448 - 1024;
    p.nstage = (budget - 1024 - CH_NPATCH * p.patch_bytes - 2 * CH_STG_BYTES) / p.b_bytes;
    if (p.nstage > CH_NSTAGE) p.nstage = CH_NSTAGE;
    if (p.nstage < 3) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: shared-memory plan too large"); return 8; }
    const int smem = CH_NPATCH * p.patch_bytes + p.nstage * p.b_bytes + 2 * CH_STG_BYTES + 1024;
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.ld * 2, (cuuint64_t)a.W * a.ld * 2, (cuuint64_t)a.H * a.W * a.ld * 2};
        cuuint32_t box[4] = {64, CH_TW, (cuuint32_t)p.patch_h, 1};
        CUresult r = enc(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: encode(A) failed: %d", (int)r); return 2; }
    }
    {
        const cuuint64_t Ktot = (cuuint64_t)a.ks * a.ks * a.Cpad;
        cuuint64_t dims[2] = {Ktot, (cuuint64_t)a.N};
        cuuint64_t strides[1] = {Ktot * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: encode(B) failed: %d", (int)r); return 2; }
    }
    {
        const size_t ld = (size_t)a.out_ld;
        cuuint64_t dims[4] = {(cuuint64_t)a.N, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {ld * 2, (cuuint64_t)a.W * ld * 2, (cuuint64_t)a.H * a.W * ld * 2};
        cuuint32_t box[4] = {64, CH_TW, CH_TH, 1};
        CUresult r = enc(&tm.o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.out, dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo: encode(out) failed: %d", (int)r); return 2; }
    }
    static bool attr[64] = {};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    if (!attr[dev]) {
        cudaError_t er = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
        if (er != cudaSuccess) { snprintf(g_ch_err, sizeof g_ch_err, "cudaFuncSetAttribute(conv_halo): %s", cudaGetErrorString(er)); return 4; }
        attr[dev] = true;
    }
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.ntiles < sms ? p.ntiles : sms));
    cfg.blockDim = dim3((unsigned)CH_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    static const int dbgmode = getenv("MLIC_TC_DEBUG") ? atoi(getenv("MLIC_TC_DEBUG")) : 0;
    unsigned long long* dbg = nullptr;
    if (dbgmode & 32) {
        static unsigned long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc((void**)&dbuf, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dbuf, 0, 16 * sizeof(unsigned long long), s);
        dbg = dbuf;
    }
    cudaError_t er = cudaLaunchKernelEx(&cfg, conv_halo_kernel, tm, p, dbg);
    if (dbg && er == cudaSuccess) {
        static int printed = 0;
        unsigned long long h[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        if (printed++ < 4) {
            const double tiles = (double)((p.ntiles + cfg.gridDim.x - 1) / cfg.gridDim.x);
            fprintf(stderr, "[conv halo dbg] tiles/cta %.0f ks %d chunks %d BN %d | per tile: tma total %.0f wait-patch-free %.0f wait-stage-free %.0f | mma total %.0f wait-acc-free %.0f wait-operands %.0f | epilogue w0 total %.0f wait-acc %.0f\n",
                    tiles, p.ks, p.kchunks, p.BN, h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles, h[6] / tiles, h[7] / tiles);
        }
    }
    if (er != cudaSuccess) { snprintf(g_ch_err, sizeof g_ch_err, "conv_halo launch: %s (smem %d)", cudaGetErrorString(er), smem); return 5; }
    return 0;
}

}  // namespace mlic
