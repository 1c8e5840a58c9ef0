// Three chained per-pixel layers in ONE launch, bf16 fast mode: the intermediate activations never leave the SM.
//
//   C3_EP    : the tail of EntropyParameters (modules/transform/entropy.py:13-17): Conv1x1(320 -> 256) GELU Conv1x1(256 -> 128) GELU
//              Conv1x1(128 -> 2C), fp32 output (the mu / sigma rows of the checkerboard-squeezed half image)
//   C3_LOCAL : the tail of LocalContext (modules/transform/context.py:108-110) on the squeezed non-anchor pixels:
//              p = proj(fusion(windows)) [one folded GEMM, K = 25 C];  out = p + fc2(GELU(fc1(LayerNorm(p)))), bf16 output
//
// Per 128-row tile: GEMM 1 streams A (the rows) and W1 through a TMA ring into a TMEM accumulator; the epilogue warps turn it into the
// bf16 A OPERAND OF THE NEXT GEMM IN TENSOR MEMORY (bias, GELU or LayerNorm, tcgen05.st: row = lane, 2 K elements per 32-bit column), so
// GEMM 2 and GEMM 3 are tcgen05.mma with A from TMEM and B (W2 resident in shared memory, W3 through the ring) from shared memory.
// Four launches (three GEMMs + LayerNorm) and their HBM round trips become one; the launch reads its input rows once and writes 2C
// values per row.  TMEM columns: acc1 [0, 256), H1 [256, 384), acc2 [384, 512), H2 [256, 320) and acc3 [320, 384) (H1 is dead by then).
//
// 640 threads: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM allocation, warps 4..19 epilogue (TMEM lane quarter = warp % 4, four warps
// per quarter share the columns of each phase).
#include "kernels.h"
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlic {

enum { C3_EP = 0, C3_LOCAL = 1 };
constexpr int C3_THREADS = 128 + 16 * 32;
constexpr int C3_MAXSTAGE = 6;
constexpr int C3_A_BYTES = 128 * 128;
constexpr uint32_t C3_ACC1 = 0, C3_H1 = 256, C3_ACC2 = 384, C3_H2 = 256, C3_ACC3 = 320;

struct C3Params {
    int M, ntiles;
    int K1, kch1;                   // real K of GEMM 1, its 64-wide chunks
    int N1, N2, N3;
    int nstage, stage_bytes;
    const float *b1, *b2, *b3, *ln_g, *ln_b;
    void* out; int out_ld;
    float ln_eps;
    int unsq_H, unsq_W;             // LOCAL: != 0: row r is the r-th NON-ANCHOR pixel of a [.., H, W] grid (utils/ckbd.py squeeze order); `out` is that grid
};
struct C3Maps { CUtensorMap a, w1, w2, w3; };

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void c3_tmem_st8(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void c3_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t c3_pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// acc columns [c, c + 32) of this thread's row (already in `raw`) -> bias, GELU, bf16 pairs -> 16 columns of the next A operand
__device__ __forceinline__ void c3_gelu_block(const uint32_t raw[32], const float* __restrict__ sb, uint32_t dst) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(&sb[2 * j]);
        float2 v = __fadd2_rn(make_float2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1])), b);
        v = gelu2(v);
        pk[j] = c3_pack(v.x, v.y);
    }
    c3_tmem_st8(dst, pk);
    c3_tmem_st8(dst + 8, pk + 8);
}

template <int MODE>
__global__ void __launch_bounds__(C3_THREADS, 1)
chain3_kernel(const __grid_constant__ C3Maps tm, const C3Params p, unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t c3_smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)c3_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[C3_MAXSTAGE], empty[C3_MAXSTAGE];
    __shared__ uint64_t w2_bar, acc1_full, h1_full, acc2_full, h2_full, acc3_full, tile_done;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float sB1[256], sB2[128], sB3[64], sG[64], sBt[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long tw[6] = {0, 0, 0, 0, 0, 0};
    const long long t_start = dbg ? clock64() : 0;
#define C3_TIMED(acc, stmt) do { if (dbg) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)
    const int kch2 = p.N1 >> 6, kch3 = p.N2 >> 6;
    const int w2_chunk = p.N2 * 128, w3_chunk = p.N3 * 128;
    // GEMM 3 of a tile is issued after the first `split` K chunks of GEMM 1 of the next tile (about the time phase 2 takes), the rest of
    // that GEMM 1 behind it: the tensor pipe runs in issue order, and GEMM 3 must not wait behind a whole operand-feed-bound GEMM 1
    const int split = min(p.kch1, MODE == C3_LOCAL ? 3 : 2);
    uint8_t* w2s = base;
    uint8_t* ring = base + (size_t)kch2 * w2_chunk;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.w2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.w3) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&w2_bar, 1); mbar_init(&acc1_full, 1); mbar_init(&acc2_full, 1); mbar_init(&acc3_full, 1);
        mbar_init(&h1_full, 16); mbar_init(&h2_full, 16); mbar_init(&tile_done, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += C3_THREADS) sB1[i] = i < p.N1 ? p.b1[i] : 0.f;
    for (int i = threadIdx.x; i < 128; i += C3_THREADS) sB2[i] = i < p.N2 ? p.b2[i] : 0.f;
    for (int i = threadIdx.x; i < 64; i += C3_THREADS) {
        sB3[i] = i < p.N3 ? p.b3[i] : 0.f;
        sG[i] = (MODE == C3_LOCAL && i < p.N1) ? p.ln_g[i] : 1.f;
        sBt[i] = (MODE == C3_LOCAL && i < p.N1) ? p.ln_b[i] : 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&w2_bar, (uint32_t)(kch2 * w2_chunk));
            for (int k = 0; k < kch2; ++k) tma_load_2d(w2s + (size_t)k * w2_chunk, &tm.w2, &w2_bar, k * 64, 0);
            int s = 0;
            uint32_t ph = 0;
            // ring order = the order the MMA thread consumes it: GEMM 1 of the first tile, then per tile GEMM 1 of the NEXT tile and W3 of this one
            auto load_l1 = [&](int t, int c0, int c1) {
                for (int cc = c0; cc < c1; ++cc) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + (size_t)s * p.stage_bytes;
                    mbar_expect_tx(&full[s], (uint32_t)(C3_A_BYTES + p.N1 * 128));
                    tma_load_2d(st, &tm.a, &full[s], cc * 64, t * 128);
                    tma_load_2d(st + C3_A_BYTES, &tm.w1, &full[s], cc * 64, 0);
                    if (++s == p.nstage) { s = 0; ph ^= 1; }
                }
            };
            if ((int)blockIdx.x < p.ntiles) load_l1((int)blockIdx.x, 0, p.kch1);
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int tn = t + (int)gridDim.x;
                if (tn < p.ntiles) load_l1(tn, 0, split);
                for (int k = 0; k < kch3; ++k) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)w3_chunk);
                    tma_load_2d(ring + (size_t)s * p.stage_bytes, &tm.w3, &full[s], k * 64, 0);
                    if (++s == p.nstage) { s = 0; ph ^= 1; }
                }
                if (tn < p.ntiles) load_l1(tn, split, p.kch1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
            const uint32_t idesc1 = idesc0 | ((uint32_t)(p.N1 >> 3) << 17), idesc2 = idesc0 | ((uint32_t)(p.N2 >> 3) << 17), idesc3 = idesc0 | ((uint32_t)(p.N3 >> 3) << 17);
            int s = 0, it = 0;
            uint32_t ph = 0;
            // GEMM 1 of tile it + 1 is issued between GEMM 2 and GEMM 3 of tile it: its accumulator (EP: columns [0, 256), free once phase 1 of
            // tile it has read them; LOCAL: the other 64-column slot, last read by phase 3 of tile it - 1) fills while the epilogue warps
            // work on tile it, and the tensor pipe executes the MMAs in issue order
            auto issue_l1 = [&](int itn, int c0, int c1) {
                const uint32_t acc1 = tmem_base + C3_ACC1 + (MODE == C3_LOCAL ? (uint32_t)((itn & 1) * 64) : 0u);
                for (int cc = c0; cc < c1; ++cc) {
                    const int nk = min(4, (p.K1 - cc * 64 + 15) >> 4);
                    C3_TIMED(tw[0], mbar_wait(&full[s], ph));
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)s * p.stage_bytes);
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + (uint32_t)C3_A_BYTES);
                    for (int kk = 0; kk < nk; ++kk) umma_bf16(acc1, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc1, (cc | kk) ? 1u : 0u);
                    tcgen05_commit(&empty[s]);
                    if (++s == p.nstage) { s = 0; ph ^= 1; }
                }
                if (c1 == p.kch1) tcgen05_commit(&acc1_full);
            };
            if ((int)blockIdx.x < p.ntiles) issue_l1(0, 0, p.kch1);
            mbar_wait(&w2_bar, 0);
            tcgen05_fence_after();
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const uint32_t tp = (uint32_t)it & 1u;
                C3_TIMED(tw[1], mbar_wait(&h1_full, tp));
                tcgen05_fence_after();
                for (int k = 0; k < kch2; ++k) {
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(w2s + (size_t)k * w2_chunk));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)          // 16 bf16 of K = 8 TMEM columns per step
                        umma_ts(tmem_base + C3_ACC2, tmem_base + C3_H1 + (uint32_t)((k * 4 + kk) * 8), bdesc + (uint64_t)(kk * 2), idesc2, (k | kk) ? 1u : 0u);
                }
                tcgen05_commit(&acc2_full);
                const bool more = t + (int)gridDim.x < p.ntiles;
                if (more) issue_l1(it + 1, 0, split);
                C3_TIMED(tw[2], mbar_wait(&h2_full, tp));
                tcgen05_fence_after();
                for (int k = 0; k < kch3; ++k) {
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + (size_t)s * p.stage_bytes));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_ts(tmem_base + C3_ACC3, tmem_base + C3_H2 + (uint32_t)((k * 4 + kk) * 8), bdesc + (uint64_t)(kk * 2), idesc3, (k | kk) ? 1u : 0u);
                    tcgen05_commit(&empty[s]);
                    if (++s == p.nstage) { s = 0; ph ^= 1; }
                }
                tcgen05_commit(&acc3_full);
                if (more) issue_l1(it + 1, split, p.kch1);
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = (unsigned long long)(clock64() - t_start); dbg[1] = (unsigned long long)tw[0]; dbg[2] = (unsigned long long)tw[1]; dbg[3] = (unsigned long long)tw[2]; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, sub = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
            const uint32_t tp = (uint32_t)it & 1u;
            const long long grow = (long long)t * 128 + row;
            // ---- phase 1: acc1 -> H1.  H1 overlays H2 / acc3 of the previous tile: no warp may store before every warp has finished
            // its phase 3 (tile_done doubles as the barrier among the epilogue warps)
            const uint32_t acc1c = C3_ACC1 + (MODE == C3_LOCAL ? (uint32_t)((it & 1) * 64) : 0u);
            if (it > 0) C3_TIMED(tw[3], mbar_wait(&tile_done, tp ^ 1u));
            C3_TIMED(tw[0], mbar_wait(&acc1_full, tp));
            tcgen05_fence_after();
            if constexpr (MODE == C3_EP) {
                // columns of this warp: [sub * N1 / 4, + N1 / 4) in blocks of 32; the loads of block i + 1 are in flight while block i is computed
                const int w1c = p.N1 >> 2;
                uint32_t raw[2][32];
                int c = sub * w1c;
                const int cend = c + w1c;
                tmem_ld16(lane_base + acc1c + (uint32_t)c, raw[0]);
                tmem_ld16(lane_base + acc1c + (uint32_t)(c + 16), raw[0] + 16);
                tmem_ld_wait();
#pragma unroll 1
                for (; c < cend; c += 64) {
                    const bool more = c + 32 < cend;
                    if (more) {
                        tmem_ld16(lane_base + acc1c + (uint32_t)(c + 32), raw[1]);
                        tmem_ld16(lane_base + acc1c + (uint32_t)(c + 48), raw[1] + 16);
                    }
                    c3_gelu_block(raw[0], sB1 + c, lane_base + C3_H1 + (uint32_t)(c >> 1));
                    if (more) {
                        tmem_ld_wait();
                        if (c + 64 < cend) {
                            tmem_ld16(lane_base + acc1c + (uint32_t)(c + 64), raw[0]);
                            tmem_ld16(lane_base + acc1c + (uint32_t)(c + 80), raw[0] + 16);
                        }
                        c3_gelu_block(raw[1], sB1 + c + 32, lane_base + C3_H1 + (uint32_t)((c + 32) >> 1));
                        if (c + 64 < cend) tmem_ld_wait();
                    }
                }
            } else {
                // LayerNorm over the N1 = 64 columns of the row: every warp of the quarter reads the whole row once (64 registers), computes
                // its statistics in two passes over the registers, then normalises its own 16 columns (re-read: `sub` is not a constant)
                uint32_t raw[64];
                tmem_ld16(lane_base + acc1c, raw);
                tmem_ld16(lane_base + acc1c + 16u, raw + 16);
                tmem_ld16(lane_base + acc1c + 32u, raw + 32);
                tmem_ld16(lane_base + acc1c + 48u, raw + 48);
                tmem_ld_wait();
                float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 64; ++j) { const float x = __uint_as_float(raw[j]) + sB1[j]; raw[j] = __float_as_uint(x); m4[j & 3] += x; }
                const float mean = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.f / 64.f);
                float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 64; ++j) { const float d = __uint_as_float(raw[j]) - mean; v4[j & 3] = fmaf(d, d, v4[j & 3]); }
                const float rstd = rsqrtf(((v4[0] + v4[1]) + (v4[2] + v4[3])) * (1.f / 64.f) + p.ln_eps);
                const int c = sub * 16;
                uint32_t own[16], pk[8];
                tmem_ld16(lane_base + acc1c + (uint32_t)c, own);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x0 = (__uint_as_float(own[2 * j]) + sB1[c + 2 * j]) - mean, x1 = (__uint_as_float(own[2 * j + 1]) + sB1[c + 2 * j + 1]) - mean;
                    pk[j] = c3_pack(fmaf(x0 * rstd, sG[c + 2 * j], sBt[c + 2 * j]), fmaf(x1 * rstd, sG[c + 2 * j + 1], sBt[c + 2 * j + 1]));
                }
                c3_tmem_st8(lane_base + C3_H1 + (uint32_t)(c >> 1), pk);
            }
            c3_tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full);
            // ---- phase 2: acc2 -> H2 (bias, GELU)
            C3_TIMED(tw[1], mbar_wait(&acc2_full, tp));
            tcgen05_fence_after();
            {
                const int c = sub * 32;                             // N2 = 128: one block of 32 columns per warp
                uint32_t raw[32];
                tmem_ld16(lane_base + C3_ACC2 + (uint32_t)c, raw);
                tmem_ld16(lane_base + C3_ACC2 + (uint32_t)(c + 16), raw + 16);
                tmem_ld_wait();
                c3_gelu_block(raw, sB2 + c, lane_base + C3_H2 + (uint32_t)(c >> 1));
            }
            c3_tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_full);
            // ---- phase 3: acc3 (+ the folded projection, LOCAL) -> memory; N3 = 64: 16 columns per warp
            C3_TIMED(tw[2], mbar_wait(&acc3_full, tp));
            tcgen05_fence_after();
            {
                const int c = sub * 16;
                uint32_t raw[16];
                float v[16];
                tmem_ld16(lane_base + C3_ACC3 + (uint32_t)c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + sB3[c + j];
                if constexpr (MODE == C3_LOCAL) {
                    tmem_ld16(lane_base + acc1c + (uint32_t)c, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(raw[j]) + sB1[c + j];
                }
                if (grow < p.M) {
                    if constexpr (MODE == C3_EP) {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)grow * p.out_ld + c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        long long orow = grow;
                        if (p.unsq_W) {                             // squeezed row -> its pixel: w = 2 j + (h & 1)
                            const int W2 = p.unsq_W >> 1;
                            const long long bh = grow / W2;
                            const int j = (int)(grow - bh * W2), h = (int)(bh % p.unsq_H);
                            orow = bh * p.unsq_W + 2 * j + (h & 1);
                        }
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (size_t)orow * p.out_ld + c);
                        o[0] = make_uint4(c3_pack(v[0], v[1]), c3_pack(v[2], v[3]), c3_pack(v[4], v[5]), c3_pack(v[6], v[7]));
                        o[1] = make_uint4(c3_pack(v[8], v[9]), c3_pack(v[10], v[11]), c3_pack(v[12], v[13]), c3_pack(v[14], v[15]));
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_done);
        }
        if (dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { dbg[4] = (unsigned long long)(clock64() - t_start); dbg[5] = (unsigned long long)tw[0]; dbg[6] = (unsigned long long)tw[1]; dbg[7] = (unsigned long long)tw[2]; dbg[8] = (unsigned long long)tw[3]; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled_c3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local char g_c3_err[256] = "";
const char* chain3_last_error() { return g_c3_err; }

bool chain3_supported(const Chain3Args& a) {
    if (!a.in || !a.w1 || !a.w2 || !a.w3 || !a.b1 || !a.b2 || !a.b3 || !a.out) return false;
    if (a.M <= 0 || a.K1 < 16 || (a.K1 % 8) != 0 || a.K1pad % 64 != 0 || a.K1pad < a.K1 || (a.ld % 8) != 0) return false;
    if (a.N2 != 128 || a.N3 != 64) return false;
    if (a.mode == 0) { if (a.N1 % 128 != 0 || a.N1 < 128 || a.N1 > 256 || (a.out_ld % 4) != 0 || ((uintptr_t)a.out % 16) != 0) return false; }
    else if (a.mode == 1) { if (a.N1 != 64 || !a.ln_g || !a.ln_b || (a.out_ld % 8) != 0 || ((uintptr_t)a.out % 16) != 0) return false; }
    else return false;
    if (((uintptr_t)a.in % 16) != 0 || ((uintptr_t)a.w1 % 16) != 0 || ((uintptr_t)a.w2 % 16) != 0 || ((uintptr_t)a.w3 % 16) != 0) return false;
    return true;
}

int launch_chain3(const Chain3Args& a, cudaStream_t s) {
    if (tc_init()) { snprintf(g_c3_err, sizeof g_c3_err, "%s", tc_last_error()); return 1; }
    if (!chain3_supported(a)) { snprintf(g_c3_err, sizeof g_c3_err, "chain3: unsupported layer chain"); return 1; }
    PFN_encodeTiled_c3 enc = (PFN_encodeTiled_c3)tc_encode_fn();
    C3Maps tm;
    memset(&tm, 0, sizeof tm);
    C3Params p;
    memset(&p, 0, sizeof p);
    p.M = a.M; p.ntiles = (a.M + 127) / 128;
    p.K1 = a.K1; p.kch1 = a.K1pad / 64; p.N1 = a.N1; p.N2 = a.N2; p.N3 = a.N3;
    p.b1 = a.b1; p.b2 = a.b2; p.b3 = a.b3; p.ln_g = a.ln_g; p.ln_b = a.ln_b; p.ln_eps = a.ln_eps;
    p.out = a.out; p.out_ld = a.out_ld; p.unsq_H = a.unsq_H; p.unsq_W = a.unsq_W;
    p.stage_bytes = C3_A_BYTES + a.N1 * 128;
    const int resident = (a.N1 / 64) * a.N2 * 128;
    const int budget = 232448 - 4096 - 1024;            // (static shared memory: barriers + biases, ~2.5 KB)
    p.nstage = (budget - resident) / p.stage_bytes;
    if (p.nstage > C3_MAXSTAGE) p.nstage = C3_MAXSTAGE;
    if (p.nstage < 2) { snprintf(g_c3_err, sizeof g_c3_err, "chain3: shared-memory plan too large"); return 8; }
    const int smem = resident + p.nstage * p.stage_bytes + 1024;
    const cuuint32_t estr[2] = {1, 1};
    auto enc2 = [&](CUtensorMap* m, const void* ptr, cuuint64_t inner, cuuint64_t rows, cuuint64_t pitch_elems, cuuint32_t box_rows, CUtensorMapL2promotion prom) {
        cuuint64_t dims[2] = {inner, rows};
        cuuint64_t strides[1] = {pitch_elems * 2};
        cuuint32_t box[2] = {64, box_rows};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = enc2(&tm.a, a.in, (cuuint64_t)a.K1, (cuuint64_t)a.M, (cuuint64_t)a.ld, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (r == CUDA_SUCCESS) r = enc2(&tm.w1, a.w1, (cuuint64_t)a.K1pad, (cuuint64_t)a.N1, (cuuint64_t)a.K1pad, (cuuint32_t)a.N1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r == CUDA_SUCCESS) r = enc2(&tm.w2, a.w2, (cuuint64_t)a.N1, (cuuint64_t)a.N2, (cuuint64_t)a.N1, (cuuint32_t)a.N2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r == CUDA_SUCCESS) r = enc2(&tm.w3, a.w3, (cuuint64_t)a.N2, (cuuint64_t)a.N3, (cuuint64_t)a.N2, (cuuint32_t)a.N3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r != CUDA_SUCCESS) { snprintf(g_c3_err, sizeof g_c3_err, "chain3: cuTensorMapEncodeTiled failed: %d", (int)r); return 2; }
    static bool attr[64][2] = {};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    auto fn = a.mode == 0 ? chain3_kernel<0> : chain3_kernel<1>;
    if (!attr[dev][a.mode]) {
        cudaError_t er = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 4096);
        if (er != cudaSuccess) { snprintf(g_c3_err, sizeof g_c3_err, "cudaFuncSetAttribute(chain3): %s", cudaGetErrorString(er)); return 4; }
        attr[dev][a.mode] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.ntiles < sms ? p.ntiles : sms));
    cfg.blockDim = dim3((unsigned)C3_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    static const int dbgmode = getenv("MLIC_TC_DEBUG") ? atoi(getenv("MLIC_TC_DEBUG")) : 0;
    unsigned long long* dbg = nullptr;
    if (dbgmode & 32) {
        static unsigned long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc((void**)&dbuf, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dbuf, 0, 16 * sizeof(unsigned long long), s);
        dbg = dbuf;
    }
    cudaError_t er = cudaLaunchKernelEx(&cfg, fn, tm, p, dbg);
    if (dbg && er == cudaSuccess) {
        static int printed[2] = {0, 0};
        unsigned long long h[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        if (printed[a.mode]++ < 3) {
            const double tiles = (double)((p.ntiles + cfg.gridDim.x - 1) / cfg.gridDim.x);
            fprintf(stderr, "[chain3 dbg mode %d] tiles/cta %.0f stages %d | per tile: mma total %.0f wait-ring %.0f wait-h1 %.0f wait-h2 %.0f | epilogue w4 total %.0f wait-acc1 %.0f wait-acc2 %.0f wait-acc3 %.0f wait-tile-done %.0f\n",
                    a.mode, tiles, p.nstage, h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles, h[6] / tiles, h[7] / tiles, h[8] / tiles);
        }
    }
    if (er != cudaSuccess) { snprintf(g_c3_err, sizeof g_c3_err, "chain3 launch: %s (smem %d)", cudaGetErrorString(er), smem); return 5; }
    return 0;
}

}  // namespace mlic
