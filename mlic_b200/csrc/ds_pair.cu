// Two-SM (tcgen05 cta_group::2) fused blocks of the full- and half-resolution transforms, bf16 NHWC, C = 64 * NCH channels
// in and out (MLICPP_L: 192).  One launch runs, per 8 x 16 pixel tile,
//
//   MODE_DS   : out = act(pw(dw3x3(x) + bd) + b1) [+ res]                      DepthWiseConv (modules/layers/conv.py:46-63)
//                                                                               with the GELU / skip of res_blk.py:142-154
//   MODE_TAIL : v = pw(dw3x3(x) + bd) + b1 ;  out = v * (r)sqrt(gamma v^2 + beta) + res
//               the tail of ResidualBlockWithStride (res_blk.py:88-93: conv2 -> GDN -> + skip) and of
//               ResidualBlockUpsample (res_blk.py:116-121: conv -> IGDN -> + upsample): v and v^2 never leave the SM
//               (three HBM passes -- x, res, out -- instead of the seven of conv kernel + GDN kernel).
//
// Why a CTA PAIR: both 192 x 192 weight matrices must stay resident (re-streaming 72 KB per 128-pixel tile would put
// 9.4 GB per launch on L2), and two of them (144 KB) do not fit next to the halo ring, the A stages and the staging
// blocks of one SM.  With cta_group::2 the MMA is M = 256 (128 pixels per CTA), N = C, and each CTA holds only ITS HALF
// of the N rows of every weight matrix (36 KB each): the pair reads the other half through the tensor core's peer path.
//
// Per CTA, 768 threads: warp 0 TMA (weights once, then the halo patches {64 ch, 18, 10} of its tile), warp 1 MMA issuer
// (leader CTA only: one thread issues for the pair), warp 2 TMEM allocation, warps 4..11 depthwise producers (lane =
// channel pair, FFMA2 against 9 taps in registers) writing the K-major SWIZZLE_128B A stages, warps 12.. epilogue (one
// 4-warp group per 64 output columns: TMEM -> registers -> swizzled staging block -> TMA store; the residual block is
// TMA-loaded into the same staging block while the accumulator is still being computed).
// MODE_TAIL keeps v (fp32) in TMEM columns [0, C), writes v^2 as the bf16 A operand of the GDN GEMM into TMEM columns
// [2C, 2C + C/2) (tcgen05.st; the second MMA takes A from TMEM), accumulates gamma v^2 into columns [C, 2C) and reads
// both accumulators back for x * rsqrt(.) -- v is used in fp32, where the two-kernel path re-read it as bf16.
#include "kernels.h"
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlic {

__device__ __forceinline__ void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// non-volatile shared loads for the producer's inner loop (the compiler may schedule them; the caller pins them behind
// the barrier wait by laundering the base address through an asm statement that follows the wait)
__device__ __forceinline__ uint32_t lds32_nv(uint32_t a) { uint32_t v; asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f2_nv(uint32_t a) { float2 v; asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
// bf16x2 word -> two fp32, both on the ALU pipe (PRMT / LOP3): the FMA pipe is the scarce one in these kernels
__device__ __forceinline__ float2 bf2_unpack_alu(uint32_t w) {
    uint32_t lo;
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(w));
    return make_float2(__uint_as_float(lo), __uint_as_float(w & 0xffff0000u));
}

enum { DP_DS = 0, DP_TAIL = 1 };
constexpr int DP_TH = 8, DP_TW = 16;
constexpr int DP_RAW_BYTES = (DP_TH + 2) * (DP_TW + 2) * 128;       // one halo patch of a 64-channel chunk: 23 040 B
constexpr int DP_RAW_SLOTS_MAX = 4;
constexpr int DP_A_BYTES = 128 * 128;
constexpr int DP_STG_BYTES = 128 * 128;
#ifndef MLIC_DP_PROD_WARPS
#define MLIC_DP_PROD_WARPS 8
#endif
// 8: one warp = 2 columns x 8 rows of the tile; 16: 2 columns x 4 rows (1024 threads, registers traded between the roles with
// setmaxnreg).  Measured at 4 x 544 x 960 x 192 (profiles/r02_pair_variants.txt): 16 warps make the producers fast enough that the
// MMA never waits for A, but they take their issue slots from the epilogue warps, which then pace the tile: 460 -> 491 us without
// and 514 -> 527 us with a residual.  The kernel is issue-bound (19 instructions per output value, ~50 % of the issue slots), not
// short of one role's warps.
constexpr int DP_PROD_WARPS = MLIC_DP_PROD_WARPS;
constexpr int DP_PROD_ROWS = DP_TH * 8 / DP_PROD_WARPS;            // output rows per producer warp
// register budgets per role (setmaxnreg, warpgroup-granular): 1024 threads leave 64 registers per thread on average
#ifndef MLIC_DP_REGS_MISC
#define MLIC_DP_REGS_MISC 40
#endif
#ifndef MLIC_DP_REGS_EPI
#define MLIC_DP_REGS_EPI 80
#endif
#ifndef MLIC_DP_REGS_PROD
#define MLIC_DP_REGS_PROD 56
#endif

struct DpParams {
    int tilesH, tilesW, ntiles, npairs;
    int Hout, Wout;
    const float* dw_w9;     // [9][C]
    const float* dw_bias;   // [C]
    const float* b1;        // pointwise bias [C]
    const float* b2;        // GDN beta [C] (DP_TAIL)
};
struct DpMaps {
    CUtensorMap raw;        // input, box {64, 18, 10, 1}, no swizzle
    CUtensorMap w1, w2;     // weights [C][C] bf16, box {64, C/2}, 128B swizzle
    CUtensorMap out, res;   // box {64, 16, 8, 1}, 128B swizzle
};

template <int NCH, int MODE, bool RES = false> struct DpCfg {
    static constexpr int C = 64 * NCH;
    static constexpr int WCH_BYTES = (C / 2) * 128;                 // one K chunk of this CTA's half of a weight matrix
    static constexpr int W_BYTES = NCH * WCH_BYTES;
    static constexpr int NA = MODE == DP_TAIL ? NCH : (NCH < 2 ? NCH : 2);     // A stages
    // staging slots per epilogue group / halo patches in flight.  DS with a residual: two slots, so that the residual block of the
    // next tile lands while this one is computed (the wait for it and for the drain of the previous store were 3.1 k of the
    // 9.2 k clocks of a tile), paid for with two of the four halo slots; TAIL: shared memory is full
    static constexpr int NRING = (MODE == DP_DS && RES) ? 2 : 1;
    static constexpr int RAW_SLOTS = (MODE == DP_TAIL || RES) ? 2 : 4;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = W_BYTES;
    static constexpr int OFF_A = (MODE == DP_TAIL ? 2 : 1) * W_BYTES;
    static constexpr int OFF_STG = OFF_A + NA * DP_A_BYTES;
    static constexpr int OFF_RAW = OFF_STG + NCH * NRING * DP_STG_BYTES;
    static constexpr int OFF_DW = OFF_RAW + RAW_SLOTS * DP_RAW_BYTES;
    static constexpr int OFF_B1 = OFF_DW + 10 * C * 4;
    static constexpr int OFF_B2 = OFF_B1 + C * 4;
    static constexpr int SMEM = OFF_B2 + C * 4 + 1024;              // + alignment slack
    static_assert(SMEM <= 232448, "shared-memory plan exceeds 227 KB");
    static constexpr int EPI_WARPS = 4 * NCH;
    static constexpr int THREADS = 128 + (DP_PROD_WARPS + EPI_WARPS) * 32;
    // 1024 threads: 64 registers per thread at launch; the roles trade them (misc 40 / epilogue 80 / producers 56 per thread)
    static constexpr bool SETREG = THREADS > 896;
    static_assert(!SETREG || 4 * MLIC_DP_REGS_MISC + EPI_WARPS * MLIC_DP_REGS_EPI + DP_PROD_WARPS * MLIC_DP_REGS_PROD <= 2048, "register plan");
    static constexpr int EW0 = 4;                                   // epilogue warps 4 .. 4 + EPI_WARPS - 1 (multiple of 4: TMEM lane quarter = warp % 4)
    static constexpr int PW0 = 4 + EPI_WARPS;                       // producers get the highest warp ids: the issue arbiter prefers them, and they pace the tile
    // TMEM columns (512 allocated): DS: two accumulator stages at 0 and 256; TAIL: v at [0, C), the bf16 v^2 A operand at
    // [C, C + C/2), gamma v^2 at [320, 320 + C)
    static constexpr int TM_D2 = MODE == DP_TAIL ? 320 : 256;
    static constexpr int TM_A2 = C;
    static_assert(MODE != DP_TAIL || (C + C / 2 <= 320 && 320 + C <= 512), "TMEM plan");
};

template <int NCH, int MODE, int ACT, bool RES, int GDN>
__global__ void __launch_bounds__(DpCfg<NCH, MODE, RES>::THREADS, 1)
ds_pair_kernel(const __grid_constant__ DpMaps tm, const DpParams p, unsigned long long* __restrict__ dbg) {
    using Cfg = DpCfg<NCH, MODE, RES>;
    constexpr int C = Cfg::C;
    extern __shared__ uint8_t dp_smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)dp_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t raw_full[DP_RAW_SLOTS_MAX], raw_empty[DP_RAW_SLOTS_MAX];
    __shared__ uint64_t a_full[NCH], a_empty[NCH];              // a_full: leader's instance is the live one
    __shared__ uint64_t w_full, w_ready;                        // w_ready (leader): both CTAs hold their weights
    __shared__ uint64_t d_full[2], d_empty[2];                  // DS: accumulator stages; TAIL: [0] = v ready / all read, [1] = gamma v^2 ready
    __shared__ uint64_t a2_full;                                // TAIL (leader): v^2 operand written by both CTAs
    __shared__ uint64_t stg_bar[NCH][2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    float* sDw = reinterpret_cast<float*>(base + Cfg::OFF_DW);      // [chunk][tap 0..8 | bias][64]
    float* sB1 = reinterpret_cast<float*>(base + Cfg::OFF_B1);
    float* sB2 = reinterpret_cast<float*>(base + Cfg::OFF_B2);

    // GELU layers: the accumulator is made to hold HALF the pre-activation (depthwise taps, depthwise bias and pointwise bias scaled by
    // 1/2 -- exact), which gelu2_half() turns into the same bits as gelu2() of the full value with one FP32x2 multiply less
    constexpr bool HALF = (MODE == DP_DS && ACT == ACT_GELU && MLIC_GELU_FORM == 0);
    constexpr float PS = HALF ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < 10 * C; i += Cfg::THREADS) {
        const int k = i / 640, rem = i - k * 640, tap = rem >> 6, ch = rem & 63;
        sDw[i] = PS * (tap < 9 ? p.dw_w9[tap * C + k * 64 + ch] : p.dw_bias[k * 64 + ch]);
    }
    for (int i = threadIdx.x; i < C; i += Cfg::THREADS) { sB1[i] = PS * p.b1[i]; sB2[i] = (MODE == DP_TAIL) ? p.b2[i] : 0.f; }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.raw) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::RAW_SLOTS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], DP_PROD_WARPS); }
        for (int s = 0; s < NCH; ++s) { mbar_init(&a_full[s], 2 * DP_PROD_WARPS); mbar_init(&a_empty[s], 1); }
        mbar_init(&w_full, 1); mbar_init(&w_ready, 2);
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 2 * Cfg::EPI_WARPS); }
        mbar_init(&a2_full, 2 * Cfg::EPI_WARPS);
        for (int i = 0; i < NCH; ++i) { mbar_init(&stg_bar[i][0], 1); mbar_init(&stg_bar[i][1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers are initialised before anyone arrives on them remotely
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // programmatic dependent launch: everything above touches parameters only
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // development (MLIC_TC_DEBUG & 32): per-role wait / work clocks of CTA 0's producer warp 4, first epilogue warp and MMA thread
    const long long t_start = dbg ? clock64() : 0;
#define DP_TIMED(acc, stmt) do { if (dbg) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)
    const int tiles_per_img = p.tilesH * p.tilesW;
    const int pair_first = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);
    // tile of this CTA in pair `pp`: 2 pp + rank (an odd tile count leaves the last pair's second CTA a duplicate of the last tile, not stored)
#define DP_TILE(pp)                                                                           \
    int tix = 2 * (pp) + (int)rank;                                                           \
    const bool tvalid = tix < p.ntiles;                                                       \
    if (!tvalid) tix = p.ntiles - 1;                                                          \
    const int img = tix / tiles_per_img;                                                      \
    const int trem = tix - img * tiles_per_img;                                               \
    const int th = trem / p.tilesW, tw = trem - th * p.tilesW;                                \
    const int h0 = th * DP_TH, w0 = tw * DP_TW;                                               \
    (void)tvalid; (void)h0; (void)w0; (void)img

    // the role split is ONE if / else chain and every setmaxnreg sits at the top of its arm: ptxas takes the smallest budget that can
    // reach a point, so an arm that joined a `dec` path would be compiled for that path's registers
    if (warp < 4) {
      if constexpr (Cfg::SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MLIC_DP_REGS_MISC));
      if (warp == 0) {
        if (lane == 0) {
            // this CTA's half of the N rows of each weight matrix, all K chunks, once
            mbar_expect_tx(&w_full, (uint32_t)((MODE == DP_TAIL ? 2 : 1) * Cfg::W_BYTES));
            for (int k = 0; k < NCH; ++k) {
                tma_load_2d(base + Cfg::OFF_W1 + k * Cfg::WCH_BYTES, &tm.w1, &w_full, k * 64, (int)rank * (C / 2));
                if (MODE == DP_TAIL) tma_load_2d(base + Cfg::OFF_W2 + k * Cfg::WCH_BYTES, &tm.w2, &w_full, k * 64, (int)rank * (C / 2));
            }
            int rslot = 0;
            uint32_t rphase = 0;
            for (int pp = pair_first; pp < p.npairs; pp += pair_step) {
                DP_TILE(pp);
                for (int k = 0; k < NCH; ++k) {
                    mbar_wait(&raw_empty[rslot], rphase ^ 1);
                    mbar_expect_tx(&raw_full[rslot], (uint32_t)DP_RAW_BYTES);
                    tma_load_4d(base + Cfg::OFF_RAW + rslot * DP_RAW_BYTES, &tm.raw, &raw_full[rslot], k * 64, w0 - 1, h0 - 1, img);
                    if (++rslot == Cfg::RAW_SLOTS) { rslot = 0; rphase ^= 1; }
                }
            }
        }
      } else if (warp == 1) {
        if (lane == 0) {
            long long tw0 = 0, tw1 = 0, tw2 = 0;
            // both CTAs report their weights to the leader
            mbar_wait(&w_full, 0);
            mbar_arrive_cluster(mapa_u32(smem_u32(&w_ready), 0));
            if (leader) {
                // instruction descriptor: D = f32, A = B = bf16, K-major, N = C, M = 256 (128 rows per CTA)
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((256u >> 4) << 24);
                mbar_wait_cl(&w_ready, 0);
                tcgen05_fence_after();
                const uint32_t w1s = smem_u32(base + Cfg::OFF_W1), w2s = smem_u32(base + Cfg::OFF_W2), as = smem_u32(base + Cfg::OFF_A);
                int it = 0;
                int stage = 0;
                uint32_t sphase = 0;
                for (int pp = pair_first; pp < p.npairs; pp += pair_step, ++it) {
                    if constexpr (MODE == DP_DS) {
                        const int acc = it & 1;
                        DP_TIMED(tw0, mbar_wait_cl(&d_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u));
                        tcgen05_fence_after();
                        const uint32_t dcol = tmem_base + (uint32_t)(acc * Cfg::TM_D2);
                        for (int k = 0; k < NCH; ++k) {
                            DP_TIMED(tw1, mbar_wait_cl(&a_full[stage], sphase));
                            tcgen05_fence_after();
                            const uint64_t adesc = umma_desc_sw128(as + (uint32_t)(stage * DP_A_BYTES));
                            const uint64_t bdesc = umma_desc_sw128(w1s + (uint32_t)(k * Cfg::WCH_BYTES));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma2_ss(dcol, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                            commit_pair(&a_empty[stage]);
                            if (++stage == Cfg::NA) { stage = 0; sphase ^= 1; }
                        }
                        commit_pair(&d_full[acc]);
                    } else {
                        const uint32_t ph = (uint32_t)it & 1u;
                        DP_TIMED(tw0, mbar_wait_cl(&d_empty[0], ph ^ 1u));             // both accumulators of the previous tile read by both CTAs
                        tcgen05_fence_after();
                        for (int k = 0; k < NCH; ++k) {
                            DP_TIMED(tw1, mbar_wait_cl(&a_full[k], ph));
                            tcgen05_fence_after();
                            const uint64_t adesc = umma_desc_sw128(as + (uint32_t)(k * DP_A_BYTES));
                            const uint64_t bdesc = umma_desc_sw128(w1s + (uint32_t)(k * Cfg::WCH_BYTES));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma2_ss(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                            commit_pair(&a_empty[k]);
                        }
                        commit_pair(&d_full[0]);
                        DP_TIMED(tw2, mbar_wait_cl(&a2_full, ph));                     // v^2 (bf16) in TMEM of both CTAs
                        tcgen05_fence_after();
                        for (int k = 0; k < NCH; ++k) {
                            const uint64_t bdesc = umma_desc_sw128(w2s + (uint32_t)(k * Cfg::WCH_BYTES));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)              // 16 bf16 of K = 8 TMEM columns per step
                                umma2_ts(tmem_base + (uint32_t)Cfg::TM_D2, tmem_base + (uint32_t)(Cfg::TM_A2 + (k * 4 + kk) * 8), bdesc + (uint64_t)(kk * 2), idesc,
                                         1u);                 // onto beta, which phase 1 of the epilogue has written into the accumulator
                        }
                        commit_pair(&d_full[1]);
                    }
                }
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = (unsigned long long)(clock64() - t_start); dbg[1] = (unsigned long long)tw0; dbg[2] = (unsigned long long)tw1; dbg[3] = (unsigned long long)tw2; }
        }
      }
    } else if (warp >= Cfg::PW0) {
        // ---- depthwise 3x3 producers: halo patch -> A stage (bf16, K-major, 128B swizzle).  Warp pw: columns 2 cp, 2 cp + 1 and
        // output rows [row0, row0 + DP_PROD_ROWS) of the 8 x 16 tile; lane = channel pair
        if constexpr (Cfg::SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MLIC_DP_REGS_PROD));
        long long tw0 = 0, tw1 = 0;
        const int pw = warp - Cfg::PW0;
        const int cp = pw & 7, row0 = (pw >> 3) * DP_PROD_ROWS;
        const uint32_t a_full_leader0 = mapa_u32(smem_u32(&a_full[0]), 0);
        int stage = 0, rslot = 0;
        uint32_t sphase = 0, rphase = 0;
        // STS offsets of this lane inside an A stage: row r = oy * 16 + 2 cp + c, (r & 7) = (2 cp + c) & 7 whatever oy
        uint32_t st_off[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int r0 = 2 * cp + c;
            st_off[c] = (uint32_t)(row0 * 16 * 128 + r0 * 128 + (((lane >> 2) ^ (r0 & 7)) << 4) + ((lane & 3) << 2));
        }
        for (int pp = pair_first; pp < p.npairs; pp += pair_step) {
            for (int k = 0; k < NCH; ++k) {
                DP_TIMED(tw0, mbar_wait(&raw_full[rslot], rphase));
                DP_TIMED(tw1, mbar_wait(&a_empty[stage], sphase ^ 1));
                uint32_t rp = smem_u32(base + Cfg::OFF_RAW) + (uint32_t)(rslot * DP_RAW_BYTES + (row0 * (DP_TW + 2) + 2 * cp) * 128 + lane * 4);
                uint32_t wk = smem_u32(sDw) + (uint32_t)((k * 640 + 2 * lane) * 4);
                asm volatile("" : "+r"(rp), "+r"(wk) :: "memory");        // the loads below depend on rp / wk: they stay behind the waits
                const uint32_t sa = smem_u32(base + Cfg::OFF_A) + (uint32_t)(stage * DP_A_BYTES);
                float2 w2[9];
#pragma unroll
                for (int tp = 0; tp < 9; ++tp) w2[tp] = lds_f2_nv(wk + (uint32_t)(tp * 256));
                const float2 b2 = lds_f2_nv(wk + 9u * 256u);
                float2 acc[DP_PROD_ROWS][2];
#pragma unroll
                for (int oy = 0; oy < DP_PROD_ROWS; ++oy) { acc[oy][0] = b2; acc[oy][1] = b2; }
#pragma unroll
                for (int iy = 0; iy < DP_PROD_ROWS + 2; ++iy) {
                    float2 x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = bf2_unpack_alu(lds32_nv(rp + (uint32_t)((iy * (DP_TW + 2) + j) * 128)));
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int oy = iy - ky;
#ifdef MLIC_DP_SKIP_FMA          // development: timing attribution only (wrong results): one tap instead of nine
                        if (ky != 1) continue;
#endif
                        if (oy >= 0 && oy < DP_PROD_ROWS) {
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
                            __nv_bfloat162 hv = __floats2bfloat162_rn(acc[oy][c].x, acc[oy][c].y);
                            sts32(sa + st_off[c] + (uint32_t)(oy * 16 * 128), *reinterpret_cast<uint32_t*>(&hv));
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> UMMA (async proxy) reads, possibly issued by the peer
                __syncwarp();
                if (lane == 0) { mbar_arrive_cluster(a_full_leader0 + (uint32_t)(stage * 8)); mbar_arrive(&raw_empty[rslot]); }
                if (++rslot == Cfg::RAW_SLOTS) { rslot = 0; rphase ^= 1; }
                if (++stage == Cfg::NA) { stage = 0; sphase ^= 1; }
            }
        }
        if (dbg && blockIdx.x == 0 && pw == 0 && lane == 0) { dbg[4] = (unsigned long long)(clock64() - t_start); dbg[5] = (unsigned long long)tw0; dbg[6] = (unsigned long long)tw1; }
    } else {
        // ---- epilogue: group eb = 64 output columns, q = TMEM lane quarter of this warp, one pixel per thread
        if constexpr (Cfg::SETREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MLIC_DP_REGS_EPI));
        long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
        const int q = warp & 3, eb = (warp - Cfg::EW0) >> 2;
        const int r = q * 32 + lane;
        const bool gissuer = (q == 0 && lane == 0);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t d_empty_leader0 = mapa_u32(smem_u32(&d_empty[0]), 0);
        const uint32_t a2_full_leader = mapa_u32(smem_u32(&a2_full), 0);
        (void)a2_full_leader;
        const uint32_t sB1s = smem_u32(sB1) + (uint32_t)(eb * 64 * 4), sB2s = smem_u32(sB2) + (uint32_t)(eb * 64 * 4);
        int it = 0;
        for (int pp = pair_first; pp < p.npairs; pp += pair_step, ++it) {
            DP_TILE(pp);
            const int slot = Cfg::NRING == 1 ? 0 : (it & 1);
            uint8_t* sb = base + Cfg::OFF_STG + (size_t)(eb * Cfg::NRING + slot) * DP_STG_BYTES;
            const uint32_t sb_s = smem_u32(sb);
            uint64_t* sbar = &stg_bar[eb][slot];
            const uint32_t sbar_ph = Cfg::NRING == 1 ? ((uint32_t)it & 1u) : ((uint32_t)(it >> 1) & 1u);
            // RES with two slots: the residual block of tile it + 1 is fetched into the other slot while this tile is computed (issued
            // below, after the first accumulator read), so neither the drain of the previous store nor the latency of the load is
            // on the epilogue's path; the first tile's block is fetched here
            constexpr bool RESPF = RES && Cfg::NRING == 2;
            if constexpr (RESPF) {
                if (it == 0 && gissuer) {
                    mbar_expect_tx(sbar, (uint32_t)DP_STG_BYTES);
                    tma_load_4d(sb, &tm.res, sbar, eb * 64, w0, h0, img);
                }
            } else {
                // the slot's previous TMA store (NRING tiles ago) must have finished reading it
                DP_TIMED(tw3, { if (gissuer) tma_store_wait_read(Cfg::NRING - 1); asm volatile("bar.sync %0, 128;" ::"r"(eb + 1) : "memory"); });
                if constexpr (RES) {
                    if (gissuer) {
                        mbar_expect_tx(sbar, (uint32_t)DP_STG_BYTES);
                        tma_load_4d(sb, &tm.res, sbar, eb * 64, w0, h0, img);
                    }
                }
            }
            if constexpr (MODE == DP_DS) {
                const int acc = it & 1;
                DP_TIMED(tw0, mbar_wait(&d_full[acc], (uint32_t)(it >> 1) & 1u));
                tcgen05_fence_after();
                if constexpr (RES) DP_TIMED(tw1, mbar_wait(sbar, sbar_ph));
                const uint32_t trow = lane_base + (uint32_t)(acc * Cfg::TM_D2 + eb * 64);
#pragma unroll 1
                for (int pr = 0; pr < 4; ++pr) {
                    uint32_t raw[16];
                    tmem_ld16(trow + (uint32_t)(pr * 16), raw);
                    tmem_ld_wait();
                    if constexpr (RESPF) {
                        if (pr == 0 && gissuer && pp + pair_step < p.npairs) {
                            int tn = 2 * (pp + pair_step) + (int)rank;
                            if (tn >= p.ntiles) tn = p.ntiles - 1;
                            const int imgn = tn / tiles_per_img, tremn = tn - imgn * tiles_per_img;
                            const int thn = tremn / p.tilesW, twn = tremn - thn * p.tilesW;
                            tma_store_wait_read(0);             // the other slot's store (previous tile) has read its block
                            mbar_expect_tx(&stg_bar[eb][slot ^ 1], (uint32_t)DP_STG_BYTES);
                            tma_load_4d(base + Cfg::OFF_STG + (size_t)(eb * Cfg::NRING + (slot ^ 1)) * DP_STG_BYTES, &tm.res, &stg_bar[eb][slot ^ 1], eb * 64,
                                        twn * DP_TW, thn * DP_TH, imgn);
                        }
                    }
                    if (pr == 3) {                  // accumulator fully read by this warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(d_empty_leader0 + (uint32_t)(acc * 8));
                    }
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int jj = pr * 2 + sub;
                        const uint32_t off = (uint32_t)(r * 128 + ((jj ^ (r & 7)) << 4));
                        const uint4 bau = lds128(sB1s + (uint32_t)(jj * 32)), bbu = lds128(sB1s + (uint32_t)(jj * 32 + 16));
                        const float4 ba = make_float4(__uint_as_float(bau.x), __uint_as_float(bau.y), __uint_as_float(bau.z), __uint_as_float(bau.w));
                        const float4 bb = make_float4(__uint_as_float(bbu.x), __uint_as_float(bbu.y), __uint_as_float(bbu.z), __uint_as_float(bbu.w));
                        float2 v[4];
                        v[0] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 0]), __uint_as_float(raw[sub * 8 + 1])), make_float2(ba.x, ba.y));
                        v[1] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 2]), __uint_as_float(raw[sub * 8 + 3])), make_float2(ba.z, ba.w));
                        v[2] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 4]), __uint_as_float(raw[sub * 8 + 5])), make_float2(bb.x, bb.y));
                        v[3] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 6]), __uint_as_float(raw[sub * 8 + 7])), make_float2(bb.z, bb.w));
#ifndef MLIC_DP_SKIP_GELU        // development: timing attribution only (wrong results)
                        if constexpr (ACT == ACT_GELU) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = HALF ? gelu2_half(v[j]) : gelu2(v[j]);
                        }
#endif
                        if constexpr (RES) {
                            const uint4 t = lds128(sb_s + off);
                            v[0] = __fadd2_rn(v[0], bf2_to_f2(t.x)); v[1] = __fadd2_rn(v[1], bf2_to_f2(t.y));
                            v[2] = __fadd2_rn(v[2], bf2_to_f2(t.z)); v[3] = __fadd2_rn(v[3], bf2_to_f2(t.w));
                        }
                        uint4 o;
                        { __nv_bfloat162 h0b = __floats2bfloat162_rn(v[0].x, v[0].y), h1b = __floats2bfloat162_rn(v[1].x, v[1].y),
                                         h2b = __floats2bfloat162_rn(v[2].x, v[2].y), h3b = __floats2bfloat162_rn(v[3].x, v[3].y);
                          o.x = *reinterpret_cast<uint32_t*>(&h0b); o.y = *reinterpret_cast<uint32_t*>(&h1b);
                          o.z = *reinterpret_cast<uint32_t*>(&h2b); o.w = *reinterpret_cast<uint32_t*>(&h3b); }
                        sts128(sb_s + off, o);
                    }
                }
            } else {
                // ---- phase 1: v = acc1 + b1, written back in place; v^2 -> bf16 A operand of the GDN GEMM, in TMEM; the GDN accumulator is
                // preloaded with beta (the second MMA accumulates onto it).  Phase 2 is the serial stretch of the tile (MMA 1 -> phase 1 -> MMA 2
                // -> phase 2, one accumulator set): both biases are paid here, 6 instructions per 16 values, instead of 16 there.
                const uint32_t ph = (uint32_t)it & 1u;
                DP_TIMED(tw0, mbar_wait(&d_full[0], ph));
                tcgen05_fence_after();
                const uint32_t trow1 = lane_base + (uint32_t)(eb * 64);
                const uint32_t trowa = lane_base + (uint32_t)(Cfg::TM_A2 + eb * 32);
                const uint32_t trow2 = lane_base + (uint32_t)(Cfg::TM_D2 + eb * 64);
#pragma unroll 1
                for (int pr = 0; pr < 4; ++pr) {
                    uint32_t raw[16], sq[8];
                    tmem_ld16(trow1 + (uint32_t)(pr * 16), raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 b = lds_f2(sB1s + (uint32_t)((pr * 16 + 2 * j) * 4));
                        float2 v = __fadd2_rn(make_float2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1])), b);
                        raw[2 * j] = __float_as_uint(v.x); raw[2 * j + 1] = __float_as_uint(v.y);
                        v = __fmul2_rn(v, v);
                        __nv_bfloat162 hv = __floats2bfloat162_rn(v.x, v.y);
                        sq[j] = *reinterpret_cast<uint32_t*>(&hv);
                    }
                    tmem_st8(trowa + (uint32_t)(pr * 8), sq);
                    tmem_st16(trow1 + (uint32_t)(pr * 16), raw);
                    uint32_t be[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 t = lds128(sB2s + (uint32_t)((pr * 16 + 4 * j) * 4));
                        be[4 * j] = t.x; be[4 * j + 1] = t.y; be[4 * j + 2] = t.z; be[4 * j + 3] = t.w;
                    }
                    tmem_st16(trow2 + (uint32_t)(pr * 16), be);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(a2_full_leader);
                // ---- phase 2: out = v * (r)sqrt(gamma v^2 + beta) + res
                DP_TIMED(tw2, mbar_wait(&d_full[1], ph));
                tcgen05_fence_after();
                if constexpr (RES) DP_TIMED(tw1, mbar_wait(sbar, sbar_ph));
#pragma unroll 1
                for (int pr = 0; pr < 4; ++pr) {
                    uint32_t rv[16], rn[16];
                    tmem_ld16(trow1 + (uint32_t)(pr * 16), rv);
                    tmem_ld16(trow2 + (uint32_t)(pr * 16), rn);
                    tmem_ld_wait();
                    if (pr == 3) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(d_empty_leader0);
                    }
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int jj = pr * 2 + sub;
                        const uint32_t off = (uint32_t)(r * 128 + ((jj ^ (r & 7)) << 4));
                        uint32_t ow[4];
                        uint4 t = make_uint4(0u, 0u, 0u, 0u);
                        if constexpr (RES) t = lds128(sb_s + off);
                        const uint32_t tw4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 v = make_float2(__uint_as_float(rv[sub * 8 + 2 * j]), __uint_as_float(rv[sub * 8 + 2 * j + 1]));
                            float s0, s1;
                            if constexpr (GDN == GDN_FWD) {
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "r"(rn[sub * 8 + 2 * j]));
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "r"(rn[sub * 8 + 2 * j + 1]));
                            } else {
                                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "r"(rn[sub * 8 + 2 * j]));
                                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "r"(rn[sub * 8 + 2 * j + 1]));
                            }
                            float2 o = __fmul2_rn(v, make_float2(s0, s1));
                            if constexpr (RES) o = __fadd2_rn(o, bf2_to_f2(tw4[j]));
                            __nv_bfloat162 hv = __floats2bfloat162_rn(o.x, o.y);
                            ow[j] = *reinterpret_cast<uint32_t*>(&hv);
                        }
                        sts128(sb_s + off, make_uint4(ow[0], ow[1], ow[2], ow[3]));
                    }
                }
            }
            DP_TIMED(tw3, { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); asm volatile("bar.sync %0, 128;" ::"r"(eb + 1) : "memory"); });
            if (gissuer) {
                if (tvalid) tma_store_4d(&tm.out, sb, eb * 64, w0, h0, img);
                tma_store_commit();
            }
        }
        if (gissuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (dbg && blockIdx.x == 0 && warp == Cfg::EW0 && lane == 0) {
            dbg[8] = (unsigned long long)(clock64() - t_start); dbg[9] = (unsigned long long)tw0; dbg[10] = (unsigned long long)tw1; dbg[11] = (unsigned long long)tw2; dbg[12] = (unsigned long long)tw3;
        }
    }
#undef DP_TILE
#undef DP_TIMED
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still be reading this CTA's weights / TMEM through the pair MMA
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled_dp)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local char g_dp_err[256] = "";
const char* ds_pair_last_error() { return g_dp_err; }

bool ds_pair_supported(const DsPairArgs& a) {
    if (!a.in || !a.out || !a.w1 || !a.b1 || !a.dw_w9 || !a.dw_bias) return false;
    if (a.C != 192 && a.C != 128) return false;
    if (a.gdn != GDN_NONE && (!a.w2 || !a.b2 || a.act != ACT_NONE)) return false;
    if (a.gdn == GDN_NONE && a.act != ACT_NONE && a.act != ACT_GELU) return false;
    if (a.B <= 0 || a.H <= 0 || a.W <= 0) return false;
    auto ok16 = [](const void* q, int ld) { return q == nullptr || ((((uintptr_t)q) % 16 == 0) && ((ld * 2) % 16 == 0)); };
    if (!ok16(a.in, a.ld) || !ok16(a.out, a.out_ld) || !ok16(a.res, a.res_ld)) return false;
    if (((uintptr_t)a.w1 % 16) != 0 || ((uintptr_t)a.w2 % 16) != 0) return false;
    return true;
}

template <int NCH, int MODE, int ACT, bool RES, int GDN>
static int dp_launch(const DpMaps& tm, const DpParams& p, int nclusters, cudaStream_t s) {
    using Cfg = DpCfg<NCH, MODE, RES>;
    auto fn = ds_pair_kernel<NCH, MODE, ACT, RES, GDN>;
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) {
        cudaError_t er = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (er != cudaSuccess) { snprintf(g_dp_err, sizeof g_dp_err, "cudaFuncSetAttribute(ds_pair, %d B): %s", Cfg::SMEM, cudaGetErrorString(er)); return 4; }
        attr[dev] = true;
    }
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3((unsigned)Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 2 : 1;
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
        static int printed = 0;
        unsigned long long h[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        if (printed++ < 4) {
            const double tiles = (double)((p.npairs + nclusters - 1) / nclusters);
            fprintf(stderr, "[pair dbg mode=%d res=%d] tiles/cta %.0f | per tile: total %.0f | mma: wait-acc-free %.0f wait-A %.0f wait-A2 %.0f | producer w4: wait-raw %.0f wait-stage-free %.0f"
                            " | epilogue w0: wait-acc %.0f wait-res %.0f wait-acc2 %.0f barriers+store-drain %.0f\n", MODE, (int)RES, tiles, h[0] / tiles,
                    h[1] / tiles, h[2] / tiles, h[3] / tiles, h[5] / tiles, h[6] / tiles, h[9] / tiles, h[10] / tiles, h[11] / tiles, h[12] / tiles);
        }
    }
    if (er != cudaSuccess) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair launch: %s (smem %d)", cudaGetErrorString(er), Cfg::SMEM); return 5; }
    return 0;
}

template <int NCH>
static int dp_dispatch(const DsPairArgs& a, const DpMaps& tm, const DpParams& p, int nclusters, cudaStream_t s) {
    const bool res = a.res != nullptr;
    if (a.gdn == GDN_FWD) return res ? dp_launch<NCH, DP_TAIL, ACT_NONE, true, GDN_FWD>(tm, p, nclusters, s) : dp_launch<NCH, DP_TAIL, ACT_NONE, false, GDN_FWD>(tm, p, nclusters, s);
    if (a.gdn == GDN_INV) return res ? dp_launch<NCH, DP_TAIL, ACT_NONE, true, GDN_INV>(tm, p, nclusters, s) : dp_launch<NCH, DP_TAIL, ACT_NONE, false, GDN_INV>(tm, p, nclusters, s);
    if (a.act == ACT_GELU) return res ? dp_launch<NCH, DP_DS, ACT_GELU, true, GDN_NONE>(tm, p, nclusters, s) : dp_launch<NCH, DP_DS, ACT_GELU, false, GDN_NONE>(tm, p, nclusters, s);
    return res ? dp_launch<NCH, DP_DS, ACT_NONE, true, GDN_NONE>(tm, p, nclusters, s) : dp_launch<NCH, DP_DS, ACT_NONE, false, GDN_NONE>(tm, p, nclusters, s);
}

int launch_ds_pair(const DsPairArgs& a, cudaStream_t s) {
    if (tc_init()) { snprintf(g_dp_err, sizeof g_dp_err, "%s", tc_last_error()); return 1; }
    if (!ds_pair_supported(a)) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair: unsupported layer"); return 1; }
    PFN_encodeTiled_dp enc = (PFN_encodeTiled_dp)tc_encode_fn();
    DpMaps tm;
    memset(&tm, 0, sizeof tm);
    DpParams p;
    memset(&p, 0, sizeof p);
    p.Hout = a.H; p.Wout = a.W;
    p.tilesH = (a.H + DP_TH - 1) / DP_TH; p.tilesW = (a.W + DP_TW - 1) / DP_TW;
    const long long nt = (long long)a.B * p.tilesH * p.tilesW;
    if (nt <= 0 || nt > 0x3fffffffLL) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair: tile count out of range"); return 3; }
    p.ntiles = (int)nt; p.npairs = (p.ntiles + 1) / 2;
    p.dw_w9 = a.dw_w9; p.dw_bias = a.dw_bias; p.b1 = a.b1; p.b2 = a.b2;
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.ld * 2, (cuuint64_t)a.W * a.ld * 2, (cuuint64_t)a.H * a.W * a.ld * 2};
        cuuint32_t box[4] = {64, DP_TW + 2, DP_TH + 2, 1};
        CUresult r = enc(&tm.raw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair: encode(raw) failed: %d", (int)r); return 2; }
    }
    for (int m = 0; m < 2; ++m) {
        const void* w = m ? a.w2 : a.w1;
        if (!w) continue;
        cuuint64_t dims[2] = {(cuuint64_t)a.C, (cuuint64_t)a.C};
        cuuint64_t strides[1] = {(cuuint64_t)a.C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(a.C / 2)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(m ? &tm.w2 : &tm.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair: encode(w%d) failed: %d", m + 1, (int)r); return 2; }
    }
    for (int m = 0; m < 2; ++m) {
        const void* q = m ? a.res : a.out;
        const int ld = m ? a.res_ld : a.out_ld;
        if (!q) continue;
        cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)a.W * ld * 2, (cuuint64_t)a.H * a.W * ld * 2};
        cuuint32_t box[4] = {64, DP_TW, DP_TH, 1};
        CUresult r = enc(m ? &tm.res : &tm.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_dp_err, sizeof g_dp_err, "ds_pair: encode(%s) failed: %d", m ? "res" : "out", (int)r); return 2; }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 2) sms = 148;
    const int nclusters = p.npairs < sms / 2 ? p.npairs : sms / 2;
    if (a.C == 192) return dp_dispatch<3>(a, tm, p, nclusters, s);
    return dp_dispatch<2>(a, tm, p, nclusters, s);
}

}  // namespace mlic
