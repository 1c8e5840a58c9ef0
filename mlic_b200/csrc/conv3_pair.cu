// Two-SM (tcgen05 cta_group::2) dense 3x3 convolution, bf16 NHWC, for the wide sub-pixel convolutions of g_s
// (modules/layers/res_blk.py:107-121: `subpel_conv` and `upsample` of ResidualBlockUpsample, Conv2d(Cin -> 4 Cout, 3, pad 1) +
// PixelShuffle(2); 433 of the 876 GMAC of an MLICPP_L forward).
//
// Why a CTA PAIR.  With one CTA per 128 x 256 tile every K step of 64 channels needs 16 KB of activations AND 32 KB of
// weights in shared memory for 512 clocks of MMA: 94 B/clock per SM, where the L2 delivers ~45-50 (chip-wide ~6.3 KB/clock,
// B300_MICROARCH.md "LTS throughput cap"; measured in round 1 as the TMA thread busy ~900 clocks per K step).  With
// cta_group::2 the MMA is M = 256 (128 pixels of each CTA), N = 256, and each CTA stages only ITS 128 rows of the weight tile:
// 32 KB per K step and SM, 64 B/clock -- the weight traffic per FLOP is halved, the activation traffic unchanged.
//
// Per CTA, 640 threads: warp 0 TMA (implicit im2col: the A operand of K step (tap, 64-channel chunk) is the 4-D box
// {64, 16, 8, 1} of the NHWC input at (c0, w0 + kx - 1, h0 + ky - 1, b), zero-filled outside the image; B is the 2-D box
// {64, 128} of the packed weights [N][tap][Cpad]), both CTAs' loads complete on the LEADER's `full` barrier; warp 1 of the
// leader issues the MMAs for the pair and releases stages / accumulators with multicast commits; warp 2 owns TMEM (two
// accumulator stages of 256 columns: the epilogue of a tile overlaps the main loop of the next); 16 epilogue warps (one group of
// 4 per 64 output columns): TMEM -> bias (+ GELU) -> bf16 -> swizzled staging block -> TMA store through the tensor map of the
// column block's PixelShuffle group (the shuffle is an address pattern of the store).
#include "kernels.h"
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlic {

// (tile shape TH x TW = 128 pixels: CpParams, chosen per launch)
constexpr int CP_A_BYTES = 128 * 128, CP_B_BYTES = 128 * 128;
constexpr int CP_STAGE_BYTES = CP_A_BYTES + CP_B_BYTES;
constexpr int CP_NSTAGE = 5;
constexpr int CP_STG_BYTES = 128 * 128;
constexpr int CP_EPI_WARPS = 16;
constexpr int CP_THREADS = 128 + CP_EPI_WARPS * 32;
constexpr int CP_SMEM = CP_NSTAGE * CP_STAGE_BYTES + 4 * CP_STG_BYTES + 1024;
static_assert(CP_SMEM + 512 <= 232448, "shared-memory plan exceeds 227 KB");

struct CpParams {
    int tilesH, tilesW, ntiles, npairs;     // TH x TW pixel tiles (TH * TW = 128); a pair = tiles 2 pp, 2 pp + 1
    int TH, TW;
    int tilesN, nitems;                     // 256-column tiles; work items = npairs * tilesN (column tile fastest)
    int kchunks;                            // Cpad / 64
    int Cq;                                 // columns per output tensor map (N / 4 with PixelShuffle, else N)
    const float* bias;                      // [N]
    int ks, pad;                            // 3 / 1 (the sub-pixel convs) or 1 / 0 (wide 1x1 GEMMs of the entropy model)
    int BN, N;                              // columns per tile (multiple of 16, <= 256; each CTA stages BN / 2 weight rows) and in all
    int ck;                                 // 1 | 2: A rows are the anchor | non-anchor pixels of the input, squeezed (5-D map, gemm_tc.cu TcConv::ck)
};
struct CpMaps {
    CUtensorMap a, b;
    CUtensorMap o[4];
};

// .cta_group::2: the mbarrier that receives complete_tx may live in the peer CTA of the pair (here: always the leader's)
__device__ __forceinline__ void tma_load_4d_cl(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_cl(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_5d_cl(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

template <int ACT, bool RES>
__global__ void __launch_bounds__(CP_THREADS, 1)
conv3_pair_kernel(const __grid_constant__ CpMaps tm, const CpParams p, unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t cp_smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)cp_smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[CP_NSTAGE], empty[CP_NSTAGE];      // full: the leader's instance is the live one
    __shared__ uint64_t d_full[2], d_empty[2];                  // d_empty: the leader's instance is the live one
    __shared__ uint64_t stg_bar[4];                             // RES: the tile's current output block has landed in the staging block
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    uint8_t* stg = base + CP_NSTAGE * CP_STAGE_BYTES;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.o[0]) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < CP_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], (uint32_t)(2 * 4 * ((p.BN + 63) / 64))); }
        for (int s = 0; s < 4; ++s) mbar_init(&stg_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers are initialised before anything arrives on them remotely
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // programmatic dependent launch: everything above touches parameters only
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tiles_per_img = p.tilesH * p.tilesW;
    const int item_first = (int)(blockIdx.x >> 1), item_step = (int)(gridDim.x >> 1);
    const int ksteps = p.ks * p.ks * p.kchunks;
    // work item -> (pixel-tile pair, column tile); tile of this CTA = 2 pp + rank (an odd tile count leaves the last pair's second CTA
    // a duplicate of the last tile, computed but not stored)
#define CP_ITEM(item)                                                                         \
    const int pp = (item) / p.tilesN, nt = (item) - pp * p.tilesN;                            \
    int tix = 2 * pp + (int)rank;                                                             \
    const bool tvalid = tix < p.ntiles;                                                       \
    if (!tvalid) tix = p.ntiles - 1;                                                          \
    const int img = tix / tiles_per_img;                                                      \
    const int trem = tix - img * tiles_per_img;                                               \
    const int th = trem / p.tilesW, tw = trem - th * p.tilesW;                                \
    const int h0 = th * p.TH, w0 = tw * p.TW;                                                 \
    (void)tvalid; (void)h0; (void)w0; (void)img; (void)nt

    long long tw0 = 0, tw1 = 0;
    const long long t_start = dbg ? clock64() : 0;
#define CP_TIMED(acc, stmt) do { if (dbg) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t full_leader0 = mapa_u32(smem_u32(&full[0]), 0);
            const uint32_t sbase = smem_u32(base);
            int s = 0;
            uint32_t ph = 0;
            for (int item = item_first; item < p.nitems; item += item_step) {
                CP_ITEM(item);
                const int nrow = nt * p.BN + (int)rank * (p.BN >> 1);
                const uint32_t stage_tx = (uint32_t)(CP_A_BYTES + (p.BN >> 1) * 128);
                int ks = 0;
                for (int ky = 0; ky < p.ks; ++ky)
                    for (int kx = 0; kx < p.ks; ++kx)
                        for (int k = 0; k < p.kchunks; ++k, ++ks) {
                            CP_TIMED(tw0, mbar_wait(&empty[s], ph ^ 1));
                            if (leader) mbar_expect_tx(&full[s], 2u * stage_tx);      // the loads of BOTH CTAs land on this barrier
                            const uint32_t dst = sbase + (uint32_t)(s * CP_STAGE_BYTES);
                            if (p.ck) tma_load_5d_cl(dst, &tm.a, full_leader0 + (uint32_t)(s * 8), k * 64, w0, 0, h0 >> 1, img);
                            else tma_load_4d_cl(dst, &tm.a, full_leader0 + (uint32_t)(s * 8), k * 64, w0 + kx - p.pad, h0 + ky - p.pad, img);
                            tma_load_2d_cl(dst + CP_A_BYTES, &tm.b, full_leader0 + (uint32_t)(s * 8), ks * 64, nrow);
                            if (++s == CP_NSTAGE) { s = 0; ph ^= 1; }
                        }
            }
            if (dbg && blockIdx.x == 0) { dbg[0] = (unsigned long long)(clock64() - t_start); dbg[1] = (unsigned long long)tw0; }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // instruction descriptor: D = f32, A = B = bf16, K-major, N = BN, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((256u >> 4) << 24);
            const uint32_t sbase = smem_u32(base);
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int item = item_first; item < p.nitems; item += item_step, ++it) {
                const int acc = it & 1;
                CP_TIMED(tw0, mbar_wait_cl(&d_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u));
                tcgen05_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(acc * 256);
                for (int ks = 0; ks < ksteps; ++ks) {
                    CP_TIMED(tw1, mbar_wait_cl(&full[s], ph));
                    tcgen05_fence_after();
                    const uint64_t adesc = umma_desc_sw128(sbase + (uint32_t)(s * CP_STAGE_BYTES));
                    const uint64_t bdesc = umma_desc_sw128(sbase + (uint32_t)(s * CP_STAGE_BYTES + CP_A_BYTES));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma2_ss(dcol, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (ks | kk) ? 1u : 0u);
                    commit_pair(&empty[s]);
                    if (++s == CP_NSTAGE) { s = 0; ph ^= 1; }
                }
                commit_pair(&d_full[acc]);
            }
            if (dbg && blockIdx.x == 0) { dbg[2] = (unsigned long long)(clock64() - t_start); dbg[3] = (unsigned long long)tw0; dbg[4] = (unsigned long long)tw1; }
        }
    } else if (warp >= 4) {
        // ---- epilogue: group eb = 64 output columns of the tile, q = TMEM lane quarter of this warp, one pixel per thread
        const int q = warp & 3, eb = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const bool gissuer = (q == 0 && lane == 0);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t d_empty_leader0 = mapa_u32(smem_u32(&d_empty[0]), 0);
        uint8_t* sb = stg + (size_t)eb * CP_STG_BYTES;
        const uint32_t sb_s = smem_u32(sb);
        const int ncol = min(64, p.BN - eb * 64);                  // accumulator columns of this group (<= 0: the group owns none)
        int it = 0;
        for (int item = item_first; item < p.nitems && ncol > 0; item += item_step, ++it) {
            CP_ITEM(item);
            const int acc = it & 1;
            const int col0 = nt * p.BN + eb * 64;
            // the previous TMA store of this group has finished reading the staging block
            if (gissuer) tma_store_wait_read(0);
            asm volatile("bar.sync %0, 128;" ::"r"(eb + 1) : "memory");
            const int g = col0 / p.Cq, cc = col0 - g * p.Cq;            // PixelShuffle group of this column block and its offset inside it
            if constexpr (RES) {
                // out += conv(x): the block the tile will overwrite is fetched into the staging block while the main loop still runs
                if (gissuer) {
                    mbar_expect_tx(&stg_bar[eb], (uint32_t)CP_STG_BYTES);
                    tma_load_4d(sb, &tm.o[g], &stg_bar[eb], cc, w0, h0, img);
                }
            }
            CP_TIMED(tw0, mbar_wait(&d_full[acc], (uint32_t)(it >> 1) & 1u));
            tcgen05_fence_after();
            if constexpr (RES) mbar_wait(&stg_bar[eb], (uint32_t)it & 1u);
            const uint32_t trow = lane_base + (uint32_t)(acc * 256 + eb * 64);
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col0);
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
                    if (lane == 0) mbar_arrive_cluster(d_empty_leader0 + (uint32_t)(acc * 8));
                }
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int jj = pr * 2 + sub;
                    const uint32_t off = (uint32_t)(r * 128 + ((jj ^ (r & 7)) << 4));
                    const bool cin = col0 + jj * 8 + 8 <= p.N;                 // (a ragged last column tile: nothing beyond N is read or stored)
                    const float4 ba = cin ? __ldg(bias4 + jj * 2) : make_float4(0.f, 0.f, 0.f, 0.f), bb = cin ? __ldg(bias4 + jj * 2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float2 v[4];
                    v[0] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 0]), __uint_as_float(raw[sub * 8 + 1])), make_float2(ba.x, ba.y));
                    v[1] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 2]), __uint_as_float(raw[sub * 8 + 3])), make_float2(ba.z, ba.w));
                    v[2] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 4]), __uint_as_float(raw[sub * 8 + 5])), make_float2(bb.x, bb.y));
                    v[3] = __fadd2_rn(make_float2(__uint_as_float(raw[sub * 8 + 6]), __uint_as_float(raw[sub * 8 + 7])), make_float2(bb.z, bb.w));
                    if constexpr (ACT == ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = gelu2(v[j]);
                    }
                    if constexpr (RES) {
                        const uint4 t4 = lds128(sb_s + off);
                        v[0] = __fadd2_rn(v[0], bf2_to_f2(t4.x)); v[1] = __fadd2_rn(v[1], bf2_to_f2(t4.y));
                        v[2] = __fadd2_rn(v[2], bf2_to_f2(t4.z)); v[3] = __fadd2_rn(v[3], bf2_to_f2(t4.w));
                    }
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
                if (tvalid && col0 < p.N) tma_store_4d(&tm.o[g], sb, cc, w0, h0, img);
                tma_store_commit();
            }
        }
        if (gissuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { dbg[5] = (unsigned long long)(clock64() - t_start); dbg[6] = (unsigned long long)tw0; }
    }
#undef CP_ITEM
#undef CP_TIMED
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still be reading this CTA's operands / TMEM through the pair MMA
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled_cp)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local char g_cp_err[256] = "";
const char* conv3_pair_last_error() { return g_cp_err; }

static int cp_pick_bn(int N) {
    // one tile: N rounded up to 32 (16 weight rows per CTA; the last 64-column store block is clipped by the output map).  Several tiles:
    // a multiple of 64, so that every epilogue group stores a whole 64-column block of ITS tile, the one that wastes the fewest columns
    if (N <= 256) return (N + 31) / 32 * 32;
    int best = 256, waste = (N + 255) / 256 * 256 - N;
    for (int bn = 192; bn >= 128; bn -= 64) {
        const int w = (N + bn - 1) / bn * bn - N;
        if (w < waste) { waste = w; best = bn; }
    }
    return best;
}

bool conv3_pair_supported(const Conv3PairArgs& a) {
    if (!a.in || !a.w || !a.bias || !a.out) return false;
    if (a.B <= 0 || a.H <= 0 || a.W <= 0) return false;
    if (a.ks != 1 && a.ks != 3) return false;
    if (a.Cin < 8 || (a.Cin % 8) != 0 || a.Cpad < a.Cin || (a.Cpad % 64) != 0 || a.Cpad > 2048) return false;
    if (a.ks == 3 && (a.Cpad != a.Cin || (a.N % 256) != 0)) return false;
    if (a.N < 64 || (a.N % 8) != 0) return false;
    if (a.act != ACT_NONE && a.act != ACT_GELU) return false;
    if (a.shuffle && (a.ks != 3 || ((a.N / 4) % 64) != 0)) return false;
    if (a.res_inplace && a.ks != 3) return false;
    if (a.ck && (a.ks != 1 || (a.ck != 1 && a.ck != 2) || (a.H % 2) != 0 || (a.W % 2) != 0)) return false;
    if (((uintptr_t)a.in % 16) != 0 || (a.ld % 8) != 0 || ((uintptr_t)a.out % 16) != 0 || (a.out_ld % 8) != 0 || ((uintptr_t)a.w % 16) != 0) return false;
    if (((uintptr_t)a.bias % 16) != 0) return false;
    return true;
}

template <int ACT, bool RES>
static int cp_launch(const CpMaps& tm, const CpParams& p, int nclusters, cudaStream_t s) {
    auto fn = conv3_pair_kernel<ACT, RES>;
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) {
        cudaError_t er = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM);
        if (er != cudaSuccess) { snprintf(g_cp_err, sizeof g_cp_err, "cudaFuncSetAttribute(conv3_pair, %d B): %s", CP_SMEM, cudaGetErrorString(er)); return 4; }
        attr[dev] = true;
    }
    static const int pdl = getenv("MLIC_PDL") ? atoi(getenv("MLIC_PDL")) : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3((unsigned)CP_THREADS);
    cfg.dynamicSmemBytes = CP_SMEM;
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
            const double items = (double)((p.nitems + nclusters - 1) / nclusters);
            fprintf(stderr, "[conv3 pair dbg] items/pair %.0f ksteps %d | per item: tma total %.0f wait-empty %.0f | mma total %.0f wait-acc-free %.0f wait-full %.0f | epilogue w0 total %.0f wait-acc %.0f\n",
                    items, 9 * p.kchunks, h[0] / items, h[1] / items, h[2] / items, h[3] / items, h[4] / items, h[5] / items, h[6] / items);
        }
    }
    if (er != cudaSuccess) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair launch: %s (smem %d)", cudaGetErrorString(er), CP_SMEM); return 5; }
    return 0;
}

int launch_conv3_pair(const Conv3PairArgs& a, cudaStream_t s) {
    if (tc_init()) { snprintf(g_cp_err, sizeof g_cp_err, "%s", tc_last_error()); return 1; }
    if (!conv3_pair_supported(a)) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: unsupported layer"); return 1; }
    PFN_encodeTiled_cp enc = (PFN_encodeTiled_cp)tc_encode_fn();
    CpMaps tm;
    memset(&tm, 0, sizeof tm);
    CpParams p;
    memset(&p, 0, sizeof p);
    // GEMM rows: the output pixels; checkerboard mode: the anchor | non-anchor pixels of the input in squeezed order [B][H][W/2]
    const int Wout = a.ck ? a.W / 2 : a.W;
    {   // tile shape: the one with the fewest tiles (a flat [1, 1, M] matrix takes 1 x 128; checkerboard rows need an even TH)
        const int cand[4][2] = {{8, 16}, {4, 32}, {2, 64}, {1, 128}};
        long long best = -1;
        for (int i = 0; i < (a.ck ? 3 : 4); ++i) {
            const long long t = (long long)((a.H + cand[i][0] - 1) / cand[i][0]) * ((Wout + cand[i][1] - 1) / cand[i][1]);
            if (best < 0 || t < best) { best = t; p.TH = cand[i][0]; p.TW = cand[i][1]; }
        }
    }
    p.tilesH = (a.H + p.TH - 1) / p.TH; p.tilesW = (Wout + p.TW - 1) / p.TW;
    const long long nt = (long long)a.B * p.tilesH * p.tilesW;
    p.ks = a.ks; p.pad = a.ks / 2; p.N = a.N; p.ck = a.ck;
    p.BN = a.ks == 3 ? 256 : cp_pick_bn(a.N);
    p.tilesN = (a.N + p.BN - 1) / p.BN;
    if (nt <= 0 || nt * p.tilesN > 0x3fffffffLL) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: tile count out of range"); return 3; }
    p.ntiles = (int)nt; p.npairs = (p.ntiles + 1) / 2;
    p.nitems = p.npairs * p.tilesN;
    p.kchunks = a.Cpad / 64;
    p.Cq = a.shuffle ? a.N / 4 : a.N;
    p.bias = a.bias;
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    if (a.ck) {
        // anchor (ck 1): row h keeps w = 2j + 1 - (h & 1); non-anchor (ck 2): w = 2j + (h & 1)   (anchor = (h + w) odd); dims (c, j, h & 1, h >> 1, b)
        const size_t ld = (size_t)a.ld;
        const bf16* basep = reinterpret_cast<const bf16*>(a.in) + (a.ck == 1 ? ld : 0);
        const size_t hp_stride = a.ck == 1 ? ((size_t)a.W - 1) * ld : ((size_t)a.W + 1) * ld;
        cuuint64_t dims[5] = {(cuuint64_t)a.Cin, (cuuint64_t)(a.W / 2), 2, (cuuint64_t)(a.H / 2), (cuuint64_t)a.B};
        cuuint64_t strides[4] = {(cuuint64_t)(2 * ld) * 2, (cuuint64_t)hp_stride * 2, (cuuint64_t)(2 * (size_t)a.W * ld) * 2, (cuuint64_t)a.H * a.W * ld * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)p.TW, 2, (cuuint32_t)(p.TH / 2), 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(basep), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: encode(A, checkerboard) failed: %d", (int)r); return 2; }
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.ld * 2, (cuuint64_t)a.W * a.ld * 2, (cuuint64_t)a.H * a.W * a.ld * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        CUresult r = enc(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: encode(A) failed: %d", (int)r); return 2; }
    }
    {
        const cuuint64_t Ktot = (cuuint64_t)a.ks * a.ks * a.Cpad;
        cuuint64_t dims[2] = {Ktot, (cuuint64_t)a.N};
        cuuint64_t strides[1] = {Ktot * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / 2)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: encode(B) failed: %d", (int)r); return 2; }
    }
    {
        bf16* ob = reinterpret_cast<bf16*>(a.out);
        const size_t ld = (size_t)a.out_ld;
        const int ng = a.shuffle ? 4 : 1;
        for (int g = 0; g < ng; ++g) {          // group g = 2r + s -> output pixel (2h + r, 2w + s)
            const size_t OW = 2 * (size_t)a.W, OH = 2 * (size_t)a.H;
            bf16* bp = a.shuffle ? ob + ((size_t)(g >> 1) * OW + (size_t)(g & 1)) * ld : ob;
            cuuint64_t dims[4] = {(cuuint64_t)p.Cq, (cuuint64_t)Wout, (cuuint64_t)a.H, (cuuint64_t)a.B};
            cuuint64_t strides[3];
            if (a.shuffle) { strides[0] = 2 * ld * 2; strides[1] = 2 * OW * ld * 2; strides[2] = OH * OW * ld * 2; }
            else { strides[0] = ld * 2; strides[1] = (cuuint64_t)Wout * ld * 2; strides[2] = (cuuint64_t)a.H * Wout * ld * 2; }
            cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
            CUresult r = enc(&tm.o[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, bp, dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { snprintf(g_cp_err, sizeof g_cp_err, "conv3_pair: encode(out %d) failed: %d", g, (int)r); return 2; }
        }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 2) sms = 148;
    const int nclusters = p.nitems < sms / 2 ? p.nitems : sms / 2;
    if (a.res_inplace) return a.act == ACT_GELU ? cp_launch<ACT_GELU, true>(tm, p, nclusters, s) : cp_launch<ACT_NONE, true>(tm, p, nclusters, s);
    if (a.act == ACT_GELU) return cp_launch<ACT_GELU, false>(tm, p, nclusters, s);
    return cp_launch<ACT_NONE, false>(tm, p, nclusters, s);
}

}  // namespace mlic
