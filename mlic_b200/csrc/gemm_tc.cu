// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
//   D[m][n] = sum_{tap, c} X[b, oh + ky - pad, ow + kx - pad, c] * Wt[n][tap * Cpad + c]
//
// One CTA computes a 128-pixel x BN-column output tile.  The 128 pixels are a TH x TW patch of the
// output grid (TH * TW = 128) so that the A operand of every (tap, 64-channel chunk) K step is ONE 4-D
// TMA box {64 ch, TW, TH, 1} of the NHWC activation: the conv halo / zero padding is the TMA's
// out-of-bounds zero fill, and there is no im2col buffer.  The box lands in shared memory as 128 rows of
// 128 B with the 128-byte swizzle, which is exactly the canonical K-major SWIZZLE_128B UMMA layout.
// B (weights, [N][taps * Cpad] bf16, K-major) is a 2-D TMA box {64, BN}.
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected
// lane, tcgen05.mma cta_group::1 kind::f16, M = 128, N = BN, K = 16), warp 2 = TMEM allocator,
// warps 4..7 = epilogue (tcgen05.ld 32x32b, one output pixel per thread, the shared `Epi` epilogue:
// bias / (I)GDN / GELU / tanh / parity mask / residual / pixel-shuffle store).
// Two CTAs are resident per SM (shared-memory budget below), so one CTA's epilogue overlaps the other's
// main loop.  Every mbarrier wait is bounded and traps instead of hanging.
#include "kernels.h"

#include <cuda.h>
#include <stdio.h>
#include <string.h>

namespace mlic {

static char g_tc_err[512] = "";
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

// ------------------------------------------------------------------------------------------ device PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: fail the launch instead of hanging the GPU
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
//   [0,14) start address >> 4 ; [16,30) leading byte offset >> 4 (unused for swizzled K-major, set 1) ;
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups ; [46,48) version = 1 ; [61,64) layout = 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcParams {
    int tilesH, tilesW;     // output patches per image
    int TH, TW;             // patch shape, TH * TW = 128
    int ks, pad;
    int kchunks;            // Cpad / 64
    int Cpad;
    int BN;                 // columns per CTA (multiple of 16, <= 256)
    int stages;
    int tmem_cols;          // power of two >= max(32, BN)
};

constexpr int TC_A_BYTES = 128 * 128;       // 128 rows x 64 bf16

__global__ void __launch_bounds__(256) conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, TcParams p,
                                                           Epi e, int vec) {
    extern __shared__ uint8_t smem_raw[];
    // carve: [barriers | pad to 1024] [stage0 A | stage0 B] ...
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.BN * 128;
    const int stage_bytes = TC_A_BYTES + b_bytes;      // multiple of 1024 (BN multiple of 16 -> 2048 B steps)
    __shared__ uint64_t full_bar[8], empty_bar[8], accum_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tilesH * p.tilesW;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
    const int h0 = th * p.TH, w0 = tw * p.TW;
    const int n0 = blockIdx.y * p.BN;
    const int ksteps = p.ks * p.ks * p.kchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int tap = k / p.kchunks, cc = k - tap * p.kchunks;
                const int ky = tap / p.ks, kx = tap - ky * p.ks;
                uint8_t* sa = base + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + TC_A_BYTES;
                mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                tma_load_4d(sa, &tmA, &full_bar[stage], cc * 64, w0 + kx - p.pad, h0 + ky - p.pad, img);
                tma_load_2d(sb, &tmB, &full_bar[stage], tap * p.Cpad + cc * 64, n0);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < ksteps; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(base + (size_t)stage * stage_bytes);
                const uint64_t adesc = umma_desc_sw128(sa);
                const uint64_t bdesc = umma_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)      // 4 x K=16 inside the 128-byte swizzle row: +32 B per step
                    umma_bf16(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                              (k | kk) ? 1u : 0u);
                tcgen05_commit(&empty_bar[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(&accum_bar);
        }
    } else if (warp >= 4) {
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int h = h0 + r / p.TW, w = w0 + r % p.TW;
        const bool valid = h < e.Hout && w < e.Wout;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < p.BN; c += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c, v);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) epi_store4<bf16>(e, img, h, w, n0 + c + j, v + j, vec != 0);
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
static int pick_bn(int N) {
    int nt = (N + 255) / 256;
    int bn = ((N + nt - 1) / nt + 15) / 16 * 16;
    return bn < 16 ? 16 : bn;
}

bool tc_conv_supported(const TcConv& c, const Epi& e) {
    if (!c.in || !c.w) return false;
    if (((uintptr_t)c.in) % 16 != 0 || (c.ld % 8) != 0) return false;      // TMA: 16-byte base / strides
    if (c.sW <= 0 || c.sH <= 0 || (c.sW % 8) != 0 || (c.sH % 8) != 0 || (c.sB % 8) != 0) return false;
    if (c.Cin < 8 || c.Cpad % 64 != 0) return false;
    if (e.N < 8) return false;
    if (c.H <= 0 || c.W <= 0 || c.B <= 0) return false;
    return true;
}

int launch_conv_gemm_tc(const TcConv& c, const Epi& e, int vec, cudaStream_t s) {
    if (tc_init()) return 1;
    TcParams p;
    // output patch shape: minimise the number of tiles
    const int cand[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
    long long best = -1;
    for (int i = 0; i < 5; ++i) {
        long long t = (long long)((e.Hout + cand[i][0] - 1) / cand[i][0]) * ((e.Wout + cand[i][1] - 1) / cand[i][1]);
        if (best < 0 || t < best) { best = t; p.TH = cand[i][0]; p.TW = cand[i][1]; }
    }
    p.tilesH = (e.Hout + p.TH - 1) / p.TH;
    p.tilesW = (e.Wout + p.TW - 1) / p.TW;
    p.ks = c.ks; p.pad = c.pad; p.Cpad = c.Cpad; p.kchunks = c.Cpad / 64;
    p.BN = pick_bn(e.N);
    const int stage_bytes = TC_A_BYTES + p.BN * 128;
    p.stages = (100 * 1024) / stage_bytes;
    if (p.stages < 2) p.stages = 2;
    if (p.stages > 6) p.stages = 6;
    const int ksteps = p.ks * p.ks * p.kchunks;
    if (p.stages > ksteps) p.stages = ksteps < 1 ? 1 : ksteps;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.Cin, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.sW * 2, (cuuint64_t)c.sH * 2, (cuuint64_t)c.sB * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(c.in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
        CUresult r = g_encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(c.w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(B) failed: %d (K=%llu N=%d BN=%d)", (int)r,
                     (unsigned long long)Ktot, e.N, p.BN);
            return 3;
        }
    }
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t er = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (er != cudaSuccess) {
            snprintf(g_tc_err, sizeof g_tc_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(er));
            return 4;
        }
        smem_set = 200 * 1024;
    }
    dim3 grid((unsigned)(c.B * p.tilesH * p.tilesW), (unsigned)((e.N + p.BN - 1) / p.BN));
    conv_gemm_tc_kernel<<<grid, 256, smem, s>>>(tmA, tmB, p, e, vec);
    cudaError_t er = cudaGetLastError();
    if (er != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "conv_gemm_tc launch: %s", cudaGetErrorString(er));
        return 5;
    }
    return 0;
}

}  // namespace mlic
