// Shared device-side vocabulary of the MLIC++ engine: NHWC activation views, the GEMM
// epilogue description, exact-erf GELU, checkerboard parity, typed vector load/store.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mlic {

typedef __nv_bfloat16 bf16;

// NHWC view: p points at channel 0 of the view, ld = elements between consecutive pixels.
struct Act {
    void* p;
    int B, H, W, C, ld;
};

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_HALF_TANH = 2 };
enum { PAR_NONE = 0, PAR_ANCHOR = 1, PAR_NONANCHOR = 2 };
enum { GDN_NONE = 0, GDN_FWD = 1, GDN_INV = 2 };

// Programmatic dependent launch: every kernel starts with pdl_wait() -- it blocks until the preceding grids of the stream have
// completed and flushed (a no-op for a plain launch), then lets the next grid of the stream be scheduled so that its own launch
// latency and prologue overlap this grid's execution (that grid again waits here for this one to finish).
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// anchor = (row + col) odd  (reference: MLIC++/utils/ckbd.py:35-45)
__device__ __forceinline__ bool parity_keep(int par, int h, int w) {
    return par == PAR_NONE || (((h + w) & 1) == (par == PAR_ANCHOR ? 1 : 0));
}
struct Epi;
__device__ __forceinline__ int premask_of(const Epi& e, int n);

// What happens to one accumulator row segment on its way to memory.  Order:
//   v = premask ? (keep ? acc : 0) : acc;  v += bias;  gdn: v = x * (r)sqrt(v);  act;  postmask;  v += res;
//   store out (optionally pixel-shuffled r=2);  out2 = v*v (optional, same addressing as out).
struct Epi {
    const float* bias;      // [N] in GEMM column order, may be null
    int act, premask, postmask;
    int pm_w, pm_codes;     // pm_w != 0: the premask of column n is (pm_codes >> 2 (n / pm_w)) & 3 instead of `premask` (one GEMM for the
                            // q | k | v projections of LinearGlobalIntraContext, whose inputs carry different parity masks); pm_w % 8 == 0
    const void* res; int res_ld;       // residual, addressed like the OUTPUT
    int gdn; const void* gdn_x; int gdn_ld;   // x operand of (I)GDN, addressed like the GEMM-space pixel
    int shuffle;            // 0 | 1: PixelShuffle(2); GEMM column n' = (2r+s)*Cq + c  (weights permuted at pack time)
    void* out; int out_ld;
    int out_f32;            // 1: `out` is float* even when activations are bf16 (y, entropy parameters)
    int nchw;               // 1 (with out_f32): `out` is a dense fp32 NCHW tensor [B][C][OH][OW] (x_hat); scalar stores
    void* out2; int out2_ld;
    int Hout, Wout, N;      // GEMM-space output grid and column count
};
__device__ __forceinline__ int premask_of(const Epi& e, int n) { return e.pm_w ? ((e.pm_codes >> (2 * (n / e.pm_w))) & 3) : e.premask; }

__device__ __forceinline__ float gelu_erf(float x) {          // nn.GELU(approximate='none')
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// Fast GELU of the bf16 path (packed fp32x2; see the epilogue notes in gemm_tc.cu).
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifndef MLIC_GELU_FORM
#define MLIC_GELU_FORM 0
#endif
__device__ __forceinline__ float2 gelu2(float2 x) {
    float2 t = __fmul2_rn(x, x);
    t.x = fminf(t.x, 64.0f); t.y = fminf(t.y, 64.0f);
#if MLIC_GELU_FORM == 0      // one MUFU.TANH (2^-11 relative on tanh)
    float2 q = __ffma2_rn(t, make_float2(-3.51516790e-04f, -3.51516790e-04f), make_float2(3.70056460e-02f, 3.70056460e-02f));
    q = __ffma2_rn(q, t, make_float2(7.97507884e-01f, 7.97507884e-01f));
    const float2 u = __fmul2_rn(x, q);
    const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, th, hx);
#else                        // Phi = 1 / (1 + 2^(-2 log2(e) u)): MUFU.EX2 + MUFU.RCP, both ~2^-22
    constexpr float K = -2.0f * 1.4426950408889634f;
    float2 q = __ffma2_rn(t, make_float2(K * -3.51516790e-04f, K * -3.51516790e-04f), make_float2(K * 3.70056460e-02f, K * 3.70056460e-02f));
    q = __ffma2_rn(q, t, make_float2(K * 7.97507884e-01f, K * 7.97507884e-01f));
    const float2 u = __fmul2_rn(x, q);
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(u.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(u.y));
    const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
    return __fmul2_rn(x, make_float2(r0, r1));
#endif
}
// GELU(2 h) from h = x / 2 (tanh form only): with the 1/2 folded into the producer's operands (exact: a power of two) the epilogue saves
// the multiply by 0.5 -- one of its 6 FP32x2 instructions per value pair, and the FP32x2 pipe (1.3-1.7 instructions per clock and SM,
// tools/ubench/pipes.cu) is what paces the fused DepthWiseConv kernels.  Bit-identical to gelu2(2 h): every rescaled quantity differs
// from the original by a power of two (s = t / 4, q' = 2 q with coefficients 32 c2, 8 c1, 2 c0, u = h q' = x q).
__device__ __forceinline__ float2 gelu2_half(float2 h) {
    float2 s = __fmul2_rn(h, h);
    s.x = fminf(s.x, 16.0f); s.y = fminf(s.y, 16.0f);
    float2 q = __ffma2_rn(s, make_float2(32.0f * -3.51516790e-04f, 32.0f * -3.51516790e-04f), make_float2(8.0f * 3.70056460e-02f, 8.0f * 3.70056460e-02f));
    q = __ffma2_rn(q, s, make_float2(2.0f * 7.97507884e-01f, 2.0f * 7.97507884e-01f));
    const float2 u = __fmul2_rn(h, q);
    const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
    return __ffma2_rn(h, th, h);
}
__device__ __forceinline__ float gelu_fast(float x) { return gelu2(make_float2(x, x)).x; }
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }


template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector access (16 B for float, 8 B for bf16); pointers must be so aligned.
__device__ __forceinline__ void load4(const float* p, float v[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float v[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
}

// Apply the epilogue to 4 consecutive GEMM columns n..n+3 of GEMM-space pixel (b,h,w) and store.
// `vec` = host-verified that every pointer/ld involved allows 4-wide vector access.
template <typename T>
__device__ __forceinline__ void epi_store4(const Epi& e, int b, int h, int w, int n, float v[4], bool vec) {
    if (n >= e.N) return;
    const bool keep_pre = parity_keep(premask_of(e, n), h, w);      // (n is a multiple of 4, groups of pm_w % 8 == 0 columns)
    const bool keep_post = parity_keep(e.postmask, h, w);
    int oh = h, ow = w, oc = n, OH = e.Hout, OW = e.Wout;
    bool straddle = false;
    if (e.shuffle) {
        const int Cq = e.N >> 2;
        const int g = n / Cq;
        oc = n - g * Cq;
        oh = 2 * h + (g >> 1);
        ow = 2 * w + (g & 1);
        OH *= 2; OW *= 2;
        straddle = (Cq & 3) != 0;       // the 4 columns may cross a shuffle group (final 3-channel conv)
    }
    const size_t gpix = ((size_t)b * e.Hout + h) * e.Wout + w;
    float x4[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.gdn) {
        const T* xp = reinterpret_cast<const T*>(e.gdn_x) + gpix * e.gdn_ld + n;
        if (vec && n + 3 < e.N) load4(xp, x4);
        else for (int j = 0; j < 4; ++j) if (n + j < e.N) x4[j] = to_f<T>(xp[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = keep_pre ? v[j] : 0.0f;
        if (e.bias && n + j < e.N) a += e.bias[n + j];
        if (e.gdn == GDN_FWD) a = x4[j] * rsqrtf(a);
        else if (e.gdn == GDN_INV) a = x4[j] * sqrtf(a);
        if (e.act == ACT_GELU) a = gelu_erf(a);
        else if (e.act == ACT_HALF_TANH) a = 0.5f * tanhf(a);
        if (!keep_post) a = 0.0f;
        v[j] = a;
    }
    if (!straddle && n + 3 < e.N && !e.nchw) {
        const size_t opix = ((size_t)b * OH + oh) * OW + ow;
        if (e.res) {
            const T* rp = reinterpret_cast<const T*>(e.res) + opix * e.res_ld + oc;
            float r4[4];
            if (vec) load4(rp, r4);
            else for (int j = 0; j < 4; ++j) r4[j] = to_f<T>(rp[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += r4[j];
        }
        if (e.out_f32) {
            float* op = reinterpret_cast<float*>(e.out) + opix * e.out_ld + oc;
            if (vec) store4(op, v);
            else for (int j = 0; j < 4; ++j) op[j] = v[j];
        } else {
            T* op = reinterpret_cast<T*>(e.out) + opix * e.out_ld + oc;
            if (vec) store4(op, v);
            else for (int j = 0; j < 4; ++j) op[j] = from_f<T>(v[j]);
        }
        if (e.out2) {
            float s4[4] = {v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3]};
            T* qp = reinterpret_cast<T*>(e.out2) + opix * e.out2_ld + oc;
            if (vec) store4(qp, s4);
            else for (int j = 0; j < 4; ++j) qp[j] = from_f<T>(s4[j]);
        }
    } else {
        // scalar tail: columns may cross a shuffle group or the end of N
        for (int j = 0; j < 4; ++j) {
            const int nj = n + j;
            if (nj >= e.N) break;
            int ohj = h, owj = w, ocj = nj;
            if (e.shuffle) {
                const int Cq = e.N >> 2;
                const int g = nj / Cq;
                ocj = nj - g * Cq;
                ohj = 2 * h + (g >> 1);
                owj = 2 * w + (g & 1);
            }
            const size_t opix = ((size_t)b * OH + ohj) * OW + owj;
            float a = v[j];
            if (e.res) a += to_f<T>(reinterpret_cast<const T*>(e.res)[opix * e.res_ld + ocj]);
            if (e.nchw) {
                const int Cc = e.shuffle ? (e.N >> 2) : e.N;
                reinterpret_cast<float*>(e.out)[(((size_t)b * Cc + ocj) * OH + ohj) * OW + owj] = a;
            } else if (e.out_f32) reinterpret_cast<float*>(e.out)[opix * e.out_ld + ocj] = a;
            else reinterpret_cast<T*>(e.out)[opix * e.out_ld + ocj] = from_f<T>(a);
            if (e.out2) reinterpret_cast<T*>(e.out2)[opix * e.out2_ld + ocj] = from_f<T>(a * a);
        }
    }
}

}  // namespace mlic
