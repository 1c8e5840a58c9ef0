// Host-side launch API of the engine's kernels.  `bf` selects the activation element type
// (0 = fp32 validation mode, 1 = bf16 fast mode); arithmetic is fp32 everywhere.
#pragma once
#include "common.cuh"

namespace mlic {

// Launch of a CUDA-core kernel with programmatic stream serialization (MLIC_PDL=0: plain launches); the kernel's first statement is
// pdl_wait() (common.cuh).  ~190 of the 565 launches of an MLICPP_L forward are such kernels (depthwise convs, LayerNorm, the
// attention kernels, quantisation): without the attribute each of them pays the full launch gap behind a tcgen05 kernel.
bool pdl_enabled();
template <typename... KA, typename... A>
inline cudaError_t launch_k(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A... a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, static_cast<KA>(a)...);
}

struct ConvGeom {
    int B, H, W, Cin, ld;       // input view
    int Hout, Wout;
    int ks, stride, pad;
    int Ktot;                   // row length of the fp32 weight matrix [N][Ktot], Ktot = ks*ks*Cin
};

// ---- CUDA-core kernels (kernels.cu) ----
void launch_conv_gemm_simt(int bf, const void* in, const ConvGeom& g, const float* Wt, const Epi& e, int vec,
                           cudaStream_t s);
void launch_dwconv3x3(int bf, const Act& in, const Act& out, const float* w9, const float* bias, int stride, int act,
                      cudaStream_t s);
// g_a stage-0 head (bf16 mode): fp32 NCHW image -> GELU(pw(dw3x3_s2 x)) and skip1x1_s2(x), both bf16 NHWC
bool ga_head_supported(int H, int W, int N, const Act& t, const Act& sk);
void launch_ga_head(const float* x, int B, int H, int W, const float* dw9, const float* dwb, const float* w1, const float* b1,
                    const float* wsk, const float* bsk, int N, const Act& t, const Act& sk, cudaStream_t s);
void launch_nchw_to_nhwc(int bf, const float* src, const Act& dst, int Csrc, cudaStream_t s);
void launch_nhwc_to_nchw(int bf, const Act& src, float* dst, cudaStream_t s);
void launch_nhwc_f32_to_nchw(const float* src, int ld, int B, int H, int W, int C, float* dst, cudaStream_t s);
void launch_copy_channels(int bf, const Act& src, const Act& dst, cudaStream_t s);
void launch_copy_f32_to_act(int bf, const float* src, int ld, const Act& dst, cudaStream_t s);
void launch_fill_zero(void* p, size_t bytes, cudaStream_t s);

// EntropyBottleneck: z (NHWC act) -> z_hat (NHWC act), z_lik (fp32 NCHW), z_sym (int32 NCHW, optional)
void launch_entropy_bottleneck(int bf, const Act& z, const Act& z_hat, const float* packed /*[N][58]*/,
                               const float* medians, float* z_lik_nchw, int32_t* z_sym_nchw, cudaStream_t s, float qs = 1.0f);

void launch_zsym_to_zhat(int bf, const int32_t* z_sym_nchw, const float* medians, const Act& z_hat, cudaStream_t s, float qs = 1.0f);

// LocalContext windowed attention: F[pix][3C] fp32 (q|k|v) -> O[pix][25][C] (activation type); C = 32 or 64.
// returns non-zero for an unsupported C.
int launch_local_attn(int bf, const float* F, int B, int H, int W, int C, const float* rel_bias /*[2][25][25]*/,
                      void* O, cudaStream_t s);
// bf16 fast mode: the same attention on the warp-level tensor path for the NON-ANCHOR pixels only.
//   F: bf16 NHWC [B,H,W,96], channels head-major (q_h0 q_h1 k_h0 k_h1 v_h0 v_h1); O: bf16 [B*H*(W/2)][25*32] squeezed
//   (w = 2j + (h & 1)).  Non-zero: unsupported geometry.
int launch_local_attn_mma(const Act& F, const float* rel_bias, void* O, cudaStream_t s);
// squeezed non-anchor rows [B*H*(W/2)][C] -> NHWC view [B,H,W,C]; anchor pixels are written as zeros
void launch_unsqueeze_nonanchor(const Act& src, const Act& dst, cudaStream_t s);
void launch_layernorm(int bf, const Act& x, const float* g, const float* b, const Act& out, cudaStream_t s);

// Linear (kernelised) global attention: K softmax over positions, Q softmax over head channels.
//   qkv view: channels [0,D) = Q, [D,2D) = K, [2D,3D) = V.  par_kv / par_q: parity filters (intra context).
size_t lin_attn_scratch_floats(int B, int heads, int hd, int HW);
int launch_lin_attn(int bf, const Act& qkv, int D, int heads, int hd, int par_kv, int par_q, float* scratch,
                    const Act& out, cudaStream_t s);      // non-zero: unsupported head dim

// Fused quantise / likelihood / CDF-index kernels.
struct QuantArgs {
    const float* y; int y_ld;       // latent slice (fp32 NHWC view at channel i*C)
    const float* pa;                // anchor entropy parameters [pix][2C] (scales | means), fp32
    const float* pn;                // non-anchor entropy parameters [pix][2C] (may be null in the anchor pass)
    Act slot;                       // y_hat slice i inside the LRP/concat workspace (activation type)
    int B, H, W, C;
    int mode;                       // 0 forward, 1 compress, 2 decoder walk, 3 decompress: index list only, 4 decompress: symbols -> y_hat
    int vbr; float gain, rgain;     // VBR: gain g and 1/g as computed on the host in fp32
    float* lik; int lik_ld;         // forward: likelihood slice inside the fp32 NHWC [pix][M] buffer
    int32_t* sym; int32_t* idx;     // compress: base of this half-slice (flattened over [B,C,H,W/2])
    const float* table; int levels; // scale table (fp32, as stored in gaussian_conditional.scale_table)
    int sq;                         // 1: pa / pn hold only the anchor / non-anchor pixels, squeezed ([B][H][W/2][2C])
};
void launch_quant_anchor(int bf, const QuantArgs& a, cudaStream_t s);
void launch_quant_nonanchor(int bf, const QuantArgs& a, cudaStream_t s);

void launch_gc_flat(const float* y, const float* scales, const float* means, size_t n, float* y_hat, float* lik,
                    int32_t* sym, int32_t* idx, const float* table, int levels, cudaStream_t s);

// Rate/distortion sums: mode 0: *out += sum log2(a[i]) ; mode 1: *out += sum (a[i]-b[i])^2  (double, deterministic)
constexpr int RD_BLOCKS = 1024;
void launch_reduce(const float* a, const float* b, long long n, int mode, double* partial /*[RD_BLOCKS]*/, double* out,
                   cudaStream_t s);

// ---- tcgen05 / TMA implicit-GEMM convolution (gemm_tc.cu), bf16 activations only ----
struct TcConv {
    const void* in;             // bf16 NHWC view base
    int B, H, W, Cin, ld;       // input view as the conv sees it (for a stride-2 1x1: the sub-sampled grid)
    int sW, sH, sB;             // element strides of the W / H / B dims of that view
    int ks, pad;                // stride-1 taps around the output pixel
    const void* w;              // bf16 [N][ks*ks*Cpad], Cpad = roundup(Cin,64), zero padded per tap
    int Cpad;
    int prod;                   // 0: A = the input (TMA); 1: A = depthwise3x3(input) + dw bias; 2: A = input^2  (1x1 GEMMs only)
    const float* dw_w9;         // prod 1: depthwise weights [9][Cin] fp32, bias [Cin]
    const float* dw_bias;
    int ck;                     // 1 | 2 (1x1 GEMMs): the GEMM rows are only the ANCHOR | NON-ANCHOR pixels of the input view, in the squeezed
                                //    order of utils/ckbd.py:47-59 ([B][H][W/2], w = 2j + ((h + ck) & 1)); H / W describe the FULL view, Epi.Hout = H,
                                //    Epi.Wout = W / 2.  The gather is one 5-D TMA map (c, j, h & 1, h >> 1, b): no squeezed copy of the input exists.
    int ss;                     // 1: shift-sum form of a 3x3 (pad 1) subpel conv with 12 outputs: `w` is [9*12][Cpad] (row = tap*12 + n,
                                //    n in pixel-shuffle column order), Epi.N = 108, bias[0..11], out = fp32 NCHW [B][3][2H][2W]; ks = 1
};
// TMA-fed persistent depthwise 3x3 (bf16 NHWC); non-zero: not taken
int launch_dwconv3x3_tma(const Act& in, const Act& out, const float* w9, const float* bias, int stride, int act, cudaStream_t s);
bool tc_conv_supported(const TcConv& c, const Epi& e);
// returns cudaError_t-like int (0 ok)
int launch_conv_gemm_tc(const TcConv& c, const Epi& e, int vec, cudaStream_t s);
int tc_init();                   // resolves cuTensorMapEncodeTiled; 0 on success
void* tc_encode_fn();            // the resolved cuTensorMapEncodeTiled entry point (after tc_init)
const char* tc_last_error();

// ---- two-SM (cta_group::2) fused DepthWiseConv / (I)GDN-tail blocks (ds_pair.cu), bf16 NHWC, C channels in and out
struct DsPairArgs {
    const void* in; int B, H, W, ld;        // input view (stride-1 conv: output grid = input grid)
    int C;                                  // 192 | 128
    const float* dw_w9; const float* dw_bias;       // depthwise 3x3 weights [9][C] fp32, bias [C]
    const void* w1; const float* b1;        // pointwise weights bf16 [C][C] (K-major), bias [C]
    int act;                                // gdn == GDN_NONE: ACT_NONE | ACT_GELU
    int gdn;                                // GDN_NONE: out = act(pw(dw(x))) + res; GDN_FWD | GDN_INV: out = v * (r)sqrt(w2 v^2 + b2) + res, v = pw(dw(x))
    const void* w2; const float* b2;        // (I)GDN gamma bf16 [C][C], beta [C]
    const void* res; int res_ld;            // residual, addressed like the output (may be null)
    void* out; int out_ld;
};
bool ds_pair_supported(const DsPairArgs& a);
int launch_ds_pair(const DsPairArgs& a, cudaStream_t s);      // 0 ok
const char* ds_pair_last_error();

// ---- two-SM (cta_group::2) dense convolution (conv3_pair.cu), bf16 NHWC, stride 1: 3x3 pad 1 (Cin % 64 == 0, N % 256 == 0: the sub-pixel
//      convs of g_s) or 1x1 (any Cin % 8, N % 8: the wide GEMMs of the entropy model whose weights do not fit one SM)
struct Conv3PairArgs {
    const void* in; int B, H, W, Cin, ld;   // input view
    int ks, Cpad;                           // 3 | 1; channels per tap in `w` (Cin rounded up to 64)
    int ck;                                 // 1x1 only: 1 | 2 = GEMM rows are the anchor | non-anchor pixels of the input, squeezed ([B][H][W/2]), out likewise
    const void* w;                          // bf16 [N][ks * ks * Cpad], K-major, tap-major (the packing of TcConv::w; PixelShuffle folded into the column order)
    const float* bias;                      // [N]
    int N, act;                             // ACT_NONE | ACT_GELU
    int shuffle;                            // 1: PixelShuffle(2): out is [B][2H][2W][N / 4], (N / 4) % 64 == 0
    void* out; int out_ld;
    int res_inplace;                        // 1: out += act(conv(in) + bias): the output's current contents are the residual (read through the store maps)
};
bool conv3_pair_supported(const Conv3PairArgs& a);
int launch_conv3_pair(const Conv3PairArgs& a, cudaStream_t s);      // 0 ok
const char* conv3_pair_last_error();

// ---- halo-patch k x k convolution with a narrow output (conv_halo.cu), bf16 NHWC, stride 1, pad k / 2, k = 3 | 5, N <= 128
struct ConvHaloArgs {
    const void* in; int B, H, W, Cin, ld;   // input view
    const void* w; int Cpad;                // bf16 [N][ks * ks * Cpad] (the packing of TcConv::w)
    const float* bias;                      // [N] or null
    int N, ks;
    void* out; int out_ld;
    int swap;                               // 1: roles swapped (weights as the M operand, 256 pixels as N: conv_halo_t_kernel)
};
bool conv_halo_supported(const ConvHaloArgs& a);
int launch_conv_halo(const ConvHaloArgs& a, cudaStream_t s);        // 0 ok
const char* conv_halo_last_error();

// ---- three chained per-pixel layers in one launch (chain3.cu), bf16: rows [M][K1] -> [M][N3]
struct Chain3Args {
    int mode;                               // 0: EntropyParameters tail (GELU, GELU, fp32 out); 1: LocalContext tail (LayerNorm, GELU, + projection, bf16 out)
    const void* in; int M, K1, ld;          // bf16 rows
    const void* w1; int K1pad;              // bf16 [N1][K1pad]
    const void* w2;                         // bf16 [N2][N1]
    const void* w3;                         // bf16 [N3][N2]
    const float *b1, *b2, *b3;
    const float *ln_g, *ln_b; float ln_eps; // mode 1
    int N1, N2, N3;
    void* out; int out_ld;                  // mode 0: float [M][out_ld]; mode 1: bf16 [M][out_ld]
    int unsq_H, unsq_W;                     // mode 1, != 0: rows are the non-anchor pixels of a [.., H, W] grid in squeeze order and `out` is that grid (NHWC, out_ld per pixel): each row is stored at its pixel
};
bool chain3_supported(const Chain3Args& a);
int launch_chain3(const Chain3Args& a, cudaStream_t s);             // 0 ok
const char* chain3_last_error();

}  // namespace mlic
