/* mlic_b200 -- C ABI of the B200-native MLIC++ network-forward engine.
 *
 * The reference (LuZWCHA/MLIC) has no FFI: its seam is the Python model object returned by
 * models.get_model(name) (MLIC++/models/model_loader.py:4-18).  This header is what a binding for
 * that seam calls; mlic_b200/models.py is the ctypes binding and mirrors the reference classes.
 * Each entry point cites the reference method it replaces (paths relative to /root/reference/MLIC++).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * status otherwise, with a human-readable message available from mlic_last_error().  All image /
 * likelihood tensors are fp32 NCHW, exactly the reference's layout.  "dev" pointers are CUDA device
 * pointers on the engine's device; the caller owns every buffer.  No entry point synchronises the
 * stream unless it says so.  An engine is not re-entrant: use one engine per host thread / stream.
 */
#ifndef MLIC_B200_H_
#define MLIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlic_engine mlic_engine;

enum { MLIC_KIND_BASE = 0, MLIC_KIND_SD = 1, MLIC_KIND_VBR = 2, MLIC_KIND_SD_VBR = 3 };   /* MLICPlusPlus / ...SD / ...Vbr / ...SDVbr (bit 0: small decoder, bit 1: gains) */
enum { MLIC_PREC_FP32 = 0, MLIC_PREC_BF16 = 1 };                    /* validation mode / fast mode   */
enum { MLIC_MODE_FORWARD = 0, MLIC_MODE_COMPRESS = 1, MLIC_MODE_DECODER = 2, MLIC_MODE_DECOMPRESS = 3 /* mlic_decompress only */ };

/* Buffers of one call.  Unused outputs may be NULL.  B images of H x W (multiples of 64). */
typedef struct {
    const float* x;          /* in : [B,3,H,W]                      (forward, compress)                          */
    float* x_hat;            /* out: [B,3,H,W]                                                                    */
    float* y_likelihoods;    /* out: [B,M,H/16,W/16]                (forward)                                     */
    float* z_likelihoods;    /* out: [B,N,H/64,W/64]                (forward)                                     */
    int32_t* symbols;        /* out: [2*slices][B,C,H/16,W/32] flat, order A0,N0,A1,N1,...   (compress)           */
    int32_t* indexes;        /* out: same shape: CDF indexes into the 64-entry scale table   (compress)           */
    int32_t* z_symbols;      /* out: [B,N,H/64,W/64] round(z - median)                       (compress)           */
    float* y;                /* out (optional tap): g_a output [B,M,H/16,W/16]; INPUT when option "stages" has no g_a */
    float* y_hat;            /* out (optional tap): quantised latent after LRP [B,M,H/16,W/16]; INPUT for stages = 4 */
    double* rd_sums;         /* out (optional, forward): [2] = { sum log2(likelihoods), sum (x - x_hat)^2 }       */
} mlic_buffers;

/* Engine for one model configuration (config/config.py:19-62): N, M, slice_num and the class kind. */
int mlic_engine_create(int N, int M, int slice_num, int kind, mlic_engine** out);
void mlic_engine_destroy(mlic_engine* e);

/* One state_dict entry, reference name and shape (host fp32, copied).  Replaces load_state_dict
 * (models/mlicpp.py:461-468). */
int mlic_engine_set_param(mlic_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim);

/* Packs the weights onto the current CUDA device (NCHW fp32 -> GEMM layouts, bf16 copies, folded GDN /
 * EntropyBottleneck re-parametrisations, scale table of utils/func.py:16-19).  Call after the last
 * set_param and again after any parameter change; replaces update() (models/mlicpp.py:470-475). */
int mlic_engine_finalize(mlic_engine* e);

/* Knobs: "tensor_cores" (1 = tcgen05 implicit-GEMM in bf16 mode [default], 0 = CUDA-core GEMM only);
 *        "profile" (1 = bracket every tcgen05 GEMM launch with a CUDA-event pair on the launch stream);
 *        "fuse" (1 = depthwise 3x3 / x^2 produced inside the tcgen05 GEMM [default]); "trace" (see mlic_trace_dump);
 *        "stages" (bit mask, default 7): 1 = g_a, 2 = h_a + EntropyBottleneck + h_s + the slice loop, 4 = g_s.  Row-band
 *        sharding of one large image runs the three apart:  stages 1: x band -> `y` (H a multiple of 16);
 *        stages 2 | 6: `y` is an INPUT (the gathered bands), outputs as usual plus the `y_hat` tap; stages 4: `y_hat` is an
 *        INPUT (a band of latent rows, H = 16 x rows) -> x_hat band.  Stage subsets take device buffers (mlic_run) only. */
int mlic_engine_set_option(mlic_engine* e, const char* name, int value);
/* Float knobs: "z_qstep" -- quantisation step of the hyper prior z for the calls that follow (MLICPlusPlusVbr / MLICPlusPlusSDVbr with
 * vr_entbttlnck=True: EntropyBottleneckVbr(z, qs), models/mlicpp_vbr.py:253-259,553-559): z_hat = round((z - median) / qs) qs + median,
 * likelihood over [z_hat - qs / 2, z_hat + qs / 2], z symbols = round((z - median) / qs).  1 (default) is the plain EntropyBottleneck. */
int mlic_engine_set_option_f(mlic_engine* e, const char* name, float value);

/* Device workspace needed by one call of the given mode / precision / shape. */
int mlic_workspace_bytes(mlic_engine* e, int mode, int precision, int B, int H, int W, size_t* bytes);

/* forward(x) -> {x_hat, likelihoods}  (models/mlicpp.py:79-185; VBR: models/mlicpp_vbr.py:137-336 with
 * stage 2; `gain` is the level's Gain value or `inputscale`, ignored for non-VBR kinds).
 * mode = MLIC_MODE_COMPRESS: the network walk of compress() up to the rANS coder
 *   (models/mlicpp.py:199-277, utils/ckbd.py:123-144; VBR utils/ckbd.py:76-90,146-158) -> symbols / indexes.
 * mode = MLIC_MODE_DECODER: net_decoder_forward (models/mlicpp.py:380-459); `x` is not read. */
int mlic_run(mlic_engine* e, int mode, int precision, int B, int H, int W, float gain, const mlic_buffers* dev,
             void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Same call on HOST buffers: copies x in, runs, copies the requested outputs back and synchronises.
 * Workspace and device staging are owned and cached by the engine.  `pinned` != 0 promises that the host
 * pointers are page-locked (async copies). */
int mlic_run_host(mlic_engine* e, int mode, int precision, int B, int H, int W, float gain, const mlic_buffers* host,
                  int pinned);

/* decompress(strings, shape) (models/mlicpp.py:292-378; VBR models/mlicpp_vbr.py decompress with `gain`): the decoder-side
 * walk with the range decoder in the slice loop.  `z_symbols` (device, [B,N,H/64,W/64]) are the values the caller decoded
 * from z_strings (EntropyBottleneck.decompress before `+ medians`); `y_stream` (host) is the single y string of the batch.
 * Per half-slice the engine writes the CDF index list (utils/ckbd.py:195-229 order), copies it to a pinned mailbox, runs
 * mlic_rans_decode_stream on the calling thread against the tables of mlic_engine_set_cdf (gaussian_conditional's
 * _quantized_cdf / _cdf_length / _offset), and copies the symbols back: 2 x slice_num stream synchronisations per call.
 * x_hat: device [B,3,H,W]; y_hat: optional device tap.  workspace: mlic_workspace_bytes(mode = MLIC_MODE_DECOMPRESS). */
int mlic_engine_set_cdf(mlic_engine* e, const int32_t* cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                        int n_tables);
int mlic_decompress(mlic_engine* e, int precision, int B, int H, int W, float gain, const uint8_t* y_stream, size_t y_bytes,
                    const int32_t* z_symbols, float* x_hat, float* y_hat, void* workspace, size_t workspace_bytes,
                    void* cuda_stream);

/* Number of kernels the engine launched in the last mlic_run / mlic_run_host call. */
int64_t mlic_last_launch_count(const mlic_engine* e);

/* Live profile of the dominant kernel (the tcgen05 implicit GEMM) since the last reset:
 * out3 = { summed launch duration in ms, summed algorithmic FLOPs (2*M*N*K), launches }.  Synchronises. */
int mlic_profile_read(mlic_engine* e, double* out3, int reset);
/* Same bracket, restricted to the launches of the heaviest GEMM shape seen since the last reset (most algorithmic FLOPs
 * per launch; for MLICPP_L the 3x3 192 -> 768 sub-pixel convolutions): out3 = {summed ms, FLOPs PER LAUNCH, launches}. */
int mlic_profile_read_top(mlic_engine* e, double* out3, int reset);

/* Per-launch trace: with option "trace" = 1 every launch of the following calls is followed by a CUDA event on the
 * launch stream; this writes "label<TAB>microseconds" per launch to `path` (synchronises) and clears the trace. */
int mlic_trace_dump(mlic_engine* e, const char* path);

/* Stand-alone convolution on an NHWC activation tensor (fp32 or bf16 per `precision`; weights / bias are HOST
 * fp32 in the reference's nn.Conv2d layout [N][Cin][ks][ks]): out = act(conv(in) + bias) (+ residual), optionally
 * pixel-shuffled by 2 (CompressAI subpel_conv3x3).  tensor_cores = 1 selects the tcgen05 implicit-GEMM kernel where
 * it applies, 2 additionally the two-SM (cta_group::2) kernel for wide 3x3 convs, 0 the CUDA-core kernel.  Runs `iters` launches and reports the average duration of launches 2..iters
 * (CUDA events); synchronises.  Kernel-level test / micro-benchmark hook for nn.Conv2d call sites such as
 * modules/layers/conv.py:55-60 and res_blk.py:107-111. */
int mlic_conv2d_nhwc(int precision, int tensor_cores, const void* in, int B, int H, int W, int Cin, const float* weight,
                     const float* bias, int N, int ks, int stride, int pad, int act, int shuffle, const void* residual,
                     void* out, int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone depthwise 3x3 convolution (pad 1, stride 1 | 2) + bias (+ GELU when act = 1) on an NHWC activation
 * tensor; weight: HOST fp32 [C][1][3][3] (nn.Conv2d(groups=C) layout, modules/layers/conv.py:49-54), bias: HOST [C].
 * Timing as mlic_conv2d_nhwc.  Kernel-level test / micro-benchmark hook. */
int mlic_dwconv3x3_nhwc(int precision, const void* in, int B, int H, int W, int C, const float* weight, const float* bias,
                        int stride, int act, void* out, int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone DepthWiseConv (modules/layers/conv.py:46-63): depthwise 3x3 (pad 1, `stride`, bias) followed by a 1x1
 * convolution (bias), then act / + residual, on an NHWC activation tensor.  dw_weight: HOST fp32 [Cin][1][3][3],
 * pw_weight: HOST fp32 [N][Cin][1][1].  fuse = 1 lets the bf16 path produce the depthwise result inside the tcgen05
 * GEMM kernel (A-operand producer) where the layer shape allows; 2 additionally allows the two-SM (cta_group::2) kernel
 * for Cin = N = 192 | 128; 0 runs the two kernels back to back.  Timing as mlic_conv2d_nhwc.  Kernel-level test /
 * micro-benchmark hook. */
int mlic_dsconv_nhwc(int precision, int fuse, const void* in, int B, int H, int W, int Cin, const float* dw_weight,
                     const float* dw_bias, const float* pw_weight, const float* pw_bias, int N, int stride, int act,
                     const void* residual, void* out, int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone tail of ResidualBlockWithStride / ResidualBlockUpsample (modules/layers/res_blk.py:88-93,116-121), bf16:
 * v = DepthWiseConv(in) (C -> C); out = v * rsqrt(gamma v^2 + beta) (GDN; `inverse`: v * sqrt(.), IGDN) + residual.
 * in / residual / out: DEVICE bf16 NHWC [B,H,W,C]; dw_weight HOST [C][1][3][3]; pw_weight HOST [C][C]; gamma HOST [C][C] and
 * beta HOST [C] are the EFFECTIVE (re-parametrised, CompressAI NonNegativeParametrizer) values.  fuse = 2: one two-SM kernel,
 * v and v^2 stay on chip (fails if that kernel does not take the shape); 1: DepthWiseConv kernel + GDN GEMM that squares its
 * operand on chip; 0: unfused.  Timing as mlic_conv2d_nhwc.  Kernel-level test / micro-benchmark hook. */
int mlic_ds_gdn_nhwc(int fuse, const void* in, int B, int H, int W, int C, const float* dw_weight, const float* dw_bias,
                     const float* pw_weight, const float* pw_bias, const float* gamma, const float* beta, int inverse,
                     const void* residual, void* out, int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone final synthesis layer (subpel_conv3x3(C, 3, 2) of g_s, modules/transform/synthesis.py:67 with CompressAI's
 * subpel_conv3x3 = Conv2d(C, 12, 3, pad 1) + PixelShuffle(2)) of the bf16 path: in DEVICE bf16 NHWC [B,H,W,Cin], weight HOST
 * [12][Cin][3][3], bias HOST [12], out DEVICE fp32 NCHW [B,3,2H,2W].  impl 0: implicit-GEMM conv (halo-patch A operand);
 * impl 1: shift-sum form (one 1x1 GEMM with 9*12 columns, the taps summed in the epilogue).  Timing as mlic_conv2d_nhwc. */
int mlic_final_subpel(int impl, const void* in, int B, int H, int W, int Cin, const float* weight, const float* bias, float* out,
                      int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone LocalContext windowed attention (modules/transform/context.py:80-107; 5x5 window, 2 heads of 16, C = 32).
 * rel_bias: DEVICE fp32 [2][25][25] (relative_position_table gathered through relative_position_index).
 *   impl 0: fp32 CUDA-core kernel.  F: DEVICE fp32 [B*H*W][96], channels q|k|v with the reference's interleaved head
 *           split (c = d*2 + head); O: DEVICE fp32 [B*H*W][25][32] (channel = head*16 + d), every pixel.
 *   impl 1: as 0 with O in bf16.
 *   impl 2: bf16 warp-level tensor-core kernel.  F: DEVICE bf16 [B*H*W][96] HEAD-MAJOR (q_h0 q_h1 k_h0 k_h1 v_h0 v_h1);
 *           O: DEVICE bf16 [B*H*(W/2)][25][32]: the NON-ANCHOR pixels only, squeezed as utils/ckbd.py:47-59
 *           (w = 2j + (h & 1)).
 * Timing as mlic_conv2d_nhwc.  Kernel-level test / micro-benchmark hook. */
int mlic_local_attn(int impl, const void* F, int B, int H, int W, const float* rel_bias, void* O, int iters, float* avg_ms,
                    void* cuda_stream);

/* Stand-alone kernelised ("linear") global attention of LinearGlobalInterContext / LinearGlobalIntraContext
 * (modules/transform/context.py:169-193,226-245): qkv is a DEVICE NHWC tensor [B,H,W,3*D] (fp32 or bf16 per `precision`) whose
 * channels are Q | K | V (D each, `heads` heads of `hd` = D / heads channels, hd = 32 | 16); per head K is soft-maxed over the
 * positions, Q over the head's channels, out[p] = (K^ V^T)^T Q^[p] -> DEVICE NHWC [B,H,W,D] of the same element type.
 * par_kv / par_q: 0 none, 1 anchor, 2 non-anchor pixels only (the intra context: keys / values on anchors, queries on
 * non-anchors; filtered-out positions give zeros).  Timing as mlic_conv2d_nhwc.  Kernel-level test / micro-benchmark hook. */
int mlic_lin_attn(int precision, const void* qkv, int B, int H, int W, int D, int heads, int par_kv, int par_q, void* out, int iters,
                  float* avg_ms, void* cuda_stream);

/* Stand-alone three-layer per-pixel chain of the bf16 fast mode in ONE launch (chain3.cu), replacing per row of `in`
 *   mode 0: EntropyParameters layers 1..3 (modules/transform/entropy.py:13-17): out = W3 GELU(W2 GELU(W1 in + b1) + b2) + b3, fp32
 *   mode 1: the LocalContext tail (modules/transform/context.py:108-110 with proj folded into fusion): p = W1 in + b1,
 *           out = p + W3 GELU(W2 LayerNorm(p) + b2) + b3, bf16
 * in DEVICE bf16 [M][K1]; w1 [N1][K1], w2 [N2][N1], w3 [N3][N2], biases, ln_gamma / ln_beta [N1] HOST fp32;
 * N2 = 128, N3 = 64, N1 = 128 | 256 (mode 0) or 64 (mode 1); out DEVICE fp32 (mode 0) / bf16 (mode 1) [M][N3].
 * Timing as mlic_conv2d_nhwc.  Kernel-level test / micro-benchmark hook. */
int mlic_chain3(int mode, const void* in, int M, int K1, const float* w1, const float* b1, int N1, const float* w2, const float* b2, int N2,
                const float* w3, const float* b3, int N3, const float* ln_gamma, const float* ln_beta, void* out, int iters, float* avg_ms,
                void* cuda_stream);

/* Stand-alone g_a stage-0 head of the bf16 path (ResidualBlockWithStride(3 -> N, stride 2), modules/layers/res_blk.py:82-93
 * with DepthWiseConv, conv.py:46-63): x DEVICE fp32 NCHW [B,3,H,W] ->
 *   t_out    = GELU(point_conv(depth_conv_s2(x)))   DEVICE bf16 NHWC [B,H/2,W/2,N]
 *   skip_out = skip_1x1_s2(x)                        DEVICE bf16 NHWC [B,H/2,W/2,N]
 * dw_weight HOST [3][1][3][3], pw_weight / skip_weight HOST [N][3][1][1], biases HOST.  Timing as mlic_conv2d_nhwc. */
int mlic_ga_head(const float* x, int B, int H, int W, const float* dw_weight, const float* dw_bias, const float* pw_weight,
                 const float* pw_bias, const float* skip_weight, const float* skip_bias, int N, void* t_out, void* skip_out,
                 int iters, float* avg_ms, void* cuda_stream);

/* Stand-alone fused quantise / likelihood / CDF-index kernel on NCHW fp32 device tensors of one slice
 * (CompressAI GaussianConditional.forward / quantize / build_indexes; call sites models/mlicpp.py:132-134,
 * utils/ckbd.py:128-129).  y, scales, means, y_hat, lik: [n]; sym, idx: [n] (any may be NULL among outputs).
 * scale_table64: the 64-entry fp32 scale table on the device (gaussian_conditional.scale_table), or NULL for the
 * engine's own exp(linspace(log .11, log 256, 64)). */
int mlic_gaussian_conditional(const float* y, const float* scales, const float* means, size_t n, const float* scale_table64,
                              float* y_hat, float* lik, int32_t* sym, int32_t* idx, void* cuda_stream);

/* ---- host entropy coder (mlic_b200/csrc/rans.cpp): the range-ANS coder behind compress() / decompress().
 * Replaces compressai.ans.BufferedRansEncoder.encode_with_indexes + flush, RansDecoder.set_stream + decode_stream and
 * compressai._CXX.pmf_to_quantized_cdf as the reference calls them (MLIC++/models/mlicpp.py:212-216,279-280,303-304;
 * MLIC++/utils/ckbd.py:195-229).  Tables: `cdfs` is [n_tables][cdf_stride] int32 (16-bit precision), `cdf_sizes[t]` the
 * used length of row t, `offsets[t]` the symbol value of its first bin.  All return 0 on success. */
typedef struct mlic_rans_decoder mlic_rans_decoder;
int mlic_pmf_to_quantized_cdf(const float* pmf, int n, int32_t* cdf_out /* n + 1 */);
size_t mlic_rans_encode_bound(size_t n_symbols);
int mlic_rans_encode(const int32_t* symbols, const int32_t* indexes, size_t n, const int32_t* cdfs, int cdf_stride,
                     const int32_t* cdf_sizes, const int32_t* offsets, int n_tables, uint8_t* out, size_t out_cap,
                     size_t* out_bytes);
mlic_rans_decoder* mlic_rans_decoder_create(const uint8_t* stream, size_t nbytes);
void mlic_rans_decoder_destroy(mlic_rans_decoder* d);
int mlic_rans_decode_stream(mlic_rans_decoder* d, const int32_t* indexes, size_t n, const int32_t* cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int n_tables, int32_t* out);

const char* mlic_last_error(void);
const char* mlic_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MLIC_B200_H_ */
