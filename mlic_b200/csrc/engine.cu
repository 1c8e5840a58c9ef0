// The MLIC++ network walk (g_a, h_a, EntropyBottleneck, h_s, the 10-slice x (anchor, non-anchor)
// multi-reference entropy model, g_s) over the kernels of kernels.cu / gemm_tc.cu, and the C ABI of
// include/mlic_b200.h.
//
// Layout in HBM (one call, B images, latent grid h x w = H/16 x W/16, all activations NHWC):
//   LRPW  [B,h,w, Me + M]   : hyper_means | y_hat slice 0 | ... | slice S-1   -> every LRP / context input is a
//                             channel-prefix view, no torch.cat               (models/mlicpp.py:119,143-144)
//   EPW   [B,h,w, 10C + 2Me]: local 2C | intra 2C | inter 2C | channel 4C | hyper_params 2Me
//                             -> both EntropyParameters inputs are channel-suffix views (mlicpp.py:146,160)
//   y32   [B,h,w, M] fp32, lik [B,h,w, M] fp32, pa / pn [B,h,w, 2C] fp32 (entropy parameters never leave fp32)
// Activations are fp32 (validation mode) or bf16 (fast mode); accumulation is fp32 in both.
#include <cuda_runtime.h>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mlic_b200.h"
#include "kernels.h"

using namespace mlic;

static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}
#define CUDA_OK(x)                                                                             \
    do {                                                                                       \
        cudaError_t _e = (x);                                                                  \
        if (_e != cudaSuccess) return fail("%s: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {

struct HostT {
    std::vector<float> v;
    std::vector<int64_t> shape;
    int64_t dim(int i) const { return shape[i]; }
};

struct ConvW {          // GEMM-packed convolution / linear weights
    float* w32 = nullptr;       // [N][ks*ks*Cin], k = tap*Cin + c
    bf16* wbf = nullptr;        // [N][ks*ks*Cpad], zero padded per tap
    float* bias = nullptr;      // [N] (GEMM column order)
    int Cin = 0, N = 0, ks = 1, Cpad = 0, shuffle = 0;
};
struct DwW {
    float* w9 = nullptr;        // [9][C]
    float* bias = nullptr;      // [C]
    int C = 0;
};
struct LnW { float* g = nullptr; float* b = nullptr; int C = 0; };

struct EpiOpt {
    int act = ACT_NONE, premask = PAR_NONE, postmask = PAR_NONE;
    const Act* res = nullptr;
    int gdn = GDN_NONE;
    const Act* gdn_x = nullptr;
    const Act* out2 = nullptr;
    float* out_f32 = nullptr;   // fp32 destination [pix][out_f32_ld] instead of `out`
    int out_f32_ld = 0;
    int nchw = 0;               // out_f32 is a dense NCHW tensor (x_hat)
    int ck = 0;                 // 1 | 2: the GEMM rows are the anchor | non-anchor pixels of `in`, squeezed (TcConv::ck); tcgen05 path only
    int pm_w = 0, pm_codes = 0; // per-column-group premask (Epi::pm_w)
};

}  // namespace

// Host-buffer call (mlic_run_host): copies are pipelined against the walk image by image.  x[b] is uploaded on `s_in`
// while g_a works on image b-1; the likelihoods are downloaded on `s_out` while g_s runs; x_hat[b] is downloaded while
// g_s works on image b+1.
struct HostPipe {
    const float* hx = nullptr;        // host x [B,3,H,W]
    float* hx_hat = nullptr;          // host x_hat
    float* hy_lik = nullptr;          // host y_likelihoods
    float* hz_lik = nullptr;          // host z_likelihoods
    cudaStream_t s_in = nullptr, s_out = nullptr;
};

struct mlic_engine {
    int N, M, S, C, kind;
    bool sd, vbr;
    int Me;                  // channels of hyper_means as the entropy model sees them (M, or M/4 for SD)
    std::map<std::string, HostT> params;
    bool finalized = false;
    int use_tc = 1;
    // decompress (MLIC_MODE_DECOMPRESS): quantised CDF tables of gaussian_conditional and the pinned mailbox of the range decoder
    std::vector<int32_t> cdf_tab, cdf_sizes, cdf_offsets;
    int cdf_stride = 0;
    mlic_rans_decoder* rans = nullptr;     // attached for the duration of one mlic_decompress call
    int32_t* h_mail = nullptr;       // pinned: [indexes | symbols] of one half-slice
    size_t h_mail_n = 0;
    int stages = 7;          // bit 0: g_a, bit 1: h_a + EntropyBottleneck + h_s + slice loop, bit 2: g_s (row-band sharding runs them apart)
    int fuse = 1;            // bf16 + tensor cores: depthwise 3x3 and x^2 computed inside the GEMM kernel (A-operand producers)
    int pair = 1;            // fuse: DepthWiseConv / (I)GDN-tail blocks with C = N = 192 | 128 on the two-SM kernel (ds_pair.cu)
    int halo5 = getenv("MLIC_HALO5") ? atoi(getenv("MLIC_HALO5")) : 2;      // 5x5 convs with N <= 128 on the halo-patch kernels (conv_halo.cu): 1 pixels as M, 2 roles swapped (weights as M, 256 pixels as N)
    int folds = getenv("MLIC_FOLDS") ? atoi(getenv("MLIC_FOLDS")) : 1;      // algebraic folds of the bf16 fast path (pack_folds, pack_fusion_proj)
    int chain = getenv("MLIC_CHAIN") ? atoi(getenv("MLIC_CHAIN")) : 1;      // EntropyParameters / LocalContext tails as one launch (chain3.cu)
    int wide_pair = getenv("MLIC_WIDE_PAIR") ? atoi(getenv("MLIC_WIDE_PAIR")) : 1;      // wide 1x1 GEMMs on the two-SM kernel (conv3_pair.cu)
    float z_qstep = 1.0f;    // quantisation step of the hyper prior (EntropyBottleneckVbr, vr_entbttlnck=True); 1: the plain EntropyBottleneck

    std::vector<void*> dev_allocs;
    std::unordered_map<std::string, ConvW> convs;
    std::unordered_map<std::string, DwW> dws;
    std::unordered_map<std::string, LnW> lns;
    std::unordered_map<std::string, float*> misc;
    float* eb_packed = nullptr;
    float* eb_medians = nullptr;
    float* scale_table = nullptr;
    int scale_levels = 64;

    // per-call state
    int bf = 0;
    bool dry = false;
    cudaStream_t st = nullptr;
    uint8_t* ws_base = nullptr;
    size_t ws_size = 0, ws_off = 0, ws_peak = 0;
    int64_t launches = 0;
    int rc = 0;              // sticky error of the current walk

    // live profiling of the dominant kernel (tcgen05 implicit GEMM): CUDA-event pairs on the launch stream
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<double> ev_flops;
    double prof_ms = 0, prof_flops = 0, prof_launches = 0;
    double top_flops = 0, top_ms = 0, top_n = 0;
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        return ev_pool[ev_used++];
    }
    int profile_collect() {          // synchronises the events recorded so far and folds them into the sums
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float ms = 0;
            cudaError_t e = cudaEventSynchronize(ev_pool[i + 1]);
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]);
            if (e != cudaSuccess) return fail("profile: %s", cudaGetErrorString(e));
            prof_ms += ms; prof_flops += ev_flops[i / 2]; prof_launches += 1;
            // the heaviest launch shape (most algorithmic FLOPs per launch) is tracked separately: bench.py's roofline line
            const double f = ev_flops[i / 2];
            if (f > top_flops * (1.0 + 1e-9)) { top_flops = f; top_ms = ms; top_n = 1; }
            else if (f >= top_flops * (1.0 - 1e-9)) { top_ms += ms; top_n += 1; }
        }
        ev_used = 0; ev_flops.clear();
        return 0;
    }

    // per-launch trace (option "trace"): one event after every launch, labelled with the layer key
    int trace = 0;
    std::vector<cudaEvent_t> tr_ev;
    std::vector<std::string> tr_lab;
    void tr(const std::string& label) {
        if (!trace || dry) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
        tr_ev.push_back(e); tr_lab.push_back(label);
    }
    int trace_dump(const char* path) {
        FILE* f = fopen(path, "w");
        if (!f) return fail("cannot open %s", path);
        for (size_t i = 1; i < tr_ev.size(); ++i) {
            float ms = 0;
            cudaEventSynchronize(tr_ev[i]);
            cudaEventElapsedTime(&ms, tr_ev[i - 1], tr_ev[i]);
            fprintf(f, "%s\t%.3f\n", tr_lab[i].c_str(), ms * 1e3);
        }
        fclose(f);
        for (cudaEvent_t e : tr_ev) cudaEventDestroy(e);
        tr_ev.clear(); tr_lab.clear();
        return 0;
    }

    // host-call staging (mlic_run_host)
    void* h_ws = nullptr; size_t h_ws_bytes = 0;
    void* h_io = nullptr; size_t h_io_bytes = 0;
    std::map<std::array<int, 8>, size_t> ws_cache;       // mlic_workspace_bytes results
    cudaStream_t h_stream = nullptr, h_in = nullptr, h_out = nullptr;
    // Two independent branches of a slice (inter || channel context, intra || local context: mlicpp.py:140-150) run on two streams:
    // each is a chain of 4-10 launches of 20-500 us that keeps the SMs busy only part of the time (pipeline fill, tails, latency-bound
    // attention kernels), so the two chains fill each other's gaps.  MLIC_OVERLAP=0 runs them one after the other.
    std::map<int, cudaStream_t> side_streams;      // per device
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    template <class FA, class FB> void fork2(FA fa, FB fb) {
        static const int overlap = getenv("MLIC_OVERLAP") ? atoi(getenv("MLIC_OVERLAP")) : 1;
        // the second branch allocates ABOVE everything the first one touched, in the dry run and the sequential modes too (one
        // workspace plan whatever the schedule)
        const size_t base = ws_off, peak_before = ws_peak;
        const bool two = overlap && go() && !trace && !profile;
        cudaStream_t s2 = nullptr;
        if (two) {
            int dev = 0;
            cudaGetDevice(&dev);
            auto it = side_streams.find(dev);
            if (it == side_streams.end()) {
                cudaStream_t ns = nullptr;
                if (cudaStreamCreateWithFlags(&ns, cudaStreamNonBlocking) != cudaSuccess) { rc = fail("side stream: %s", cudaGetErrorString(cudaGetLastError())); return; }
                it = side_streams.emplace(dev, ns).first;
            }
            s2 = it->second;
            if (!ev_fork) { cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming); }
            cudaEventRecord(ev_fork, st);
            cudaStreamWaitEvent(s2, ev_fork, 0);
        }
        ws_peak = ws_off;
        fa();
        const size_t top_a = ws_peak;
        ws_off = top_a;
        cudaStream_t keep = st;
        if (two) st = s2;
        fb();
        st = keep;
        if (two) {
            cudaEventRecord(ev_join, s2);
            cudaStreamWaitEvent(st, ev_join, 0);
        }
        if (ws_peak < peak_before) ws_peak = peak_before;
        ws_off = base;
    }
    std::vector<cudaEvent_t> pipe_ev;
    size_t pipe_used = 0;
    cudaEvent_t pipe_event() {
        if (pipe_used == pipe_ev.size()) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); pipe_ev.push_back(e); }
        return pipe_ev[pipe_used++];
    }

    ~mlic_engine() {
        for (void* p : dev_allocs) cudaFree(p);
        for (cudaEvent_t ev : ev_pool) cudaEventDestroy(ev);
        if (h_ws) cudaFree(h_ws);
        if (h_io) cudaFree(h_io);
        if (h_mail) cudaFreeHost(h_mail);
        if (h_stream) cudaStreamDestroy(h_stream);
        if (h_in) cudaStreamDestroy(h_in);
        if (h_out) cudaStreamDestroy(h_out);
        for (cudaEvent_t ev : pipe_ev) cudaEventDestroy(ev);
        for (auto& kv : side_streams) cudaStreamDestroy(kv.second);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }

    // ------------------------------------------------------------------ parameter access / packing
    const HostT* get(const std::string& name) {
        auto it = params.find(name);
        if (it == params.end()) { rc = fail("missing parameter '%s'", name.c_str()); return nullptr; }
        return &it->second;
    }
    template <typename T> T* upload(const std::vector<T>& h) {
        void* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(h.size() * sizeof(T), 16)) != cudaSuccess) { rc = fail("cudaMalloc failed"); return nullptr; }
        dev_allocs.push_back(d);
        if (!h.empty()) cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
        return (T*)d;
    }
    static uint16_t f2bf(float f) {            // round-to-nearest-even, as __float2bfloat16_rn
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
        u += 0x7fffu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    }
    // w: [N][Cin][ks][ks] (+ bias [N]); shuffle: PixelShuffle(2) column permutation n' = (2r+s)*Cq + c <- o = 4c + 2r + s
    void pack_conv_raw(const std::string& key, const float* w, const float* b, int N_, int Cin, int ks, int shuffle) {
        ConvW cw;
        cw.Cin = Cin; cw.N = N_; cw.ks = ks; cw.shuffle = shuffle;
        cw.Cpad = (Cin + 63) / 64 * 64;
        const int taps = ks * ks;
        std::vector<float> w32((size_t)N_ * taps * Cin), bias(N_, 0.f);
        std::vector<uint16_t> wbf((size_t)N_ * taps * cw.Cpad, 0);
        for (int o = 0; o < N_; ++o) {
            int n = o;
            if (shuffle) { int Cq = N_ / 4; n = (o & 3) * Cq + (o >> 2); }
            for (int c = 0; c < Cin; ++c)
                for (int t = 0; t < taps; ++t) {
                    float v = w[((size_t)o * Cin + c) * taps + t];
                    w32[((size_t)n * taps + t) * Cin + c] = v;
                    wbf[((size_t)n * taps + t) * cw.Cpad + c] = f2bf(v);
                }
            if (b) bias[n] = b[o];
        }
        cw.w32 = upload(w32);
        cw.wbf = (bf16*)upload(wbf);
        cw.bias = upload(bias);
        convs[key] = cw;
    }
    void pack_conv(const std::string& p, int shuffle = 0) {
        const HostT* w = get(p + ".weight");
        const HostT* b = get(p + ".bias");
        if (!w || !b) return;
        int ks = w->shape.size() == 4 ? (int)w->dim(2) : 1;
        pack_conv_raw(p, w->v.data(), b->v.data(), (int)w->dim(0), (int)w->dim(1), ks, shuffle);
    }
    void pack_dw_list(const std::string& key, const std::vector<std::string>& ps) {     // concatenated depthwise convs
        int Ct = 0;
        for (auto& p : ps) { const HostT* w = get(p + ".weight"); if (!w) return; Ct += (int)w->dim(0); }
        std::vector<float> w9((size_t)9 * Ct), bias(Ct);
        int c0 = 0;
        for (auto& p : ps) {
            const HostT* w = get(p + ".weight");
            const HostT* b = get(p + ".bias");
            if (!w || !b) return;
            int Cc = (int)w->dim(0);
            for (int c = 0; c < Cc; ++c) {
                for (int t = 0; t < 9; ++t) w9[(size_t)t * Ct + c0 + c] = w->v[(size_t)c * 9 + t];
                bias[c0 + c] = b->v[c];
            }
            c0 += Cc;
        }
        DwW d; d.C = Ct; d.w9 = upload(w9); d.bias = upload(bias);
        dws[key] = d;
    }
    void pack_ds(const std::string& p) { pack_dw_list(p + ".depth_conv", {p + ".depth_conv"}); pack_conv(p + ".point_conv"); }
    void pack_c3(const std::string& p, bool dense) { if (dense) pack_conv(p); else pack_ds(p); }
    void pack_gdn(const std::string& p) {
        const HostT* beta = get(p + ".beta");
        const HostT* gamma = get(p + ".gamma");
        if (!beta || !gamma) return;
        const int Cc = (int)beta->dim(0);
        // NonNegativeParametrizer (CompressAI): eff = max(p, bound)^2 - pedestal, all in fp32 as torch computes it
        const float ped = (float)ldexp(1.0, -36);
        const float bound_b = sqrtf(1e-6f + ped), bound_g = sqrtf(0.0f + ped);
        std::vector<float> g((size_t)Cc * Cc), b(Cc);
        for (int i = 0; i < Cc; ++i) { float t = fmaxf(beta->v[i], bound_b); b[i] = t * t - ped; }
        for (size_t i = 0; i < g.size(); ++i) { float t = fmaxf(gamma->v[i], bound_g); g[i] = t * t - ped; }
        pack_conv_raw(p, g.data(), b.data(), Cc, Cc, 1, 0);
    }
    void pack_ln(const std::string& p) {
        const HostT* g = get(p + ".weight");
        const HostT* b = get(p + ".bias");
        if (!g || !b) return;
        LnW l; l.C = (int)g->dim(0); l.g = upload(g->v); l.b = upload(b->v);
        lns[p] = l;
    }
    void pack_rb(const std::string& p, bool dense) {
        pack_c3(p + ".conv1", dense); pack_c3(p + ".conv2", dense);
        if (params.count(p + ".skip.weight")) pack_conv(p + ".skip");
    }
    void pack_qkv(const std::string& p, bool fuse_pw) {
        if (fuse_pw) {      // one GEMM producing [Q | K | V]
            const HostT *wq = get(p + ".queries.0.weight"), *wk = get(p + ".keys.0.weight"), *wv = get(p + ".values.0.weight");
            const HostT *bq = get(p + ".queries.0.bias"), *bk = get(p + ".keys.0.bias"), *bv = get(p + ".values.0.bias");
            if (!wq || !wk || !wv || !bq || !bk || !bv) return;
            std::vector<float> w(wq->v), b(bq->v);
            w.insert(w.end(), wk->v.begin(), wk->v.end()); w.insert(w.end(), wv->v.begin(), wv->v.end());
            b.insert(b.end(), bk->v.begin(), bk->v.end()); b.insert(b.end(), bv->v.begin(), bv->v.end());
            int D = (int)wq->dim(0);
            pack_conv_raw(p + ".qkv_pw", w.data(), b.data(), 3 * D, D, 1, 0);
        } else {
            pack_conv(p + ".queries.0"); pack_conv(p + ".keys.0"); pack_conv(p + ".values.0");
        }
        pack_dw_list(p + ".qkv_dw", {p + ".queries.1", p + ".keys.1", p + ".values.1"});
    }
    // Algebraic folds of the bf16 fast path (the fp32 validation mode keeps the reference's chain of layers, op for op):
    //  * global_inter_context: skip(att) + mlp.4(h) is ONE GEMM over the channel concatenation [att | h] (context.py:245);
    //  * global_intra_context: the q / k / v 1x1 convs read adjacent channel ranges of one tensor (previous slice | current slot):
    //    one GEMM with the block matrix [[Wq 0] [Wk 0] [0 Wv]] and a premask per 32-column group (context.py:172-174).
    void pack_folds(const std::string& gi, const std::string& ga) {
        const HostT *ws = get(gi + ".skip.weight"), *bs = get(gi + ".skip.bias"), *w4 = get(gi + ".mlp.4.weight"), *b4 = get(gi + ".mlp.4.bias");
        if (ws && bs && w4 && b4 && ws->dim(0) == w4->dim(0)) {
            const int No = (int)ws->dim(0), Ka = (int)ws->dim(1), Kh = (int)w4->dim(1);
            std::vector<float> w((size_t)No * (Ka + Kh)), b(No);
            for (int o = 0; o < No; ++o) {
                for (int c = 0; c < Ka; ++c) w[(size_t)o * (Ka + Kh) + c] = ws->v[(size_t)o * Ka + c];
                for (int c = 0; c < Kh; ++c) w[(size_t)o * (Ka + Kh) + Ka + c] = w4->v[(size_t)o * Kh + c];
                b[o] = bs->v[o] + b4->v[o];
            }
            pack_conv_raw(gi + ".skip_mlp4", w.data(), b.data(), No, Ka + Kh, 1, 0);
        }
        const HostT *wq = get(ga + ".queries.0.weight"), *wk = get(ga + ".keys.0.weight"), *wv = get(ga + ".values.0.weight");
        const HostT *bq = get(ga + ".queries.0.bias"), *bk = get(ga + ".keys.0.bias"), *bv = get(ga + ".values.0.bias");
        if (wq && wk && wv && bq && bk && bv) {
            const int D = (int)wq->dim(0);
            if ((int)wq->dim(1) == D && (D % 8) == 0) {
                std::vector<float> w((size_t)3 * D * 2 * D, 0.f), b(3 * D);
                for (int o = 0; o < D; ++o)
                    for (int c = 0; c < D; ++c) {
                        w[(size_t)o * 2 * D + c] = wq->v[(size_t)o * D + c];
                        w[(size_t)(D + o) * 2 * D + c] = wk->v[(size_t)o * D + c];
                        w[(size_t)(2 * D + o) * 2 * D + D + c] = wv->v[(size_t)o * D + c];
                    }
                for (int o = 0; o < D; ++o) { b[o] = bq->v[o]; b[D + o] = bk->v[o]; b[2 * D + o] = bv->v[o]; }
                pack_conv_raw(ga + ".qkv_blk", w.data(), b.data(), 3 * D, 2 * D, 1, 0);
            }
        }
    }
    //  * local_context: proj(fusion(windows)) has no non-linearity in between (context.py:108-109): W' = Wp Wf, b' = Wp bf + bp (in double)
    void pack_fusion_proj(const std::string& lc) {
        const HostT *wf = get(lc + ".fusion.weight"), *bf_ = get(lc + ".fusion.bias"), *wp = get(lc + ".proj.weight"), *bp = get(lc + ".proj.bias");
        if (!wf || !bf_ || !wp || !bp) return;
        const int No = (int)wf->dim(0), Ci = (int)wf->dim(1), Np = (int)wp->dim(0);
        if ((int)wp->dim(1) != No || wf->shape.size() != 4 || wf->dim(2) != 5 || wf->dim(3) != 5) return;
        std::vector<float> wl((size_t)Np * 25 * Ci), bl(Np);
        for (int o = 0; o < Np; ++o) {
            double bacc = bp->v[o];
            for (int m = 0; m < No; ++m) bacc += (double)wp->v[(size_t)o * No + m] * (double)bf_->v[m];
            bl[o] = (float)bacc;
            for (int c = 0; c < Ci; ++c)
                for (int t = 0; t < 25; ++t) {
                    double acc = 0.0;
                    for (int m = 0; m < No; ++m) acc += (double)wp->v[(size_t)o * No + m] * (double)wf->v[((size_t)m * Ci + c) * 25 + t];
                    wl[((size_t)o * 25 + t) * Ci + c] = (float)acc;        // K = tap * C + c, as the `fusion` packing
                }
        }
        pack_conv_raw(lc + ".fusion_proj", wl.data(), bl.data(), Np, 25 * Ci, 1, 0);
    }
    void pack_dw_mlp(const std::string& p) {
        pack_conv(p + ".0"); pack_dw_list(p + ".2", {p + ".2"}); pack_conv(p + ".4");
    }

    int finalize() {
        for (void* p : dev_allocs) cudaFree(p);
        dev_allocs.clear(); convs.clear(); dws.clear(); lns.clear(); misc.clear(); ws_cache.clear();
        rc = 0;
        // g_a / h_a
        std::string p = "g_a.analysis_transform.";
        for (int i : {0, 2, 4}) {
            std::string q = p + std::to_string(i);
            pack_c3(q + ".conv1", sd); pack_c3(q + ".conv2", sd); pack_gdn(q + ".gdn"); pack_conv(q + ".skip");
            pack_rb(p + std::to_string(i + 1), sd);
        }
        pack_c3(p + "6", sd);
        for (int i : {0, 2, 4, 6, 8}) pack_c3("h_a.reduction." + std::to_string(i), sd);
        // g_s / h_s
        p = "g_s.synthesis_transform.";
        pack_rb(p + "0", false);
        for (int i : {1, 3, 5}) {
            std::string q = p + std::to_string(i);
            pack_conv(q + ".subpel_conv.0", 1); pack_ds(q + ".conv"); pack_gdn(q + ".igdn"); pack_conv(q + ".upsample.0", 1);
            pack_rb(p + std::to_string(i + 1), false);
        }
        pack_conv(p + "7.0", 1);
        {   // shift-sum form of the final subpel conv (gemm_tc.cu STORE_SS): row tap*12 + n' <- w[o][c][tap], n' = (o & 3)*3 + (o >> 2)
            const HostT* w = get(p + "7.0.weight");
            const HostT* b = get(p + "7.0.bias");
            if (w && b && w->shape.size() == 4 && w->dim(0) == 12 && w->dim(2) == 3 && w->dim(3) == 3) {
                const int Ci = (int)w->dim(1);
                std::vector<float> ws((size_t)108 * Ci), bs(108, 0.f);
                for (int o = 0; o < 12; ++o) {
                    const int n = (o & 3) * 3 + (o >> 2);
                    for (int c = 0; c < Ci; ++c)
                        for (int t = 0; t < 9; ++t) ws[(size_t)(t * 12 + n) * Ci + c] = w->v[((size_t)o * Ci + c) * 9 + t];
                    bs[n] = b->v[o];
                }
                pack_conv_raw(p + "7.0_ss", ws.data(), bs.data(), 108, Ci, 1, 0);
            }
        }
        p = "h_s.increase.";
        pack_ds(p + "0"); pack_conv(p + "2.0", 1); pack_ds(p + "4"); pack_conv(p + "6.0", 1); pack_ds(p + "8");
        // entropy model
        for (int i = 0; i < S; ++i) {
            std::string is = std::to_string(i);
            std::string lc = "local_context." + is;
            pack_ln(lc + ".norm1"); pack_ln(lc + ".norm2");
            pack_conv(lc + ".qkv_proj"); pack_conv(lc + ".proj"); pack_conv(lc + ".mlp.fc1"); pack_conv(lc + ".mlp.fc2");
            if (C == 32) {   // head-major copy of qkv_proj for the tensor-core attention kernel: row u*C + hh*16 + dd <- u*C + dd*2 + hh
                const HostT* w = get(lc + ".qkv_proj.weight");
                const HostT* b = get(lc + ".qkv_proj.bias");
                if (w && b && w->dim(0) == 3 * C) {
                    const int Ci = (int)w->dim(1);
                    std::vector<float> wp((size_t)3 * C * Ci), bp(3 * C);
                    for (int u = 0; u < 3; ++u)
                        for (int hh = 0; hh < 2; ++hh)
                            for (int dd = 0; dd < 16; ++dd) {
                                const int o = u * C + dd * 2 + hh, n = u * C + hh * 16 + dd;
                                for (int c = 0; c < Ci; ++c) wp[(size_t)n * Ci + c] = w->v[(size_t)o * Ci + c];
                                bp[n] = b->v[o];
                            }
                    pack_conv_raw(lc + ".qkv_proj_hm", wp.data(), bp.data(), 3 * C, Ci, 1, 0);
                }
            }
            {   // fusion Conv2d(C,2C,k=5) applied to [C,5,5] windows == linear over K = tap*C + c
                const HostT* w = get(lc + ".fusion.weight");
                const HostT* b = get(lc + ".fusion.bias");
                if (w && b) {
                    int No = (int)w->dim(0), Ci = (int)w->dim(1);
                    std::vector<float> wl((size_t)No * 25 * Ci);
                    for (int o = 0; o < No; ++o)
                        for (int c = 0; c < Ci; ++c)
                            for (int t = 0; t < 25; ++t) wl[((size_t)o * 25 + t) * Ci + c] = w->v[((size_t)o * Ci + c) * 25 + t];
                    pack_conv_raw(lc + ".fusion", wl.data(), b->v.data(), No, 25 * Ci, 1, 0);
                }
                // relative position bias [2][25][25] = table[index[a][b]][head]   (context.py:95-100)
                const HostT* tab = get(lc + ".relative_position_table");
                const HostT* idx = get(lc + ".relative_position_index");
                if (tab && idx) {
                    std::vector<float> rb(2 * 625);
                    for (int hh = 0; hh < 2; ++hh)
                        for (int ab = 0; ab < 625; ++ab) rb[hh * 625 + ab] = tab->v[(size_t)((int)idx->v[ab]) * 2 + hh];
                    misc[lc + ".rel_bias"] = upload(rb);
                }
            }
            for (const char* tag : {"anchor", "nonanchor"}) {
                std::string ep = std::string("entropy_parameters_") + tag + "." + is + ".fusion.";
                for (int j : {0, 2, 4, 6}) pack_conv(ep + std::to_string(j));
                std::string lr = std::string("lrp_") + tag + "." + is + ".lrp_transform.";
                for (int j = 0; j < (sd ? 4 : 3); ++j) pack_ds(lr + std::to_string(2 * j));
            }
            if (i > 0) {
                std::string cc = "channel_context." + is + ".fushion.";
                for (int j : {0, 2, 4}) pack_c3(cc + std::to_string(j), sd);
                std::string gi = "global_inter_context." + is;
                pack_qkv(gi, true); pack_conv(gi + ".reprojection"); pack_dw_mlp(gi + ".mlp"); pack_conv(gi + ".skip");
                std::string ga = "global_intra_context." + is;
                pack_qkv(ga, false); pack_conv(ga + ".reprojection"); pack_dw_mlp(ga + ".mlp");
                pack_folds(gi, ga);
            }
            pack_fusion_proj(lc);
        }
        // EntropyBottleneck: softplus(matrices), tanh(factors) folded
        {
            std::vector<float> packed((size_t)N * 58), med(N);
            const int moff[5] = {0, 3, 12, 21, 30}, msz[5] = {3, 9, 9, 9, 3};
            const int boff[5] = {33, 36, 39, 42, 45}, bsz[5] = {3, 3, 3, 3, 1};
            const int foff[4] = {46, 49, 52, 55};
            for (int i = 0; i < 5; ++i) {
                const HostT* m = get("entropy_bottleneck.matrices." + std::to_string(i));
                const HostT* b = get("entropy_bottleneck.biases." + std::to_string(i));
                if (!m || !b) break;
                for (int c = 0; c < N; ++c) {
                    for (int j = 0; j < msz[i]; ++j) {
                        float x = m->v[(size_t)c * msz[i] + j];
                        packed[(size_t)c * 58 + moff[i] + j] = x > 20.f ? x : log1pf(expf(x));     // F.softplus
                    }
                    for (int j = 0; j < bsz[i]; ++j) packed[(size_t)c * 58 + boff[i] + j] = b->v[(size_t)c * bsz[i] + j];
                }
                if (i < 4) {
                    const HostT* f = get("entropy_bottleneck.factors." + std::to_string(i));
                    if (!f) break;
                    for (int c = 0; c < N; ++c)
                        for (int j = 0; j < 3; ++j) packed[(size_t)c * 58 + foff[i] + j] = tanhf(f->v[(size_t)c * 3 + j]);
                }
            }
            const HostT* q = get("entropy_bottleneck.quantiles");
            if (q) for (int c = 0; c < N; ++c) med[c] = q->v[(size_t)c * 3 + 1];
            eb_packed = upload(packed);
            eb_medians = upload(med);
        }
        {   // utils/func.py:16-19 -- exp(linspace(log .11, log 256, 64)) in fp32 (torch.linspace is symmetric around the middle)
            // (a table of any length set through update(scale_table=...) is used as it is: build_indexes only counts the
            // entries below sigma; the 0.11 lower bound of the kernels is GaussianConditional's scale_bound, not table[0])
            std::vector<float> tab(64);
            auto it = params.find("gaussian_conditional.scale_table");
            if (it != params.end() && !it->second.v.empty()) {
                tab = it->second.v;
                for (size_t k = 1; k < tab.size(); ++k)
                    if (!(tab[k] > tab[k - 1])) return fail("gaussian_conditional.scale_table must be strictly increasing (entry %zu)", k);
            } else {
                const float lo = logf(0.11f), hi = logf(256.0f);
                const float step = (hi - lo) / 63.0f;
                for (int k = 0; k < 64; ++k) {
                    float v = k < 32 ? lo + step * (float)k : hi - step * (float)(63 - k);
                    tab[k] = expf(v);
                }
            }
            scale_table = upload(tab);
            scale_levels = (int)tab.size();
        }
        if (rc) return rc;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return fail("finalize: %s", cudaGetErrorString(e));
        finalized = true;
        return 0;
    }

    // ------------------------------------------------------------------ workspace arena
    int esz() const { return bf ? 2 : 4; }
    void* ws_alloc(size_t bytes) {
        size_t o = (ws_off + 255) & ~(size_t)255;
        ws_off = o + bytes;
        if (ws_off > ws_peak) ws_peak = ws_off;
        if (!dry && ws_off > ws_size) { if (!rc) rc = fail("workspace too small: need > %zu, have %zu", ws_off, ws_size); return ws_base; }
        return ws_base + o;
    }
    Act act(int B, int H, int W, int Cc) {
        Act a; a.B = B; a.H = H; a.W = W; a.C = Cc; a.ld = (Cc + 7) / 8 * 8;
        a.p = ws_alloc((size_t)B * H * W * a.ld * esz());
        return a;
    }
    float* f32(size_t n) { return (float*)ws_alloc(n * 4); }
    Act view(const Act& a, int c0, int Cc) const {
        Act v = a; v.p = (uint8_t*)a.p + (size_t)c0 * esz(); v.C = Cc;
        return v;
    }
    bool go() const { return !dry && !rc; }
    void after_launch(const char* what) {
        ++launches;
        if (trace) tr(what);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && !rc) rc = fail("%s launch: %s", what, cudaGetErrorString(e));
    }

    // ------------------------------------------------------------------ primitive layers
    const ConvW* cw(const std::string& k) {
        auto it = convs.find(k);
        if (it == convs.end()) { if (!rc) rc = fail("conv '%s' not packed", k.c_str()); return nullptr; }
        return &it->second;
    }
    const DwW* dw(const std::string& k) {
        auto it = dws.find(k);
        if (it == dws.end()) { if (!rc) rc = fail("dwconv '%s' not packed", k.c_str()); return nullptr; }
        return &it->second;
    }
    bool al4(const void* p) const { return ((uintptr_t)p % (size_t)(4 * esz())) == 0; }

    // out = epilogue(conv(in, W)); `out` is an activation view of w->N (or N/4 when shuffled) channels.
    // prod != 0: tcgen05 kernel with a fused A-operand producer (1: depthwise 3x3 `dwp` of `in`, 2: in^2); returns false
    // without launching anything when that kernel does not take the layer (the caller then runs the unfused sequence).
    bool gemm(const Act& in, const std::string& key, int stride, int pad, const Act* out, const EpiOpt& o, int prod = 0,
              const DwW* dwp = nullptr) {
        const ConvW* w = cw(key);
        if (!w) return true;
        if (in.C != w->Cin) { if (!rc) rc = fail("gemm '%s': input has %d channels, weights expect %d", key.c_str(), in.C, w->Cin); return true; }
        Epi e;
        memset(&e, 0, sizeof e);
        e.bias = w->bias; e.act = o.act; e.premask = o.premask; e.postmask = o.postmask;
        if (o.pm_w) { e.pm_w = o.pm_w; e.pm_codes = o.pm_codes; e.premask = PAR_NONANCHOR; }     // (premask != 0 keeps the layer off the kernels without a premask)
        e.N = w->N; e.shuffle = w->shuffle;
        e.Hout = (in.H + 2 * pad - w->ks) / stride + 1;
        e.Wout = (in.W + 2 * pad - w->ks) / stride + 1;
        if (o.ck) {
            if (!(bf && use_tc && w->ks == 1 && stride == 1 && pad == 0 && (in.W % 2) == 0 && (in.H % 2) == 0)) { if (!rc) rc = fail("gemm '%s': checkerboard rows need the tcgen05 1x1 path", key.c_str()); return true; }
            e.Wout = in.W / 2;
        }
        bool vec = (w->N % 4 == 0);
        if (o.nchw) {
            if (!dry && !o.out_f32) { if (!rc) rc = fail("gemm '%s': no output", key.c_str()); return true; }
            e.out = o.out_f32; e.out_ld = 0; e.out_f32 = 1; e.nchw = 1; vec = false;
        }
        else if (o.out_f32) { e.out = o.out_f32; e.out_ld = o.out_f32_ld; e.out_f32 = 1; vec = vec && (o.out_f32_ld % 4 == 0) && ((uintptr_t)o.out_f32 % 16 == 0); }
        else if (out) { e.out = out->p; e.out_ld = out->ld; vec = vec && (out->ld % 4 == 0) && al4(out->p); }
        else { if (!rc) rc = fail("gemm '%s': no output", key.c_str()); return true; }
        if (o.res) { e.res = o.res->p; e.res_ld = o.res->ld; vec = vec && (o.res->ld % 4 == 0) && al4(o.res->p); }
        if (o.gdn) { e.gdn = o.gdn; e.gdn_x = o.gdn_x->p; e.gdn_ld = o.gdn_x->ld; vec = vec && (o.gdn_x->ld % 4 == 0) && al4(o.gdn_x->p); }
        if (o.out2) { e.out2 = o.out2->p; e.out2_ld = o.out2->ld; vec = vec && (o.out2->ld % 4 == 0) && al4(o.out2->p); }
        if (w->shuffle && ((w->N / 4) % 4 != 0)) vec = false;
        if (prod && !(bf && use_tc && fuse && stride == 1)) return false;
        if (bf && use_tc && (stride == 1 || w->ks == 1)) {
            TcConv t;
            memset(&t, 0, sizeof t);
            t.prod = prod;
            if (prod == 1) { t.dw_w9 = dwp->w9; t.dw_bias = dwp->bias; }
            t.in = in.p; t.B = in.B; t.Cin = in.C; t.ld = in.ld; t.ks = w->ks; t.pad = pad; t.w = w->wbf; t.Cpad = w->Cpad;
            t.H = e.Hout + (w->ks - 1) - 2 * pad;     // == in.H for stride 1
            t.W = e.Wout + (w->ks - 1) - 2 * pad;
            if (stride == 1) { t.H = in.H; t.W = in.W; }
            else { t.H = e.Hout; t.W = e.Wout; }      // 1x1 stride s: sub-sampled grid
            t.sW = in.ld * stride; t.sH = in.W * in.ld * stride; t.sB = in.H * in.W * in.ld;
            t.ck = o.ck;
            const bool sup = tc_conv_supported(t, e);
            if (prod && !sup) return false;
            if (o.ck && !sup) { if (!rc) rc = fail("gemm '%s': checkerboard rows not supported for this layer", key.c_str()); return true; }
            if (!go()) return true;
            // 5x5 convs with a narrow output (the re-projections of the global contexts): activations staged once per tile as a halo patch,
            // only the weights stream (conv_halo.cu)
            if (sup && halo5 && !prod && !o.ck && w->ks == 5 && stride == 1 && pad == 2 && !e.res && !e.gdn && !e.out2 && !e.out_f32 && !e.nchw &&
                !e.premask && !e.postmask && e.act == ACT_NONE && !w->shuffle) {
                ConvHaloArgs a;
                memset(&a, 0, sizeof a);
                a.in = in.p; a.B = in.B; a.H = in.H; a.W = in.W; a.Cin = in.C; a.ld = in.ld;
                a.w = w->wbf; a.Cpad = w->Cpad; a.bias = w->bias; a.N = w->N; a.ks = 5; a.out = e.out; a.out_ld = e.out_ld;
                // roles swapped (256 pixels per tile) once the 128-pixel tiles would need a second wave; the two kernels agree bit for bit
                a.swap = halo5 == 3 || (halo5 == 2 && (long long)in.B * ((in.H + 7) / 8) * ((in.W + 15) / 16) > 148);
                if (conv_halo_supported(a)) {
                    cudaEvent_t ev1 = nullptr;
                    if (profile) {
                        cudaEventRecord(next_event(), st);
                        ev1 = next_event();
                        ev_flops.push_back(2.0 * (double)in.B * e.Hout * e.Wout * (double)w->N * (double)(25 * w->Cin));
                    }
                    int r = launch_conv_halo(a, st);
                    if (ev1) cudaEventRecord(ev1, st);
                    if (r) { if (!rc) rc = fail("halo conv '%s': %s", key.c_str(), conv_halo_last_error()); return true; }
                    ++launches;
                    if (trace) {
                        char lab[256];
                        snprintf(lab, sizeof lab, "%s [halo conv5 M=%d N=%d K=25x%d]", key.c_str(), in.B * e.Hout * e.Wout, w->N, w->Cin);
                        tr(lab);
                    }
                    return true;
                }
            }
            // wide dense 3x3 convs (the sub-pixel convs of g_s): two-SM kernel, each CTA stages half of the weight tile (conv3_pair.cu)
            const bool res_inplace = e.res && e.res == e.out && e.res_ld == e.out_ld;
            // ... and the wide 1x1 GEMMs whose weight matrix does not fit one SM (EntropyParameters / LRP first layers, the q|k|v projection):
            // re-streaming the weights per 128-pixel tile is what bounds them (L2 -> shared memory), a CTA pair stages half each
            // (up to two column tiles: with more, the activation tile is re-read per column tile and the one-SM kernel's wider spread wins: q|k|v, N = 864)
            const bool wide1 = w->ks == 1 && stride == 1 && pad == 0 && !e.res && w->N >= 192 && w->N <= 384 && w->Cin >= 256 && !w->shuffle &&
                               (size_t)w->Cpad * w->N * 2 > 80 * 1024 && in.B * e.Hout * e.Wout >= 4 * 148 * 128 && wide_pair;
            const bool conv3 = w->ks == 3 && stride == 1 && pad == 1 && (!e.res || res_inplace) && w->Cpad == w->Cin && !o.ck;
            if (sup && pair && !prod && (conv3 || wide1) && !e.gdn && !e.out2 && !e.out_f32 && !e.nchw && !e.premask && !e.postmask) {
                Conv3PairArgs a;
                memset(&a, 0, sizeof a);
                a.in = in.p; a.B = in.B; a.H = in.H; a.W = in.W; a.Cin = in.C; a.ld = in.ld;
                a.ks = w->ks; a.Cpad = w->Cpad; a.ck = o.ck;
                a.w = w->wbf; a.bias = w->bias; a.N = w->N; a.act = e.act; a.shuffle = w->shuffle; a.out = e.out; a.out_ld = e.out_ld;
                a.res_inplace = res_inplace ? 1 : 0;
                if (conv3_pair_supported(a)) {
                    cudaEvent_t ev1 = nullptr;
                    if (profile) {
                        cudaEventRecord(next_event(), st);
                        ev1 = next_event();
                        ev_flops.push_back(2.0 * (double)in.B * e.Hout * e.Wout * (double)w->N * (double)(w->ks * w->ks * w->Cin));
                    }
                    int r = launch_conv3_pair(a, st);
                    if (ev1) cudaEventRecord(ev1, st);
                    if (r) { if (!rc) rc = fail("two-SM conv '%s': %s", key.c_str(), conv3_pair_last_error()); return true; }
                    ++launches;
                    if (trace) {
                        char lab[256];
                        snprintf(lab, sizeof lab, "%s [pair conv%d M=%d N=%d K=%dx%d]", key.c_str(), w->ks, in.B * e.Hout * e.Wout, w->N, w->ks * w->ks, w->Cin);
                        tr(lab);
                    }
                    return true;
                }
            }
            if (sup) {
                cudaEvent_t ev1 = nullptr;
                if (profile) {
                    cudaEventRecord(next_event(), st);
                    ev1 = next_event();
                    ev_flops.push_back(2.0 * (double)in.B * e.Hout * e.Wout * (double)w->N * (double)(w->ks * w->ks * w->Cin));
                }
                int r = launch_conv_gemm_tc(t, e, vec ? 1 : 0, st);
                if (ev1) cudaEventRecord(ev1, st);
                if (r) { if (!rc) rc = fail("tcgen05 conv '%s': %s", key.c_str(), tc_last_error()); return true; }
                ++launches;
                if (trace) {
                    char lab[256];
                    snprintf(lab, sizeof lab, "%s [tc M=%d N=%d K=%dx%d prod=%d]", key.c_str(), in.B * e.Hout * e.Wout, w->N, w->ks * w->ks, w->Cin, prod);
                    tr(lab);
                }
                return true;
            }
        }
        if (!go()) return true;
        ConvGeom g;
        g.B = in.B; g.H = in.H; g.W = in.W; g.Cin = in.C; g.ld = in.ld; g.Hout = e.Hout; g.Wout = e.Wout;
        g.ks = w->ks; g.stride = stride; g.pad = pad; g.Ktot = w->ks * w->ks * w->Cin;
        int avec = vec && (in.ld % 4 == 0) && al4(in.p);
        // the kernel's `vec` covers both the A loads and the epilogue accesses
        launch_conv_gemm_simt(bf, in.p, g, w->w32, e, (avec && vec) ? 1 : 0, st);
        after_launch(key.c_str());
        return true;
    }
    // final 3x3 subpel conv in shift-sum form (tcgen05 kernel, STORE_SS); false: not taken, nothing launched
    bool gemm_ss(const Act& in, const std::string& key, float* xhat_nchw) {
        const ConvW* w = cw(key);
        if (!w || w->N != 108 || in.C != w->Cin) return false;
        Epi e;
        memset(&e, 0, sizeof e);
        e.bias = w->bias; e.N = 108; e.shuffle = 1; e.Hout = in.H; e.Wout = in.W;
        e.out = xhat_nchw; e.out_f32 = 1; e.nchw = 1;
        TcConv t;
        memset(&t, 0, sizeof t);
        t.ss = 1; t.in = in.p; t.B = in.B; t.H = in.H; t.W = in.W; t.Cin = in.C; t.ld = in.ld; t.ks = 1; t.pad = 0; t.w = w->wbf; t.Cpad = w->Cpad;
        t.sW = in.ld; t.sH = in.W * in.ld; t.sB = in.H * in.W * in.ld;
        if (!tc_conv_supported(t, e)) return false;
        if (!go()) return true;
        if (!xhat_nchw) { if (!rc) rc = fail("gemm '%s': no output", key.c_str()); return true; }
        cudaEvent_t ev1 = nullptr;
        if (profile) {
            cudaEventRecord(next_event(), st);
            ev1 = next_event();
            ev_flops.push_back(2.0 * (double)in.B * in.H * in.W * 108.0 * (double)w->Cin);
        }
        int r = launch_conv_gemm_tc(t, e, 0, st);
        if (ev1) cudaEventRecord(ev1, st);
        if (r) { if (!rc) rc = fail("tcgen05 conv '%s': %s", key.c_str(), tc_last_error()); return true; }
        ++launches;
        if (trace) {
            char lab[256];
            snprintf(lab, sizeof lab, "%s [tc shift-sum M=%d N=108 K=1x%d]", key.c_str(), in.B * in.H * in.W, w->Cin);
            tr(lab);
        }
        return true;
    }
    void dwconv(const Act& in, const std::string& key, int stride, int actv, const Act& out) {
        const DwW* d = dw(key);
        if (!d) return;
        if (d->C != in.C) { if (!rc) rc = fail("dwconv '%s': %d channels vs %d", key.c_str(), in.C, d->C); return; }
        if (!go()) return;
        launch_dwconv3x3(bf, in, out, d->w9, d->bias, stride, actv, st);
        after_launch(key.c_str());
    }
    // Two-SM kernel (ds_pair.cu): DepthWiseConv `p` (+ GELU, + residual), or with `gdn` the whole tail
    // v = DepthWiseConv(in); out = v * (r)sqrt(gamma v^2 + beta) + res.  false: not taken, nothing launched.
    bool ds_pair(const Act& in, const std::string& p, const Act* out, const EpiOpt& o, const std::string* gdn) {
        if (!(bf && use_tc && fuse && pair) || !out || o.out2 || o.out_f32 || o.nchw || o.ck || o.premask || o.postmask) return false;
        const DwW* d = dw(p + ".depth_conv");
        const ConvW* w = cw(p + ".point_conv");
        const ConvW* g = gdn ? cw(*gdn) : nullptr;
        if (!d || !w || (gdn && !g)) return false;
        if (d->C != in.C || w->Cin != in.C || w->N != in.C || w->Cpad != in.C || w->ks != 1 || w->shuffle || out->C != in.C) return false;
        if (g && (g->Cin != in.C || g->N != in.C || g->Cpad != in.C || g->ks != 1)) return false;
        if (!gdn && o.gdn) return false;
        if (gdn && (!o.gdn || o.act != ACT_NONE)) return false;
        DsPairArgs a;
        memset(&a, 0, sizeof a);
        a.in = in.p; a.B = in.B; a.H = in.H; a.W = in.W; a.ld = in.ld; a.C = in.C;
        a.dw_w9 = d->w9; a.dw_bias = d->bias; a.w1 = w->wbf; a.b1 = w->bias; a.act = o.act;
        a.gdn = gdn ? o.gdn : GDN_NONE;
        if (g) { a.w2 = g->wbf; a.b2 = g->bias; }
        if (o.res) { a.res = o.res->p; a.res_ld = o.res->ld; }
        a.out = out->p; a.out_ld = out->ld;
        if (!dry && !ds_pair_supported(a)) return false;
        if (dry) return (in.C == 192 || in.C == 128) && (in.ld % 8) == 0 && (out->ld % 8) == 0 && (!o.res || (o.res->ld % 8) == 0);
        if (rc) return true;
        cudaEvent_t ev1 = nullptr;
        if (profile) {
            cudaEventRecord(next_event(), st);
            ev1 = next_event();
            ev_flops.push_back((gdn ? 4.0 : 2.0) * (double)in.B * in.H * in.W * (double)in.C * (double)in.C);
        }
        const int r = launch_ds_pair(a, st);
        if (ev1) cudaEventRecord(ev1, st);
        if (r) { if (!rc) rc = fail("two-SM block '%s': %s", p.c_str(), ds_pair_last_error()); return true; }
        ++launches;
        if (trace) {
            char lab[256];
            snprintf(lab, sizeof lab, "%s [pair %s M=%d C=%d]", p.c_str(), gdn ? "dw+pw+gdn" : "dw+pw", in.B * in.H * in.W, in.C);
            tr(lab);
        }
        return true;
    }
    // DepthWiseConv (modules/layers/conv.py:46-63): dw3x3(stride) -> pw1x1 with epilogue
    void dsconv(const Act& in, const std::string& p, int stride, const Act* out, const EpiOpt& o) {
        if (stride == 1 && !o.gdn && ds_pair(in, p, out, o, nullptr)) return;
        if (bf && use_tc && fuse && stride == 1 && !o.out2) {
            const DwW* d = dw(p + ".depth_conv");
            if (d && d->C == in.C && gemm(in, p + ".point_conv", 1, 0, out, o, 1, d)) return;
        }
        size_t mark = ws_off;
        Act t = act(in.B, (in.H - 1) / stride + 1, (in.W - 1) / stride + 1, in.C);
        dwconv(in, p + ".depth_conv", stride, ACT_NONE, t);
        gemm(t, p + ".point_conv", 1, 0, out, o);
        ws_off = mark;
    }
    void c3(const Act& in, const std::string& p, int stride, bool dense, const Act* out, const EpiOpt& o) {
        if (dense) gemm(in, p, stride, 1, out, o);
        else dsconv(in, p, stride, out, o);
    }
    int out_ch(const std::string& p, bool dense) {
        const ConvW* w = cw(dense ? p : p + ".point_conv");
        return w ? w->N : 0;
    }
    void layernorm(const Act& x, const std::string& p, const Act& out) {
        auto it = lns.find(p);
        if (it == lns.end()) { if (!rc) rc = fail("layernorm '%s' not packed", p.c_str()); return; }
        if (!go()) return;
        launch_layernorm(bf, x, it->second.g, it->second.b, out, st);
        after_launch(p.c_str());
    }

    // ------------------------------------------------------------------ blocks
    // v = conv(t); out = (I)GDN(v) [+ res]: norm = conv1x1(v^2, gamma) + beta.  Fast path: the GDN GEMM squares its A
    // operand on chip (PROD_SQ); otherwise conv's epilogue writes v^2 as a side tensor for a plain GEMM.
    void gdn_block(const Act& t, const std::string& conv, bool dense, const Act& v, const std::string& gdn, const Act& out,
                   const EpiOpt& og) {
        if (!dense && ds_pair(t, conv, &out, og, &gdn)) return;      // conv -> (I)GDN -> + res in one two-SM kernel, v stays on chip
        if (bf && use_tc && fuse) {
            size_t mark = ws_off;
            c3(t, conv, 1, dense, &v, EpiOpt());
            if (gemm(v, gdn, 1, 0, &out, og, 2)) { ws_off = mark; return; }
            // (not taken by the fused kernel: recompute through the side tensor below; conv is re-run, rare path)
            ws_off = mark;
        }
        Act sq = act(v.B, v.H, v.W, v.C);
        EpiOpt o2; o2.out2 = &sq;
        c3(t, conv, 1, dense, &v, o2);
        gemm(sq, gdn, 1, 0, &out, og);
    }
    // ResidualBlock (res_blk.py:142-154): GELU(conv2(GELU(conv1 x))) + skip(x)
    void rb(const Act& x, const std::string& p, bool dense, const Act& out) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        Act t = act(x.B, x.H, x.W, out.C);
        c3(x, p + ".conv1", 1, dense, &t, g);
        EpiOpt g2; g2.act = ACT_GELU;
        Act idn = x;
        if (convs.count(p + ".skip")) {
            idn = act(x.B, x.H, x.W, out.C);
            gemm(x, p + ".skip", 1, 0, &idn, EpiOpt());
        }
        g2.res = &idn;
        c3(t, p + ".conv2", 1, dense, &out, g2);
        ws_off = mark;
    }
    // ResidualBlockWithStride (res_blk.py:82-93): GDN(conv2(GELU(conv1_s2 x))) + skip_1x1_s2(x)
    // x_nchw != null: x is the fp32 NCHW image itself (3 channels) and conv1 + skip run as the fused head kernel.
    bool head_fused() const { return bf && use_tc && fuse && !sd; }
    void rbws(const Act& x, const std::string& p, bool dense, const Act& out, const float* x_nchw = nullptr, bool head = false) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        Act t = act(out.B, out.H, out.W, out.C);
        Act v = act(out.B, out.H, out.W, out.C);
        Act sk = act(out.B, out.H, out.W, out.C);
        if (head) {
            const DwW* d = dw(p + ".conv1.depth_conv");
            const ConvW* w1 = cw(p + ".conv1.point_conv");
            const ConvW* wk = cw(p + ".skip");
            if (d && w1 && wk && go()) {
                if (d->C != 3 || w1->Cin != 3 || wk->Cin != 3 || w1->N != out.C || wk->N != out.C || !ga_head_supported(x.H, x.W, out.C, t, sk) || !x_nchw) {
                    if (!rc) rc = fail("g_a head: unsupported geometry");
                } else {
                    launch_ga_head(x_nchw, x.B, x.H, x.W, d->w9, d->bias, w1->w32, w1->bias, wk->w32, wk->bias, out.C, t, sk, st);
                    after_launch((p + ".head").c_str());
                }
            }
        } else {
            c3(x, p + ".conv1", 2, dense, &t, g);
            gemm(x, p + ".skip", 2, 0, &sk, EpiOpt());
        }
        EpiOpt og; og.gdn = GDN_FWD; og.gdn_x = &v; og.res = &sk;
        gdn_block(t, p + ".conv2", dense, v, p + ".gdn", out, og);
        ws_off = mark;
    }
    // ResidualBlockUpsample (res_blk.py:113-121): IGDN(conv(GELU(subpel x))) + upsample(x)
    void rbu(const Act& x, const std::string& p, const Act& out) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        Act t = act(out.B, out.H, out.W, out.C);
        gemm(x, p + ".subpel_conv.0", 1, 1, &t, g);
        Act v = act(out.B, out.H, out.W, out.C);
        // bf16 + two-SM kernels: the `+ upsample(x)` of res_blk.py:121 moves from the memory-bound tail kernel (one tensor pass less there)
        // into the epilogue of the tensor-bound upsample conv, which has the bandwidth to spare: out = IGDN(conv(t)); out += upsample(x)
        const ConvW* wu = bf && use_tc && pair ? cw(p + ".upsample.0") : nullptr;
        const bool up_adds = wu && wu->ks == 3 && wu->shuffle && wu->Cpad == wu->Cin && (wu->Cin % 64) == 0 && wu->Cin >= 64 && (wu->N % 256) == 0 &&
                             ((wu->N / 4) % 64) == 0 && (out.ld % 8) == 0 && ((uintptr_t)out.p % 16) == 0 && (x.ld % 8) == 0 && ((uintptr_t)x.p % 16) == 0;
        if (up_adds) {
            EpiOpt og; og.gdn = GDN_INV; og.gdn_x = &v;
            gdn_block(t, p + ".conv", false, v, p + ".igdn", out, og);
            EpiOpt ou; ou.res = &out;
            gemm(x, p + ".upsample.0", 1, 1, &out, ou);
            ws_off = mark;
            return;
        }
        Act up = act(out.B, out.H, out.W, out.C);
        gemm(x, p + ".upsample.0", 1, 1, &up, EpiOpt());
        EpiOpt og; og.gdn = GDN_INV; og.gdn_x = &v; og.res = &up;
        gdn_block(t, p + ".conv", false, v, p + ".igdn", out, og);
        ws_off = mark;
    }

    // g_a (transform/analysis.py:9-17): x [B,H,W,3] -> y fp32 [pix][M]
    void g_a(const Act& x, float* y32, const float* x_nchw = nullptr) {
        size_t mark = ws_off;
        const std::string p = "g_a.analysis_transform.";
        Act cur = x;
        for (int i : {0, 2, 4}) {
            Act a = act(cur.B, cur.H / 2, cur.W / 2, N);
            rbws(cur, p + std::to_string(i), sd, a, x_nchw, i == 0 && head_fused());
            Act b = act(a.B, a.H, a.W, N);
            rb(a, p + std::to_string(i + 1), sd, b);
            cur = b;
        }
        EpiOpt o; o.out_f32 = y32; o.out_f32_ld = M;
        c3(cur, p + "6", 2, sd, nullptr, o);
        ws_off = mark;
    }
    // h_a (transform/analysis.py:33-43)
    void h_a(const Act& y, const Act& z) {
        size_t mark = ws_off;
        const std::string p = "h_a.reduction.";
        const int strides[5] = {1, 1, 2, 1, 2};
        Act cur = y;
        for (int j = 0; j < 5; ++j) {
            EpiOpt o; o.act = j < 4 ? ACT_GELU : ACT_NONE;
            int s = strides[j];
            Act nx = j < 4 ? act(cur.B, (cur.H - 1) / s + 1, (cur.W - 1) / s + 1, N) : z;
            c3(cur, p + std::to_string(2 * j), s, sd, &nx, o);
            cur = nx;
        }
        ws_off = mark;
    }
    // h_s (transform/synthesis.py:18-28): z_hat -> hyper_params (written into `out`, 2*Me channels)
    void h_s(const Act& zh, const Act& out) {
        size_t mark = ws_off;
        const std::string p = "h_s.increase.";
        EpiOpt g; g.act = ACT_GELU;
        int c0 = out_ch(p + "0", false);
        Act a = act(zh.B, zh.H, zh.W, c0);
        dsconv(zh, p + "0", 1, &a, g);
        Act b = act(zh.B, zh.H * 2, zh.W * 2, c0);
        gemm(a, p + "2.0", 1, 1, &b, g);
        int c4 = out_ch(p + "4", false);
        Act c = act(b.B, b.H, b.W, c4);
        dsconv(b, p + "4", 1, &c, g);
        Act d = act(b.B, b.H * 2, b.W * 2, c4);
        gemm(c, p + "6.0", 1, 1, &d, g);
        dsconv(d, p + "8", 1, &out, EpiOpt());
        ws_off = mark;
    }
    // g_s (transform/synthesis.py:59-68): y_hat view [B,h,w,M] -> x_hat fp32 NHWC [B,H,W,3]
    void g_s(const Act& yh, float* xhat_nchw) {
        size_t mark = ws_off;
        const std::string p = "g_s.synthesis_transform.";
        Act cur = act(yh.B, yh.H, yh.W, M);
        rb(yh, p + "0", false, cur);
        for (int i : {1, 3, 5}) {
            const ConvW* w = cw(p + std::to_string(i) + ".subpel_conv.0");
            int co = w ? w->N / 4 : 0;
            Act a = act(cur.B, cur.H * 2, cur.W * 2, co);
            rbu(cur, p + std::to_string(i), a);
            Act b = act(a.B, a.H, a.W, co);
            rb(a, p + std::to_string(i + 1), false, b);
            cur = b;
        }
        EpiOpt o; o.out_f32 = xhat_nchw; o.nchw = 1;
        if (!(bf && use_tc && fuse && convs.count(p + "7.0_ss") && gemm_ss(cur, p + "7.0_ss", xhat_nchw)))
            gemm(cur, p + "7.0", 1, 1, nullptr, o);
        ws_off = mark;
    }

    // EntropyParameters (transform/entropy.py:10-29): 1x1 chain in->320->256->128->2C, GELU between; fp32 out
    // ck != 0 (bf16 fast path): only the anchor (1) / non-anchor (2) pixels are evaluated -- the stack is per-pixel and its
    // output is multiplied by that mask (mlicpp.py:112-117,148-152); out32 is then squeezed, [B*h*(w/2)][2C].
    bool ep_squeezed(int Hh, int Ww) const { return bf && use_tc && fuse && (Hh % 2) == 0 && (Ww % 2) == 0; }
    // three chained layers in one launch (chain3.cu); false: not taken, nothing launched
    bool chain3(int mode, const Act& rows, const std::string& k1, const std::string& k2, const std::string& k3, const LnW* ln, void* out, int out_ld, const char* what, int unsq_H = 0, int unsq_W = 0) {
        if (!(bf && use_tc && fuse && chain)) return false;
        const ConvW *w1 = cw(k1), *w2 = cw(k2), *w3 = cw(k3);
        if (!w1 || !w2 || !w3 || w1->ks != 1 || w2->ks != 1 || w3->ks != 1) return false;
        if (w1->Cin != rows.C || w2->Cin != w1->N || w3->Cin != w2->N || w2->Cpad != w1->N || w3->Cpad != w2->N) return false;
        Chain3Args a;
        memset(&a, 0, sizeof a);
        a.mode = mode; a.in = rows.p; a.M = rows.B * rows.H * rows.W; a.K1 = rows.C; a.ld = rows.ld;
        a.w1 = w1->wbf; a.K1pad = w1->Cpad; a.w2 = w2->wbf; a.w3 = w3->wbf; a.b1 = w1->bias; a.b2 = w2->bias; a.b3 = w3->bias;
        a.N1 = w1->N; a.N2 = w2->N; a.N3 = w3->N;
        if (ln) { a.ln_g = ln->g; a.ln_b = ln->b; a.ln_eps = 1e-5f; }
        a.out = out; a.out_ld = out_ld; a.unsq_H = unsq_H; a.unsq_W = unsq_W;
        if (dry) { a.in = a.out = (void*)16; }                     // (the dry run has no buffers: geometry only)
        if (!chain3_supported(a)) return false;
        if (!go()) return true;
        cudaEvent_t ev1 = nullptr;
        if (profile) {
            cudaEventRecord(next_event(), st);
            ev1 = next_event();
            ev_flops.push_back(2.0 * (double)a.M * ((double)a.K1 * a.N1 + (double)a.N1 * a.N2 + (double)a.N2 * a.N3));
        }
        const int r = launch_chain3(a, st);
        if (ev1) cudaEventRecord(ev1, st);
        if (r) { if (!rc) rc = fail("chain '%s': %s", k1.c_str(), chain3_last_error()); return true; }
        ++launches;
        if (trace) {
            char lab[256];
            snprintf(lab, sizeof lab, "%s [chain3 M=%d K=%d N=%d-%d-%d]", what, a.M, a.K1, a.N1, a.N2, a.N3);
            tr(lab);
        }
        return true;
    }
    void ep(const Act& in, const std::string& p, float* out32, int ck = 0) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        Act cur = in;
        const int Mh = in.B * in.H * (in.W / 2);
        if (ck && bf && use_tc && fuse && chain) {
            // squeezed rows: layer 0 as before (wide GEMM over the 5-D checkerboard gather), layers 1..3 in one launch
            const ConvW* w0 = cw(p + ".fusion.0");
            Act h0 = act(1, 1, Mh, w0 ? w0->N : 0);
            EpiOpt g0 = g; g0.ck = ck;
            const size_t mark2 = ws_off;
            (void)mark2;
            gemm(cur, p + ".fusion.0", 1, 0, &h0, g0);
            if (chain3(0, h0, p + ".fusion.2", p + ".fusion.4", p + ".fusion.6", nullptr, out32, 2 * C, (p + ".fusion.2-6").c_str())) { ws_off = mark; return; }
            cur = h0;
            for (int j : {2, 4}) {
                const ConvW* w = cw(p + ".fusion." + std::to_string(j));
                Act nx = act(1, 1, Mh, w ? w->N : 0);
                gemm(cur, p + ".fusion." + std::to_string(j), 1, 0, &nx, g);
                cur = nx;
            }
            EpiOpt o; o.out_f32 = out32; o.out_f32_ld = 2 * C;
            gemm(cur, p + ".fusion.6", 1, 0, nullptr, o);
            ws_off = mark;
            return;
        }
        for (int j : {0, 2, 4}) {
            const ConvW* w = cw(p + ".fusion." + std::to_string(j));
            Act nx = ck ? act(1, 1, Mh, w ? w->N : 0) : act(in.B, in.H, in.W, w ? w->N : 0);
            EpiOpt gj = g;
            if (ck && j == 0) gj.ck = ck;
            gemm(cur, p + ".fusion." + std::to_string(j), 1, 0, &nx, gj);
            cur = nx;
        }
        EpiOpt o; o.out_f32 = out32; o.out_f32_ld = 2 * C;
        gemm(cur, p + ".fusion.6", 1, 0, nullptr, o);
        ws_off = mark;
    }
    // LatentResidualPrediction (transform/quantization.py:30-45; :9-28 for SD): slot += mask * 0.5*tanh(stack(in))
    void lrp(const Act& in, const std::string& p, const Act& slot, int parity) {
        size_t mark = ws_off;
        const int nl = sd ? 4 : 3;
        EpiOpt g; g.act = ACT_GELU;
        Act cur = in;
        for (int j = 0; j < nl - 1; ++j) {
            std::string q = p + ".lrp_transform." + std::to_string(2 * j);
            Act nx = act(in.B, in.H, in.W, out_ch(q, false));
            dsconv(cur, q, 1, &nx, g);
            cur = nx;
        }
        EpiOpt o; o.act = ACT_HALF_TANH; o.postmask = parity; o.res = &slot;
        dsconv(cur, p + ".lrp_transform." + std::to_string(2 * (nl - 1)), 1, &slot, o);
        ws_off = mark;
    }
    // ChannelContext (transform/context.py:118-138; context_old.py:120-126 dense for SD)
    void channel_ctx(const Act& in, const std::string& p, const Act& out) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        Act a = act(in.B, in.H, in.W, out_ch(p + ".fushion.0", sd));
        c3(in, p + ".fushion.0", 1, sd, &a, g);
        Act b = act(in.B, in.H, in.W, out_ch(p + ".fushion.2", sd));
        c3(a, p + ".fushion.2", 1, sd, &b, g);
        c3(b, p + ".fushion.4", 1, sd, &out, EpiOpt());
        ws_off = mark;
    }
    // 1x1 -> GELU -> dw3x3 -> GELU -> 1x1 (+ residual)   (context.py:161-167,217-223)
    void dw_mlp(const Act& in, const std::string& p, const Act& out, const Act* res) {
        size_t mark = ws_off;
        EpiOpt g; g.act = ACT_GELU;
        const ConvW* w0 = cw(p + ".0");
        int hid = w0 ? w0->N : 0;
        Act a = act(in.B, in.H, in.W, hid), b = act(in.B, in.H, in.W, hid);
        gemm(in, p + ".0", 1, 0, &a, g);
        dwconv(a, p + ".2", 1, ACT_GELU, b);
        EpiOpt o; o.res = res;
        gemm(b, p + ".4", 1, 0, &out, o);
        ws_off = mark;
    }
    void lin_attn(const Act& qkv, int D, int heads, int hd, int par_kv, int par_q, const Act& out) {
        float* scratch = f32(lin_attn_scratch_floats(qkv.B, heads, hd, qkv.H * qkv.W));
        if (!go()) return;
        if (launch_lin_attn(bf, qkv, D, heads, hd, par_kv, par_q, scratch, out, st)) { if (!rc) rc = fail("linear attention: unsupported head dim %d", hd); return; }
        launches += 3;
        after_launch("lin_attn");
    }
    // LinearGlobalInterContext (transform/context.py:226-245)
    void inter_ctx(const Act& X, const std::string& p, const Act& out) {
        size_t mark = ws_off;
        const int D = X.C;
        Act pw = act(X.B, X.H, X.W, 3 * D), qkv = act(X.B, X.H, X.W, 3 * D);
        gemm(X, p + ".qkv_pw", 1, 0, &pw, EpiOpt());
        dwconv(pw, p + ".qkv_dw", 1, ACT_NONE, qkv);
        Act O = act(X.B, X.H, X.W, D);
        lin_attn(qkv, D, D / 32, 32, PAR_NONE, PAR_NONE, O);
        const ConvW* wr = cw(p + ".reprojection");
        const ConvW* w0 = cw(p + ".mlp.0");
        if (bf && use_tc && fuse && folds && wr && w0 && convs.count(p + ".skip_mlp4") && (wr->N % 8) == 0) {
            // skip(att) + mlp.4(h): one GEMM over [att | h], which the re-projection and the depthwise conv write side by side
            Act AH = act(X.B, X.H, X.W, wr->N + w0->N);
            Act A = view(AH, 0, wr->N), h2 = view(AH, wr->N, w0->N);
            gemm(O, p + ".reprojection", 1, 2, &A, EpiOpt());
            Act a = act(X.B, X.H, X.W, w0->N);
            EpiOpt g; g.act = ACT_GELU;
            gemm(A, p + ".mlp.0", 1, 0, &a, g);
            dwconv(a, p + ".mlp.2", 1, ACT_GELU, h2);
            gemm(AH, p + ".skip_mlp4", 1, 0, &out, EpiOpt());
            ws_off = mark;
            return;
        }
        Act A = act(X.B, X.H, X.W, wr ? wr->N : 0);
        gemm(O, p + ".reprojection", 1, 2, &A, EpiOpt());
        gemm(A, p + ".skip", 1, 0, &out, EpiOpt());
        dw_mlp(A, p + ".mlp", out, &out);
        ws_off = mark;
    }
    // LinearGlobalIntraContext (transform/context.py:169-193)
    void intra_ctx(const Act& x1, const Act& x2, const std::string& p, const Act& out) {
        size_t mark = ws_off;
        const int D = x1.C;
        Act pw = act(x1.B, x1.H, x1.W, 3 * D), qkv = act(x1.B, x1.H, x1.W, 3 * D);
        Act vq = view(pw, 0, D), vk = view(pw, D, D), vv = view(pw, 2 * D, D);
        EpiOpt oq; oq.premask = PAR_NONANCHOR;
        EpiOpt ok; ok.premask = PAR_ANCHOR;
        if (bf && use_tc && fuse && folds && convs.count(p + ".qkv_blk") && x1.ld == x2.ld && (uint8_t*)x2.p == (uint8_t*)x1.p + (size_t)D * esz()) {
            // x1 | x2 are adjacent channel ranges: q | k | v in one GEMM (block weights, premask per D-column group)
            EpiOpt ob; ob.pm_w = D; ob.pm_codes = PAR_NONANCHOR | (PAR_ANCHOR << 2) | (PAR_NONE << 4);
            Act x12 = x1; x12.C = 2 * D;
            gemm(x12, p + ".qkv_blk", 1, 0, &pw, ob);
        } else {
        gemm(x1, p + ".queries.0", 1, 0, &vq, oq);
        gemm(x1, p + ".keys.0", 1, 0, &vk, ok);
        gemm(x2, p + ".values.0", 1, 0, &vv, EpiOpt());
        }
        dwconv(pw, p + ".qkv_dw", 1, ACT_NONE, qkv);
        Act O = act(x1.B, x1.H, x1.W, D);
        lin_attn(qkv, D, 2, D / 2, PAR_ANCHOR, PAR_NONANCHOR, O);
        Act A = act(x1.B, x1.H, x1.W, 2 * D);
        gemm(O, p + ".reprojection", 1, 2, &A, EpiOpt());
        dw_mlp(A, p + ".mlp", out, &A);
        ws_off = mark;
    }
    // LocalContext (transform/context.py:67-112)
    void local_ctx(const Act& x, const std::string& p, const Act& out) {
        size_t mark = ws_off;
        const int Cc = x.C;
        const size_t npix = (size_t)x.B * x.H * x.W;
        Act t = act(x.B, x.H, x.W, Cc);
        layernorm(x, p + ".norm1", t);
        if (bf && use_tc && fuse && Cc == 32 && (x.W % 2) == 0 && convs.count(p + ".qkv_proj_hm")) {
            // bf16 fast path: attention on the warp-level tensor path for the non-anchor pixels only (the anchor half of the
            // LocalContext output is never observed, mlicpp.py:146-150); the per-pixel tail runs on the squeezed matrix.
            const int Mh = (int)(npix / 2);
            Act Fb = act(x.B, x.H, x.W, 3 * Cc);
            gemm(t, p + ".qkv_proj_hm", 1, 0, &Fb, EpiOpt());
            Act O = act(1, 1, Mh, 25 * Cc);
            if (go()) {
                if (launch_local_attn_mma(Fb, misc[p + ".rel_bias"], O.p, st)) { if (!rc) rc = fail("local attention (mma): unsupported geometry"); }
                after_launch("local_attn");
            }
            Act fu = act(1, 1, Mh, 2 * Cc), pr = act(1, 1, Mh, 2 * Cc), n2 = act(1, 1, Mh, 2 * Cc), os = act(1, 1, Mh, 2 * Cc);
            {
                auto itn = lns.find(p + ".norm2");
                if (folds && convs.count(p + ".fusion_proj") && itn != lns.end() && itn->second.C == 2 * Cc) {
                    // with squeezed EntropyParameters rows only the non-anchor pixels of `out` are ever read: the chain stores each row at its pixel
                    const bool direct = ep_squeezed(x.H, x.W) && (out.ld % 8) == 0 && ((uintptr_t)out.p % 16) == 0;
                    if (direct ? chain3(1, O, p + ".fusion_proj", p + ".mlp.fc1", p + ".mlp.fc2", &itn->second, out.p, out.ld, (p + ".tail").c_str(), x.H, x.W)
                               : chain3(1, O, p + ".fusion_proj", p + ".mlp.fc1", p + ".mlp.fc2", &itn->second, os.p, os.ld, (p + ".tail").c_str())) {
                        if (!direct && go()) { launch_unsqueeze_nonanchor(os, out, st); after_launch("local_unsqueeze"); }
                        ws_off = mark;
                        return;
                    }
                }
            }
            if (folds && convs.count(p + ".fusion_proj")) gemm(O, p + ".fusion_proj", 1, 0, &pr, EpiOpt());
            else {
            gemm(O, p + ".fusion", 1, 0, &fu, EpiOpt());
            gemm(fu, p + ".proj", 1, 0, &pr, EpiOpt());
            }
            layernorm(pr, p + ".norm2", n2);
            Act h = act(1, 1, Mh, 4 * Cc);
            EpiOpt g; g.act = ACT_GELU;
            gemm(n2, p + ".mlp.fc1", 1, 0, &h, g);
            EpiOpt o; o.res = &pr;
            gemm(h, p + ".mlp.fc2", 1, 0, &os, o);
            if (go()) { launch_unsqueeze_nonanchor(os, out, st); after_launch("local_unsqueeze"); }
            ws_off = mark;
            return;
        }
        float* F = f32(npix * 3 * Cc);
        EpiOpt of; of.out_f32 = F; of.out_f32_ld = 3 * Cc;
        gemm(t, p + ".qkv_proj", 1, 0, nullptr, of);
        Act O = act(x.B, x.H, x.W, 25 * Cc);
        if (go()) {
            if (launch_local_attn(bf, F, x.B, x.H, x.W, Cc, misc[p + ".rel_bias"], O.p, st)) { if (!rc) rc = fail("local attention: unsupported slice width %d", Cc); }
            after_launch("local_attn");
        }
        Act fu = act(x.B, x.H, x.W, 2 * Cc), pr = act(x.B, x.H, x.W, 2 * Cc), n2 = act(x.B, x.H, x.W, 2 * Cc);
        gemm(O, p + ".fusion", 1, 0, &fu, EpiOpt());
        gemm(fu, p + ".proj", 1, 0, &pr, EpiOpt());
        layernorm(pr, p + ".norm2", n2);
        Act h = act(x.B, x.H, x.W, 4 * Cc);
        EpiOpt g; g.act = ACT_GELU;
        gemm(n2, p + ".mlp.fc1", 1, 0, &h, g);
        EpiOpt o; o.res = &pr;
        gemm(h, p + ".mlp.fc2", 1, 0, &out, o);
        ws_off = mark;
    }

    // host-buffer calls: images per upload / g_a / g_s / download group (copies of group k overlap the kernels of group k -+ 1; one
    // image per group keeps the copy engines busiest but runs the transforms at their least efficient launch size)
    static int pipe_group(int B) {
        static const int forced = getenv("MLIC_PIPE_GROUP") ? atoi(getenv("MLIC_PIPE_GROUP")) : 0;
        if (forced > 0) return std::min(forced, B);
        return B >= 16 ? 4 : (B >= 8 ? 2 : 1);
    }
    // Image groups of the host-buffer pipeline: `pg` images per group, but the groups at the exposed end of the pipe -- the first uploads
    // (g_a waits for them) or the last downloads (nothing runs behind them) -- ramp 1, 1, 2, ...: the copy that nothing overlaps is one
    // image, not four (32 images: ~3 ms of a 112 ms step).  MLIC_PIPE_RAMP=0: uniform groups.
    static std::vector<std::pair<int, int>> pipe_groups(int B, bool ramp_front) {
        static const int ramp = getenv("MLIC_PIPE_RAMP") ? atoi(getenv("MLIC_PIPE_RAMP")) : 1;
        const int pg = pipe_group(B);
        std::vector<int> sizes;
        int left = B;
        if (ramp && pg >= 4 && B >= 4 * pg) { for (int r : {1, 1, 2}) { sizes.push_back(r); left -= r; } }
        while (left > 0) { const int n = std::min(pg, left); sizes.push_back(n); left -= n; }
        if (!ramp_front) std::reverse(sizes.begin(), sizes.end());
        std::vector<std::pair<int, int>> g;
        int b = 0;
        for (int n : sizes) { g.emplace_back(b, n); b += n; }
        return g;
    }

    // ------------------------------------------------------------------ the whole call
    int run(int mode, int precision, int B, int H, int W, float gain, const mlic_buffers* io, void* ws, size_t ws_bytes,
            cudaStream_t stream, bool dry_run, const HostPipe* hp = nullptr) {
        if (!finalized) return fail("engine not finalized");
        if (mode < 0 || mode > 3) return fail("bad mode %d", mode);
        const bool decomp = mode == MLIC_MODE_DECOMPRESS;
        if (decomp && (stages != 7 || hp)) return fail("decompress runs all stages on device buffers");
        // stage subsets (row bands): g_a alone ends at the `y` tap; without g_a `y` is an INPUT; g_s alone reads `y_hat`
        const int stg = ((mode == MLIC_MODE_DECODER || decomp) && stages != 4) ? (stages | 1) : stages;   // the decoder walk has no g_a
        if (stg != 7 && stg != 1 && stg != 6 && stg != 2 && stg != 4 && stg != 3) return fail("bad stage mask %d", stages);
        if (stg != 7 && hp) return fail("stage subsets take device buffers (mlic_run)");
        const int hmul = (stg & 2) ? 64 : 16;         // a band of g_a / g_s rows only has to keep the 16x sampling phase
        if (B <= 0 || H <= 0 || W <= 0 || (H % hmul) || (W % 64)) return fail("B=%d H=%d W=%d: H must be a positive multiple of %d and W of 64", B, H, W, hmul);
        bf = precision == MLIC_PREC_BF16;
        dry = dry_run; st = stream; rc = 0; launches = 0;
        if (hp && !dry) pipe_used = 0;
        if (!dry && profile && ev_used > 200000) { if (profile_collect()) return 1; }
        ws_base = (uint8_t*)ws; ws_size = ws_bytes; ws_off = 0; ws_peak = 0;
        if (trace) tr("start");
        if (!dry && ((uintptr_t)ws % 256)) return fail("workspace must be 256-byte aligned");
        if (!dry && bf && use_tc && tc_init()) return fail("%s", tc_last_error());
        static const mlic_buffers none = {};
        if (!io) io = &none;
        const int h = H / 16, w = W / 16, hz = H / 64, wz = W / 64;
        const size_t npix = (size_t)B * h * w;
        const int use_gain = (vbr && gain != 0.0f) ? 1 : 0;
        const float rgain = use_gain ? 1.0f / gain : 1.0f;

        Act LRPW = act(B, h, w, Me + M);
        Act EPW = act(B, h, w, 10 * C + 2 * Me);
        float* y32 = f32(npix * M);
        float* lik = mode == MLIC_MODE_FORWARD ? f32(npix * M) : nullptr;
        float* pa = f32(npix * 2 * C);
        float* pn = f32(npix * 2 * C);
        Act zh = act(B, hz, wz, N);
        Act hyper = view(EPW, 10 * C, 2 * Me);
        Act yhat = view(LRPW, Me, M);

        if (stg == 4) {                                   // g_s alone: the quantised latent comes from the caller
            if (!dry && !io->y_hat) return fail("stage g_s alone needs y_hat as input");
            if (go()) { launch_nchw_to_nhwc(bf, io->y_hat, yhat, M, st); after_launch("y_hat_in"); }
            g_s(yhat, io->x_hat);
            return rc;
        }
        if (decomp) {                                     // z_hat from the decoded z symbols (EntropyBottleneck.decompress)
            if (!dry && !io->z_symbols) return fail("decompress needs z_symbols");
            if (go()) { launch_zsym_to_zhat(bf, io->z_symbols, eb_medians, zh, st, z_qstep); after_launch("z_hat_in"); }
        } else if (mode != MLIC_MODE_DECODER) {
            size_t mark = ws_off;
            if (!(stg & 1)) {                             // g_a ran elsewhere (its row bands were gathered): y is an input
                if (!dry && !io->y) return fail("stages without g_a need y as input");
                Act yin; yin.p = y32; yin.B = B; yin.H = h; yin.W = w; yin.C = M; yin.ld = M;
                if (go()) { launch_nchw_to_nhwc(0, io->y, yin, M, st); after_launch("y_in"); }
            } else if (!dry && !io->x) { return fail("x is NULL");
            } else if (hp && hp->hx && head_fused() && !dry) {
                // host call: upload in groups of `pg` images, g_a on a group as soon as it has landed
                const size_t img = (size_t)3 * H * W;
                const auto groups = pipe_groups(B, true);
                std::vector<cudaEvent_t> ev;
                for (const auto& gr : groups) {
                    const int b = gr.first, nb = gr.second;
                    cudaMemcpyAsync(const_cast<float*>(io->x) + b * img, hp->hx + b * img, nb * img * 4, cudaMemcpyHostToDevice, hp->s_in);
                    ev.push_back(pipe_event());
                    cudaEventRecord(ev.back(), hp->s_in);
                }
                for (size_t gi = 0; gi < groups.size() && !rc; ++gi) {
                    const int b = groups[gi].first, nb = groups[gi].second;
                    cudaStreamWaitEvent(st, ev[gi], 0);
                    Act x; x.p = nullptr; x.B = nb; x.H = H; x.W = W; x.C = 3; x.ld = 3;
                    g_a(x, y32 + (size_t)b * h * w * M, io->x + b * img);
                }
            } else if (head_fused()) {            // the head kernel reads the NCHW image directly
                if (hp && hp->hx && !dry) cudaMemcpyAsync(const_cast<float*>(io->x), hp->hx, (size_t)B * 3 * H * W * 4, cudaMemcpyHostToDevice, st);
                Act x; x.p = nullptr; x.B = B; x.H = H; x.W = W; x.C = 3; x.ld = 3;
                g_a(x, y32, io->x);
            } else {
                if (hp && hp->hx && !dry) cudaMemcpyAsync(const_cast<float*>(io->x), hp->hx, (size_t)B * 3 * H * W * 4, cudaMemcpyHostToDevice, st);
                Act x = act(B, H, W, 3);
                if (go()) { launch_nchw_to_nhwc(bf, io->x, x, 3, st); after_launch("nchw_to_nhwc"); }
                g_a(x, y32);
            }
            ws_off = mark;
            if ((stg & 1) && io->y && go()) { launch_nhwc_f32_to_nchw(y32, M, B, h, w, M, io->y, st); after_launch("y_tap"); }
            if (!(stg & 2)) return rc;                    // g_a alone
            Act ya = act(B, h, w, M);
            if (go()) { launch_copy_f32_to_act(bf, y32, M, ya, st); after_launch("copy_y"); }
            Act z = act(B, hz, wz, N);
            h_a(ya, z);
            if (go()) {
                launch_entropy_bottleneck(bf, z, zh, eb_packed, eb_medians, mode == MLIC_MODE_FORWARD ? io->z_likelihoods : nullptr,
                                          mode == MLIC_MODE_COMPRESS ? io->z_symbols : nullptr, st, z_qstep);
                after_launch("entropy_bottleneck");
            }
            ws_off = mark;
        } else if (go()) {
            launch_fill_zero(zh.p, (size_t)B * hz * wz * zh.ld * esz(), st);      // z_hat = 0 (mlicpp.py:390-394)
        }
        h_s(zh, hyper);
        if (go()) { launch_copy_channels(bf, view(EPW, 10 * C + Me, Me), view(LRPW, 0, Me), st); after_launch("copy_hyper_means"); }

        const size_t half = (size_t)B * C * h * (w / 2);
        int32_t* d_mail = decomp ? (int32_t*)ws_alloc(2 * half * sizeof(int32_t)) : nullptr;      // device [indexes | symbols]
        if (decomp && !dry) {
            if (!rans) return fail("decompress: no range decoder attached");
            if (cdf_tab.empty()) return fail("decompress: mlic_engine_set_cdf was not called");
            if (h_mail_n < 2 * half) {
                if (h_mail) cudaFreeHost(h_mail);
                h_mail = nullptr; h_mail_n = 0;
                if (cudaMallocHost((void**)&h_mail, 2 * half * sizeof(int32_t)) != cudaSuccess) return fail("cudaMallocHost failed");
                h_mail_n = 2 * half;
            }
        }
        // one half-slice of decompress_anchor / decompress_nonanchor (utils/ckbd.py:195-229): index list to the host, the range
        // decoder reads `half` symbols, symbols back, y_hat = symbols (* 1/gain) + means on this parity
        auto decode_half = [&](QuantArgs q, bool anchor) {
            q.mode = 3; q.idx = d_mail; q.sym = d_mail + half;
            if (go()) { if (anchor) launch_quant_anchor(bf, q, st); else launch_quant_nonanchor(bf, q, st); after_launch(anchor ? "index_anchor" : "index_nonanchor"); }
            if (!dry && !rc) {
                cudaMemcpyAsync(h_mail, d_mail, half * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
                if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail("decompress: %s", cudaGetErrorString(cudaGetLastError())); return; }
                const int r = mlic_rans_decode_stream(rans, h_mail, half, cdf_tab.data(), cdf_stride, cdf_sizes.data(), cdf_offsets.data(),
                                                      (int)cdf_sizes.size(), h_mail + half);
                if (r) { rc = fail("range decoder failed (%d) in half-slice", r); return; }
                cudaMemcpyAsync(d_mail + half, h_mail + half, half * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            }
            q.mode = 4;
            if (go()) { if (anchor) launch_quant_anchor(bf, q, st); else launch_quant_nonanchor(bf, q, st); after_launch(anchor ? "dequant_anchor" : "dequant_nonanchor"); }
        };
        for (int i = 0; i < S && !rc; ++i) {
            const std::string is = std::to_string(i);
            Act slot = view(LRPW, Me + i * C, C);
            Act lrp_in = view(LRPW, 0, Me + (i + 1) * C);
            Act s_inter = view(EPW, 4 * C, 2 * C), s_chan = view(EPW, 6 * C, 4 * C), s_intra = view(EPW, 2 * C, 2 * C);
            Act s_local = i ? view(EPW, 0, 2 * C) : view(EPW, 8 * C, 2 * C);
            Act ep_a = i ? view(EPW, 4 * C, 6 * C + 2 * Me) : hyper;
            Act ep_n = i ? view(EPW, 0, 10 * C + 2 * Me) : view(EPW, 8 * C, 2 * C + 2 * Me);
            if (i) {
                Act prev = view(LRPW, Me, i * C);
                fork2([&] { inter_ctx(prev, "global_inter_context." + is, s_inter); }, [&] { channel_ctx(prev, "channel_context." + is, s_chan); });
            }
            const bool esq = ep_squeezed(h, w);
            ep(ep_a, "entropy_parameters_anchor." + is, pa, esq ? PAR_ANCHOR : 0);
            QuantArgs q;
            memset(&q, 0, sizeof q);
            q.y = y32 + (size_t)i * C; q.y_ld = M; q.pa = pa; q.pn = pn; q.slot = slot;
            q.B = B; q.H = h; q.W = w; q.C = C; q.mode = mode; q.vbr = use_gain; q.gain = gain; q.rgain = rgain;
            q.lik = lik ? lik + (size_t)i * C : nullptr; q.lik_ld = M;
            q.table = scale_table; q.levels = scale_levels; q.sq = esq ? 1 : 0;
            if (mode == MLIC_MODE_COMPRESS) {
                if (!dry && (!io->symbols || !io->indexes)) return fail("compress needs symbols and indexes buffers");
                q.sym = io->symbols ? io->symbols + (size_t)(2 * i) * half : nullptr;
                q.idx = io->indexes ? io->indexes + (size_t)(2 * i) * half : nullptr;
            }
            if (decomp) decode_half(q, true);
            else if (go()) { launch_quant_anchor(bf, q, st); after_launch("quant_anchor"); }
            lrp(lrp_in, "lrp_anchor." + is, slot, PAR_ANCHOR);
            if (i) fork2([&] { local_ctx(slot, "local_context." + is, s_local); },
                         [&] { intra_ctx(view(LRPW, Me + (i - 1) * C, C), slot, "global_intra_context." + is, s_intra); });
            else local_ctx(slot, "local_context." + is, s_local);
            ep(ep_n, "entropy_parameters_nonanchor." + is, pn, esq ? PAR_NONANCHOR : 0);
            if (mode == MLIC_MODE_COMPRESS) { q.sym += half; q.idx += half; }
            if (decomp) decode_half(q, false);
            else if (go()) { launch_quant_nonanchor(bf, q, st); after_launch("quant_nonanchor"); }
            lrp(lrp_in, "lrp_nonanchor." + is, slot, PAR_NONANCHOR);
        }
        if (rc) return rc;
        if (mode == MLIC_MODE_FORWARD && io->y_likelihoods && go()) {
            launch_nhwc_f32_to_nchw(lik, M, B, h, w, M, io->y_likelihoods, st);
            after_launch("lik_nchw");
        }
        if (io->y_hat && go()) { launch_nhwc_to_nchw(bf, yhat, io->y_hat, st); after_launch("y_hat_tap"); }
        if (hp && go() && mode == MLIC_MODE_FORWARD && (hp->hy_lik || hp->hz_lik)) {      // likelihoods go home while g_s runs
            cudaEvent_t ev = pipe_event();
            cudaEventRecord(ev, st);
            cudaStreamWaitEvent(hp->s_out, ev, 0);
            if (hp->hy_lik && io->y_likelihoods) cudaMemcpyAsync(hp->hy_lik, io->y_likelihoods, npix * M * 4, cudaMemcpyDeviceToHost, hp->s_out);
            if (hp->hz_lik && io->z_likelihoods) cudaMemcpyAsync(hp->hz_lik, io->z_likelihoods, (size_t)B * hz * wz * N * 4, cudaMemcpyDeviceToHost, hp->s_out);
        }
        if (hp && hp->hx_hat && io->x_hat && go()) {          // (hp implies all stages)
            const size_t img = (size_t)3 * H * W;
            for (const auto& gr : pipe_groups(B, false)) {
                if (rc) break;
                const int b = gr.first, nb = gr.second;
                Act yb = yhat; yb.B = nb; yb.p = (uint8_t*)yhat.p + (size_t)b * h * w * yhat.ld * esz();
                g_s(yb, io->x_hat + b * img);
                cudaEvent_t ev = pipe_event();
                cudaEventRecord(ev, st);
                cudaStreamWaitEvent(hp->s_out, ev, 0);
                cudaMemcpyAsync(hp->hx_hat + b * img, io->x_hat + b * img, nb * img * 4, cudaMemcpyDeviceToHost, hp->s_out);
            }
        } else if ((stg & 4) && (io->x_hat || dry)) g_s(yhat, io->x_hat);
        if (mode == MLIC_MODE_FORWARD && (stg & 4) && (io->rd_sums || dry)) {
            double* partial = (double*)ws_alloc(RD_BLOCKS * sizeof(double));
            if (go() && io->rd_sums) {
                if (!io->x_hat || !io->y_likelihoods || !io->z_likelihoods) return fail("rd_sums needs x_hat and both likelihood buffers");
                cudaMemsetAsync(io->rd_sums, 0, 2 * sizeof(double), st);
                launch_reduce(io->y_likelihoods, nullptr, (long long)npix * M, 0, partial, io->rd_sums, st);
                launch_reduce(io->z_likelihoods, nullptr, (long long)B * hz * wz * N, 0, partial, io->rd_sums, st);
                launch_reduce(io->x, io->x_hat, (long long)B * 3 * H * W, 1, partial, io->rd_sums + 1, st);
                launches += 6;
                after_launch("rd_sums");
            }
        }
        return rc;
    }
};

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

const char* mlic_last_error(void) { return g_err; }
const char* mlic_version(void) { return "mlic_b200 0.1 (sm_100a)"; }

int mlic_engine_create(int N, int M, int slice_num, int kind, mlic_engine** out) {
    if (!out) return fail("out is NULL");
    if (N <= 0 || M <= 0 || slice_num <= 0 || M % slice_num) return fail("M must be divisible by slice_num");     // mlicpp.py:21
    if (kind < 0 || kind > 3) return fail("bad kind %d", kind);
    int C = M / slice_num;
    if (C != 32 && C != 64) return fail("slice width %d unsupported (32 or 64)", C);
    mlic_engine* e = new mlic_engine();
    e->N = N; e->M = M; e->S = slice_num; e->C = C; e->kind = kind;
    e->sd = (kind & MLIC_KIND_SD) != 0; e->vbr = (kind & MLIC_KIND_VBR) != 0;
    e->Me = e->sd ? M / 4 : M;
    *out = e;
    return 0;
}
void mlic_engine_destroy(mlic_engine* e) { delete e; }

int mlic_engine_set_param(mlic_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    if (!e || !name || ndim < 0 || ndim > 8) return fail("bad arguments");
    HostT t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
    if (n && !host_data) return fail("NULL data for '%s'", name);
    t.v.assign(host_data, host_data + n);
    e->params[name] = std::move(t);
    e->finalized = false;
    return 0;
}
int mlic_engine_finalize(mlic_engine* e) {
    if (!e) return fail("engine is NULL");
    return e->finalize();
}
int mlic_engine_set_option(mlic_engine* e, const char* name, int value) {
    if (!e || !name) return fail("bad arguments");
    if (!strcmp(name, "tensor_cores")) { e->use_tc = value; return 0; }
    if (!strcmp(name, "profile")) { e->profile = value; return 0; }
    if (!strcmp(name, "fuse")) { e->fuse = value; return 0; }
    if (!strcmp(name, "pair")) { e->pair = value; return 0; }
    if (!strcmp(name, "trace")) { e->trace = value; return 0; }
    if (!strcmp(name, "stages")) { e->stages = value & 7; return 0; }
    return fail("unknown option '%s'", name);
}

int mlic_engine_set_option_f(mlic_engine* e, const char* name, float value) {
    if (!e || !name) return fail("bad arguments");
    if (!strcmp(name, "z_qstep")) {
        if (!(value > 0.0f) || !std::isfinite(value)) return fail("z_qstep must be a positive number");
        e->z_qstep = value;
        return 0;
    }
    return fail("unknown float option '%s'", name);
}
int mlic_workspace_bytes(mlic_engine* e, int mode, int precision, int B, int H, int W, size_t* bytes) {
    if (!e || !bytes) return fail("bad arguments");
    // the dry walk costs ~1 ms of host time: remember its result per call geometry (finalize() clears the cache)
    const std::array<int, 8> key = {mode, precision, B, H, W, e->use_tc, e->fuse + 2 * e->pair, e->stages};
    auto it = e->ws_cache.find(key);
    if (it != e->ws_cache.end()) { *bytes = it->second; return 0; }
    int r = e->run(mode, precision, B, H, W, 1.0f, nullptr, nullptr, 0, nullptr, true);
    if (r) return r;
    *bytes = e->ws_peak + 256;
    e->ws_cache[key] = *bytes;
    return 0;
}
int mlic_run(mlic_engine* e, int mode, int precision, int B, int H, int W, float gain, const mlic_buffers* dev,
             void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!e) return fail("engine is NULL");
    return e->run(mode, precision, B, H, W, gain, dev, workspace, workspace_bytes, (cudaStream_t)cuda_stream, false);
}
int64_t mlic_last_launch_count(const mlic_engine* e) { return e ? e->launches : 0; }

int mlic_engine_set_cdf(mlic_engine* e, const int32_t* cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                        int n_tables) {
    if (!e || !cdfs || !cdf_sizes || !offsets || cdf_stride <= 0 || n_tables <= 0) return fail("bad arguments");
    for (int t = 0; t < n_tables; ++t)
        if (cdf_sizes[t] < 2 || cdf_sizes[t] > cdf_stride) return fail("cdf table %d: size %d does not fit stride %d", t, cdf_sizes[t], cdf_stride);
    e->cdf_tab.assign(cdfs, cdfs + (size_t)n_tables * cdf_stride);
    e->cdf_sizes.assign(cdf_sizes, cdf_sizes + n_tables);
    e->cdf_offsets.assign(offsets, offsets + n_tables);
    e->cdf_stride = cdf_stride;
    return 0;
}

int mlic_decompress(mlic_engine* e, int precision, int B, int H, int W, float gain, const uint8_t* y_stream, size_t y_bytes,
                    const int32_t* z_symbols, float* x_hat, float* y_hat, void* workspace, size_t workspace_bytes,
                    void* cuda_stream) {
    if (!e) return fail("engine is NULL");
    if (!y_stream || !z_symbols || !x_hat) return fail("decompress needs y_stream, z_symbols and x_hat");
    mlic_rans_decoder* d = mlic_rans_decoder_create(y_stream, y_bytes);
    if (!d) return fail("y stream of %zu bytes is not a range-coder stream", y_bytes);
    mlic_buffers io;
    memset(&io, 0, sizeof io);
    io.z_symbols = const_cast<int32_t*>(z_symbols); io.x_hat = x_hat; io.y_hat = y_hat;
    e->rans = d;
    const int r = e->run(MLIC_MODE_DECOMPRESS, precision, B, H, W, gain, &io, workspace, workspace_bytes, (cudaStream_t)cuda_stream, false);
    e->rans = nullptr;
    mlic_rans_decoder_destroy(d);
    return r;
}

int mlic_profile_read(mlic_engine* e, double* out3, int reset) {
    if (!e || !out3) return fail("bad arguments");
    int r = e->profile_collect();
    if (r) return r;
    out3[0] = e->prof_ms; out3[1] = e->prof_flops; out3[2] = e->prof_launches;
    if (reset) { e->prof_ms = e->prof_flops = e->prof_launches = 0; }
    return 0;
}
int mlic_profile_read_top(mlic_engine* e, double* out3, int reset) {
    if (!e || !out3) return fail("bad arguments");
    int r = e->profile_collect();
    if (r) return r;
    out3[0] = e->top_ms; out3[1] = e->top_flops; out3[2] = e->top_n;
    if (reset) { e->top_ms = e->top_flops = e->top_n = 0; }
    return 0;
}

int mlic_trace_dump(mlic_engine* e, const char* path) {
    if (!e || !path) return fail("bad arguments");
    return e->trace_dump(path);
}

int mlic_run_host(mlic_engine* e, int mode, int precision, int B, int H, int W, float gain, const mlic_buffers* host,
                  int pinned) {
    if (!e || !host) return fail("bad arguments");
    (void)pinned;
    size_t need = 0;
    int r = mlic_workspace_bytes(e, mode, precision, B, H, W, &need);
    if (r) return r;
    if (!e->h_stream) CUDA_OK(cudaStreamCreateWithFlags(&e->h_stream, cudaStreamNonBlocking));
    if (!e->h_in) CUDA_OK(cudaStreamCreateWithFlags(&e->h_in, cudaStreamNonBlocking));
    if (!e->h_out) CUDA_OK(cudaStreamCreateWithFlags(&e->h_out, cudaStreamNonBlocking));
    if (need > e->h_ws_bytes) {
        if (e->h_ws) cudaFree(e->h_ws);
        e->h_ws = nullptr; e->h_ws_bytes = 0;
        CUDA_OK(cudaMalloc(&e->h_ws, need));
        e->h_ws_bytes = need;
    }
    const int h = H / 16, w = W / 16, hz = H / 64, wz = W / 64;
    const size_t n_x = (size_t)B * 3 * H * W, n_y = (size_t)B * e->M * h * w, n_z = (size_t)B * e->N * hz * wz;
    const size_t n_sym = (size_t)2 * e->S * B * e->C * h * (w / 2);
    // staging layout (bytes, 256-aligned): x | x_hat | y_lik | z_lik | sym | idx | z_sym | y | y_hat | rd
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    size_t o_x = take(n_x * 4), o_xh = take(n_x * 4), o_yl = take(n_y * 4), o_zl = take(n_z * 4), o_sym = take(n_sym * 4),
           o_idx = take(n_sym * 4), o_zs = take(n_z * 4), o_y = take(n_y * 4), o_yh = take(n_y * 4), o_rd = take(16);
    if (off > e->h_io_bytes) {
        if (e->h_io) cudaFree(e->h_io);
        e->h_io = nullptr; e->h_io_bytes = 0;
        CUDA_OK(cudaMalloc(&e->h_io, off));
        e->h_io_bytes = off;
    }
    uint8_t* d = (uint8_t*)e->h_io;
    cudaStream_t s = e->h_stream;
    mlic_buffers dev;
    memset(&dev, 0, sizeof dev);
    HostPipe hp;
    hp.s_in = e->h_in; hp.s_out = e->h_out;
    if (mode != MLIC_MODE_DECODER) {
        if (!host->x) return fail("x is NULL");
        hp.hx = host->x;                  // uploaded inside the walk (image by image on the copy stream)
        dev.x = (const float*)(d + o_x);
    }
    hp.hx_hat = host->x_hat;
    if (mode == MLIC_MODE_FORWARD) { hp.hy_lik = host->y_likelihoods; hp.hz_lik = host->z_likelihoods; }
    const char* pe = getenv("MLIC_HOST_PIPE");       // development switch: 0 = copies and walk strictly one after the other
    const bool piped = !(pe && atoi(pe) == 0);
    if (!piped && hp.hx) CUDA_OK(cudaMemcpyAsync(d + o_x, host->x, n_x * 4, cudaMemcpyHostToDevice, s));
    const bool want_rd = host->rd_sums && mode == MLIC_MODE_FORWARD;
    if (host->x_hat || want_rd) dev.x_hat = (float*)(d + o_xh);
    if (mode == MLIC_MODE_FORWARD) {
        if (host->y_likelihoods || want_rd) dev.y_likelihoods = (float*)(d + o_yl);
        if (host->z_likelihoods || want_rd) dev.z_likelihoods = (float*)(d + o_zl);
        if (want_rd) dev.rd_sums = (double*)(d + o_rd);
    }
    if (mode == MLIC_MODE_COMPRESS) {
        dev.symbols = (int32_t*)(d + o_sym); dev.indexes = (int32_t*)(d + o_idx);
        if (host->z_symbols) dev.z_symbols = (int32_t*)(d + o_zs);
    }
    if (host->y) dev.y = (float*)(d + o_y);
    if (host->y_hat) dev.y_hat = (float*)(d + o_yh);
    r = e->run(mode, precision, B, H, W, gain, &dev, e->h_ws, e->h_ws_bytes, s, false, piped ? &hp : nullptr);
    if (!r && !piped) {
        if (host->x_hat) CUDA_OK(cudaMemcpyAsync(host->x_hat, dev.x_hat, n_x * 4, cudaMemcpyDeviceToHost, s));
        if (host->y_likelihoods && dev.y_likelihoods) CUDA_OK(cudaMemcpyAsync(host->y_likelihoods, dev.y_likelihoods, n_y * 4, cudaMemcpyDeviceToHost, s));
        if (host->z_likelihoods && dev.z_likelihoods) CUDA_OK(cudaMemcpyAsync(host->z_likelihoods, dev.z_likelihoods, n_z * 4, cudaMemcpyDeviceToHost, s));
    }
    if (r) { cudaStreamSynchronize(s); cudaStreamSynchronize(e->h_in); cudaStreamSynchronize(e->h_out); return r; }
    // x_hat and the likelihoods were downloaded on the copy stream inside the walk; the rest follows the walk here
    if (host->symbols && dev.symbols) CUDA_OK(cudaMemcpyAsync(host->symbols, dev.symbols, n_sym * 4, cudaMemcpyDeviceToHost, s));
    if (host->indexes && dev.indexes) CUDA_OK(cudaMemcpyAsync(host->indexes, dev.indexes, n_sym * 4, cudaMemcpyDeviceToHost, s));
    if (host->z_symbols && dev.z_symbols) CUDA_OK(cudaMemcpyAsync(host->z_symbols, dev.z_symbols, n_z * 4, cudaMemcpyDeviceToHost, s));
    if (host->y && dev.y) CUDA_OK(cudaMemcpyAsync(host->y, dev.y, n_y * 4, cudaMemcpyDeviceToHost, s));
    if (host->y_hat && dev.y_hat) CUDA_OK(cudaMemcpyAsync(host->y_hat, dev.y_hat, n_y * 4, cudaMemcpyDeviceToHost, s));
    if (want_rd) CUDA_OK(cudaMemcpyAsync(host->rd_sums, dev.rd_sums, 16, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    CUDA_OK(cudaStreamSynchronize(e->h_out));
    CUDA_OK(cudaStreamSynchronize(e->h_in));
    return 0;
}

int mlic_conv2d_nhwc(int precision, int tensor_cores, const void* in, int B, int H, int W, int Cin, const float* weight,
                     const float* bias, int N, int ks, int stride, int pad, int act, int shuffle, const void* residual,
                     void* out, int iters, float* avg_ms, void* cuda_stream) {
    if (!in || !weight || !out || iters < 1) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0;
    e.rc = 0;
    e.pack_conv_raw("w", weight, bias, N, Cin, ks, shuffle);
    if (e.rc) return e.rc;
    e.bf = precision == MLIC_PREC_BF16; e.use_tc = tensor_cores != 0; e.pair = tensor_cores >= 2; e.halo5 = tensor_cores == 2 ? 1 : tensor_cores == 3 ? 3 : 0; e.dry = false; e.st = (cudaStream_t)cuda_stream;      // 2: the two-SM kernels where they apply
    if (e.bf && e.use_tc && tc_init()) return fail("%s", tc_last_error());
    Act a; a.p = const_cast<void*>(in); a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin;
    const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
    Act o; o.p = out; o.B = B;
    if (shuffle) { o.H = 2 * Ho; o.W = 2 * Wo; o.C = N / 4; } else { o.H = Ho; o.W = Wo; o.C = N; }
    o.ld = o.C;
    Act r = o; r.p = const_cast<void*>(residual);
    EpiOpt eo; eo.act = act; if (residual) eo.res = &r;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    e.gemm(a, "w", stride, pad, &o, eo);          // warm-up / correctness launch
    CUDA_OK(cudaEventRecord(e0, e.st));
    for (int i = 1; i < iters; ++i) e.gemm(a, "w", stride, pad, &o, eo);
    CUDA_OK(cudaEventRecord(e1, e.st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e.rc) return e.rc;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_dwconv3x3_nhwc(int precision, const void* in, int B, int H, int W, int Cc, const float* weight, const float* bias,
                        int stride, int act, void* out, int iters, float* avg_ms, void* cuda_stream) {
    if (!in || !weight || !bias || !out || iters < 1 || (stride != 1 && stride != 2)) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    HostT w, b;
    w.shape = {Cc, 1, 3, 3}; w.v.assign(weight, weight + (size_t)Cc * 9);
    b.shape = {Cc}; b.v.assign(bias, bias + Cc);
    e.params["d.weight"] = w; e.params["d.bias"] = b;
    e.pack_dw_list("d", {"d"});
    if (e.rc) return e.rc;
    e.bf = precision == MLIC_PREC_BF16; e.dry = false; e.st = (cudaStream_t)cuda_stream;
    Act a; a.p = const_cast<void*>(in); a.B = B; a.H = H; a.W = W; a.C = Cc; a.ld = Cc;
    Act o; o.p = out; o.B = B; o.H = (H - 1) / stride + 1; o.W = (W - 1) / stride + 1; o.C = Cc; o.ld = Cc;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    e.dwconv(a, "d", stride, act, o);
    CUDA_OK(cudaEventRecord(e0, e.st));
    for (int i = 1; i < iters; ++i) e.dwconv(a, "d", stride, act, o);
    CUDA_OK(cudaEventRecord(e1, e.st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e.rc) return e.rc;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_dsconv_nhwc(int precision, int fuse, const void* in, int B, int H, int W, int Cin, const float* dw_weight,
                     const float* dw_bias, const float* pw_weight, const float* pw_bias, int N, int stride, int act,
                     const void* residual, void* out, int iters, float* avg_ms, void* cuda_stream) {
    if (!in || !dw_weight || !dw_bias || !pw_weight || !out || iters < 1 || (stride != 1 && stride != 2)) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    HostT w, b;
    w.shape = {Cin, 1, 3, 3}; w.v.assign(dw_weight, dw_weight + (size_t)Cin * 9);
    b.shape = {Cin}; b.v.assign(dw_bias, dw_bias + Cin);
    e.params["d.depth_conv.weight"] = w; e.params["d.depth_conv.bias"] = b;
    e.pack_dw_list("d.depth_conv", {"d.depth_conv"});
    e.pack_conv_raw("d.point_conv", pw_weight, pw_bias, N, Cin, 1, 0);
    if (e.rc) return e.rc;
    e.bf = precision == MLIC_PREC_BF16; e.use_tc = 1; e.fuse = fuse ? 1 : 0; e.pair = fuse == 2; e.dry = false; e.st = (cudaStream_t)cuda_stream;
    if (e.bf && tc_init()) return fail("%s", tc_last_error());
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    void* ws = nullptr;
    const size_t ws_bytes = (size_t)B * Ho * Wo * ((Cin + 7) / 8 * 8) * 4 + 1024;
    CUDA_OK(cudaMalloc(&ws, ws_bytes));
    e.dev_allocs.push_back(ws);
    e.ws_base = (uint8_t*)ws; e.ws_size = ws_bytes; e.ws_off = 0;
    Act a; a.p = const_cast<void*>(in); a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin;
    Act o; o.p = out; o.B = B; o.H = Ho; o.W = Wo; o.C = N; o.ld = N;
    Act r = o; r.p = const_cast<void*>(residual);
    EpiOpt eo; eo.act = act; if (residual) eo.res = &r;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    e.dsconv(a, "d", stride, &o, eo);
    CUDA_OK(cudaEventRecord(e0, e.st));
    for (int i = 1; i < iters; ++i) e.dsconv(a, "d", stride, &o, eo);
    CUDA_OK(cudaEventRecord(e1, e.st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e.rc) return e.rc;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_ds_gdn_nhwc(int fuse, const void* in, int B, int H, int W, int Cc, const float* dw_weight, const float* dw_bias,
                     const float* pw_weight, const float* pw_bias, const float* gamma, const float* beta, int inverse,
                     const void* residual, void* out, int iters, float* avg_ms, void* cuda_stream) {
    if (!in || !dw_weight || !dw_bias || !pw_weight || !pw_bias || !gamma || !beta || !out || iters < 1) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    HostT w, b;
    w.shape = {Cc, 1, 3, 3}; w.v.assign(dw_weight, dw_weight + (size_t)Cc * 9);
    b.shape = {Cc}; b.v.assign(dw_bias, dw_bias + Cc);
    e.params["d.depth_conv.weight"] = w; e.params["d.depth_conv.bias"] = b;
    e.pack_dw_list("d.depth_conv", {"d.depth_conv"});
    e.pack_conv_raw("d.point_conv", pw_weight, pw_bias, Cc, Cc, 1, 0);
    e.pack_conv_raw("g", gamma, beta, Cc, Cc, 1, 0);         // effective (re-parametrised) gamma [C][C] and beta [C]
    if (e.rc) return e.rc;
    e.bf = 1; e.use_tc = 1; e.fuse = fuse ? 1 : 0; e.pair = fuse == 2; e.dry = false; e.st = (cudaStream_t)cuda_stream;
    if (tc_init()) return fail("%s", tc_last_error());
    void* ws = nullptr;
    const size_t ws_bytes = (size_t)B * H * W * ((Cc + 7) / 8 * 8) * 2 * 3 + 4096;
    CUDA_OK(cudaMalloc(&ws, ws_bytes));
    e.dev_allocs.push_back(ws);
    e.ws_base = (uint8_t*)ws; e.ws_size = ws_bytes; e.ws_off = 0;
    Act a; a.p = const_cast<void*>(in); a.B = B; a.H = H; a.W = W; a.C = Cc; a.ld = Cc;
    Act o = a; o.p = out;
    Act r = a; r.p = const_cast<void*>(residual);
    Act v = e.act(B, H, W, Cc);
    EpiOpt og; og.gdn = inverse ? GDN_INV : GDN_FWD; og.gdn_x = &v; if (residual) og.res = &r;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    const int64_t l0 = e.launches;
    e.gdn_block(a, "d", false, v, "g", o, og);
    if (fuse == 2 && e.launches - l0 != 1 && !e.rc) return fail("two-SM kernel did not take the layer");
    CUDA_OK(cudaEventRecord(e0, e.st));
    for (int i = 1; i < iters; ++i) e.gdn_block(a, "d", false, v, "g", o, og);
    CUDA_OK(cudaEventRecord(e1, e.st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e.rc) return e.rc;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_final_subpel(int impl, const void* in, int B, int H, int W, int Cin, const float* weight, const float* bias, float* out,
                      int iters, float* avg_ms, void* cuda_stream) {
    if (!in || !weight || !bias || !out || iters < 1 || impl < 0 || impl > 1) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    e.pack_conv_raw("w", weight, bias, 12, Cin, 3, 1);
    {
        std::vector<float> ws((size_t)108 * Cin), bs(108, 0.f);
        for (int o = 0; o < 12; ++o) {
            const int n = (o & 3) * 3 + (o >> 2);
            for (int c = 0; c < Cin; ++c)
                for (int t = 0; t < 9; ++t) ws[(size_t)(t * 12 + n) * Cin + c] = weight[((size_t)o * Cin + c) * 9 + t];
            bs[n] = bias[o];
        }
        e.pack_conv_raw("w_ss", ws.data(), bs.data(), 108, Cin, 1, 0);
    }
    if (e.rc) return e.rc;
    e.bf = 1; e.use_tc = 1; e.dry = false; e.st = (cudaStream_t)cuda_stream;
    if (tc_init()) return fail("%s", tc_last_error());
    Act a; a.p = const_cast<void*>(in); a.B = B; a.H = H; a.W = W; a.C = Cin; a.ld = Cin;
    EpiOpt o; o.out_f32 = out; o.nchw = 1;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    for (int i = 0; i < iters; ++i) {
        if (i == 1) CUDA_OK(cudaEventRecord(e0, e.st));
        if (impl == 1) { if (!e.gemm_ss(a, "w_ss", out)) return fail("shift-sum conv: unsupported geometry"); }
        else e.gemm(a, "w", 1, 1, nullptr, o);
    }
    if (iters == 1) CUDA_OK(cudaEventRecord(e0, e.st));
    CUDA_OK(cudaEventRecord(e1, e.st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e.rc) return e.rc;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_local_attn(int impl, const void* F, int B, int H, int W, const float* rel_bias, void* O, int iters, float* avg_ms,
                    void* cuda_stream) {
    if (!F || !rel_bias || !O || iters < 1 || impl < 0 || impl > 2) return fail("bad arguments");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    int r = 0;
    for (int i = 0; i < iters && !r; ++i) {
        if (i == 1) CUDA_OK(cudaEventRecord(e0, st));
        if (impl == 2) { Act f; f.p = const_cast<void*>(F); f.B = B; f.H = H; f.W = W; f.C = 96; f.ld = 96; r = launch_local_attn_mma(f, rel_bias, O, st); }
        else r = launch_local_attn(impl, (const float*)F, B, H, W, 32, rel_bias, O, st);
    }
    if (iters == 1) CUDA_OK(cudaEventRecord(e0, st));
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (r) return fail("local attention: unsupported geometry");
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_lin_attn(int precision, const void* qkv, int B, int H, int W, int D, int heads, int par_kv, int par_q, void* out, int iters,
                  float* avg_ms, void* cuda_stream) {
    if (!qkv || !out || iters < 1 || heads < 1 || D % heads || B < 1 || H < 1 || W < 1) return fail("bad arguments");
    const int hd = D / heads, bf = precision == MLIC_PREC_BF16;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    float* scratch = nullptr;
    CUDA_OK(cudaMalloc((void**)&scratch, lin_attn_scratch_floats(B, heads, hd, H * W) * sizeof(float)));
    Act q; q.p = const_cast<void*>(qkv); q.B = B; q.H = H; q.W = W; q.C = 3 * D; q.ld = 3 * D;
    Act o; o.p = out; o.B = B; o.H = H; o.W = W; o.C = D; o.ld = D;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    int r = 0;
    for (int i = 0; i < iters && !r; ++i) {
        if (i == 1) CUDA_OK(cudaEventRecord(e0, st));
        r = launch_lin_attn(bf, q, D, heads, hd, par_kv, par_q, scratch, o, st);
    }
    if (iters == 1) CUDA_OK(cudaEventRecord(e0, st));
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(scratch);
    if (r) return fail("linear attention: unsupported head dim %d", hd);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_chain3(int mode, const void* in, int M, int K1, const float* w1, const float* b1, int N1, const float* w2, const float* b2, int N2,
                const float* w3, const float* b3, int N3, const float* ln_gamma, const float* ln_beta, void* out, int iters, float* avg_ms,
                void* cuda_stream) {
    if (!in || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !out || iters < 1 || M < 1 || (mode == 1 && (!ln_gamma || !ln_beta))) return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    e.pack_conv_raw("l1", w1, b1, N1, K1, 1, 0);
    e.pack_conv_raw("l2", w2, b2, N2, N1, 1, 0);
    e.pack_conv_raw("l3", w3, b3, N3, N2, 1, 0);
    if (e.rc) return e.rc;
    const ConvW &c1 = e.convs["l1"], &c2 = e.convs["l2"], &c3 = e.convs["l3"];
    Chain3Args a;
    memset(&a, 0, sizeof a);
    a.mode = mode; a.in = in; a.M = M; a.K1 = K1; a.ld = K1;
    a.w1 = c1.wbf; a.K1pad = c1.Cpad; a.w2 = c2.wbf; a.w3 = c3.wbf; a.b1 = c1.bias; a.b2 = c2.bias; a.b3 = c3.bias;
    a.N1 = N1; a.N2 = N2; a.N3 = N3; a.out = out; a.out_ld = N3; a.ln_eps = 1e-5f;
    if (mode == 1) {
        a.ln_g = e.upload(std::vector<float>(ln_gamma, ln_gamma + N1));
        a.ln_b = e.upload(std::vector<float>(ln_beta, ln_beta + N1));
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int r = 0;
    if (c2.Cpad != N1 || c3.Cpad != N2 || !chain3_supported(a)) r = 1;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    for (int i = 0; i < iters && !r; ++i) {
        if (i == 1) CUDA_OK(cudaEventRecord(e0, st));
        r = launch_chain3(a, st);
    }
    if (iters == 1) CUDA_OK(cudaEventRecord(e0, st));
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (void* q : e.dev_allocs) cudaFree(q);
    e.dev_allocs.clear();
    if (r) return fail("chain3: %s", r == 1 && !chain3_last_error()[0] ? "unsupported layer chain" : chain3_last_error());
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_ga_head(const float* x, int B, int H, int W, const float* dw_weight, const float* dw_bias, const float* pw_weight,
                 const float* pw_bias, const float* skip_weight, const float* skip_bias, int N, void* t_out, void* skip_out,
                 int iters, float* avg_ms, void* cuda_stream) {
    if (!x || !dw_weight || !dw_bias || !pw_weight || !pw_bias || !skip_weight || !skip_bias || !t_out || !skip_out || iters < 1)
        return fail("bad arguments");
    mlic_engine e;
    e.N = e.M = e.S = e.C = 0; e.kind = 0; e.sd = e.vbr = false; e.Me = 0; e.rc = 0;
    std::vector<float> w9(27);
    for (int c = 0; c < 3; ++c) for (int t = 0; t < 9; ++t) w9[t * 3 + c] = dw_weight[c * 9 + t];
    float* d_w9 = e.upload(w9);
    float* d_db = e.upload(std::vector<float>(dw_bias, dw_bias + 3));
    float* d_w1 = e.upload(std::vector<float>(pw_weight, pw_weight + (size_t)N * 3));
    float* d_b1 = e.upload(std::vector<float>(pw_bias, pw_bias + N));
    float* d_ws = e.upload(std::vector<float>(skip_weight, skip_weight + (size_t)N * 3));
    float* d_bs = e.upload(std::vector<float>(skip_bias, skip_bias + N));
    if (e.rc) return e.rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Act t; t.p = t_out; t.B = B; t.H = H / 2; t.W = W / 2; t.C = N; t.ld = N;
    Act k = t; k.p = skip_out;
    if (!ga_head_supported(H, W, N, t, k)) return fail("g_a head: unsupported geometry");
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    launch_ga_head(x, B, H, W, d_w9, d_db, d_w1, d_b1, d_ws, d_bs, N, t, k, st);
    CUDA_OK(cudaEventRecord(e0, st));
    for (int i = 1; i < iters; ++i) launch_ga_head(x, B, H, W, d_w9, d_db, d_w1, d_b1, d_ws, d_bs, N, t, k, st);
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (avg_ms) *avg_ms = iters > 1 ? ms / (iters - 1) : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mlic_gaussian_conditional(const float* y, const float* scales, const float* means, size_t n, const float* scale_table64,
                              float* y_hat, float* lik, int32_t* sym, int32_t* idx, void* cuda_stream) {
    static float* table = nullptr;          // utils/func.py:16-19, fp32
    if (!table) {
        float tab[64];
        const float lo = logf(0.11f), hi = logf(256.0f), step = (hi - lo) / 63.0f;
        for (int k = 0; k < 64; ++k) tab[k] = expf(k < 32 ? lo + step * (float)k : hi - step * (float)(63 - k));
        CUDA_OK(cudaMalloc((void**)&table, sizeof tab));
        CUDA_OK(cudaMemcpy(table, tab, sizeof tab, cudaMemcpyHostToDevice));
    }
    launch_gc_flat(y, scales, means, n, y_hat, lik, sym, idx, scale_table64 ? scale_table64 : table, 64, (cudaStream_t)cuda_stream);
    CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
