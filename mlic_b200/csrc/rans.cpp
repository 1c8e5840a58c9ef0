// Host-side range-ANS entropy coder of the compress / decompress path.
//
// The reference hands its symbol / index lists to CompressAI's `BufferedRansEncoder.encode_with_indexes` / `flush` and
// reads them back with `RansDecoder.set_stream` / `decode_stream` (MLIC++/models/mlicpp.py:212-216,279-280,303-304;
// MLIC++/utils/ckbd.py:195-229).  compressai==1.2.6 (requirements.txt:31) is an un-vendored dependency that is absent
// from /root/reference and cannot be installed here, so this file RESTATES its published algorithm
// (compressai/cpp_exts/rans/rans_interface.cpp on top of ryg_rans' 64-bit rANS with 32-bit renormalisation):
//   * 16-bit probability precision; state in [2^31, 2^63); words are emitted backwards, so the decoder reads forwards;
//   * symbol value = symbol - offset[index]; values outside [0, max_value) are coded as the escape symbol max_value
//     (= cdf_size - 2) followed by a 4-bit bypass code of 2|v| - 1 (negative) or 2 (v - max_value) (non-negative):
//     the nibble count in unary-of-15s, then the nibbles, least significant first;
//   * pmf -> quantised CDF: round(p * 2^16), rescale to sum 2^16, steal from the smallest frequency > 1 for zero bins.
// PARITY UNPINNED against CompressAI's bytes (no copy of the package to run here); pinned by construction only:
// encode -> decode round trips, escape paths, and CDF invariants (tests/test_coder.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/mlic_b200.h"

namespace {

constexpr int kPrecision = 16;
constexpr int kBypassBits = 4;
constexpr uint32_t kMaxBypass = (1u << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31;


// x / freq without a divide (the 64-bit division is most of an encoder step): Alverson reciprocals as in ryg_rans'
// rans64.h -- rcp = ceil(2^(shift+63) / freq), shift = ceil(log2 freq); q = mulhi(x, rcp) >> (shift - 1) is exact for
// every state x < 2^63.  One entry per frequency 1..2^16, built once (tests/test_coder.py checks it against `/`).
struct Rcp {
    uint64_t rcp;
    uint32_t shift;
};
const Rcp* rcp_table() {
    static const std::vector<Rcp> table = [] {
        std::vector<Rcp> t((size_t)(1u << kPrecision) + 1);
        t[0] = {0, 0};
        t[1] = {~0ull, 0};                               // q = x - 1, folded into the bias below
        for (uint32_t f = 2; f <= (1u << kPrecision); ++f) {
            uint32_t sh = 0;
            while (f > (1u << sh)) ++sh;
            const unsigned __int128 num = ((unsigned __int128)1 << (sh + 63)) + f - 1;
            t[f] = {(uint64_t)(num / f), sh - 1};
        }
        return t;
    }();
    return table.data();
}
inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

inline void enc_put(uint64_t& x, uint32_t*& p, uint32_t start, uint32_t freq, int bits, const Rcp* rt) {
    const uint64_t x_max = ((kRansL >> bits) << 32) * freq;
    if (x >= x_max) { *--p = (uint32_t)x; x >>= 32; }
    // ((x / freq) << bits) + (x % freq) + start  ==  x + start + q * (2^bits - freq),  q = x / freq
    const Rcp r = rt[freq];
    if (freq == 1) x = x + start + ((1u << bits) - 1) + (x - 1) * (uint64_t)((1u << bits) - 1);
    else x = x + start + (mulhi64(x, r.rcp) >> r.shift) * (uint64_t)((1u << bits) - freq);
}
inline void enc_put_bits(uint64_t& x, uint32_t*& p, uint32_t val, int nbits) {
    const uint64_t freq = 1ull << (kPrecision - nbits);
    const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
    if (x >= x_max) { *--p = (uint32_t)x; x >>= 32; }
    x = (x << nbits) | val;
}

}  // namespace

struct mlic_rans_decoder {
    std::vector<uint32_t> words;
    size_t pos = 0;
    uint64_t x = 0;
    // per decode_stream call and table, built on first use: first[b] = the symbol whose interval holds the cumulative
    // value b << 8, b = 0..256 -- a 16-bit value is then searched among the few symbols of its 1/256 bucket only
    std::vector<uint16_t> first;
    std::vector<uint8_t> ready;
    bool overrun = false;      // a word was requested past the end of the stream: truncated or mismatched (tables / gain) stream
    uint32_t next() { if (pos < words.size()) return words[pos++]; overrun = true; return 0u; }      // never reads past the stream
};

extern "C" {

int mlic_pmf_to_quantized_cdf(const float* pmf, int n, int32_t* cdf_out) {
    if (!pmf || !cdf_out || n <= 0) return 1;
    for (int i = 0; i < n; ++i) if (!(pmf[i] >= 0.0f) || !std::isfinite(pmf[i])) return 2;
    std::vector<uint32_t> cdf((size_t)n + 1);
    cdf[0] = 0;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) { cdf[i + 1] = (uint32_t)std::lround((double)pmf[i] * (1 << kPrecision)); total += cdf[i + 1]; }
    if (total == 0) return 3;
    for (auto& v : cdf) v = (uint32_t)((((uint64_t)1 << kPrecision) * v) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << kPrecision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t best_freq = ~0u;
        int best = -1;
        for (int j = 0; j < n; ++j) {
            const uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < best_freq) { best_freq = f; best = j; }
        }
        if (best < 0) return 4;
        if (best < i) { for (int j = best + 1; j <= i; ++j) cdf[j]--; }
        else { for (int j = i + 1; j <= best; ++j) cdf[j]++; }
    }
    for (int i = 0; i <= n; ++i) cdf_out[i] = (int32_t)cdf[i];
    return 0;
}

size_t mlic_rans_encode_bound(size_t n) { return 4 * (2 * n + 16) + 64; }

int mlic_rans_encode(const int32_t* symbols, const int32_t* indexes, size_t n, const int32_t* cdfs, int cdf_stride,
                     const int32_t* cdf_sizes, const int32_t* offsets, int n_tables, uint8_t* out, size_t out_cap,
                     size_t* out_bytes) {
    if ((!symbols || !indexes) && n) return 1;
    if (!cdfs || !cdf_sizes || !offsets || !out || !out_bytes || cdf_stride <= 0) return 1;
    // rANS is last-in first-out: walk the symbols backwards and write the words backwards, so that the decoder reads
    // both forwards.  No intermediate symbol list: per symbol one table lookup, one reciprocal multiply; an escape is
    // its bypass nibbles (in reverse), then the escape symbol itself.
    // worst case: every symbol an 8-nibble escape.  The words are written backwards from the end of the caller's buffer when
    // it holds the worst case (mlic_rans_encode_bound does) and moved to its front afterwards; only touched pages are paid for.
    const size_t worst = 2 * n + 16;
    std::unique_ptr<uint32_t[]> tmp;
    uint32_t* end;
    if (((uintptr_t)out & 3u) == 0 && out_cap / 4 >= worst) end = (uint32_t*)out + out_cap / 4;
    else { tmp.reset(new uint32_t[worst]); end = tmp.get() + worst; }
    uint32_t* p = end;
    uint64_t x = kRansL;
    const Rcp* const rt = rcp_table();
    for (size_t i = n; i-- > 0;) {
        const int t = indexes[i];
        if (t < 0 || t >= n_tables) return 2;
        const int32_t* cdf = cdfs + (size_t)t * cdf_stride;
        const int32_t max_value = cdf_sizes[t] - 2;
        if (max_value < 0 || cdf_sizes[t] > cdf_stride) return 3;
        int32_t value = symbols[i] - offsets[t];
        if (value < 0 || value >= max_value) {
            const uint32_t raw = value < 0 ? (uint32_t)(-2 * (int64_t)value - 1) : (uint32_t)(2 * ((int64_t)value - max_value));
            value = max_value;
            int n_bypass = 0;
            while (n_bypass < 8 && (raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;      // (8 nibbles cover 32 bits: no shift by 32)
            // forward order: [escape symbol] [nibble count, in units of at most 15] [nibbles, least significant first]
            for (int j = n_bypass; j-- > 0;) enc_put_bits(x, p, (raw >> (j * kBypassBits)) & kMaxBypass, kBypassBits);
            uint32_t cnt[2];
            int nc = 0;
            int32_t val = n_bypass;
            while (val >= (int32_t)kMaxBypass) { cnt[nc++] = kMaxBypass; val -= kMaxBypass; }
            cnt[nc++] = (uint32_t)val;
            while (nc-- > 0) enc_put_bits(x, p, cnt[nc], kBypassBits);
        }
        const uint32_t start = (uint32_t)cdf[value], range = (uint32_t)(cdf[value + 1] - cdf[value]);
        if (range == 0 || range >= (1u << kPrecision)) return 4;      // zero-frequency symbol: the table was not built by pmf_to_quantized_cdf
        enc_put(x, p, start, range, kPrecision, rt);
    }
    p -= 2;
    p[0] = (uint32_t)x;
    p[1] = (uint32_t)(x >> 32);
    const size_t nbytes = (size_t)(end - p) * 4;
    *out_bytes = nbytes;
    if (nbytes > out_cap) return 5;
    memmove(out, p, nbytes);
    return 0;
}

mlic_rans_decoder* mlic_rans_decoder_create(const uint8_t* stream, size_t nbytes) {
    if (!stream || nbytes < 8 || (nbytes % 4)) return nullptr;
    mlic_rans_decoder* d = new mlic_rans_decoder();
    d->words.resize(nbytes / 4);
    memcpy(d->words.data(), stream, nbytes);
    d->x = (uint64_t)d->words[0] | ((uint64_t)d->words[1] << 32);
    d->pos = 2;
    return d;
}
void mlic_rans_decoder_destroy(mlic_rans_decoder* d) { delete d; }

int mlic_rans_decode_stream(mlic_rans_decoder* d, const int32_t* indexes, size_t n, const int32_t* cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int n_tables, int32_t* out) {
    if (!d || (!indexes && n) || !cdfs || !cdf_sizes || !offsets || (!out && n)) return 1;
    const uint32_t mask = (1u << kPrecision) - 1;
    auto get_bits = [&](int nbits) {
        uint64_t x = d->x;
        const uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
        x >>= nbits;
        if (x < kRansL) x = (x << 32) | d->next();
        d->x = x;
        return val;
    };
    constexpr int kBuckets = 256, kBucketShift = kPrecision - 8;
    if (n_tables <= 0) return n ? 2 : 0;
    d->first.resize((size_t)n_tables * (kBuckets + 1));
    d->ready.assign((size_t)n_tables, 0);                 // the caller may pass other tables on every call
    for (size_t i = 0; i < n; ++i) {
        const int t = indexes[i];
        if (t < 0 || t >= n_tables) return 2;
        const int32_t* cdf = cdfs + (size_t)t * cdf_stride;
        const int size = cdf_sizes[t];
        const int32_t max_value = size - 2;
        if (max_value < 0 || size > cdf_stride) return 3;
        uint16_t* const first = d->first.data() + (size_t)t * (kBuckets + 1);
        if (!d->ready[t]) {
            // a table is 0 = cdf[0] <= cdf[1] <= ... <= cdf[size - 1] = 2^16 (pmf_to_quantized_cdf); anything else cannot decode
            if (cdf[0] != 0 || cdf[size - 1] != (1 << kPrecision) || size - 1 > 0xffff) return 4;
            for (int j = 1; j < size; ++j) if (cdf[j] < cdf[j - 1]) return 4;
            int sidx = 0;
            for (int b = 0; b <= kBuckets; ++b) {
                const int32_t v = b << kBucketShift;
                while (sidx + 1 <= max_value && cdf[sidx + 1] <= v) ++sidx;        // largest s <= max_value with cdf[s] <= v
                first[b] = (uint16_t)sidx;
            }
            d->ready[t] = 1;
        }
        const uint32_t cum = (uint32_t)(d->x & mask);
        // the symbol s with cdf[s] <= cum < cdf[s + 1] (the last one of a run of equal entries, as a search for the first
        // entry above cum finds it)
        // (one path for every table size: a size test per symbol mispredicts on mixed streams and costs more than it saves)
        int lo = first[cum >> kBucketShift], hi = first[(cum >> kBucketShift) + 1];
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((uint32_t)cdf[mid] <= cum) lo = mid; else hi = mid - 1; }
        const int s = lo;
        const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        uint64_t x = d->x;
        x = (uint64_t)freq * (x >> kPrecision) + (x & mask) - start;
        if (x < kRansL) x = (x << 32) | d->next();
        d->x = x;
        int32_t value = s;
        if (value == max_value) {
            uint32_t val = get_bits(kBypassBits);
            int n_bypass = (int)val;
            while (val == kMaxBypass) { val = get_bits(kBypassBits); n_bypass += (int)val; }
            if (n_bypass > 8) return 5;                  // a 32-bit raw value has at most 8 nibbles: corrupt stream
            uint32_t raw = 0;
            for (int j = 0; j < n_bypass; ++j) raw |= get_bits(kBypassBits) << (j * kBypassBits);
            value = (int32_t)(raw >> 1);
            if (raw & 1) value = -value - 1;
            else value += max_value;
        }
        out[i] = value + offsets[t];
    }
    return d->overrun ? 6 : 0;       // symbols decoded from words the stream does not hold are not data
}

}  // extern "C"
