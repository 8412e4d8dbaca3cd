// Measurement support: synthetic corpora generated ON THE DEVICE from a counter-based
// RNG, so that a 1 GB corpus of long documents or 1.25 GB of Arabic-script text per GPU never crosses PCIe.  A document
// is a pure function of (seed, global document index): every rank can generate the same corpus, or only its own range
// of it, and the result does not depend on how the documents are spread over GPUs.
//
// Shape (the reference's inputs are S2ORC abstracts, main_analyze_s2orc.py:253-255, and the MADAR lexicon,
// dialect_arabic.py:24,53-89; dptok/synth.py is the host generator with the same surface statistics): words drawn with
// Zipf(1.0) frequencies from a lexicon uploaded by the caller (word strings + a 32-bit cumulative table), single spaces,
// sentences of geometric length ending in '.', commas, capitalised sentence starts (ASCII lexicons), 3 % numbers /
// percentages / parentheses / hyphenated pairs.  Two lexicons may be mixed per document (English / Arabic script).
//
// Not part of the tokenization path: nothing here is called by the encode entry points.
#include <cuda_runtime.h>

#include <string>

#include "../../include/dptok.h"
#include "kernels.h"

namespace dpt {

struct SynthLex {
    const uint8_t* bytes;
    const int64_t* offs;
    const uint32_t* cdf;  // cdf[k] = floor(2^32 * P(rank <= k)) (saturated), ascending
    int32_t n;
};

__device__ __forceinline__ uint64_t sy_mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ int32_t sy_draw(const SynthLex& L, uint32_t u) {  // first k with cdf[k] >= u
    int32_t lo = 0, hi = L.n - 1;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(L.cdf + mid) < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int sy_put_lex(const SynthLex& L, int32_t k, bool cap, uint8_t* out) {
    const int64_t a = __ldg(L.offs + k), b = __ldg(L.offs + k + 1);
    const int len = (int)(b - a);
    if (out) {
        for (int q = 0; q < len; ++q) out[q] = __ldg(L.bytes + a + q);
        if (cap && len > 0 && out[0] >= 'a' && out[0] <= 'z') out[0] = (uint8_t)(out[0] - 32);
    }
    return len;
}
__device__ __forceinline__ int sy_put_num(uint32_t v, uint8_t* out) {  // decimal, no leading zeros
    uint8_t d[10];
    int n = 0;
    do {
        d[n++] = (uint8_t)('0' + v % 10u);
        v /= 10u;
    } while (v);
    if (out)
        for (int q = 0; q < n; ++q) out[q] = d[n - 1 - q];
    return n;
}

__device__ __forceinline__ bool sy_sentence_end(uint64_t seed, int64_t doc, int32_t w, int32_t n_words, uint32_t sentence_mean) {
    if (w < 0) return true;
    if (w == n_words - 1) return true;
    return (uint32_t)(sy_mix(seed ^ 0xA5A5A5A5ull ^ ((uint64_t)doc << 24) ^ (uint64_t)(uint32_t)w) >> 33) % sentence_mean == 0u;
}

// bytes of word w of document `doc` (with its punctuation, without the separating space); out == nullptr: count only
__device__ int sy_word(const SynthLex& L, const dpt_synth_params& sp, int64_t doc, int32_t w, int32_t n_words, uint8_t* out) {
    const uint64_t r = sy_mix(sp.seed ^ ((uint64_t)doc * 0x100000001B3ull) ^ ((uint64_t)(uint32_t)w << 1));
    const uint64_t r2 = sy_mix(r);
    const bool ascii_lex = (sp.flags & 1) != 0;
    const bool cap = ascii_lex && sy_sentence_end(sp.seed, doc, w - 1, n_words, (uint32_t)sp.sentence_mean);
    int n = 0;
    const uint32_t kind = (uint32_t)(r2 & 0xFFFFu);
    const int32_t k = sy_draw(L, (uint32_t)(r >> 32));
    if (kind < 1966u) {  // 3 %
        const uint32_t v = (uint32_t)(r2 >> 16) & 3u;
        if (v == 0) {
            n += sy_put_num((uint32_t)(r2 >> 20) % 3000u, out ? out + n : nullptr);
        } else if (v == 1) {
            n += sy_put_num((uint32_t)(r2 >> 20) % 100u, out ? out + n : nullptr);
            if (out) out[n] = '.';
            ++n;
            n += sy_put_num((uint32_t)(r2 >> 40) % 10u, out ? out + n : nullptr);
            if (out) out[n] = '%';
            ++n;
        } else if (v == 2) {
            if (out) out[n] = '(';
            ++n;
            n += sy_put_lex(L, k, false, out ? out + n : nullptr);
            if (out) out[n] = ')';
            ++n;
        } else {
            n += sy_put_lex(L, k, cap, out ? out + n : nullptr);
            if (out) out[n] = '-';
            ++n;
            n += sy_put_lex(L, sy_draw(L, (uint32_t)(r2 >> 32)), false, out ? out + n : nullptr);
        }
    } else {
        n += sy_put_lex(L, k, cap, out ? out + n : nullptr);
        if (sp.suffix_prob) {  // redundancy sweep: make this occurrence a word of its own
            const uint64_t r3 = sy_mix(r2 ^ 0x5AFFull);
            if ((uint32_t)(r3 >> 32) < sp.suffix_prob || sp.suffix_prob == 0xFFFFFFFFu) {
                if (out)
                    for (int q = 0; q < 4; ++q) out[n + q] = (uint8_t)('a' + (uint32_t)((r3 >> (5 * q)) & 31u) % 26u);
                n += 4;
            }
        }
    }
    if (sy_sentence_end(sp.seed, doc, w, n_words, (uint32_t)sp.sentence_mean)) {
        if (out) out[n] = '.';
        ++n;
    } else if (((r2 >> 48) & 15u) == 1u) {
        if (out) out[n] = (r2 >> 52) & 7u ? ',' : ';';
        ++n;
    }
    return n;
}

__device__ __forceinline__ int32_t sy_doc_words(const dpt_synth_params& sp, int64_t doc) {
    const uint32_t span = (uint32_t)(sp.words_hi - sp.words_lo + 1);
    return sp.words_lo + (int32_t)((uint32_t)(sy_mix(sp.seed ^ 0x5EEDull ^ ((uint64_t)doc << 1)) >> 32) % span);
}
__device__ __forceinline__ bool sy_doc_uses_b(const dpt_synth_params& sp, int64_t doc) {
    return (uint32_t)(sy_mix(sp.seed ^ 0xB10Bull ^ ((uint64_t)doc * 3ull)) >> 32) < sp.frac_b;
}

// one thread per document; d_doc_offs == nullptr: lengths only
__global__ void __launch_bounds__(128) k_synth(const SynthLex A, const SynthLex B, const dpt_synth_params sp, int64_t doc_base,
                                               int64_t n_docs, int64_t* d_doc_len, const int64_t* d_doc_offs, uint8_t* d_text) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const int64_t doc = doc_base + d;
    const bool use_b = B.n > 0 && sy_doc_uses_b(sp, doc);
    dpt_synth_params q = sp;
    if (use_b) q.flags = sp.flags >> 1;  // bit 0 = lexicon A is ASCII (capitalise), bit 1 = lexicon B is
    const SynthLex& L = use_b ? B : A;
    const int32_t nw = sy_doc_words(sp, doc);
    uint8_t* out = d_doc_offs ? d_text + d_doc_offs[d] : nullptr;
    int64_t n = 0;
    for (int32_t w = 0; w < nw; ++w) {
        if (w) {
            if (out) out[n] = ' ';
            ++n;
        }
        n += sy_word(L, q, doc, w, nw, out ? out + n : nullptr);
    }
    if (!d_doc_offs) d_doc_len[d] = n;
}

static SynthLex make_lex(const uint8_t* b, const int64_t* o, const uint32_t* c, int32_t n) {
    SynthLex L;
    L.bytes = b;
    L.offs = o;
    L.cdf = c;
    L.n = n;
    return L;
}

int synth_run(const uint8_t* a_bytes, const int64_t* a_offs, const uint32_t* a_cdf, int32_t a_n, const uint8_t* b_bytes,
              const int64_t* b_offs, const uint32_t* b_cdf, int32_t b_n, const dpt_synth_params* sp, int64_t doc_base,
              int64_t n_docs, int64_t* d_doc_len, const int64_t* d_doc_offs, uint8_t* d_text, cudaStream_t st,
              std::string& err) {
    if (!a_bytes || !a_offs || !a_cdf || a_n <= 0 || !sp || n_docs <= 0 || sp->words_lo <= 0 || sp->words_hi < sp->words_lo ||
        sp->sentence_mean <= 0 || (!d_doc_offs && !d_doc_len) || (d_doc_offs && !d_text)) {
        err = "synth: bad argument";
        return DPT_EINVAL;
    }
    const SynthLex A = make_lex(a_bytes, a_offs, a_cdf, a_n);
    const SynthLex B = make_lex(b_bytes, b_offs, b_cdf, b_bytes ? b_n : 0);
    k_synth<<<(unsigned)((n_docs + 127) / 128), 128, 0, st>>>(A, B, *sp, doc_base, n_docs, d_doc_len, d_doc_offs, d_text);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("synth: ") + cudaGetErrorString(e);
        return DPT_ECUDA;
    }
    return DPT_OK;
}

}  // namespace dpt
