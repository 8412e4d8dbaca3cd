// TEST-ONLY host build of the device headers (dpt_common.h / dpt_dp_core.h / dpt_rules.h).
// Lets the CPU test-suite exercise the exact per-word code the CUDA kernels run, against the oracle,
// before any GPU time is spent.  Never linked into the product library; nothing under
// dp-tokenization_b200/ loads it.
#include <cstring>
#include <string>
#include <vector>

#include "dpt_dp_core.h"
#include "dpt_rules.h"
#include "vocab.h"

extern "C" {

void* sim_vocab_create(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n, int32_t unit_mode) {
    dpt_vocab* v = nullptr;
    std::string err;
    if (dpt_vocab_build(bytes, offs, ids, n, unit_mode, &v, err)) return nullptr;
    return v;
}
void sim_vocab_destroy(void* v) { delete (dpt_vocab*)v; }

int32_t sim_lookup(void* vv, const uint8_t* s, int32_t n) {
    dpt_vocab* v = (dpt_vocab*)vv;
    DptHashState h = dpt_hash_init(v->ph_salt);
    for (int i = 0; i < n; ++i) dpt_hash_byte(h, s[i]);
    return dpt_ph_lookup(v->h_view, h);
}

// returns number of ids (or -1 if untokenizable); *word_len = len_dp[n]
int32_t sim_word(void* vv, const uint8_t* s, int32_t n, const uint8_t* unit_starts, int32_t* out_ids, int32_t cap,
                 int32_t* word_len) {
    dpt_vocab* v = (dpt_vocab*)vv;
    std::vector<uint64_t> best(n + 1);
    std::vector<uint16_t> A(n + 1), B(n + 1);
    dpt_forward<true>(v->h_view, s, n, unit_starts, best.data(), A.data(), B.data());
    *word_len = (int32_t)dpt_key_len(best[n]);
    if (!dpt_backward_emit(v->h_view, s, n, best.data(), A.data(), B.data(), out_ids, cap)) return -1;
    return *word_len;
}

void sim_info(void* vv, int32_t* out) {
    dpt_vocab* v = (dpt_vocab*)vv;
    out[0] = v->n_tokens;
    out[1] = v->n_nodes;
    out[2] = (int32_t)v->da.size();
    out[3] = (int32_t)v->lmax;
    out[4] = (int32_t)v->ph_seed.size();
    out[5] = (int32_t)v->ph_id.size();
    out[6] = v->marker_leading_only;
    out[7] = v->byte_fallback;
}
// Sequential walk over the raw text with exactly the per-character functions k_spm_count/k_spm_write use.
// Returns normalised byte count; *n_words_out = number of words; word_offs gets n_words+1 entries.
int64_t sim_spm_normalise(void* vv, const uint8_t* text, int64_t n, const int64_t* doc_offs, int64_t n_docs, uint8_t* out,
                          int64_t out_cap, int64_t* word_offs, int64_t word_cap, int64_t* n_words_out, uint8_t* doc_flags) {
    dpt_vocab* v = (dpt_vocab*)vv;
    std::vector<uint32_t> bits(n / 32 + 2, 0u);
    for (int64_t d = 0; d < n_docs; ++d) bits[doc_offs[d] >> 5] |= 1u << (doc_offs[d] & 31);
    int64_t ob = 0, ow = 0, od = 0;
    for (int64_t p = 0; p < n; ++p) {
        if (!dpt_spm_is_char_start(text, bits.data(), p)) continue;
        const DptSpmChar ch = dpt_spm_classify(v->h_view, text, n, bits.data(), p);
        if (dpt_bit_test(bits.data(), p)) {
            if (ow < word_cap) word_offs[ow] = ob;
            if (ow + 1 < word_cap) word_offs[ow + 1] = ob + 3;
            if (ob + 6 <= out_cap) {
                const uint8_t pre[6] = {'<', 's', '>', DPT_MARK0, DPT_MARK1, DPT_MARK2};
                memcpy(out + ob, pre, 6);
            }
            if (ch.marker && doc_flags) doc_flags[od] = 1;
            ob += 6;
            ow += 2;
            od += 1;
        } else if (ch.marker) {
            if (!dpt_spm_prev_is_marker(text, bits.data(), p)) {
                if (ow < word_cap) word_offs[ow] = ob;
                ow += 1;
            } else if (doc_flags) {
                doc_flags[od - 1] = 1;
            }
        }
        if (ob + ch.out_len <= out_cap) dpt_spm_write_char(text, p, ch, out + ob);
        ob += ch.out_len;
    }
    if (ow < word_cap + 1) word_offs[ow] = ob;
    *n_words_out = ow;
    return ob;
}
}
