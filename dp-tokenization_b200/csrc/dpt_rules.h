// Pre-tokenisation rules, host/device.  Each rule restates what the reference obtains from the
// tokenizer object (tokenizer_utils.py:24-31 `pretokenize_with_llama`, :157-159 `pre_tokenize_str`);
// the oracle for these is the installed `tokenizers` 0.22.2.
#pragma once
#include "dpt_common.h"

// ---- document-start bitmap -----------------------------------------------------------------
DPT_HD bool dpt_bit_test(const uint32_t* bits, int64_t p) { return (bits[p >> 5] >> (p & 31)) & 1u; }

// ---- SPM_LLAMA ------------------------------------------------------------------------------
// HF Llama tokenizer: normaliser Prepend(U+2581) + Replace(' ', U+2581), no pre-tokenizer, BPE with
// byte_fallback; BOS '<s>' added by the post-processor.  pretokenize_with_llama then glues the
// token STRINGS back together and starts a new word at every token that starts with U+2581
// (tokenizer_utils.py:12-17), so a word is   U+2581 + non-marker run   with every character that is
// not itself a vocabulary entry replaced by the literal text "<0xHH>" of each of its bytes, and
// '<s>' is a word of its own.  This is exact whenever markers come singly (the default BPE can then
// only start a token at each marker because no vocabulary entry has a marker after a non-marker
// character - checked by the compiler as `marker_leading_only`); runs of >= 2 markers depend on the
// BPE merge order and are flagged DPT_DF_AMBIGUOUS for the caller to pre-split on the host.
struct DptSpmChar {
    int32_t src_len;   // bytes of this character in the raw text
    int32_t out_len;   // bytes it contributes to the normalised text (without doc prefix)
    bool marker;       // ' ' or U+2581
    bool literal;      // expanded to "<0xHH>" per byte
};

// p must be a character start (dpt_spm_is_char_start).  n = total bytes; doc_bits marks doc starts.
DPT_HD DptSpmChar dpt_spm_classify(const DptVocabView& V, const uint8_t* text, int64_t n, const uint32_t* doc_bits, int64_t p) {
    DptSpmChar r;
    const uint32_t c0 = text[p];
    int64_t e = p + 1;
    while (e < n && !dpt_is_cp_start(text[e]) && !dpt_bit_test(doc_bits, e)) ++e;
    r.src_len = (int32_t)(e - p);
    r.marker = (c0 == 0x20u) ||
               (r.src_len == 3 && c0 == DPT_MARK0 && text[p + 1] == DPT_MARK1 && text[p + 2] == DPT_MARK2);
    r.literal = false;
    if (r.marker) {
        r.out_len = 3;
        return r;
    }
    bool in_vocab;
    if (r.src_len == 1 && c0 < 128u) {
        in_vocab = (V.ascii_single[c0 >> 5] >> (c0 & 31)) & 1u;
    } else {
        uint32_t entry = DPT_DA_ROOT_ENTRY;
        in_vocab = true;
        for (int64_t q = p; q < e && in_vocab; ++q) in_vocab = dpt_da_step(V.da, entry, text[q]);
        in_vocab = in_vocab && (entry & DPT_DA_TERMINAL);
    }
    r.literal = !in_vocab;
    r.out_len = in_vocab ? r.src_len : 6 * r.src_len;
    return r;
}

DPT_HD bool dpt_spm_is_char_start(const uint8_t* text, const uint32_t* doc_bits, int64_t p) {
    return dpt_is_cp_start(text[p]) || dpt_bit_test(doc_bits, p);
}

// Is the character that ENDS just before p (p > doc start) a marker?
DPT_HD bool dpt_spm_prev_is_marker(const uint8_t* text, const uint32_t* doc_bits, int64_t p) {
    if (text[p - 1] == 0x20u) return true;
    if (text[p - 1] != DPT_MARK2) return false;
    if (p < 3 || dpt_bit_test(doc_bits, p - 1)) return false;
    if (text[p - 2] != DPT_MARK1 || dpt_bit_test(doc_bits, p - 2)) return false;
    return text[p - 3] == DPT_MARK0;
}

DPT_HD void dpt_spm_write_char(const uint8_t* text, int64_t p, const DptSpmChar& c, uint8_t* out) {
    if (c.marker) {
        out[0] = DPT_MARK0;
        out[1] = DPT_MARK1;
        out[2] = DPT_MARK2;
    } else if (!c.literal) {
        for (int32_t k = 0; k < c.src_len; ++k) out[k] = text[p + k];
    } else {
        for (int32_t k = 0; k < c.src_len; ++k) {
            const uint32_t b = text[p + k];
            out[6 * k + 0] = '<';
            out[6 * k + 1] = '0';
            out[6 * k + 2] = 'x';
            out[6 * k + 3] = (uint8_t)((b >> 4) < 10 ? '0' + (b >> 4) : 'A' + (b >> 4) - 10);
            out[6 * k + 4] = (uint8_t)((b & 15) < 10 ? '0' + (b & 15) : 'A' + (b & 15) - 10);
            out[6 * k + 5] = '>';
        }
    }
}
