// Token ids -> text, one token at a time (host/device): what tokenizer.decode gives for the ids of one document
// (tokenizer_utils.py:82-84 SentencePiece: U+2581 -> ' ', "<0xHH>" -> the byte, the leading dummy-prefix space dropped;
// :176-179 byte-level: the token's raw bytes), cut so that a warp can check a document in parallel: every lane asks
// for the decoded LENGTH of its token, a scan gives each token its place in the text, every lane compares its bytes.
// k_roundtrip (kernels.cu) and the host emulation (tests/host_sim/sim.cpp) both run exactly these functions.
#pragma once
#include "dpt_common.h"

// token id -> [a, b) in V.tok_bytes; false for an id that is no token
DPT_HD bool dpt_tok_span(const DptVocabView& V, int32_t id, int64_t& a, int64_t& b) {
    const int32_t r = (id >= 0 && id < V.id_space) ? V.id_rank[id] : -1;
    if (r < 0) return false;
    a = V.tok_offs[r];
    b = V.tok_offs[r + 1];
    return true;
}

DPT_HD int dpt_hex_val(uint32_t c) {
    if (c >= '0' && c <= '9') return (int)c - '0';
    if (c >= 'A' && c <= 'F') return (int)c - 'A' + 10;
    return -1;
}

// "<0xHH>" (byte_fallback token of a SentencePiece vocabulary): its byte value, else -1
DPT_HD int dpt_tok_byte_value(const DptVocabView& V, int64_t a, int64_t b) {
    const uint8_t* t = V.tok_bytes + a;
    if (b - a != 6 || t[0] != '<' || t[1] != '0' || t[2] != 'x' || t[5] != '>') return -1;
    const int hi = dpt_hex_val(t[3]), lo = dpt_hex_val(t[4]);
    return (hi < 0 || lo < 0) ? -1 : hi * 16 + lo;
}

// One pass over the token's characters: decoded length, and (text != nullptr) whether text[p .. p + length) holds them
// (never reading at or beyond pe).  doc_first: the document's first decoded token - a leading space (the marker the
// normaliser prepends) is not part of the text.
DPT_HD int32_t dpt_tok_walk(const DptVocabView& V, int64_t a, int64_t b, bool spm, bool doc_first, const uint8_t* text,
                            int64_t p, int64_t pe, bool& same) {
    same = true;
    if (spm) {
        const int byte = dpt_tok_byte_value(V, a, b);
        if (byte >= 0) {
            if (text) same = p < pe && text[p] == (uint32_t)byte;
            return 1;
        }
    }
    int32_t n = 0;
    bool first_char = doc_first;
    for (int64_t q = a; q < b;) {
        uint32_t c = V.tok_bytes[q];
        int adv = 1;
        if (spm && c == DPT_MARK0 && q + 2 < b && V.tok_bytes[q + 1] == DPT_MARK1 && V.tok_bytes[q + 2] == DPT_MARK2) {
            c = 0x20u;
            adv = 3;
        }
        q += adv;
        if (spm && first_char) {
            first_char = false;
            if (c == 0x20u) continue;  // the Prepend(U+2581) marker
        }
        if (text && same) same = p + n < pe && text[p + n] == c;
        ++n;
    }
    return n;
}

DPT_HD int32_t dpt_tok_decoded_len(const DptVocabView& V, int64_t a, int64_t b, bool spm, bool doc_first) {
    bool same;
    return dpt_tok_walk(V, a, b, spm, doc_first, nullptr, 0, 0, same);
}

DPT_HD bool dpt_tok_matches(const DptVocabView& V, int64_t a, int64_t b, bool spm, bool doc_first, const uint8_t* text,
                            int64_t p, int64_t pe) {
    bool same;
    dpt_tok_walk(V, a, b, spm, doc_first, text, p, pe, same);
    return same;
}
