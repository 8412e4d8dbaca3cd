// Byte-level split rules (GPT-2, Llama-3, BLOOM), host/device.
//
// The reference obtains its pieces from the tokenizer object: `pre_tokenize_str` (tokenizer_utils.py:157-159) runs
// the tokenizer's split regex over the text; each piece is then solved by the DP on its own (:165-168).  These
// functions restate the two regexes SURVEY.md 9.1 lists as sequential scanners with exactly the regex engine's
// semantics (leftmost match, alternatives tried in order, greedy quantifiers with backtracking):
//
//   GPT-2    's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
//   Llama-3  (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
//   BLOOM    Split( ?[^(\s|[.,!?…。，、।۔،])]+, isolated): the bracket expression is ONE negated class (the inner
//            [...] is a nested class), i.e. "not whitespace and none of ( ) | . , ! ? U+2026 U+3002 U+FF0C U+3001
//            U+0964 U+06D4 U+060C"; the regex does not cover the text, and `isolated` turns every maximal
//            uncovered stretch into a piece of its own (probed with tokenizers 0.22.2: "a. .b" -> a | ". ." | b)
//
// dpt_piece_end(rule, ...) returns where the piece that starts at p ends.  Code-point classes come from the table
// generated out of the installed `tokenizers` itself (tools/gen_unicode_tables.py), so \p{L}, \p{N} and \s mean what
// the reference's regex engine means by them; the case-insensitive contraction letters ('ſ' folds to 's') were
// probed the same way.  The oracle for these rules is `tokenizers`' pre_tokenize_str (tests).
//
// Parallelisation (kernel A): a space followed by a non-whitespace character of the same document is ALWAYS a
// piece start under the GPT-2 and Llama-3 regexes (whitespace alternatives never consume the last whitespace
// character in front of a non-space; only a piece that starts AT that space can take it as its optional prefix), so
// those positions and the document starts are synchronisation points: each thread scans the stretch between two
// consecutive ones.  BLOOM: the same holds for a space followed by a character of the regex's class (a match never
// runs through a space, and an uncovered stretch ends where the next match begins).
#pragma once
#include "dpt_common.h"

#define DPT_CLS_O 0u
#define DPT_CLS_L 1u
#define DPT_CLS_N 2u
#define DPT_CLS_S 3u

struct DptUniView {
    const uint8_t* stage1;  // DPT_UNI_STAGE1_LEN block indices (code point >> 8)
    const uint8_t* stage2;  // 64-byte blocks, 2 bits per code point
};

DPT_HD uint32_t dpt_cp_class(const DptUniView& U, uint32_t cp) {
    if (cp >= 0x110000u) return DPT_CLS_O;
    const uint32_t blk = U.stage1[cp >> 8];
    return (U.stage2[blk * 64u + ((cp & 255u) >> 2)] >> ((cp & 3u) * 2u)) & 3u;
}

// Decode the character that starts at text[p] (p < end).  Malformed sequences decode as a 1-byte character of
// class "other" (the reference only ever sees valid UTF-8: Python str).
struct DptChar {
    uint32_t cp;
    int32_t len;
    uint32_t cls;
};
DPT_HD DptChar dpt_char_at(const DptUniView& U, const uint8_t* text, int64_t p, int64_t end) {
    DptChar c;
    const uint32_t b0 = text[p];
    c.cp = b0;
    c.len = 1;
    if (b0 < 0x80u) {
        // ASCII: letters, digits, the six whitespace characters
        c.cls = ((b0 | 0x20u) - 'a' < 26u) ? DPT_CLS_L
                : (b0 - '0' < 10u)        ? DPT_CLS_N
                : (b0 == 0x20u || (b0 - 9u) < 5u) ? DPT_CLS_S
                                                  : DPT_CLS_O;
        return c;
    }
    int32_t need = (b0 & 0xE0u) == 0xC0u ? 2 : (b0 & 0xF0u) == 0xE0u ? 3 : (b0 & 0xF8u) == 0xF0u ? 4 : 0;
    uint32_t cp = need == 2 ? (b0 & 0x1Fu) : need == 3 ? (b0 & 0x0Fu) : (b0 & 0x07u);
    bool ok = need != 0 && p + need <= end;
    for (int32_t k = 1; ok && k < need; ++k) {
        const uint32_t bk = text[p + k];
        ok = (bk & 0xC0u) == 0x80u;
        cp = (cp << 6) | (bk & 0x3Fu);
    }
    if (!ok) {
        c.cls = DPT_CLS_O;
        return c;
    }
    c.cp = cp;
    c.len = need;
    c.cls = dpt_cp_class(U, cp);
    return c;
}

// Where the scanners get their characters from.  DptTextSrc decodes UTF-8 and looks the class up (two dependent table
// loads per multi-byte character); kernel A answers from a per-byte CODE array of its tile that one data-parallel pass
// filled (dpt_pipe.h: PaCodeSrc, dpt_char_code below) - the scanners themselves are sequential per stretch, so what a
// character costs them is what the kernel costs.  at(p, end): the character that starts at p; byte(p): raw byte.
struct DptTextSrc {
    DptUniView U;
    const uint8_t* text;
    DPT_HD DptChar at(int64_t p, int64_t end) const { return dpt_char_at(U, text, p, end); }
    DPT_HD uint32_t byte(int64_t p) const { return text[p]; }
};

DPT_HD bool dpt_bloom_excluded_cp(uint32_t cls, uint32_t cp) {
    if (cls == DPT_CLS_S) return true;
    switch (cp) {
        case '(': case ')': case '|': case '.': case ',': case '!': case '?':
        case 0x2026u: case 0x3002u: case 0xFF0Cu: case 0x3001u: case 0x0964u: case 0x06D4u: case 0x060Cu:
            return true;
        default:
            return false;
    }
}

// One byte per text position: what dpt_char_at would answer there.  bits 0-1 class, bits 2-3 length - 1, 0x10 newline
// (CR / LF), 0x20 the space U+0020, 0x40 a non-whitespace character BLOOM's class excludes.  Everything else the
// scanners ask of a code point is one of these three tests.
#define DPT_CODE_NL 0x10u
#define DPT_CODE_SP 0x20u
#define DPT_CODE_BX 0x40u
DPT_HD uint32_t dpt_char_code(const DptChar& c) {
    uint32_t code = c.cls | ((uint32_t)(c.len - 1) << 2);
    if (c.cp == 0x0Au || c.cp == 0x0Du) code |= DPT_CODE_NL;
    if (c.cp == 0x20u) code |= DPT_CODE_SP;
    if (c.cls != DPT_CLS_S && dpt_bloom_excluded_cp(c.cls, c.cp)) code |= DPT_CODE_BX;
    return code;
}
DPT_HD DptChar dpt_char_of_code(uint32_t code) {
    DptChar c;
    c.cls = code & 3u;
    c.len = (int32_t)((code >> 2) & 3u) + 1;
    c.cp = (code & DPT_CODE_NL) ? 0x0Au : (code & DPT_CODE_SP) ? 0x20u : (code & DPT_CODE_BX) ? (uint32_t)'.' : 0xFFFFu;
    return c;
}

// Optional accelerator of the scanners: ascii_letters(p, end), p a character start, returns the end of a run of whole
// characters of class L that starts at p (p itself if it knows of none), never beyond `end`; it need not be the
// maximal run.  Kernel A answers it from a one-bit-per-byte mask of its tile (dpt_pipe.h: PaLetterSkip - every byte of
// the letters of the tile); the default does nothing and the scanner walks the run character by character.
struct DptNoSkip {
    DPT_HD int64_t ascii_letters(int64_t p, int64_t) const { return p; }
};

// end of the maximal run of characters of class `cls` starting at p
template <class Src, class Skip>
DPT_HD int64_t dpt_run_end(const Src& T, int64_t p, int64_t end, uint32_t cls, const Skip& skip) {
    if (cls == DPT_CLS_L) p = skip.ascii_letters(p, end);
    while (p < end) {
        const DptChar c = T.at(p, end);
        if (c.cls != cls) break;
        p += c.len;
    }
    return p;
}

// Whitespace alternatives shared by both regexes, from a piece start p whose character is whitespace:
//   [Llama-3 only]  \s*[\r\n]+   through the last newline of the run
//   \s+(?!\S)       the run without its last character when a non-space follows, the whole run at the end
//   \s+             a single whitespace character in front of a non-space
template <class Src>
DPT_HD int64_t dpt_ws_piece_end(const Src& T, int64_t p, int64_t end, bool newline_alt) {
    int64_t e = p, last = p, after_nl = -1;
    while (e < end) {
        const DptChar c = T.at(e, end);
        if (c.cls != DPT_CLS_S) break;
        last = e;
        e += c.len;
        if (c.cp == 0x0Au || c.cp == 0x0Du) after_nl = e;
    }
    if (newline_alt && after_nl >= 0) return after_nl;
    if (e >= end) return e;
    return last > p ? last : e;
}

template <class Src, class Skip>
DPT_HD int64_t dpt_piece_end_gpt2(const Src& T, int64_t p, int64_t end, const Skip& skip) {
    const uint32_t c0 = T.byte(p);
    if (c0 == '\'' && p + 1 < end) {  // 's|'t|'re|'ve|'m|'ll|'d  (case-sensitive)
        const uint32_t c1 = T.byte(p + 1);
        if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') return p + 2;
        if (p + 2 < end) {
            const uint32_t c2 = T.byte(p + 2);
            if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) return p + 3;
        }
    }
    int64_t q = p;
    if (c0 == 0x20u && p + 1 < end) q = p + 1;  // the optional space of  ' ?X+'
    const DptChar c = T.at(q, end);
    if (c.cls != DPT_CLS_S) return dpt_run_end(T, q + c.len, end, c.cls, skip);
    return dpt_ws_piece_end(T, p, end, false);
}

DPT_HD bool dpt_is_newline(uint32_t b) { return b == 0x0Au || b == 0x0Du; }

template <class Src, class Skip>
DPT_HD int64_t dpt_piece_end_llama3(const Src& T, int64_t p, int64_t end, const Skip& skip) {
    const uint32_t c0 = T.byte(p);
    if (c0 == '\'' && p + 1 < end) {  // (?i:'s|'t|'re|'ve|'m|'ll|'d); U+017F (C5 BF) folds to 's'
        const uint32_t r1 = T.byte(p + 1);
        const uint32_t c1 = r1 | 0x20u;
        if (r1 < 0x80u && (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd')) return p + 2;
        if (p + 2 < end) {
            const uint32_t r2 = T.byte(p + 2);
            if (r1 == 0xC5u && r2 == 0xBFu) return p + 3;
            const uint32_t c2 = r2 | 0x20u;
            if (r1 < 0x80u && r2 < 0x80u && (((c1 == 'r' || c1 == 'v') && c2 == 'e') || (c1 == 'l' && c2 == 'l'))) return p + 3;
        }
    }
    const DptChar k0 = T.at(p, end);
    // [^\r\n\p{L}\p{N}]?\p{L}+
    if (k0.cls == DPT_CLS_L) return dpt_run_end(T, p + k0.len, end, DPT_CLS_L, skip);
    if (k0.cls != DPT_CLS_N && !dpt_is_newline(k0.cp) && p + k0.len < end) {
        const DptChar k1 = T.at(p + k0.len, end);
        if (k1.cls == DPT_CLS_L) return dpt_run_end(T, p + k0.len + k1.len, end, DPT_CLS_L, skip);
    }
    // \p{N}{1,3}
    if (k0.cls == DPT_CLS_N) {
        int64_t e = p + k0.len;
        for (int n = 1; n < 3 && e < end; ++n) {
            const DptChar c = T.at(e, end);
            if (c.cls != DPT_CLS_N) break;
            e += c.len;
        }
        return e;
    }
    // ' ?[^\s\p{L}\p{N}]+[\r\n]*'
    int64_t q = -1;
    if (k0.cls == DPT_CLS_O) {
        q = p;
    } else if (c0 == 0x20u && p + 1 < end) {
        const DptChar k1 = T.at(p + 1, end);
        if (k1.cls == DPT_CLS_O) q = p + 1;
    }
    if (q >= 0) {
        int64_t e = dpt_run_end(T, q, end, DPT_CLS_O, skip);
        while (e < end && dpt_is_newline(T.byte(e))) ++e;
        return e;
    }
    // whitespace: \s*[\r\n]+ | \s+(?!\S) | \s+
    return dpt_ws_piece_end(T, p, end, true);
}

// BLOOM: characters the split regex's class excludes
DPT_HD bool dpt_bloom_excluded(const DptChar& c) { return dpt_bloom_excluded_cp(c.cls, c.cp); }

template <class Src, class Skip>
DPT_HD int64_t dpt_piece_end_bloom(const Src& T, int64_t p, int64_t end, const Skip& skip) {
    const DptChar c0 = T.at(p, end);
    int64_t e = -1;  // first byte after the first class character of a match that starts at p
    if (!dpt_bloom_excluded(c0)) {
        e = p + c0.len;
    } else if (c0.cp == 0x20u && p + 1 < end) {  // the optional space
        const DptChar c1 = T.at(p + 1, end);
        if (!dpt_bloom_excluded(c1)) e = p + 1 + c1.len;
    }
    if (e >= 0) {  // a match: the run of class characters (letters are in the class: letter runs are skipped over)
        while (e < end) {
            const int64_t e2 = skip.ascii_letters(e, end);
            if (e2 >= end) return end;
            const DptChar c = T.at(e2, end);
            if (dpt_bloom_excluded(c)) return e2;
            e = e2 + c.len;
        }
        return e;
    }
    // no match starts here: the uncovered stretch runs to where the next match begins
    e = p + c0.len;
    while (e < end) {
        const DptChar c = T.at(e, end);
        if (!dpt_bloom_excluded(c)) break;
        if (c.cp == 0x20u && e + 1 < end && !dpt_bloom_excluded(T.at(e + 1, end))) break;
        e += c.len;
    }
    return e;
}

template <class Src, class Skip>
DPT_HD int64_t dpt_piece_end_src(int32_t rule, const Src& T, int64_t p, int64_t end, const Skip& skip) {
    return rule == 3 /* DPT_RULE_LLAMA3 */  ? dpt_piece_end_llama3(T, p, end, skip)
           : rule == 4 /* DPT_RULE_BLOOM */ ? dpt_piece_end_bloom(T, p, end, skip)
                                            : dpt_piece_end_gpt2(T, p, end, skip);
}
template <class Skip>
DPT_HD int64_t dpt_piece_end(int32_t rule, const DptUniView& U, const uint8_t* text, int64_t p, int64_t end, const Skip& skip) {
    return dpt_piece_end_src(rule, DptTextSrc{U, text}, p, end, skip);
}
DPT_HD int64_t dpt_piece_end(int32_t rule, const DptUniView& U, const uint8_t* text, int64_t p, int64_t end) {
    return dpt_piece_end(rule, U, text, p, end, DptNoSkip{});
}

// Is p (doc_start < p < end, text[p] == ' ') a synchronisation point: a space whose next character, in the same
// document, is a non-whitespace character (GPT-2, Llama-3) / a character of the regex's class (BLOOM)?
DPT_HD bool dpt_is_sync_space(int32_t rule, const DptUniView& U, const uint8_t* text, int64_t p, int64_t end) {
    if (text[p] != 0x20u || p + 1 >= end) return false;
    const DptChar c = dpt_char_at(U, text, p + 1, end);
    return rule == 4 /* DPT_RULE_BLOOM */ ? !dpt_bloom_excluded(c) : c.cls != DPT_CLS_S;
}
