// Corpus pipeline of the shortest-tokenization path, host/device source.
//
//   A  scan + dedup   one CTA per 4 KB tile of RAW corpus bytes: coalesced 16-byte loads into shared memory,
//                     one-bit-per-byte class masks, the tokenizer's boundary rule -> word starts, and for every
//                     word one probe of an HBM/L2-resident hash table keyed by the word's bytes.  The first
//                     occurrence claims the slot with one 64-bit CAS (the tag holds hash, length and the byte
//                     offset of that occurrence, so later occurrences verify against the immutable corpus text
//                     and never wait on another thread).  Output: one 32-bit ref per word, in corpus order
//                     (decoupled look-back over per-tile word counts), and the list of distinct words.
//   B  DP             ONE shortest-tokenization DP per DISTINCT word (dp_tokenize.py:24-84 in the closed form of
//                     dpt_dp_core.h): normalise (U+2581 marker, "<0xHH>" expansion of out-of-vocabulary
//                     characters), forward DP with the reference's tie order, backward select, ids -> the
//                     word's 16-byte result record.
//   C  emit           one thread per 4 words: ref -> result record -> token count; block scan + decoupled
//                     look-back over tiles -> final offsets; ids, per-word lengths, flags, document offsets and
//                     counters written once, in corpus order.
//
// The reference runs the O(n^2) DP for every occurrence of every word (tokenizer_utils.py:70-75); natural text
// repeats words (Zipf), so B does ~5 % of that work on the S2ORC-shaped benchmark corpus while A and C stream.
// The table lives in the caller's workspace and is rebuilt by every call: nothing is cached between calls.
//
// Written against a small "block" interface (tid/sync/scan/atomics/look-back) so the same source runs as CUDA
// kernels (pipe.cu, DevBlk) and under a std::thread emulation (tests/host_sim, HostBlk).
#pragma once
#include "dpt_common.h"
#include "dpt_dp_core.h"

#if defined(__CUDACC__)
#define DPT_PIPE_FN __device__ __forceinline__
#else
#define DPT_PIPE_FN inline
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
#endif

namespace dpt {

constexpr int PA_T = 4096;                    // raw bytes per tile of kernel A
constexpr int PA_HALO = 32;                   // look-behind (multiple of 32 keeps mask words aligned)
constexpr int PA_LA = 96;                     // look-ahead: a word that ends within it is handled in-tile
constexpr int PA_R = PA_HALO + PA_T + PA_LA;  // region bytes = 4224 = 132 * 32
constexpr int PA_NW = PA_R / 32;
constexpr int PA_THREADS = 256;
constexpr int PA_MAXLEN = 63;                 // longest word body (bytes) that goes through the dedup table
constexpr int PA_PROBES = 8;
constexpr int PB_THREADS = 128;
constexpr int PB_LOCAL = 72;                  // normalised bytes solved with per-thread local state in kernel B
constexpr int PC_THREADS = 256;
constexpr int PC_PER = 4;                     // words per thread in kernel C
constexpr int PC_TILE = PC_THREADS * PC_PER;

constexpr uint32_t REF_BOS = 0xFFFFFFFFu;     // the '<s>' word in front of every SPM_LLAMA document
constexpr uint32_t REF_ODD = 0x80000000u;     // | index into the odd-word list (not deduplicated)
constexpr unsigned long long PD_MASK = (1ull << 62) - 1;

constexpr uint32_t RES_UNTOK = 1u << 24;      // result meta: word_len (24 bits) | flags
constexpr uint32_t RES_POOLED = 1u << 25;     // ids live in the pool at rec.y (more than 3 ids)
constexpr uint32_t RES_LONG = 1u << 26;       // solved by the long-word kernel

struct OddWord {
    int64_t pos;   // global offset of the word's first raw byte
    int32_t len;   // raw bytes
    int32_t virt;  // 1: first word of its document (gets the Prepend(U+2581) marker)
};

struct PipeCtl {  // device-side counters, zeroed by the launcher
    unsigned int ticket_a, ticket_c;
    unsigned int n_pending, n_odd, n_long, pad;
    unsigned long long pool_used, lp_used, n_words, n_untok, n_too_long;
};

struct PipeParams {
    DptVocabView V;
    const uint8_t* text;
    int64_t n_bytes;
    const int64_t* doc_offs;  // n_docs + 1, doc_offs[0] == 0, doc_offs[n_docs] == n_bytes
    int64_t n_docs;
    int32_t* ids;
    int64_t ids_cap;
    int32_t* word_lens;
    uint8_t* word_flags;
    int64_t word_cap;
    int64_t* doc_tok_offs;  // n_docs + 1
    uint8_t* doc_flags;     // n_docs (zeroed by the launcher) or null
    unsigned long long* counters;  // 4
    int64_t* n_out;                // 8
    // workspace
    uint32_t* refs;               // word_cap
    int64_t* doc_first_word;      // n_docs + 1
    unsigned long long* tags;     // n_slots, zeroed by the launcher
    uint4* res;                   // n_slots
    uint32_t* pending;            // n_slots
    OddWord* odd;                 // odd_cap
    uint4* odd_res;               // odd_cap
    int32_t* pool;                // pool_cap ids of words with more than 3 tokens
    uint32_t* longq;              // n_slots + odd_cap
    uint8_t* lp_norm;             // long-word scratch: lp_cap positions
    uint64_t* lp_best;
    uint16_t* lp_a;
    uint16_t* lp_b;
    PipeCtl* ctl;
    unsigned long long* desc_w;   // n_tiles look-back descriptors of kernel A (zeroed)
    unsigned long long* desc_t;   // n_ctiles look-back descriptors of kernel C (zeroed)
    int64_t odd_cap, pool_cap, lp_cap;
    uint32_t slot_mask;
    int32_t n_tiles, n_ctiles;
    int32_t spm;  // 1: SPM_LLAMA rule; 0: byte-level rules
    int32_t rule;
};

struct ASmem {
    alignas(16) uint8_t text[PA_R + 64];
    uint32_t mDS[PA_NW + 2];  // document starts (and the end-of-text sentinel)
    uint32_t mCS[PA_NW + 2];  // code-point start bytes
    uint32_t mSP[PA_NW + 2];  // ' '
    uint32_t mM3[PA_NW + 2];  // E2 96 81 candidates
    uint32_t mCF[PA_NW + 2];  // character starts = CS | DS
    uint32_t mWS[PA_NW + 2];  // word starts
    uint32_t mCX[PA_NW + 2];  // positions that make their word "odd" (solved from the raw text, not deduplicated)
    uint32_t cnt[PA_NW + 2];
    uint16_t wlist[2 * PA_T];  // region index of every word that starts in this tile (| 0x8000: its '<s>' word)
    uint32_t pend[PA_T];       // table slots claimed by this tile
    uint32_t scan[40];
    int32_t tile, d_first, n_entries;
    uint32_t n_pend, pend_base;
    unsigned long long base_w;
};

struct CSmem {
    uint32_t scan[40];
    int32_t tile;
    uint32_t tile_tot, n_untok;
    unsigned long long base_t;
};

// ---- small helpers ------------------------------------------------------------------------------------
DPT_HD int pp_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
DPT_HD int pp_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// smallest set bit index >= from and < limit in mask m, or `limit`
DPT_HD int pp_mask_next(const uint32_t* m, int from, int limit) {
    if (from >= limit) return limit;
    int w = from >> 5;
    const int wl = (limit - 1) >> 5;
    uint32_t x = m[w] & (~0u << (from & 31));
    while (!x) {
        if (++w > wl) return limit;
        x = m[w];
    }
    const int r = (w << 5) + pp_ctz(x);
    return r < limit ? r : limit;
}
DPT_HD bool pp_bit(const uint32_t* m, int r) { return (m[r >> 5] >> (r & 31)) & 1u; }
DPT_HD uint32_t pp_range_mask(int w, int lo, int hi) {  // bits of mask word w inside [lo, hi)
    const int a = lo - (w << 5), b = hi - (w << 5);
    if (b <= 0 || a >= 32) return 0u;
    const uint32_t ma = a <= 0 ? ~0u : (~0u << a);
    const uint32_t mb = b >= 32 ? ~0u : ((1u << b) - 1u);
    return ma & mb;
}
DPT_HD bool pp_any_in_range(const uint32_t* m, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi - 1) >> 5; ++w)
        if (m[w] & pp_range_mask(w, lo, hi)) return true;
    return false;
}
DPT_HD int64_t pp_lower_bound(const int64_t* a, int64_t n, int64_t x) {  // first i with a[i] >= x
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
DPT_HD int64_t pp_upper_bound(const int64_t* a, int64_t n, int64_t x) {  // first i with a[i] > x
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

DPT_HD unsigned long long pp_tag(uint32_t hash, int len, int64_t pos) {
    return ((unsigned long long)(hash & 0xFFFFFu) << 44) | ((unsigned long long)len << 38) | (unsigned long long)(pos + 1);
}
DPT_HD int pp_tag_len(unsigned long long t) { return (int)((t >> 38) & 63u); }
DPT_HD int64_t pp_tag_pos(unsigned long long t) { return (int64_t)(t & ((1ull << 38) - 1)) - 1; }

// End (global offset) of the SPM word starting at g_ws, scanning the raw text: the next marker character that
// is not preceded by a marker, or the end of the document (dpt_rules.h's rule, sequentially).
DPT_PIPE_FN int64_t pp_spm_word_end_global(const PipeParams& P, int64_t g_ws, int ml, bool virt) {
    const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, g_ws);
    const int64_t doc_end = d <= P.n_docs ? P.doc_offs[d] : P.n_bytes;
    int64_t p = g_ws + (virt ? 0 : ml);
    bool prevm = true;  // the initial marker (virtual or real)
    while (p < doc_end) {
        int64_t e = p + 1;
        while (e < doc_end && !dpt_is_cp_start(P.text[e])) ++e;
        const uint32_t c0 = P.text[p];
        const bool mk = (c0 == 0x20u) || (e - p == 3 && c0 == DPT_MARK0 && P.text[p + 1] == DPT_MARK1 && P.text[p + 2] == DPT_MARK2);
        if (mk && !prevm) break;
        prevm = mk;
        p = e;
    }
    return p;
}

// =========================================================================================================
// Kernel A: scan + dedup
// =========================================================================================================
template <class Blk>
DPT_PIPE_FN void pa_run_tile(Blk& blk, const PipeParams& P, ASmem& S, const int tile) {
    const int tid = blk.tid(), nt = blk.nthreads();
    const int64_t t0 = (int64_t)tile * PA_T;
    const int64_t g0 = t0 - PA_HALO;  // global offset of region index 0
    const int64_t n = P.n_bytes;
    const int tvalid = (int)((n - t0) < PA_T ? (n - t0) : PA_T);
    const int own_lo = PA_HALO, own_hi = PA_HALO + tvalid;
    const bool spm = P.spm != 0;

    // ---- load: coalesced 16-byte loads of the region ------------------------------------------------------
    {
        const bool aligned = (((uintptr_t)P.text) & 15u) == 0;
        for (int i = tid; i < PA_R / 16; i += nt) {
            const int64_t g = g0 + 16 * (int64_t)i;
            if (aligned && g >= 0 && g + 16 <= n) {
                *reinterpret_cast<uint4*>(&S.text[16 * i]) = *reinterpret_cast<const uint4*>(P.text + g);
            } else {
                for (int k = 0; k < 16; ++k) {
                    const int64_t q = g + k;
                    S.text[16 * i + k] = (q >= 0 && q < n) ? P.text[q] : (uint8_t)0;
                }
            }
        }
        for (int i = tid; i < 64; i += nt) S.text[PA_R + i] = 0;
        for (int w = tid; w < PA_NW + 2; w += nt) {
            S.mDS[w] = 0;
            if (w >= PA_NW) S.mCS[w] = S.mSP[w] = S.mM3[w] = S.mCF[w] = S.mWS[w] = S.mCX[w] = S.cnt[w] = 0;
        }
        if (tid == 0) {
            S.d_first = (int32_t)pp_lower_bound(P.doc_offs, P.n_docs + 1, g0 < 0 ? 0 : g0);
            S.n_pend = 0;
        }
    }
    blk.sync();

    // ---- byte-class masks (one bit per byte), document starts -----------------------------------------------
    for (int w = tid; w < PA_NW; w += nt) {
        uint32_t cs = 0, sp = 0, m3 = 0;
        const uint8_t* t = &S.text[32 * w];
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const uint32_t b = t[k];
            cs |= (uint32_t)((b & 0xC0u) != 0x80u) << k;
            sp |= (uint32_t)(b == 0x20u) << k;
            m3 |= (uint32_t)(b == DPT_MARK0 && t[k + 1] == DPT_MARK1 && t[k + 2] == DPT_MARK2) << k;
        }
        S.mCS[w] = cs;
        S.mSP[w] = spm ? sp : 0u;
        S.mM3[w] = spm ? m3 : 0u;
    }
    for (int64_t k = (int64_t)S.d_first + tid; k <= P.n_docs; k += nt) {
        const int64_t o = P.doc_offs[k];
        if (o >= g0 + PA_R) break;
        const int r = (int)(o - g0);
        blk.atomic_or(&S.mDS[r >> 5], 1u << (r & 31));
    }
    blk.sync();

    // ---- boundary rule -> word starts ---------------------------------------------------------------------
    for (int w = tid; w < PA_NW; w += nt) {
        const uint32_t ds = S.mDS[w], dsn = S.mDS[w + 1], dsp = w ? S.mDS[w - 1] : 0u;
        const uint32_t cs = S.mCS[w], csn = (w + 1 < PA_NW) ? S.mCS[w + 1] : ~0u, csp = w ? S.mCS[w - 1] : 0u;
        // raw U+2581 at p: E2 96 81 inside one document, followed by a character start
        const uint32_t m3 = S.mM3[w] & ~((ds >> 1) | (dsn << 31)) & ~((ds >> 2) | (dsn << 30)) &
                            (((cs | ds) >> 3) | ((csn | dsn) << 29));
        const uint32_t m3p = w ? (S.mM3[w - 1] & ~((dsp >> 1) | (ds << 31)) & ~((dsp >> 2) | (ds << 30)) &
                                  (((csp | dsp) >> 3) | ((cs | ds) << 29)))
                               : 0u;
        const uint32_t sp = S.mSP[w], spp = w ? S.mSP[w - 1] : 0u;
        const uint32_t mk = sp | m3;
        const uint32_t pm = (sp << 1) | (spp >> 31) | (m3 << 3) | (m3p >> 29);  // previous character is a marker
        uint32_t ws, cx, amb = 0;
        if (spm) {
            ws = (mk & ~pm) | ds;
            cx = amb = mk & (pm | ds);  // a marker after a marker: word split depends on the BPE merge order
            // malformed UTF-8: continuation bytes glued to a space (the character rule swallows them into the
            // marker) -> solve the word from the raw text with the general character rule
            cx |= sp & ~(((cs | ds) >> 1) | ((csn | dsn) << 31));
        } else {
            ws = ds;  // byte-level rules add their own word starts; documents always split
            cx = 0;
        }
        S.mCF[w] = cs | ds;
        S.mWS[w] = ws;
        S.mCX[w] = cx;
        const uint32_t rm = pp_range_mask(w, own_lo, own_hi);
        S.cnt[w] = (uint32_t)(pp_popc(ws & rm) + (spm ? pp_popc(ds & rm) : 0));
        amb &= rm;
        while (amb && P.doc_flags) {
            const int r = (w << 5) + pp_ctz(amb);
            amb &= amb - 1;
            const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, g0 + r) - 1;
            if (d >= 0 && d < P.n_docs) P.doc_flags[d] = 1;  // DPT_DF_AMBIGUOUS
        }
    }
    blk.sync();

    // ---- word list (corpus order) ---------------------------------------------------------------------------
    {
        const int chunk = (PA_NW + nt - 1) / nt;
        const int w0 = tid * chunk, w1 = (w0 + chunk) < PA_NW ? (w0 + chunk) : PA_NW;
        uint32_t mine = 0;
        for (int w = w0; w < w1; ++w) mine += S.cnt[w];
        uint32_t total;
        uint32_t off = blk.exclusive_scan(mine, S.scan, total);
        for (int w = w0; w < w1; ++w) {
            uint32_t bits = S.mWS[w] & pp_range_mask(w, own_lo, own_hi);
            while (bits) {
                const int r = (w << 5) + pp_ctz(bits);
                bits &= bits - 1;
                if (spm && pp_bit(S.mDS, r)) S.wlist[off++] = (uint16_t)(r | 0x8000);
                S.wlist[off++] = (uint16_t)r;
            }
        }
        if (tid == 0) S.n_entries = (int32_t)total;
    }
    blk.sync();
    const int ne = S.n_entries;
    blk.lookback(P.desc_w, tile, (unsigned long long)ne, &S.base_w);
    blk.sync();
    const int64_t base_w = (int64_t)S.base_w;

    // ---- one table probe per word ------------------------------------------------------------------------------
    for (int k = tid; k < ne; k += nt) {
        const uint32_t e = S.wlist[k];
        const int ws = (int)(e & 0x7FFFu);
        const int64_t gw = base_w + k;
        uint32_t ref;
        if (e & 0x8000u) {
            ref = REF_BOS;
            const int64_t d = pp_lower_bound(P.doc_offs, P.n_docs + 1, g0 + ws);
            if (d < P.n_docs) P.doc_first_word[d] = gw;
        } else {
            const bool ds = pp_bit(S.mDS, ws);
            if (!spm && ds) {
                const int64_t d = pp_lower_bound(P.doc_offs, P.n_docs + 1, g0 + ws);
                if (d < P.n_docs) P.doc_first_word[d] = gw;
            }
            const int we = pp_mask_next(S.mWS, ws + 1, PA_R);
            const int ml = !spm ? 0 : ds ? 0 : (S.text[ws] == 0x20u ? 1 : 3);
            const int b = ws + ml, len = we - b;
            const bool open = we >= PA_R;
            bool odd = open || len > PA_MAXLEN || len < 0 || pp_any_in_range(S.mCX, ws, we);
            ref = 0;
            if (!odd) {
                uint32_t h = 0x811C9DC5u;
                for (int q = b; q < we; ++q) h = (h ^ S.text[q]) * 0x01000193u;
                h ^= h >> 15;
                h *= 0x2C1B3C6Du;
                h ^= h >> 13;
                const unsigned long long mine = pp_tag(h >> 12, len, g0 + b);
                uint32_t slot = h & P.slot_mask;
                bool done = false;
                for (int probe = 0; probe < PA_PROBES && !done; ++probe, slot = (slot + 1) & P.slot_mask) {
                    unsigned long long t = blk.load_relaxed(&P.tags[slot]);
                    if (t == 0) {
                        t = blk.cas_u64(&P.tags[slot], 0ull, mine);
                        if (t == 0) {  // first occurrence of this word: claim the slot, queue the DP
                            const uint32_t li = blk.atomic_add_ret(&S.n_pend, 1u);
                            S.pend[li] = slot;
                            ref = slot;
                            done = true;
                            break;
                        }
                    }
                    if ((t >> 38) == (mine >> 38)) {  // same hash bits and length: verify against the corpus text
                        const uint8_t* rep = P.text + pp_tag_pos(t);
                        bool same = true;
                        for (int q = 0; q < len && same; ++q) same = rep[q] == S.text[b + q];
                        if (same) {
                            ref = slot;
                            done = true;
                        }
                    }
                }
                odd = !done;  // neighbourhood full: solve this occurrence on its own
            }
            if (odd) {
                const int64_t g_ws = g0 + ws;
                const int64_t g_we = open ? (spm ? pp_spm_word_end_global(P, g_ws, ml, ds) : g_ws + 1) : g0 + we;
                const uint32_t j = blk.atomic_add_ret(&P.ctl->n_odd, 1u);
                if ((int64_t)j < P.odd_cap) {
                    OddWord o;
                    o.pos = g_ws;
                    o.len = (int32_t)(g_we - g_ws);
                    o.virt = (spm && ds) ? 1 : 0;
                    P.odd[j] = o;
                }
                ref = REF_ODD | j;
            }
        }
        if (gw < P.word_cap) P.refs[gw] = ref;
    }
    blk.sync();
    if (tid == 0) {
        S.pend_base = S.n_pend ? blk.atomic_add_ret(&P.ctl->n_pending, S.n_pend) : 0u;
        if (tile == P.n_tiles - 1) {
            P.ctl->n_words = (unsigned long long)(base_w + ne);
            P.doc_first_word[P.n_docs] = base_w + ne;
        }
    }
    blk.sync();
    for (uint32_t i = tid; i < S.n_pend; i += nt) P.pending[S.pend_base + i] = S.pend[i];
    blk.sync();
}

template <class Blk>
DPT_PIPE_FN void pa_kernel(Blk& blk, const PipeParams& P, ASmem& S) {
    // tiles are handed out in corpus order by an atomic ticket, so the look-back only ever waits for tiles held
    // by CTAs that are already running
    for (;;) {
        if (blk.tid() == 0) S.tile = (int32_t)blk.atomic_add_ret(&P.ctl->ticket_a, 1u);
        blk.sync();
        const int tile = S.tile;
        if (tile >= P.n_tiles) break;
        pa_run_tile(blk, P, S, tile);
        if (!blk.persistent()) break;
    }
}

// =========================================================================================================
// Kernel B: one DP per distinct word
// =========================================================================================================
// SPM_LLAMA normalisation of raw bytes [p, e) of one document into out (capacity cap); `marker` = emit the
// word-initial U+2581 first (deduplicated bodies and document-first words).  Returns the normalised length,
// or -1 if it does not fit.  Character rule = dpt_rules.h (dpt_spm_classify / dpt_spm_write_char).
DPT_PIPE_FN int32_t pb_normalise(const PipeParams& P, int64_t p, int64_t e_end, bool marker, uint8_t* out, int32_t cap) {
    const DptVocabView& V = P.V;
    int32_t n = 0;
    if (!P.spm) {
        if (e_end - p > cap) return -1;
        for (; p < e_end; ++p) out[n++] = P.text[p];
        return n;
    }
    if (marker) {
        if (cap < 3) return -1;
        out[0] = DPT_MARK0; out[1] = DPT_MARK1; out[2] = DPT_MARK2;
        n = 3;
    }
    while (p < e_end) {
        int64_t e = p + 1;
        while (e < e_end && !dpt_is_cp_start(P.text[e])) ++e;
        const uint32_t c0 = P.text[p];
        const int32_t src = (int32_t)(e - p);
        const bool mk = (c0 == 0x20u) || (src == 3 && c0 == DPT_MARK0 && P.text[p + 1] == DPT_MARK1 && P.text[p + 2] == DPT_MARK2);
        if (mk) {
            if (n + 3 > cap) return -1;
            out[n] = DPT_MARK0; out[n + 1] = DPT_MARK1; out[n + 2] = DPT_MARK2;
            n += 3;
        } else {
            bool in_vocab;
            if (src == 1 && c0 < 128u) {
                in_vocab = (V.ascii_single[c0 >> 5] >> (c0 & 31)) & 1u;
            } else {
                uint32_t entry = DPT_DA_ROOT_ENTRY;
                in_vocab = true;
                for (int64_t q = p; q < e && in_vocab; ++q) in_vocab = dpt_da_step(V.da, entry, P.text[q]);
                in_vocab = in_vocab && (entry & DPT_DA_TERMINAL);
            }
            if (in_vocab) {
                if (n + src > cap) return -1;
                for (int64_t q = p; q < e; ++q) out[n++] = P.text[q];
            } else {
                if (n + 6 * src > cap) return -1;
                for (int64_t q = p; q < e; ++q) {
                    const uint32_t b = P.text[q];
                    out[n + 0] = '<'; out[n + 1] = '0'; out[n + 2] = 'x';
                    out[n + 3] = (uint8_t)((b >> 4) < 10 ? '0' + (b >> 4) : 'A' + (b >> 4) - 10);
                    out[n + 4] = (uint8_t)((b & 15) < 10 ? '0' + (b & 15) : 'A' + (b & 15) - 10);
                    out[n + 5] = '>';
                    n += 6;
                }
            }
        }
        p = e;
    }
    return n;
}

// item i of the DP work list: i < n_pending -> table slot pending[i]; else odd word i - n_pending
struct PbItem {
    int64_t pos, end;
    bool marker;
    uint4* out;
};
DPT_PIPE_FN PbItem pb_item(const PipeParams& P, uint32_t i, uint32_t n_pending) {
    PbItem it;
    if (i < n_pending) {
        const uint32_t slot = P.pending[i];
        const unsigned long long t = P.tags[slot];
        it.pos = pp_tag_pos(t);
        it.end = it.pos + pp_tag_len(t);
        it.marker = P.spm != 0;
        it.out = &P.res[slot];
    } else {
        const OddWord o = P.odd[i - n_pending];
        it.pos = o.pos;
        it.end = o.pos + o.len;
        it.marker = o.virt != 0;
        it.out = &P.odd_res[i - n_pending];
    }
    return it;
}

template <class Blk>
DPT_PIPE_FN void pb_store(Blk& blk, const PipeParams& P, uint4* out, uint32_t word_len, bool reach, const int32_t* ids, uint32_t extra) {
    uint4 r;
    r.x = (word_len & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | extra;
    r.y = r.z = r.w = 0;
    if (reach) {
        if (word_len <= 3) {
            r.y = (uint32_t)ids[0];
            if (word_len > 1) r.z = (uint32_t)ids[1];
            if (word_len > 2) r.w = (uint32_t)ids[2];
        } else {
            const unsigned long long off = blk.atomic_add_u64_ret(&P.ctl->pool_used, (unsigned long long)word_len);
            r.x |= RES_POOLED;
            r.y = (uint32_t)(off & 0xFFFFFFFFull);
            r.z = (uint32_t)(off >> 32);
            if ((int64_t)(off + word_len) <= P.pool_cap)
                for (uint32_t k = 0; k < word_len; ++k) P.pool[off + k] = ids[k];
        }
    }
    *out = r;
}

// one thread per distinct word, state in registers / local memory
template <class Blk>
DPT_PIPE_FN void pb_thread(Blk& blk, const PipeParams& P, int64_t gtid, int64_t gthreads) {
    const uint32_t n_pending = P.ctl->n_pending;
    const uint32_t n_odd = P.ctl->n_odd < (uint32_t)P.odd_cap ? P.ctl->n_odd : (uint32_t)P.odd_cap;
    const uint64_t total = (uint64_t)n_pending + n_odd;
    for (uint64_t i = (uint64_t)gtid; i < total; i += (uint64_t)gthreads) {
        const PbItem it = pb_item(P, (uint32_t)i, n_pending);
        uint8_t norm[PB_LOCAL + 8];
        const int32_t nlen = pb_normalise(P, it.pos, it.end, it.marker, norm, PB_LOCAL);
        if (nlen < 0) {
            const uint32_t q = blk.atomic_add_ret(&P.ctl->n_long, 1u);
            P.longq[q] = (uint32_t)i;
            continue;
        }
        if (nlen == 0) {  // cannot happen (documents are non-empty); keep the record defined
            uint4 z; z.x = RES_UNTOK; z.y = z.z = z.w = 0;
            *it.out = z;
            continue;
        }
        uint64_t best[PB_LOCAL + 1];
        uint16_t A[PB_LOCAL + 1], B[PB_LOCAL + 1];
        dpt_forward<true>(P.V, norm, nlen, nullptr, best, A, B);
        const uint64_t kn = best[nlen];
        const uint32_t word_len = dpt_key_len(kn);
        const bool reach = dpt_key_reach(kn);
        int32_t ids[PB_LOCAL + 1];
        if (reach) dpt_backward_emit(P.V, norm, nlen, best, A, B, ids, PB_LOCAL + 1);
        pb_store(blk, P, it.out, word_len, reach, ids, 0u);
    }
}

// long words: one thread each, state in a global scratch pool (13 bytes per normalised position)
template <class Blk>
DPT_PIPE_FN void pb_long_thread(Blk& blk, const PipeParams& P, int64_t gtid, int64_t gthreads) {
    const uint32_t n_pending = P.ctl->n_pending;
    const uint32_t n_long = P.ctl->n_long;
    for (uint64_t k = (uint64_t)gtid; k < n_long; k += (uint64_t)gthreads) {
        const PbItem it = pb_item(P, P.longq[k], n_pending);
        const int64_t raw = it.end - it.pos;
        const int64_t need = (P.spm ? 6 * raw + 3 : raw) + 2;
        const unsigned long long off = blk.atomic_add_u64_ret(&P.ctl->lp_used, (unsigned long long)need);
        if ((int64_t)(off + need) > P.lp_cap || need >= (1ll << 31)) {  // reported through n_out; caller retries bigger
            uint4 z; z.x = RES_UNTOK | RES_LONG; z.y = z.z = z.w = 0;
            *it.out = z;
            blk.atomic_add_u64_ret(&P.ctl->n_too_long, 1ull);
            continue;
        }
        uint8_t* norm = P.lp_norm + off;
        uint64_t* best = P.lp_best + off;
        uint16_t* A = P.lp_a + off;
        uint16_t* B = P.lp_b + off;
        const int32_t nlen = pb_normalise(P, it.pos, it.end, it.marker, norm, (int32_t)(need - 2));
        dpt_forward<true>(P.V, norm, nlen, nullptr, best, A, B);
        const uint64_t kn = best[nlen];
        const uint32_t word_len = dpt_key_len(kn);
        const bool reach = dpt_key_reach(kn);
        uint4 r;
        r.x = (word_len & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | RES_LONG;
        r.y = r.z = r.w = 0;
        if (reach) {
            const unsigned long long po = blk.atomic_add_u64_ret(&P.ctl->pool_used, (unsigned long long)word_len);
            r.x |= RES_POOLED;
            r.y = (uint32_t)(po & 0xFFFFFFFFull);
            r.z = (uint32_t)(po >> 32);
            if ((int64_t)(po + word_len) <= P.pool_cap) dpt_backward_emit(P.V, norm, nlen, best, A, B, P.pool + po, (int64_t)word_len);
        }
        *it.out = r;
    }
}

// =========================================================================================================
// Kernel C: scan + emit
// =========================================================================================================
DPT_PIPE_FN uint4 pc_record(const PipeParams& P, uint32_t ref) {
    uint4 r;
    if (ref == REF_BOS) {
        r.x = (uint32_t)P.V.bos_len | (P.V.bos_ntok ? 0u : RES_UNTOK);
        r.y = (uint32_t)P.V.bos_ids[0];
        r.z = (uint32_t)P.V.bos_ids[1];
        r.w = (uint32_t)P.V.bos_ids[2];
        return r;
    }
    const uint4* src;
    if (ref & REF_ODD) {
        const uint32_t j = ref & 0x7FFFFFFFu;
        if ((int64_t)j >= P.odd_cap) {
            r.x = RES_UNTOK; r.y = r.z = r.w = 0;
            return r;
        }
        src = &P.odd_res[j];
    } else {
        src = &P.res[ref];
    }
#if defined(__CUDA_ARCH__)
    return __ldg(src);
#else
    return *src;
#endif
}

template <class Blk>
DPT_PIPE_FN void pc_run_tile(Blk& blk, const PipeParams& P, CSmem& S, const int tile) {
    const int tid = blk.tid();
    const int64_t n_words = (int64_t)P.ctl->n_words < P.word_cap ? (int64_t)P.ctl->n_words : P.word_cap;
    const int64_t w0 = (int64_t)tile * PC_TILE + (int64_t)tid * PC_PER;
    uint4 rec[PC_PER];
    uint32_t ntok[PC_PER];
    uint32_t mine = 0, untok = 0;
#pragma unroll
    for (int k = 0; k < PC_PER; ++k) {
        ntok[k] = 0;
        if (w0 + k < n_words) {
            rec[k] = pc_record(P, P.refs[w0 + k]);
            if (rec[k].x & RES_UNTOK) ++untok; else ntok[k] = rec[k].x & 0xFFFFFFu;
            mine += ntok[k];
        }
    }
    if (tid == 0) S.n_untok = 0;
    uint32_t total;
    uint32_t off = blk.exclusive_scan(mine, S.scan, total);
    if (untok) blk.atomic_add(&S.n_untok, untok);
    blk.sync();
    blk.lookback(P.desc_t, tile, (unsigned long long)total, &S.base_t);
    blk.sync();
    int64_t gt = (int64_t)S.base_t + off;
#pragma unroll
    for (int k = 0; k < PC_PER; ++k) {
        const int64_t w = w0 + k;
        if (w >= n_words) break;
        const uint32_t meta = rec[k].x;
        P.word_lens[w] = (int32_t)(meta & 0xFFFFFFu);
        P.word_flags[w] = (uint8_t)(((meta & RES_UNTOK) ? 1u : 0u) | ((meta & RES_LONG) ? 4u : 0u));
        const uint32_t ref = P.refs[w0 + k];
        const bool doc_first = P.spm ? (ref == REF_BOS) : false;
        if (doc_first) {
            const int64_t d = pp_lower_bound(P.doc_first_word, P.n_docs, w);
            if (d < P.n_docs && P.doc_first_word[d] == w) P.doc_tok_offs[d] = gt;
        }
        const uint32_t nk = ntok[k];
        if (nk) {
            if (meta & RES_POOLED) {
                const int64_t po = (int64_t)rec[k].y | ((int64_t)rec[k].z << 32);
                for (uint32_t q = 0; q < nk; ++q)
                    if (gt + q < P.ids_cap && po + q < P.pool_cap) P.ids[gt + q] = P.pool[po + q];
            } else {
                if (gt < P.ids_cap) P.ids[gt] = (int32_t)rec[k].y;
                if (nk > 1 && gt + 1 < P.ids_cap) P.ids[gt + 1] = (int32_t)rec[k].z;
                if (nk > 2 && gt + 2 < P.ids_cap) P.ids[gt + 2] = (int32_t)rec[k].w;
            }
        }
        gt += nk;
    }
    if (tid == 0) {
        if (S.n_untok) blk.atomic_add_u64_ret(&P.ctl->n_untok, (unsigned long long)S.n_untok);
    }
    blk.sync();
}

// byte-level rules: every document start is a word start, not a separate word: document token offsets
template <class Blk>
DPT_PIPE_FN void pc_kernel(Blk& blk, const PipeParams& P, CSmem& S) {
    for (;;) {
        if (blk.tid() == 0) S.tile = (int32_t)blk.atomic_add_ret(&P.ctl->ticket_c, 1u);
        blk.sync();
        const int tile = S.tile;
        const int64_t n_words = (int64_t)P.ctl->n_words < P.word_cap ? (int64_t)P.ctl->n_words : P.word_cap;
        if (tile >= P.n_ctiles || (int64_t)tile * PC_TILE >= n_words) break;
        pc_run_tile(blk, P, S, tile);
        if (!blk.persistent()) break;
    }
}

// final counters, written by one thread after kernel C (stream order)
DPT_PIPE_FN void pd_finish(const PipeParams& P) {
    const int64_t n_words_true = (int64_t)P.ctl->n_words;
    const int64_t n_words = n_words_true < P.word_cap ? n_words_true : P.word_cap;
    const int64_t n_ctiles = (n_words + PC_TILE - 1) / PC_TILE;
    const unsigned long long tot = n_ctiles > 0 ? (P.desc_t[n_ctiles - 1] & PD_MASK) : 0ull;
    P.counters[0] = (unsigned long long)P.n_bytes;
    P.counters[1] = (unsigned long long)n_words_true;
    P.counters[2] = tot;
    P.counters[3] = P.ctl->n_untok;
    P.n_out[0] = (int64_t)tot;             // DPT_NOUT_IDS
    P.n_out[1] = n_words_true;             // DPT_NOUT_WORDS
    P.n_out[2] = (int64_t)P.ctl->lp_used;  // DPT_NOUT_POOL_REQ  (long-word scratch positions)
    P.n_out[3] = P.lp_cap;                 // DPT_NOUT_POOL_CAP
    P.n_out[4] = (int64_t)P.ctl->pool_used;  // ids pool required
    P.n_out[5] = P.pool_cap;
    P.n_out[6] = (int64_t)P.ctl->n_odd;    // odd words required
    P.n_out[7] = P.odd_cap;
    P.doc_tok_offs[P.n_docs] = (int64_t)tot;
}

}  // namespace dpt
