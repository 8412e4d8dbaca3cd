// Fused single-pass tile pipeline: boundary rule -> trie walks -> shortest-tokenization DP -> tie-break select ->
// compaction, one CTA per 4 KB tile of RAW corpus bytes, one kernel launch per corpus.
//
// Replaces, for a whole corpus at once, the per-document loop of /root/reference/packages/tokenizer_utils.py:66-80
// (pretokenize -> compute_shortest_tokenizations -> obtain_longest_token -> ids) whose inner operations are
// dp_tokenize.py:35-47 (forward), :49-70 (backtrace) and :72-84 (selection).
//
// The code is written against a tiny "block" interface (tid/sync/scan/atomics/look-back) so that the SAME
// source runs (a) as the CUDA kernel in fused.cu (DevBlk: __syncthreads, shuffles, relaxed gpu-scope loads) and
// (b) under a std::thread emulation in tests/host_sim (HostBlk) for CPU parity tests against the oracle.
//
// Per tile (region = 32 B look-behind + 4096 B tile + 256 B look-ahead, all in shared memory):
//   P0  coalesced 16-byte loads of the region; binary search of the first document in the region
//   P1  one-bit-per-byte masks: code-point starts, spaces, raw U+2581, document starts -> word starts (WS),
//       "complex" positions (CX: non-initial markers, out-of-vocabulary characters, over-long tokens)
//   P2  ONE THREAD PER BYTE POSITION walks the double-array trie (hot slots staged in shared memory) from its
//       position to the end of its word and records the 32-bit end mask E[p] (bit k: text[p..p+k] is a token)
//   P3  one thread per word: push-form forward DP over (len, longest-token) keys kept in shared memory
//       (dpt_dp_core.h's ordering packed into 32 bits), backpointers A/B as 1-byte distances
//   P4  block scan of (words, tokens) + decoupled look-back across tiles -> global offsets, no second pass
//   P5  one thread per word: pointer-chase the selected segmentation, re-walk each token for its id, write
//       ids / per-word lengths / flags / document token offsets straight to their final place
// A word is "simple" when every character is itself a vocabulary entry (no "<0xHH>" expansion), it holds no
// marker after its first character, no walk ran past 32 bytes and it ends inside the region; then every DP
// position is reachable, the phantom initialisation of dp_tokenize.py:28 can never undercut a real path, and
// the 32-bit key suffices.  Every other word ("complex") is normalised into a per-CTA global arena and solved
// there with the general 64-bit-key code of dpt_dp_core.h (exact for unreachable positions, any token length).
// Words longer than the arena slot are counted in n_out[DPT_NOUT_FALLBACK]; the host then reruns the batch
// through the general multi-kernel path (kernels.cu).
#pragma once
#include "dpt_common.h"
#include "dpt_dp_core.h"

#if defined(__CUDACC__)
#define DPT_TILE_FN __device__ __forceinline__
#else
#define DPT_TILE_FN inline
#endif

#if !defined(__CUDACC__)
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
#endif

namespace dpt {

constexpr int TL_T = 4096;                       // raw bytes per tile
constexpr int TL_HALO = 32;                      // look-behind (multiple of 32: mask words stay aligned)
constexpr int TL_LA = 256;                       // look-ahead for the tail of the tile's last word
constexpr int TL_R = TL_HALO + TL_T + TL_LA;     // region bytes = 4384 = 137 * 32
constexpr int TL_NW = TL_R / 32;                 // mask words
constexpr int TL_KC = 32768;                     // trie slots staged in shared memory (128 KB)
constexpr int TL_THREADS = 512;
constexpr int TL_BIGTAIL = 2048;                 // raw bytes a word may run past the region and stay in-kernel
constexpr int TL_ARENA_POS = 10 * TL_T + 6 * (TL_LA + TL_BIGTAIL) + 128;  // normalised positions per CTA
constexpr int TL_ARENA_BYTES = TL_ARENA_POS * 13;                         // norm 1 + best 8 + A 2 + B 2
constexpr int TL_WALK_BITS = 32;

constexpr uint32_t TL_KEY_INF = 0xFFFFFFFFu;
constexpr uint32_t TL_REC_COMPLEX = 0x80000000u;
constexpr uint32_t TL_REC_UNTOK = 0x40000000u;
constexpr unsigned long long TL_DESC_MASK = (1ull << 62) - 1;

// index 6 of the int64[8] status vector: words the fused kernel could not solve in-kernel
#define DPT_NOUT_FALLBACK_IDX 6

struct TileParams {
    DptVocabView V;
    const uint8_t* text;
    int64_t n_bytes;
    const int64_t* doc_offs;  // n_docs + 1 entries, doc_offs[0] == 0, doc_offs[n_docs] == n_bytes
    int64_t n_docs;
    int32_t* ids;
    int64_t ids_cap;
    int32_t* word_lens;
    uint8_t* word_flags;
    int64_t word_cap;
    int64_t* doc_tok_offs;  // n_docs + 1
    uint8_t* doc_flags;     // n_docs, zeroed by the launcher; may be null
    unsigned long long* counters;  // 4, zeroed by the launcher
    int64_t* n_out;                // 8, zeroed by the launcher
    unsigned long long* desc_w;    // n_tiles look-back descriptors, zeroed by the launcher
    unsigned long long* desc_t;
    unsigned int* ticket;          // zeroed by the launcher
    uint8_t* arena_norm;           // n_ctas * TL_ARENA_POS each
    uint64_t* arena_best;
    uint16_t* arena_a;
    uint16_t* arena_b;
    int32_t n_tiles;
    int32_t kc;    // trie slots staged in shared memory: min(TL_KC, n_slots)
    int32_t spm;   // 1: SPM_LLAMA rule (markers, "<s>" words, code-point units); 0: byte-level rules
    int32_t rule;
};

struct TileSmem {
    uint32_t da_cache[TL_KC];
    uint32_t E[TL_R + 32];
    uint32_t key[TL_R + 32];
    uint8_t A[TL_R + 32];
    uint8_t B[TL_R + 32];
    alignas(16) uint8_t text[TL_R + 64];
    uint32_t mDS[TL_NW + 2];  // document starts (and the end-of-text sentinel)
    uint32_t mCS[TL_NW + 2];  // code-point start bytes
    uint32_t mSP[TL_NW + 2];  // ' '
    uint32_t mM3[TL_NW + 2];  // E2 96 81 candidates
    uint32_t mMK[TL_NW + 2];  // marker characters (space or raw U+2581)
    uint32_t mCF[TL_NW + 2];  // character starts = CS | DS
    uint32_t mWS[TL_NW + 2];  // word starts
    uint32_t mCX[TL_NW + 2];  // positions that make their word complex
    uint16_t wlist[TL_T];     // region index of every word that starts in this tile, in order
    uint32_t scan[40];
    int32_t tile;
    int32_t d_first;
    int32_t n_words_tile;
    int32_t last_end;
    uint32_t tile_tot;  // words << 16 | tokens
    uint32_t n_untok;
    uint32_t n_fallback;
    unsigned long long base_w, base_t;
};

// ---- small helpers ---------------------------------------------------------------------------------
DPT_HD int tl_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
DPT_HD int tl_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
DPT_HD int tl_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// smallest set bit index >= from and < limit in mask m, or `limit`
DPT_HD int tl_mask_next(const uint32_t* m, int from, int limit) {
    if (from >= limit) return limit;
    int w = from >> 5;
    const int wl = (limit - 1) >> 5;
    uint32_t x = m[w] & (~0u << (from & 31));
    while (!x) {
        if (++w > wl) return limit;
        x = m[w];
    }
    const int r = (w << 5) + tl_ctz(x);
    return r < limit ? r : limit;
}
DPT_HD bool tl_bit(const uint32_t* m, int r) { return (m[r >> 5] >> (r & 31)) & 1u; }
// bits of mask word w that lie in [lo, hi)
DPT_HD uint32_t tl_range_mask(int w, int lo, int hi) {
    const int a = lo - (w << 5), b = hi - (w << 5);
    if (b <= 0 || a >= 32) return 0u;
    const uint32_t ma = a <= 0 ? ~0u : (~0u << a);
    const uint32_t mb = b >= 32 ? ~0u : ((1u << b) - 1u);
    return ma & mb;
}
// set bits of m in [a, i), 0 < i - a <= 32
DPT_HD int tl_count(const uint32_t* m, int a, int i) {
    const uint64_t two = (uint64_t)m[a >> 5] | ((uint64_t)m[(a >> 5) + 1] << 32);
    const uint64_t x = (two >> (a & 31)) & ((1ull << (i - a)) - 1ull);
    return tl_popc64(x);
}
DPT_HD int64_t tl_lower_bound(const int64_t* a, int64_t n, int64_t x) {  // first i with a[i] >= x
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
DPT_HD int64_t tl_upper_bound(const int64_t* a, int64_t n, int64_t x) {  // first i with a[i] > x
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

DPT_TILE_FN uint32_t tl_da(const TileSmem& S, const TileParams& P, uint32_t idx) {
    if (idx < (uint32_t)P.kc) return S.da_cache[idx];
#if defined(__CUDA_ARCH__)
    return __ldg(P.V.da + idx);
#else
    return P.V.da[idx];
#endif
}

// Walk from `entry` over region bytes [start, start + limit).  Bit s of the result: a token ends after s+1
// bytes.  `slot` follows the node (for id lookup); `alive`: all `limit` bytes matched and the node has children.
DPT_TILE_FN uint32_t tl_walk(const TileSmem& S, const TileParams& P, uint32_t entry, int start, int limit, bool& alive) {
    uint32_t m = 0;
    int s = 0;
    for (; s < limit; ++s) {
        const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
        if (!base) break;
        const uint32_t c = S.text[start + s];
        const uint32_t e = tl_da(S, P, base + c);
        if ((e & DPT_DA_MATCH_MASK) != (DPT_DA_OCCUPIED | c)) break;
        entry = e;
        if (e & DPT_DA_TERMINAL) m |= 1u << s;
    }
    alive = (s == limit) && (entry >> DPT_DA_BASE_SHIFT) != 0;
    return m;
}
// id of the token made of (optional marker +) region bytes [j, i); the token is known to exist.
DPT_TILE_FN int32_t tl_token_id(const TileSmem& S, const TileParams& P, bool from_marker, int j, int i) {
    uint32_t entry = from_marker ? P.V.marker_entry : DPT_DA_ROOT_ENTRY;
    uint32_t slot = from_marker ? P.V.marker_slot : 0u;
    for (int p = j; p < i; ++p) {
        slot = (entry >> DPT_DA_BASE_SHIFT) + S.text[p];
        entry = tl_da(S, P, slot);
    }
#if defined(__CUDA_ARCH__)
    return __ldg(P.V.slot_id + slot);
#else
    return P.V.slot_id[slot];
#endif
}

DPT_HD uint32_t tl_key_extend(uint32_t kj, uint32_t cl) {
    const uint32_t lowj = kj & 0xFFu, lowe = 0xFFu - cl;
    return (kj & 0xFFFFFF00u) + 0x100u + (lowj < lowe ? lowj : lowe);
}

// ---- complex words: normalise into the arena and run the general DP ---------------------------------------
// Word = raw bytes [g_ws, g_we) of one document; `virt` = the word is the first of its document and gets the
// Prepend(U+2581) marker.  Character rules restate dpt_rules.h (dpt_spm_classify / dpt_spm_write_char).
DPT_TILE_FN uint32_t tl_complex_forward(const TileParams& P, int64_t cta, int apos, int64_t g_ws, int64_t g_we, bool virt) {
    const DptVocabView& V = P.V;
    const int64_t ab = cta * (int64_t)TL_ARENA_POS + apos;
    uint8_t* norm = P.arena_norm + ab;
    uint64_t* best = P.arena_best + ab;
    uint16_t* A = P.arena_a + ab;
    uint16_t* B = P.arena_b + ab;
    int32_t n = 0;
    if (P.spm) {
        if (virt) {
            norm[0] = DPT_MARK0; norm[1] = DPT_MARK1; norm[2] = DPT_MARK2;
            n = 3;
        }
        for (int64_t p = g_ws; p < g_we;) {
            int64_t e = p + 1;
            while (e < g_we && !dpt_is_cp_start(P.text[e])) ++e;
            const uint32_t c0 = P.text[p];
            const int32_t src = (int32_t)(e - p);
            const bool marker = (c0 == 0x20u) || (src == 3 && c0 == DPT_MARK0 && P.text[p + 1] == DPT_MARK1 && P.text[p + 2] == DPT_MARK2);
            if (marker) {
                norm[n] = DPT_MARK0; norm[n + 1] = DPT_MARK1; norm[n + 2] = DPT_MARK2;
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
                    for (int64_t q = p; q < e; ++q) norm[n++] = P.text[q];
                } else {
                    for (int64_t q = p; q < e; ++q) {
                        const uint32_t b = P.text[q];
                        norm[n + 0] = '<'; norm[n + 1] = '0'; norm[n + 2] = 'x';
                        norm[n + 3] = (uint8_t)((b >> 4) < 10 ? '0' + (b >> 4) : 'A' + (b >> 4) - 10);
                        norm[n + 4] = (uint8_t)((b & 15) < 10 ? '0' + (b & 15) : 'A' + (b & 15) - 10);
                        norm[n + 5] = '>';
                        n += 6;
                    }
                }
            }
            p = e;
        }
    } else {
        for (int64_t p = g_ws; p < g_we; ++p) norm[n++] = P.text[p];
    }
    dpt_forward<true>(V, norm, n, nullptr, best, A, B);
    B[0] = (uint16_t)n;  // B[0] is never read by the backward pass: keeps the normalised length for P5
    const uint64_t kn = best[n];
    return TL_REC_COMPLEX | (dpt_key_reach(kn) ? 0u : TL_REC_UNTOK) | (dpt_key_len(kn) & 0x3FFFFFFFu);
}

DPT_TILE_FN void tl_complex_emit(const TileParams& P, int64_t cta, int apos, int64_t gt) {
    const int64_t ab = cta * (int64_t)TL_ARENA_POS + apos;
    const uint16_t* B = P.arena_b + ab;
    const int32_t n = B[0];
    const int64_t cap = P.ids_cap > gt ? P.ids_cap - gt : 0;
    dpt_backward_emit(P.V, P.arena_norm + ab, n, P.arena_best + ab, P.arena_a + ab, B, P.ids + gt, cap);
}

// End (global offset) of the SPM word that starts at g_ws, found by scanning the raw text: the next marker
// character that is not preceded by a marker, or the end of the document.  Returns -1 when the word is longer
// than `max_len` raw bytes.
DPT_TILE_FN int64_t tl_spm_word_end_global(const TileParams& P, int64_t g_ws, int ml, bool virt, int64_t max_len) {
    const int64_t d = tl_upper_bound(P.doc_offs, P.n_docs + 1, g_ws);
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
        if (p - g_ws > max_len) return -1;
    }
    return p;
}

// ---- the tile ------------------------------------------------------------------------------------------
template <class Blk>
DPT_TILE_FN void tl_stage_trie(Blk& blk, const TileParams& P, TileSmem& S) {
    for (int i = blk.tid(); i < P.kc; i += blk.nthreads()) S.da_cache[i] = P.V.da[i];
    blk.sync();
}

template <class Blk>
DPT_TILE_FN void tl_run_tile(Blk& blk, const TileParams& P, TileSmem& S, const int tile, const int64_t cta) {
    const int tid = blk.tid(), nt = blk.nthreads();
    const int64_t t0 = (int64_t)tile * TL_T;
    const int64_t g0 = t0 - TL_HALO;  // global offset of region index 0
    const int64_t n = P.n_bytes;
    const int tvalid = (int)((n - t0) < TL_T ? (n - t0) : TL_T);
    const int own_lo = TL_HALO, own_hi = TL_HALO + tvalid;
    const DptVocabView& V = P.V;
    const bool spm = P.spm != 0;

    // ---- P0: load ------------------------------------------------------------------------------------
    {
        const bool aligned = (((uintptr_t)P.text) & 15u) == 0;
        for (int i = tid; i < TL_R / 16; i += nt) {
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
        for (int i = tid; i < 64; i += nt) S.text[TL_R + i] = 0;
        for (int w = tid; w < TL_NW + 2; w += nt) {
            S.mDS[w] = 0;
            S.mCX[w] = 0;
            if (w >= TL_NW) S.mCS[w] = S.mSP[w] = S.mM3[w] = S.mMK[w] = S.mCF[w] = S.mWS[w] = 0;
        }
        if (tid == 0) {
            S.d_first = (int32_t)tl_lower_bound(P.doc_offs, P.n_docs + 1, g0 < 0 ? 0 : g0);
            S.n_untok = 0;
            S.n_fallback = 0;
        }
    }
    blk.sync();

    // ---- P1a: byte-class masks, document starts --------------------------------------------------------
    for (int w = tid; w < TL_NW; w += nt) {
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
        if (o >= g0 + TL_R) break;
        const int r = (int)(o - g0);
        blk.atomic_or(&S.mDS[r >> 5], 1u << (r & 31));
    }
    blk.sync();

    // ---- P1b: word starts ------------------------------------------------------------------------------
    for (int w = tid; w < TL_NW; w += nt) {
        const uint32_t ds = S.mDS[w], dsn = S.mDS[w + 1], dsp = w ? S.mDS[w - 1] : 0u;
        const uint32_t cs = S.mCS[w], csn = (w + 1 < TL_NW) ? S.mCS[w + 1] : ~0u, csp = w ? S.mCS[w - 1] : 0u;
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
            cx = amb = mk & (pm | ds);
            // malformed UTF-8: continuation bytes glued to a space (the general path's character rule swallows
            // them into the marker) -> solve the word with the general code so both paths agree byte for byte
            cx |= sp & ~(((cs | ds) >> 1) | ((csn | dsn) << 31));
        } else {
            ws = ds;  // byte-level rules add their own word starts (tl_rule_*), documents always split
            cx = 0;
        }
        S.mMK[w] = mk;
        S.mCF[w] = cs | ds;
        S.mWS[w] = ws;
        S.mCX[w] = cx;
        const uint32_t rm = tl_range_mask(w, own_lo, own_hi);
        S.key[w] = (uint32_t)tl_popc(ws & rm);  // words per mask word (scratch use of key[])
        // SPM_LLAMA: a marker right after a marker makes the reference's split depend on the BPE merge order
        amb &= rm;
        while (amb && P.doc_flags) {
            const int r = (w << 5) + tl_ctz(amb);
            amb &= amb - 1;
            const int64_t d = tl_upper_bound(P.doc_offs, P.n_docs + 1, g0 + r) - 1;
            if (d >= 0 && d < P.n_docs) P.doc_flags[d] = 1;  // DPT_DF_AMBIGUOUS
        }
    }
    blk.sync();

    // ---- word list ---------------------------------------------------------------------------------------
    {
        const int chunk = (TL_NW + nt - 1) / nt;
        const int w0 = tid * chunk, w1 = (w0 + chunk) < TL_NW ? (w0 + chunk) : TL_NW;
        uint32_t mine = 0;
        for (int w = w0; w < w1; ++w) mine += S.key[w];
        uint32_t total;
        uint32_t off = blk.exclusive_scan(mine, S.scan, total);
        for (int w = w0; w < w1; ++w) {
            uint32_t bits = S.mWS[w] & tl_range_mask(w, own_lo, own_hi);
            while (bits) {
                S.wlist[off++] = (uint16_t)((w << 5) + tl_ctz(bits));
                bits &= bits - 1;
            }
        }
        if (tid == 0) {
            S.n_words_tile = (int32_t)total;
            S.last_end = tl_mask_next(S.mWS, own_hi, TL_R);
        }
    }
    blk.sync();
    const int nw = S.n_words_tile;

    // ---- P2: trie walks, one thread per byte position -------------------------------------------------------
    if (nw > 0) {
        const int lo = S.wlist[0], hi = S.last_end;
        for (int r = lo + tid; r < hi; r += nt) {
            const bool is_ws = tl_bit(S.mWS, r), is_ds = tl_bit(S.mDS, r), is_mk = tl_bit(S.mMK, r);
            const bool is_cf = tl_bit(S.mCF, r);
            uint32_t e = 0;
            if (spm) {
                const int we = tl_mask_next(S.mWS, r + 1, TL_R);
                if (is_ws && !is_ds) {  // initial marker: walk the body behind it from the marker node
                    const int b = r + (S.text[r] == 0x20u ? 1 : 3);
                    int limit = we - b;
                    if (limit < 0) limit = 0;
                    const bool capped = limit > TL_WALK_BITS - 1;
                    if (capped) limit = TL_WALK_BITS - 1;
                    bool alive;
                    e = tl_walk(S, P, V.marker_entry, b, limit, alive) << 1;
                    if (V.marker_entry & DPT_DA_TERMINAL) e |= 1u;
                    if (capped && alive) blk.atomic_or(&S.mCX[r >> 5], 1u << (r & 31));
                } else if (is_cf && !is_mk) {  // body character
                    int limit = we - r;
                    const bool capped = limit > TL_WALK_BITS;
                    if (capped) limit = TL_WALK_BITS;
                    bool alive;
                    e = tl_walk(S, P, DPT_DA_ROOT_ENTRY, r, limit, alive);
                    const int clen = tl_mask_next(S.mCF, r + 1, we) - r;
                    const bool in_vocab = clen <= TL_WALK_BITS && ((e >> (clen - 1)) & 1u);
                    if (!in_vocab || (capped && alive)) blk.atomic_or(&S.mCX[r >> 5], 1u << (r & 31));
                }
            } else {
                const int we = tl_mask_next(S.mWS, r + 1, TL_R);
                int limit = we - r;
                const bool capped = limit > TL_WALK_BITS;
                if (capped) limit = TL_WALK_BITS;
                bool alive;
                e = tl_walk(S, P, DPT_DA_ROOT_ENTRY, r, limit, alive);
                if (capped && alive) blk.atomic_or(&S.mCX[r >> 5], 1u << (r & 31));
            }
            S.E[r] = e;
            S.key[r] = TL_KEY_INF;
        }
    }
    blk.sync();

    // ---- P3: forward DP, one thread per word ----------------------------------------------------------------
    const int wpt = nw > 0 ? (nw + nt - 1) / nt : 0;
    const int k0 = tid * wpt, k1 = (k0 + wpt) < nw ? (k0 + wpt) : nw;
    uint32_t my_tot = 0;  // words << 16 | tokens
    for (int k = k0; k < k1; ++k) {
        const int ws = S.wlist[k];
        const bool ds = tl_bit(S.mDS, ws);
        int we = tl_mask_next(S.mWS, ws + 1, TL_R);
        const bool open = we >= TL_R;
        const int ml = !spm ? 0 : ds ? 0 : (S.text[ws] == 0x20u ? 1 : 3);
        const int b = ws + ml;
        bool complex_word = open || !V.fast_ok;
        if (!complex_word) {
            for (int w = ws >> 5; w <= (we - 1) >> 5 && !complex_word; ++w)
                complex_word = (S.mCX[w] & tl_range_mask(w, ws, we)) != 0;
        }
        uint32_t rec;
        if (!complex_word) {
            uint32_t kend = TL_KEY_INF, aend = 0, bend = 0;
            // pushes in ascending predecessor order: ties keep the LARGEST predecessor (dp_tokenize.py:58 explores
            // the largest split first)
            auto push = [&](int i, uint32_t knew, uint32_t dist) {
                if (i == we) {
                    if ((knew >> 8) <= (kend >> 8)) aend = dist;
                    if (knew <= kend) { kend = knew; bend = dist; }
                } else {
                    const uint32_t bi = S.key[i];
                    if ((knew >> 8) <= (bi >> 8)) S.A[i] = (uint8_t)dist;
                    if (knew <= bi) { S.key[i] = knew; S.B[i] = (uint8_t)dist; }
                }
            };
            if (spm) {
                uint32_t em;
                if (ml) {
                    em = S.E[ws];
                } else {  // virtual marker of a document's first word: nobody walked it in P2
                    int limit = we - b;
                    if (limit > TL_WALK_BITS - 1) limit = TL_WALK_BITS - 1;
                    bool alive;
                    em = (tl_walk(S, P, V.marker_entry, b, limit, alive) << 1) | ((V.marker_entry & DPT_DA_TERMINAL) ? 1u : 0u);
                    if (alive && we - b > TL_WALK_BITS - 1) complex_word = true;
                }
                while (em && !complex_word) {
                    const int kbit = tl_ctz(em);
                    em &= em - 1;
                    const int i = b + kbit;
                    const uint32_t cl = 1u + (kbit ? (uint32_t)tl_count(S.mCF, b, i) : 0u);
                    push(i, tl_key_extend(0xFFu, cl), 0u);
                }
            } else {
                S.key[b] = 0xFFu;  // origin
            }
            if (!complex_word) {
                for (int j = b; j < we; j = spm ? tl_mask_next(S.mCF, j + 1, we) : j + 1) {
                    const uint32_t kj = S.key[j];
                    uint32_t ej = S.E[j];
                    while (ej) {
                        const int kbit = tl_ctz(ej);
                        ej &= ej - 1;
                        const int i = j + kbit + 1;
                        const uint32_t cl = spm ? (uint32_t)tl_count(S.mCF, j, i) : (uint32_t)(kbit + 1);
                        push(i, tl_key_extend(kj, cl), (uint32_t)(kbit + 1));
                    }
                }
                if (kend == TL_KEY_INF) {  // cannot happen for well-formed input; never emit garbage
                    complex_word = true;
                } else {
                    const uint32_t ntok = kend >> 8, longest = 0xFFu - (kend & 0xFFu);
                    rec = ntok | (longest << 13) | (aend << 19) | (bend << 25);
                    my_tot += (1u << 16) | ntok;
                }
            }
        }
        if (complex_word) {
            int64_t g_ws = g0 + ws, g_we = g0 + we;
            bool ok = true;
            if (open) {
                g_we = spm ? tl_spm_word_end_global(P, g_ws, ml, ds, TL_LA + TL_BIGTAIL) : -1;
                ok = g_we >= 0;
            }
            if (ok) {
                rec = tl_complex_forward(P, cta, 10 * (ws - TL_HALO), g_ws, g_we, spm && ds);
            } else {
                rec = TL_REC_COMPLEX | TL_REC_UNTOK;
                blk.atomic_add(&S.n_fallback, 1u);
            }
            if (rec & TL_REC_UNTOK) {
                blk.atomic_add(&S.n_untok, 1u);
                my_tot += (1u << 16);
            } else {
                my_tot += (1u << 16) | (rec & 0xFFFFu);
            }
        }
        if (spm && ds) {  // the '<s>' word in front of every document (tokenizer_utils.py:26-30)
            my_tot += (1u << 16) | (uint32_t)V.bos_ntok;
            if (!V.bos_ntok) blk.atomic_add(&S.n_untok, 1u);
        }
        S.E[ws] = rec;
    }

    // ---- P4: offsets inside the tile, then across tiles (decoupled look-back) -------------------------------
    {
        uint32_t total;
        uint32_t off = blk.exclusive_scan(my_tot, S.scan, total);
        for (int k = k0; k < k1; ++k) {
            const int ws = S.wlist[k];
            S.key[ws] = off;
            const uint32_t rec = S.E[ws];
            const uint32_t ntok = (rec & TL_REC_COMPLEX) ? ((rec & TL_REC_UNTOK) ? 0u : (rec & 0xFFFFu)) : (rec & 0x1FFFu);
            off += (1u << 16) | ntok;
            if (spm && tl_bit(S.mDS, ws)) off += (1u << 16) | (uint32_t)V.bos_ntok;
        }
        if (tid == 0) S.tile_tot = total;
    }
    blk.sync();
    blk.lookback(P, S, tile);
    blk.sync();
    const unsigned long long base_w = S.base_w, base_t = S.base_t;

    // ---- P5: select + emit, one thread per word --------------------------------------------------------------
    for (int k = k0; k < k1; ++k) {
        const int ws = S.wlist[k];
        const uint32_t rec = S.E[ws], off = S.key[ws];
        int64_t gw = (int64_t)base_w + (off >> 16);
        int64_t gt = (int64_t)base_t + (off & 0xFFFFu);
        const bool ds = tl_bit(S.mDS, ws);
        if (spm && ds) {
            const int64_t d = tl_lower_bound(P.doc_offs, P.n_docs + 1, g0 + ws);
            if (d < P.n_docs) P.doc_tok_offs[d] = gt;
            if (gw < P.word_cap) {
                P.word_lens[gw] = V.bos_len;
                P.word_flags[gw] = V.bos_ntok ? 0 : 1;  // DPT_WF_UNTOKENIZABLE
            }
            for (int q = 0; q < V.bos_ntok; ++q)
                if (gt + q < P.ids_cap) P.ids[gt + q] = V.bos_ids[q];
            gw += 1;
            gt += V.bos_ntok;
        } else if (!spm && ds) {
            const int64_t d = tl_lower_bound(P.doc_offs, P.n_docs + 1, g0 + ws);
            if (d < P.n_docs) P.doc_tok_offs[d] = gt;
        }
        if (rec & TL_REC_COMPLEX) {
            if (gw < P.word_cap) {
                P.word_lens[gw] = (int32_t)(rec & 0x3FFFFFFFu);
                P.word_flags[gw] = (uint8_t)(((rec & TL_REC_UNTOK) ? 1u : 0u) | 4u);  // | DPT_WF_LONG: general code path
            }
            if (!(rec & TL_REC_UNTOK)) tl_complex_emit(P, cta, 10 * (ws - TL_HALO), gt);
            continue;
        }
        const int we = tl_mask_next(S.mWS, ws + 1, TL_R);
        const int ml = !spm ? 0 : ds ? 0 : (S.text[ws] == 0x20u ? 1 : 3);
        const int b = ws + ml;
        const int ntok = (int)(rec & 0x1FFFu);
        const uint32_t target = (rec >> 13) & 63u;
        if (gw < P.word_cap) {
            P.word_lens[gw] = ntok;
            P.word_flags[gw] = 0;
        }
        bool got = false;
        int i = we;
        for (int s = ntok - 1; s >= 0; --s) {
            uint32_t d;
            if (i == we) d = got ? ((rec >> 19) & 63u) : ((rec >> 25) & 63u);
            else d = got ? S.A[i] : S.B[i];
            int32_t id;
            uint32_t cl;
            if (spm && d == 0) {  // token = marker + body[b, i)
                id = tl_token_id(S, P, true, b, i);
                cl = 1u + (i > b ? (uint32_t)tl_count(S.mCF, b, i) : 0u);
            } else {
                const int j = i - (int)d;
                id = tl_token_id(S, P, false, j, i);
                cl = spm ? (uint32_t)tl_count(S.mCF, j, i) : d;
                i = j;
            }
            if (!got && cl == target) got = true;
            if (gt + s < P.ids_cap) P.ids[gt + s] = id;
        }
    }

    // ---- tile epilogue ---------------------------------------------------------------------------------------
    if (tid == 0) {
        if (S.n_untok) blk.atomic_add_u64(&P.counters[3], (unsigned long long)S.n_untok);
        if (S.n_fallback) blk.atomic_add_u64((unsigned long long*)&P.n_out[DPT_NOUT_FALLBACK_IDX], (unsigned long long)S.n_fallback);
        if (tile == P.n_tiles - 1) {
            const unsigned long long tw = base_w + (S.tile_tot >> 16), tt = base_t + (S.tile_tot & 0xFFFFu);
            P.counters[0] = (unsigned long long)n;
            P.counters[1] = tw;
            P.counters[2] = tt;
            P.n_out[0] = (int64_t)tt;  // DPT_NOUT_IDS
            P.n_out[1] = (int64_t)tw;  // DPT_NOUT_WORDS
            P.doc_tok_offs[P.n_docs] = (int64_t)tt;
        }
    }
    blk.sync();
}

// persistent loop: tiles are handed out in corpus order by an atomic ticket, so a tile only ever waits (in the
// look-back) for tiles held by CTAs that are already running
template <class Blk>
DPT_TILE_FN void tl_loop(Blk& blk, const TileParams& P, TileSmem& S, int64_t cta) {
    tl_stage_trie(blk, P, S);
    for (;;) {
        if (blk.tid() == 0) S.tile = (int32_t)blk.take_ticket(P.ticket);
        blk.sync();
        const int tile = S.tile;
        if (tile >= P.n_tiles) break;
        tl_run_tile(blk, P, S, tile, cta);
    }
}

}  // namespace dpt
