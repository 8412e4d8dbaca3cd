// Corpus pipeline of the shortest-tokenization path, host/device source.
//
//   A  scan + dedup   one CTA per 3968-byte tile of RAW corpus bytes (4 KB region with halo and look-ahead): coalesced 16-byte loads into shared memory,
//                     one-bit-per-byte class masks, the tokenizer's boundary rule -> word starts, and for every
//                     word one probe of an HBM/L2-resident hash table keyed by the word's bytes.  The first
//                     occurrence claims the slot with one 64-bit CAS (the tag holds hash, length and the byte
//                     offset of that occurrence, so later occurrences verify against the immutable corpus text
//                     and never wait on another thread).  Output: one 32-bit ref per word, in corpus order
//                     (decoupled look-back over per-tile word counts), and the list of distinct words.
//   B  DP             ONE shortest-tokenization DP per DISTINCT word (dp_tokenize.py:24-84 in the closed form of
//                     dpt_dp_core.h): normalise (U+2581 marker, "<0xHH>" expansion of out-of-vocabulary
//                     characters), forward DP with the reference's tie order, backward select, ids -> the
//                     word's 32-byte result record (one sector: count + up to 7 ids inline).
//   C  emit           one thread per 8 words: ref -> result record -> token count; block scan + decoupled
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
#include "dpt_split_rules.h"

#if defined(__CUDACC__)
#define DPT_PIPE_FN __device__ __forceinline__
#else
#define DPT_PIPE_FN inline
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
#endif

namespace dpt {

#ifndef DPT_PA_T
#define DPT_PA_T 6016  // byte-level rules (PaGeom below has the SentencePiece tile): 6144-byte regions, 57 KB of dynamic shared memory
#endif
#ifndef DPT_PA_THREADS
#define DPT_PA_THREADS 256
#endif
#ifndef DPT_PA_WIN
#define DPT_PA_WIN 1152
#endif
constexpr int PA_T = DPT_PA_T;                // raw bytes per tile of kernel A
constexpr int PA_HALO = 32;                   // look-behind (multiple of 32 keeps mask words aligned)
constexpr int PA_LA = 96;                     // look-ahead: a word that ends within it is handled in-tile
constexpr int PA_R = PA_HALO + PA_T + PA_LA;  // region bytes, a multiple of 32 (mask words) and of 16 (one 16-byte load per thread and round)
static_assert(PA_HALO == 32 && PA_LA == 96, "PaGeom spells the halo and the look-ahead out");
constexpr int PA_NW = PA_R / 32;
constexpr int PA_THREADS = DPT_PA_THREADS;
// Tile geometry per boundary rule.  The byte-level kernels take 6016-byte tiles (6144-byte regions, windows of 1152 words,
// 3 CTAs/SM at 64 registers: their shared memory - code array, sync-point and document lists - grows with the region, 57 KB
// here; 1.056 -> 0.974 ms per 100 MB of the Llama-3 mix, 1.003 -> 0.890 on en/de pairs; 8064-byte tiles: 0.999 / 0.948;
// round 1's 3968-byte tiles at 4 CTAs/SM were the best the 48 KB static limit allowed).  The SentencePiece kernel takes 8064-byte tiles
// (8192-byte region) with windows of 1152 words at 5 CTAs/SM and 48 registers: the per-tile work (document search,
// look-back, scans, barriers) is amortised over twice the bytes and nothing spills - 0.378 -> 0.352 ms on the B200; 6016
// and 11136-byte tiles, 4 or 6 CTAs/SM and 512 threads were all slower (profiles/r2_variants_tile_size.txt).  PA_T / PA_R /
// PA_NW / PA_WIN below are the byte-level geometry and, for the host's sizing, the upper bound of the tile count.
#ifndef DPT_PA_T_SPM
#define DPT_PA_T_SPM 8064
#endif
#ifndef DPT_PA_WIN_SPM
#define DPT_PA_WIN_SPM 1152
#endif
template <bool kSpm>
struct PaGeom {
    static constexpr int T = kSpm ? DPT_PA_T_SPM : DPT_PA_T;
    static constexpr int R = 32 + T + 96;  // PA_HALO + T + PA_LA
    static constexpr int NW = R / 32;
    static constexpr int WIN = kSpm ? DPT_PA_WIN_SPM : DPT_PA_WIN;
};
constexpr int PA_MAXLEN = 63;                 // longest word body (bytes) that goes through the dedup table
constexpr int PA_PROBES = 32;                // probes before a word gives up on the table and becomes an odd word (a run-time loop:
                                              // nearly every word needs one or two).  16 lost 0.9 % of the words of an all-distinct
                                              // corpus at a load of 0.45 - linear probing clusters - and their per-occurrence DP
                                              // (k_dp_distinct, 2.1 ms) was that pass's critical path; 8 lost 5e-4 at a load of 0.3
constexpr int PB_THREADS = 128;
constexpr int PB_CLASSES = 5;                 // length classes of the DP work queues (8 measured no better: lane
                                              // imbalance comes from walk depths, not from word length)
constexpr int PB_LOCAL = 72;                  // normalised bytes solved with per-thread local state in kernel B
// kernel B: finished lanes take new words once fewer lanes than this are still walking.  1 = a warp finishes its 32
// words before it takes the next 32.  Measured on the B200 (100 MB S2ORC-shaped corpus): 1 -> 0.350 ms, 12 -> 0.404,
// 20 -> 0.438, 28 -> 0.526: every refill runs the long normalise / initialise / backward code for a few lanes while
// the others wait, which costs more than the idle lanes of the hot loop
#ifndef DPT_PB_REFILL
#define DPT_PB_REFILL 1
#endif
constexpr int PB_REFILL = DPT_PB_REFILL;
#ifndef DPT_PC_THREADS
#define DPT_PC_THREADS 256
#endif
#ifndef DPT_PC_STAGE
#define DPT_PC_STAGE 6144
#endif
constexpr int PC_THREADS = DPT_PC_THREADS;
constexpr int PC_PER = 8;                     // words per thread in kernel C
constexpr int PA_WIN = DPT_PA_WIN;                  // words of a tile handled per pass.  A tile holds ~500-540 words: with 512 most
                                              // byte-level tiles needed a second, nearly empty pass and resolved their look-back
                                              // BEFORE probing (768: k_scan_dedup_bl 1.44 -> 1.38 / 1.46 -> 1.31 ms, SPM unchanged)
constexpr int PC_TILE = PC_THREADS * PC_PER;
constexpr int PC_STAGE = DPT_PC_STAGE;                // ids of one kernel-C tile staged in shared memory (more -> direct writes)

// one 32-bit ref per word: top two bits 11 = the '<s>' word in front of an SPM_LLAMA document | document index,
// 10 = index into the odd-word list (not deduplicated), otherwise the word's table slot
// (byte-level rules have no '<s>' word: bit 29 marks the first word of a document instead)
constexpr uint32_t REF_BOS = 0xC0000000u;
constexpr uint32_t REF_ODD = 0x80000000u;
constexpr uint32_t REF_KIND = 0xC0000000u;
constexpr uint32_t REF_DOCFIRST = 0x20000000u;
constexpr uint32_t REF_INDEX = 0x1FFFFFFFu;
constexpr unsigned long long PD_MASK = (1ull << 62) - 1;

constexpr uint32_t RES_UNTOK = 1u << 24;      // result meta: word_len (24 bits) | flags
constexpr uint32_t RES_POOLED = 1u << 25;     // more than RES_INLINE ids: they live in the pool at ids[0] | ids[1] << 32
constexpr int RES_INLINE = 7;
constexpr int RES_HEAD = 3;                   // ids in the first 16 bytes of a record
constexpr uint32_t RES_LONG = 1u << 26;       // solved by the long-word kernel

// Result of one distinct word: ONE 32-byte sector, indexed by the word's table slot.  Kernel C gathers the first 16 bytes
// (meta + ids[0..2]: nine of ten word OCCURRENCES have at most three tokens) and, for words of 4..7 tokens, the second 16
// bytes of the SAME sector (an L1 hit behind the first load); only words of more than 7 tokens go through the id pool.
// (Round 2 tried 16-byte records with 3 ids inline: the sector footprint in L2 is the same - one sector per live record
// either way, hit rate 47 % -> 50 % - and the pool copy of every word of 4+ tokens, a serial loop at 2.5 of 32 lanes,
// became 38 % of kernel C's instructions: 0.160 -> 0.183 ms.)
struct alignas(32) ResRec {
    uint32_t meta;           // word_len (len_dp[n], 24 bits) | RES_* flags
    int32_t ids[RES_INLINE];
};

struct OddWord {
    int64_t pos;   // global offset of the word's first raw byte
    int32_t len;   // raw bytes
    int32_t virt;  // 1: first word of its document (gets the Prepend(U+2581) marker)
};

struct PipeCtl {  // device-side counters, zeroed by the launcher
    unsigned int ticket_a, ticket_c;
    unsigned int n_pending[PB_CLASSES];  // distinct words queued for the DP, by length class (keeps a warp's lanes alike)
    unsigned int n_odd, n_long;
    unsigned int n_defer, pad0;          // words the cooperative DP kernel handed to the thread-per-word kernel
    unsigned long long lp_used, n_words, n_untok, n_too_long;
    unsigned long long b_cursor;  // next unclaimed item of kernel B's work list (thread-per-word kernel, first launch)
    unsigned long long b_cursor2; // ... second launch (the words the lock-step kernel deferred)
    unsigned int lock_ticket, lock_ticket2;  // next unclaimed chunk of the lock-step kernels (<= 31 units / 32..63 units)
};

struct PipePersist {  // survives the launches of one chunked call (same lifetime as the word table)
    unsigned long long pool_used;
};

struct PipeParams {
    DptVocabView V;
    const uint8_t* text;      // the WHOLE corpus buffer; this launch tokenizes [byte_begin, byte_end)
    int64_t n_bytes;
    const int64_t* doc_offs;  // n_docs + 1, doc_offs[0] == 0, doc_offs[n_docs] == n_bytes (global offsets)
    int64_t n_docs;
    // the range of this launch: documents [doc_begin, doc_begin + n_docs_local) = bytes [byte_begin, byte_end).
    // Outputs are range-local (ids from 0, doc_tok_offs[d - doc_begin]); table tags hold corpus-global offsets, so a
    // chunked call can keep its word table across launches.
    int64_t byte_begin, byte_end, doc_begin, n_docs_local;
    PipePersist* persist;
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
    ResRec* res;                  // n_slots
    uint32_t* pending;            // 4 length classes x pend_stride (private to the range: ranges may overlap in time)
    OddWord* odd;                 // odd_cap
    ResRec* odd_res;              // odd_cap
    int32_t* pool;                // pool_cap ids of words with more than 7 tokens (persistent, like the table)
    uint32_t* longq;              // n_slots + odd_cap
    uint32_t* defer;              // pend_stride table slots (cooperative DP kernel -> thread-per-word kernel)
    int32_t coop;                 // work list of the thread-per-word kernel: 0 = everything (host emulation); 1 = odd words + words
                                  // of more than 31 units (beside the lock-step kernel); 2 = the words the lock-step kernel deferred
    uint8_t* lp_norm;             // long-word scratch: lp_cap positions
    uint64_t* lp_best;
    uint16_t* lp_a;
    uint16_t* lp_b;
    PipeCtl* ctl;
    unsigned long long* desc_w;   // n_tiles look-back descriptors of kernel A (zeroed)
    unsigned long long* desc_t;   // n_ctiles look-back descriptors of kernel C (zeroed)
    int64_t odd_cap, pool_cap, lp_cap;
    uint32_t slot_mask;           // n_slots - 1 (power of two, <= 2^30)
    int32_t n_tiles, n_ctiles;
    int32_t tile_first;           // global index (byte offset / PA_T) of this launch's first tile
    int64_t pend_stride;          // capacity of one length class of `pending`
    int32_t spm;  // 1: SPM_LLAMA rule; 0: byte-level rules
    int32_t rule;
    int32_t vec_ok;  // word_lens / word_flags are aligned for 16- / 8-byte stores
};

template <bool kSpm>
struct ASmemT {
    static constexpr int PA_R = PaGeom<kSpm>::R, PA_NW = PaGeom<kSpm>::NW, PA_WIN = PaGeom<kSpm>::WIN;  // (this rule's geometry)
    // list capacity: SPM needs one window of entries; byte-level rules also list their synchronisation points
    // (up to one per byte of the region)
    static constexpr int WL_CAP = kSpm ? PA_WIN : PA_R + 32;
    alignas(16) uint8_t text[PA_R + 64];
    uint32_t mDS[PA_NW + 2];  // document starts (and the end-of-text sentinel)
    uint32_t mCS[PA_NW + 2];  // code-point start bytes
    uint32_t mSP[PA_NW + 2];  // ' '
    uint32_t mM3[PA_NW + 2];  // E2 96 81 candidates
    uint32_t mWS[PA_NW + 2];  // word starts
    uint32_t mCX[PA_NW + 2];  // positions that make their word "odd" (solved from the raw text, not deduplicated)
    uint32_t mSY[PA_NW + 2];  // byte-level rules: synchronisation points of the split scanner
    uint32_t mAL[kSpm ? 2 : PA_NW + 2];  // byte-level rules: every byte of a letter (the split scanner skips over their runs)
    uint32_t mNW[kSpm ? 2 : PA_NW + 2];  // byte-level rules: positions whose character can follow a synchronising space
    alignas(16) uint8_t code[kSpm ? 16 : PA_R + 64];  // byte-level rules: dpt_char_code of the character at every position
    alignas(16) uint8_t lut[kSpm ? 16 : 128];         // ... and of the 128 ASCII characters
    uint32_t dsn[PA_NW + 2];  // byte-level rules: document starts in front of each mask word
    uint16_t wlist[WL_CAP];   // region index of the words of the current window (| 0x8000: a document's '<s>' word)
    uint16_t dslist[kSpm ? 2 : PA_R + 32];  // byte-level rules: region indices of the document starts, in order
    uint32_t pend[PA_WIN];    // table slots claimed in the current window
    uint32_t stage[PA_WIN];   // refs of the current window, written out coalesced once the word offset is known
    uint32_t scan[40];
    int32_t tile, d_first, n_entries;
    uint32_t any_cx;          // some mCX bit is set in this tile (rare: the per-word range test is skipped otherwise)
    uint32_t n_pend, n_pend_c[PB_CLASSES], cur_c[PB_CLASSES], base_c[PB_CLASSES];
    int32_t n_sync, n_ds;
    long long region_doc_end, first_sync_global;
    unsigned long long base_w;
};

struct CSmem {
    int32_t ids[PC_STAGE];
    uint32_t scan[40];
    int32_t tile;
    uint32_t tile_tot, n_untok;
    uint32_t n_docfirst;       // byte-level rules: documents that start in this tile
    long long doc0;            // ... and the index (in the range) of the first of them
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
DPT_HD int pp_clz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return __builtin_clz(x);
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

// 4 bytes at byte offset `off` of a 4-byte-aligned buffer (little endian), two aligned loads + funnel shift
DPT_HD uint32_t pp_load4(const uint8_t* base4, int64_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base4 + (off & ~(int64_t)3));
    const uint32_t lo = w[0], hi = w[1];
    const uint32_t sh = (uint32_t)(off & 3) * 8u;
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
// the first min(len,16) bytes at byte offset `off` of a 4-byte-aligned buffer as four little-endian words,
// zero-padded: five aligned loads + four funnel shifts, no loop
DPT_HD void pp_load16(const uint8_t* base4, int64_t off, int len, uint32_t v[4]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base4 + (off & ~(int64_t)3));
    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
    const uint32_t sh = (uint32_t)(off & 3) * 8u;
#if defined(__CUDA_ARCH__)
    v[0] = __funnelshift_r(a0, a1, sh);
    v[1] = __funnelshift_r(a1, a2, sh);
    v[2] = __funnelshift_r(a2, a3, sh);
    v[3] = __funnelshift_r(a3, a4, sh);
#else
    v[0] = sh ? (a0 >> sh) | (a1 << (32u - sh)) : a0;
    v[1] = sh ? (a1 >> sh) | (a2 << (32u - sh)) : a1;
    v[2] = sh ? (a2 >> sh) | (a3 << (32u - sh)) : a2;
    v[3] = sh ? (a3 >> sh) | (a4 << (32u - sh)) : a3;
#endif
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // keep min(max(len - 4k, 0), 4) bytes of word k
        int rem = len - 4 * k;
        rem = rem < 0 ? 0 : rem > 4 ? 4 : rem;
        v[k] &= (uint32_t)((1ull << (8 * rem)) - 1ull);
    }
}
// PP_TAIL_LUT[len]: byte masks of the first min(len,16) bytes of a 16-byte window (four little-endian words)
#define PP_TL(n) {(n) >= 4 ? 0xFFFFFFFFu : (n) <= 0 ? 0u : ((1u << (8 * ((n) > 0 && (n) < 4 ? (n) : 0))) - 1u),                 \
                  (n) >= 8 ? 0xFFFFFFFFu : (n) <= 4 ? 0u : ((1u << (8 * ((n) > 4 && (n) < 8 ? (n) - 4 : 0))) - 1u),             \
                  (n) >= 12 ? 0xFFFFFFFFu : (n) <= 8 ? 0u : ((1u << (8 * ((n) > 8 && (n) < 12 ? (n) - 8 : 0))) - 1u),           \
                  (n) >= 16 ? 0xFFFFFFFFu : (n) <= 12 ? 0u : ((1u << (8 * ((n) > 12 && (n) < 16 ? (n) - 12 : 0))) - 1u)}
#if defined(__CUDACC__)
static __device__ const uint4 PP_TAIL_LUT[17] = {
#else
static const uint4 PP_TAIL_LUT[17] = {
#endif
    PP_TL(0), PP_TL(1), PP_TL(2),  PP_TL(3),  PP_TL(4),  PP_TL(5),  PP_TL(6),  PP_TL(7), PP_TL(8),
    PP_TL(9), PP_TL(10), PP_TL(11), PP_TL(12), PP_TL(13), PP_TL(14), PP_TL(15), PP_TL(16)};
#undef PP_TL
DPT_PIPE_FN uint4 pp_tail_lut(int len) {
#if defined(__CUDA_ARCH__)
    return __ldg(&PP_TAIL_LUT[len < 16 ? len : 16]);
#else
    return PP_TAIL_LUT[len < 16 ? len : 16];
#endif
}
// 16 bytes at byte offset `off` of a 4-byte-aligned buffer, unmasked (the caller masks with pp_tail_lut(len))
DPT_HD void pp_load16_raw(const uint8_t* base4, int64_t off, uint32_t v[4]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base4 + (off & ~(int64_t)3));
    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
    const uint32_t sh = (uint32_t)(off & 3) * 8u;
#if defined(__CUDA_ARCH__)
    v[0] = __funnelshift_r(a0, a1, sh);
    v[1] = __funnelshift_r(a1, a2, sh);
    v[2] = __funnelshift_r(a2, a3, sh);
    v[3] = __funnelshift_r(a3, a4, sh);
#else
    v[0] = sh ? (a0 >> sh) | (a1 << (32u - sh)) : a0;
    v[1] = sh ? (a1 >> sh) | (a2 << (32u - sh)) : a1;
    v[2] = sh ? (a2 >> sh) | (a3 << (32u - sh)) : a2;
    v[3] = sh ? (a3 >> sh) | (a4 << (32u - sh)) : a3;
#endif
}
DPT_HD uint32_t pp_hash_step(uint32_t h, uint32_t v) {
    h = (h ^ v) * 0x9E3779B1u;
    return h ^ (h >> 15);
}

// 4-bit mask of the bytes of x equal to the bytes of c4 (exact SWAR zero-byte test)
DPT_HD uint32_t pp_eq4(uint32_t x, uint32_t c4) {
    const uint32_t z = x ^ c4;
    uint32_t t = (z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    t = ~(t | z | 0x7F7F7F7Fu);                 // 0x80 in every byte of z that is zero
    return ((t >> 7) * 0x01020408u) >> 24;      // gather the four flag bits
}
// 4-bit mask of the bytes of x that are ASCII letters (same test as dpt_char_at: ((b | 0x20) - 'a') < 26, b < 0x80)
DPT_HD uint32_t pp_letters4(uint32_t x) {
    const uint32_t y = (x | 0x20202020u) & 0x7F7F7F7Fu;
    const uint32_t ge = y + 0x1F1F1F1Fu;  // bit 7 of a byte: y >= 'a'
    const uint32_t gt = y + 0x05050505u;  // bit 7 of a byte: y > 'z'
    const uint32_t t = ge & ~gt & ~x & 0x80808080u;
    return ((t >> 7) * 0x01020408u) >> 24;
}
// the split scanners' accelerator (dpt_split_rules.h: DptNoSkip) over a tile's letter mask (bytes of ASCII letters and of
// the multi-byte letters the tile could classify: a run of set bits ends on a character boundary); region index = p - g0
struct PaLetterSkip {
    const uint32_t* mask;
    int64_t g0;
    DPT_HD int64_t ascii_letters(int64_t p, int64_t end) const {
        int64_t r = p - g0;
        if (r < 0) return p;  // in front of the region: the scanner reads global memory character by character
        while (p < end) {
            const int sh = (int)(r & 31), avail = 32 - sh;
            const uint32_t inv = ~(mask[r >> 5] >> sh);  // the bits shifted in from the top read as "not a letter"
            int n = inv ? pp_ctz(inv) : 32;
            if (n > avail) n = avail;
            p += n;
            r += n;
            if (n < avail) break;
        }
        return p < end ? p : end;
    }
};
// the split scanners' character source over a tile's code array (dpt_split_rules.h: DptTextSrc): index = global offset
struct PaCodeSrc {
    const uint8_t* code;  // S.code - g0
    const uint8_t* text;  // S.text - g0
    DPT_HD DptChar at(int64_t p, int64_t) const { return dpt_char_of_code(code[p]); }
    DPT_HD uint32_t byte(int64_t p) const { return text[p]; }
};
// Length class of a word by its `units` (body bytes, + 1 for the SPM marker): 0..2 -> the lock-step DP kernel with a
// 32-byte register window (the CTA sorts its words by exact length anyway; the classes only keep A's queues apart),
// 3 -> its 64-byte instantiation, 4 -> the thread-per-word kernel.
DPT_HD int pp_len_class(int units) {
    return units <= 8 ? 0 : units <= 16 ? 1 : units <= 31 ? 2 : units <= 63 ? 3 : 4;
}
DPT_HD uint32_t pp_tail_mask(int nbytes) { return nbytes >= 4 ? ~0u : ((1u << (8 * nbytes)) - 1u); }

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
template <class Blk, bool kSpm>
DPT_PIPE_FN void pa_run_tile(Blk& blk, const PipeParams& P, ASmemT<kSpm>& S, const int tile) {
    // this rule's tile geometry (the names shadow the namespace-level byte-level constants on purpose)
    constexpr int PA_T = PaGeom<kSpm>::T, PA_R = PaGeom<kSpm>::R, PA_NW = PaGeom<kSpm>::NW, PA_WIN = PaGeom<kSpm>::WIN;
    static_assert(PA_R % 32 == 0 && PA_R < 32768 && PA_WIN <= 65535, "region indices are 15 bits, counts 16");
    const int tid = blk.tid(), nt = blk.nthreads();
    const int64_t t0 = ((int64_t)P.tile_first + tile) * PA_T;
    const int64_t g0 = t0 - PA_HALO;  // global offset of region index 0
    const int64_t n = P.byte_end;  // nothing at or beyond the end of the range is looked at (it may not be there yet)
    const int tvalid = (int)((n - t0) < PA_T ? (n - t0) : PA_T);
    const int own_lo = PA_HALO + (int)(P.byte_begin > t0 ? P.byte_begin - t0 : 0), own_hi = PA_HALO + tvalid;
    constexpr bool spm = kSpm;
    // bits of mask word w that belong to this tile's own bytes [own_lo, own_hi) / to [0, own_hi): for a full tile
    // (all but the first and last of a range) these are whole words
    const bool full_tile = own_lo == PA_HALO && own_hi == PA_HALO + PA_T;
    auto own_mask = [&](int w) -> uint32_t {
        if (full_tile) return ((uint32_t)(w - PA_HALO / 32) < (uint32_t)(PA_T / 32)) ? ~0u : 0u;
        return pp_range_mask(w, own_lo, own_hi);
    };
    auto upto_mask = [&](int w) -> uint32_t {
        if (full_tile) return w < (PA_HALO + PA_T) / 32 ? ~0u : 0u;
        return pp_range_mask(w, 0, own_hi);
    };

    // ---- load: coalesced 16-byte loads of the region ------------------------------------------------------
    {
        const bool aligned = (((uintptr_t)P.text) & 15u) == 0;
        if (aligned && g0 >= 0 && g0 + PA_R <= n) {  // the whole region is there (every tile but the first and the last few)
            const uint4* src = reinterpret_cast<const uint4*>(P.text + g0);
            for (int i = tid; i < PA_R / 16; i += nt) *reinterpret_cast<uint4*>(&S.text[16 * i]) = src[i];
        } else if (!aligned && g0 >= 4 && g0 + PA_R + 8 <= n) {
            // a corpus buffer that does not start on a 16-byte boundary (a shard cut out of a larger buffer): five aligned
            // 4-byte loads and four funnel shifts per 16 bytes instead of sixteen single-byte loads
            const uint8_t* base4 = P.text - ((uintptr_t)P.text & 3u);
            const int64_t o4 = g0 + (int64_t)((uintptr_t)P.text & 3u);
            for (int i = tid; i < PA_R / 16; i += nt) {
                uint32_t v[4];
                pp_load16_raw(base4, o4 + 16 * (int64_t)i, v);
                *reinterpret_cast<uint4*>(&S.text[16 * i]) = uint4{v[0], v[1], v[2], v[3]};
            }
        } else {
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
        }
        // small fixed jobs go to the last threads of the block (the first warp is busy with the document search)
        if (tid >= nt - 4) *reinterpret_cast<uint4*>(&S.text[PA_R + 16 * (nt - 1 - tid)]) = uint4{0u, 0u, 0u, 0u};
        if (tid >= nt - 6 && tid < nt - 4) {  // the two mask words behind the region
            const int w = PA_NW + (nt - 5 - tid);
            S.mCS[w] = S.mSP[w] = S.mM3[w] = S.mWS[w] = S.mCX[w] = S.mSY[w] = 0;
        }
        if (!spm) {
            for (int w = tid; w < PA_NW + 2; w += nt) {
                S.mSY[w] = 0;
                S.mWS[w] = 0;
                S.mAL[w] = 0;
            }
            for (int i = tid; i < 128 + 4; i += nt) {
                if (i < 128) {  // codes of the ASCII characters (conflict-free look-ups in the mask pass: one LUT word per bank)
                    const uint8_t b = (uint8_t)i;
                    S.lut[i] = (uint8_t)dpt_char_code(dpt_char_at(DptUniView{P.V.uni1, P.V.uni2}, &b, 0, 1));
                } else {
                    *reinterpret_cast<uint4*>(&S.code[PA_R + 16 * (i - 128)]) = uint4{0u, 0u, 0u, 0u};
                }
            }
        }
        // Document starts of the region, by the first warp alone while the text loads are in flight: a 32-ary search
        // for the first document (4 rounds of loads instead of the 16 dependent ones of a binary search, which was the
        // longest thing in front of the first barrier: ncu v12), then one offset per lane.
        if (blk.in_first_warp()) {
            const int lane = blk.lane(), ww = blk.warp_width();
            for (int w = lane; w < PA_NW + 2; w += ww) S.mDS[w] = 0;
            blk.reconverge();
            const int64_t d_first = blk.warp_lower_bound(P.doc_offs, P.n_docs + 1, g0 < 0 ? 0 : g0);
            for (int64_t k = d_first + lane; k <= P.n_docs; k += ww) {
                const int64_t o = P.doc_offs[k];
                if (o >= g0 + PA_R) break;
                const int r = (int)(o - g0);
                blk.atomic_or(&S.mDS[r >> 5], 1u << (r & 31));
            }
            if (tid == 0) {
                S.d_first = (int32_t)d_first;
                S.n_pend = 0;
                S.any_cx = 0;
                for (int c = 0; c < PB_CLASSES; ++c) S.n_pend_c[c] = S.cur_c[c] = 0;
            }
        }
    }
    blk.sync();

    // ---- byte-class masks (one bit per byte), document starts -----------------------------------------------
    for (int hw = tid; hw < 2 * PA_NW; hw += nt) {  // 16 bytes -> 16 mask bits per thread
        const uint8_t* t = &S.text[16 * hw];
        const uint4 x = *reinterpret_cast<const uint4*>(t);
        uint32_t cs = 0xFFFFu, sp = 0, m3 = 0;
        sp = pp_eq4(x.x, 0x20202020u) | (pp_eq4(x.y, 0x20202020u) << 4) | (pp_eq4(x.z, 0x20202020u) << 8) |
             (pp_eq4(x.w, 0x20202020u) << 12);
        if ((x.x | x.y | x.z | x.w) & 0x80808080u) {  // non-ASCII bytes: continuation bytes, raw U+2581 candidates
            const uint32_t cont = pp_eq4(x.x & 0xC0C0C0C0u, 0x80808080u) | (pp_eq4(x.y & 0xC0C0C0C0u, 0x80808080u) << 4) |
                                  (pp_eq4(x.z & 0xC0C0C0C0u, 0x80808080u) << 8) | (pp_eq4(x.w & 0xC0C0C0C0u, 0x80808080u) << 12);
            cs = ~cont & 0xFFFFu;
            if (spm) {  // raw U+2581 candidates: only where the lead byte E2 occurs
                uint32_t e2 = pp_eq4(x.x, 0x01010101u * DPT_MARK0) | (pp_eq4(x.y, 0x01010101u * DPT_MARK0) << 4) |
                              (pp_eq4(x.z, 0x01010101u * DPT_MARK0) << 8) | (pp_eq4(x.w, 0x01010101u * DPT_MARK0) << 12);
                while (e2) {
                    const int k = pp_ctz(e2);
                    e2 &= e2 - 1;
                    m3 |= (uint32_t)(t[k + 1] == DPT_MARK1 && t[k + 2] == DPT_MARK2) << k;
                }
            }
        }
        reinterpret_cast<uint16_t*>(S.mCS)[hw] = (uint16_t)cs;
        reinterpret_cast<uint16_t*>(S.mSP)[hw] = (uint16_t)sp;
        reinterpret_cast<uint16_t*>(S.mM3)[hw] = (uint16_t)(spm ? m3 : 0u);
        if (!spm) {
            // What the split scanners will ask about every position of these 16 bytes, answered once, by all threads at
            // once: the character's class, length and the three code-point tests (dpt_char_code).  The scanners run one
            // thread per stretch, so every instruction a character costs THEM is paid at a fraction of a warp's lanes:
            // decoding UTF-8 and walking the two-stage class table there was most of this kernel (ncu, round 2: 758 M
            // warp-instructions per 100 MB at 16.7 lanes).
            const bool bloom = P.rule == 4;
            uint32_t cw[4] = {0u, 0u, 0u, 0u};
            uint32_t nw = 0;
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
            // ASCII bytes: one conflict-free LUT look-up each (a non-ASCII byte looks up its low 7 bits and is patched below)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t c = S.lut[(xs[k >> 2] >> (8 * (k & 3))) & 0x7Fu];
                cw[k >> 2] |= c << (8 * (k & 3));
                nw |= (uint32_t)((c & 3u) != DPT_CLS_S && !(bloom && (c & DPT_CODE_BX))) << k;
            }
            // al: 16 bits + the spill of a letter that runs into the next 16 bytes
            uint32_t al = pp_letters4(x.x) | (pp_letters4(x.y) << 4) | (pp_letters4(x.z) << 8) | (pp_letters4(x.w) << 12);
            uint32_t hi = ((((x.x & 0x80808080u) >> 7) * 0x01020408u) >> 24) | (((((x.y & 0x80808080u) >> 7) * 0x01020408u) >> 24) << 4) |
                          (((((x.z & 0x80808080u) >> 7) * 0x01020408u) >> 24) << 8) | (((((x.w & 0x80808080u) >> 7) * 0x01020408u) >> 24) << 12);
            if (hi) {  // non-ASCII bytes, one at a time (ONE copy of the decoder in the kernel: unrolled it thrashed the instruction cache)
                const DptUniView U{P.V.uni1, P.V.uni2};
                const int r0 = 16 * hw;
                // document starts at r0 + 1 .. r0 + 18: a character cut by a document boundary is malformed
                const uint64_t dsw = (((uint64_t)S.mDS[r0 >> 5] | ((uint64_t)S.mDS[(r0 >> 5) + 1] << 32)) >> (r0 & 31)) >> 1;
                uint32_t c4[4] = {0u, 0u, 0u, 0u};  // codes of the non-ASCII positions, and the bytes they replace
                uint32_t keep = 0xFFFFu;
                do {
                    const int k = pp_ctz(hi);
                    const uint32_t near = (uint32_t)(dsw >> k) & 7u;  // document starts at r + 1, r + 2, r + 3
                    const int lim = near ? pp_ctz(near) + 1 : 4;
                    DptChar ch;
                    uint32_t c;
                    const uint32_t b0 = t[k], b1 = t[k + 1];
                    if ((b0 & 0xE0u) == 0xC0u && lim >= 2 && (b1 & 0xC0u) == 0x80u) {
                        // a 2-byte character (Latin-1 .. Arabic): its code straight from the table by code point - the same
                        // value dpt_char_code(dpt_char_at(...)) gives, without the decode and the two-stage class look-up
                        c = P.V.code2[((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu)];
                        ch.len = 2;
                        ch.cls = c & 3u;
                        ch.cp = 0;
                    } else {
                        ch = dpt_char_at(U, t, k, k + lim);
                        c = dpt_char_code(ch);
                    }
                    const uint32_t span = ((1u << ch.len) - 1u) << k;  // the bytes of this character: the scanners only ever
                    hi &= ~span;                                       // ask about its first
                    keep &= ~(1u << k);
#pragma unroll
                    for (int q = 0; q < 4; ++q)  // (static indices: c4 stays in registers)
                        if ((k >> 2) == q) c4[q] |= c << (8 * (k & 3));
                    const uint32_t isnw = (uint32_t)((c & 3u) != DPT_CLS_S && !(bloom && (c & DPT_CODE_BX)));
                    nw = (nw & ~(1u << k)) | (isnw << k);
                    if (ch.cls == DPT_CLS_L) al |= span;
                } while (hi);
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // byte-select: patched positions take c4, the others keep the LUT code
                    const uint32_t kb = (keep >> (4 * q)) & 15u;
                    const uint32_t km = ((kb & 1u) * 0xFFu) | ((kb & 2u) * (0xFF00u >> 1)) | ((kb & 4u) * (0xFF0000u >> 2)) | ((kb & 8u) * (0xFF000000u >> 3));
                    cw[q] = (cw[q] & km) | (c4[q] & ~km);
                }
            }
            *reinterpret_cast<uint4*>(&S.code[16 * hw]) = uint4{cw[0], cw[1], cw[2], cw[3]};
            reinterpret_cast<uint16_t*>(S.mNW)[hw] = (uint16_t)nw;
            // letters reach up to 3 bytes into the next thread's bits: everyone ORs into the zeroed mask
            if (al) {
                const uint64_t a64 = (uint64_t)al << (16 * (hw & 1));
                blk.atomic_or(&S.mAL[hw >> 1], (uint32_t)a64);
                if (a64 >> 32) blk.atomic_or(&S.mAL[(hw >> 1) + 1], (uint32_t)(a64 >> 32));
            }
        }
    }
    blk.sync();

    // ---- boundary rule -> word starts ---------------------------------------------------------------------
    const int chunk = (PA_NW + nt - 1) / nt;  // mask words per thread: [w0, w1)
    const int w0 = tid * chunk, w1 = (w0 + chunk) < PA_NW ? (w0 + chunk) : PA_NW;
    uint32_t my_cnt = 0;  // entries that start in this thread's mask words | document starts up to them << 16
    if (!spm) {
        // Byte-level rules (dpt_split_rules.h).  (0) for every mask word the next document start at or after it
        // (documents are long: without this every "where does this document end" question scans the whole mask).
        const DptUniView U{P.V.uni1, P.V.uni2};
        {
            const int chunk = (PA_NW + nt - 1) / nt;
            const int w0 = tid * chunk, w1 = (w0 + chunk) < PA_NW ? (w0 + chunk) : PA_NW;
            uint32_t mine = 0;
            for (int w = w0; w < w1; ++w) mine += (uint32_t)pp_popc(S.mDS[w]);
            uint32_t total;
            uint32_t off = blk.exclusive_scan(mine, S.scan, total);
            for (int w = w0; w < w1; ++w) {
                S.dsn[w] = off;  // document starts in front of mask word w
                uint32_t bits = S.mDS[w];
                while (bits) {
                    S.dslist[off++] = (uint16_t)((w << 5) + pp_ctz(bits));
                    bits &= bits - 1;
                }
            }
            if (tid == 0) {
                S.n_ds = (int32_t)total;
                const int64_t di = (int64_t)S.d_first + (int64_t)total;
                S.region_doc_end = di <= P.n_docs ? P.doc_offs[di] : n;  // end of the document open at the region's end
            }
        }
        blk.sync();
        const int n_ds = S.n_ds;
        // next document start strictly after region index r, or PA_R
        auto next_ds = [&](int r) -> int {
            const int w = r >> 5;
            const uint32_t hi = (r & 31) == 31 ? 0u : (S.mDS[w] & (~0u << ((r & 31) + 1)));
            if (hi) return (w << 5) + pp_ctz(hi);
            const int k = (int)S.dsn[w] + pp_popc(S.mDS[w]);
            return k < n_ds ? (int)S.dslist[k] : PA_R;
        };
        // (1) synchronisation points: document starts and every space whose next character is a non-whitespace
        // character of the same document
        for (int w = tid; w < PA_NW; w += nt) {
            // a space, the next position in the same document, and a character there that is no whitespace (BLOOM: a
            // character of the class) - dpt_is_sync_space, 32 positions at a time on the masks of the code pass
            const uint32_t ds = S.mDS[w], dsn = S.mDS[w + 1];
            const uint32_t nwn = w + 1 < PA_NW ? S.mNW[w + 1] : 0u;
            uint32_t sy = S.mSP[w] & ((S.mNW[w] >> 1) | (nwn << 31)) & ~((ds >> 1) | (dsn << 31));
            // within 6 bytes of the region's end the next character may not be loaded completely: no sync, scanned through
            sy &= pp_range_mask(w, 0, PA_R - 5);
            S.mSY[w] = sy | ds;
            S.mCX[w] = 0;
        }
        blk.sync();
        // (2) the stretch that covers the first byte of the tile starts at the last sync point at or before it: every
        // thread finds it for itself in the one or two mask words of the look-behind (a thread-0 section and a barrier
        // here kept 255 threads waiting)
        int sf = -1;
        for (int w = own_lo >> 5; w >= 0 && sf < 0; --w) {
            const uint32_t m = S.mSY[w] & pp_range_mask(w, 0, own_lo + 1);
            if (m) sf = (w << 5) + 31 - pp_clz(m);
        }
        if (sf < 0) {  // (block-uniform) no sync point in the look-behind: walk back through the text (rare: a piece-free run > 32 B)
            if (tid == 0) {
                const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, t0) - 1;
                const int64_t dstart = P.doc_offs[d < 0 ? 0 : d];
                const int nd = next_ds(own_lo);
                const int64_t dend = nd < PA_R ? g0 + nd : (int64_t)S.region_doc_end;
                int64_t q = g0 - 1;
                while (q > dstart && !dpt_is_sync_space(P.rule, U, P.text, q, dend)) --q;
                S.first_sync_global = q < dstart ? dstart : q;
            }
            blk.sync();
        }
        // (3) one thread per stretch between consecutive sync points: sequential regex scanner, marks piece starts
        {
            const int chunk = (PA_NW + nt - 1) / nt;
            // stretch list = sync bits in (max(sf,-1) .. PA_R); the first stretch starts at sf (or in global memory)
            const int w0 = tid * chunk, w1 = (w0 + chunk) < PA_NW ? (w0 + chunk) : PA_NW;
            uint32_t mine = 0;
            for (int w = w0; w < w1; ++w) mine += (uint32_t)pp_popc(S.mSY[w] & pp_range_mask(w, sf < 0 ? 0 : sf, PA_R));
            uint32_t total;
            uint32_t off = blk.exclusive_scan(mine, S.scan, total);
            for (int w = w0; w < w1; ++w) {
                uint32_t bits = S.mSY[w] & pp_range_mask(w, sf < 0 ? 0 : sf, PA_R);
                while (bits) {
                    S.wlist[off++] = (uint16_t)((w << 5) + pp_ctz(bits));
                    bits &= bits - 1;
                }
            }
            if (tid == 0) S.n_sync = (int32_t)total;
            blk.sync();
            const int ns = S.n_sync;
            const int extra = sf < 0 ? 1 : 0;  // the stretch that starts before the region
            for (int k = tid; k < ns + extra; k += nt) {
                const int idx = k - extra;
                int64_t p = idx < 0 ? (int64_t)S.first_sync_global : g0 + S.wlist[idx];
                const int64_t limit = idx + 1 < ns ? g0 + S.wlist[idx + 1] : g0 + PA_R;
                int64_t dend;
                {
                    const int nd = next_ds(idx < 0 ? own_lo : (int)S.wlist[idx]);
                    dend = nd < PA_R ? g0 + nd : (int64_t)S.region_doc_end;
                }
                const int64_t stop = limit < dend ? limit : dend;
                // Inside the region the scanner reads the shared-memory copy of the text (index g -> S.text[g - g0]).
                // It looks at most two characters past the end of a piece (BLOOM: a space and the character after it),
                // so a piece end more than 8 bytes in front of the region end was decided on loaded bytes only; closer
                // than that the scan stops and the word stays open (its end is then found from global memory by the
                // probe loop).
                const int64_t rend = g0 + PA_R;
                const int64_t send = dend < rend ? dend : rend;
                const PaCodeSrc tsm{S.code - g0, S.text - g0};
                bool undecided = false;
                while (p < stop) {
                    int64_t pe;
                    if (p >= g0) {
#if defined(DPT_NO_LETTER_SKIP)  // tuning variant: the scanner walks letter runs character by character
                        pe = dpt_piece_end_src(P.rule, tsm, p, send, DptNoSkip{});
#else
                        pe = dpt_piece_end_src(P.rule, tsm, p, send, PaLetterSkip{S.mAL, g0});
#endif
                        if (send < dend && pe + 8 > rend) {
                            undecided = true;
                            const int64_t r = p - g0;
                            blk.atomic_or(&S.mWS[r >> 5], 1u << (r & 31));
                            break;
                        }
                    } else {
                        pe = dpt_piece_end(P.rule, U, P.text, p, dend);  // the stretch that starts before the region
                    }
                    const int64_t r = p - g0;
                    if (r >= 0) blk.atomic_or(&S.mWS[r >> 5], 1u << (r & 31));
                    p = pe;
                }
                if (!undecided && p >= g0 && p < g0 + PA_R && p < n) {  // the piece start the scan landed on
                    const int64_t r = p - g0;
                    blk.atomic_or(&S.mWS[r >> 5], 1u << (r & 31));
                }
            }
        }
        blk.sync();
        for (int w = w0; w < w1; ++w) {  // (this thread's own mask words: their counts stay in a register)
            const uint32_t ds = S.mDS[w];
            const uint32_t ws = S.mWS[w] | ds;
            S.mWS[w] = ws;
            my_cnt += (uint32_t)pp_popc(ws & own_mask(w)) | ((uint32_t)pp_popc(ds & upto_mask(w)) << 16);
        }
    } else
    {
        for (int w = w0; w < w1; ++w) {
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
            S.mWS[w] = ws;
            S.mCX[w] = cx;
            if (cx) S.any_cx = 1u;
            const uint32_t rm = own_mask(w);
            // words (and '<s>' words) that start in this tile | document starts up to here << 16
            my_cnt += (uint32_t)(pp_popc(ws & rm) + (spm ? pp_popc(ds & rm) : 0)) | ((uint32_t)pp_popc(ds & upto_mask(w)) << 16);
            amb &= rm;
            // (with the tokenizer's merge table kernel B splits these runs itself: pb_segment_split; nothing is ambiguous)
            while (amb && P.doc_flags && !P.V.merge_mask) {
                const int r = (w << 5) + pp_ctz(amb);
                amb &= amb - 1;
                const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, g0 + r) - 1;
                if (d >= P.doc_begin && d < P.doc_begin + P.n_docs_local) P.doc_flags[d - P.doc_begin] = 1;  // DPT_DF_AMBIGUOUS
            }
        }
    }
    // (no barrier here: every thread wrote the word-start masks of its own mask words and kept their counts; the scan's
    // barriers order the masks for the probe phase, which reads its neighbours')

    // ---- entries (words and '<s>' words) that start in this tile, in corpus order ---------------------------------
    uint32_t my_off, my_dord;
    int ne;
    {
        uint32_t total;
        const uint32_t ex = blk.exclusive_scan(my_cnt, S.scan, total);
        my_off = ex & 0xFFFFu;
        my_dord = ex >> 16;  // document starts in the region before this thread's chunk
        ne = (int)(total & 0xFFFFu);
    }
    // publish this tile's word count at once (successors never wait long); resolve the prefix as late as possible:
    // after the probes when the tile fits one window (the normal case)
    blk.lookback_publish(P.desc_w, tile, (unsigned long long)ne);
    const bool single = ne <= PA_WIN;
    if (!single) {
        blk.lookback_resolve(P.desc_w, tile, (unsigned long long)ne, &S.base_w);
        blk.sync();
    }
    int lo = 0;
    do {
        const int hi = (lo + PA_WIN) < ne ? (lo + PA_WIN) : ne;
        const int nwin = hi - lo;
        // ---- list of this window ------------------------------------------------------------------------------
        {
            uint32_t off = my_off, dord = my_dord;
            for (int w = w0; w < w1; ++w) {
                uint32_t bits = S.mWS[w] & own_mask(w);
                const uint32_t dsw = S.mDS[w] & upto_mask(w);
                while (bits) {
                    const int r = (w << 5) + pp_ctz(bits);
                    bits &= bits - 1;
                    if (spm && ((dsw >> (r & 31)) & 1u)) {
                        if ((int)off >= lo && (int)off < hi) {
                            // index of this document = first document of the region + document starts before r
                            const uint32_t before = dord + (uint32_t)pp_popc(dsw & ((1u << (r & 31)) - 1u));
                            S.stage[off - lo] = REF_BOS | ((uint32_t)((int64_t)S.d_first + before - P.doc_begin) & REF_INDEX);
                            S.wlist[off - lo] = (uint16_t)(r | 0x8000);
                        }
                        ++off;
                    }
                    if ((int)off >= lo && (int)off < hi) S.wlist[off - lo] = (uint16_t)r;
                    ++off;
                }
                dord += (uint32_t)pp_popc(dsw);
            }
        }
        blk.sync();

        // ---- one table probe per word ----------------------------------------------------------------------------
            for (int k = tid; k < nwin; k += nt) {
            const uint32_t e = S.wlist[k];
            const int ws = (int)(e & 0x7FFFu);
            uint32_t ref;
            if (e & 0x8000u) {
                continue;  // staged with its document index by the list build
            } else {
                const bool ds = pp_bit(S.mDS, ws);
                // the word ends where the next entry of the list starts (a '<s>' entry sits on its document's first
                // word); the last entry of a window looks the end up in the word-start mask
                const int we = k + 1 < nwin ? (int)(S.wlist[k + 1] & 0x7FFFu) : pp_mask_next(S.mWS, ws + 1, PA_R);
                const int ml = !spm ? 0 : ds ? 0 : (S.text[ws] == 0x20u ? 1 : 3);
                const int b = ws + ml, len = we - b;
                const bool open = we >= PA_R;
                bool odd = open || len > PA_MAXLEN || len < 0 || (S.any_cx && pp_any_in_range(S.mCX, ws, we));
                ref = 0;
                if (!odd) {
                    // hash of the length and the first 16 bytes of the body (four zero-padded words, straight-line: every
                    // lane active).  Longer bodies hash like their 16-byte prefix: the table verifies the bytes anyway, and
                    // bodies of one length that share 16 bytes are too few to lengthen the probe sequences.
                    uint32_t wv[4];
                    pp_load16_raw(S.text, b, wv);
                    const uint4 lm = pp_tail_lut(len);
                    wv[0] &= lm.x; wv[1] &= lm.y; wv[2] &= lm.z; wv[3] &= lm.w;
                    uint32_t h = 0x811C9DC5u ^ (uint32_t)len;
                    h = pp_hash_step(pp_hash_step(pp_hash_step(pp_hash_step(h, wv[0]), wv[1]), wv[2]), wv[3]);
                    h *= 0x2C1B3C6Du;
                    h ^= h >> 13;
                    const int64_t g_b = g0 + b;
                    const unsigned long long mine = pp_tag(h >> 12, len, g_b);
                    uint32_t slot = h & P.slot_mask;
                    bool done = false;
                    for (int probe = 0; probe < PA_PROBES && !done; ++probe, slot = (slot + 1) & P.slot_mask) {
                        unsigned long long t = blk.load_relaxed(&P.tags[slot]);
                        if (t == 0) {
                            t = blk.cas_u64(&P.tags[slot], 0ull, mine);
                            if (t == 0) {  // first occurrence of this word: claim the slot, queue the DP
                                const uint32_t li = blk.atomic_add_ret(&S.n_pend, 1u);
                                const uint32_t cls = (uint32_t)pp_len_class(len + (spm ? 1 : 0));
                                S.pend[li] = slot | (cls << 29);
                                blk.atomic_add(&S.n_pend_c[cls], 1u);
                                ref = slot;
                                done = true;
                                break;
                            }
                        }
                        if ((t >> 38) == (mine >> 38)) {  // same hash bits and length: verify against the corpus text
                            const int64_t rp = pp_tag_pos(t);
                            bool same;
                            if (rp + len + 24 <= n) {  // words straight off the 4-byte-aligned corpus base
                                const uint8_t* base4 = P.text - ((uintptr_t)P.text & 3u);
                                const int64_t ro = rp + (int64_t)((uintptr_t)P.text & 3u);
                                uint32_t rv[4];
                                pp_load16_raw(base4, ro, rv);
                                same = (((rv[0] & lm.x) ^ wv[0]) | ((rv[1] & lm.y) ^ wv[1]) | ((rv[2] & lm.z) ^ wv[2]) |
                                        ((rv[3] & lm.w) ^ wv[3])) == 0;
                                for (int q = 16; q < len && same; q += 4)
                                    same = ((pp_load4(base4, ro + q) ^ pp_load4(S.text, b + q)) & pp_tail_mask(len - q)) == 0;
                            } else {
                                same = true;
                                for (int q = 0; q < len && same; ++q) same = P.text[rp + q] == S.text[b + q];
                            }
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
                    int64_t g_we = g0 + we;
                    if (open) {
                        if (spm) {
                            g_we = pp_spm_word_end_global(P, g_ws, ml, ds);
                        } else {
                            const DptUniView U{P.V.uni1, P.V.uni2};
                            const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, g_ws);
                            g_we = dpt_piece_end(P.rule, U, P.text, g_ws, d <= P.n_docs ? P.doc_offs[d] : n);
                        }
                    }
                    const uint32_t j = blk.atomic_add_ret(&P.ctl->n_odd, 1u);
                    if ((int64_t)j < P.odd_cap) {
                        OddWord o;
                        o.pos = g_ws;
                        o.len = (int32_t)(g_we - g_ws);
                        o.virt = (spm && ds) ? 1 : 0;
                        P.odd[j] = o;
                    }
                    ref = REF_ODD | (j & REF_INDEX);
                }
                if (!spm && ds) ref |= REF_DOCFIRST;
            }
            S.stage[k] = ref;
        }
        if (single) blk.lookback_resolve(P.desc_w, tile, (unsigned long long)ne, &S.base_w);  // warp 0, after its probes
        blk.sync();
        const int64_t base = (int64_t)S.base_w + lo;
        for (int k = tid; k < nwin; k += nt) {
            const uint32_t ref = S.stage[k];
            if (base + k < P.word_cap) P.refs[base + k] = ref;
            if (!spm && (ref & REF_DOCFIRST)) {
                // index of this document = first document of the region + document starts in front of the word
                const int r = (int)(S.wlist[k] & 0x7FFFu);
                const int64_t d = (int64_t)S.d_first + S.dsn[r >> 5] + pp_popc(S.mDS[r >> 5] & ((1u << (r & 31)) - 1u));
                if (d >= P.doc_begin && d < P.doc_begin + P.n_docs_local) P.doc_first_word[d - P.doc_begin] = base + k;
            }
        }
        // ---- distinct words claimed in this window -> the DP queues (by length class) -------------------------------
        for (int c = tid; c < PB_CLASSES; c += nt)
            S.base_c[c] = S.n_pend_c[c] ? blk.atomic_add_ret(&P.ctl->n_pending[c], S.n_pend_c[c]) : 0u;
        blk.sync();
        const uint32_t npend = S.n_pend;
        for (uint32_t i = tid; i < npend; i += nt) {
            const uint32_t v = S.pend[i], cls = v >> 29;
            const uint32_t r = blk.atomic_add_ret(&S.cur_c[cls], 1u);
            // (a queue only overflows when the corpus holds more words than word_cap: reported, the caller retries)
            if ((int64_t)(S.base_c[cls] + r) < P.pend_stride)
                P.pending[(size_t)cls * (size_t)P.pend_stride + S.base_c[cls] + r] = v & REF_INDEX;
        }
        blk.sync();
        if (tid == 0) {
            S.n_pend = 0;
            for (int c = 0; c < PB_CLASSES; ++c) S.n_pend_c[c] = S.cur_c[c] = 0;
        }
        lo += PA_WIN;
        if (lo < ne) blk.sync();
    } while (lo < ne);
    if (tid == 0 && tile == P.n_tiles - 1) P.ctl->n_words = (unsigned long long)((int64_t)S.base_w + ne);
    if (tid == 0 && tile == 0) P.counters[3] = 0;  // untokenizable words: kernel C adds to it tile by tile
    blk.sync();
}

template <class Blk, bool kSpm>
DPT_PIPE_FN void pa_kernel(Blk& blk, const PipeParams& P, ASmemT<kSpm>& S) {
    // The look-back of a tile waits for its predecessors, so a tile must only start once every earlier tile has.
    // On the device the hardware dispatches CTAs in blockIdx order (the assumption CUB's decoupled look-back scan is
    // built on), so tile = blockIdx.x and no CTA begins with a global atomic round trip; a persistent block (the host
    // emulation) takes tiles in corpus order from an atomic ticket.
    if (!blk.persistent()) {
        const int tile = blk.block_index();
        if (tile < P.n_tiles) pa_run_tile<Blk, kSpm>(blk, P, S, tile);
        return;
    }
    for (;;) {
        if (blk.tid() == 0) S.tile = (int32_t)blk.atomic_add_ret(&P.ctl->ticket_a, 1u);
        blk.sync();
        const int tile = S.tile;
        if (tile >= P.n_tiles) break;
        pa_run_tile<Blk, kSpm>(blk, P, S, tile);
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
        const uint32_t c0 = P.text[p];
        if (c0 < 128u && c0 != 0x20u && ((V.ascii_single[c0 >> 5] >> (c0 & 31)) & 1u) &&
            (p + 1 >= e_end || dpt_is_cp_start(P.text[p + 1]))) {  // the common case, one straight line
            if (n + 1 > cap) return -1;
            out[n++] = (uint8_t)c0;
            ++p;
            continue;
        }
        int64_t e = p + 1;
        while (e < e_end && !dpt_is_cp_start(P.text[e])) ++e;
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

// ---- SentencePiece marker runs (tokenizer_utils.py:7-31) ------------------------------------------------
// The reference's words are the DEFAULT tokenizer's tokens glued together up to the next token that starts with U+2581.
// Between two non-marker characters separated by ONE marker that is the marker itself; inside a RUN of markers ("a   b",
// leading spaces, "\n\n  indented") it depends on which markers the default BPE merged with each other ("▁▁") and which
// with the text behind them ("▁b"), i.e. on the merge ranks.  Kernel A keeps such a run and the text up to the next
// single marker as ONE odd word; here it is cut the way the tokenizer cuts it: the characters become the tokenizer's
// initial symbols (a character that is no vocabulary entry: its "<0xHH>" byte tokens), the merge table is applied in rank
// order, leftmost first (`tokenizers` BPE: Word::merge_all), and a new word starts at every resulting token whose first
// symbol is a marker.  No merge crosses the ends of the segment: no vocabulary entry has a marker behind another
// character (marker_leading_only, a precondition of the device rule).
constexpr int PB_SEG_MAX = 512;  // symbols; longer segments (whitespace art) are flagged for the host split
// true when raw bytes [p, e) hold a marker character (' ' or U+2581) behind their first character (virt: anywhere)
DPT_PIPE_FN bool pb_raw_inner_marker(const PipeParams& P, int64_t p, int64_t e, bool virt) {
    int64_t q = p;
    if (!virt) {  // skip the leading marker character of the word itself
        q = p + 1;
        while (q < e && !dpt_is_cp_start(P.text[q])) ++q;
    }
    for (; q < e; ++q) {
        const uint32_t c = P.text[q];
        if (c == 0x20u) return true;
        if (c == DPT_MARK0 && q + 2 < e && P.text[q + 1] == DPT_MARK1 && P.text[q + 2] == DPT_MARK2) return true;
    }
    return false;
}
// symbol = id << 32 | raw offset of its first byte (relative to the segment start) << 1 | its first symbol is a marker.
// Returns the number of word starts written to cut[] (offsets relative to p; cut[0] = 0), or -1: too long for PB_SEG_MAX.
DPT_PIPE_FN int32_t pb_segment_split(const PipeParams& P, int64_t p, int64_t e_end, bool virt, unsigned long long* sym,
                                     int32_t sym_cap, uint32_t* cut) {
    const DptVocabView& V = P.V;
    const int32_t marker_id = V.slot_id[V.marker_slot];
    int32_t n = 0;
    if (virt) sym[n++] = ((unsigned long long)(uint32_t)marker_id << 32) | 1ull;  // Prepend(U+2581): no raw byte of its own
    for (int64_t q = p; q < e_end;) {
        int64_t e = q + 1;
        while (e < e_end && !dpt_is_cp_start(P.text[e])) ++e;
        const uint32_t c0 = P.text[q];
        const bool mk = (c0 == 0x20u) || (e - q == 3 && c0 == DPT_MARK0 && P.text[q + 1] == DPT_MARK1 && P.text[q + 2] == DPT_MARK2);
        const unsigned long long off = (unsigned long long)(q - p) << 1;
        if (mk) {
            if (n + 1 > sym_cap) return -1;
            sym[n++] = ((unsigned long long)(uint32_t)marker_id << 32) | off | 1ull;
        } else {
            uint32_t entry = DPT_DA_ROOT_ENTRY, slot = 0;
            bool hit = true;
            for (int64_t k = q; k < e && hit; ++k) hit = dpt_da_step_idx(V.da, entry, P.text[k], slot);
            if (hit && (entry & DPT_DA_TERMINAL)) {
                if (n + 1 > sym_cap) return -1;
                sym[n++] = ((unsigned long long)(uint32_t)V.slot_id[slot] << 32) | off;
            } else {  // byte fallback: one symbol per byte (they take no part in merges of their own)
                if (n + (int32_t)(e - q) > sym_cap) return -1;
                for (int64_t k = q; k < e; ++k) sym[n++] = ((unsigned long long)(uint32_t)V.byte_ids[P.text[k]] << 32) | off;
            }
        }
        q = e;
    }
    // BPE: the lowest-ranked adjacent pair, leftmost on ties, until none merges
    for (;;) {
        uint32_t best = 0xFFFFFFFFu;
        int32_t bi = -1, bid = 0;
        for (int32_t i = 0; i + 1 < n; ++i) {
            uint32_t rank;
            int32_t merged;
            if (dpt_merge_lookup(V, (int32_t)(uint32_t)(sym[i] >> 32), (int32_t)(uint32_t)(sym[i + 1] >> 32), rank, merged) &&
                rank < best) {
                best = rank;
                bi = i;
                bid = merged;
            }
        }
        if (bi < 0) break;
        sym[bi] = ((unsigned long long)(uint32_t)bid << 32) | (sym[bi] & 0xFFFFFFFFull);
        for (int32_t i = bi + 1; i + 1 < n; ++i) sym[i] = sym[i + 1];
        --n;
    }
    int32_t nc = 0;
    for (int32_t i = 0; i < n; ++i)
        if (i == 0 || (sym[i] & 1ull)) cut[nc++] = (uint32_t)((sym[i] & 0xFFFFFFFFull) >> 1);
    return nc;
}

// entries of length class c that kernel A could store (see the enqueue at the end of pa_run_tile)
DPT_PIPE_FN uint32_t pb_queue_len(const PipeParams& P, int c) {
    const uint32_t n = P.ctl->n_pending[c];
    return (int64_t)n < P.pend_stride ? n : (uint32_t)P.pend_stride;
}
// item i of the DP work list: i < n_pending -> table slot pending[i]; else odd word i - n_pending
struct PbItem {
    int64_t pos, end;
    bool marker;
    ResRec* out;
};
// Work list of the thread-per-word kernel.  On the device the lock-step kernel (dpt_dp_lock.cuh) solves the length
// classes 0..3; this kernel runs twice: P.coop == 1, BESIDE the lock-step kernel on a side stream: the odd words and the
// words of more than 63 units (class 4); P.coop == 2, after it: the words the lock-step kernels deferred
// (out-of-vocabulary characters that expand to "<0xHH>" text).  P.coop == 0 (host emulation): odd words, then every
// length class from longest to shortest.
DPT_PIPE_FN uint64_t pb_list_len(const PipeParams& P, const uint32_t* npc, uint32_t n_odd, uint32_t n_defer) {
    if (P.coop == 2) return n_defer;
    uint64_t total = n_odd;
    if (P.coop == 1) return total + npc[PB_CLASSES - 1];
    for (int c = 0; c < PB_CLASSES; ++c) total += npc[c];
    return total;
}
// a word of the long-word queue: bit 31 = odd word index, else table slot
DPT_PIPE_FN PbItem pb_item_tagged(const PipeParams& P, uint32_t v) {
    PbItem it;
    if (v & 0x80000000u) {
        const uint32_t i = v & 0x7FFFFFFFu;
        const OddWord o = P.odd[i];
        it.pos = o.pos;
        it.end = o.pos + o.len;
        it.marker = o.virt != 0;
        it.out = &P.odd_res[i];
        return it;
    }
    const unsigned long long t = P.tags[v];
    it.pos = pp_tag_pos(t);
    it.end = it.pos + pp_tag_len(t);
    it.marker = P.spm != 0;
    it.out = &P.res[v];
    return it;
}
// item i of the work list; *tagged = its long-word-queue form
DPT_PIPE_FN PbItem pb_item(const PipeParams& P, uint64_t i, const uint32_t* npc, uint32_t n_odd, uint32_t* tagged) {
    PbItem it;
    if (P.coop != 2 && i < n_odd) {
        *tagged = 0x80000000u | (uint32_t)i;
        return pb_item_tagged(P, *tagged);
    }
    if (P.coop != 2) i -= n_odd;
    uint32_t slot;
    if (P.coop == 2) {
        slot = P.defer[i];
    } else if (P.coop == 1) {
        slot = P.pending[(size_t)(PB_CLASSES - 1) * (size_t)P.pend_stride + i];
    } else {
        uint32_t cls = PB_CLASSES - 1;
        while (cls > 0 && i >= npc[cls]) {
            i -= npc[cls];
            --cls;
        }
        slot = P.pending[(size_t)cls * (size_t)P.pend_stride + i];
    }
    *tagged = slot;
    return pb_item_tagged(P, slot);
}

// One thread per distinct word, DP state in local memory; the forward pass is the resumable state machine of
// dpt_dp_core.h.  The lanes of a warp step their words in lock step; words differ a lot in their step counts, so a
// warp that simply waited for its longest word ran the hot loop with 11 of 32 lanes (ncu, v11).  Here a lane that has
// finished its word leaves the hot loop only when fewer than PB_REFILL lanes are still walking (or every 4 steps when
// nothing is left to take): the finished lanes then write their records and take new words from the work list
// (normalise, initialise) while the unfinished lanes keep their walk state in registers, and all re-enter the loop.
template <class Blk>
DPT_PIPE_FN void pb_finish_word(Blk& blk, const PipeParams& P, const uint8_t* norm, int32_t n, const uint32_t* best,
                                const uint32_t* Ap, const uint32_t* Bp, const uint8_t* upos, ResRec* out) {
    ResRec rec;
    for (int k = 0; k < RES_INLINE; ++k) rec.ids[k] = 0;
    const uint32_t kn = best[n];
    const uint32_t word_len = dpt_k32_len(kn);
    const bool reach = dpt_k32_reach(kn);
    rec.meta = (word_len & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK);
    if (reach) {
        if (word_len <= (uint32_t)RES_INLINE) {
            dpt_backward_chase(P.V, norm, n, word_len, dpt_k32_longest(kn), Ap, Bp, rec.ids, RES_INLINE, upos);
        } else {
            const unsigned long long off = blk.atomic_add_u64_ret(&P.persist->pool_used, (unsigned long long)word_len);
            rec.meta |= RES_POOLED;
            rec.ids[0] = (int32_t)(uint32_t)(off & 0xFFFFFFFFull);
            rec.ids[1] = (int32_t)(uint32_t)(off >> 32);
            if ((int64_t)(off + word_len) <= P.pool_cap)
                dpt_backward_chase(P.V, norm, n, word_len, dpt_k32_longest(kn), Ap, Bp, P.pool + off, (int64_t)word_len, upos);
        }
    }
    *out = rec;
}

template <class Blk>
DPT_PIPE_FN void pb_thread(Blk& blk, const PipeParams& P) {
    uint32_t npc[PB_CLASSES];
    for (int c = 0; c < PB_CLASSES; ++c) npc[c] = pb_queue_len(P, c);
    const uint32_t n_odd = P.ctl->n_odd < (uint32_t)P.odd_cap ? P.ctl->n_odd : (uint32_t)P.odd_cap;
    const uint64_t total = pb_list_len(P, npc, n_odd, P.ctl->n_defer);
    // Few words (behind the cooperative kernel this list holds the odd words and the words longer than a warp: a few
    // thousand): one word per WARP instead of one per lane.  32 unrelated long words in lock step serialise their
    // divergent walks (measured 0.14 ms for 2.7 k words); alone in its warp a word runs at the speed of its own chain.
    const bool spread = total * 2 <= (uint64_t)blk.grid_warps();
    uint8_t norm[PB_LOCAL + 8];
    uint32_t best[PB_LOCAL + 1], Ap[PB_LOCAL + 1], Bp[PB_LOCAL + 1];
    uint8_t upos[PB_LOCAL + 8];
    DptFlat32 st;
    st.j = st.i = 0;
    st.entry = st.cl = st.kj = 0;
    ResRec* out = nullptr;
    int32_t n = 0;      // 0: no word (dpt_flat32_running is false)
    bool have = false;  // this lane holds a word whose forward pass is not complete
    bool more = true;   // warp-uniform: the work list may still hold items
    for (;;) {
        if (more) {
            // lanes without a word take the next items of the list (one atomic per warp)
            const bool ask = !have && (!spread || blk.lane() == 0);
            const uint64_t idx = blk.warp_take_n(P.coop == 2 ? &P.ctl->b_cursor2 : &P.ctl->b_cursor, ask);
            const bool got = ask && idx < total;
            more = !blk.warp_any(ask && !got);
            bool fresh = false;
            if (got) {
                uint32_t tagged;
                const PbItem it = pb_item(P, idx, npc, n_odd, &tagged);
                int32_t nlen = pb_normalise(P, it.pos, it.end, it.marker, norm, PB_LOCAL);
                // an odd word with a marker run inside: the long-word kernel cuts it into the tokenizer's words (pb_segment_split)
                if (nlen > 0 && P.spm && P.V.merge_mask && (tagged & 0x80000000u) && pb_raw_inner_marker(P, it.pos, it.end, it.marker))
                    nlen = -1;
                if (nlen < 0) {  // too long for the local state: the long-word kernel solves it
                    const uint32_t q = blk.atomic_add_ret(&P.ctl->n_long, 1u);
                    P.longq[q] = tagged;
                } else if (nlen == 0) {  // cannot happen (documents are non-empty); keep the record defined
                    ResRec rec;
                    for (int k = 0; k < RES_INLINE; ++k) rec.ids[k] = 0;
                    rec.meta = RES_UNTOK;
                    *it.out = rec;
                } else {
                    out = it.out;
                    n = nlen;
                    fresh = true;
                }
            }
            // the lanes leave the normalisation at different times: bring them together again before the
            // initialisation loop (ncu, v12: it ran with 4 of 32 lanes inside the divergent region)
            blk.reconverge();
            if (fresh) {
                dpt_flat32_init(P.V, norm, n, best, Ap, Bp, st, upos);
                have = true;
            }
            blk.reconverge();
        }
        if (!blk.warp_any(have)) {
            if (!more) break;
            continue;
        }
        // the hot loop: one trie step per lane and iteration
        const int thresh = more ? PB_REFILL : 1;
        do {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (dpt_flat32_running(st, n)) dpt_flat32_step(P.V, norm, n, best, Ap, Bp, st);
        } while (blk.warp_count(dpt_flat32_running(st, n)) >= thresh);
        blk.reconverge();
        if (have && !dpt_flat32_running(st, n)) {
            pb_finish_word(blk, P, norm, n, best, Ap, Bp, upos, out);
            have = false;
            n = 0;
            st.j = 0;
        }
        blk.reconverge();
    }
}

// long words: one thread each, state in a global scratch pool (13 bytes per normalised position)
template <class Blk>
DPT_PIPE_FN void pb_long_thread(Blk& blk, const PipeParams& P, int64_t gtid, int64_t gthreads) {
    const uint32_t n_long = P.ctl->n_long;
    for (uint64_t k = (uint64_t)gtid; k < n_long; k += (uint64_t)gthreads) {
        const PbItem it = pb_item_tagged(P, P.longq[k]);
        const int64_t raw = it.end - it.pos;
        const int64_t need_dp = (P.spm ? 6 * raw + 3 : raw) + 2;
        // a marker-run segment (see pb_segment_split) also needs its symbols and its cuts: 1.5 x (raw + 2) positions of lp_best
        const bool seg = P.spm && P.V.merge_mask && raw + 1 <= PB_SEG_MAX && pb_raw_inner_marker(P, it.pos, it.end, it.marker);
        const int64_t need = need_dp + (seg ? raw + 2 + (raw + 2) / 2 + 1 : 0);
        const unsigned long long off = blk.atomic_add_u64_ret(&P.ctl->lp_used, (unsigned long long)need);
        ResRec rec;
        for (int q = 0; q < RES_INLINE; ++q) rec.ids[q] = 0;
        if ((int64_t)(off + need) > P.lp_cap || need >= (1ll << 31)) {  // reported through n_out; caller retries bigger
            rec.meta = RES_UNTOK | RES_LONG;
            *it.out = rec;
            blk.atomic_add_u64_ret(&P.ctl->n_too_long, 1ull);
            continue;
        }
        uint8_t* norm = P.lp_norm + off;
        uint64_t* best = P.lp_best + off;
        uint16_t* A = P.lp_a + off;
        uint16_t* B = P.lp_b + off;
        if (P.spm && P.V.merge_mask && !seg && pb_raw_inner_marker(P, it.pos, it.end, it.marker) && P.doc_flags) {
            // a marker run too long to split here: the document goes back to the host split (DPT_DF_AMBIGUOUS)
            const int64_t d = pp_upper_bound(P.doc_offs, P.n_docs + 1, it.pos) - 1;
            if (d >= P.doc_begin && d < P.doc_begin + P.n_docs_local) P.doc_flags[d - P.doc_begin] = 1;
        }
        if (seg) {
            unsigned long long* sym = reinterpret_cast<unsigned long long*>(best + need_dp);
            uint32_t* cut = reinterpret_cast<uint32_t*>(sym + raw + 2);
            const int32_t nc = pb_segment_split(P, it.pos, it.end, it.marker, sym, (int32_t)(raw + 2), cut);
            // two passes over the words of the segment: token counts, then (one pool allocation) the ids
            uint32_t total = 0;
            bool reach_all = nc > 0;
            unsigned long long po = 0;
            for (int pass = 0; pass < 2 && reach_all; ++pass) {
                uint32_t at = 0;
                for (int32_t w = 0; w < nc; ++w) {
                    const int64_t a = it.pos + cut[w], b = w + 1 < nc ? it.pos + cut[w + 1] : it.end;
                    const int32_t nlen = pb_normalise(P, a, b, it.marker && w == 0, norm, (int32_t)(need_dp - 2));
                    dpt_forward<true>(P.V, norm, nlen, nullptr, best, A, B);
                    const uint64_t kn = best[nlen];
                    const uint32_t wl = dpt_key_len(kn);
                    if (!dpt_key_reach(kn)) {
                        reach_all = false;
                        total += wl;
                        continue;
                    }
                    if (pass == 0) {
                        total += wl;
                    } else {
                        if ((int64_t)(po + at + wl) <= P.pool_cap) dpt_backward_emit(P.V, norm, nlen, best, A, B, P.pool + po + at, (int64_t)wl);
                        at += wl;
                    }
                }
                if (pass == 0 && reach_all) po = blk.atomic_add_u64_ret(&P.persist->pool_used, (unsigned long long)total);
            }
            rec.meta = (total & 0xFFFFFFu) | (reach_all ? 0u : RES_UNTOK) | RES_LONG;
            if (reach_all) {
                rec.meta |= RES_POOLED;
                rec.ids[0] = (int32_t)(uint32_t)(po & 0xFFFFFFFFull);
                rec.ids[1] = (int32_t)(uint32_t)(po >> 32);
            }
            *it.out = rec;
            continue;
        }
        const int32_t nlen = pb_normalise(P, it.pos, it.end, it.marker, norm, (int32_t)(need - 2));
        dpt_forward<true>(P.V, norm, nlen, nullptr, best, A, B);
        const uint64_t kn = best[nlen];
        const uint32_t word_len = dpt_key_len(kn);
        const bool reach = dpt_key_reach(kn);
        rec.meta = (word_len & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | RES_LONG;
        if (reach) {
            const unsigned long long po = blk.atomic_add_u64_ret(&P.persist->pool_used, (unsigned long long)word_len);
            rec.meta |= RES_POOLED;
            rec.ids[0] = (int32_t)(uint32_t)(po & 0xFFFFFFFFull);
            rec.ids[1] = (int32_t)(uint32_t)(po >> 32);
            if ((int64_t)(po + word_len) <= P.pool_cap) dpt_backward_emit(P.V, norm, nlen, best, A, B, P.pool + po, (int64_t)word_len);
        }
        *it.out = rec;
    }
}

// =========================================================================================================
// Kernel C: scan + emit
// =========================================================================================================
// streaming (evict-first) accesses for data touched once, so the result records keep their L2 lines
DPT_PIPE_FN uint4 pc_ld_stream(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldcs(p);
#else
    return *p;
#endif
}
DPT_PIPE_FN void pc_st_stream(uint4* p, uint4 v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}
DPT_PIPE_FN void pc_st_stream(int32_t* p, int32_t v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}

DPT_PIPE_FN uint4 pc_ld_head(const ResRec* r) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4*>(r));
#else
    return *reinterpret_cast<const uint4*>(r);
#endif
}
DPT_PIPE_FN uint4 pc_ld_tail(const ResRec* r) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4*>(r) + 1);
#else
    return reinterpret_cast<const uint4*>(r)[1];
#endif
}
DPT_PIPE_FN const ResRec* pc_record_ptr(const PipeParams& P, uint32_t ref) {
    if ((ref & REF_KIND) == REF_ODD) {
        const uint32_t j = ref & REF_INDEX;
        return (int64_t)j < P.odd_cap ? &P.odd_res[j] : nullptr;
    }
    return &P.res[ref & REF_INDEX];
}
DPT_PIPE_FN uint32_t pc_meta(const PipeParams& P, uint32_t ref) {
    if ((ref & REF_KIND) == REF_BOS) return (uint32_t)P.V.bos_len | (P.V.bos_ntok ? 0u : RES_UNTOK);
    const ResRec* r = pc_record_ptr(P, ref);
    if (!r) return RES_UNTOK;
#if defined(__CUDA_ARCH__)
    return __ldg(&r->meta);
#else
    return r->meta;
#endif
}

// Final counters and the capacity report, written by thread 0 of kernel C's LAST tile, which knows the grand total of
// tokens from its own look-back.  Everything else here was final before kernel C started; counters[3] (untokenizable
// words) is zeroed by kernel A and added to by every tile of kernel C.
DPT_PIPE_FN void pd_finish(const PipeParams& P, unsigned long long tot) {
    const int64_t n_words_true = (int64_t)P.ctl->n_words;
    P.counters[0] = (unsigned long long)(P.byte_end - P.byte_begin);
    P.counters[1] = (unsigned long long)n_words_true;
    P.counters[2] = tot;
    P.n_out[0] = (int64_t)tot;             // DPT_NOUT_IDS
    P.n_out[1] = n_words_true;             // DPT_NOUT_WORDS
    P.n_out[2] = (int64_t)P.ctl->lp_used;  // DPT_NOUT_POOL_REQ  (long-word scratch positions)
    P.n_out[3] = P.lp_cap;                 // DPT_NOUT_POOL_CAP
    P.n_out[4] = (int64_t)P.persist->pool_used;  // ids pool required
    P.n_out[5] = P.pool_cap;
    P.n_out[6] = (int64_t)P.ctl->n_odd;    // odd words required
    P.n_out[7] = P.odd_cap;
    P.doc_tok_offs[P.n_docs_local] = (int64_t)tot;
}

template <class Blk>
DPT_PIPE_FN void pc_run_tile(Blk& blk, const PipeParams& P, CSmem& S, const int tile) {
    const int tid = blk.tid();
    const int64_t n_words = (int64_t)P.ctl->n_words < P.word_cap ? (int64_t)P.ctl->n_words : P.word_cap;
    const int64_t w0 = (int64_t)tile * PC_TILE + (int64_t)tid * PC_PER;
    uint32_t ref[PC_PER], meta[PC_PER];
    uint4 head[PC_PER];  // every word's result record: meta + ids[0..2] (pooled: meta + pool offset)
    uint32_t mine = 0, untok = 0;
    if (w0 + PC_PER <= n_words) {  // two 16-byte loads of 8 refs
        const uint4 r0 = pc_ld_stream(reinterpret_cast<const uint4*>(P.refs + w0));
        const uint4 r1 = pc_ld_stream(reinterpret_cast<const uint4*>(P.refs + w0 + 4));
        ref[0] = r0.x; ref[1] = r0.y; ref[2] = r0.z; ref[3] = r0.w;
        ref[4] = r1.x; ref[5] = r1.y; ref[6] = r1.z; ref[7] = r1.w;
    } else {
#pragma unroll
        for (int k = 0; k < PC_PER; ++k) ref[k] = w0 + k < n_words ? P.refs[w0 + k] : REF_BOS;
    }
    // one 16-byte load per word (eight independent ones in flight per thread): words of up to three tokens - nine of ten -
    // need nothing else from their record, so the id copy below does not wait for memory a second time
#pragma unroll
    for (int k = 0; k < PC_PER; ++k) {
        uint4 h;
        h.x = RES_UNTOK;
        h.y = h.z = h.w = 0u;
        if (w0 + k < n_words) {
            if ((ref[k] & REF_KIND) == REF_BOS) {
                h.x = (uint32_t)P.V.bos_len | (P.V.bos_ntok ? 0u : RES_UNTOK);
                h.y = (uint32_t)P.V.bos_ids[0];
                h.z = (uint32_t)P.V.bos_ids[1];
                h.w = (uint32_t)P.V.bos_ids[2];
            } else {
                const ResRec* r = pc_record_ptr(P, ref[k]);
                if (r) h = pc_ld_head(r);
            }
        }
        head[k] = h;
    }
#pragma unroll
    for (int k = 0; k < PC_PER; ++k) {
        meta[k] = head[k].x;
        if (w0 + k < n_words) {
            if (meta[k] & RES_UNTOK) ++untok; else mine += meta[k] & 0xFFFFFFu;
        }
    }
    // byte-level rules: first words of documents among this thread's words (a '<s>' ref carries its document index)
    uint32_t my_docfirst = 0;
    if (!P.spm) {
#pragma unroll
        for (int k = 0; k < PC_PER; ++k)
            if (w0 + k < n_words && (ref[k] & REF_KIND) != REF_BOS && (ref[k] & REF_DOCFIRST)) ++my_docfirst;
    }
    if (tid == 0) S.n_untok = S.n_docfirst = 0;
    uint32_t total;
    const uint32_t off = blk.exclusive_scan(mine, S.scan, total);
    blk.lookback_publish(P.desc_t, tile, (unsigned long long)total);
    if (untok) blk.atomic_add(&S.n_untok, untok);
    if (my_docfirst) blk.atomic_add(&S.n_docfirst, my_docfirst);
    // per-word outputs that do not need the token offset
    if (w0 + PC_PER <= n_words && P.vec_ok) {
        uint4 l0, l1;
        l0.x = meta[0] & 0xFFFFFFu; l0.y = meta[1] & 0xFFFFFFu; l0.z = meta[2] & 0xFFFFFFu; l0.w = meta[3] & 0xFFFFFFu;
        l1.x = meta[4] & 0xFFFFFFu; l1.y = meta[5] & 0xFFFFFFu; l1.z = meta[6] & 0xFFFFFFu; l1.w = meta[7] & 0xFFFFFFu;
        pc_st_stream(reinterpret_cast<uint4*>(P.word_lens + w0), l0);
        pc_st_stream(reinterpret_cast<uint4*>(P.word_lens + w0 + 4), l1);
        unsigned long long f = 0;
#pragma unroll
        for (int k = 0; k < PC_PER; ++k)
            f |= (unsigned long long)(((meta[k] & RES_UNTOK) ? 1u : 0u) | ((meta[k] & RES_LONG) ? 4u : 0u)) << (8 * k);
        *reinterpret_cast<unsigned long long*>(P.word_flags + w0) = f;
    } else {
#pragma unroll
        for (int k = 0; k < PC_PER; ++k)
            if (w0 + k < n_words) {
                P.word_lens[w0 + k] = (int32_t)(meta[k] & 0xFFFFFFu);
                P.word_flags[w0 + k] = (uint8_t)(((meta[k] & RES_UNTOK) ? 1u : 0u) | ((meta[k] & RES_LONG) ? 4u : 0u));
            }
    }
    // ids: staged in shared memory at their tile-local offsets (every thread copies its words' ids out of the
    // records), then written to their final place by the whole CTA with fully coalesced stores.  The look-back is
    // resolved in between, as late as possible.
    const bool staged = total <= (uint32_t)PC_STAGE;
    if (!staged) {
        blk.lookback_resolve(P.desc_t, tile, (unsigned long long)total, &S.base_t);
        blk.sync();
    }
    {
        int64_t gt = staged ? (int64_t)off : (int64_t)S.base_t + off;  // tile-local when staged
        int32_t* dst = staged ? S.ids : P.ids;
        const int64_t cap = staged ? (int64_t)PC_STAGE : P.ids_cap;
#pragma unroll
        for (int k = 0; k < PC_PER; ++k) {
            if (w0 + k >= n_words) break;
            if (meta[k] & RES_UNTOK) continue;
            const uint32_t nk = meta[k] & 0xFFFFFFu;
            if (nk == 0) continue;
            if (meta[k] & RES_POOLED) {
                const int64_t po = (int64_t)head[k].y | ((int64_t)head[k].z << 32);
                for (uint32_t q = 0; q < nk; ++q)
                    if (gt + q < cap && po + q < P.pool_cap) dst[gt + q] = P.pool[po + q];
            } else {
                if (gt < cap) dst[gt] = (int32_t)head[k].y;
                if (nk > 1 && gt + 1 < cap) dst[gt + 1] = (int32_t)head[k].z;
                if (nk > 2 && gt + 2 < cap) dst[gt + 2] = (int32_t)head[k].w;
                if (nk > (uint32_t)RES_HEAD) {  // ids[3..6]: the second half of the record's sector
                    const ResRec* r = pc_record_ptr(P, ref[k]);
                    const uint4 t = r ? pc_ld_tail(r) : uint4{0u, 0u, 0u, 0u};
                    if (gt + 3 < cap) dst[gt + 3] = (int32_t)t.x;
                    if (nk > 4 && gt + 4 < cap) dst[gt + 4] = (int32_t)t.y;
                    if (nk > 5 && gt + 5 < cap) dst[gt + 5] = (int32_t)t.z;
                    if (nk > 6 && gt + 6 < cap) dst[gt + 6] = (int32_t)t.w;
                }
            }
            gt += nk;
        }
    }
    if (staged) blk.lookback_resolve(P.desc_t, tile, (unsigned long long)total, &S.base_t);
    blk.sync();
    {
        const int64_t base_t = (int64_t)S.base_t;
        if (staged)
            for (uint32_t q = (uint32_t)tid; q < total; q += (uint32_t)blk.nthreads())
                if (base_t + q < P.ids_cap) pc_st_stream(P.ids + base_t + q, S.ids[q]);
        // document token offsets.  Byte-level rules: documents are numbered by their first words, so the index of a
        // first word = documents that start in front of this tile (ONE search of doc_first_word per tile; a search
        // per document made kernel C twice as slow on short documents) + first words in front of it within the tile.
        int64_t d_next = 0;
        if (!P.spm && S.n_docfirst) {  // block-uniform
            uint32_t tile_docs;
            const uint32_t before = blk.exclusive_scan(my_docfirst, S.scan, tile_docs);
            if (tid == 0) S.doc0 = (long long)pp_lower_bound(P.doc_first_word, P.n_docs_local, (int64_t)tile * PC_TILE);
            blk.sync();
            d_next = (int64_t)S.doc0 + before;
        }
        int64_t gt = base_t + off;
#pragma unroll
        for (int k = 0; k < PC_PER; ++k) {
            if (w0 + k >= n_words) break;
            const uint32_t kind = ref[k] & REF_KIND;
            if (kind == REF_BOS) {
                const int64_t d = (int64_t)(ref[k] & REF_INDEX);
                if (d < P.n_docs_local) P.doc_tok_offs[d] = gt;
            } else if (ref[k] & REF_DOCFIRST) {  // byte-level rules: first word of a document
                const int64_t d = d_next++;
                if (d < P.n_docs_local) P.doc_tok_offs[d] = gt;
            }
            if (!(meta[k] & RES_UNTOK)) gt += meta[k] & 0xFFFFFFu;
        }
    }
    if (tid == 0) {
        if (S.n_untok) blk.atomic_add_u64_ret(&P.counters[3], (unsigned long long)S.n_untok);
        // the last tile knows the grand total from its own look-back: it writes the counters and the capacity report
        // (everything else in them was final before this kernel started)
        if ((int64_t)(tile + 1) * PC_TILE >= n_words) pd_finish(P, (unsigned long long)S.base_t + total);
    }
    blk.sync();
}

template <class Blk>
DPT_PIPE_FN void pc_kernel(Blk& blk, const PipeParams& P, CSmem& S) {
    if (!blk.persistent()) {  // tile = blockIdx.x, see pa_kernel
        const int tile = blk.block_index();
        const int64_t n_words = (int64_t)P.ctl->n_words < P.word_cap ? (int64_t)P.ctl->n_words : P.word_cap;
        if (tile < P.n_ctiles && (int64_t)tile * PC_TILE < n_words) pc_run_tile(blk, P, S, tile);
        else if (tile == 0 && blk.tid() == 0) pd_finish(P, 0ull);  // no words (cannot happen: documents are non-empty)
        return;
    }
    for (;;) {
        if (blk.tid() == 0) S.tile = (int32_t)blk.atomic_add_ret(&P.ctl->ticket_c, 1u);
        blk.sync();
        const int tile = S.tile;
        const int64_t n_words = (int64_t)P.ctl->n_words < P.word_cap ? (int64_t)P.ctl->n_words : P.word_cap;
        if (tile >= P.n_ctiles || (int64_t)tile * PC_TILE >= n_words) {
            if (tile == 0 && blk.tid() == 0) pd_finish(P, 0ull);
            break;
        }
        pc_run_tile(blk, P, S, tile);
        if (!blk.persistent()) break;
    }
}

}  // namespace dpt
