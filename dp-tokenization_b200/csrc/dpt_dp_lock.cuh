// Kernel B of the corpus pipeline: the shortest-tokenization DP (dp_tokenize.py:24-84 in the closed form of
// dpt_dp_core.h) for every distinct word of at most 31 units - one THREAD per word, the 32 words of a warp in LOCK STEP.
//
// Why this shape (measured, profiles/README.md):
//   * round 1's kernel stepped a per-lane state machine (one trie step of the lane's own (start, end) pair per
//     iteration) with its DP state in 1 KB of local memory per thread: ~110 instructions per trie step, 417 MB of
//     DRAM traffic for the spilled state.
//   * a sub-warp-cooperative kernel (lane j walks from start j, redux.min relaxation, ballot/clz selection; kept as
//     profiles/rejected/r2_dp_coop.patch) has no local memory but runs dense n x T lane-slots per word:
//     512 M warp-instructions, 0.73 ms against 0.28 ms.
//   * here the loops are warp-uniform - every lane is at start position s of its own word, the walk loop runs until
//     no lane's walk is alive - so the lanes never diverge, the word's bytes sit in a REGISTER window that shifts by one
//     byte per start (static byte extraction, no indexed loads), and the state of a position is ONE 32-bit word
//     (16-bit ordered key | A distance | B distance) in shared memory laid out [position][thread]: conflict-free,
//     16 KB per CTA, nothing spills.
//   * a CTA takes chunks of PBL_CHUNK words from a ticket and counting-sorts each chunk by word length in shared
//     memory before its warps solve it: the outer loop of a warp runs to the LONGEST of its 32 words.
//
// Ordered 16-bit key of a position (words of at most 31 units):  len << 7 | notreach << 6 | (63 - longest token);
// relaxing an edge is one unsigned min, exactly like the 32/64-bit keys of dpt_dp_core.h.  Key 0 = "inside a
// character" (no real key is 0: the origin is 63): such a row never wins a comparison, so the relaxation needs no test
// for it.  A = distance to the largest predecessor with minimal (len, notreach), B = distance to the largest
// predecessor with the minimal key: the backward pass follows B until a token of the target length has been taken, then A
// (dp_tokenize.py:57-69,82-84), and turns every chosen edge into an id by re-walking its bytes (trie lines are L1-hot).
//
// SPM rule: the word-initial U+2581 is one unit in front of the body (start 0 begins at the marker's trie node); a word
// with a character that is no vocabulary entry (its normalised text is the "<0xHH>" spelling, tokenizer_utils.py:26-29)
// is handed to the thread-per-word kernel through the `defer` list, as are words of more than 31 units (class 3) and the
// odd words.
#pragma once
#include "dpt_pipe.h"

namespace dpt {

constexpr int PBL_THREADS = 128;
// Two instantiations: NW = 8 window registers (words of at most 31 units: length classes 0..2, nearly all words) and
// NW = 16 (32..63 units: class 3, a few thousand words, on the side stream).  Rows = positions 0 .. 4 NW - 1.
template <int NW>
struct PblCfg {
    static constexpr int ROWS = 4 * NW;            // DP positions 0 .. ROWS - 1: words of at most ROWS - 1 units
    static constexpr int CHUNK = NW == 8 ? 256 : 128;  // words a CTA sorts and solves at a time (a multiple of PBL_THREADS)
};

template <int NW>
struct PblSmem {
    uint32_t st[PblCfg<NW>::ROWS * PBL_THREADS];    // DP state, [position][thread]
    unsigned long long tag[PblCfg<NW>::CHUNK];      // the chunk's words sorted by length: table tag (position, length) ...
    uint32_t slot[PblCfg<NW>::CHUNK];               // ... and table slot
    uint32_t hist[PblCfg<NW>::ROWS + 1];            // words per length, then the bins' start offsets
    uint32_t chunk;
};

template <int NW>
__device__ __forceinline__ uint32_t pbl_byte(const uint32_t (&wb)[NW], const int k) {  // byte k of the window (k static)
    return (wb[k >> 2] >> (8 * (k & 3))) & 0xFFu;
}
// e = pred ? *p : 0 as ONE predicated load (the compiler turns the C form into a divergent branch: BSSY / BRA / BSYNC)
__device__ __forceinline__ uint32_t pbl_ldg_if(const uint32_t* p, bool pred) {
    uint32_t e;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
        : "=r"(e)
        : "l"(p), "r"((uint32_t)pred));
    return e;
}

// One batch: 32 words (one per lane) of the sorted chunk.
template <bool kSpm, int NW>
__device__ __forceinline__ void pbl_batch(const PipeParams& P, PblSmem<NW>& S, const int first, const int count, const int lane) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int m = kSpm ? 1 : 0;
    constexpr int PBL_ROWS = PblCfg<NW>::ROWS;
    bool valid = first + lane < count;
    uint32_t slot_item = 0;
    int64_t pos = 0;
    int len = 0;
    if (valid) {
        slot_item = S.slot[first + lane];
        const unsigned long long tag = S.tag[first + lane];
        pos = pp_tag_pos(tag);
        len = pp_tag_len(tag);
        if (len + m > PBL_ROWS - 1) {  // cannot happen (the class says so); never index out of the rows
            valid = false;
            len = 0;
        }
    }
    const int n = valid ? len + m : 0;  // units of the word = DP positions 0..n

    // ---- the body bytes, zero-padded to 32, in eight registers --------------------------------------------------
    uint32_t wb[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) wb[q] = 0;
    if (valid) {
        if (pos + 4 * NW + 16 <= P.n_bytes) {
            const uint8_t* base4 = P.text - ((uintptr_t)P.text & 3u);
            const int64_t ro = pos + (int64_t)((uintptr_t)P.text & 3u);
#pragma unroll
            for (int g = 0; g < NW / 4; ++g) {
                if (g == 0 || len > 16 * g) {
                    uint32_t v[4];
                    pp_load16_raw(base4, ro + 16 * g, v);
                    wb[4 * g] = v[0]; wb[4 * g + 1] = v[1]; wb[4 * g + 2] = v[2]; wb[4 * g + 3] = v[3];
                }
            }
        } else {
#pragma unroll
            for (int qw = 0; qw < NW; ++qw) {  // (static indices: the window stays in registers)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (4 * qw + b < len) wb[qw] |= (uint32_t)P.text[pos + 4 * qw + b] << (8 * b);
            }
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {  // zero the bytes behind the word
            const int rem = len - 4 * q;
            if (rem < 4) wb[q] = rem <= 0 ? 0u : (wb[q] & ((1u << (8 * rem)) - 1u));
        }
    }
    const int nmax = __reduce_max_sync(FULL, n);
    uint32_t* const col = S.st + threadIdx.x;  // this thread's column: row p at col[p * PBL_THREADS]

    // ---- rows: phantom keys (len_dp[i] = i, dp_tokenize.py:28; not reachable), 0 inside a character ------------------
    {
        uint32_t u = 0;  // unit index of the row being written
#pragma unroll
        for (int p = 0; p < PBL_ROWS; ++p) {
            bool bnd;
            if (!kSpm) {
                bnd = true;
            } else if (p <= 1 || p == n) {
                bnd = true;  // in front of the marker, in front of the body (a stray continuation byte starts a character there), end
            } else {
                bnd = (pbl_byte<NW>(wb, p >= 1 ? p - 1 : 0) & 0xC0u) != 0x80u;
            }
            uint32_t key = 0;
            if (p <= n && bnd) key = p == 0 ? 63u : ((u << 7) | 0x7Fu);
            if (p <= nmax) col[p * PBL_THREADS] = key;
            if (p <= n && bnd) ++u;
        }
    }

    // ---- forward: warp-uniform loop over the start positions ----------------------------------------------------
    const uint32_t* __restrict__ da = P.V.da;
    uint32_t Eprev = 0;     // SPM: end positions relaxed from the previous active start (out-of-vocabulary test)
    unsigned long long Eprev2 = 0;  // ... the same for the 64-position instantiation
    bool have_prev = false, oov = false;
#pragma unroll 1
    for (int s = 0; s < nmax; ++s) {
        const uint32_t kj = (s < n ? col[s * PBL_THREADS] : 0u) & 0xFFFFu;
        const bool active = kj != 0u;  // a unit boundary of this lane's word
        if (kSpm && active) {
            // the previous character [prev start, s) must itself be a vocabulary entry (else the normalised text spells it "<0xHH>")
            if (have_prev && !(NW == 8 ? (Eprev >> s) & 1u : (uint32_t)(Eprev2 >> s) & 1u)) oov = true;
            Eprev = 0;
            Eprev2 = 0;
            have_prev = s >= 1;  // the marker (start 0) is always there (rule precondition)
        }
        const uint32_t hi1 = (kj & 0xFFC0u) + 0x80u, low = kj & 0x3Fu;
        uint32_t entry = DPT_DA_ROOT_ENTRY, cl = 0;
        int p = s;
        bool alive = active;
        // relax the edge s -> p (cl units) for the lanes with `term`; rows inside a character hold key 0 and never lose
        auto relax = [&](const bool term) {
            if (__any_sync(FULL, term)) {
                uint32_t* const rp = col + p * PBL_THREADS;
                const uint32_t row = *rp;
                const uint32_t bi = row & 0xFFFFu;
                const uint32_t lowe = cl ^ 0x3Fu;  // 63 - cl
                const uint32_t k = hi1 + (low < lowe ? low : lowe);
                const uint32_t d = (uint32_t)(p - s);
                const bool pa = term && k <= (bi | 0x3Fu);  // (k >> 6) <= (bi >> 6)
                const bool pb = term && k <= bi;
                uint32_t nr = (row & 0xFF00FFFFu) | (d << 16);          // A = d
                if (pb) nr = (nr & 0x00FF0000u) | k | (d << 24);       // key = k, B = d
                if (pa) *rp = nr;
                // (the edge exists whether or not it improves the row: the out-of-vocabulary test asks for existence)
                if (kSpm && NW == 8 && term && bi != 0u) Eprev |= 1u << p;
                if (kSpm && NW == 16 && term && bi != 0u) Eprev2 |= 1ull << p;
            }
        };
        if (kSpm && s == 0) {  // start 0 begins behind U+2581 (its trie node is part of the compiled vocabulary)
            entry = P.V.marker_entry;
            cl = 1;
            p = 1;
            relax(alive && (entry & DPT_DA_TERMINAL));
        }
        // The walk: groups of four steps peel the bytes off a running copy of the window (a fully unrolled 31-step walk let
        // the compiler keep every step's derived values live at once: > 128 registers).
        uint32_t ww[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) ww[q] = wb[q];
        bool open = true;
#pragma unroll 1
        for (int g = 0; g < NW && open; ++g) {
            uint32_t wcur = ww[0];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = wcur & 0xFFu;
                wcur >>= 8;
                const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
                alive = alive && p < n && base != 0;
                const uint32_t e = pbl_ldg_if(da + base + c, alive);
                alive = alive && (e & DPT_DA_MATCH_MASK) == (DPT_DA_OCCUPIED | c);
                if (!__any_sync(FULL, alive)) {
                    open = false;
                    break;
                }
                if (alive) {
                    entry = e;
                    ++p;
                    if (kSpm) cl += (c & 0xC0u) != 0x80u ? 1u : 0u;
                }
                if (!kSpm) cl = (uint32_t)(p - s);
                relax(alive && (e & DPT_DA_TERMINAL));
            }
#pragma unroll
            for (int q = 0; q < NW - 1; ++q) ww[q] = ww[q + 1];
            ww[NW - 1] = 0;
        }
        if (!(kSpm && s == 0)) {  // the window moves on by one byte (start 0 of an SPM word reads the body from its first byte, like start 1)
#pragma unroll
            for (int q = 0; q < NW - 1; ++q) wb[q] = __funnelshift_r(wb[q], wb[q + 1], 8);
            wb[NW - 1] >>= 8;
        }
    }
    if (kSpm && valid && have_prev && !(NW == 8 ? (Eprev >> n) & 1u : (uint32_t)(Eprev2 >> n) & 1u)) oov = true;
    if (kSpm) {  // hand the words with out-of-vocabulary characters to the thread-per-word kernel
        oov = oov && valid;
        const uint32_t mo = __ballot_sync(FULL, oov);
        if (mo) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&P.ctl->n_defer, (uint32_t)__popc(mo));
            base = __shfl_sync(FULL, base, 0);
            if (oov) P.defer[base + (uint32_t)__popc(mo & ((1u << lane) - 1u))] = slot_item;
        }
        if (oov) valid = false;
    }

    // ---- result: length, flags, ids ------------------------------------------------------------------------------------
    const uint32_t kn = valid ? (col[n * PBL_THREADS] & 0xFFFFu) : 0xFFFFu;
    const uint32_t wl = kn >> 7;
    const bool reach = valid && !(kn & 0x40u);
    const uint32_t target = 63u - (kn & 0x3Fu);
    const bool pooled = reach && wl > (uint32_t)RES_INLINE;
    unsigned long long poff = 0;
    {   // one pool allocation per warp
        uint32_t inc = pooled ? wl : 0u;
        const uint32_t mine = inc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.persist->pool_used, (unsigned long long)total);
            base = __shfl_sync(FULL, base, 0);
            poff = base + (inc - mine);
        }
    }
    ResRec* const rec = &P.res[slot_item];
    int i = reach ? n : 0, o = (int)wl;
    bool got = false;
    const uint8_t* __restrict__ body = P.text + pos;
#pragma unroll 1
    while (__any_sync(FULL, i > 0)) {
        uint32_t d = 0;
        if (i > 0) {
            const uint32_t ab = col[i * PBL_THREADS] >> 16;
            d = got ? (ab & 0xFFu) : (ab >> 8);
            if (d == 0 || (int)d > i) {  // cannot happen on a reachable path; never spin on corrupt state
                i = 0;
                d = 0;
            }
        }
        const int j = i - (int)d;
        // re-walk the token [j, i): its trie slot gives the id
        uint32_t entry = DPT_DA_ROOT_ENTRY, slot = 0, cl = 0;
        int q = j - m, qe = i - m;  // body byte range
        if (kSpm && j == 0 && d) {
            entry = P.V.marker_entry;
            slot = P.V.marker_slot;
            cl = 1;
            q = 0;
        }
        const int steps = d ? qe - q : 0;
        const int smax = __reduce_max_sync(FULL, steps);
        for (int r = 0; r < smax; ++r) {
            if (r < steps) {
                const uint32_t c = body[q + r];
                slot = (entry >> DPT_DA_BASE_SHIFT) + c;
                entry = __ldg(da + slot);
                cl += (!kSpm || (c & 0xC0u) != 0x80u) ? 1u : 0u;
            }
        }
        if (d) {
            const int32_t id = __ldg(P.V.slot_id + slot);
            --o;
            if (pooled) {
                if ((int64_t)(poff + (unsigned long long)o) < P.pool_cap) P.pool[poff + (unsigned long long)o] = id;
            } else if (o >= 0 && o < RES_INLINE) {
                rec->ids[o] = id;
            }
            if (!got && cl == target) got = true;
            i = j;
        }
    }
    if (valid) {
        rec->meta = (wl & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | (pooled ? RES_POOLED : 0u);
        if (pooled) {
            rec->ids[0] = (int32_t)(uint32_t)(poff & 0xFFFFFFFFull);
            rec->ids[1] = (int32_t)(uint32_t)(poff >> 32);
        }
    }
}

// Persistent CTAs: a chunk of CHUNK words of the concatenated queues (NW = 8: classes 2, 1, 0; NW = 16: class 3) from a
// ticket -> counting sort by length in shared memory -> the warps solve its batches of 32, interleaved so every warp gets
// long and short ones.
template <bool kSpm, int NW>
__device__ __forceinline__ void pbl_kernel(const PipeParams& P, PblSmem<NW>& S) {
    constexpr int ROWS = PblCfg<NW>::ROWS, CHUNK = PblCfg<NW>::CHUNK, PER = CHUNK / PBL_THREADS, BINS_PER_LANE = ROWS / 32;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nq2 = NW == 8 ? pb_queue_len(P, 2) : pb_queue_len(P, 3);
    const uint32_t nq1 = NW == 8 ? pb_queue_len(P, 1) : 0u, nq0 = NW == 8 ? pb_queue_len(P, 0) : 0u;
    const int cls_first = NW == 8 ? 2 : 3;
    const uint32_t total = nq2 + nq1 + nq0;
    // Words per chunk: CHUNK when there is work for every CTA; with fewer words (a 16 MB range of a chunked call adds
    // ~30 k distinct words, a single document a few hundred) smaller chunks - down to one batch of 32 - so that the words
    // spread over all CTAs and the kernel takes the time of ONE batch instead of the 8 a full chunk holds (0.17 ms per
    // launch however few the words were: the chunked host path paid it once per range, profiles/r2_e2e_timeline_n1.txt).
    uint32_t cw = (uint32_t)CHUNK;
    if (total < gridDim.x * (uint32_t)CHUNK) {
        cw = ((total + gridDim.x - 1) / gridDim.x + 31u) & ~31u;
        cw = cw < 32u ? 32u : cw > (uint32_t)CHUNK ? (uint32_t)CHUNK : cw;
    }
    const uint32_t n_chunks = (total + cw - 1) / cw;
    unsigned int* const ticket = NW == 8 ? &P.ctl->lock_ticket : &P.ctl->lock_ticket2;
    for (;;) {
        if (tid == 0) S.chunk = atomicAdd(ticket, 1u);
        if (tid <= ROWS) S.hist[tid] = 0;
        __syncthreads();
        const uint32_t chunk = S.chunk;
        if (chunk >= n_chunks) break;
        const uint32_t c0 = chunk * cw;
        const int count = (int)(total - c0 < cw ? total - c0 : cw);
        // this thread's words of the chunk
        uint32_t slot[PER];
        unsigned long long tag[PER];
        uint32_t rank[PER];
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int k = tid + r * PBL_THREADS;
            slot[r] = 0;
            tag[r] = 0;
            rank[r] = 0;
            if (k < count) {
                uint32_t g = c0 + (uint32_t)k;
                int cls = cls_first;
                if (g >= nq2 + nq1) {
                    cls = 0;
                    g -= nq2 + nq1;
                } else if (g >= nq2) {
                    cls = 1;
                    g -= nq2;
                }
                slot[r] = P.pending[(size_t)cls * (size_t)P.pend_stride + g];
            }
        }
#pragma unroll
        for (int r = 0; r < PER; ++r)
            if (tid + r * PBL_THREADS < count) tag[r] = P.tags[slot[r]];
#pragma unroll
        for (int r = 0; r < PER; ++r)
            if (tid + r * PBL_THREADS < count) {
                const int len = pp_tag_len(tag[r]);
                rank[r] = atomicAdd(&S.hist[len < ROWS ? len : ROWS], 1u);
            }
        __syncthreads();
        if (warp == 0) {  // bins -> start offsets, longest words first: lane l owns the lengths ROWS-1 - BINS_PER_LANE l - q
            uint32_t h[BINS_PER_LANE], mine = 0;
#pragma unroll
            for (int q = 0; q < BINS_PER_LANE; ++q) {
                h[q] = S.hist[ROWS - 1 - (lane * BINS_PER_LANE + q)];
                mine += h[q];
            }
            uint32_t inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += o;
            }
            uint32_t at = inc - mine;
#pragma unroll
            for (int q = 0; q < BINS_PER_LANE; ++q) {
                S.hist[ROWS - 1 - (lane * BINS_PER_LANE + q)] = at;
                at += h[q];
            }
            if (lane == 31) S.hist[ROWS] = inc;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PER; ++r)
            if (tid + r * PBL_THREADS < count) {
                const int len = pp_tag_len(tag[r]);
                const uint32_t at = S.hist[len < ROWS ? len : ROWS] + rank[r];
                if (at < (uint32_t)CHUNK) {
                    S.slot[at] = slot[r];
                    S.tag[at] = tag[r];
                }
            }
        __syncthreads();
        for (int b = warp; b * 32 < count; b += PBL_THREADS / 32) pbl_batch<kSpm, NW>(P, S, b * 32, count, lane);
        __syncthreads();
    }
}

}  // namespace dpt
