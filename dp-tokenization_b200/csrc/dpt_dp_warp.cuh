// Kernel B for the LONG words (32..63 units: length class 3, a few thousand distinct words per 100 MB): one WARP per
// word, lane t owns the units t and t + 32.
//
// The lock-step kernel (dpt_dp_lock.cuh) keeps one word per thread; a word of 60 bytes is then a serial chain of
// ~60 starts x ~8 trie steps x ~50 instructions = 0.13 ms however few such words there are (measured, also beside the
// main kernel on a side stream: the chain only gets slower when it shares the issue slots).  Here the chain is cut by
// working on all starts of the word at once:
//   walks     lane t walks the trie from its two start positions (all 64 walks of the word in flight; the next byte of
//             a walk comes from its owner lane with a shuffle) and keeps the ends of the vocabulary entries it passes as
//             64-bit masks E (bit i: s[start:i] in V);
//   forward   for i = 1..n: every lane offers  extend(best_start, units(start, i))  for its starts with bit i set, ONE
//             full-warp redux.sync.min gives best_i = min(phantom_i, offers) (dp_tokenize.py:27-47 with the phantom
//             initialisation of :28); 32-bit ordered keys  len << 17 | notreach << 16 | (0xFFFF - longest token);
//   backward  from i = n: predecessor = the HIGHEST start whose offer equals best_i (full key until a token of the
//             target length has been taken, len|notreach after: the DFS order and first-maximum rule of
//             dp_tokenize.py:57-69,82-84) - two ballots + clz - and the token's id from re-walking its bytes (uniform
//             across the warp).
// This is the shape north_star sketches for the whole DP; for SHORT words it loses to the lock-step kernel by 3x (dense
// n x 32 lane-slots per word, profiles/rejected/r2_dp_coop.patch), for the long ones it is what removes their latency.
#pragma once
#include "dpt_pipe.h"

namespace dpt {

constexpr int PBW_THREADS = 128;

__device__ __forceinline__ uint32_t pbw_extend(uint32_t kj, uint32_t cl) {
    const uint32_t lowj = kj & 0xFFFFu, lowe = 0xFFFFu - cl;
    return (kj & 0xFFFF0000u) + (1u << 17) + (lowj < lowe ? lowj : lowe);
}
__device__ __forceinline__ uint32_t pbw_units_before(unsigned long long Bm, int p) {  // unit boundaries in front of position p
    return (uint32_t)__popcll(Bm & ((1ull << p) - 1ull));
}

template <bool kSpm>
__device__ __forceinline__ void pbw_word(const PipeParams& P, const uint32_t slot_item, const int lane) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int m = kSpm ? 1 : 0;
    const unsigned long long tag = P.tags[slot_item];
    const int64_t pos = pp_tag_pos(tag);
    const int n = pp_tag_len(tag) + m;  // units = DP positions 0..n, n <= 63
    if (n > 63) return;                 // cannot happen (the class says so)
    // ---- units: lane t owns unit t (h = 0) and unit t + 32 (h = 1) ---------------------------------------------------------
    uint32_t byte[2];
    bool isstart[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int u = lane + 32 * h;
        byte[h] = (u >= m && u < n) ? (uint32_t)P.text[pos + u - m] : 0u;
        // characters start at code-point start bytes and at the first byte of the body (pb_normalise cuts there too)
        isstart[h] = u < n && (u <= m || !kSpm || (byte[h] & 0xC0u) != 0x80u);
    }
    const unsigned long long Bm = (unsigned long long)__ballot_sync(FULL, isstart[0]) |
                                  ((unsigned long long)__ballot_sync(FULL, isstart[1]) << 32) | (1ull << n);
    // ---- walks ---------------------------------------------------------------------------------------------
    const uint32_t* __restrict__ da = P.V.da;
    uint32_t entry[2] = {DPT_DA_ROOT_ENTRY, DPT_DA_ROOT_ENTRY};
    unsigned long long E[2] = {0ull, 0ull};
    int nxt[2] = {lane, lane + 32};
    bool alive[2] = {isstart[0], isstart[1]};
    if (kSpm && lane == 0) {  // unit 0 = U+2581: its walk starts behind the marker's trie node
        entry[0] = P.V.marker_entry;
        nxt[0] = 1;
        if (entry[0] & DPT_DA_TERMINAL) E[0] = 2ull;
    }
#pragma unroll 1
    for (int k = 0; k < 64; ++k) {
        bool any = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t c0 = __shfl_sync(FULL, byte[0], nxt[h] & 31);
            const uint32_t c1 = __shfl_sync(FULL, byte[1], nxt[h] & 31);
            const uint32_t c = (nxt[h] & 32) ? c1 : c0;
            const uint32_t base = entry[h] >> DPT_DA_BASE_SHIFT;
            alive[h] = alive[h] && nxt[h] < n && base != 0;
            uint32_t e = 0;
            if (alive[h]) e = __ldg(da + base + c);
            alive[h] = alive[h] && (e & DPT_DA_MATCH_MASK) == (DPT_DA_OCCUPIED | c);
            if (alive[h]) {
                entry[h] = e;
                ++nxt[h];
                if (e & DPT_DA_TERMINAL) E[h] |= 1ull << nxt[h];
            }
            any = any || alive[h];
        }
        if (!__any_sync(FULL, any)) break;
    }
    E[0] &= Bm;  // a vocabulary entry never ends inside a character
    E[1] &= Bm;
    // ---- SPM rule: a character that is no vocabulary entry is spelled "<0xHH>" -> the thread-per-word kernel --------------
    if (kSpm) {
        bool oov = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int u = lane + 32 * h;
            if (isstart[h] && u >= 1) {
                const unsigned long long above = Bm >> (u + 1);  // bit n is set: never zero for u < n
                const int nb = u + 1 + (__ffsll((long long)above) - 1);
                if (!((E[h] >> nb) & 1ull)) oov = true;
            }
        }
        if (__any_sync(FULL, oov)) {
            if (lane == 0) {
                const uint32_t q = atomicAdd(&P.ctl->n_defer, 1u);
                P.defer[q] = slot_item;
            }
            return;
        }
    }
    // ---- forward relaxation ------------------------------------------------------------------------------------
    const uint32_t Ut[2] = {pbw_units_before(Bm, lane), pbw_units_before(Bm, lane + 32)};
    uint32_t best[2] = {lane == 0 ? 0xFFFFu : DPT_K32_NONE, DPT_K32_NONE};  // origin: len 0, reachable, longest 0
    uint32_t bestN = DPT_K32_NONE;
#pragma unroll 1
    for (int i = 1; i <= n; ++i) {
        const uint32_t Ui = pbw_units_before(Bm, i);
        uint32_t cand = DPT_K32_NONE;
        if ((E[0] >> i) & 1ull) cand = pbw_extend(best[0], Ui - Ut[0]);
        if ((E[1] >> i) & 1ull) {
            const uint32_t c1 = pbw_extend(best[1], Ui - Ut[1]);
            cand = c1 < cand ? c1 : cand;
        }
        const uint32_t kmin = __reduce_min_sync(FULL, cand);
        const uint32_t ph = ((Bm >> i) & 1ull) ? ((Ui << 17) | 0x1FFFFu) : DPT_K32_NONE;  // phantom: len = unit index, not reachable
        const uint32_t nb = kmin < ph ? kmin : ph;
        if (lane == (i & 31)) {
            if (i < 32) best[0] = nb; else best[1] = nb;
        }
        if (i == n) bestN = nb;
    }
    // ---- backward selection + ids ------------------------------------------------------------------------------
    const uint32_t wl = dpt_k32_len(bestN);
    const bool reach = dpt_k32_reach(bestN);
    const uint32_t target = dpt_k32_longest(bestN);
    const bool pooled = reach && wl > (uint32_t)RES_INLINE;
    unsigned long long poff = 0;
    if (pooled) {
        if (lane == 0) poff = atomicAdd(&P.persist->pool_used, (unsigned long long)wl);
        poff = __shfl_sync(FULL, poff, 0);
    }
    ResRec* const rec = &P.res[slot_item];
    int i = reach ? n : 0, o = (int)wl;
    bool got = false;
    uint32_t cur = bestN;
#pragma unroll 1
    while (i > 0) {  // (everything here is warp-uniform: one word per warp)
        const uint32_t Ui = pbw_units_before(Bm, i);
        bool sel[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t cand = DPT_K32_NONE;
            if ((E[h] >> i) & 1ull) cand = pbw_extend(best[h], Ui - Ut[h]);
            sel[h] = cand != DPT_K32_NONE && (got ? (cand >> 16) == (cur >> 16) : cand == cur);
        }
        const unsigned long long mm = (unsigned long long)__ballot_sync(FULL, sel[0]) |
                                      ((unsigned long long)__ballot_sync(FULL, sel[1]) << 32);
        if (!mm) break;  // cannot happen on a reachable path; never spin on corrupt state
        const int j = 63 - __clzll((long long)mm);
        const uint32_t b0 = __shfl_sync(FULL, best[0], j & 31), b1 = __shfl_sync(FULL, best[1], j & 31);
        // the token [j, i): re-walk its bytes (every lane the same walk: uniform loads)
        uint32_t entry_t = DPT_DA_ROOT_ENTRY, slot = 0;
        int q = j;
        if (kSpm && j == 0) {
            entry_t = P.V.marker_entry;
            slot = P.V.marker_slot;
            q = 1;
        }
        for (; q < i; ++q) {
            const uint32_t c0 = __shfl_sync(FULL, byte[0], q & 31), c1 = __shfl_sync(FULL, byte[1], q & 31);
            const uint32_t c = (q & 32) ? c1 : c0;
            slot = (entry_t >> DPT_DA_BASE_SHIFT) + c;
            entry_t = __ldg(da + slot);
        }
        --o;
        if (lane == 0) {
            const int32_t id = __ldg(P.V.slot_id + slot);
            if (pooled) {
                if ((int64_t)(poff + (unsigned long long)o) < P.pool_cap) P.pool[poff + (unsigned long long)o] = id;
            } else if (o >= 0 && o < RES_INLINE) {
                rec->ids[o] = id;
            }
        }
        if (!got && Ui - pbw_units_before(Bm, j) == target) got = true;
        cur = (j & 32) ? b1 : b0;
        i = j;
    }
    if (lane == 0) {
        rec->meta = (wl & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | (pooled ? RES_POOLED : 0u);
        if (pooled) {
            rec->ids[0] = (int32_t)(uint32_t)(poff & 0xFFFFFFFFull);
            rec->ids[1] = (int32_t)(uint32_t)(poff >> 32);
        }
    }
}

template <bool kSpm>
__device__ __forceinline__ void pbw_kernel(const PipeParams& P) {
    const int lane = (int)(threadIdx.x & 31);
    const uint32_t nq = pb_queue_len(P, 3);
    const uint32_t nw = gridDim.x * (blockDim.x >> 5);
    for (uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < nq; g += nw) {
        pbw_word<kSpm>(P, P.pending[(size_t)3 * (size_t)P.pend_stride + g], lane);
        __syncwarp();
    }
}

}  // namespace dpt
