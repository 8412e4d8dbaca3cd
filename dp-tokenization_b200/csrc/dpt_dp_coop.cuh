// Kernel B of the corpus pipeline: the shortest-tokenization DP (dp_tokenize.py:24-84 in the closed form of
// dpt_dp_core.h), SUB-WARP COOPERATIVE.  One tile of T = 8 / 16 / 32 lanes per distinct word (4 / 2 / 1 words per warp),
// everything in registers and 4 KB of shared memory per warp, no local memory:
//
//   lanes      lane t of a tile owns the t-th byte of the word (SPM rule: lane 0 owns the word-initial U+2581 marker
//              as ONE lane, the body bytes follow; byte-level rules: lane t = byte t).  Position t = "in front of lane
//              t"; position n = the end of the word.
//   walks      lane t walks the double-array trie from position t - one 4-byte load per step, all starts of all
//              words of the warp in flight at once, the next byte fetched from its owner lane with one shuffle -
//              and records the ends of the vocabulary entries it passes as a bit mask E_t (bit i: s[t:i] in V) and
//              the trie slot of every step in shared memory (column = lane: no conflicts, no synchronisation).
//   forward    for i = 1..n: every lane offers  extend(best_t, units(t,i))  if bit i of E_t is set; one
//              redux.sync.min over the tile gives best_i = min(phantom_i, ...) (phantom = the reference's
//              len_dp[i] = i initialisation, dp_tokenize.py:28); lane i keeps it as the value of its own start.
//              The ordered 32-bit key is  len << 17 | notreach << 16 | (0xFFFF - longest token)  (dpt_dp_core.h).
//   backward   from i = n: the predecessor is the HIGHEST lane whose offer equals best_i (before a token of the
//              target length has been taken: full key, "B" pointer; after: len|notreach only, "A" pointer) - one
//              ballot + clz per token, i.e. the DFS order "largest split first, first maximum wins" of
//              dp_tokenize.py:57-69,82-84 - and the owning lane turns its edge into an id: slot_id[slot of that step].
//
// Words with an out-of-vocabulary character under the SPM rule (normalised text differs from marker + raw bytes: the
// character is spelled "<0xHH>" per byte, tokenizer_utils.py:26-29) are handed to the thread-per-word kernel through
// the `defer` list; so are words longer than a warp (length class 3) and the odd words.
#pragma once
#include "dpt_pipe.h"

namespace dpt {

constexpr int PBC_THREADS = 128;  // 4 warps; 16 KB of slot columns per CTA

__device__ __forceinline__ uint32_t pbc_extend(uint32_t kj, uint32_t cl) {
    const uint32_t lowj = kj & 0xFFFFu, lowe = 0xFFFFu - cl;
    return (kj & 0xFFFF0000u) + (1u << 17) + (lowj < lowe ? lowj : lowe);
}

// One batch: 32 / T words, one per tile.  `slot_item` = the word's table slot (tile-uniform), `valid` = the tile holds a word.
constexpr unsigned long long PBC_POOL_CHUNK = 512;  // ids a warp takes from the pool per atomic (holes are harmless: records hold offsets)

template <int T>
__device__ __forceinline__ void pbc_batch(const PipeParams& P, uint32_t* __restrict__ slotcol, const uint32_t slot_item,
                                          bool valid, const int lane, unsigned long long& pool_cur,
                                          unsigned long long& pool_end) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t TM = T == 32 ? 0xFFFFFFFFu : ((1u << T) - 1u);
    const int t = lane & (T - 1), tb = lane & ~(T - 1);
    const uint32_t tmask = TM << tb;
    const bool spm = P.spm != 0;
    const bool cp = P.V.unit_mode != 0;
    const int m = spm ? 1 : 0;

    // ---- the word: raw bytes [pos, pos + len); lanes in use n = len + m ----------------------------------------
    int64_t pos = 0;
    int n = 0;
    if (valid) {
        const unsigned long long tag = P.tags[slot_item];  // same address across the tile: one transaction
        pos = pp_tag_pos(tag);
        n = pp_tag_len(tag) + m;
        if (n > (T == 32 ? 31 : T)) {  // cannot happen (the class says so); never index out of the tile
            valid = false;
            n = 0;
        }
    }
    uint32_t byte = 0;
    if (t >= m && t < n) byte = P.text[pos + t - m];
    const bool isstart = t < n && (t < m || !cp || (byte & 0xC0u) != 0x80u);
    // unit boundaries of the tile (bit p: position p is one), position n included
    const uint32_t Bm = ((__ballot_sync(FULL, isstart) >> tb) & TM) | (valid ? (1u << n) : 0u);

    // ---- walks ---------------------------------------------------------------------------------------------
    uint32_t entry = DPT_DA_ROOT_ENTRY, E = 0;
    int nxt = t;  // tile-local index of the next byte this lane's walk consumes
    bool alive = isstart;
    if (spm && t == 0 && valid) {  // the marker lane starts behind U+2581 (its trie node is part of the compiled vocabulary)
        entry = P.V.marker_entry;
        nxt = 1;
        if (entry & DPT_DA_TERMINAL) E = 2u;  // edge 0 -> 1: the bare marker
    }
    const int nxt0 = nxt;
    const uint32_t* __restrict__ da = P.V.da;
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
        const uint32_t c = __shfl_sync(FULL, byte, tb + (nxt & (T - 1)));
        const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
        alive = alive && nxt < n && base != 0;
        const uint32_t slot = base + c;
        uint32_t e = 0;
        if (alive) e = __ldg(da + slot);
        alive = alive && (e & DPT_DA_MATCH_MASK) == (DPT_DA_OCCUPIED | c);
        if (alive) {
            entry = e;
            ++nxt;
            if (e & DPT_DA_TERMINAL) E |= 1u << nxt;
            slotcol[k * 32 + lane] = slot;
        }
        if (!__any_sync(FULL, alive)) break;
    }
    E &= Bm;  // a vocabulary entry never ends inside a character (same test as best[i] != NONE in dpt_flat32_step)

    // ---- SPM rule: a character that is no vocabulary entry is spelled "<0xHH>" in the normalised text -> not this kernel ----
    if (spm) {
        bool oov = false;
        if (isstart && t >= 1) {
            const uint32_t above = Bm >> (t + 1);  // bit n is set: never zero for t < n
            const int nb = t + 1 + (__ffs((int)above) - 1);
            oov = ((E >> nb) & 1u) == 0u;
        }
        const uint32_t any_oov = __ballot_sync(FULL, oov) & tmask;
        if (any_oov && valid) {
            if (t == 0) {
                const uint32_t q = atomicAdd(&P.ctl->n_defer, 1u);
                P.defer[q] = slot_item;  // (capacity pend_stride: every deferred word owns a queue entry)
            }
            valid = false;
        }
    }

    // ---- forward relaxation ------------------------------------------------------------------------------------
    const uint32_t Ut = (uint32_t)__popc(Bm & ((1u << t) - 1u));  // unit index of position t
    uint32_t best = t == 0 ? 0xFFFFu : DPT_K32_NONE;                // origin: len 0, reachable, longest 0
    uint32_t bestN = DPT_K32_NONE;
    const int nmax = __reduce_max_sync(FULL, n);
#pragma unroll 1
    for (int i = 1; i <= nmax; ++i) {
        const uint32_t Ui = (uint32_t)__popc(Bm & ((1u << i) - 1u));
        uint32_t cand = DPT_K32_NONE;
        if ((E >> i) & 1u) cand = pbc_extend(best, Ui - Ut);
        const uint32_t kmin = __reduce_min_sync(tmask, cand);
        const uint32_t ph = ((Bm >> i) & 1u) ? ((Ui << 17) | 0x1FFFFu) : DPT_K32_NONE;  // phantom: len = unit index, not reachable
        const uint32_t nb = kmin < ph ? kmin : ph;
        if (t == i) best = nb;
        if (i == n) bestN = nb;
    }

    // ---- backward selection + ids ------------------------------------------------------------------------------
    const uint32_t wl = dpt_k32_len(bestN);
    const bool reach = valid && dpt_k32_reach(bestN);
    const uint32_t target = dpt_k32_longest(bestN);
    const bool pooled = reach && wl > (uint32_t)RES_INLINE;
    // ids of words with more than RES_INLINE tokens go to the pool: the warp sub-allocates from a private chunk (one
    // global atomic per PBC_POOL_CHUNK ids instead of one per word: 0.3 M words would queue on one L2 address)
    unsigned long long poff = 0;
    {
        uint32_t before = 0, need = 0;  // ids of the pooled tiles in front of this one / of all tiles of the warp
#pragma unroll
        for (int q = 0; q < 32 / T; ++q) {
            const uint32_t w = __shfl_sync(FULL, pooled ? wl : 0u, q * T);
            if (q * T < tb) before += w;
            need += w;
        }
        if (need) {  // warp-uniform
            if (pool_cur + need > pool_end) {
                const unsigned long long take = need > PBC_POOL_CHUNK ? (unsigned long long)need : PBC_POOL_CHUNK;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&P.persist->pool_used, take);
                pool_cur = __shfl_sync(FULL, base, 0);
                pool_end = pool_cur + take;
            }
            poff = pool_cur + before;
            pool_cur += need;
        }
    }
    ResRec* const rec = &P.res[slot_item];
    int i = reach ? n : 0;
    int o = (int)wl;
    bool got = false;
    uint32_t cur = bestN;
    while (__any_sync(FULL, i > 0)) {
        const uint32_t Ui = (uint32_t)__popc(Bm & ((1u << i) - 1u));
        uint32_t cand = DPT_K32_NONE;
        if (i > 0 && ((E >> i) & 1u)) cand = pbc_extend(best, Ui - Ut);
        const bool sel = cand != DPT_K32_NONE && (got ? (cand >> 16) == (cur >> 16) : cand == cur);
        const uint32_t mm = (__ballot_sync(FULL, sel) >> tb) & TM;
        const int j = mm ? 31 - __clz((int)mm) : 0;  // mm != 0 on a reachable path; never spin on corrupt state
        const uint32_t bj = __shfl_sync(FULL, best, tb + j);
        if (i > 0) {
            --o;
            if (mm && t == j) {  // this lane owns the chosen edge j -> i
                const uint32_t slot = (spm && t == 0 && i == 1) ? P.V.marker_slot : slotcol[(i - nxt0 - 1) * 32 + lane];
                const int32_t id = __ldg(P.V.slot_id + slot);
                if (pooled) {
                    if ((int64_t)(poff + (unsigned long long)o) < P.pool_cap) P.pool[poff + (unsigned long long)o] = id;
                } else {
                    rec->ids[o] = id;
                }
            }
            const uint32_t Uj = (uint32_t)__popc(Bm & ((1u << j) - 1u));
            if (!got && Ui - Uj == target) got = true;
            cur = bj;
            i = mm ? j : 0;
        }
    }
    if (valid && t == 0) {
        uint32_t meta = (wl & 0xFFFFFFu) | (reach ? 0u : RES_UNTOK) | (pooled ? RES_POOLED : 0u);
        rec->meta = meta;
        if (pooled) {
            rec->ids[0] = (int32_t)(uint32_t)(poff & 0xFFFFFFFFull);
            rec->ids[1] = (int32_t)(uint32_t)(poff >> 32);
        }
    }
}

template <int T>
__device__ __forceinline__ void pbc_class(const PipeParams& P, uint32_t* slotcol, const int cls, const uint32_t nq,
                                          const uint32_t batch, const int lane, unsigned long long& pool_cur,
                                          unsigned long long& pool_end) {
    constexpr int TPW = 32 / T;
    const uint32_t idx = batch * TPW + (uint32_t)(lane / T);
    const bool valid = idx < nq;
    const uint32_t slot = valid ? P.pending[(size_t)cls * (size_t)P.pend_stride + idx] : 0u;
    pbc_batch<T>(P, slotcol, slot, valid, lane, pool_cur, pool_end);
}

// Persistent warps over the batches of the three tile widths, widest first (their words cost the most: the tail of the
// kernel is made of the cheapest).  Static striding: the words of one class cost about the same, and a claim per batch
// would be 300 k atomics on three addresses.
__device__ __forceinline__ void pbc_kernel(const PipeParams& P, uint32_t* slot_smem) {
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    uint32_t* slotcol = slot_smem + warp * (32 * 32);
    const uint32_t nq2 = pb_queue_len(P, 2), nq1 = pb_queue_len(P, 1), nq0 = pb_queue_len(P, 0);
    const uint32_t nb2 = nq2, nb1 = (nq1 + 1) / 2, nb0 = (nq0 + 3) / 4;
    const uint32_t total = nb2 + nb1 + nb0;
    const uint32_t nw = gridDim.x * (blockDim.x >> 5);
    unsigned long long pool_cur = 0, pool_end = 0;
    for (uint32_t g = blockIdx.x * (blockDim.x >> 5) + warp; g < total; g += nw) {
        if (g < nb2) pbc_class<32>(P, slotcol, 2, nq2, g, lane, pool_cur, pool_end);
        else if (g < nb2 + nb1) pbc_class<16>(P, slotcol, 1, nq1, g - nb2, lane, pool_cur, pool_end);
        else pbc_class<8>(P, slotcol, 0, nq0, g - nb2 - nb1, lane, pool_cur, pool_end);
    }
}

}  // namespace dpt
