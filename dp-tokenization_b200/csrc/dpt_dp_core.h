// The shortest-tokenization DP on one word, host/device.
//
// Closed-form restatement of /root/reference/packages/dp_tokenize.py:24-84:
//   forward  len_dp[i] = min(i, min_{j<i, s[j:i] in V} len_dp[j]+1)          (:27-47, phantom init :28)
//   reach[i] = some optimal predecessor chain reaches 0                        (paths dropped at :66-69)
//   M[i]     = longest single token (in units of `charlen`) on any optimal complete path to i
//   select   = first segmentation in the DFS order of :57-69 (largest split first) whose longest
//              token equals M[n]                                               (:82-84)
// All three forward quantities are packed into ONE ordered 64-bit key per position
//      key = len << 32 | notreach << 31 | (0x7FFFFFFF - M)
// so the relaxation over an incoming edge is a single unsigned min.  Two predecessor distances are
// kept per position:  A = largest j among edges with minimal (len, notreach)   -> "got" branch
//                     B = largest j among edges attaining the minimal key      -> "not yet got" branch
// which makes the backward selection a pointer chase independent of M[n].
#pragma once
#include "dpt_common.h"

#define DPT_KEY_LOW 0x7FFFFFFFull
#define DPT_KEY_NOTREACH 0x80000000ull

DPT_HD uint64_t dpt_key_phantom(uint32_t unit_index) { return ((uint64_t)unit_index << 32) | 0xFFFFFFFFull; }
DPT_HD uint64_t dpt_key_origin() { return DPT_KEY_LOW; }
DPT_HD uint64_t dpt_key_extend(uint64_t kj, uint32_t charlen) {
    const uint64_t lowj = kj & DPT_KEY_LOW;
    const uint64_t lowe = DPT_KEY_LOW - charlen;
    return (kj & ~DPT_KEY_LOW) + (1ull << 32) + (lowj < lowe ? lowj : lowe);
}
DPT_HD uint32_t dpt_key_len(uint64_t k) { return (uint32_t)(k >> 32); }
DPT_HD bool dpt_key_reach(uint64_t k) { return (k & DPT_KEY_NOTREACH) == 0; }
DPT_HD uint32_t dpt_key_longest(uint64_t k) { return (uint32_t)(DPT_KEY_LOW - (k & DPT_KEY_LOW)); }

// Forward pass over the normalised bytes s[0..n).  best/A/B have n+1 entries.
// kEmit=false skips A/B (count pass).
template <bool kEmit>
DPT_HD void dpt_forward(const DptVocabView& V, const uint8_t* s, int32_t n, const uint8_t* unit_starts,
                        uint64_t* best, uint16_t* A, uint16_t* B) {
    const bool bytes_mode = V.unit_mode == 0 && unit_starts == nullptr;
    uint32_t u = 0;
    for (int32_t p = 0; p <= n; ++p) {
        const bool b = (p == 0 || p == n) ? true
                       : unit_starts      ? unit_starts[p] != 0
                       : bytes_mode       ? true
                                          : dpt_is_cp_start(s[p]);
        best[p] = b ? dpt_key_phantom(u) : ~0ull;
        if (b) ++u;
        if (kEmit) {
            A[p] = 0;
            B[p] = 0;
        }
    }
    best[0] = dpt_key_origin();
    for (int32_t j = 0; j < n; ++j) {
        const uint64_t kj = best[j];
        if (kj == ~0ull) continue;  // not a unit boundary
        uint32_t entry = DPT_DA_ROOT_ENTRY;
        uint32_t cl = 0;
        const int32_t stop = (n - j) < (int32_t)V.lmax ? n : j + (int32_t)V.lmax;
        for (int32_t i = j + 1; i <= stop; ++i) {
            const uint32_t c = s[i - 1];
            if (!dpt_da_step(V.da, entry, c)) break;
            cl += (V.unit_mode == 0 || dpt_is_cp_start(c)) ? 1u : 0u;
            if ((entry & DPT_DA_TERMINAL) && best[i] != ~0ull) {
                const uint64_t k = dpt_key_extend(kj, cl);
                const uint64_t bi = best[i];
                if (kEmit && (k >> 31) <= (bi >> 31)) A[i] = (uint16_t)(i - j);
                if (k <= bi) {
                    best[i] = k;
                    if (kEmit) B[i] = (uint16_t)(i - j);
                }
            }
        }
    }
}

// Backward selection + id emission.  Writes len tokens to out_ids[0..len) in text order (filled from
// the back).  Returns false (and writes nothing) when the word is untokenizable.
DPT_HD bool dpt_backward_emit(const DptVocabView& V, const uint8_t* s, int32_t n, const uint64_t* best,
                              const uint16_t* A, const uint16_t* B, int32_t* out_ids, int64_t out_cap) {
    const uint64_t kn = best[n];
    if (!dpt_key_reach(kn)) return false;
    const uint32_t target = dpt_key_longest(kn);
    int64_t slot = (int64_t)dpt_key_len(kn) - 1;
    bool got = false;
    int32_t i = n;
    while (i > 0) {
        const int32_t d = got ? A[i] : B[i];
        if (d <= 0 || d > i) break;  // cannot happen on a reachable path; never spin on corrupt state
        const int32_t j = i - d;
        DptHashState h = dpt_hash_init(V.ph_salt);
        uint32_t cl = 0;
        for (int32_t p = j; p < i; ++p) {
            const uint32_t c = s[p];
            dpt_hash_byte(h, c);
            cl += (V.unit_mode == 0 || dpt_is_cp_start(c)) ? 1u : 0u;
        }
        if (!got && cl == target) got = true;
        if (slot >= 0 && slot < out_cap) out_ids[slot] = dpt_ph_lookup(V, h);
        --slot;
        i = j;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Warp-friendly form of the same DP.  dpt_forward above is a nest of data-dependent loops (start position x trie
// walk); on the GPU the lanes of a warp leave those loops at different times and run the rest of the word
// serialised (measured: 1.7 of 32 lanes active).  Here the forward pass is ONE loop whose every iteration is one
// trie step of the lane's current (start, end) pair - a state machine - so the lanes of a warp stay converged and
// a lane only idles for the difference between its total step count and the warp's maximum.  The slot of the
// winning edge is kept beside each back-pointer (As/Bs), so the backward pass needs no re-hash of the token
// bytes: id = slot_id[slot].
// ---------------------------------------------------------------------------------------------------------
DPT_HD void dpt_forward_flat(const DptVocabView& V, const uint8_t* s, int32_t n, uint64_t* best, uint16_t* A, uint16_t* B,
                             uint32_t* As, uint32_t* Bs) {
    const bool cp_mode = V.unit_mode != 0;
    uint32_t u = 0;
    for (int32_t p = 0; p <= n; ++p) {
        const bool b = (p == 0 || p == n || !cp_mode) ? true : dpt_is_cp_start(s[p]);
        best[p] = b ? dpt_key_phantom(u) : ~0ull;
        if (b) ++u;
        A[p] = 0;
        B[p] = 0;
    }
    if (n > 0) best[0] = dpt_key_origin();
    const uint32_t* __restrict__ da = V.da;
    int32_t j = -1, i = 0;
    uint32_t entry = 0, cl = 0;
    uint64_t kj = 0;
    bool walking = false;
    for (;;) {
        if (!walking) {
            if (++j >= n) break;
            kj = best[j];
            entry = DPT_DA_ROOT_ENTRY;
            i = j;
            cl = 0;
            walking = kj != ~0ull;  // not a unit boundary: nothing starts here
            if (!walking) continue;
        }
        const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
        const uint32_t c = i < n ? (uint32_t)s[i] : 0x100u;
        const uint32_t slot = base + (c & 0xFFu);
        uint32_t e = 0;
        if (base != 0 && c < 0x100u) {
#if defined(__CUDA_ARCH__)
            e = __ldg(da + slot);
#else
            e = da[slot];
#endif
        }
        if ((e & DPT_DA_MATCH_MASK) != (DPT_DA_OCCUPIED | c)) {
            walking = false;
            continue;
        }
        entry = e;
        ++i;
        cl += (!cp_mode || dpt_is_cp_start(c)) ? 1u : 0u;
        if (e & DPT_DA_TERMINAL) {
            const uint64_t bi = best[i];
            if (bi != ~0ull) {
                const uint64_t k = dpt_key_extend(kj, cl);
                const uint32_t packed = slot | ((cl < 1023u ? cl : 1023u) << 22);  // slots < 2^22 (DPT_DA_MAX_SLOTS)
                if ((k >> 31) <= (bi >> 31)) {
                    A[i] = (uint16_t)(i - j);
                    As[i] = packed;
                }
                if (k <= bi) {
                    best[i] = k;
                    B[i] = (uint16_t)(i - j);
                    Bs[i] = packed;
                }
            }
        }
    }
}

// Backward selection for dpt_forward_flat: writes the len ids to out_ids[0..len) in text order.  Returns false
// (and writes nothing) when the word is untokenizable.
DPT_HD bool dpt_backward_flat(const DptVocabView& V, const uint8_t* s, int32_t n, const uint64_t* best, const uint16_t* A,
                              const uint16_t* B, const uint32_t* As, const uint32_t* Bs, int32_t* out_ids, int64_t out_cap) {
    const uint64_t kn = best[n];
    if (!dpt_key_reach(kn)) return false;
    const bool cp_mode = V.unit_mode != 0;
    const uint32_t target = dpt_key_longest(kn);
    int64_t slot_out = (int64_t)dpt_key_len(kn) - 1;
    bool got = false;
    int32_t i = n;
    while (i > 0 && slot_out >= 0) {
        const int32_t d = got ? A[i] : B[i];
        if (d <= 0 || d > i) break;  // cannot happen on a reachable path; never spin on corrupt state
        const uint32_t ts = got ? As[i] : Bs[i];
        const int32_t j = i - d;
        if (!got) {
            uint32_t cl = ts >> 22;
            if (cl == 1023u) {  // saturated: count (tokens of >= 1023 units only)
                cl = (uint32_t)d;
                if (cp_mode) {
                    cl = 0;
                    for (int32_t p = j; p < i; ++p) cl += dpt_is_cp_start(s[p]) ? 1u : 0u;
                }
            }
            if (cl == target) got = true;
        }
        if (slot_out < out_cap) out_ids[slot_out] = V.slot_id[ts & 0x3FFFFFu];
        --slot_out;
        i = j;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Compact-state variant for words of at most DPT_FLAT32_MAX normalised bytes (kernel B's per-thread local state):
// the same ordered key in 32 bits   len << 17 | notreach << 16 | (0xFFFF - M),   back-pointers as 1-byte
// distances, 14 bytes per position instead of 20.
// ---------------------------------------------------------------------------------------------------------
#define DPT_FLAT32_MAX 255
#define DPT_K32_NONE 0xFFFFFFFFu
DPT_HD uint32_t dpt_k32_extend(uint32_t kj, uint32_t cl) {
    const uint32_t lowj = kj & 0xFFFFu, lowe = 0xFFFFu - cl;
    return (kj & 0xFFFF0000u) + (1u << 17) + (lowj < lowe ? lowj : lowe);
}
DPT_HD uint32_t dpt_k32_len(uint32_t k) { return k >> 17; }
DPT_HD bool dpt_k32_reach(uint32_t k) { return (k & 0x10000u) == 0; }
DPT_HD uint32_t dpt_k32_longest(uint32_t k) { return 0xFFFFu - (k & 0xFFFFu); }

// trie slot load: read-only path, L1 evict_last (the DP state streaming through L1 as local memory must not push the
// trie out: the trie lookup is the loop-carried dependency of the forward pass)
DPT_HD uint32_t dpt_da_load(const uint32_t* da, uint32_t slot) {
#if defined(__CUDA_ARCH__)
    uint32_t e;
    asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(e) : "l"(da + slot));
    return e;
#else
    return da[slot];
#endif
}

// The forward pass as a resumable state machine: dpt_flat32_init sets up the per-position arrays, every call of
// dpt_flat32_step advances the (start j, end i) walk by at most one trie step; the pass is complete when j == n.
// Kernel B keeps one of these per lane and steps all lanes of a warp in lock step (dpt_pipe.h: pb_thread).
// No flags: entry == 0 means "no walk open" (an open walk always holds an occupied trie entry).
struct DptFlat32 {
    int32_t j, i;          // current start position, current end position
    uint32_t entry, cl, kj;
};
DPT_HD bool dpt_flat32_running(const DptFlat32& st, int32_t n) { return st.j < n; }
// Ap/Bp: packed back-pointers  slot | distance << 22  (distance <= DPT_FLAT32_MAX, slots < 2^22): one 4-byte store
// per back-pointer and relaxation instead of a distance and a slot each
// upos[p] (optional) = unit index of position p (n <= DPT_FLAT32_MAX): the backward chase reads a token's unit count off
// it instead of counting code-point starts byte by byte
DPT_HD void dpt_flat32_init(const DptVocabView& V, const uint8_t* s, int32_t n, uint32_t* best, uint32_t* Ap, uint32_t* Bp,
                            DptFlat32& st, uint8_t* upos = nullptr) {
    const bool cp_mode = V.unit_mode != 0;
    uint32_t u = 0;
    for (int32_t p = 0; p <= n; ++p) {
        const bool b = (p == 0 || p == n || !cp_mode) ? true : dpt_is_cp_start(s[p]);
        best[p] = b ? ((u << 17) | 0x1FFFFu) : DPT_K32_NONE;  // phantom: len = unit index, not reachable
        if (upos) upos[p] = (uint8_t)u;
        if (b) ++u;
        Ap[p] = 0;
        Bp[p] = 0;
    }
    if (n > 0) best[0] = 0xFFFFu;  // origin: len 0, reachable, longest 0
    st.j = -1;
    st.i = 0;
    st.entry = 0;
    st.cl = 0;
    st.kj = 0;
}
DPT_HD void dpt_flat32_step(const DptVocabView& V, const uint8_t* s, int32_t n, uint32_t* best, uint32_t* Ap, uint32_t* Bp,
                            DptFlat32& st) {
    const bool cp_mode = V.unit_mode != 0;
    if (st.entry == 0) {  // next start position (same iteration as its first trie step)
        if (++st.j >= n) return;
        st.kj = best[st.j];
        if (st.kj == DPT_K32_NONE) return;  // not a unit boundary: nothing starts here
        st.entry = DPT_DA_ROOT_ENTRY;
        st.i = st.j;
        st.cl = 0;
    }
    const uint32_t base = st.entry >> DPT_DA_BASE_SHIFT;
    const uint32_t c = (uint32_t)s[st.i];  // i < n whenever a walk is open
    const uint32_t slot = base + c;
    uint32_t e = 0;
    if (base != 0) e = dpt_da_load(V.da, slot);
    if ((e & DPT_DA_MATCH_MASK) != (DPT_DA_OCCUPIED | c)) {
        st.entry = 0;
        return;
    }
    const int32_t i = ++st.i;
    st.cl += (!cp_mode || dpt_is_cp_start(c)) ? 1u : 0u;
    if (e & DPT_DA_TERMINAL) {
        const uint32_t bi = best[i];
        if (bi != DPT_K32_NONE) {
            const uint32_t k = dpt_k32_extend(st.kj, st.cl);
            const uint32_t packed = slot | ((uint32_t)(i - st.j) << 22);
            if ((k >> 16) <= (bi >> 16)) Ap[i] = packed;
            if (k <= bi) {
                best[i] = k;
                Bp[i] = packed;
            }
        }
    }
    st.entry = i >= n ? 0u : e;  // end of the word: the walk from j is over
}
DPT_HD void dpt_forward_flat32(const DptVocabView& V, const uint8_t* s, int32_t n, uint32_t* best, uint32_t* Ap,
                               uint32_t* Bp) {
    DptFlat32 st;
    dpt_flat32_init(V, s, n, best, Ap, Bp, st);
    while (dpt_flat32_running(st, n)) dpt_flat32_step(V, s, n, best, Ap, Bp, st);
}

// Backward chase for dpt_forward_flat32: word_len ids into out_ids[0..word_len) in text order.  The code-point count of
// a token is only needed until the token of length `target` has been taken.
DPT_HD void dpt_backward_chase(const DptVocabView& V, const uint8_t* s, int32_t n, uint32_t word_len, uint32_t target,
                               const uint32_t* Ap, const uint32_t* Bp, int32_t* out_ids, int64_t out_cap,
                               const uint8_t* upos = nullptr) {
    const bool cp_mode = V.unit_mode != 0;
    int64_t slot_out = (int64_t)word_len - 1;
    bool got = false;
    int32_t i = n;
    while (i > 0 && slot_out >= 0) {
        const uint32_t ts = got ? Ap[i] : Bp[i];
        const int32_t d = (int32_t)(ts >> 22);
        if (d <= 0 || d > i) break;  // cannot happen on a reachable path; never spin on corrupt state
        if (!got) {
            uint32_t cl = (uint32_t)d;
            if (cp_mode) {
                if (upos) {  // tokens start and end on unit boundaries
                    cl = (uint32_t)upos[i] - (uint32_t)upos[i - d];
                } else {
                    cl = 0;
                    for (int32_t p = i - d; p < i; ++p) cl += dpt_is_cp_start(s[p]) ? 1u : 0u;
                }
            }
            if (cl == target) got = true;
        }
        if (slot_out < out_cap) out_ids[slot_out] = V.slot_id[ts & 0x3FFFFFu];
        --slot_out;
        i -= d;
    }
}
