// CUDA kernels (sm_100a) + host launch orchestration for the reference-shaped pipeline:
//   boundary/normalise  ->  per-word DP (ids into a stash)  ->  scan  ->  placement of the ids  ->  counters.
// This is the general path: it accepts ANY word (any length, any unit boundaries, out-of-vocab
// characters expanded to "<0xHH>" text).  It runs one DP per word OCCURRENCE (the ids wait in a stash at the word's byte
// offset until the scan over the token counts has placed them) and is what
// dpt_encode_words (pre-split words from a host pre-tokenizer) and dpt_encode_corpus_general (the cross-check of the
// deduplicating corpus pipeline in pipe.cu) are made of; the throughput path is the pipeline.  k_lattice, k_min_tokens
// (one serial thread) are known-answer entry points, not throughput paths; k_roundtrip checks one document per warp.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dptok.h"
#include "dpt_decode.h"
#include "dpt_dp_core.h"
#include "dpt_rules.h"
#include "kernels.h"
#include "vocab.h"

namespace dpt {

std::atomic<int64_t> g_launches{0};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// block-wide helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// exclusive scan over the block (blockDim.x multiple of 32, <= 1024); total returned to all threads.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem /* >= 33 entries */, T& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const T inc = warp_inclusive_scan(v);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarps ? smem[lane] : T(0);
        const T winc = warp_inclusive_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    const T base = smem[warp];
    total = smem[32];
    __syncthreads();
    return base + inc - v;
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_doc_start_bits(const int64_t* __restrict__ doc_offs, int64_t n_docs, uint32_t* __restrict__ bits) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const int64_t p = doc_offs[d];
    atomicOr(&bits[p >> 5], 1u << (p & 31));
}

// In-place exclusive scan of `cols` interleaved int64 columns over `n` rows (row-major [n][cols]),
// single block.  Totals go to totals[c].  n is a tile count (thousands), so one block is plenty.
__global__ void k_scan_rows(int64_t* __restrict__ rows, int64_t n, int cols, int64_t* __restrict__ totals) {
    __shared__ int64_t sm[33];
    for (int c = 0; c < cols; ++c) {
        int64_t carry = 0;
        for (int64_t base = 0; base < n; base += blockDim.x) {
            const int64_t i = base + threadIdx.x;
            const int64_t v = i < n ? rows[i * cols + c] : 0;
            int64_t tot;
            const int64_t ex = block_exclusive_scan<int64_t>(v, sm, tot);
            if (i < n) rows[i * cols + c] = carry + ex;
            carry += tot;
        }
        if (threadIdx.x == 0) totals[c] = carry;
    }
}

// ---------------------------------------------------------------------------------------------
// SPM_LLAMA normaliser: raw documents -> "<s>" + U+2581 + text with ' '->U+2581 and out-of-vocab
// characters spelled "<0xHH>"; emits word offsets into the normalised text (dpt_rules.h).
// Tile = SPM_THREADS x SPM_PER bytes; pass 1 counts (bytes, words, docs) per tile, pass 2 writes.
// ---------------------------------------------------------------------------------------------
constexpr int SPM_THREADS = 256;
constexpr int SPM_PER = 8;
constexpr int SPM_TILE = SPM_THREADS * SPM_PER;

struct SpmCounts {
    int32_t bytes, words, docs;
};

__device__ __forceinline__ SpmCounts spm_chunk_counts(const DptVocabView& V, const uint8_t* text, int64_t n,
                                                      const uint32_t* doc_bits, int64_t p0) {
    SpmCounts c{0, 0, 0};
    for (int k = 0; k < SPM_PER; ++k) {
        const int64_t p = p0 + k;
        if (p >= n) break;
        if (!dpt_spm_is_char_start(text, doc_bits, p)) continue;
        const DptSpmChar ch = dpt_spm_classify(V, text, n, doc_bits, p);
        c.bytes += ch.out_len;
        if (dpt_bit_test(doc_bits, p)) {
            c.bytes += 6;  // "<s>" + U+2581
            c.words += 2;
            c.docs += 1;
        } else if (ch.marker && !dpt_spm_prev_is_marker(text, doc_bits, p)) {
            c.words += 1;
        }
    }
    return c;
}

__global__ void __launch_bounds__(SPM_THREADS)
k_spm_count(DptVocabView V, const uint8_t* __restrict__ text, int64_t n, const uint32_t* __restrict__ doc_bits,
            int64_t* __restrict__ tile_sums /* [tiles][3] */) {
    __shared__ int32_t sm[3][SPM_THREADS / 32];
    const int64_t p0 = (int64_t)blockIdx.x * SPM_TILE + (int64_t)threadIdx.x * SPM_PER;
    SpmCounts c = spm_chunk_counts(V, text, n, doc_bits, p0);
    int32_t v[3] = {c.bytes, c.words, c.docs};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int32_t x = v[q];
#pragma unroll
        for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0) sm[q][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int64_t s = 0;
        for (int w = 0; w < SPM_THREADS / 32; ++w) s += sm[threadIdx.x][w];
        tile_sums[(int64_t)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(SPM_THREADS)
k_spm_write(DptVocabView V, const uint8_t* __restrict__ text, int64_t n, const uint32_t* __restrict__ doc_bits,
            const int64_t* __restrict__ tile_base /* scanned [tiles][3] */, uint8_t* __restrict__ out, int64_t out_cap,
            int64_t* __restrict__ word_offs, int64_t word_cap, int64_t* __restrict__ norm_doc_offs,
            int64_t* __restrict__ doc_first_word, uint8_t* __restrict__ doc_flags) {
    __shared__ int32_t sm[33];
    const int64_t p0 = (int64_t)blockIdx.x * SPM_TILE + (int64_t)threadIdx.x * SPM_PER;
    const SpmCounts c = spm_chunk_counts(V, text, n, doc_bits, p0);
    int32_t tot;
    int64_t ob = tile_base[(int64_t)blockIdx.x * 3 + 0] + block_exclusive_scan<int32_t>(c.bytes, sm, tot);
    int64_t ow = tile_base[(int64_t)blockIdx.x * 3 + 1] + block_exclusive_scan<int32_t>(c.words, sm, tot);
    int64_t od = tile_base[(int64_t)blockIdx.x * 3 + 2] + block_exclusive_scan<int32_t>(c.docs, sm, tot);
    for (int k = 0; k < SPM_PER; ++k) {
        const int64_t p = p0 + k;
        if (p >= n) break;
        if (!dpt_spm_is_char_start(text, doc_bits, p)) continue;
        const DptSpmChar ch = dpt_spm_classify(V, text, n, doc_bits, p);
        const bool doc_start = dpt_bit_test(doc_bits, p);
        if (doc_start) {
            norm_doc_offs[od] = ob;
            doc_first_word[od] = ow;
            if (ow < word_cap) word_offs[ow] = ob;
            if (ow + 1 < word_cap) word_offs[ow + 1] = ob + 3;
            if (ob + 6 <= out_cap) {
                out[ob + 0] = '<';
                out[ob + 1] = 's';
                out[ob + 2] = '>';
                out[ob + 3] = DPT_MARK0;
                out[ob + 4] = DPT_MARK1;
                out[ob + 5] = DPT_MARK2;
            }
            if (ch.marker && doc_flags) doc_flags[od] = DPT_DF_AMBIGUOUS;
            ob += 6;
            ow += 2;
            od += 1;
        } else if (ch.marker) {
            if (!dpt_spm_prev_is_marker(text, doc_bits, p)) {
                if (ow < word_cap) word_offs[ow] = ob;
                ow += 1;
            } else if (doc_flags) {
                // benign race: several threads may set the same bit of the same byte
                doc_flags[od - 1] = DPT_DF_AMBIGUOUS;
            }
        }
        if (ob + ch.out_len <= out_cap) dpt_spm_write_char(text, p, ch, out + ob);
        ob += ch.out_len;
    }
}

__global__ void k_spm_finish(const int64_t* __restrict__ totals /* bytes, words, docs */, int64_t* __restrict__ word_offs,
                             int64_t word_cap, int64_t* __restrict__ norm_doc_offs, int64_t n_docs,
                             int64_t* __restrict__ n_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (totals[1] <= word_cap) word_offs[totals[1]] = totals[0];
        norm_doc_offs[n_docs] = totals[0];
        n_out[0] = totals[1];
        n_out[1] = totals[0];
    }
}

// ---------------------------------------------------------------------------------------------
// per-word DP, one thread per word (short words: state in local memory; long: global scratch)
// ---------------------------------------------------------------------------------------------
constexpr int DP_THREADS = 128;
constexpr int DP_LOCAL_CAP = 64;  // bytes; longer words take the long path

struct LongCtl {
    unsigned long long n_long;       // number of long words
    unsigned long long pool_used;    // scratch positions handed out
};

// One DP per word: forward pass with back-pointers, token count and flags out, and the selected ids into
// stash[a .. a + len) (a = the word's byte offset: a word of n bytes has at most n tokens, so the stash of one int32 per
// text byte never overlaps).  k_dp_emit only moves them once the scan has placed the word.  Words lying beyond the
// stash (a caller whose n_text_bytes understates the text) are solved again in k_dp_emit instead.
__global__ void __launch_bounds__(DP_THREADS)
k_dp_count(DptVocabView V, const uint8_t* __restrict__ text, const int64_t* __restrict__ word_offs, int64_t n_words,
           int32_t* __restrict__ lens, uint8_t* __restrict__ flags, const uint8_t* __restrict__ flags_in,
           int32_t* __restrict__ long_list, int64_t* __restrict__ long_scratch, LongCtl* ctl,
           int32_t* __restrict__ stash, int64_t stash_cap) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const int64_t a = word_offs[w], b = word_offs[w + 1];
    const int64_t n = b - a;
    uint8_t f = flags_in ? (flags_in[w] & DPT_WF_DOC_FIRST) : 0;
    if (n <= 0) {
        lens[w] = 0;
        flags[w] = f | DPT_WF_UNTOKENIZABLE;
        return;
    }
    if (n > DP_LOCAL_CAP) {
        const unsigned long long slot = atomicAdd(&ctl->n_long, 1ull);
        const unsigned long long off = atomicAdd(&ctl->pool_used, (unsigned long long)(n + 1));
        long_list[slot] = (int32_t)w;
        long_scratch[slot] = (int64_t)off;
        lens[w] = 0;
        flags[w] = f | DPT_WF_LONG;
        return;
    }
    uint64_t best[DP_LOCAL_CAP + 1];
    uint16_t A[DP_LOCAL_CAP + 1], B[DP_LOCAL_CAP + 1];
    dpt_forward<true>(V, text + a, (int32_t)n, nullptr, best, A, B);
    const uint64_t kn = best[n];
    lens[w] = (int32_t)dpt_key_len(kn);
    flags[w] = f | (dpt_key_reach(kn) ? 0 : DPT_WF_UNTOKENIZABLE);
    if (a >= 0 && a + n <= stash_cap) dpt_backward_emit(V, text + a, (int32_t)n, best, A, B, stash + a, n);
}

// long words: one thread each, state in the global pool (12 bytes per position).
__global__ void __launch_bounds__(DP_THREADS)
k_dp_long(DptVocabView V, const uint8_t* __restrict__ text, const int64_t* __restrict__ word_offs,
          const int32_t* __restrict__ long_list, const int64_t* __restrict__ long_scratch, const LongCtl* ctl,
          uint64_t* __restrict__ pool_best, uint16_t* __restrict__ pool_a, uint16_t* __restrict__ pool_b,
          int64_t pool_cap, int32_t* __restrict__ lens, uint8_t* __restrict__ flags, int emit,
          const int64_t* __restrict__ tok_offs, int32_t* __restrict__ ids, int64_t ids_cap) {
    const unsigned long long n_long = ctl->n_long;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n_long;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const int64_t w = long_list[k];
        const int64_t a = word_offs[w], n = word_offs[w + 1] - a;
        const int64_t off = long_scratch[k];
        if (off + n + 1 > pool_cap) {  // pool overflow: reported through n_out[3]; word left flagged
            if (!emit) {
                lens[w] = 0;
                flags[w] |= DPT_WF_UNTOKENIZABLE;
            }
            continue;
        }
        uint64_t* best = pool_best + off;
        uint16_t* A = pool_a + off;
        uint16_t* B = pool_b + off;
        if (!emit) {
            dpt_forward<true>(V, text + a, (int32_t)n, nullptr, best, A, B);
            const uint64_t kn = best[n];
            lens[w] = (int32_t)dpt_key_len(kn);
            if (!dpt_key_reach(kn)) flags[w] |= DPT_WF_UNTOKENIZABLE;
        } else if (!(flags[w] & DPT_WF_UNTOKENIZABLE)) {
            const int64_t o = tok_offs[w];
            dpt_backward_emit(V, text + a, (int32_t)n, best, A, B, ids + o, ids_cap - o);
        }
    }
}

// token-count tile sums: one block per DP_SCAN_TILE words; also accumulates the counters.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER;

__device__ __forceinline__ int32_t word_tokens(const int32_t* lens, const uint8_t* flags, int64_t w) {
    return (flags[w] & DPT_WF_UNTOKENIZABLE) ? 0 : lens[w];
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_len_tile_sums(const int32_t* __restrict__ lens, const uint8_t* __restrict__ flags, int64_t n_words,
                int64_t* __restrict__ tile_sums, unsigned long long* __restrict__ counters) {
    __shared__ int32_t sm[2][SCAN_THREADS / 32];
    int32_t t = 0, u = 0;
    const int64_t w0 = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t w = w0 + (int64_t)k * SCAN_THREADS;
        if (w < n_words) {
            t += word_tokens(lens, flags, w);
            u += (flags[w] & DPT_WF_UNTOKENIZABLE) ? 1 : 0;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, d);
        u += __shfl_xor_sync(0xffffffffu, u, d);
    }
    if ((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = t;
        sm[1][threadIdx.x >> 5] = u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t ts = 0, us = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) {
            ts += sm[0][k];
            us += sm[1][k];
        }
        tile_sums[blockIdx.x] = ts;
        if (us) atomicAdd(&counters[DPT_CTR_UNTOKENIZABLE], (unsigned long long)us);
    }
}

// expands scanned tile bases into per-word token offsets (int64[n_words+1])
__global__ void __launch_bounds__(SCAN_THREADS)
k_tok_offsets(const int32_t* __restrict__ lens, const uint8_t* __restrict__ flags, int64_t n_words,
              const int64_t* __restrict__ tile_base, const int64_t* __restrict__ total, int64_t* __restrict__ tok_offs) {
    __shared__ int32_t sm[33];
    // thread handles SCAN_PER consecutive words so the scan is in word order
    const int64_t w0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER;
    int32_t v[SCAN_PER];
    int32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t w = w0 + k;
        v[k] = w < n_words ? word_tokens(lens, flags, w) : 0;
        s += v[k];
    }
    int32_t tot;
    int64_t o = tile_base[blockIdx.x] + block_exclusive_scan<int32_t>(s, sm, tot);
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t w = w0 + k;
        if (w < n_words) tok_offs[w] = o;
        o += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) tok_offs[n_words] = total[0];
}

__global__ void __launch_bounds__(DP_THREADS)
k_dp_emit(DptVocabView V, const uint8_t* __restrict__ text, const int64_t* __restrict__ word_offs, int64_t n_words,
          const int32_t* __restrict__ lens, const uint8_t* __restrict__ flags, const int64_t* __restrict__ tok_offs,
          const int32_t* __restrict__ stash, int64_t stash_cap, int32_t* __restrict__ ids, int64_t ids_cap) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    if (flags[w] & (DPT_WF_UNTOKENIZABLE | DPT_WF_LONG)) return;
    const int64_t a = word_offs[w];
    const int32_t n = (int32_t)(word_offs[w + 1] - a);
    const int64_t o = tok_offs[w];
    if (a >= 0 && a + n <= stash_cap) {  // the ids k_dp_count left at the word's byte offset
        const int32_t len = lens[w];
        for (int32_t k = 0; k < len && o + k < ids_cap; ++k) ids[o + k] = stash[a + k];
        return;
    }
    uint64_t best[DP_LOCAL_CAP + 1];
    uint16_t A[DP_LOCAL_CAP + 1], B[DP_LOCAL_CAP + 1];
    dpt_forward<true>(V, text + a, n, nullptr, best, A, B);
    dpt_backward_emit(V, text + a, n, best, A, B, ids + o, ids_cap - o);
}

__global__ void k_finish_counters(unsigned long long* __restrict__ counters, int64_t n_bytes, int64_t n_words,
                                  const int64_t* __restrict__ total_tokens, const LongCtl* ctl, int64_t pool_cap,
                                  int64_t* __restrict__ n_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        counters[DPT_CTR_BYTES] = (unsigned long long)n_bytes;
        counters[DPT_CTR_WORDS] = (unsigned long long)n_words;
        counters[DPT_CTR_TOKENS] = (unsigned long long)total_tokens[0];
        n_out[DPT_NOUT_IDS] = total_tokens[0];
        n_out[DPT_NOUT_WORDS] = n_words;
        n_out[DPT_NOUT_POOL_REQ] = (int64_t)ctl->pool_used;
        n_out[DPT_NOUT_POOL_CAP] = pool_cap;
    }
}

__global__ void k_doc_tok_offs(const int64_t* __restrict__ doc_first_word, int64_t n_docs,
                               const int64_t* __restrict__ tok_offs, int64_t n_words, int64_t* __restrict__ doc_tok_offs) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    doc_tok_offs[d] = d == n_docs ? tok_offs[n_words] : tok_offs[doc_first_word[d]];
}

// Length-only DP with infinity initialisation (inspect_tokenizer.py:77-86 `min_tokens_for_string`): differs from the
// lattice's len_dp only where the phantom initialisation of dp_tokenize.py:28 shows (untokenizable input).  One serial
// thread like k_lattice: a known-answer entry point, not a throughput path.  out[0] = dp[n], -1 = infinity.
__global__ void k_min_tokens(DptVocabView V, const uint8_t* __restrict__ s, int32_t n, const uint8_t* __restrict__ unit_starts,
                             int32_t* __restrict__ out, int32_t* __restrict__ dp /* n+1 scratch, by byte position */) {
    if (threadIdx.x || blockIdx.x) return;
    const int32_t INF = 0x3FFFFFFF;
    for (int32_t p = 0; p <= n; ++p) {
        const bool b = (p == 0 || p == n) ? true
                       : unit_starts      ? unit_starts[p] != 0
                       : V.unit_mode == 0 ? true
                                          : dpt_is_cp_start(s[p]);
        dp[p] = b ? INF : -1;  // -1: not a unit boundary
    }
    dp[0] = 0;
    for (int32_t j = 0; j < n; ++j) {
        if (dp[j] < 0 || dp[j] >= INF) continue;
        uint32_t entry = DPT_DA_ROOT_ENTRY;
        for (int32_t i = j + 1; i <= n; ++i) {
            if (!dpt_da_step(V.da, entry, s[i - 1])) break;
            if ((entry & DPT_DA_TERMINAL) && dp[i] >= 0 && dp[j] + 1 < dp[i]) dp[i] = dp[j] + 1;
        }
    }
    out[0] = dp[n] >= INF ? -1 : dp[n];
}


// ---------------------------------------------------------------------------------------------
// lattice of one word (enumerate-all API, dp_tokenize.py:27-47): single thread, tiny inputs
// ---------------------------------------------------------------------------------------------
__global__ void k_lattice(DptVocabView V, const uint8_t* __restrict__ s, int32_t n, const uint8_t* __restrict__ unit_starts,
                          int32_t* __restrict__ len_dp, int32_t* __restrict__ pred_offs, int32_t* __restrict__ pred,
                          int32_t pred_cap, int32_t* __restrict__ n_out, int32_t* __restrict__ unit_of /* n+1 scratch */) {
    if (threadIdx.x || blockIdx.x) return;
    // unit index of every boundary byte position
    int32_t nu = 0;
    for (int32_t p = 0; p <= n; ++p) {
        const bool b = (p == 0 || p == n) ? true
                       : unit_starts      ? unit_starts[p] != 0
                       : V.unit_mode == 0 ? true
                                          : dpt_is_cp_start(s[p]);
        unit_of[p] = b ? nu : -1;
        if (b) {
            len_dp[nu] = nu;
            ++nu;
        }
    }
    const int32_t n_units = nu - 1;
    // pass 1: len_dp (push relaxations; predecessor order does not matter for the minimum)
    for (int32_t j = 0; j < n; ++j) {
        if (unit_of[j] < 0) continue;
        uint32_t entry = DPT_DA_ROOT_ENTRY;
        for (int32_t i = j + 1; i <= n; ++i) {
            if (!dpt_da_step(V.da, entry, s[i - 1])) break;
            if ((entry & DPT_DA_TERMINAL) && unit_of[i] >= 0) {
                const int32_t c = len_dp[unit_of[j]] + 1;
                if (c < len_dp[unit_of[i]]) len_dp[unit_of[i]] = c;
            }
        }
    }
    // pass 2 counts predecessors per end unit, pass 3 fills them; the outer loop runs j ascending so
    // every list comes out ascending like segment_index_dp (dp_tokenize.py:38-47).
    for (int32_t u = 0; u <= n_units + 1; ++u) pred_offs[u] = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int32_t j = 0; j < n; ++j) {
            if (unit_of[j] < 0) continue;
            uint32_t entry = DPT_DA_ROOT_ENTRY;
            for (int32_t i = j + 1; i <= n; ++i) {
                if (!dpt_da_step(V.da, entry, s[i - 1])) break;
                if ((entry & DPT_DA_TERMINAL) && unit_of[i] >= 0 &&
                    len_dp[unit_of[j]] + 1 == len_dp[unit_of[i]]) {
                    const int32_t ui = unit_of[i];
                    if (pass == 0) {
                        pred_offs[ui + 1] += 1;
                    } else {
                        const int32_t q = pred_offs[ui]++;  // cursor of unit ui
                        if (q < pred_cap) pred[q] = unit_of[j];
                    }
                }
            }
        }
        if (pass == 0) {
            for (int32_t u = 0; u <= n_units; ++u) pred_offs[u + 1] += pred_offs[u];  // pred_offs[u] = start(u)
            n_out[1] = pred_offs[n_units + 1];
        }
    }
    // each cursor now sits at the end of its list == start of the next: shift back into CSR starts
    for (int32_t u = n_units + 1; u >= 1; --u) pred_offs[u] = pred_offs[u - 1];
    pred_offs[0] = 0;
    n_out[0] = n_units;
}

// ---------------------------------------------------------------------------------------------
// decode + round-trip check on device (tokenizer_utils.py:82-84,176-179; asserts at
// main_analyze_s2orc.py:85, main_biomed_translation.py:78).  One WARP per document, 32 tokens per round: every lane
// takes the decoded length of its token (dpt_decode.h: U+2581 -> ' ', "<0xHH>" -> the byte, the one leading space of the
// Prepend normaliser dropped), a warp scan places the tokens in the raw document, every lane compares its bytes.
// ---------------------------------------------------------------------------------------------
constexpr int RT_THREADS = 128;

__global__ void __launch_bounds__(RT_THREADS)
k_roundtrip(DptVocabView V, const int32_t* __restrict__ ids, const int64_t* __restrict__ doc_tok_offs,
            const uint8_t* __restrict__ text, const int64_t* __restrict__ doc_offs, int64_t n_docs, int32_t skip_bos,
            uint8_t* __restrict__ ok) {
    const unsigned lane = threadIdx.x & 31;
    const int64_t d = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;  // the whole warp
    const bool spm = V.unit_mode == 1;
    const int64_t t0 = doc_tok_offs[d] + (skip_bos ? 1 : 0), t1 = doc_tok_offs[d + 1];
    int64_t p = doc_offs[d];
    const int64_t pe = doc_offs[d + 1];
    bool good = true;
    for (int64_t base = t0; base < t1 && good; base += 32) {
        const int64_t t = base + lane;
        int64_t a = 0, b = 0;
        bool fine = true;
        int32_t dl = 0;
        if (t < t1) {
            fine = dpt_tok_span(V, ids[t], a, b);
            if (fine) dl = dpt_tok_decoded_len(V, a, b, spm, t == t0);
        }
        const int32_t inc = warp_inclusive_scan(dl);
        if (fine && t < t1) fine = dpt_tok_matches(V, a, b, spm, t == t0, text, p + inc - dl, pe);
        good = __all_sync(0xffffffffu, fine);
        p += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) ok[d] = (good && p == pe) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// training-data feed on device: token stream -> padded int64 batch.
// One block per row.  Row r = ids of document doc_begin + r of stream A, optionally followed by the ids of the
// same document of stream B (the custom collator of main_biomed_translation.py:104-124 concatenates
// input_ids + labels), truncated to row_len, padded with pad_id on the right (or left); attention_mask is 1 on
// real tokens (main_analyze_s2orc.py:88-89) and 0 on padding (DataCollatorWithPadding).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pad_batch(const int32_t* __restrict__ ids_a, const int64_t* __restrict__ offs_a, const int32_t* __restrict__ ids_b,
            const int64_t* __restrict__ offs_b, int64_t doc_begin, int64_t row_len, int64_t pad_id, int32_t pad_left,
            int64_t* __restrict__ input_ids, int64_t* __restrict__ attention_mask, int64_t* __restrict__ row_lens) {
    const int64_t r = blockIdx.x, d = doc_begin + r;
    const int64_t a0 = offs_a[d], na = offs_a[d + 1] - a0;
    const int64_t b0 = ids_b ? offs_b[d] : 0, nb = ids_b ? offs_b[d + 1] - b0 : 0;
    int64_t n = na + nb;
    if (n > row_len) n = row_len;
    const int64_t shift = pad_left ? row_len - n : 0;
    int64_t* out = input_ids + r * row_len;
    int64_t* am = attention_mask ? attention_mask + r * row_len : nullptr;
    for (int64_t p = threadIdx.x; p < row_len; p += blockDim.x) {
        const int64_t q = p - shift;
        const bool real = q >= 0 && q < n;
        int64_t v = pad_id;
        if (real) v = q < na ? (int64_t)ids_a[a0 + q] : (int64_t)ids_b[b0 + (q - na)];
        out[p] = v;
        if (am) am[p] = real ? 1 : 0;
    }
    if (threadIdx.x == 0 && row_lens) row_lens[r] = n;
}

// ---------------------------------------------------------------------------------------------
// Compact output: int32 ids -> uint16 (vocabularies of at most 65,536 entries).  8 ids per thread: two 16-byte loads,
// one 16-byte store; the count is read from device memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_narrow_u16(const int32_t* __restrict__ ids, const int64_t* __restrict__ d_n, int64_t cap, uint16_t* __restrict__ out,
             int64_t* __restrict__ overflow) {
    int64_t n = *d_n;
    if (n > cap) n = cap;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    uint32_t v[8];
    const bool vec = i0 + 8 <= n && ((((uintptr_t)ids) & 15u) == 0) && ((((uintptr_t)out) & 15u) == 0);
    if (vec) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(ids + i0));
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(ids + i0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i0 + k < n ? (uint32_t)ids[i0 + k] : 0u;
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (v[k] > 0xFFFFu) {
            v[k] = 0xFFFFu;
            ++bad;
        }
    if (bad && overflow) atomicAdd(reinterpret_cast<unsigned long long*>(overflow), (unsigned long long)bad);
    if (vec) {
        uint4 o;
        o.x = v[0] | (v[1] << 16); o.y = v[2] | (v[3] << 16); o.z = v[4] | (v[5] << 16); o.w = v[6] | (v[7] << 16);
        __stcs(reinterpret_cast<uint4*>(out + i0), o);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k < n) out[i0 + k] = (uint16_t)v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

ProfScope::ProfScope(const char* n, cudaStream_t s) : name(n), st(s) {
    if (g_prof_on.load(std::memory_order_relaxed)) {
        cudaEventCreate(&a);
        cudaEventRecord(a, st);
    }
}
ProfScope::~ProfScope() {
    if (a) {
        cudaEvent_t b;
        cudaEventCreate(&b);
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back({name, a, b});
    }
}

void profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

// "name launches total_ms\n" per kernel, sorted by total time; clears the records.  Synchronises.
std::string profile_report() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<int64_t, double>> agg;
    const bool list = getenv("DPT_PROF_LIST") != nullptr;  // development: every launch, in order, on stderr
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (list) fprintf(stderr, "[dpt prof] %-22s %.4f ms\n", r.name, ms);
        auto& e = agg[r.name];
        e.first += 1;
        e.second += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    std::vector<std::pair<std::string, std::pair<int64_t, double>>> v(agg.begin(), agg.end());
    std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.second.second > y.second.second; });
    std::string out;
    char line[256];
    for (auto& e : v) {
        snprintf(line, sizeof line, "%s %lld %.6f\n", e.first.c_str(), (long long)e.second.first, e.second.second);
        out += line;
    }
    return out;
}

#define DPT_LAUNCH(kernel, grid, block, stream, ...)                        \
    do {                                                                    \
        ProfScope _prof(#kernel, (stream));                                 \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);              \
        ++g_launches;                                                       \
    } while (0)

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

struct Carver {
    char* base;
    int64_t cap, used;
    template <typename T>
    T* take(int64_t count) {
        used = align_up(used, 256);
        T* p = (T*)(base + used);
        used += count * (int64_t)sizeof(T);
        return p;
    }
    int64_t remaining() const { return cap - align_up(used, 256); }
};

int64_t encode_words_workspace_fixed(int64_t n_words, int64_t n_text_bytes) {
    const int64_t tiles = (n_words + SCAN_TILE - 1) / SCAN_TILE + 1;
    int64_t b = 0;
    b += align_up(tiles * 8, 256);             // tile sums
    b += align_up(64, 256);                    // totals + ctl
    b += align_up((n_words + 1) * 8, 256);     // tok_offs (when the caller gives none)
    b += align_up(n_words * 4, 256);           // long_list
    b += align_up(n_words * 8, 256);           // long_scratch
    b += align_up((n_text_bytes > 0 ? n_text_bytes : 0) * 4, 256);  // id stash: one int32 per text byte
    return b + 1024;
}

int encode_words(const dpt_vocab* v, const uint8_t* d_text, const int64_t* d_word_offs, int64_t n_words,
                 int64_t n_bytes_for_counter, int64_t text_extent, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens,
                 uint8_t* d_word_flags, int64_t* d_word_tok_offs, int64_t* d_counters, int64_t* d_n_out,
                 void* d_ws, int64_t ws_bytes, cudaStream_t st, std::string& err) {
    if (n_words < 0 || !d_counters || !d_n_out) {
        err = "encode_words: bad argument";
        return DPT_EINVAL;
    }
    if (n_words > 0 && (!d_text || !d_word_offs || !d_word_lens || !d_word_flags || !d_ids)) {
        err = "encode_words: null buffer";
        return DPT_EINVAL;
    }
    if (n_words >= (1ll << 31)) {
        err = "encode_words: more than 2^31-1 words in one call; split the batch";
        return DPT_EINVAL;
    }
    if (text_extent < 0) text_extent = 0;
    const int64_t fixed = encode_words_workspace_fixed(n_words, text_extent);
    if (!d_ws || ws_bytes < fixed) {
        err = "encode_words: workspace too small (see dpt_encode_words_workspace)";
        return DPT_ECAPACITY;
    }
    Carver cv{(char*)d_ws, ws_bytes, 0};
    const int64_t tiles = (n_words + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* tile_sums = cv.take<int64_t>(tiles + 1);
    int64_t* totals = cv.take<int64_t>(4);
    LongCtl* ctl = (LongCtl*)cv.take<int64_t>(2);
    int64_t* tok_offs = d_word_tok_offs ? d_word_tok_offs : cv.take<int64_t>(n_words + 1);
    int32_t* long_list = cv.take<int32_t>(n_words);
    int64_t* long_scratch = cv.take<int64_t>(n_words);
    int32_t* stash = cv.take<int32_t>(text_extent);
    // remaining workspace is the long-word pool: 12 bytes per position
    const int64_t pool_cap = cv.remaining() > 0 ? (cv.remaining() - 1024) / 12 : 0;
    uint64_t* pool_best = cv.take<uint64_t>(pool_cap > 0 ? pool_cap : 0);
    uint16_t* pool_a = cv.take<uint16_t>(pool_cap > 0 ? pool_cap : 0);
    uint16_t* pool_b = cv.take<uint16_t>(pool_cap > 0 ? pool_cap : 0);

    cudaMemsetAsync(d_counters, 0, 4 * sizeof(int64_t), st);
    cudaMemsetAsync(d_n_out, 0, 8 * sizeof(int64_t), st);
    cudaMemsetAsync(ctl, 0, sizeof(LongCtl), st);
    cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), st);
    if (n_words > 0) {
        const DptVocabView& V = v->d_view;
        DPT_LAUNCH(k_dp_count, blocks_for(n_words, DP_THREADS), DP_THREADS, st, V, d_text, d_word_offs, n_words,
                   d_word_lens, d_word_flags, (const uint8_t*)nullptr, long_list, long_scratch, ctl, stash, text_extent);
        DPT_LAUNCH(k_dp_long, 148 * 2, DP_THREADS, st, V, d_text, d_word_offs, long_list, long_scratch, ctl, pool_best,
                   pool_a, pool_b, pool_cap, d_word_lens, d_word_flags, 0, (const int64_t*)nullptr, (int32_t*)nullptr,
                   (int64_t)0);
        DPT_LAUNCH(k_len_tile_sums, (unsigned)tiles, SCAN_THREADS, st, d_word_lens, d_word_flags, n_words, tile_sums,
                   (unsigned long long*)d_counters);
        DPT_LAUNCH(k_scan_rows, 1, 1024, st, tile_sums, tiles, 1, totals);
        DPT_LAUNCH(k_tok_offsets, (unsigned)tiles, SCAN_THREADS, st, d_word_lens, d_word_flags, n_words, tile_sums, totals,
                   tok_offs);
        DPT_LAUNCH(k_dp_emit, blocks_for(n_words, DP_THREADS), DP_THREADS, st, V, d_text, d_word_offs, n_words,
                   d_word_lens, d_word_flags, tok_offs, stash, text_extent, d_ids, ids_cap);
        DPT_LAUNCH(k_dp_long, 148 * 2, DP_THREADS, st, V, d_text, d_word_offs, long_list, long_scratch, ctl, pool_best,
                   pool_a, pool_b, pool_cap, d_word_lens, d_word_flags, 1, tok_offs, d_ids, ids_cap);
    }
    DPT_LAUNCH(k_finish_counters, 1, 32, st, (unsigned long long*)d_counters, n_bytes_for_counter, n_words, totals, ctl,
               pool_cap > 0 ? pool_cap : 0, d_n_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("encode_words: ") + cudaGetErrorString(e);
        return DPT_ECUDA;
    }
    return DPT_OK;
}

int64_t pretokenize_workspace(int64_t n_bytes, int64_t n_docs) {
    const int64_t tiles = (n_bytes + SPM_TILE - 1) / SPM_TILE + 1;
    return align_up((n_bytes / 32 + 2) * 4, 256) + align_up(tiles * 24, 256) + align_up((n_docs + 1) * 8, 256) + 2048;
}

int pretokenize_spm(const dpt_vocab* v, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs, int64_t n_docs,
                    uint8_t* d_norm, int64_t norm_cap, int64_t* d_norm_doc_offs, int64_t* d_word_offs, int64_t word_cap,
                    int64_t* d_doc_first_word, uint8_t* d_doc_flags, int64_t* d_n_out, void* d_ws, int64_t ws_bytes,
                    cudaStream_t st, std::string& err) {
    if (!v->byte_fallback || !v->marker_entry) {
        err = "pretokenize(SPM_LLAMA): vocabulary lacks U+2581 or the 256 <0xHH> byte tokens; pre-split on the host";
        return DPT_EINVAL;
    }
    if (n_bytes <= 0 || n_docs <= 0 || !d_text || !d_doc_offs || !d_norm || !d_norm_doc_offs || !d_word_offs || !d_n_out) {
        err = "pretokenize(SPM_LLAMA): bad argument";
        return DPT_EINVAL;
    }
    if (ws_bytes < pretokenize_workspace(n_bytes, n_docs)) {
        err = "pretokenize: workspace too small";
        return DPT_ECAPACITY;
    }
    Carver cv{(char*)d_ws, ws_bytes, 0};
    const int64_t nbits = n_bytes / 32 + 2;
    const int64_t tiles = (n_bytes + SPM_TILE - 1) / SPM_TILE;
    uint32_t* doc_bits = cv.take<uint32_t>(nbits);
    int64_t* tile_sums = cv.take<int64_t>(tiles * 3 + 3);
    int64_t* totals = cv.take<int64_t>(4);
    int64_t* first_word = d_doc_first_word ? d_doc_first_word : cv.take<int64_t>(n_docs + 1);
    cudaMemsetAsync(doc_bits, 0, nbits * 4, st);
    if (d_doc_flags) cudaMemsetAsync(d_doc_flags, 0, n_docs, st);
    const DptVocabView& V = v->d_view;
    DPT_LAUNCH(k_doc_start_bits, blocks_for(n_docs, 256), 256, st, d_doc_offs, n_docs, doc_bits);
    DPT_LAUNCH(k_spm_count, (unsigned)tiles, SPM_THREADS, st, V, d_text, n_bytes, doc_bits, tile_sums);
    DPT_LAUNCH(k_scan_rows, 1, 1024, st, tile_sums, tiles, 3, totals);
    DPT_LAUNCH(k_spm_write, (unsigned)tiles, SPM_THREADS, st, V, d_text, n_bytes, doc_bits, tile_sums, d_norm, norm_cap,
               d_word_offs, word_cap, d_norm_doc_offs, first_word, d_doc_flags);
    DPT_LAUNCH(k_spm_finish, 1, 32, st, totals, d_word_offs, word_cap, d_norm_doc_offs, n_docs, d_n_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("pretokenize: ") + cudaGetErrorString(e);
        return DPT_ECUDA;
    }
    return DPT_OK;
}

int doc_tok_offsets(const int64_t* d_doc_first_word, int64_t n_docs, const int64_t* d_tok_offs, int64_t n_words,
                    int64_t* d_doc_tok_offs, cudaStream_t st) {
    DPT_LAUNCH(k_doc_tok_offs, blocks_for(n_docs + 1, 256), 256, st, d_doc_first_word, n_docs, d_tok_offs, n_words,
               d_doc_tok_offs);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

int roundtrip_check(const dpt_vocab* v, const int32_t* d_ids, const int64_t* d_doc_tok_offs, const uint8_t* d_text,
                    const int64_t* d_doc_offs, int64_t n_docs, int32_t skip_bos, uint8_t* d_ok, cudaStream_t st) {
    DPT_LAUNCH(k_roundtrip, blocks_for(n_docs * 32, RT_THREADS), RT_THREADS, st, v->d_view, d_ids, d_doc_tok_offs, d_text, d_doc_offs, n_docs,
               skip_bos, d_ok);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

int narrow_ids_u16(const int32_t* d_ids, const int64_t* d_n, int64_t cap, uint16_t* d_out, int64_t* d_overflow, cudaStream_t st) {
    const int64_t threads = (cap + 7) / 8;
    DPT_LAUNCH(k_narrow_u16, (unsigned)((threads + 255) / 256), 256, st, d_ids, d_n, cap, d_out, d_overflow);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

int pad_batch(const int32_t* d_ids_a, const int64_t* d_offs_a, const int32_t* d_ids_b, const int64_t* d_offs_b,
              int64_t doc_begin, int64_t n_rows, int64_t row_len, int64_t pad_id, int32_t pad_left, int64_t* d_input_ids,
              int64_t* d_attention_mask, int64_t* d_row_lens, cudaStream_t st) {
    DPT_LAUNCH(k_pad_batch, (unsigned)n_rows, 128, st, d_ids_a, d_offs_a, d_ids_b, d_offs_b, doc_begin, row_len, pad_id, pad_left,
               d_input_ids, d_attention_mask, d_row_lens);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

int lattice_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts,
                 int32_t* d_len_dp, int32_t* d_pred_offs, int32_t* d_pred, int32_t pred_cap, int32_t* d_n_out,
                 int32_t* d_unit_of, cudaStream_t st) {
    DPT_LAUNCH(k_lattice, 1, 32, st, v->d_view, d_text, n_bytes, d_unit_starts, d_len_dp, d_pred_offs, d_pred, pred_cap,
               d_n_out, d_unit_of);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

int min_tokens_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts, int32_t* d_out,
                    int32_t* d_scratch, cudaStream_t st) {
    DPT_LAUNCH(k_min_tokens, 1, 32, st, v->d_view, d_text, n_bytes, d_unit_starts, d_out, d_scratch);
    return cudaGetLastError() == cudaSuccess ? DPT_OK : DPT_ECUDA;
}

}  // namespace dpt
