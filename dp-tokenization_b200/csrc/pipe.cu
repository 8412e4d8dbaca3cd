// CUDA kernels (sm_100a) of the corpus pipeline in dpt_pipe.h:  clear -> A scan+dedup -> B DP per distinct word (+ B'
// long words) -> C scan+emit (its last tile writes the counters).  Five launches per corpus, no host synchronisation
// in between.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/dptok.h"
#include "dpt_pipe.h"
#include "dpt_dp_lock.cuh"
#include "dpt_dp_warp.cuh"
#include "kernels.h"
#include "vocab.h"

// resident CTAs per SM the kernels are compiled for (register budgets; tuned on the B200, profiles/README.md)
#ifndef DPT_PA_CTAS
#define DPT_PA_CTAS 5  // SentencePiece kernel A: 8064-byte tiles, 48 registers, 30 KB of shared memory
#endif
#ifndef DPT_PABL_CTAS
#define DPT_PABL_CTAS 3  // byte-level kernel A: 6016-byte tiles, 57 KB of shared memory, <= 64 registers
#endif
#ifndef DPT_PB_CTAS
#define DPT_PB_CTAS 12
#endif
#ifndef DPT_PBL_CTAS
#define DPT_PBL_CTAS 8
#endif
#ifndef DPT_PC_CTAS
#define DPT_PC_CTAS 4
#endif

namespace dpt {

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct DevBlk {
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
#if defined(DPT_TICKET_ORDER)  // tuning variant: tiles from an atomic ticket instead of blockIdx order
    __device__ __forceinline__ bool persistent() const { return true; }
#else
    __device__ __forceinline__ bool persistent() const { return false; }
#endif
    __device__ __forceinline__ int block_index() const { return (int)blockIdx.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) const { atomicOr(p, v); }
    __device__ __forceinline__ void atomic_add(uint32_t* p, uint32_t v) const { atomicAdd(p, v); }
    __device__ __forceinline__ uint32_t atomic_add_ret(uint32_t* p, uint32_t v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ unsigned long long atomic_add_u64_ret(unsigned long long* p, unsigned long long v) const {
        return atomicAdd(p, v);
    }
    __device__ __forceinline__ unsigned long long load_relaxed(const unsigned long long* p) const { return ld_relaxed_gpu(p); }
    __device__ __forceinline__ unsigned long long cas_u64(unsigned long long* p, unsigned long long expect,
                                                          unsigned long long desired) const {
        return atomicCAS(p, expect, desired);
    }

    // block-wide exclusive scan of one uint32 per thread; every thread must call.  sm: >= 34 words.
#if defined(DPT_SCAN_ONE_BARRIER)
    // Tuning variant, measured and not adopted: ONE barrier per call - the warp totals go to one of two alternating 16-word
    // buffers and EVERY warp scans them for itself with shuffles.  Two barriers fewer per scan, ~25 more instructions per
    // warp: the kernels are issue-bound, not barrier-bound (k_scan_dedup 0.383 -> 0.400 ms).
    mutable uint32_t scan_phase = 0;
    __device__ __forceinline__ uint32_t exclusive_scan(uint32_t v, uint32_t* sm, uint32_t& total) const {
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        uint32_t* const buf = sm + scan_phase;
        scan_phase ^= 16u;
        if (lane == 31) buf[warp] = inc;
        __syncthreads();
        const uint32_t w = lane < nwarps ? buf[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {  // at most 16 warps per block
            const uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += o;
        }
        total = __shfl_sync(0xffffffffu, winc, 15);
        const uint32_t base = __shfl_sync(0xffffffffu, winc - w, warp);
        return base + inc - v;
    }
#else
    __device__ __forceinline__ uint32_t exclusive_scan(uint32_t v, uint32_t* sm, uint32_t& total) const {
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = lane < nwarps ? sm[lane] : 0u;
            uint32_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= (unsigned)d) winc += o;
            }
            sm[lane] = winc - w;
            if (lane == 31) sm[32] = winc;
        }
        __syncthreads();
        const uint32_t base = sm[warp];
        total = sm[32];
        __syncthreads();
        return base + inc - v;
    }
#endif

    __device__ __forceinline__ void reconverge() const { __syncwarp(); }
    __device__ __forceinline__ bool in_first_warp() const { return threadIdx.x < 32; }
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int warp_width() const { return 32; }
    __device__ __forceinline__ unsigned grid_warps() const { return gridDim.x * (blockDim.x >> 5); }
    // first i in [0, n] with a[i] >= x (a sorted), by the 32 lanes of one warp: every round probes 32 split points
    __device__ __forceinline__ int64_t warp_lower_bound(const int64_t* a, int64_t n, int64_t x) const {
        const int lane = (int)(threadIdx.x & 31);
        int64_t lo = 0, hi = n;
        while (hi > lo) {
            const int64_t step = (hi - lo + 32) / 33;
            const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
            const bool less = idx < hi && __ldg(a + idx) < x;
            const int k = __popc(__ballot_sync(0xffffffffu, less));  // a is sorted: the first k probes are < x
            const int64_t cut = lo + (int64_t)(k + 1) * step - 1;     // probe k: >= x, or beyond hi
            lo += (int64_t)k * step;
            if (k < 32 && cut < hi) hi = cut;
        }
        return lo;
    }
    // kernel B: lane 0 claims 32 consecutive work items for its warp; returns this lane's item index
    __device__ __forceinline__ unsigned long long warp_take(unsigned long long* cursor) const {
        unsigned long long base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(cursor, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        return base + (threadIdx.x & 31);
    }
    __device__ __forceinline__ bool warp_any(bool p) const { return __any_sync(0xffffffffu, p); }
    __device__ __forceinline__ int warp_count(bool p) const { return __popc(__ballot_sync(0xffffffffu, p)); }
    // kernel B: the lanes with `ask` take consecutive work items (one atomic per warp); other lanes get a don't-care
    __device__ __forceinline__ unsigned long long warp_take_n(unsigned long long* cursor, bool ask) const {
        const unsigned m = __ballot_sync(0xffffffffu, ask);
        const unsigned lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == 0 && m) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
    }

    // decoupled look-back, split in two so a tile can publish early and resolve late.  descriptor = status << 62 |
    // value (1 = this tile's aggregate, 2 = inclusive prefix).  publish: thread 0.  resolve: warp 0 sums the
    // predecessors' aggregates back to the nearest inclusive prefix, then publishes its own inclusive prefix;
    // *out = exclusive prefix of this tile (shared memory; visible after the next block sync).
    __device__ __forceinline__ void lookback_publish(unsigned long long* desc, int tile, unsigned long long agg) const {
        if (threadIdx.x == 0) st_relaxed_gpu(&desc[tile], ((tile == 0 ? 2ull : 1ull) << 62) | agg);
    }
    __device__ __forceinline__ void lookback_resolve(unsigned long long* desc, int tile, unsigned long long agg,
                                                     unsigned long long* out) const {
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        unsigned long long excl = 0;
        if (tile > 0) {
            // every lane inspects 4 consecutive predecessors per round (128 per warp round): a tile typically has a
            // few hundred predecessors in flight, and each round costs an L2 round trip
            int base = tile - 1;
            for (;;) {
                unsigned long long v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // four independent loads in flight
                    const int idx = base - (lane * 4 + j);
                    v[j] = idx >= 0 ? ld_relaxed_gpu(&desc[idx]) : (2ull << 62);  // in front of tile 0: inclusive prefix 0
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // older tiles publish earlier, so every one of these becomes non-zero
                    const int idx = base - (lane * 4 + j);
                    while ((v[j] >> 62) == 0) {  // (a waiting lane gives its issue slots to the warps that are working)
#if !defined(DPT_NO_LOOKBACK_SLEEP)
                        __nanosleep(40);
#endif
                        v[j] = ld_relaxed_gpu(&desc[idx]);
                    }
                }
                unsigned long long c = 0;
                bool found = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (!found) {
                        c += v[j] & PD_MASK;
                        found = (v[j] >> 62) == 2;
                    }
                }
                const unsigned incl = __ballot_sync(0xffffffffu, found);
                const int first = incl ? __ffs((int)incl) - 1 : 31;
                if (lane > first) c = 0;
#pragma unroll
                for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                excl += c;
                if (incl) break;
                base -= 128;
            }
            if (lane == 0) st_relaxed_gpu(&desc[tile], (2ull << 62) | (excl + agg));
        }
        if (lane == 0) *out = excl;
    }
};

// SPM_LLAMA rule: 8 KB regions, 30 KB of shared memory, <= 48 registers -> 5 CTAs (40 warps) per SM (PaGeom, dpt_pipe.h)
__global__ void __launch_bounds__(PA_THREADS, DPT_PA_CTAS) k_scan_dedup(const __grid_constant__ PipeParams P) {
    __shared__ ASmemT<true> S;
    DevBlk blk;
    pa_kernel<DevBlk, true>(blk, P, S);
}
// byte-level rules (GPT-2, Llama-3): the split scanner needs more registers and the sync-point list more memory
// (its shared memory - 37 KB at 3968-byte tiles - is DYNAMIC, so that tile sizes beyond the 48 KB static limit can be used)
__global__ void __launch_bounds__(PA_THREADS, DPT_PABL_CTAS) k_scan_dedup_bl(const __grid_constant__ PipeParams P) {
    extern __shared__ __align__(16) unsigned char pa_bl_smem[];
    ASmemT<false>& S = *reinterpret_cast<ASmemT<false>*>(pa_bl_smem);
    DevBlk blk;
    pa_kernel<DevBlk, false>(blk, P, S);
}

// the lock-step DP (dpt_dp_lock.cuh): words of at most 31 units (length classes 0..2) ...
__global__ void __launch_bounds__(PBL_THREADS, DPT_PBL_CTAS) k_dp_lock_spm(const __grid_constant__ PipeParams P) {
    __shared__ PblSmem<8> S;
    pbl_kernel<true, 8>(P, S);
}
__global__ void __launch_bounds__(PBL_THREADS, DPT_PBL_CTAS) k_dp_lock_bl(const __grid_constant__ PipeParams P) {
    __shared__ PblSmem<8> S;
    pbl_kernel<false, 8>(P, S);
}
// ... and of 32..63 units (class 3: a few thousand words): one warp per word (dpt_dp_warp.cuh).  The lock-step kernel's
// 64-row instantiation (pbl_kernel<., 16>) solved them in 0.127 ms - one serial chain per word - this one in ~0.02 ms; with
// ~45 k such words (German compounds of the sentence-pair corpus) the two are level (0.27 / 0.31 ms), so there is one path.
__global__ void __launch_bounds__(PBW_THREADS) k_dp_warp_spm(const __grid_constant__ PipeParams P) { pbw_kernel<true>(P); }
__global__ void __launch_bounds__(PBW_THREADS) k_dp_warp_bl(const __grid_constant__ PipeParams P) { pbw_kernel<false>(P); }

// thread-per-word DP with local-memory state: odd words, words longer than a warp, words the cooperative kernel deferred
__global__ void __launch_bounds__(PB_THREADS, DPT_PB_CTAS) k_dp_distinct(const __grid_constant__ PipeParams P) {
    DevBlk blk;
    pb_thread(blk, P);
}

__global__ void __launch_bounds__(PB_THREADS) k_dp_distinct_long(const __grid_constant__ PipeParams P) {
    DevBlk blk;
    pb_long_thread(blk, P, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(PC_THREADS, DPT_PC_CTAS) k_emit(const __grid_constant__ PipeParams P) {
    __shared__ CSmem S;
    DevBlk blk;
    pc_kernel(blk, P, S);
}

// Everything a launch sequence needs zeroed, in ONE launch (three cudaMemsetAsync calls before: each is a launch of its
// own with its own gap on the stream): up to three 16-byte-aligned regions, sizes in 16-byte units.
struct ClearJob {
    uint4* p[2];
    unsigned long long n16[2];
    uint8_t* bytes;  // a region of any alignment and size (the caller's doc_flags)
    unsigned long long n_bytes;
};
__global__ void __launch_bounds__(256) k_pipe_clear(const ClearJob J) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int r = 0; r < 2; ++r)
        for (unsigned long long i = t; i < J.n16[r]; i += stride) J.p[r][i] = z;
    for (unsigned long long i = t; i < J.n_bytes; i += stride) J.bytes[i] = 0;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// One side stream + fork/join events per host thread and device (created on first use, kept for the life of the thread).
struct SideStream {
    int device = -1;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
    bool ok() const { return stream != nullptr; }
};
static SideStream& side_stream() {
    static thread_local SideStream cache[16];
    static SideStream none;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return none;
    SideStream& s = cache[dev];
    if (s.device != dev) {
        const char* off = getenv("DPT_NO_SIDE_STREAM");
        s.device = dev;
        if (!(off && off[0] == '1')) {
            if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) s.stream = nullptr;
            if (s.stream && cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking) != cudaSuccess) {
                cudaStreamDestroy(s.stream);
                s.stream = nullptr;
            }
            if (s.stream && (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
                             cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess ||
                             cudaEventCreateWithFlags(&s.join2, cudaEventDisableTiming) != cudaSuccess)) {
                cudaStreamDestroy(s.stream);
                cudaStreamDestroy(s.stream2);
                s.stream = nullptr;
            }
        }
    }
    return s;
}

// Workspace = a TABLE part (word table, result records, id pool, DP queues: lives as long as the table, i.e. one
// call or the chunks of one chunked call) + a RANGE part (everything private to one launch sequence).
struct TableSizes {
    int64_t n_slots, pool_cap;
};
struct RangeSizes {
    int64_t n_tiles, n_ctiles, odd_cap, lp_cap;
};

static TableSizes table_sizes(int64_t n_bytes_total, int64_t word_cap_total, int worst) {
    TableSizes z;
    // worst: 0 = typical text (natural language repeats its words: ~1 distinct word per 150 bytes), 1 = every word may be
    // distinct, 2 = "roomy": a table for 1 distinct word per 10 bytes with the typical sizes for everything else - for
    // corpora that are mostly distinct words (identifiers, hashes): their first occurrences then go through the lock-step DP
    // kernel instead of overflowing into the per-occurrence odd-word path
    // (never more than two slots per word the caller makes room for: every word distinct is a load of 1/2, where 16 probes
    // practically always find a place - at one slot per word 6 % of the words of an all-distinct corpus became odd words)
    int64_t want = worst == 1 ? 2 * word_cap_total : worst == 2 ? n_bytes_total / 10 : n_bytes_total / 48;
    if (worst == 2 && want > 2 * word_cap_total) want = 2 * word_cap_total;
    if (want < 4096) want = 4096;
    int64_t s = 4096;
    while (s < want) s <<= 1;
    z.n_slots = s;
    z.pool_cap = worst == 1 ? 8 * n_bytes_total + 5 * word_cap_total + 64 : n_bytes_total / 4 + 65536;
    return z;
}
static RangeSizes range_sizes(int64_t range_bytes, int64_t word_cap, int worst) {
    RangeSizes z;
    z.n_tiles = (range_bytes + PA_T - 1) / PA_T + 1;  // +1: a range need not start on a tile boundary
    z.n_ctiles = (word_cap + PC_TILE - 1) / PC_TILE;
    z.odd_cap = worst == 1 ? word_cap + 16 : range_bytes / 64 + 4096;
    z.lp_cap = worst == 1 ? 6 * range_bytes + 8 * word_cap + 64 : range_bytes / 4 + 262144;
    return z;
}

int64_t corpus_table_workspace(int64_t n_bytes_total, int64_t word_cap_total, int32_t worst) {
    const TableSizes z = table_sizes(n_bytes_total, word_cap_total, worst);
    int64_t b = 0;
    b += align_up(sizeof(PipePersist), 256);
    b += align_up(z.n_slots * 8, 256);    // tags
    b += align_up(z.n_slots * (int64_t)sizeof(ResRec), 256);   // res
    b += align_up(z.pool_cap * 4, 256);   // pool
    return b + 1024;
}

int64_t corpus_range_workspace(int64_t range_bytes, int64_t range_docs, int64_t word_cap, int32_t worst) {
    const RangeSizes z = range_sizes(range_bytes, word_cap, worst);
    int64_t b = 0;
    b += align_up(sizeof(PipeCtl), 256);
    b += align_up(z.n_tiles * 8 + 8, 256) + align_up(z.n_ctiles * 8 + 8, 256);
    b += align_up(word_cap * 4 + 64, 256);                   // refs
    b += align_up(word_cap * 4 * PB_CLASSES + 64, 256);      // pending (length classes x min(word_cap, n_slots); upper bound)
    b += align_up((range_docs + 1) * 8, 256);                // doc_first_word
    b += align_up(z.odd_cap * 16, 256) + align_up(z.odd_cap * (int64_t)sizeof(ResRec), 256);  // odd, odd_res
    b += align_up((word_cap + z.odd_cap) * 4, 256);          // longq
    b += align_up(word_cap * 4 + 64, 256);                   // defer (upper bound of pend_stride entries)
    b += align_up(z.lp_cap, 256) + align_up(z.lp_cap * 8, 256) + 2 * align_up(z.lp_cap * 2, 256);
    return b + 4096;
}

int64_t encode_corpus_pipe_workspace(int64_t n_bytes, int64_t n_docs, int64_t word_cap, int32_t worst) {
    return corpus_table_workspace(n_bytes, word_cap, worst) + corpus_range_workspace(n_bytes, n_docs, word_cap, worst);
}

int encode_corpus_range(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes_total,
                        const int64_t* d_doc_offs, int64_t n_docs_total, int64_t byte_begin, int64_t byte_end,
                        int64_t doc_begin, int64_t doc_end, int32_t reset_table, int64_t table_bytes_total,
                        int64_t table_word_cap, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                        int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                        int64_t* d_n_out, void* d_table_ws, int64_t table_ws_bytes, void* d_ws, int64_t ws_bytes,
                        int32_t worst, int32_t phases, cudaStream_t st, std::string& err) {
    const int64_t range_bytes = byte_end - byte_begin, range_docs = doc_end - doc_begin;
    if (n_bytes_total <= 0 || n_docs_total <= 0 || word_cap <= 0 || range_bytes <= 0 || range_docs <= 0 || byte_begin < 0 ||
        byte_end > n_bytes_total || doc_begin < 0 || doc_end > n_docs_total || !d_text || !d_doc_offs || !d_doc_tok_offs ||
        !d_counters || !d_n_out || !d_ids || !d_word_lens || !d_word_flags) {
        err = "encode_corpus: bad argument";
        return DPT_EINVAL;
    }
    if (rule == DPT_RULE_SPM_LLAMA && (!v->byte_fallback || !v->marker_entry || v->unit_mode != DPT_UNIT_CODEPOINTS)) {
        err = "encode_corpus(SPM_LLAMA): vocabulary lacks U+2581 or the 256 <0xHH> byte tokens, or is not a code-point "
              "vocabulary; pre-split on the host";
        return DPT_EINVAL;
    }
    if (rule != DPT_RULE_SPM_LLAMA && v->unit_mode != DPT_UNIT_BYTES) {
        err = "encode_corpus(GPT2/LLAMA3): byte-level rules need a byte-unit vocabulary (DPT_UNIT_BYTES)";
        return DPT_EINVAL;
    }
    if (n_bytes_total >= (1ll << 37) || word_cap >= (1ll << 29) || range_docs >= (1ll << 29)) {
        err = "encode_corpus: batch too large (>= 128 GiB, >= 2^29 words or >= 2^29 documents); split it";
        return DPT_EINVAL;
    }
    if (!d_ws || !d_table_ws || ws_bytes < corpus_range_workspace(range_bytes, range_docs, word_cap, worst) ||
        table_ws_bytes < corpus_table_workspace(table_bytes_total, table_word_cap, worst)) {
        err = "encode_corpus: workspace too small (see dpt_encode_corpus_workspace / dpt_corpus_table_workspace)";
        return DPT_ECAPACITY;
    }
    const TableSizes tz = table_sizes(table_bytes_total, table_word_cap, worst);
    const RangeSizes z = range_sizes(range_bytes, word_cap, worst);
    if (tz.n_slots > (1ll << 29)) {
        err = "encode_corpus: word table too large; split the batch";
        return DPT_EINVAL;
    }
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    ClearJob clear{};
    PipeParams P{};
    P.V = v->d_view;
    P.text = d_text;
    P.n_bytes = n_bytes_total;
    P.doc_offs = d_doc_offs;
    P.n_docs = n_docs_total;
    P.byte_begin = byte_begin;
    P.byte_end = byte_end;
    P.doc_begin = doc_begin;
    P.n_docs_local = range_docs;
    P.ids = d_ids;
    P.ids_cap = ids_cap;
    P.word_lens = d_word_lens;
    P.word_flags = d_word_flags;
    P.word_cap = word_cap;
    P.doc_tok_offs = d_doc_tok_offs;
    P.doc_flags = d_doc_flags;
    P.counters = (unsigned long long*)d_counters;
    P.n_out = d_n_out;
    {   // table part: PipePersist + tags first, so a reset is one contiguous region of k_pipe_clear
        char* base = (char*)d_table_ws;
        int64_t used = 0;
        auto take = [&](int64_t bytes) {
            used = align_up(used, 256);
            char* p = base + used;
            used += bytes;
            return p;
        };
        P.persist = (PipePersist*)take(sizeof(PipePersist));
        P.tags = (unsigned long long*)take(tz.n_slots * 8);
        const int64_t zero_bytes = used;
        P.res = (ResRec*)take(tz.n_slots * (int64_t)sizeof(ResRec));
        P.pool = (int32_t*)take(tz.pool_cap * 4);
        P.pool_cap = tz.pool_cap;
        P.slot_mask = (uint32_t)(tz.n_slots - 1);
        if (reset_table) {
            clear.p[0] = (uint4*)base;
            clear.n16[0] = (unsigned long long)((zero_bytes + 15) / 16);  // (the region ends on a 256-byte boundary)
        }
    }
    const bool do_scan = (phases & 1) != 0, do_dp = (phases & 2) != 0, do_emit = (phases & 4) != 0;
    {   // range part: everything that must start zeroed is contiguous
        char* base = (char*)d_ws;
        int64_t used = 0;
        auto take = [&](int64_t bytes) {
            used = align_up(used, 256);
            char* p = base + used;
            used += bytes;
            return p;
        };
        P.ctl = (PipeCtl*)take(sizeof(PipeCtl));
        P.desc_w = (unsigned long long*)take(z.n_tiles * 8 + 8);
        P.desc_t = (unsigned long long*)take(z.n_ctiles * 8 + 8);
        const int64_t zero_bytes = used;
        P.refs = (uint32_t*)take(word_cap * 4 + 64);
        P.pend_stride = word_cap < tz.n_slots ? word_cap : tz.n_slots;  // a pending word owns a table slot
        P.pending = (uint32_t*)take(P.pend_stride * 4 * PB_CLASSES + 64);
        P.doc_first_word = (int64_t*)take((range_docs + 1) * 8);
        P.odd = (OddWord*)take(z.odd_cap * 16);
        P.odd_res = (ResRec*)take(z.odd_cap * (int64_t)sizeof(ResRec));
        P.longq = (uint32_t*)take((word_cap + z.odd_cap) * 4);
        P.defer = (uint32_t*)take(P.pend_stride * 4 + 64);
        P.coop = 1;
        P.lp_norm = (uint8_t*)take(z.lp_cap);
        P.lp_best = (uint64_t*)take(z.lp_cap * 8);
        P.lp_a = (uint16_t*)take(z.lp_cap * 2);
        P.lp_b = (uint16_t*)take(z.lp_cap * 2);
        P.odd_cap = z.odd_cap;
        P.lp_cap = z.lp_cap;
        if (do_scan) {
            clear.p[1] = (uint4*)base;
            clear.n16[1] = (unsigned long long)((zero_bytes + 15) / 16);  // (up to 8 bytes of the padding in front of refs)
        }
    }
    const int64_t pa_t = rule == DPT_RULE_SPM_LLAMA ? PaGeom<true>::T : PaGeom<false>::T;  // (<= the sizing's PA_T tiles either way)
    P.tile_first = (int32_t)(byte_begin / pa_t);
    P.n_tiles = (int32_t)((byte_end + pa_t - 1) / pa_t - byte_begin / pa_t);
    P.n_ctiles = (int32_t)z.n_ctiles;
    P.spm = rule == DPT_RULE_SPM_LLAMA ? 1 : 0;
    P.rule = rule;
    P.vec_ok = ((((uintptr_t)d_word_lens) & 15u) == 0 && (((uintptr_t)d_word_flags) & 7u) == 0) ? 1 : 0;
    if (do_scan && d_doc_flags) {
        clear.bytes = d_doc_flags;
        clear.n_bytes = (unsigned long long)range_docs;
    }
    if (clear.n16[0] || clear.n16[1] || clear.n_bytes) {
        if (((uintptr_t)clear.p[0] | (uintptr_t)clear.p[1]) & 15u) {  // a workspace that is not 16-byte aligned
            if (clear.n16[0]) cudaMemsetAsync(clear.p[0], 0, (size_t)clear.n16[0] * 16, st);
            if (clear.n16[1]) cudaMemsetAsync(clear.p[1], 0, (size_t)clear.n16[1] * 16, st);
            if (clear.n_bytes) cudaMemsetAsync(clear.bytes, 0, (size_t)clear.n_bytes, st);
        } else {
            ProfScope prof("k_pipe_clear", st);
            k_pipe_clear<<<(unsigned)(sm_count * 8), 256, 0, st>>>(clear);
            ++g_launches;
        }
    }
    if (do_scan) {
    {
        ProfScope prof(P.spm ? "k_scan_dedup" : "k_scan_dedup_bl", st);
        if (P.spm) {
            k_scan_dedup<<<(unsigned)P.n_tiles, PA_THREADS, 0, st>>>(P);
        } else {
            static bool bl_attr = false;  // (per process; the attribute is per function and device - set again after a device change is harmless)
            static int bl_dev = -1;
            int dev = 0;
            cudaGetDevice(&dev);
            if (!bl_attr || bl_dev != dev) {
                cudaFuncSetAttribute(k_scan_dedup_bl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ASmemT<false>));
                bl_attr = true;
                bl_dev = dev;
            }
            k_scan_dedup_bl<<<(unsigned)P.n_tiles, PA_THREADS, sizeof(ASmemT<false>), st>>>(P);
        }
        ++g_launches;
    }
    }
    if (do_dp) {
        // The few long words (32..63 units: one warp per word; odd words and words of more than 63 units: the
        // thread-per-word kernel, a serial chain per word) run BESIDE the lock-step kernel on a side stream of the library
        // (fork / join with events: from the caller's point of view everything is ordered on `st`).
        // (TWO side streams: on one, the thread-per-word kernel waited behind the warp kernel - each stretched to ~0.2 ms by
        // sharing the SMs with the lock-step kernel - and their sum, not the lock-step kernel, was the critical path.)
        SideStream& side = side_stream();
        const bool forked = side.ok();
        cudaStream_t s1 = forked ? side.stream : st;
        cudaStream_t s2 = forked ? side.stream2 : st;
        if (forked) {
            cudaEventRecord(side.fork, st);
            cudaStreamWaitEvent(s1, side.fork, 0);
            cudaStreamWaitEvent(s2, side.fork, 0);
        }
        {
            ProfScope prof(P.spm ? "k_dp_warp_spm" : "k_dp_warp_bl", s1);
            if (P.spm)
                k_dp_warp_spm<<<(unsigned)(sm_count * 4), PBW_THREADS, 0, s1>>>(P);
            else
                k_dp_warp_bl<<<(unsigned)(sm_count * 4), PBW_THREADS, 0, s1>>>(P);
            ++g_launches;
        }
        if (forked) cudaEventRecord(side.join, s1);
        {
            ProfScope prof("k_dp_distinct", s2);
            PipeParams P1 = P;
            P1.coop = 1;
            k_dp_distinct<<<(unsigned)(sm_count * 2), PB_THREADS, 0, s2>>>(P1);
            ++g_launches;
        }
        if (forked) cudaEventRecord(side.join2, s2);
        {
            ProfScope prof(P.spm ? "k_dp_lock_spm" : "k_dp_lock_bl", st);
            static int bl_ctas = 0;  // CTAs per SM of the lock-step kernel's grid (development knob: DPT_BL_GRID)
            if (!bl_ctas) {
                const char* e = getenv("DPT_BL_GRID");
                bl_ctas = e && atoi(e) > 0 ? atoi(e) : DPT_PBL_CTAS;
            }
            if (P.spm)
                k_dp_lock_spm<<<(unsigned)(sm_count * bl_ctas), PBL_THREADS, 0, st>>>(P);
            else
                k_dp_lock_bl<<<(unsigned)(sm_count * bl_ctas), PBL_THREADS, 0, st>>>(P);
            ++g_launches;
        }
        if (forked) {  // (the side streams' kernels defer words too)
            cudaStreamWaitEvent(st, side.join, 0);
            cudaStreamWaitEvent(st, side.join2, 0);
        }
        if (P.spm) {  // the words the lock-step kernels deferred (byte-level rules never defer)
            ProfScope prof("k_dp_distinct_deferred", st);
            PipeParams P2 = P;
            P2.coop = 2;
            k_dp_distinct<<<(unsigned)(sm_count * 2), PB_THREADS, 0, st>>>(P2);
            ++g_launches;
        }
        {
            ProfScope prof("k_dp_distinct_long", st);
            k_dp_distinct_long<<<(unsigned)(sm_count * 2), PB_THREADS, 0, st>>>(P);
            ++g_launches;
        }
    }
    if (do_emit) {
    {
        ProfScope prof("k_emit", st);
        k_emit<<<(unsigned)z.n_ctiles, PC_THREADS, 0, st>>>(P);
        ++g_launches;
    }
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("encode_corpus: ") + cudaGetErrorString(e);
        return DPT_ECUDA;
    }
    return DPT_OK;
}

int encode_corpus_pipe(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs,
                       int64_t n_docs, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                       int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                       int64_t* d_n_out, void* d_ws, int64_t ws_bytes, int32_t worst, cudaStream_t st, std::string& err) {
    if (n_bytes <= 0 || n_docs <= 0 || word_cap <= 0 || !d_ws) {
        err = "encode_corpus: bad argument";
        return DPT_EINVAL;
    }
    const int64_t tb = corpus_table_workspace(n_bytes, word_cap, worst);
    if (ws_bytes < tb + corpus_range_workspace(n_bytes, n_docs, word_cap, worst)) {
        err = "encode_corpus: workspace too small (see dpt_encode_corpus_workspace)";
        return DPT_ECAPACITY;
    }
    const int64_t tb_al = align_up(tb, 256);
    return encode_corpus_range(v, rule, d_text, n_bytes, d_doc_offs, n_docs, 0, n_bytes, 0, n_docs, 1, n_bytes, word_cap, d_ids,
                               ids_cap, d_word_lens, d_word_flags, word_cap, d_doc_tok_offs, d_doc_flags, d_counters, d_n_out,
                               d_ws, tb_al, (char*)d_ws + tb_al, ws_bytes - tb_al, worst, 7, st, err);
}

}  // namespace dpt
