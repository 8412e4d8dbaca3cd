// The fused tile kernel (sm_100a): one persistent CTA per SM runs dpt_tile.h's pipeline over 4 KB tiles of raw
// corpus bytes handed out in order by an atomic ticket; the double-array trie's hot slots live in 128 KB of
// shared memory; token and word offsets come from a decoupled look-back across tiles, so the corpus is read
// once and every output is written once, by one launch.
#include <cuda_runtime.h>

#include <string>

#include "../../include/dptok.h"
#include "dpt_tile.h"
#include "kernels.h"
#include "vocab.h"

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
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) const { atomicOr(p, v); }
    __device__ __forceinline__ void atomic_add(uint32_t* p, uint32_t v) const { atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_add_u64(unsigned long long* p, unsigned long long v) const { atomicAdd(p, v); }
    __device__ __forceinline__ unsigned take_ticket(unsigned int* p) const { return atomicAdd(p, 1u); }

    // block-wide exclusive scan of one uint32 per thread; every thread must call.  sm: >= 34 words.
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

    // decoupled look-back: warp 0 resolves the word offset, warp 1 the token offset of this tile.
    // descriptor = status << 62 | value; status 1 = this tile's aggregate, 2 = inclusive prefix.
    __device__ __forceinline__ void lookback(const TileParams& P, TileSmem& S, int tile) const {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp >= 2) return;
        unsigned long long* desc = warp == 0 ? P.desc_w : P.desc_t;
        const unsigned long long agg = warp == 0 ? (unsigned long long)(S.tile_tot >> 16) : (unsigned long long)(S.tile_tot & 0xFFFFu);
        unsigned long long excl = 0;
        if (tile > 0) {
            if (lane == 0) st_relaxed_gpu(&desc[tile], (1ull << 62) | agg);
            int base = tile - 1;
            for (;;) {
                const int idx = base - lane;
                unsigned long long v = 2ull << 62;  // in front of tile 0: inclusive prefix 0
                if (idx >= 0) {
                    do {
                        v = ld_relaxed_gpu(&desc[idx]);
                    } while ((v >> 62) == 0);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                const int first = incl ? __ffs((int)incl) - 1 : 31;
                unsigned long long c = lane <= first ? (v & TL_DESC_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                excl += c;
                if (incl) break;
                base -= 32;
            }
        }
        if (lane == 0) {
            st_relaxed_gpu(&desc[tile], (2ull << 62) | (excl + agg));
            if (warp == 0) S.base_w = excl; else S.base_t = excl;
        }
    }
};

__global__ void __launch_bounds__(TL_THREADS, 1) k_fused_tiles(const __grid_constant__ TileParams P) {
    extern __shared__ __align__(16) unsigned char tl_smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(tl_smem_raw);
    DevBlk blk;
    tl_loop(blk, P, S, (int64_t)blockIdx.x);
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static constexpr int FUSED_MAX_CTAS = 160;

static int64_t fused_ctas(int64_t n_tiles) { return n_tiles < FUSED_MAX_CTAS ? (n_tiles > 0 ? n_tiles : 1) : FUSED_MAX_CTAS; }

int64_t encode_corpus_fused_workspace(int64_t n_bytes) {
    const int64_t n_tiles = (n_bytes + TL_T - 1) / TL_T;
    return 2 * align_up(n_tiles * 8 + 8, 256) + 256 + fused_ctas(n_tiles) * align_up((int64_t)TL_ARENA_BYTES, 256) + 4096;
}

int encode_corpus_fused(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs,
                        int64_t n_docs, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                        int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                        int64_t* d_n_out, void* d_ws, int64_t ws_bytes, cudaStream_t st, std::string& err) {
    if (n_bytes <= 0 || n_docs <= 0 || !d_text || !d_doc_offs || !d_doc_tok_offs || !d_counters || !d_n_out || !d_ids ||
        !d_word_lens || !d_word_flags) {
        err = "encode_corpus: bad argument";
        return DPT_EINVAL;
    }
    if (rule == DPT_RULE_SPM_LLAMA && (!v->byte_fallback || !v->marker_entry || v->unit_mode != DPT_UNIT_CODEPOINTS)) {
        err = "encode_corpus(SPM_LLAMA): vocabulary lacks U+2581 or the 256 <0xHH> byte tokens, or is not a code-point "
              "vocabulary; pre-split on the host";
        return DPT_EINVAL;
    }
    if (n_bytes >= (1ll << 40)) {
        err = "encode_corpus: batch too large; split it";
        return DPT_EINVAL;
    }
    const int64_t n_tiles = (n_bytes + TL_T - 1) / TL_T;
    if (n_tiles >= (1ll << 31)) {
        err = "encode_corpus: batch too large; split it";
        return DPT_EINVAL;
    }
    if (!d_ws || ws_bytes < encode_corpus_fused_workspace(n_bytes)) {
        err = "encode_corpus: workspace too small (see dpt_encode_corpus_workspace)";
        return DPT_ECAPACITY;
    }
    static int sm_count = 0;
    static bool attr_set = false;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(k_fused_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        if (e != cudaSuccess) {
            err = std::string("encode_corpus: cudaFuncSetAttribute: ") + cudaGetErrorString(e);
            return DPT_ECUDA;
        }
        attr_set = true;
    }
    int64_t ctas = fused_ctas(n_tiles);
    if (ctas > sm_count) ctas = sm_count;

    char* base = (char*)d_ws;
    int64_t used = 0;
    auto take = [&](int64_t bytes) {
        used = align_up(used, 256);
        char* p = base + used;
        used += bytes;
        return p;
    };
    TileParams P{};
    P.V = v->d_view;
    P.text = d_text;
    P.n_bytes = n_bytes;
    P.doc_offs = d_doc_offs;
    P.n_docs = n_docs;
    P.ids = d_ids;
    P.ids_cap = ids_cap;
    P.word_lens = d_word_lens;
    P.word_flags = d_word_flags;
    P.word_cap = word_cap;
    P.doc_tok_offs = d_doc_tok_offs;
    P.doc_flags = d_doc_flags;
    P.counters = (unsigned long long*)d_counters;
    P.n_out = d_n_out;
    // descriptors + ticket are contiguous: one memset
    char* zero0 = take(0);
    P.desc_w = (unsigned long long*)take(n_tiles * 8 + 8);
    P.desc_t = (unsigned long long*)take(n_tiles * 8 + 8);
    P.ticket = (unsigned int*)take(256);
    const int64_t zero_bytes = (base + used) - zero0;
    P.arena_norm = (uint8_t*)take(ctas * (int64_t)TL_ARENA_POS);
    P.arena_best = (uint64_t*)take(ctas * (int64_t)TL_ARENA_POS * 8);
    P.arena_a = (uint16_t*)take(ctas * (int64_t)TL_ARENA_POS * 2);
    P.arena_b = (uint16_t*)take(ctas * (int64_t)TL_ARENA_POS * 2);
    P.n_tiles = (int32_t)n_tiles;
    P.kc = (int32_t)(v->da.size() < (size_t)TL_KC ? v->da.size() : (size_t)TL_KC);
    P.spm = rule == DPT_RULE_SPM_LLAMA ? 1 : 0;
    P.rule = rule;

    cudaMemsetAsync(zero0, 0, (size_t)zero_bytes, st);
    cudaMemsetAsync(d_counters, 0, 4 * sizeof(int64_t), st);
    cudaMemsetAsync(d_n_out, 0, 8 * sizeof(int64_t), st);
    if (d_doc_flags) cudaMemsetAsync(d_doc_flags, 0, (size_t)n_docs, st);
    {
        ProfScope prof("k_fused_tiles", st);
        k_fused_tiles<<<(unsigned)ctas, TL_THREADS, sizeof(TileSmem), st>>>(P);
        ++g_launches;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("encode_corpus: ") + cudaGetErrorString(e);
        return DPT_ECUDA;
    }
    return DPT_OK;
}

}  // namespace dpt
