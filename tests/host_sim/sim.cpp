// TEST-ONLY host build of the device headers (dpt_common.h / dpt_dp_core.h / dpt_rules.h).
// Lets the CPU test-suite exercise the exact per-word code the CUDA kernels run, against the oracle,
// before any GPU time is spent.  Never linked into the product library; nothing under
// dp-tokenization_b200/ loads it.
#include <barrier>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "dpt_decode.h"
#include "dpt_dp_core.h"
#include "dpt_rules.h"
#include "dpt_pipe.h"
#include "vocab.h"

// ---- std::thread emulation of one CUDA block for dpt_pipe.h (the pipeline kernels' source, verbatim) --------
namespace {
struct HostShared {
    std::barrier<> bar;
    std::vector<uint32_t> vals;
    explicit HostShared(int n) : bar(n), vals(n) {}
};
struct HostBlk {
    int tid_, nt_;
    HostShared* sh;
    int tid() const { return tid_; }
    int nthreads() const { return nt_; }
    bool persistent() const { return true; }  // ONE emulated CTA walks all tiles in ticket order
    int block_index() const { return 0; }  // unused: the emulated CTA is persistent
    void sync() const { if (sh) sh->bar.arrive_and_wait(); }
    void atomic_or(uint32_t* p, uint32_t v) const { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
    void atomic_add(uint32_t* p, uint32_t v) const { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
    uint32_t atomic_add_ret(uint32_t* p, uint32_t v) const { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
    unsigned long long atomic_add_u64_ret(unsigned long long* p, unsigned long long v) const {
        return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    }
    unsigned long long load_relaxed(const unsigned long long* p) const { return __atomic_load_n(p, __ATOMIC_RELAXED); }
    unsigned long long cas_u64(unsigned long long* p, unsigned long long expect, unsigned long long desired) const {
        __atomic_compare_exchange_n(p, &expect, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
        return expect;  // old value, like atomicCAS
    }
    uint32_t exclusive_scan(uint32_t v, uint32_t*, uint32_t& total) const {
        sh->vals[tid_] = v;
        sync();
        uint32_t ex = 0, tot = 0;
        for (int k = 0; k < nt_; ++k) {
            if (k < tid_) ex += sh->vals[k];
            tot += sh->vals[k];
        }
        total = tot;
        sync();
        return ex;
    }
    void reconverge() const {}
    bool in_first_warp() const { return tid_ == 0; }  // an emulated warp is one thread
    int lane() const { return 0; }
    int warp_width() const { return 1; }
    unsigned grid_warps() const { return 0; }
    int64_t warp_lower_bound(const int64_t* a, int64_t n, int64_t x) const { return dpt::pp_lower_bound(a, n, x); }
    unsigned long long warp_take(unsigned long long* cursor) const { return __atomic_fetch_add(cursor, 1ull, __ATOMIC_RELAXED); }
    bool warp_any(bool p) const { return p; }
    int warp_count(bool p) const { return p ? 1 : 0; }
    unsigned long long warp_take_n(unsigned long long* cursor, bool ask) const {
        return ask ? __atomic_fetch_add(cursor, 1ull, __ATOMIC_RELAXED) : ~0ull;
    }
    // tiles are processed in order by the one emulated CTA: the predecessor's inclusive prefix is always there
    void lookback_publish(unsigned long long* desc, int tile, unsigned long long agg) const {
        if (tid_ == 0) desc[tile] = ((tile == 0 ? 2ull : 1ull) << 62) | agg;
    }
    void lookback_resolve(unsigned long long* desc, int tile, unsigned long long agg, unsigned long long* out) const {
        if (tid_ != 0) return;
        const unsigned long long prev = tile ? (desc[tile - 1] & dpt::PD_MASK) : 0ull;
        desc[tile] = (2ull << 62) | (prev + agg);
        *out = prev;
    }
};
}  // namespace

extern "C" {

void* sim_vocab_create(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n, int32_t unit_mode) {
    dpt_vocab* v = nullptr;
    std::string err;
    if (dpt_vocab_build(bytes, offs, ids, n, unit_mode, &v, err)) return nullptr;
    return v;
}
void sim_vocab_destroy(void* v) { delete (dpt_vocab*)v; }
int32_t sim_vocab_set_merges(void* vv, const int32_t* left, const int32_t* right, const int32_t* merged, int32_t n) {
    return ((dpt_vocab*)vv)->set_merges(left, right, merged, n);
}

int32_t sim_lookup(void* vv, const uint8_t* s, int32_t n) {
    dpt_vocab* v = (dpt_vocab*)vv;
    DptHashState h = dpt_hash_init(v->ph_salt);
    for (int i = 0; i < n; ++i) dpt_hash_byte(h, s[i]);
    return dpt_ph_lookup(v->h_view, h);
}

// returns number of ids (or -1 if untokenizable); *word_len = len_dp[n]
int32_t sim_word(void* vv, const uint8_t* s, int32_t n, const uint8_t* unit_starts, int32_t* out_ids, int32_t cap,
                 int32_t* word_len) {
    dpt_vocab* v = (dpt_vocab*)vv;
    std::vector<uint64_t> best(n + 1);
    std::vector<uint16_t> A(n + 1), B(n + 1);
    dpt_forward<true>(v->h_view, s, n, unit_starts, best.data(), A.data(), B.data());
    *word_len = (int32_t)dpt_key_len(best[n]);
    if (!dpt_backward_emit(v->h_view, s, n, best.data(), A.data(), B.data(), out_ids, cap)) return -1;
    return *word_len;
}

void sim_info(void* vv, int32_t* out) {
    dpt_vocab* v = (dpt_vocab*)vv;
    out[0] = v->n_tokens;
    out[1] = v->n_nodes;
    out[2] = (int32_t)v->da.size();
    out[3] = (int32_t)v->lmax;
    out[4] = (int32_t)v->ph_seed.size();
    out[5] = (int32_t)v->ph_id.size();
    out[6] = v->marker_leading_only;
    out[7] = v->byte_fallback;
}
// Sequential walk over the raw text with exactly the per-character functions k_spm_count/k_spm_write use.
// Returns normalised byte count; *n_words_out = number of words; word_offs gets n_words+1 entries.
int64_t sim_spm_normalise(void* vv, const uint8_t* text, int64_t n, const int64_t* doc_offs, int64_t n_docs, uint8_t* out,
                          int64_t out_cap, int64_t* word_offs, int64_t word_cap, int64_t* n_words_out, uint8_t* doc_flags) {
    dpt_vocab* v = (dpt_vocab*)vv;
    std::vector<uint32_t> bits(n / 32 + 2, 0u);
    for (int64_t d = 0; d < n_docs; ++d) bits[doc_offs[d] >> 5] |= 1u << (doc_offs[d] & 31);
    int64_t ob = 0, ow = 0, od = 0;
    for (int64_t p = 0; p < n; ++p) {
        if (!dpt_spm_is_char_start(text, bits.data(), p)) continue;
        const DptSpmChar ch = dpt_spm_classify(v->h_view, text, n, bits.data(), p);
        if (dpt_bit_test(bits.data(), p)) {
            if (ow < word_cap) word_offs[ow] = ob;
            if (ow + 1 < word_cap) word_offs[ow + 1] = ob + 3;
            if (ob + 6 <= out_cap) {
                const uint8_t pre[6] = {'<', 's', '>', DPT_MARK0, DPT_MARK1, DPT_MARK2};
                memcpy(out + ob, pre, 6);
            }
            if (ch.marker && doc_flags) doc_flags[od] = 1;
            ob += 6;
            ow += 2;
            od += 1;
        } else if (ch.marker) {
            if (!dpt_spm_prev_is_marker(text, bits.data(), p)) {
                if (ow < word_cap) word_offs[ow] = ob;
                ow += 1;
            } else if (doc_flags) {
                doc_flags[od - 1] = 1;
            }
        }
        if (ob + ch.out_len <= out_cap) dpt_spm_write_char(text, p, ch, out + ob);
        ob += ch.out_len;
    }
    if (ow < word_cap + 1) word_offs[ow] = ob;
    *n_words_out = ow;
    return ob;
}

// The corpus pipeline's code (dpt_pipe.h): kernel A on `nthreads` host threads emulating one CTA, kernel B as a
// plain loop over its threads, kernel C on PC_THREADS host threads.  n_slots (power of two, 0 = default) and
// odd_cap/pool_cap/lp_cap (0 = default) shrink the tables to exercise probe failure and capacity reporting.
// rule: 1 SPM_LLAMA, 2 GPT2, 3 LLAMA3 (include/dptok.h).  n_ranges > 1: the corpus is processed as that many
// consecutive document ranges that share ONE word table (dpt_encode_corpus_range semantics); outputs are stitched.
int32_t sim_encode_corpus_pipe(void* vv, int32_t rule, const uint8_t* text, int64_t n_bytes, const int64_t* doc_offs,
                               int64_t n_docs, int32_t* ids, int64_t ids_cap, int32_t* word_lens, uint8_t* word_flags,
                               int64_t word_cap, int64_t* doc_tok_offs, uint8_t* doc_flags, int64_t* counters,
                               int64_t* n_out, int32_t nthreads, int64_t n_slots, int64_t odd_cap, int64_t pool_cap,
                               int64_t lp_cap, int32_t n_ranges) {
    using namespace dpt;
    dpt_vocab* v = (dpt_vocab*)vv;
    if (n_slots <= 0) {
        n_slots = 4096;
        while (n_slots < n_bytes / 48) n_slots <<= 1;
    }
    if (odd_cap <= 0) odd_cap = word_cap + 16;
    if (pool_cap <= 0) pool_cap = 6 * n_bytes + 3 * word_cap + 64;
    if (lp_cap <= 0) lp_cap = 6 * n_bytes + 8 * word_cap + 64;
    if (n_ranges < 1) n_ranges = 1;
    if (n_ranges > n_docs) n_ranges = (int32_t)n_docs;
    // table part: shared by all ranges
    PipePersist persist{};
    std::vector<unsigned long long> tags(n_slots, 0);
    std::vector<ResRec> res(n_slots);
    std::vector<int32_t> pool(pool_cap);
    memset(counters, 0, 32);
    memset(n_out, 0, 64);
    int64_t ids_base = 0, words_base = 0;
    for (int32_t rg = 0; rg < n_ranges; ++rg) {
        const int64_t d0 = n_docs * rg / n_ranges, d1 = n_docs * (rg + 1) / n_ranges;
        const int64_t b0 = doc_offs[d0], b1 = doc_offs[d1], nd = d1 - d0;
        PipeParams P{};
        P.V = v->h_view;
        P.text = text;
        P.n_bytes = n_bytes;
        P.doc_offs = doc_offs;
        P.n_docs = n_docs;
        P.byte_begin = b0;
        P.byte_end = b1;
        P.doc_begin = d0;
        P.n_docs_local = nd;
        P.persist = &persist;
        // range-local outputs
        const int64_t r_ids_cap = ids_cap > ids_base ? ids_cap - ids_base : 0, r_word_cap = word_cap - words_base;
        std::vector<int32_t> r_ids(r_ids_cap + 1), r_lens(r_word_cap + 8);
        std::vector<uint8_t> r_flags(r_word_cap + 8), r_dflags(nd + 1);
        std::vector<int64_t> r_dto(nd + 1, -7);
        int64_t r_ctr[4] = {0, 0, 0, 0}, r_nout[8] = {0};
        P.ids = r_ids.data();
        P.ids_cap = r_ids_cap;
        P.word_lens = r_lens.data();
        P.word_flags = r_flags.data();
        P.word_cap = r_word_cap;
        P.doc_tok_offs = r_dto.data();
        P.doc_flags = doc_flags ? r_dflags.data() : nullptr;
        P.counters = (unsigned long long*)r_ctr;
        P.n_out = r_nout;
        const int64_t pa_t = rule == 1 ? PaGeom<true>::T : PaGeom<false>::T;  // this rule's tile (dpt_pipe.h: PaGeom)
        const int64_t n_tiles = (b1 + pa_t - 1) / pa_t - b0 / pa_t, n_ctiles = (r_word_cap + PC_TILE - 1) / PC_TILE;
        PipeCtl ctl{};
        std::vector<unsigned long long> dw(n_tiles + 1, 0), dt(n_ctiles + 1, 0);
        std::vector<ResRec> odd_res(odd_cap);
        std::vector<uint32_t> refs(r_word_cap + 16), longq(r_word_cap + odd_cap + 16), pending(PB_CLASSES * r_word_cap + 16);
        P.pend_stride = r_word_cap;
        std::vector<int64_t> dfw(nd + 1, -1);
        std::vector<OddWord> odd(odd_cap);
        std::vector<uint8_t> lpn(lp_cap);
        std::vector<uint64_t> lpb(lp_cap);
        std::vector<uint16_t> lpa(lp_cap), lpbb(lp_cap);
        P.ctl = &ctl;
        P.desc_w = dw.data();
        P.desc_t = dt.data();
        P.tags = tags.data();
        P.res = res.data();
        P.pending = pending.data();
        P.refs = refs.data();
        P.doc_first_word = dfw.data();
        P.odd = odd.data();
        P.odd_res = odd_res.data();
        P.pool = pool.data();
        P.longq = longq.data();
        P.lp_norm = lpn.data();
        P.lp_best = lpb.data();
        P.lp_a = lpa.data();
        P.lp_b = lpbb.data();
        P.odd_cap = odd_cap;
        P.pool_cap = pool_cap;
        P.lp_cap = lp_cap;
        P.slot_mask = (uint32_t)(n_slots - 1);
        P.tile_first = (int32_t)(b0 / pa_t);
        P.n_tiles = (int32_t)n_tiles;
        P.n_ctiles = (int32_t)n_ctiles;
        P.spm = rule == 1 ? 1 : 0;  // DPT_RULE_SPM_LLAMA
        P.rule = rule;
        P.vec_ok = ((((uintptr_t)P.word_lens) & 15u) == 0 && (((uintptr_t)P.word_flags) & 7u) == 0) ? 1 : 0;
        {   // kernel A
            auto run = [&](auto* S) {
                constexpr bool kSpm = std::is_same_v<std::remove_pointer_t<decltype(S)>, ASmemT<true>>;
                HostShared sh(nthreads);
                std::vector<std::thread> th;
                for (int t = 0; t < nthreads; ++t)
                    th.emplace_back([&, t] {
                        HostBlk blk{t, nthreads, &sh};
                        pa_kernel<HostBlk, kSpm>(blk, P, *S);
                    });
                for (auto& x : th) x.join();
            };
            if (P.spm) {
                auto S = std::make_unique<ASmemT<true>>();
                run(S.get());
            } else {
                auto S = std::make_unique<ASmemT<false>>();
                run(S.get());
            }
        }
        {   // kernels B (no block-level cooperation: run the threads one after the other)
            HostBlk blk{0, 1, nullptr};
            const int64_t g = 37;
            pb_thread(blk, P);
            for (int64_t t = 0; t < g; ++t) pb_long_thread(blk, P, t, g);
        }
        {   // kernel C
            auto S = std::make_unique<CSmem>();
            HostShared sh(PC_THREADS);
            std::vector<std::thread> th;
            for (int t = 0; t < PC_THREADS; ++t)
                th.emplace_back([&, t] {
                    HostBlk blk{t, PC_THREADS, &sh};
                    pc_kernel(blk, P, *S);
                });
            for (auto& x : th) x.join();
        }
        // (the counters and the capacity report were written by the thread that finished kernel C's last tile)
        // stitch
        const int64_t nid = r_nout[0] < r_ids_cap ? r_nout[0] : r_ids_cap, nw = r_nout[1] < r_word_cap ? r_nout[1] : r_word_cap;
        for (int64_t k = 0; k < nid; ++k) ids[ids_base + k] = r_ids[k];
        for (int64_t k = 0; k < nw; ++k) {
            word_lens[words_base + k] = r_lens[k];
            word_flags[words_base + k] = r_flags[k];
        }
        for (int64_t d = 0; d < nd; ++d) {
            doc_tok_offs[d0 + d] = r_dto[d] + ids_base;
            if (doc_flags) doc_flags[d0 + d] = r_dflags[d];
        }
        for (int k = 0; k < 4; ++k) counters[k] += r_ctr[k];
        n_out[0] += r_nout[0];
        n_out[1] += r_nout[1];
        for (int k = 2; k < 8; k += 2)
            if (rg == 0 || r_nout[k] - r_nout[k + 1] > n_out[k] - n_out[k + 1]) {
                n_out[k] = r_nout[k];
                n_out[k + 1] = r_nout[k + 1];
            }
        ids_base += r_nout[0];
        words_base += r_nout[1];
    }
    doc_tok_offs[n_docs] = ids_base;
    return 0;
}
// Round-trip check of k_roundtrip (kernels.cu).  mode 1: the kernel's loop with its 32 lanes run one after the other
// (length of each lane's token -> inclusive scan -> compare at the scanned place -> vote), on the functions of
// dpt_decode.h the kernel calls.  mode 0: an independent serial restatement - decode the ids of the document into a
// buffer the way tokenizer.decode does (tokenizer_utils.py:82-84, :176-179), then compare the buffer with the text.
void sim_roundtrip(void* vv, const int32_t* ids, const int64_t* doc_tok_offs, const uint8_t* text, const int64_t* doc_offs,
                   int64_t n_docs, int32_t skip_bos, uint8_t* ok, int32_t mode) {
    const DptVocabView& V = ((dpt_vocab*)vv)->h_view;
    const bool spm = V.unit_mode == 1;
    for (int64_t d = 0; d < n_docs; ++d) {
        const int64_t t0 = doc_tok_offs[d] + (skip_bos ? 1 : 0), t1 = doc_tok_offs[d + 1];
        const int64_t pe = doc_offs[d + 1];
        if (mode == 0) {
            std::string out;
            bool valid = true;
            for (int64_t t = t0; t < t1 && valid; ++t) {
                int64_t a, b;
                if (!dpt_tok_span(V, ids[t], a, b)) {
                    valid = false;
                    break;
                }
                std::string tokstr((const char*)V.tok_bytes + a, (size_t)(b - a));
                unsigned byte = 0;
                char tail = 0;
                if (spm && tokstr.size() == 6 && sscanf(tokstr.c_str(), "<0x%2X%c", &byte, &tail) == 2 && tail == '>' &&
                    !(tokstr[3] >= 'a' && tokstr[3] <= 'f') && !(tokstr[4] >= 'a' && tokstr[4] <= 'f')) {
                    out.push_back((char)byte);
                    continue;
                }
                std::string piece;
                for (size_t q = 0; q < tokstr.size();) {
                    if (spm && tokstr.compare(q, 3, "\xE2\x96\x81") == 0) {
                        piece.push_back(' ');
                        q += 3;
                    } else {
                        piece.push_back(tokstr[q++]);
                    }
                }
                if (spm && t == t0 && !piece.empty() && piece[0] == ' ') piece.erase(0, 1);
                out += piece;
            }
            const int64_t p0 = doc_offs[d];
            ok[d] = valid && (int64_t)out.size() == pe - p0 && memcmp(out.data(), text + p0, out.size()) == 0;
            continue;
        }
        int64_t p = doc_offs[d];
        bool good = true;
        for (int64_t base = t0; base < t1 && good; base += 32) {
            int64_t a[32], b[32];
            bool fine[32];
            int32_t dl[32], inc[32];
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t t = base + lane;
                a[lane] = b[lane] = 0;
                fine[lane] = true;
                dl[lane] = 0;
                if (t < t1) {
                    fine[lane] = dpt_tok_span(V, ids[t], a[lane], b[lane]);
                    if (fine[lane]) dl[lane] = dpt_tok_decoded_len(V, a[lane], b[lane], spm, t == t0);
                }
            }
            int32_t run = 0;
            for (int lane = 0; lane < 32; ++lane) inc[lane] = run += dl[lane];
            bool all = true;
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t t = base + lane;
                if (fine[lane] && t < t1)
                    fine[lane] = dpt_tok_matches(V, a[lane], b[lane], spm, t == t0, text, p + inc[lane] - dl[lane], pe);
                all = all && fine[lane];
            }
            good = all;
            p += inc[31];
        }
        ok[d] = (good && p == pe) ? 1 : 0;
    }
}
}
