// Internal interface between the C ABI (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

struct dpt_vocab;
struct dpt_synth_params;

namespace dpt {

extern std::atomic<int64_t> g_launches;
void profile_enable(int on);
std::string profile_report();

// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg)
struct ProfScope {
    const char* name;
    cudaStream_t st;
    cudaEvent_t a = nullptr;
    ProfScope(const char* n, cudaStream_t s);
    ~ProfScope();
};

// corpus pipeline (pipe.cu / dpt_pipe.h): scan+dedup -> DP per distinct word -> scan+emit; asynchronous
int64_t encode_corpus_pipe_workspace(int64_t n_bytes, int64_t n_docs, int64_t word_cap, int32_t worst);
int64_t corpus_table_workspace(int64_t n_bytes_total, int64_t word_cap_total, int32_t worst);
int64_t corpus_range_workspace(int64_t range_bytes, int64_t range_docs, int64_t word_cap, int32_t worst);
// one range [byte_begin, byte_end) = documents [doc_begin, doc_end) of a corpus resident in d_text; the word table in
// d_table_ws is kept from the previous range of the same call unless reset_table
int encode_corpus_range(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes_total,
                        const int64_t* d_doc_offs, int64_t n_docs_total, int64_t byte_begin, int64_t byte_end,
                        int64_t doc_begin, int64_t doc_end, int32_t reset_table, int64_t table_bytes_total,
                        int64_t table_word_cap, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                        int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                        int64_t* d_n_out, void* d_table_ws, int64_t table_ws_bytes, void* d_ws, int64_t ws_bytes,
                        int32_t worst, int32_t phases, cudaStream_t st, std::string& err);
int encode_corpus_pipe(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs,
                       int64_t n_docs, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                       int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                       int64_t* d_n_out, void* d_ws, int64_t ws_bytes, int32_t worst, cudaStream_t st, std::string& err);

int64_t encode_words_workspace_fixed(int64_t n_words, int64_t n_text_bytes);
int64_t pretokenize_workspace(int64_t n_bytes, int64_t n_docs);

// d_n_out: int64[8] status vector (DPT_NOUT_* in include/dptok.h).  text_extent: bytes of d_text the words lie in (sizes
// the id stash; words beyond it are solved a second time instead of being read from the stash).
int encode_words(const dpt_vocab* v, const uint8_t* d_text, const int64_t* d_word_offs, int64_t n_words,
                 int64_t n_bytes_for_counter, int64_t text_extent, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens,
                 uint8_t* d_word_flags, int64_t* d_word_tok_offs, int64_t* d_counters, int64_t* d_n_out, void* d_ws,
                 int64_t ws_bytes, cudaStream_t st, std::string& err);

// d_n_out: int64[2] = {n_words, n_norm_bytes}
int pretokenize_spm(const dpt_vocab* v, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs, int64_t n_docs,
                    uint8_t* d_norm, int64_t norm_cap, int64_t* d_norm_doc_offs, int64_t* d_word_offs, int64_t word_cap,
                    int64_t* d_doc_first_word, uint8_t* d_doc_flags, int64_t* d_n_out, void* d_ws, int64_t ws_bytes,
                    cudaStream_t st, std::string& err);

int doc_tok_offsets(const int64_t* d_doc_first_word, int64_t n_docs, const int64_t* d_tok_offs, int64_t n_words,
                    int64_t* d_doc_tok_offs, cudaStream_t st);

int roundtrip_check(const dpt_vocab* v, const int32_t* d_ids, const int64_t* d_doc_tok_offs, const uint8_t* d_text,
                    const int64_t* d_doc_offs, int64_t n_docs, int32_t skip_bos, uint8_t* d_ok, cudaStream_t st);

int narrow_ids_u16(const int32_t* d_ids, const int64_t* d_n, int64_t cap, uint16_t* d_out, int64_t* d_overflow, cudaStream_t st);
int pad_batch(const int32_t* d_ids_a, const int64_t* d_offs_a, const int32_t* d_ids_b, const int64_t* d_offs_b,
              int64_t doc_begin, int64_t n_rows, int64_t row_len, int64_t pad_id, int32_t pad_left, int64_t* d_input_ids,
              int64_t* d_attention_mask, int64_t* d_row_lens, cudaStream_t st);

int lattice_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts,
                 int32_t* d_len_dp, int32_t* d_pred_offs, int32_t* d_pred, int32_t pred_cap, int32_t* d_n_out,
                 int32_t* d_unit_of, cudaStream_t st);

int min_tokens_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts, int32_t* d_out,
                    int32_t* d_scratch, cudaStream_t st);

// synthetic corpus generator on the device (synth.cu; measurement support, not on the tokenization path)
int synth_run(const uint8_t* a_bytes, const int64_t* a_offs, const uint32_t* a_cdf, int32_t a_n, const uint8_t* b_bytes,
              const int64_t* b_offs, const uint32_t* b_cdf, int32_t b_n, const dpt_synth_params* sp, int64_t doc_base,
              int64_t n_docs, int64_t* d_doc_len, const int64_t* d_doc_offs, uint8_t* d_text, cudaStream_t st,
              std::string& err);

}  // namespace dpt
