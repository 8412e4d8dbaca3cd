/*
 * dptok.h - C ABI of the B200-native shortest-tokenization path.
 *
 * The reference (smfsamir/dp-tokenization) has no FFI layer: its boundary is
 * the Python signatures of packages/dp_tokenize.py and
 * packages/tokenizer_utils.py.  This header is the C boundary a binding for
 * that surface calls (ctypes stub: dp-tokenization_b200/dptok/_cabi.py; see
 * INTEGRATION.md).  Every entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - pointers named d_* are DEVICE pointers owned by the caller (PyTorch
 *     tensors' data_ptr()); the library owns only the dpt_vocab handle and the
 *     device copy of the compiled vocabulary hanging off it.
 *   - all device work is enqueued on the caller's `stream` (a cudaStream_t
 *     passed as void*; NULL = legacy default stream); nothing synchronises
 *     unless stated.
 *   - every function returns a dpt_status (0 = ok); dpt_last_error() gives a
 *     thread-local message.  Nothing aborts, nothing drops into a debugger
 *     (the reference calls ipdb.set_trace() at tokenizer_utils.py:72-73).
 *   - an untokenizable word is DATA, not an error: flag bit + phantom length
 *     (dp_tokenize.py:28,70) and no ids.
 *   - the handle is immutable after dpt_vocab_upload(): share it freely across
 *     host threads and streams.  One process per GPU in the sharded driver.
 *   - there is no CPU execution path behind these calls: without a CUDA device
 *     every compute entry point returns DPT_ECUDA.
 */
#ifndef DPTOK_H
#define DPTOK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpt_vocab dpt_vocab;

enum dpt_status {
    DPT_OK = 0,
    DPT_EINVAL = 1,     /* bad argument */
    DPT_ECUDA = 2,      /* CUDA runtime error / no device */
    DPT_ECAPACITY = 3,  /* a caller-provided buffer is too small; retry bigger */
    DPT_ENOMEM = 4,
    DPT_ESTATE = 5      /* e.g. vocab not uploaded to this device */
};

/* unit_mode: what one DP position and one unit of "token length" mean.
 *   BYTES      - byte-level BPE after the inverse bytes_to_unicode map
 *                (tokenizer_utils.py:149-151: one unit per mapped char).
 *   CODEPOINTS - SentencePiece-style vocab of UTF-8 strings; the DP of
 *                dp_tokenize.py:27-47 runs over code points and the tie-break
 *                of dp_tokenize.py:82 counts code points.                     */
enum dpt_unit_mode { DPT_UNIT_BYTES = 0, DPT_UNIT_CODEPOINTS = 1 };

/* Pre-tokenization rule.
 *   PRESPLIT  - caller supplies word offsets (pretokenize_with_llama /
 *               pre_tokenize_str run on the host; tokenizer_utils.py:24-31,157-159)
 *   SPM_LLAMA - normaliser Prepend(U+2581) + Replace(' ',U+2581), byte-fallback
 *               expansion of out-of-vocab characters to literal "<0xHH>" text,
 *               a word starts at every U+2581 (tokenizer_utils.py:12-17,24-31)
 *   GPT2      - ByteLevel(use_regex=True) split
 *   LLAMA3    - Llama-3 split regex
 *   BLOOM     - BLOOM split regex ' ?[^(\s|[.,!?...])]+' + isolated gaps
 *               (the pre-tokenizer tokenizer_utils.py:157-159 was written for)    */
enum dpt_rule {
    DPT_RULE_PRESPLIT = 0,
    DPT_RULE_SPM_LLAMA = 1,
    DPT_RULE_GPT2 = 2,
    DPT_RULE_LLAMA3 = 3,
    DPT_RULE_BLOOM = 4
};

/* per-word flag bits written by the encode calls */
#define DPT_WF_UNTOKENIZABLE 0x01u /* reference returns [] (dp_tokenize.py:66-70) */
#define DPT_WF_DOC_FIRST     0x02u /* first word of a document                     */
#define DPT_WF_LONG          0x04u /* went through the global-scratch long path    */

/* per-document flag bits written by dpt_encode_corpus */
#define DPT_DF_AMBIGUOUS     0x01u /* SPM_LLAMA: run of >=2 U+2581; word split depends on BPE
                                      merge order -> caller must pre-split this doc on the host */

/* indices into the int64 counters[4] vector (SURVEY.md section 5 "metrics") */
enum { DPT_CTR_BYTES = 0, DPT_CTR_WORDS = 1, DPT_CTR_TOKENS = 2, DPT_CTR_UNTOKENIZABLE = 3 };

typedef struct dpt_vocab_info {
    int32_t n_tokens;        /* distinct non-empty byte strings compiled           */
    int32_t unit_mode;
    int32_t n_nodes;         /* trie nodes                                         */
    int32_t n_slots;         /* double-array slots (4 B each)                      */
    int32_t max_token_bytes; /* Lmax                                               */
    int32_t ph_buckets;      /* perfect hash: seed table entries                   */
    int32_t ph_slots;        /* perfect hash: id table entries                     */
    int32_t marker_leading_only; /* 1 if no token has U+2581 after a non-U+2581 char
                                    (then SPM_LLAMA boundaries are exact for single spaces) */
    int32_t byte_fallback;   /* 1 if all 256 "<0xHH>" strings are tokens           */
    int32_t device;          /* device the blob is uploaded to, -1 if none         */
    int64_t blob_bytes;      /* size of the device-resident compiled vocabulary    */
} dpt_vocab_info;

/* ---- vocabulary compiler: replaces `vocab = set(tok.get_vocab())` (tokenizer_utils.py:57),
 *      `vocab_to_index` (tokenizer_utils.py:105-113) and the `in vocabulary` hash probe of
 *      dp_tokenize.py:39.  Builds a double-array byte trie + a perfect hash bytes->id.
 *      bytes/offs: token k is bytes[offs[k] .. offs[k+1]); ids[k] its id (any int32 >= 0).
 *      Host-only; no CUDA needed. */
int dpt_vocab_create(const uint8_t* bytes, const int64_t* offs, const int32_t* ids,
                     int32_t n_tokens, int32_t unit_mode, dpt_vocab** out);
void dpt_vocab_destroy(dpt_vocab* v);
int dpt_vocab_get_info(const dpt_vocab* v, dpt_vocab_info* out);
/* host-side exact lookup through the perfect hash (+ verify): id or -1 */
int dpt_vocab_lookup(const dpt_vocab* v, const uint8_t* s, int32_t len, int32_t* id_out);
/* serialise the compiled vocabulary (the only thing worth caching).
 * Call with buf=NULL to get the size in *need. */
int dpt_vocab_serialize(const dpt_vocab* v, uint8_t* buf, int64_t cap, int64_t* need);
int dpt_vocab_deserialize(const uint8_t* buf, int64_t len, dpt_vocab** out);
/* BPE merges of the tokenizer, in rank order: merge k joins token ids left[k] + right[k] into merged[k]
 * (tokenizer.json model.merges).  Optional, SPM_LLAMA rule only, before dpt_vocab_upload: with them the
 * device decides the word boundaries inside runs of U+2581 / spaces the way the reference's
 * tokenizer-driven split does (tokenizer_utils.py:7-31: a word starts at every DEFAULT-tokenizer token
 * that starts with U+2581 - inside a marker run that depends on the merge order); without them documents
 * with such runs are flagged DPT_DF_AMBIGUOUS and the caller splits them on the host.  Not serialised. */
int dpt_vocab_set_merges(dpt_vocab* v, const int32_t* left, const int32_t* right, const int32_t* merged,
                         int32_t n_merges);
/* copy the compiled vocabulary into HBM of `device` (synchronous) */
int dpt_vocab_upload(dpt_vocab* v, int device);

/* ---- status vector written by the encode calls: device int64[8]
 *      [0] ids required      (compare with ids_cap)      [1] words found (compare with word_cap)
 *      [2] long-word pool positions required             [3] long-word pool capacity
 *      [4] normalised-text bytes required (SPM_LLAMA)    [5] normalised-text capacity
 *      A value above its capacity means the corresponding outputs were truncated: retry with
 *      larger buffers / worst_case=1 workspace.                                             */
#define DPT_NOUT_IDS 0
#define DPT_NOUT_WORDS 1
#define DPT_NOUT_POOL_REQ 2
#define DPT_NOUT_POOL_CAP 3
#define DPT_NOUT_NORM_REQ 4
#define DPT_NOUT_NORM_CAP 5
#define DPT_NOUT_ODD_REQ 6  /* dpt_encode_corpus: see there */
#define DPT_NOUT_ODD_CAP 7

/* ---- workspace sizing (bytes of caller-owned device scratch).  worst_case=0 sizes the
 *      variable parts for typical text; worst_case=1 can never overflow; worst_case=2 (corpus
 *      pipeline only) = typical sizes with a word table for one distinct word per 10 bytes
 *      (text that repeats few of its words: fewer first occurrences overflow into the
 *      per-occurrence odd-word path).  The same value must be passed to the sizing call and
 *      to the encode call. */
int64_t dpt_pretokenize_workspace(int64_t n_bytes, int64_t n_docs);
int64_t dpt_encode_words_workspace(int64_t n_bytes, int64_t n_words, int32_t worst_case);
int64_t dpt_encode_corpus_workspace(int32_t rule, int64_t n_bytes, int64_t n_docs, int64_t word_cap,
                                    int32_t worst_case);
int64_t dpt_encode_corpus_general_workspace(int32_t rule, int64_t n_bytes, int64_t n_docs, int64_t word_cap,
                                            int32_t worst_case);

/* ---- boundary kernel: replaces pretokenize_with_llama / pre_tokenize_str
 *      (tokenizer_utils.py:24-31,157-159) for the rule classes above.
 *      d_text: concatenated NON-EMPTY documents, d_doc_offs[n_docs+1] their byte offsets.
 *      SPM_LLAMA: the text is normalised into d_norm_text (capacity norm_cap bytes) as
 *      "<s>" + U+2581 + text with ' '->U+2581 and out-of-vocab characters spelled "<0xHH>"
 *      per byte, exactly the concatenated token strings the reference's DP sees; '<s>' is a
 *      word of its own (tokenizer_utils.py:26-30).  d_norm_doc_offs[n_docs+1] and
 *      d_word_offs[<=word_cap+1] index d_norm_text; d_doc_first_word[n_docs] is the word index
 *      of each document's '<s>' word (may be NULL).
 *      d_n_out: device int64[2] = {n_words, n_norm_bytes} (compare with the capacities).
 *      d_doc_flags: uint8[n_docs] (DPT_DF_*), may be NULL. */
int dpt_pretokenize(const dpt_vocab* v, int32_t rule,
                    const uint8_t* d_text, int64_t n_bytes,
                    const int64_t* d_doc_offs, int64_t n_docs,
                    uint8_t* d_norm_text, int64_t norm_cap, int64_t* d_norm_doc_offs,
                    int64_t* d_word_offs, int64_t word_cap, int64_t* d_doc_first_word,
                    uint8_t* d_doc_flags, int64_t* d_n_out,
                    void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- the DP: replaces compute_shortest_tokenizations + obtain_longest_token + the id map
 *      (dp_tokenize.py:24-84, tokenizer_utils.py:70-80,165-174) for n_words pre-split words.
 *      Word w is d_text[d_word_offs[w] .. d_word_offs[w+1]) - already normalised bytes.
 *      d_ids: capacity ids_cap int32, ids of all words concatenated in word order
 *             (untokenizable words contribute none).
 *      d_word_lens[w]  = len_dp[n] of dp_tokenize.py:70 (phantom value when untokenizable)
 *      d_word_flags[w] = DPT_WF_* bits
 *      d_word_tok_offs = optional int64[n_words+1] offsets of each word's ids in d_ids (NULL ok)
 *      d_counters      = device int64[4] {bytes, words, tokens, untokenizable}, OVERWRITTEN
 *                        (bytes = n_text_bytes as given)
 *      d_n_out         = device int64[8] status vector (above)
 *      n_text_bytes    = size of d_text (>= d_word_offs[n_words]); the same value goes to
 *                        dpt_encode_words_workspace.  It sizes the id stash (one int32 per text byte)
 *                        that lets every word be solved ONCE: DP -> ids at the word's byte offset ->
 *                        scan over the token counts -> placement.  Words lying beyond n_text_bytes
 *                        are still encoded correctly, at the price of a second DP.
 *      Asynchronous: returns after enqueueing on `stream`. */
int dpt_encode_words(const dpt_vocab* v,
                     const uint8_t* d_text, const int64_t* d_word_offs, int64_t n_words,
                     int64_t n_text_bytes,
                     int32_t* d_ids, int64_t ids_cap,
                     int32_t* d_word_lens, uint8_t* d_word_flags, int64_t* d_word_tok_offs,
                     int64_t* d_counters, int64_t* d_n_out,
                     void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- corpus throughput path: boundary rule + DP + tie-break select + compaction over raw documents
 *      resident in HBM; five kernel launches (scan+dedup -> DP per DISTINCT word -> long words ->
 *      scan+emit -> counters), no host synchronisation (the per-document loops of
 *      main_analyze_s2orc.py:269-298 and main_biomed_translation.py:71-82 around
 *      tokenizer_utils.py:66-80).  d_text: concatenated NON-EMPTY documents; d_doc_offs[n_docs+1] with
 *      d_doc_offs[0] == 0 and d_doc_offs[n_docs] == n_bytes.  Outputs as dpt_encode_words plus
 *      d_doc_tok_offs[n_docs+1] (token offset of each document in d_ids; for SPM_LLAMA each
 *      document's ids start with the id of '<s>' exactly as dp_tokenize_llama's output does) and
 *      d_doc_flags[n_docs].  d_word_lens / d_word_flags have capacity word_cap.  Writes beyond a
 *      capacity are dropped and the requirement is reported in d_n_out:
 *        [0] ids required (ids_cap)            [1] words found (word_cap)
 *        [2] long-word scratch required / [3] capacity      [4] id-pool required / [5] capacity
 *        [6] not-deduplicated words required / [7] capacity
 *      a requirement above its capacity means: retry with larger buffers ([0],[1]) or with
 *      worst_case=1 workspace ([2..7]).  Nothing is cached between calls: the word table lives in
 *      the workspace and is rebuilt by every call. */
int dpt_encode_corpus(const dpt_vocab* v, int32_t rule,
                      const uint8_t* d_text, int64_t n_bytes,
                      const int64_t* d_doc_offs, int64_t n_docs,
                      int32_t* d_ids, int64_t ids_cap,
                      int32_t* d_word_lens, uint8_t* d_word_flags, int64_t word_cap,
                      int64_t* d_doc_tok_offs, uint8_t* d_doc_flags,
                      int64_t* d_counters, int64_t* d_n_out,
                      void* d_workspace, int64_t workspace_bytes, int32_t worst_case, void* stream);

/* ---- chunked corpus calls (Engine.encode_corpus_host): the corpus arrives in d_text range by range (ranges cut at
 *      document boundaries, processed in order on one stream); each call tokenizes documents
 *      [doc_begin, doc_end) = bytes [byte_begin, byte_end) and may look at bytes BEFORE byte_begin only.  The word
 *      table lives in d_table_workspace (dpt_corpus_table_workspace, sized for the WHOLE corpus) and is kept from
 *      the previous range unless reset_table != 0 (pass 1 for the first range): a word is still solved once per
 *      corpus, not once per chunk.  d_doc_offs holds corpus-global offsets; outputs are range-local exactly as
 *      dpt_encode_corpus would produce them for the range alone (ids from 0, d_doc_tok_offs[doc_end-doc_begin+1],
 *      d_doc_flags[doc_end-doc_begin]).  d_workspace: dpt_encode_corpus_range_workspace (per range).
 *      phases: bit 0 = scan + dedup (kernel A), bit 1 = DP of the words this range claimed (kernels B), bit 2 = emit
 *      (kernels C, D); 7 = the whole range.  Split calls let consecutive ranges overlap on different streams under
 *      these rules, which the caller enforces with stream dependencies:  scan(k+1) after scan(k)  (a range may only
 *      reference slots claimed by itself or an earlier range);  DP(k) after scan(k);  emit(k) after DP(j) for every
 *      j <= k.  DP(k) may run beside scan(k+1) and emit(k-1) - it is latency-bound and leaves most of the SMs idle.
 *      phases = 0 with reset_table = 1 just clears the table. */
int64_t dpt_corpus_table_workspace(int64_t n_bytes_total, int64_t word_cap_total, int32_t worst_case);
int64_t dpt_encode_corpus_range_workspace(int32_t rule, int64_t range_bytes, int64_t range_docs, int64_t word_cap,
                                          int32_t worst_case);
int dpt_encode_corpus_range(const dpt_vocab* v, int32_t rule,
                            const uint8_t* d_text, int64_t n_bytes_total,
                            const int64_t* d_doc_offs, int64_t n_docs_total,
                            int64_t byte_begin, int64_t byte_end, int64_t doc_begin, int64_t doc_end,
                            int32_t reset_table, int64_t table_word_cap,
                            int32_t* d_ids, int64_t ids_cap,
                            int32_t* d_word_lens, uint8_t* d_word_flags, int64_t word_cap,
                            int64_t* d_doc_tok_offs, uint8_t* d_doc_flags,
                            int64_t* d_counters, int64_t* d_n_out,
                            void* d_table_workspace, int64_t table_workspace_bytes,
                            void* d_workspace, int64_t workspace_bytes, int32_t worst_case, int32_t phases,
                            void* stream);

/* ---- the same contract WITHOUT deduplication: normalise -> DP count per word -> scan -> DP emit per word
 *      (what dpt_pretokenize + dpt_encode_words compose to); synchronises the stream once.  Kept as
 *      the independent cross-check of the pipeline above (tests) and for callers who want the
 *      normalised text.  Its d_n_out layout is the one documented at DPT_NOUT_*. */
int dpt_encode_corpus_general(const dpt_vocab* v, int32_t rule,
                              const uint8_t* d_text, int64_t n_bytes,
                              const int64_t* d_doc_offs, int64_t n_docs,
                              int32_t* d_ids, int64_t ids_cap,
                              int32_t* d_word_lens, uint8_t* d_word_flags, int64_t word_cap,
                              int64_t* d_doc_tok_offs, uint8_t* d_doc_flags,
                              int64_t* d_counters, int64_t* d_n_out,
                              void* d_workspace, int64_t workspace_bytes, int32_t worst_case, void* stream);

/* ---- lattice of ONE word for the enumerate-all API (dp_tokenize.py:27-47):
 *      len_dp[0..n_units] and, per unit position, the ascending predecessor list
 *      (segment_index_dp of dp_tokenize.py:30-47) in CSR form.
 *      d_unit_starts: uint8[n_bytes], 1 where a unit starts (NULL = derive from unit_mode).
 *      d_len_dp: int32[n_units+1]; d_pred_offs: int32[n_units+2]; d_pred: int32[pred_cap]
 *      (unit indices); d_scratch: int32[n_bytes+1].  d_n_out: int32[2] = {n_units, n_pred}. */
int dpt_lattice_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes,
                     const uint8_t* d_unit_starts,
                     int32_t* d_len_dp, int32_t* d_pred_offs, int32_t* d_pred, int32_t pred_cap,
                     int32_t* d_n_out, int32_t* d_scratch, void* stream);

/* ---- length-only DP with INFINITY initialisation (inspect_tokenizer.py:77-86, `min_tokens_for_string`;
 *      vectors tests/test_tokenization_algorithms.py:14-30): dp[0] = 0, dp[i] = min dp[j] + 1 over vocabulary
 *      edges.  Differs from d_len_dp of dpt_lattice_word only on untokenizable input (that one carries the
 *      phantom initialisation of dp_tokenize.py:28).  d_out: int32[1] = dp[n_units], -1 = infinity;
 *      d_scratch: int32[n_bytes+1].  One serial thread: a known-answer entry point, not a throughput path. */
int dpt_min_tokens_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes,
                        const uint8_t* d_unit_starts, int32_t* d_out, int32_t* d_scratch, void* stream);

/* ---- decode / round-trip check on device (tokenizer_utils.py:82-84,176-179; the asserts at
 *      main_analyze_s2orc.py:85 and main_biomed_translation.py:78): ids -> bytes via the
 *      id->string table (CODEPOINTS vocab: U+2581 -> ' ', "<0xHH>" tokens -> that byte, the one
 *      leading space of the Prepend normaliser dropped), compared with the RAW text of each
 *      document.  skip_bos=1 ignores the first id of every document.  d_ok: uint8[n_docs]. */
int dpt_roundtrip_check(const dpt_vocab* v, const int32_t* d_ids, const int64_t* d_doc_tok_offs,
                        const uint8_t* d_text, const int64_t* d_doc_offs, int64_t n_docs,
                        int32_t skip_bos, uint8_t* d_ok, void* stream);

/* ---- training-data feed on device: token stream -> padded int64 batch, no host round trip.
 *      Replaces tokenized_dict['input_ids'] / ['attention_mask'] = [1] * len (main_analyze_s2orc.py:87-89) followed
 *      by DataCollatorWithPadding, and - with a second stream - the custom collator of
 *      main_biomed_translation.py:104-124 that concatenates input_ids + labels and pads with pad_token_id.
 *      Row r = ids of document doc_begin + r of stream A (d_ids_a, d_doc_tok_offs_a as written by the encode calls),
 *      followed by that document's ids of stream B if d_ids_b != NULL; truncated to row_len; padded with pad_id on the
 *      right (pad_left = 0) or left.  d_input_ids / d_attention_mask: int64[n_rows * row_len] (mask may be NULL);
 *      d_row_lens: int64[n_rows] real tokens per row (may be NULL). */
int dpt_pad_batch(const int32_t* d_ids_a, const int64_t* d_doc_tok_offs_a,
                  const int32_t* d_ids_b, const int64_t* d_doc_tok_offs_b,
                  int64_t doc_begin, int64_t n_rows, int64_t row_len, int64_t pad_id, int32_t pad_left,
                  int64_t* d_input_ids, int64_t* d_attention_mask, int64_t* d_row_lens, void* stream);

/* ---- compact output: token ids as uint16 for vocabularies of at most 65,536 entries (Llama-2 32k, GPT-2 50k).
 *      The ids of a corpus leave the GPU over PCIe; at 2 bytes per id the device-to-host copy of the adapters' result
 *      (the flat List[int] of tokenizer_utils.py:76-80 / :170-174) is half as large.  Narrows the first
 *      min(*d_n, cap) entries of d_ids into d_out; d_n is a DEVICE pointer (e.g. &n_out[DPT_NOUT_IDS] of an encode
 *      call enqueued before on the same stream), so no host synchronisation is needed to learn the count.
 *      Ids that do not fit 16 bits are written as 0xFFFF and counted in *d_overflow (device int64, may be NULL;
 *      the caller zeroes it). */
int dpt_narrow_ids_u16(const int32_t* d_ids, const int64_t* d_n, int64_t cap, uint16_t* d_out, int64_t* d_overflow,
                       void* stream);

/* ---- measurement support (BASELINE.json configs[3] / configs[4]): synthetic corpora generated
 *      ON THE DEVICE from a counter-based RNG, so a 1 GB corpus of long documents or 1.25 GB of Arabic-script text per
 *      GPU never crosses PCIe.  Not part of the tokenization path (the reference reads its inputs from datasets:
 *      main_analyze_s2orc.py:253-255, main_biomed_translation.py:71-73, dialect_arabic.py:24); a document is a pure
 *      function of (seed, global document index), so every rank can generate the same corpus or its own range of it.
 *      Lexicon = word strings (d_lex_bytes / d_lex_offs[n_types+1]) + d_lex_cdf[n_types], the cumulative word
 *      probabilities scaled to 32 bits; an optional second lexicon B is used for a document with probability
 *      frac_b / 2^32 (e.g. Arabic-script documents among English ones).
 *      Two passes: d_doc_offs == NULL writes the byte length of each document to d_doc_len[n_docs]; the caller forms
 *      the offsets (exclusive prefix sum) and calls again with d_doc_offs[n_docs] and d_text to fill the bytes. */
typedef struct dpt_synth_params {
    uint64_t seed;
    int32_t words_lo, words_hi;  /* words per document, uniform                                            */
    int32_t sentence_mean;       /* a sentence ends after each word with probability 1 / sentence_mean     */
    int32_t flags;               /* bit 0: lexicon A is ASCII (capitalise sentence starts), bit 1: B is    */
    uint32_t frac_b;             /* P(document uses lexicon B) * 2^32                                      */
    uint32_t suffix_prob;        /* P(a plain word gets 4 hash-derived letters appended) * 2^32: raises the share of
                                    DISTINCT words (redundancy sweep of the dedup pipeline; 0 = off)          */
} dpt_synth_params;
int dpt_synth_corpus(const uint8_t* d_lex_a_bytes, const int64_t* d_lex_a_offs, const uint32_t* d_lex_a_cdf, int32_t n_a,
                     const uint8_t* d_lex_b_bytes, const int64_t* d_lex_b_offs, const uint32_t* d_lex_b_cdf, int32_t n_b,
                     const dpt_synth_params* params, int64_t doc_base, int64_t n_docs,
                     int64_t* d_doc_len, const int64_t* d_doc_offs, uint8_t* d_text, void* stream);

const char* dpt_last_error(void);
const char* dpt_version(void);
/* number of kernel launches issued by this library in the calling process (bench "gpu_launches") */
int64_t dpt_launch_count(void);
/* per-kernel device timing with CUDA events on the launching stream (bench.py roofline leg):
 * enable, run, then fetch "kernel_name launches total_ms" lines (call with buf=NULL for the size;
 * the report synchronises and clears the records). */
void dpt_profile_enable(int32_t on);
int dpt_profile_report(char* buf, int64_t cap, int64_t* need);

#ifdef __cplusplus
}
#endif
#endif /* DPTOK_H */
