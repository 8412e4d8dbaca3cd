// Host-side compiled vocabulary (the dpt_vocab handle behind include/dptok.h).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "dpt_common.h"

struct dpt_vocab {
    int32_t unit_mode = 0;
    int32_t n_tokens = 0;
    int32_t n_nodes = 0;
    int32_t id_space = 0;
    uint32_t lmax = 0;
    uint32_t ph_salt = 0;
    uint32_t marker_entry = 0;
    uint32_t marker_slot = 0;
    int32_t bos_len = 0, bos_ntok = 0, bos_ids[3] = {0, 0, 0};
    int32_t fast_ok = 0;
    uint32_t ascii_single[4] = {0, 0, 0, 0};
    int32_t marker_leading_only = 1;
    int32_t byte_fallback = 0;
    int32_t byte_token_id[256];

    std::vector<uint32_t> da;       // double array
    std::vector<int32_t> slot_id;   // token id per slot
    std::vector<uint32_t> ph_seed;  // per bucket
    std::vector<int32_t> ph_id;     // per slot
    std::vector<uint8_t> tok_bytes; // token strings, dense-rank order
    std::vector<int64_t> tok_offs;  // n_tokens+1
    std::vector<int32_t> tok_ids;   // dense rank -> id
    std::vector<int32_t> id_rank;   // id -> dense rank or -1
    std::vector<unsigned long long> merge_keys, merge_vals;  // BPE merges (dpt_merge_lookup); empty = none
    uint8_t code2[2048];                                     // DptVocabView::code2

    // device copy
    int device = -1;
    void* d_blob = nullptr;
    int64_t blob_bytes = 0;
    DptVocabView d_view{};  // pointers into d_blob
    DptVocabView h_view{};  // pointers into the host vectors

    void rebuild_host_view();
    int set_merges(const int32_t* left, const int32_t* right, const int32_t* merged, int32_t n);
    void derive_facts();  // marker/bos/fast_ok/byte tokens from the built arrays (also after deserialize)
};

int dpt_vocab_build(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n_tokens,
                    int32_t unit_mode, dpt_vocab** out, std::string& err);
