// Shared host/device definitions: compiled-vocabulary layout, double-array trie step,
// perfect-hash lookup.  Everything here is plain C++ so the same code runs in the CUDA kernels
// and in the host-side compiler / self-check (tests/host_sim builds it with g++).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DPT_HD __host__ __device__ __forceinline__
#else
#define DPT_HD inline
#endif

// ---- double-array trie -------------------------------------------------------------------
// One 32-bit slot per trie node:  [31:10] base  [9] occupied  [8] terminal  [7:0] label.
// Child of a node with base b on byte c lives in slot b+c and must carry label c.  Bases are
// unique across internal nodes, so "slot b+c is occupied with label c" proves parentage and no
// separate check[] array is needed: one 4-byte load per trie step.  base==0 means leaf.
// Slot 0 is reserved; the root has base 1 (children in slots 1..256) and no slot of its own.
#define DPT_DA_LABEL_MASK 0x000000FFu
#define DPT_DA_TERMINAL   0x00000100u
#define DPT_DA_OCCUPIED   0x00000200u
#define DPT_DA_MATCH_MASK 0x000002FFu
#define DPT_DA_BASE_SHIFT 10
#define DPT_DA_ROOT_ENTRY ((1u << DPT_DA_BASE_SHIFT) | DPT_DA_OCCUPIED)
#define DPT_DA_MAX_SLOTS  (1u << 22)

// U+2581 LOWER ONE EIGHTH BLOCK, the SentencePiece word marker, as UTF-8
#define DPT_MARK0 0xE2u
#define DPT_MARK1 0x96u
#define DPT_MARK2 0x81u

struct DptVocabView {
    const uint32_t* da;        // n_slots double-array slots
    const int32_t* slot_id;    // token id of the node in each slot (-1 if not terminal)
    const uint32_t* ph_seed;   // perfect hash level 1: per-bucket displacement seed
    const int32_t* ph_id;      // perfect hash level 2: slot -> token id (-1 empty)
    const uint8_t* tok_bytes;  // id-indexed token strings (decode / verify)
    const int64_t* tok_offs;   // tok_bytes offsets, indexed by dense id rank
    const int32_t* id_rank;    // token id -> dense rank (or -1), length id_space
    const uint8_t* uni1;       // code-point class table of the byte-level split rules (dpt_split_rules.h)
    const uint8_t* uni2;
    uint32_t n_slots;
    uint32_t ph_bucket_mask;
    uint32_t ph_slot_mask;
    uint32_t ph_salt;
    uint32_t lmax;             // longest token, bytes
    int32_t unit_mode;
    int32_t id_space;          // max id + 1
    uint32_t marker_entry;     // DA entry after consuming U+2581 from the root (0 if absent)
    uint32_t ascii_single[4];  // bit c set  <=>  the 1-byte string c is a token
    uint32_t marker_slot;      // slot index of that node (slot_id[marker_slot] = id of the bare marker)
    // the word "<s>" every SPM_LLAMA document starts with (tokenizer_utils.py:26-30), solved once at compile time
    int32_t bos_len;           // len_dp[n] of the word "<s>"
    int32_t bos_ntok;          // ids it contributes (0 when untokenizable)
    int32_t bos_ids[3];
    // all 256 single bytes are tokens (byte-level) / U+2581 alone is a token (code points): then no position
    // is unreachable for in-vocabulary characters and the phantom init of dp_tokenize.py:28 never undercuts
    int32_t fast_ok;
    // BPE merge table of the tokenizer (dpt_vocab_set_merges): open addressing, key = (left id + 1) << 32 | (right id + 1),
    // value = rank << 32 | merged id; merge_mask = slots - 1, 0 when the caller gave no merges.  Only the SPM_LLAMA rule's
    // marker runs need it (tokenizer_utils.py:7-31: word boundaries follow the DEFAULT tokenizer's tokens).
    const unsigned long long* merge_keys;
    const unsigned long long* merge_vals;
    uint32_t merge_mask;
    const int32_t* byte_ids;   // id of the token "<0xHH>" per byte value (-1 if absent): the default tokenizer's symbol for
                               // a character that is no vocabulary entry (byte_fallback)
    const uint8_t* code2;      // dpt_char_code of every 2-byte UTF-8 character, indexed by its 11-bit code point (2048 B):
                               // Latin-1 .. Arabic in ONE look-up in kernel A's code pass (dpt_split_rules.h)
};

// One trie step.  `entry` is the slot VALUE of the current node (it carries the base), not its
// index.  Returns false when the edge does not exist.
DPT_HD bool dpt_da_step(const uint32_t* __restrict__ da, uint32_t& entry, uint32_t c) {
    const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
    if (base == 0) return false;
    const uint32_t e = da[base + c];
    if ((e & DPT_DA_MATCH_MASK) != (DPT_DA_OCCUPIED | c)) return false;
    entry = e;
    return true;
}
// Same, also reporting the slot index (needed to read slot_id).
DPT_HD bool dpt_da_step_idx(const uint32_t* __restrict__ da, uint32_t& entry, uint32_t c, uint32_t& slot) {
    const uint32_t base = entry >> DPT_DA_BASE_SHIFT;
    if (base == 0) return false;
    const uint32_t t = base + c;
    const uint32_t e = da[t];
    if ((e & DPT_DA_MATCH_MASK) != (DPT_DA_OCCUPIED | c)) return false;
    entry = e;
    slot = t;
    return true;
}

// ---- perfect hash bytes -> id ------------------------------------------------------------
// Two 32-bit streaming hashes over the token bytes; h1 selects a bucket whose seed d displaces
// h2 into a collision-free slot (compress-hash-displace).  Device lookups are only made for
// strings known to be tokens (edges chosen by the DP), so no key comparison is needed there;
// the host lookup verifies against the token pool.
struct DptHashState {
    uint32_t h1, h2;
};
DPT_HD DptHashState dpt_hash_init(uint32_t salt) {
    DptHashState s;
    s.h1 = 0x811C9DC5u ^ salt;
    s.h2 = 0x9E3779B9u + salt * 0x85EBCA6Bu;
    return s;
}
DPT_HD void dpt_hash_byte(DptHashState& s, uint32_t b) {
    s.h1 = (s.h1 ^ b) * 0x01000193u;
    s.h2 = (s.h2 + b + 1u) * 0xCC9E2D51u;
    s.h2 = (s.h2 << 13) | (s.h2 >> 19);
}
DPT_HD uint32_t dpt_fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
DPT_HD uint32_t dpt_ph_bucket(const DptHashState& s, uint32_t bucket_mask) { return dpt_fmix32(s.h1) & bucket_mask; }
DPT_HD uint32_t dpt_ph_slot(const DptHashState& s, uint32_t seed, uint32_t slot_mask) {
    return dpt_fmix32(s.h2 ^ (seed * 0x9E3779B1u + 0x7F4A7C15u)) & slot_mask;
}
DPT_HD int32_t dpt_ph_lookup(const DptVocabView& v, const DptHashState& s) {
    const uint32_t seed = v.ph_seed[dpt_ph_bucket(s, v.ph_bucket_mask)];
    return v.ph_id[dpt_ph_slot(s, seed, v.ph_slot_mask)];
}

DPT_HD uint32_t dpt_merge_hash(int32_t l, int32_t r) {
    return dpt_fmix32((uint32_t)l * 0x9E3779B1u + dpt_fmix32((uint32_t)r + 0x7F4A7C15u));
}
// rank and merged id of the BPE merge (l, r); false when the pair does not merge
DPT_HD bool dpt_merge_lookup(const DptVocabView& v, int32_t l, int32_t r, uint32_t& rank, int32_t& merged) {
    if (!v.merge_mask || l < 0 || r < 0) return false;
    const unsigned long long key = ((unsigned long long)(uint32_t)(l + 1) << 32) | (unsigned long long)(uint32_t)(r + 1);
    uint32_t h = dpt_merge_hash(l, r) & v.merge_mask;
    for (;;) {
        const unsigned long long k = v.merge_keys[h];
        if (k == key) {
            const unsigned long long val = v.merge_vals[h];
            rank = (uint32_t)(val >> 32);
            merged = (int32_t)(uint32_t)val;
            return true;
        }
        if (k == 0ull) return false;
        h = (h + 1u) & v.merge_mask;
    }
}

DPT_HD bool dpt_is_cp_start(uint32_t b) { return (b & 0xC0u) != 0x80u; }
