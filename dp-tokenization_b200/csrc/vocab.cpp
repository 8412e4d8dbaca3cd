// Vocabulary compiler: token byte strings -> double-array trie (hot nodes first) + perfect hash.
// Replaces the Python `set`/`dict` of tokenizer_utils.py:57,105-113 whose `in` probe is the inner
// operation of dp_tokenize.py:39.  Host-only C++; the result is uploaded to HBM as one blob.
#include "vocab.h"
#include "dpt_split_rules.h"

#include "dpt_dp_core.h"
#include "dpt_unicode_tables.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>

namespace {

struct Node {
    std::vector<std::pair<uint8_t, int32_t>> kids;  // sorted by byte after build
    int32_t token = -1;                             // dense rank of the token ending here
    float hot = 0.f;
    uint32_t base = 0;
    uint32_t slot = 0;
    uint8_t label = 0;
};

struct FreeList {
    std::vector<uint32_t> nxt;
    void grow(size_t n) {
        size_t old = nxt.size();
        if (n <= old) return;
        nxt.resize(n);
        for (size_t i = old; i < n; ++i) nxt[i] = (uint32_t)i;
    }
    uint32_t find(uint32_t i) {
        if (i >= nxt.size()) grow((size_t)i + 1024);
        uint32_t r = i;
        while (nxt[r] != r) {
            r = nxt[r];
            if (r >= nxt.size()) grow((size_t)r + 1024);
        }
        while (nxt[i] != r) {
            uint32_t t = nxt[i];
            nxt[i] = r;
            i = t;
        }
        return r;
    }
    bool is_free(uint32_t i) { return find(i) == i; }
    void take(uint32_t i) {
        if (i + 1 >= nxt.size()) grow((size_t)i + 1024);
        nxt[i] = i + 1;
    }
};

int child_of(const std::vector<Node>& nodes, int32_t n, uint8_t c) {
    for (auto& kv : nodes[n].kids)
        if (kv.first == c) return kv.second;
    return -1;
}

uint32_t next_pow2(uint64_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace

void dpt_vocab::rebuild_host_view() {
    h_view.da = da.data();
    h_view.slot_id = slot_id.data();
    h_view.ph_seed = ph_seed.data();
    h_view.ph_id = ph_id.data();
    h_view.tok_bytes = tok_bytes.data();
    h_view.tok_offs = tok_offs.data();
    h_view.id_rank = id_rank.data();
    h_view.uni1 = DPT_UNI_STAGE1;
    h_view.uni2 = DPT_UNI_STAGE2;
    h_view.n_slots = (uint32_t)da.size();
    h_view.ph_bucket_mask = (uint32_t)ph_seed.size() - 1;
    h_view.ph_slot_mask = (uint32_t)ph_id.size() - 1;
    h_view.ph_salt = ph_salt;
    h_view.lmax = lmax;
    h_view.unit_mode = unit_mode;
    h_view.id_space = id_space;
    h_view.marker_entry = marker_entry;
    for (int k = 0; k < 4; ++k) h_view.ascii_single[k] = ascii_single[k];
    h_view.marker_slot = marker_slot;
    h_view.bos_len = bos_len;
    h_view.bos_ntok = bos_ntok;
    for (int k = 0; k < 3; ++k) h_view.bos_ids[k] = bos_ids[k];
    h_view.fast_ok = fast_ok;
    h_view.merge_keys = merge_keys.empty() ? nullptr : merge_keys.data();
    h_view.merge_vals = merge_vals.empty() ? nullptr : merge_vals.data();
    h_view.merge_mask = merge_keys.empty() ? 0u : (uint32_t)merge_keys.size() - 1u;
    h_view.byte_ids = byte_token_id;
    {   // (the same for every vocabulary: a function of the Unicode class tables only)
        const DptUniView U{DPT_UNI_STAGE1, DPT_UNI_STAGE2};
        for (uint32_t cp = 0; cp < 2048; ++cp) {
            DptChar c;
            c.cp = cp;
            c.len = 2;
            c.cls = dpt_cp_class(U, cp);
            code2[cp] = (uint8_t)dpt_char_code(c);
        }
    }
    h_view.code2 = code2;
}

// merges in rank order (index = rank); a pair listed twice keeps its first (lowest) rank, like the tokenizer's own map
int dpt_vocab::set_merges(const int32_t* left, const int32_t* right, const int32_t* merged, int32_t n) {
    merge_keys.clear();
    merge_vals.clear();
    if (n > 0) {
        const uint32_t slots = next_pow2((uint64_t)n * 2 + 16);
        merge_keys.assign(slots, 0ull);
        merge_vals.assign(slots, 0ull);
        for (int32_t k = 0; k < n; ++k) {
            if (left[k] < 0 || right[k] < 0 || merged[k] < 0) return 1;
            const unsigned long long key =
                ((unsigned long long)(uint32_t)(left[k] + 1) << 32) | (unsigned long long)(uint32_t)(right[k] + 1);
            uint32_t h = dpt_merge_hash(left[k], right[k]) & (slots - 1);
            while (merge_keys[h] != 0ull && merge_keys[h] != key) h = (h + 1u) & (slots - 1);
            if (merge_keys[h] == key) continue;
            merge_keys[h] = key;
            merge_vals[h] = ((unsigned long long)(uint32_t)k << 32) | (unsigned long long)(uint32_t)merged[k];
        }
    }
    rebuild_host_view();
    return 0;
}

void dpt_vocab::derive_facts() {
    dpt_vocab* v = this;
    {
        uint32_t e = DPT_DA_ROOT_ENTRY, slot = 0;
        bool ok = dpt_da_step(v->da.data(), e, DPT_MARK0) && dpt_da_step(v->da.data(), e, DPT_MARK1) &&
                  dpt_da_step_idx(v->da.data(), e, DPT_MARK2, slot);
        v->marker_entry = ok ? e : 0u;
        v->marker_slot = ok ? slot : 0u;
        for (int k = 0; k < 4; ++k) v->ascii_single[k] = 0;
        int n_single = 0;
        for (uint32_t c = 0; c < 256; ++c) {
            uint32_t e1 = DPT_DA_ROOT_ENTRY;
            if (dpt_da_step(v->da.data(), e1, c) && (e1 & DPT_DA_TERMINAL)) {
                if (c < 128) v->ascii_single[c >> 5] |= 1u << (c & 31);
                ++n_single;
            }
        }
        v->fast_ok = v->unit_mode == 0 ? (n_single == 256) : (ok && (e & DPT_DA_TERMINAL));
        v->byte_fallback = 1;
        for (int b = 0; b < 256; ++b) {
            char lit[8];
            snprintf(lit, sizeof lit, "<0x%02X>", b);
            uint32_t e2 = DPT_DA_ROOT_ENTRY, slot2 = 0;
            bool hit = true;
            for (int p = 0; p < 6 && hit; ++p) hit = dpt_da_step_idx(v->da.data(), e2, (uint8_t)lit[p], slot2);
            v->byte_token_id[b] = (hit && (e2 & DPT_DA_TERMINAL)) ? v->slot_id[slot2] : -1;
            if (v->byte_token_id[b] < 0) v->byte_fallback = 0;
        }
        v->marker_leading_only = 1;
        for (int32_t r = 0; r < v->n_tokens && v->marker_leading_only; ++r) {
            const uint8_t* s = v->tok_bytes.data() + v->tok_offs[r];
            const int32_t len = (int32_t)(v->tok_offs[r + 1] - v->tok_offs[r]);
            bool seen_other = false;
            for (int32_t p = 0; p < len;) {
                const bool mark = p + 2 < len && s[p] == DPT_MARK0 && s[p + 1] == DPT_MARK1 && s[p + 2] == DPT_MARK2;
                if (mark && seen_other) {
                    v->marker_leading_only = 0;
                    break;
                }
                if (!mark) seen_other = true;
                p += mark ? 3 : 1;
            }
        }
    }
    rebuild_host_view();
    // the word "<s>" (tokenizer_utils.py:26-30: '<s>' is a word of its own), solved with the same DP
    {
        const uint8_t bos[3] = {'<', 's', '>'};
        uint64_t best[4];
        uint16_t A[4], B[4];
        dpt_forward<true>(h_view, bos, 3, nullptr, best, A, B);
        bos_len = (int32_t)dpt_key_len(best[3]);
        bos_ntok = 0;
        bos_ids[0] = bos_ids[1] = bos_ids[2] = 0;
        if (dpt_key_reach(best[3]) && dpt_backward_emit(h_view, bos, 3, best, A, B, bos_ids, 3)) bos_ntok = bos_len;
    }
    rebuild_host_view();
}

int dpt_vocab_build(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n_in,
                    int32_t unit_mode, dpt_vocab** out, std::string& err) {
    if (!bytes || !offs || !ids || n_in <= 0 || !out) {
        err = "dpt_vocab_create: null argument or empty vocabulary";
        return 1;
    }
    if (unit_mode != 0 && unit_mode != 1) {
        err = "dpt_vocab_create: unit_mode must be 0 (bytes) or 1 (code points)";
        return 1;
    }
    // ---- collect distinct non-empty tokens, dense rank = ascending id (BPE merge order)
    std::vector<int32_t> order;
    order.reserve(n_in);
    for (int32_t k = 0; k < n_in; ++k) {
        if (offs[k + 1] < offs[k] || ids[k] < 0) {
            err = "dpt_vocab_create: offsets must be non-decreasing and ids >= 0";
            return 1;
        }
        if (offs[k + 1] - offs[k] > 65535) {
            err = "dpt_vocab_create: token longer than 65535 bytes";
            return 1;
        }
        if (offs[k + 1] > offs[k]) order.push_back(k);
    }
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ids[a] < ids[b]; });
    if (order.empty()) {
        err = "dpt_vocab_create: vocabulary has no non-empty token";
        return 1;
    }

    std::vector<Node> nodes(1);
    auto* v = new dpt_vocab();
    v->unit_mode = unit_mode;
    v->tok_offs.push_back(0);
    int32_t max_id = 0;
    for (int32_t k : order) {
        const uint8_t* s = bytes + offs[k];
        const int32_t len = (int32_t)(offs[k + 1] - offs[k]);
        int32_t cur = 0;
        for (int32_t p = 0; p < len; ++p) {
            int nx = child_of(nodes, cur, s[p]);
            if (nx < 0) {
                nx = (int)nodes.size();
                nodes.emplace_back();
                nodes[nx].label = s[p];
                nodes[cur].kids.emplace_back(s[p], nx);
            }
            cur = nx;
        }
        if (nodes[cur].token >= 0) continue;  // duplicate byte string: first (lowest id) wins
        nodes[cur].token = (int32_t)v->tok_ids.size();
        v->tok_ids.push_back(ids[k]);
        v->tok_bytes.insert(v->tok_bytes.end(), s, s + len);
        v->tok_offs.push_back((int64_t)v->tok_bytes.size());
        v->lmax = std::max<uint32_t>(v->lmax, (uint32_t)len);
        max_id = std::max(max_id, ids[k]);
    }
    v->n_tokens = (int32_t)v->tok_ids.size();
    v->n_nodes = (int32_t)nodes.size();
    v->id_space = max_id + 1;
    v->id_rank.assign((size_t)v->id_space, -1);
    for (int32_t r = 0; r < v->n_tokens; ++r)
        if (v->id_rank[v->tok_ids[r]] < 0) v->id_rank[v->tok_ids[r]] = r;
    for (auto& n : nodes) std::sort(n.kids.begin(), n.kids.end());

    // ---- hotness: how often a walk from an arbitrary text position visits each node.
    // Every substring of a frequent token is a frequent text substring; token frequency is
    // approximated by a Zipf weight on the id rank (BPE ids are merge order).
    for (int32_t r = 0; r < v->n_tokens; ++r) {
        const uint8_t* s = v->tok_bytes.data() + v->tok_offs[r];
        const int32_t len = (int32_t)(v->tok_offs[r + 1] - v->tok_offs[r]);
        const float w = 1.0f / (float)(r + 8);
        for (int32_t k = 0; k < len; ++k) {
            int32_t cur = 0;
            for (int32_t p = k; p < len; ++p) {
                int nx = child_of(nodes, cur, s[p]);
                if (nx < 0) break;
                nodes[nx].hot += w;
                cur = nx;
            }
        }
    }

    // ---- double-array placement: root first, then internal nodes by descending hotness so the
    // children of hot nodes land in the low slots (the part the kernels stage in shared memory).
    std::vector<int32_t> internal;
    for (int32_t n = 1; n < (int32_t)nodes.size(); ++n)
        if (!nodes[n].kids.empty()) internal.push_back(n);
    std::stable_sort(internal.begin(), internal.end(),
                     [&](int32_t a, int32_t b) { return nodes[a].hot > nodes[b].hot; });
    FreeList fl;
    fl.grow(nodes.size() * 2 + 1024);
    fl.take(0);
    std::vector<uint8_t> base_used(nodes.size() * 2 + 1024, 0);
    uint32_t max_slot = 0;
    auto place = [&](int32_t n, uint32_t forced_base) -> bool {
        Node& nd = nodes[n];
        const uint32_t c0 = nd.kids.front().first;
        uint32_t f = forced_base ? forced_base + c0 : fl.find(1 + c0);
        for (;;) {
            const uint32_t b = f - c0;
            if (b + 256 >= DPT_DA_MAX_SLOTS) return false;
            if (b + 256 >= base_used.size()) base_used.resize((size_t)b + 4096, 0);
            bool ok = b >= 1 && !base_used[b];
            if (ok)
                for (auto& kv : nd.kids)
                    if (!fl.is_free(b + kv.first)) {
                        ok = false;
                        break;
                    }
            if (ok) {
                base_used[b] = 1;
                nd.base = b;
                for (auto& kv : nd.kids) {
                    fl.take(b + kv.first);
                    nodes[kv.second].slot = b + kv.first;
                    max_slot = std::max(max_slot, b + kv.first);
                }
                return true;
            }
            if (forced_base) return false;
            f = fl.find(f + 1);
        }
    };
    bool placed = nodes[0].kids.empty() ? true : place(0, 1);
    base_used[1] = 1;  // the root's base even when the root has no child
    max_slot = std::max(max_slot, 256u);
    for (size_t k = 0; placed && k < internal.size(); ++k) placed = place(internal[k], 0);
    if (!placed) {
        delete v;
        err = "dpt_vocab_create: trie does not fit the 22-bit double array";
        return 1;
    }
    const uint32_t n_slots = max_slot + 1 + 256;  // pad so base+c never leaves the array
    v->da.assign(n_slots, 0u);
    v->slot_id.assign(n_slots, -1);
    for (int32_t n = 1; n < (int32_t)nodes.size(); ++n) {
        const Node& nd = nodes[n];
        uint32_t e = (nd.base << DPT_DA_BASE_SHIFT) | DPT_DA_OCCUPIED | nd.label;
        if (nd.token >= 0) {
            e |= DPT_DA_TERMINAL;
            v->slot_id[nd.slot] = v->tok_ids[nd.token];
        }
        v->da[nd.slot] = e;
    }

    // ---- perfect hash (compress, hash, displace)
    const uint32_t n = (uint32_t)v->n_tokens;
    const uint32_t nb = next_pow2(std::max<uint32_t>(1, (n + 3) / 4));
    const uint32_t ns = next_pow2((uint64_t)n + n / 3 + 8);
    bool done = false;
    for (uint32_t salt = 1; salt < 64 && !done; ++salt) {
        std::vector<DptHashState> hs(n);
        std::vector<std::vector<uint32_t>> buckets(nb);
        for (uint32_t r = 0; r < n; ++r) {
            DptHashState s = dpt_hash_init(salt);
            for (int64_t p = v->tok_offs[r]; p < v->tok_offs[r + 1]; ++p) dpt_hash_byte(s, v->tok_bytes[p]);
            hs[r] = s;
            buckets[dpt_ph_bucket(s, nb - 1)].push_back(r);
        }
        std::vector<uint32_t> bo(nb);
        std::iota(bo.begin(), bo.end(), 0u);
        std::stable_sort(bo.begin(), bo.end(), [&](uint32_t a, uint32_t b) { return buckets[a].size() > buckets[b].size(); });
        std::vector<uint32_t> seed(nb, 0);
        std::vector<int32_t> table(ns, -1);
        bool fail = false;
        std::vector<uint32_t> tmp;
        for (uint32_t bi : bo) {
            auto& bk = buckets[bi];
            if (bk.empty()) break;
            bool found = false;
            for (uint32_t d = 0; d < (1u << 20) && !found; ++d) {
                tmp.clear();
                bool ok = true;
                for (uint32_t r : bk) {
                    const uint32_t sl = dpt_ph_slot(hs[r], d, ns - 1);
                    if (table[sl] >= 0 || std::find(tmp.begin(), tmp.end(), sl) != tmp.end()) {
                        ok = false;
                        break;
                    }
                    tmp.push_back(sl);
                }
                if (ok) {
                    for (size_t k = 0; k < bk.size(); ++k) table[tmp[k]] = v->tok_ids[bk[k]];
                    seed[bi] = d;
                    found = true;
                }
            }
            if (!found) {
                fail = true;
                break;
            }
        }
        if (!fail) {
            v->ph_seed.swap(seed);
            v->ph_id.swap(table);
            v->ph_salt = salt;
            done = true;
        }
    }
    if (!done) {
        delete v;
        err = "dpt_vocab_create: perfect hash construction failed";
        return 1;
    }
    v->rebuild_host_view();
    v->derive_facts();
    *out = v;
    return 0;
}
