// extern "C" boundary (include/dptok.h).  Thin: argument checks, workspace carving, stream plumbing.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/dptok.h"
#include "kernels.h"
#include "vocab.h"
#include "dpt_unicode_tables.h"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct BlobLayout {
    int64_t da, slot_id, ph_seed, ph_id, tok_bytes, tok_offs, id_rank, uni1, uni2, merge_keys, merge_vals, byte_ids, code2, total;
};
BlobLayout layout_of(const dpt_vocab* v) {
    BlobLayout L{};
    int64_t o = 0;
    L.da = o;
    o = align_up(o + (int64_t)v->da.size() * 4, 256);
    L.slot_id = o;
    o = align_up(o + (int64_t)v->slot_id.size() * 4, 256);
    L.ph_seed = o;
    o = align_up(o + (int64_t)v->ph_seed.size() * 4, 256);
    L.ph_id = o;
    o = align_up(o + (int64_t)v->ph_id.size() * 4, 256);
    L.tok_bytes = o;
    o = align_up(o + (int64_t)v->tok_bytes.size() + 16, 256);
    L.tok_offs = o;
    o = align_up(o + (int64_t)v->tok_offs.size() * 8, 256);
    L.id_rank = o;
    o = align_up(o + (int64_t)v->id_rank.size() * 4, 256);
    L.uni1 = o;
    o = align_up(o + DPT_UNI_STAGE1_LEN, 256);
    L.uni2 = o;
    o = align_up(o + DPT_UNI_STAGE2_LEN, 256);
    L.merge_keys = o;
    o = align_up(o + (int64_t)v->merge_keys.size() * 8, 256);
    L.merge_vals = o;
    o = align_up(o + (int64_t)v->merge_vals.size() * 8, 256);
    L.byte_ids = o;
    o = align_up(o + 256 * 4, 256);
    L.code2 = o;
    o = align_up(o + 2048, 256);
    L.total = o;
    return L;
}
const uint32_t kSerialMagic = 0x31545044u;  // "DPT1"
}  // namespace

// ---- serialisation: header + raw vectors --------------------------------------------------
namespace {
template <typename T>
void put_vec(std::vector<uint8_t>& o, const std::vector<T>& v) {
    const uint64_t n = v.size();
    const uint8_t* p = (const uint8_t*)&n;
    o.insert(o.end(), p, p + 8);
    const uint8_t* q = (const uint8_t*)v.data();
    o.insert(o.end(), q, q + n * sizeof(T));
}
template <typename T>
bool get_vec(const uint8_t*& p, const uint8_t* end, std::vector<T>& v) {
    if (end - p < 8) return false;
    uint64_t n;
    std::memcpy(&n, p, 8);
    p += 8;
    if (n > (uint64_t)(end - p) / sizeof(T)) return false;  // (no n * sizeof(T): a corrupt count must not wrap)
    v.resize(n);
    std::memcpy(v.data(), p, n * sizeof(T));
    p += n * sizeof(T);
    return true;
}
struct SerialHeader {
    uint32_t magic;
    int32_t unit_mode, n_tokens, n_nodes, id_space;
    uint32_t lmax, ph_salt, marker_entry, ascii_single[4];
    int32_t marker_leading_only, byte_fallback;
    int32_t byte_token_id[256];
};
}  // namespace


extern "C" {

const char* dpt_last_error(void) { return g_err.c_str(); }
const char* dpt_version(void) { return "dptok-b200 0.1 (sm_100a)"; }
int64_t dpt_launch_count(void) { return dpt::g_launches.load(); }
void dpt_profile_enable(int32_t on) { dpt::profile_enable(on); }
int dpt_profile_report(char* buf, int64_t cap, int64_t* need) {
    static thread_local std::string pending;
    if (!need) return fail(DPT_EINVAL, "dpt_profile_report: null argument");
    if (pending.empty()) pending = dpt::profile_report();
    *need = (int64_t)pending.size() + 1;
    if (!buf) return DPT_OK;
    if (cap < *need) return fail(DPT_ECAPACITY, "dpt_profile_report: buffer too small");
    std::memcpy(buf, pending.c_str(), pending.size() + 1);
    pending.clear();
    return DPT_OK;
}

int dpt_vocab_create(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n_tokens, int32_t unit_mode,
                     dpt_vocab** out) {
    std::string err;
    const int rc = dpt_vocab_build(bytes, offs, ids, n_tokens, unit_mode, out, err);
    return rc ? fail(DPT_EINVAL, err) : DPT_OK;
}

void dpt_vocab_destroy(dpt_vocab* v) {
    if (!v) return;
    if (v->d_blob) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(v->device);
        cudaFree(v->d_blob);
        cudaSetDevice(cur);
    }
    delete v;
}

int dpt_vocab_get_info(const dpt_vocab* v, dpt_vocab_info* out) {
    if (!v || !out) return fail(DPT_EINVAL, "dpt_vocab_get_info: null argument");
    out->n_tokens = v->n_tokens;
    out->unit_mode = v->unit_mode;
    out->n_nodes = v->n_nodes;
    out->n_slots = (int32_t)v->da.size();
    out->max_token_bytes = (int32_t)v->lmax;
    out->ph_buckets = (int32_t)v->ph_seed.size();
    out->ph_slots = (int32_t)v->ph_id.size();
    out->marker_leading_only = v->marker_leading_only;
    out->byte_fallback = v->byte_fallback;
    out->device = v->device;
    out->blob_bytes = layout_of(v).total;
    return DPT_OK;
}

int dpt_vocab_lookup(const dpt_vocab* v, const uint8_t* s, int32_t len, int32_t* id_out) {
    if (!v || !id_out || (len > 0 && !s)) return fail(DPT_EINVAL, "dpt_vocab_lookup: null argument");
    *id_out = -1;
    if (len <= 0) return DPT_OK;
    DptHashState h = dpt_hash_init(v->ph_salt);
    for (int32_t k = 0; k < len; ++k) dpt_hash_byte(h, s[k]);
    const int32_t id = dpt_ph_lookup(v->h_view, h);
    if (id < 0 || id >= v->id_space) return DPT_OK;
    const int32_t r = v->id_rank[id];
    if (r < 0) return DPT_OK;
    const int64_t a = v->tok_offs[r], b = v->tok_offs[r + 1];
    if (b - a == len && std::memcmp(v->tok_bytes.data() + a, s, (size_t)len) == 0) *id_out = id;
    return DPT_OK;
}

int dpt_vocab_serialize(const dpt_vocab* v, uint8_t* buf, int64_t cap, int64_t* need) {
    if (!v || !need) return fail(DPT_EINVAL, "dpt_vocab_serialize: null argument");
    std::vector<uint8_t> o;
    SerialHeader h{};
    h.magic = kSerialMagic;
    h.unit_mode = v->unit_mode;
    h.n_tokens = v->n_tokens;
    h.n_nodes = v->n_nodes;
    h.id_space = v->id_space;
    h.lmax = v->lmax;
    h.ph_salt = v->ph_salt;
    h.marker_entry = v->marker_entry;
    std::memcpy(h.ascii_single, v->ascii_single, sizeof h.ascii_single);
    h.marker_leading_only = v->marker_leading_only;
    h.byte_fallback = v->byte_fallback;
    std::memcpy(h.byte_token_id, v->byte_token_id, sizeof h.byte_token_id);
    o.insert(o.end(), (uint8_t*)&h, (uint8_t*)&h + sizeof h);
    put_vec(o, v->da);
    put_vec(o, v->slot_id);
    put_vec(o, v->ph_seed);
    put_vec(o, v->ph_id);
    put_vec(o, v->tok_bytes);
    put_vec(o, v->tok_offs);
    put_vec(o, v->tok_ids);
    put_vec(o, v->id_rank);
    *need = (int64_t)o.size();
    if (!buf) return DPT_OK;
    if (cap < (int64_t)o.size()) return fail(DPT_ECAPACITY, "dpt_vocab_serialize: buffer too small");
    std::memcpy(buf, o.data(), o.size());
    return DPT_OK;
}

int dpt_vocab_deserialize(const uint8_t* buf, int64_t len, dpt_vocab** out) {
    if (!buf || !out || len < (int64_t)sizeof(SerialHeader)) return fail(DPT_EINVAL, "dpt_vocab_deserialize: bad buffer");
    SerialHeader h;
    std::memcpy(&h, buf, sizeof h);
    if (h.magic != kSerialMagic) return fail(DPT_EINVAL, "dpt_vocab_deserialize: bad magic");
    auto* v = new dpt_vocab();
    v->unit_mode = h.unit_mode;
    v->n_tokens = h.n_tokens;
    v->n_nodes = h.n_nodes;
    v->id_space = h.id_space;
    v->lmax = h.lmax;
    v->ph_salt = h.ph_salt;
    v->marker_entry = h.marker_entry;
    std::memcpy(v->ascii_single, h.ascii_single, sizeof h.ascii_single);
    v->marker_leading_only = h.marker_leading_only;
    v->byte_fallback = h.byte_fallback;
    std::memcpy(v->byte_token_id, h.byte_token_id, sizeof h.byte_token_id);
    const uint8_t* p = buf + sizeof h;
    const uint8_t* end = buf + len;
    bool ok = false;
    try {  // (nothing may throw across the extern "C" boundary: a corrupt count could ask resize() for the moon)
        ok = get_vec(p, end, v->da) && get_vec(p, end, v->slot_id) && get_vec(p, end, v->ph_seed) &&
             get_vec(p, end, v->ph_id) && get_vec(p, end, v->tok_bytes) && get_vec(p, end, v->tok_offs) &&
             get_vec(p, end, v->tok_ids) && get_vec(p, end, v->id_rank);
    } catch (...) {
        ok = false;
    }
    ok = ok && v->da.size() >= 257 && v->da.size() <= (size_t)DPT_DA_MAX_SLOTS + 256 && !v->ph_seed.empty() && !v->ph_id.empty() &&
         !(v->ph_seed.size() & (v->ph_seed.size() - 1)) && !(v->ph_id.size() & (v->ph_id.size() - 1)) &&
         v->slot_id.size() == v->da.size() && v->n_tokens >= 0 && v->tok_offs.size() == (size_t)v->n_tokens + 1 &&
         v->tok_ids.size() == (size_t)v->n_tokens && v->id_space >= 0 && v->id_rank.size() == (size_t)v->id_space;
    if (ok) {
        // every trie base must leave its 256 children inside the array (the kernels index da[base + byte] unchecked), every
        // token's bytes inside tok_bytes, every id inside the id space
        const size_t nda = v->da.size();
        for (size_t k = 0; ok && k < nda; ++k) ok = (size_t)(v->da[k] >> DPT_DA_BASE_SHIFT) + 256 <= nda;
        ok = ok && (size_t)(v->marker_entry >> DPT_DA_BASE_SHIFT) + 256 <= nda;
        for (size_t k = 0; ok && k < nda; ++k) ok = v->slot_id[k] < v->id_space;
        for (int32_t k = 0; ok && k < v->n_tokens; ++k)
            ok = v->tok_offs[k] >= 0 && v->tok_offs[k] <= v->tok_offs[k + 1] && v->tok_ids[k] >= 0 && v->tok_ids[k] < v->id_space;
        ok = ok && (v->n_tokens == 0 || (size_t)v->tok_offs[v->n_tokens] <= v->tok_bytes.size());
        for (size_t k = 0; ok && k < v->ph_id.size(); ++k) ok = v->ph_id[k] < v->n_tokens;
    }
    if (!ok) {
        delete v;
        return fail(DPT_EINVAL, "dpt_vocab_deserialize: truncated, corrupt or inconsistent buffer");
    }
    v->derive_facts();
    *out = v;
    return DPT_OK;
}

int dpt_vocab_upload(dpt_vocab* v, int device) {
    if (!v) return fail(DPT_EINVAL, "dpt_vocab_upload: null handle");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return fail(DPT_ECUDA, "dpt_vocab_upload: no CUDA device (this library has no CPU execution path)");
    if (device < 0 || device >= count) return fail(DPT_EINVAL, "dpt_vocab_upload: bad device index");
    if (v->d_blob && v->device == device) return DPT_OK;
    if (v->d_blob) return fail(DPT_ESTATE, "dpt_vocab_upload: handle already uploaded to another device");
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    const BlobLayout L = layout_of(v);
    char* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, (size_t)L.total);
    if (e != cudaSuccess) {
        cudaSetDevice(cur);
        return fail(DPT_ENOMEM, std::string("dpt_vocab_upload: cudaMalloc: ") + cudaGetErrorString(e));
    }
    cudaMemset(d, 0, (size_t)L.total);
    cudaMemcpy(d + L.da, v->da.data(), v->da.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.slot_id, v->slot_id.data(), v->slot_id.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.ph_seed, v->ph_seed.data(), v->ph_seed.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.ph_id, v->ph_id.data(), v->ph_id.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.tok_bytes, v->tok_bytes.data(), v->tok_bytes.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.tok_offs, v->tok_offs.data(), v->tok_offs.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.uni1, DPT_UNI_STAGE1, DPT_UNI_STAGE1_LEN, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.uni2, DPT_UNI_STAGE2, DPT_UNI_STAGE2_LEN, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.byte_ids, v->byte_token_id, 256 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + L.code2, v->code2, 2048, cudaMemcpyHostToDevice);
    if (!v->merge_keys.empty()) {
        cudaMemcpy(d + L.merge_keys, v->merge_keys.data(), v->merge_keys.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d + L.merge_vals, v->merge_vals.data(), v->merge_vals.size() * 8, cudaMemcpyHostToDevice);
    }
    e = cudaMemcpy(d + L.id_rank, v->id_rank.data(), v->id_rank.size() * 4, cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    cudaSetDevice(cur);
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(DPT_ECUDA, std::string("dpt_vocab_upload: ") + cudaGetErrorString(e));
    }
    v->d_blob = d;
    v->device = device;
    v->blob_bytes = L.total;
    v->d_view = v->h_view;
    v->d_view.da = (const uint32_t*)(d + L.da);
    v->d_view.slot_id = (const int32_t*)(d + L.slot_id);
    v->d_view.ph_seed = (const uint32_t*)(d + L.ph_seed);
    v->d_view.ph_id = (const int32_t*)(d + L.ph_id);
    v->d_view.tok_bytes = (const uint8_t*)(d + L.tok_bytes);
    v->d_view.tok_offs = (const int64_t*)(d + L.tok_offs);
    v->d_view.id_rank = (const int32_t*)(d + L.id_rank);
    v->d_view.uni1 = (const uint8_t*)(d + L.uni1);
    v->d_view.uni2 = (const uint8_t*)(d + L.uni2);
    v->d_view.merge_keys = v->merge_keys.empty() ? nullptr : (const unsigned long long*)(d + L.merge_keys);
    v->d_view.merge_vals = v->merge_vals.empty() ? nullptr : (const unsigned long long*)(d + L.merge_vals);
    v->d_view.byte_ids = (const int32_t*)(d + L.byte_ids);
    v->d_view.code2 = (const uint8_t*)(d + L.code2);
    return DPT_OK;
}

int dpt_vocab_set_merges(dpt_vocab* v, const int32_t* left, const int32_t* right, const int32_t* merged, int32_t n_merges) {
    if (!v || n_merges < 0 || (n_merges > 0 && (!left || !right || !merged)))
        return fail(DPT_EINVAL, "dpt_vocab_set_merges: bad argument");
    if (v->d_blob) return fail(DPT_ESTATE, "dpt_vocab_set_merges: call before dpt_vocab_upload");
    for (int32_t k = 0; k < n_merges; ++k)
        if (left[k] < 0 || right[k] < 0 || merged[k] < 0 || left[k] >= v->id_space || right[k] >= v->id_space ||
            merged[k] >= v->id_space)
            return fail(DPT_EINVAL, "dpt_vocab_set_merges: token id outside the vocabulary's id space");
    try {
        if (v->set_merges(left, right, merged, n_merges)) return fail(DPT_EINVAL, "dpt_vocab_set_merges: bad merge");
    } catch (...) {
        return fail(DPT_ENOMEM, "dpt_vocab_set_merges: out of memory");
    }
    return DPT_OK;
}

static int check_ready(const dpt_vocab* v, const char* who) {
    if (!v) return fail(DPT_EINVAL, std::string(who) + ": null vocab handle");
    if (!v->d_blob) return fail(DPT_ESTATE, std::string(who) + ": vocab not uploaded (dpt_vocab_upload)");
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return fail(DPT_ECUDA, std::string(who) + ": no CUDA device");
    if (cur != v->device) return fail(DPT_ESTATE, std::string(who) + ": current device differs from the vocab's device");
    return DPT_OK;
}

int64_t dpt_pretokenize_workspace(int64_t n_bytes, int64_t n_docs) { return dpt::pretokenize_workspace(n_bytes, n_docs); }

int64_t dpt_encode_words_workspace(int64_t n_bytes, int64_t n_words, int32_t worst_case) {
    const int64_t pool = worst_case ? 12 * (n_bytes + n_words) : 12 * (n_bytes / 8 + 65536);
    return dpt::encode_words_workspace_fixed(n_words, n_bytes) + pool + 4096;
}

int dpt_pretokenize(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs,
                    int64_t n_docs, uint8_t* d_norm_text, int64_t norm_cap, int64_t* d_norm_doc_offs,
                    int64_t* d_word_offs, int64_t word_cap, int64_t* d_doc_first_word, uint8_t* d_doc_flags,
                    int64_t* d_n_out, void* d_workspace, int64_t workspace_bytes, void* stream) {
    if (int rc = check_ready(v, "dpt_pretokenize")) return rc;
    std::string err;
    int rc;
    switch (rule) {
        case DPT_RULE_SPM_LLAMA:
            rc = dpt::pretokenize_spm(v, d_text, n_bytes, d_doc_offs, n_docs, d_norm_text, norm_cap, d_norm_doc_offs,
                                      d_word_offs, word_cap, d_doc_first_word, d_doc_flags, d_n_out, d_workspace,
                                      workspace_bytes, (cudaStream_t)stream, err);
            break;
        default:
            return fail(DPT_EINVAL, "dpt_pretokenize: rule not available on device in this build; pre-split on the host "
                                    "(DPT_RULE_PRESPLIT)");
    }
    return rc ? fail(rc, err) : DPT_OK;
}

int dpt_encode_words(const dpt_vocab* v, const uint8_t* d_text, const int64_t* d_word_offs, int64_t n_words,
                     int64_t n_text_bytes, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                     int64_t* d_word_tok_offs, int64_t* d_counters, int64_t* d_n_out, void* d_workspace,
                     int64_t workspace_bytes, void* stream) {
    if (int rc = check_ready(v, "dpt_encode_words")) return rc;
    std::string err;
    const int rc = dpt::encode_words(v, d_text, d_word_offs, n_words, n_text_bytes, n_text_bytes, d_ids, ids_cap, d_word_lens,
                                     d_word_flags, d_word_tok_offs, d_counters, d_n_out, d_workspace, workspace_bytes,
                                     (cudaStream_t)stream, err);
    return rc ? fail(rc, err) : DPT_OK;
}

// ---- fused corpus entry (general path): normalise -> DP -> compaction ------------------------
static int64_t corpus_norm_cap(int64_t n_bytes, int64_t n_docs, int worst) {
    return worst ? 6 * n_bytes + 6 * n_docs + 64 : n_bytes + n_bytes / 8 + 6 * n_docs + 4096;
}

int64_t dpt_encode_corpus_general_workspace(int32_t rule, int64_t n_bytes, int64_t n_docs, int64_t word_cap, int32_t worst_case) {
    (void)rule;
    int64_t b = 0;
    const int64_t norm_cap = corpus_norm_cap(n_bytes, n_docs, worst_case);
    b += align_up(norm_cap, 256);
    b += align_up((n_docs + 1) * 8, 256) * 2;  // norm_doc_offs, doc_first_word
    b += align_up((word_cap + 1) * 8, 256);    // word_offs
    b += align_up((word_cap + 1) * 8, 256);    // tok_offs
    b += align_up(64, 256);
    b += dpt::pretokenize_workspace(n_bytes, n_docs);
    b += dpt_encode_words_workspace(norm_cap, word_cap, worst_case);
    return b + 4096;
}

int dpt_encode_corpus_general(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes,
                              const int64_t* d_doc_offs, int64_t n_docs, int32_t* d_ids, int64_t ids_cap,
                              int32_t* d_word_lens, uint8_t* d_word_flags, int64_t word_cap, int64_t* d_doc_tok_offs,
                              uint8_t* d_doc_flags, int64_t* d_counters, int64_t* d_n_out, void* d_workspace,
                              int64_t workspace_bytes, int32_t worst_case, void* stream) {
    if (int rc = check_ready(v, "dpt_encode_corpus_general")) return rc;
    if (rule != DPT_RULE_SPM_LLAMA)
        return fail(DPT_EINVAL, "dpt_encode_corpus_general: only DPT_RULE_SPM_LLAMA runs its boundary rule on device in this "
                                "build; use dpt_encode_words with host pre-split words");
    if (!d_doc_tok_offs || !d_n_out || word_cap <= 0) return fail(DPT_EINVAL, "dpt_encode_corpus: bad argument");
    if (workspace_bytes < dpt_encode_corpus_general_workspace(rule, n_bytes, n_docs, word_cap, worst_case))
        return fail(DPT_ECAPACITY, "dpt_encode_corpus: workspace too small (see dpt_encode_corpus_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)d_workspace;
    int64_t used = 0;
    auto take = [&](int64_t bytes) {
        used = align_up(used, 256);
        char* p = base + used;
        used += bytes;
        return p;
    };
    const int64_t norm_cap = corpus_norm_cap(n_bytes, n_docs, worst_case);
    uint8_t* norm = (uint8_t*)take(norm_cap);
    int64_t* norm_doc_offs = (int64_t*)take((n_docs + 1) * 8);
    int64_t* first_word = (int64_t*)take((n_docs + 1) * 8);
    int64_t* word_offs = (int64_t*)take((word_cap + 1) * 8);
    int64_t* tok_offs = (int64_t*)take((word_cap + 1) * 8);
    int64_t* pre_out = (int64_t*)take(64);
    const int64_t pws = dpt::pretokenize_workspace(n_bytes, n_docs);
    void* pre_ws = take(pws);
    used = align_up(used, 256);
    void* enc_ws = base + used;
    const int64_t enc_ws_bytes = workspace_bytes - used;

    std::string err;
    int rc = dpt::pretokenize_spm(v, d_text, n_bytes, d_doc_offs, n_docs, norm, norm_cap, norm_doc_offs, word_offs, word_cap,
                                  first_word, d_doc_flags, pre_out, pre_ws, pws, st, err);
    if (rc) return fail(rc, err);
    // the word count sizes the DP grids: one 16-byte readback (the fused tile kernel has no such sync)
    int64_t h_pre[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h_pre, pre_out, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(DPT_ECUDA, std::string("dpt_encode_corpus: ") + cudaGetErrorString(e));
    const int64_t n_words = h_pre[0], n_norm = h_pre[1];
    if (n_words > word_cap || n_norm > norm_cap) {
        int64_t h_out[8] = {0, n_words, 0, 0, n_norm, norm_cap, 0, 0};
        cudaMemcpyAsync(d_n_out, h_out, sizeof h_out, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        return fail(DPT_ECAPACITY, "dpt_encode_corpus: word or normalised-text capacity exceeded (see d_n_out)");
    }
    rc = dpt::encode_words(v, norm, word_offs, n_words, n_bytes, n_norm, d_ids, ids_cap, d_word_lens, d_word_flags, tok_offs,
                           d_counters, d_n_out, enc_ws, enc_ws_bytes, st, err);
    if (rc) return fail(rc, err);
    int64_t h_tail[2] = {n_norm, norm_cap};
    cudaMemcpyAsync(d_n_out + 4, h_tail, 16, cudaMemcpyHostToDevice, st);
    rc = dpt::doc_tok_offsets(first_word, n_docs, tok_offs, n_words, d_doc_tok_offs, st);
    if (rc) return fail(rc, "dpt_encode_corpus: doc offsets kernel failed");
    return DPT_OK;
}

// ---- corpus pipeline entry (pipe.cu): asynchronous, no host synchronisation ---------------------------------
int64_t dpt_encode_corpus_workspace(int32_t rule, int64_t n_bytes, int64_t n_docs, int64_t word_cap, int32_t worst_case) {
    (void)rule;
    return dpt::encode_corpus_pipe_workspace(n_bytes, n_docs, word_cap, worst_case);
}

int dpt_encode_corpus(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes, const int64_t* d_doc_offs,
                      int64_t n_docs, int32_t* d_ids, int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags,
                      int64_t word_cap, int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters,
                      int64_t* d_n_out, void* d_workspace, int64_t workspace_bytes, int32_t worst_case, void* stream) {
    if (int rc = check_ready(v, "dpt_encode_corpus")) return rc;
    if (rule != DPT_RULE_SPM_LLAMA && rule != DPT_RULE_GPT2 && rule != DPT_RULE_LLAMA3 && rule != DPT_RULE_BLOOM)
        return fail(DPT_EINVAL, "dpt_encode_corpus: unknown rule (SPM_LLAMA, GPT2, LLAMA3, BLOOM run on the device; "
                                "pre-split anything else on the host and call dpt_encode_words)");
    std::string err;
    const int rc = dpt::encode_corpus_pipe(v, rule, d_text, n_bytes, d_doc_offs, n_docs, d_ids, ids_cap, d_word_lens,
                                           d_word_flags, word_cap, d_doc_tok_offs, d_doc_flags, d_counters, d_n_out,
                                           d_workspace, workspace_bytes, worst_case, (cudaStream_t)stream, err);
    return rc ? fail(rc, err) : DPT_OK;
}

// ---- chunked calls: one word table for all ranges of a corpus that is resident (or arriving) in one buffer ----------
int64_t dpt_corpus_table_workspace(int64_t n_bytes_total, int64_t word_cap_total, int32_t worst_case) {
    return dpt::corpus_table_workspace(n_bytes_total, word_cap_total, worst_case);
}
int64_t dpt_encode_corpus_range_workspace(int32_t rule, int64_t range_bytes, int64_t range_docs, int64_t word_cap,
                                          int32_t worst_case) {
    (void)rule;
    return dpt::corpus_range_workspace(range_bytes, range_docs, word_cap, worst_case);
}
int dpt_encode_corpus_range(const dpt_vocab* v, int32_t rule, const uint8_t* d_text, int64_t n_bytes_total,
                            const int64_t* d_doc_offs, int64_t n_docs_total, int64_t byte_begin, int64_t byte_end,
                            int64_t doc_begin, int64_t doc_end, int32_t reset_table, int64_t table_word_cap, int32_t* d_ids,
                            int64_t ids_cap, int32_t* d_word_lens, uint8_t* d_word_flags, int64_t word_cap,
                            int64_t* d_doc_tok_offs, uint8_t* d_doc_flags, int64_t* d_counters, int64_t* d_n_out,
                            void* d_table_workspace, int64_t table_workspace_bytes, void* d_workspace,
                            int64_t workspace_bytes, int32_t worst_case, int32_t phases, void* stream) {
    if (int rc = check_ready(v, "dpt_encode_corpus_range")) return rc;
    if (phases < 0 || phases > 7) return fail(DPT_EINVAL, "dpt_encode_corpus_range: phases must be 0..7");
    if (rule != DPT_RULE_SPM_LLAMA && rule != DPT_RULE_GPT2 && rule != DPT_RULE_LLAMA3 && rule != DPT_RULE_BLOOM)
        return fail(DPT_EINVAL, "dpt_encode_corpus_range: unknown rule");
    std::string err;
    const int rc = dpt::encode_corpus_range(v, rule, d_text, n_bytes_total, d_doc_offs, n_docs_total, byte_begin, byte_end,
                                            doc_begin, doc_end, reset_table, n_bytes_total, table_word_cap, d_ids, ids_cap,
                                            d_word_lens, d_word_flags, word_cap, d_doc_tok_offs, d_doc_flags, d_counters,
                                            d_n_out, d_table_workspace, table_workspace_bytes, d_workspace, workspace_bytes,
                                            worst_case, phases, (cudaStream_t)stream, err);
    return rc ? fail(rc, err) : DPT_OK;
}

int dpt_lattice_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts,
                     int32_t* d_len_dp, int32_t* d_pred_offs, int32_t* d_pred, int32_t pred_cap, int32_t* d_n_out,
                     int32_t* d_scratch, void* stream) {
    if (int rc = check_ready(v, "dpt_lattice_word")) return rc;
    if (n_bytes <= 0 || !d_text || !d_len_dp || !d_pred_offs || !d_pred || !d_n_out || !d_scratch)
        return fail(DPT_EINVAL, "dpt_lattice_word: bad argument");
    const int rc = dpt::lattice_word(v, d_text, n_bytes, d_unit_starts, d_len_dp, d_pred_offs, d_pred, pred_cap, d_n_out,
                                     d_scratch, (cudaStream_t)stream);
    return rc ? fail(rc, "dpt_lattice_word: launch failed") : DPT_OK;
}

int dpt_min_tokens_word(const dpt_vocab* v, const uint8_t* d_text, int32_t n_bytes, const uint8_t* d_unit_starts,
                        int32_t* d_out, int32_t* d_scratch, void* stream) {
    if (int rc = check_ready(v, "dpt_min_tokens_word")) return rc;
    if (n_bytes <= 0 || !d_text || !d_out || !d_scratch) return fail(DPT_EINVAL, "dpt_min_tokens_word: bad argument");
    const int rc = dpt::min_tokens_word(v, d_text, n_bytes, d_unit_starts, d_out, d_scratch, (cudaStream_t)stream);
    return rc ? fail(rc, "dpt_min_tokens_word: launch failed") : DPT_OK;
}

int dpt_pad_batch(const int32_t* d_ids_a, const int64_t* d_doc_tok_offs_a, const int32_t* d_ids_b,
                  const int64_t* d_doc_tok_offs_b, int64_t doc_begin, int64_t n_rows, int64_t row_len, int64_t pad_id,
                  int32_t pad_left, int64_t* d_input_ids, int64_t* d_attention_mask, int64_t* d_row_lens, void* stream) {
    if (n_rows <= 0 || row_len <= 0 || n_rows >= (1ll << 31) || !d_ids_a || !d_doc_tok_offs_a || !d_input_ids ||
        (d_ids_b && !d_doc_tok_offs_b))
        return fail(DPT_EINVAL, "dpt_pad_batch: bad argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fail(DPT_ECUDA, "dpt_pad_batch: no CUDA device");
    const int rc = dpt::pad_batch(d_ids_a, d_doc_tok_offs_a, d_ids_b, d_doc_tok_offs_b, doc_begin, n_rows, row_len, pad_id,
                                  pad_left, d_input_ids, d_attention_mask, d_row_lens, (cudaStream_t)stream);
    return rc ? fail(rc, "dpt_pad_batch: launch failed") : DPT_OK;
}

int dpt_narrow_ids_u16(const int32_t* d_ids, const int64_t* d_n, int64_t cap, uint16_t* d_out, int64_t* d_overflow,
                       void* stream) {
    if (!d_ids || !d_n || !d_out || cap <= 0 || cap >= (1ll << 40)) return fail(DPT_EINVAL, "dpt_narrow_ids_u16: bad argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fail(DPT_ECUDA, "dpt_narrow_ids_u16: no CUDA device");
    const int rc = dpt::narrow_ids_u16(d_ids, d_n, cap, d_out, d_overflow, (cudaStream_t)stream);
    return rc ? fail(rc, "dpt_narrow_ids_u16: launch failed") : DPT_OK;
}

int dpt_roundtrip_check(const dpt_vocab* v, const int32_t* d_ids, const int64_t* d_doc_tok_offs, const uint8_t* d_text,
                        const int64_t* d_doc_offs, int64_t n_docs, int32_t skip_bos, uint8_t* d_ok, void* stream) {
    if (int rc = check_ready(v, "dpt_roundtrip_check")) return rc;
    if (n_docs <= 0 || !d_ids || !d_doc_tok_offs || !d_text || !d_doc_offs || !d_ok)
        return fail(DPT_EINVAL, "dpt_roundtrip_check: bad argument");
    const int rc = dpt::roundtrip_check(v, d_ids, d_doc_tok_offs, d_text, d_doc_offs, n_docs, skip_bos, d_ok,
                                        (cudaStream_t)stream);
    return rc ? fail(rc, "dpt_roundtrip_check: launch failed") : DPT_OK;
}

int dpt_synth_corpus(const uint8_t* d_lex_a_bytes, const int64_t* d_lex_a_offs, const uint32_t* d_lex_a_cdf, int32_t n_a,
                     const uint8_t* d_lex_b_bytes, const int64_t* d_lex_b_offs, const uint32_t* d_lex_b_cdf, int32_t n_b,
                     const dpt_synth_params* params, int64_t doc_base, int64_t n_docs, int64_t* d_doc_len,
                     const int64_t* d_doc_offs, uint8_t* d_text, void* stream) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fail(DPT_ECUDA, "dpt_synth_corpus: no CUDA device");
    std::string err;
    const int rc = dpt::synth_run(d_lex_a_bytes, d_lex_a_offs, d_lex_a_cdf, n_a, d_lex_b_bytes, d_lex_b_offs, d_lex_b_cdf, n_b,
                                  params, doc_base, n_docs, d_doc_len, d_doc_offs, d_text, (cudaStream_t)stream, err);
    return rc ? fail(rc, err.c_str()) : DPT_OK;
}

}  // extern "C"
