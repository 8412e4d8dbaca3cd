/*
 * C restatement of the reference's shortest-tokenization DP.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Follows /root/reference/packages/dp_tokenize.py:
 *   forward  len_dp[i] = min(i, min_j len_dp[j]+1)                         (:27-47, phantom init :28)
 *   P(i)     = { j : s[j:i] in V and len_dp[j]+1 == len_dp[i] }            (:40-46)
 *   paths that cannot reach 0 are dropped                                  (:63-69)
 *   choose the first segmentation in DFS order (largest split first, :58)
 *   whose longest token (code points, :82) is maximal                      (:84)
 * in the closed form of SURVEY.md 8.1 (pull form, explicit reach/M arrays, greedy backward
 * select) - written independently of the packed-key push form used on the device.
 *
 * Parity status: PINNED by tests/test_oracle.py against tests/golden/ (outputs of the unmodified
 * reference) and against oracle/dp_oracle.py.
 *
 * Build: oracle/build.py  ->  oracle/_build/liboracle.so      (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t* bytes; /* borrowed copy below */
    uint8_t* pool;
    int64_t* offs;
    int32_t* ids;
    int32_t n;
    int32_t unit_mode;
    int32_t lmax;
    /* open addressing: slot -> token index or -1 */
    int32_t* table;
    uint64_t mask;
} orc_vocab;

static uint64_t fnv1a(const uint8_t* s, int64_t n) {
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < n; ++i) {
        h ^= s[i];
        h *= 1099511628211ull;
    }
    return h ^ (h >> 29);
}

static int32_t find(const orc_vocab* v, const uint8_t* s, int64_t n) {
    uint64_t h = fnv1a(s, n) & v->mask;
    for (;;) {
        const int32_t t = v->table[h];
        if (t < 0) return -1;
        const int64_t a = v->offs[t], b = v->offs[t + 1];
        if (b - a == n && memcmp(v->pool + a, s, (size_t)n) == 0) return t;
        h = (h + 1) & v->mask;
    }
}

void* orc_vocab_new(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n, int32_t unit_mode) {
    orc_vocab* v = (orc_vocab*)calloc(1, sizeof(orc_vocab));
    const int64_t total = offs[n];
    v->pool = (uint8_t*)malloc((size_t)total + 1);
    memcpy(v->pool, bytes, (size_t)total);
    v->offs = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n + 1));
    memcpy(v->offs, offs, sizeof(int64_t) * ((size_t)n + 1));
    v->ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(v->ids, ids, sizeof(int32_t) * (size_t)n);
    v->n = n;
    v->unit_mode = unit_mode;
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 3) cap <<= 1;
    v->mask = cap - 1;
    v->table = (int32_t*)malloc(sizeof(int32_t) * cap);
    for (uint64_t i = 0; i < cap; ++i) v->table[i] = -1;
    for (int32_t t = 0; t < n; ++t) {
        const int64_t len = offs[t + 1] - offs[t];
        if (len <= 0) continue;
        if (len > v->lmax) v->lmax = (int32_t)len;
        if (find(v, v->pool + offs[t], len) >= 0) continue; /* duplicate: first wins */
        uint64_t h = fnv1a(v->pool + offs[t], len) & v->mask;
        while (v->table[h] >= 0) h = (h + 1) & v->mask;
        v->table[h] = t;
    }
    return v;
}

void orc_vocab_free(void* p) {
    orc_vocab* v = (orc_vocab*)p;
    if (!v) return;
    free(v->pool);
    free(v->offs);
    free(v->ids);
    free(v->table);
    free(v);
}

typedef struct {
    int32_t* len;
    int32_t* longest;
    int32_t* cp;   /* code points (or bytes) before position */
    uint8_t* reach;
    uint8_t* bnd;
    int32_t* unit;
    int64_t cap;
} scratch;

static void ensure(scratch* s, int64_t n) {
    if (n + 2 <= s->cap) return;
    s->cap = n + 2 + s->cap;
    s->len = (int32_t*)realloc(s->len, sizeof(int32_t) * (size_t)s->cap);
    s->longest = (int32_t*)realloc(s->longest, sizeof(int32_t) * (size_t)s->cap);
    s->cp = (int32_t*)realloc(s->cp, sizeof(int32_t) * (size_t)s->cap);
    s->unit = (int32_t*)realloc(s->unit, sizeof(int32_t) * (size_t)s->cap);
    s->reach = (uint8_t*)realloc(s->reach, (size_t)s->cap);
    s->bnd = (uint8_t*)realloc(s->bnd, (size_t)s->cap);
}

/* one word; returns number of ids written (0 if untokenizable); *word_len = len_dp[n]; *untok */
static int32_t one_word(const orc_vocab* v, scratch* S, const uint8_t* s, int64_t n, int32_t* out, int64_t cap,
                        int32_t* word_len, uint8_t* untok) {
    ensure(S, n);
    int32_t units = 0;
    S->cp[0] = 0;
    for (int64_t p = 0; p <= n; ++p) {
        const int is_b = (p == 0 || p == n) ? 1 : (v->unit_mode == 0 ? 1 : ((s[p] & 0xC0) != 0x80));
        S->bnd[p] = (uint8_t)is_b;
        S->unit[p] = is_b ? units : -1;
        if (is_b) {
            S->len[p] = units; /* phantom init, dp_tokenize.py:28 */
            ++units;
        }
        if (p < n) S->cp[p + 1] = S->cp[p] + ((v->unit_mode == 0 || (s[p] & 0xC0) != 0x80) ? 1 : 0);
        S->reach[p] = 0;
        S->longest[p] = 0;
    }
    S->reach[0] = 1;
    for (int64_t i = 1; i <= n; ++i) {
        if (!S->bnd[i]) continue;
        const int64_t lo = i - v->lmax > 0 ? i - v->lmax : 0;
        int32_t best = S->len[i];
        for (int64_t j = lo; j < i; ++j) /* ascending j like :38 */
            if (S->bnd[j] && find(v, s + j, i - j) >= 0 && S->len[j] + 1 < best) best = S->len[j] + 1;
        S->len[i] = best;
        for (int64_t j = lo; j < i; ++j) {
            if (!S->bnd[j] || S->len[j] + 1 != best || !S->reach[j]) continue;
            if (find(v, s + j, i - j) < 0) continue;
            S->reach[i] = 1;
            const int32_t cl = S->cp[i] - S->cp[j];
            int32_t m = S->longest[j] > cl ? S->longest[j] : cl;
            if (m > S->longest[i]) S->longest[i] = m;
        }
    }
    *word_len = S->len[n];
    *untok = S->reach[n] ? 0 : 1;
    if (!S->reach[n]) return 0;
    const int32_t target = S->longest[n];
    const int32_t count = S->len[n];
    int got = 0;
    int64_t i = n;
    int32_t slot = count - 1;
    while (i > 0) {
        const int64_t lo = i - v->lmax > 0 ? i - v->lmax : 0;
        int64_t pick = -1;
        for (int64_t j = lo; j < i; ++j) { /* keep the largest qualifying j */
            if (!S->bnd[j] || S->len[j] + 1 != S->len[i] || !S->reach[j]) continue;
            if (find(v, s + j, i - j) < 0) continue;
            const int32_t cl = S->cp[i] - S->cp[j];
            if (got || cl == target || S->longest[j] == target) pick = j;
        }
        const int32_t t = find(v, s + pick, i - pick);
        if (slot >= 0 && slot < cap) out[slot] = v->ids[t];
        --slot;
        if (S->cp[i] - S->cp[pick] == target) got = 1;
        i = pick;
    }
    return count;
}

/* Encode pre-split words.  Returns the number of ids (may exceed cap; extra ids are dropped). */
int64_t orc_encode_words(void* vp, const uint8_t* text, const int64_t* word_offs, int64_t n_words, int32_t* ids,
                         int64_t cap, int32_t* lens, uint8_t* flags) {
    const orc_vocab* v = (const orc_vocab*)vp;
    scratch S;
    memset(&S, 0, sizeof S);
    int64_t total = 0;
    for (int64_t w = 0; w < n_words; ++w) {
        const int64_t a = word_offs[w], n = word_offs[w + 1] - a;
        int32_t wl = 0;
        uint8_t un = 1;
        int32_t k = 0;
        if (n > 0) k = one_word(v, &S, text + a, n, ids ? ids + total : NULL, ids ? cap - total : 0, &wl, &un);
        if (lens) lens[w] = wl;
        if (flags) flags[w] = un;
        total += k;
    }
    free(S.len);
    free(S.longest);
    free(S.cp);
    free(S.unit);
    free(S.reach);
    free(S.bnd);
    return total;
}
