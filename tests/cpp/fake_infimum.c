/* TEST INFRASTRUCTURE — never shipped, never linked into the product.
 *
 * A stand-in for the device-side entry points of libinfimum_b200.so, computed
 * with the oracle (oracle/poseidon_oracle.c via liboracle.so), so that the C++
 * host mirror (include/infimum_b200.hpp) and its reference-named tests
 * (tests/cpp/parity_tests.cpp) also run on the CPU box.  Linked BEFORE the real
 * library, its definitions take precedence for the test binary; host-only
 * functions it does not define (inf_strerror, inf_empty_ballot_roots,
 * inf_debug_dense_params) still come from the real library.  Signatures and
 * return codes follow include/infimum_b200.h.  On the GPU box the same test
 * source is linked against the real library only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "golden_vectors.h"
#include "infimum_b200.h"

int oracle_hash(int n_inputs, const uint8_t* in, const uint8_t* tag_be, uint8_t* out, int faithful);
int oracle_tree_insert_merge(int arity, int full_depth, int blank, int to_depth, const uint8_t* leaves, uint64_t n,
                             uint8_t* root, uint32_t* out_state, int faithful);

struct inf_ctx { int unused; };
struct inf_tree {
    uint32_t arity, depth;
    uint64_t counts[34];
    uint8_t* levels[34];
};

static uint64_t ipow(uint64_t a, uint32_t e) {
    unsigned __int128 r = 1;
    for (uint32_t i = 0; i < e; i++) {
        r *= a;
        if (r > (unsigned __int128)UINT64_MAX) return UINT64_MAX;
    }
    return (uint64_t)r;
}
static void rev32(uint8_t* d, const uint8_t* s) {
    uint8_t t[32];
    for (int i = 0; i < 32; i++) t[i] = s[31 - i];
    memcpy(d, t, 32);
}

/* one hash of k big- or little-endian inputs */
static void hash_k(uint32_t k, uint32_t flags, const uint8_t* tag, const uint8_t* in, uint8_t* out) {
    if (!(flags & INF_FLAG_LITTLE_ENDIAN)) {
        oracle_hash((int)k, in, tag, out, 0);
        return;
    }
    uint8_t buf[12 * 32], t[32];
    for (uint32_t i = 0; i < k; i++) rev32(buf + 32 * i, in + 32 * i);
    if (tag) rev32(t, tag);
    oracle_hash((int)k, buf, tag ? t : NULL, out, 0);
    rev32(out, out);
}

static void zero_table(uint32_t arity, uint8_t z[33][32]) {
    memcpy(z[0], arity == 2 ? G_BINARY_ZEROES[0] : G_QUINARY_ZEROES[0], 32);   /* the two seeds, zeroes.rs:2,38 */
    for (int l = 0; l < 32; l++) {
        uint8_t in[5 * 32];
        for (uint32_t k = 0; k < arity; k++) memcpy(in + 32 * k, z[l], 32);
        oracle_hash((int)arity, in, NULL, z[l + 1], 0);
    }
}

int inf_init(int device, inf_ctx** out) {
    (void)device;
    *out = (inf_ctx*)calloc(1, sizeof(inf_ctx));
    return INF_OK;
}
void inf_destroy(inf_ctx* ctx) { free(ctx); }
const char* inf_last_cuda_error(const inf_ctx* ctx) { (void)ctx; return ""; }

int inf_poseidon_hash_batch(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags, const uint8_t* tag, const uint8_t* in,
                            uint64_t n, uint8_t* out) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n_inputs < 1 || n_inputs + 1 > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    for (uint64_t i = 0; i < n; i++) hash_k(n_inputs, flags, tag, in + i * n_inputs * 32, out + i * 32);
    return INF_OK;
}

int inf_poseidon_hash_bytes(inf_ctx* ctx, uint32_t flags, const uint8_t* tag, const uint8_t* const* inputs,
                            const size_t* lens, uint32_t n_inputs, uint8_t out[32]) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    if (n_inputs < 1 || n_inputs + 1 > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    uint8_t buf[12 * 32];
    for (uint32_t i = 0; i < n_inputs; i++) {
        if (lens[i] == 0) return INF_ERR_EMPTY_INPUT;
        if (lens[i] != 32) return INF_ERR_INVALID_INPUT_LENGTH;
        memcpy(buf + 32 * i, inputs[i], 32);
    }
    hash_k(n_inputs, flags, tag, buf, out);
    return INF_OK;
}

/* ---- Poseidon with caller-supplied parameters: plain 4 x u64 arithmetic mod p -------------------- */
typedef struct { uint64_t l[4]; } fe;
static const fe P = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
static int geq(const fe* a, const fe* b) {
    for (int i = 3; i >= 0; i--)
        if (a->l[i] != b->l[i]) return a->l[i] > b->l[i];
    return 1;
}
static void sub(fe* a, const fe* b) {
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)a->l[i] - b->l[i] - bw;
        a->l[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
}
static void addm(fe* a, const fe* b) {            /* a, b < p  ->  a + b mod p (p < 2^254: no carry out) */
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a->l[i] + b->l[i];
        a->l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq(a, &P)) sub(a, &P);
}
static fe mulm(const fe* a, const fe* b) {         /* double and add */
    fe r = {{0, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        fe t = r;
        addm(&r, &t);
        if ((b->l[i / 64] >> (i % 64)) & 1) addm(&r, a);
    }
    return r;
}
static fe powm(const fe* a, uint64_t e) {
    fe r = {{1, 0, 0, 0}}, base = *a;
    for (; e; e >>= 1) {
        if (e & 1) r = mulm(&r, &base);
        base = mulm(&base, &base);
    }
    return r;
}
static fe from_be(const uint8_t* b) {
    fe r;
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[(3 - i) * 8 + k];
        r.l[i] = w;
    }
    while (geq(&r, &P)) sub(&r, &P);
    return r;
}
static void to_be(const fe* a, uint8_t* b) {
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) b[(3 - i) * 8 + k] = (uint8_t)(a->l[i] >> (56 - 8 * k));
}

int inf_poseidon_hash_batch_params(inf_ctx* ctx, uint32_t width, uint32_t full_rounds, uint32_t partial_rounds,
                                   uint64_t alpha, const uint8_t* ark, const uint8_t* mds, uint32_t flags,
                                   const uint8_t* tag, const uint8_t* in, uint64_t n, uint8_t* out) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (width < 2 || width > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    const int le = flags & INF_FLAG_LITTLE_ENDIAN;
    const uint32_t rounds = full_rounds + partial_rounds, half = full_rounds / 2;
    for (uint64_t h = 0; h < n; h++) {
        fe s[13], nx[13];
        uint8_t t[32];
        memset(&s[0], 0, sizeof(fe));
        if (tag) { if (le) rev32(t, tag); else memcpy(t, tag, 32); s[0] = from_be(t); }
        for (uint32_t i = 1; i < width; i++) {
            const uint8_t* p = in + (h * (width - 1) + (i - 1)) * 32;
            if (le) rev32(t, p); else memcpy(t, p, 32);
            s[i] = from_be(t);
        }
        for (uint32_t r = 0; r < rounds; r++) {
            const int full = r < half || r >= half + partial_rounds;
            for (uint32_t i = 0; i < width; i++) {
                fe c = from_be(ark + ((size_t)r * width + i) * 32);
                addm(&s[i], &c);
                if (full || i == 0) s[i] = powm(&s[i], alpha);
            }
            for (uint32_t i = 0; i < width; i++) {
                fe acc = {{0, 0, 0, 0}};
                for (uint32_t j = 0; j < width; j++) {
                    fe m = from_be(mds + ((size_t)i * width + j) * 32);
                    fe pr = mulm(&s[j], &m);
                    addm(&acc, &pr);
                }
                nx[i] = acc;
            }
            memcpy(s, nx, sizeof(fe) * width);
        }
        to_be(&s[0], t);
        if (le) rev32(out + 32 * h, t); else memcpy(out + 32 * h, t, 32);
    }
    return INF_OK;
}

/* ---- tables, trees ---------------------------------------------------------------------------- */
int inf_merkle_zeroes(inf_ctx* ctx, uint32_t arity, uint8_t out[33 * 32]) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    zero_table(arity == 2 ? 2 : 5, (uint8_t(*)[32])out);
    return INF_OK;
}

int inf_tree_merge(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int blank, int to_depth, const uint8_t* leaves,
                   uint64_t n, uint8_t root[32], uint32_t* insert_depth, uint32_t* root_depth, int* has_root) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    if (insert_depth) *insert_depth = 0;
    if (root_depth) *root_depth = 0;
    if (has_root) *has_root = 0;
    const uint64_t total = n + (blank ? 1 : 0), cap = ipow(arity, full_depth);
    if (total > cap) return INF_ERR_TREE_ALREADY_FULL;
    if (total == 0) return INF_OK;
    uint32_t st[3];
    uint8_t r[32];
    int rc = oracle_tree_insert_merge((int)arity, (int)full_depth, blank, to_depth, leaves, n, r, st, 0);
    uint32_t d = full_depth;
    if (!to_depth && total != cap) {
        d = 0;
        while (ipow(arity, d) < total) d++;
    }
    if (root) memcpy(root, r, 32);
    if (insert_depth) *insert_depth = st[0];
    if (root_depth) *root_depth = d;
    if (has_root) *has_root = 1;
    return rc;
}

/* the insert cascade of state.rs:176-225, literally, on a stack that may start from a stored frontier */
typedef struct {
    uint8_t lv[5 * 34];
    uint8_t hs[5 * 34][32];
    int top;
    uint32_t depth;
    int done;
} cascade;

static void cascade_push(cascade* c, uint32_t arity, uint32_t full_depth, const uint8_t* leaf) {
    c->lv[c->top] = 0;
    memcpy(c->hs[c->top], leaf, 32);
    c->top++;
    while (c->top >= (int)arity) {
        int same = 1;
        for (uint32_t k = 1; k < arity; k++) same &= c->lv[c->top - 1 - k] == c->lv[c->top - 1];
        if (!same) break;
        uint8_t in[5 * 32], h[32];
        for (uint32_t k = 0; k < arity; k++) memcpy(in + 32 * k, c->hs[c->top - arity + k], 32);
        oracle_hash((int)arity, in, NULL, h, 0);
        const uint8_t d = (uint8_t)(c->lv[c->top - 1] + 1);
        c->top -= (int)arity;
        c->lv[c->top] = d;
        memcpy(c->hs[c->top], h, 32);
        c->top++;
        if (d > c->depth) c->depth = d;
    }
    if (c->top == 1 && c->lv[0] == full_depth) c->done = 1;
}

static int cascade_out(const cascade* c, uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap, uint32_t* n_entries,
                       uint32_t* depth_out, int* has_root, uint8_t root[32]) {
    if (depth_out) *depth_out = c->depth;
    if (c->done) {
        if (has_root) *has_root = 1;
        if (root) memcpy(root, c->hs[0], 32);
        return INF_OK;
    }
    if ((uint32_t)c->top > cap) return INF_ERR_BUFFER_TOO_SMALL;
    if (c->top && (!out_levels || !out_hashes)) return INF_ERR_NULL_POINTER;
    for (int i = 0; i < c->top; i++) {
        out_levels[i] = c->lv[i];
        memcpy(out_hashes + 32 * i, c->hs[i], 32);
    }
    *n_entries = (uint32_t)c->top;
    return INF_OK;
}

int inf_tree_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int blank, const uint8_t* leaves, uint64_t n,
                      uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap, uint32_t* n_entries,
                      uint32_t* insert_depth, int* has_root, uint8_t root[32]) {
    if (!ctx || !n_entries) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    *n_entries = 0;
    if (insert_depth) *insert_depth = 0;
    if (has_root) *has_root = 0;
    const uint64_t total = n + (blank ? 1 : 0);
    if (total > ipow(arity, full_depth)) return INF_ERR_TREE_ALREADY_FULL;
    uint8_t z[33][32];
    zero_table(arity, z);
    cascade c;
    memset(&c, 0, sizeof c);
    if (blank) cascade_push(&c, arity, full_depth, z[0]);
    for (uint64_t i = 0; i < n; i++) cascade_push(&c, arity, full_depth, leaves + 32 * i);
    return cascade_out(&c, out_levels, out_hashes, cap, n_entries, insert_depth, has_root, root);
}

/* insert() x n on a stored frontier (include/infimum_b200.h: inf_tree_append) */
int inf_tree_append(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels, const uint8_t* in_hashes,
                    uint32_t n_in, uint32_t depth_in, const uint8_t* leaves, uint64_t n, uint8_t* out_levels,
                    uint8_t* out_hashes, uint32_t cap, uint32_t* n_entries, uint32_t* depth_out, int* has_root,
                    uint8_t root[32]) {
    if (!ctx || !n_entries) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    *n_entries = 0;
    if (has_root) *has_root = 0;
    cascade c;
    memset(&c, 0, sizeof c);
    c.depth = depth_in;
    uint64_t logical = 0;
    for (uint32_t i = 0; i < n_in; i++) {
        if (in_levels[i] >= full_depth || (i && in_levels[i] > in_levels[i - 1])) return INF_ERR_BAD_FRONTIER;
        c.lv[c.top] = in_levels[i];
        memcpy(c.hs[c.top], in_hashes + 32 * i, 32);
        c.top++;
        logical += ipow(arity, in_levels[i]);
    }
    if (logical + n > ipow(arity, full_depth)) return INF_ERR_TREE_ALREADY_FULL;
    for (uint64_t i = 0; i < n; i++) cascade_push(&c, arity, full_depth, leaves + 32 * i);
    return cascade_out(&c, out_levels, out_hashes, cap, n_entries, depth_out, has_root, root);
}

/* merge(to_depth) on a stored frontier, state.rs:230-281 */
int inf_tree_merge_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* levels, const uint8_t* hashes,
                            uint32_t n, int to_depth, uint8_t root[32], int* has_root, uint32_t* root_depth) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (has_root) *has_root = 0;
    if (root_depth) *root_depth = 0;
    if (n == 0) return INF_OK;
    uint8_t z[33][32];
    zero_table(arity, z);
    uint8_t lv[5 * 34], hs[5 * 34][32];
    int top = 0;
    for (uint32_t i = 0; i < n; i++) {
        lv[top] = levels[i];
        memcpy(hs[top], hashes + 32 * i, 32);
        top++;
    }
    for (;;) {
        const uint8_t d = lv[top - 1];
        if (top == 1 && (!to_depth || d == full_depth)) break;
        int run = 0;
        while (run < top && lv[top - 1 - run] == d) run++;
        uint8_t in[5 * 32], h[32];
        for (int k = 0; k < run; k++) memcpy(in + 32 * k, hs[top - run + k], 32);
        for (uint32_t k = (uint32_t)run; k < arity; k++) memcpy(in + 32 * k, z[d], 32);
        oracle_hash((int)arity, in, NULL, h, 0);
        top -= run;
        lv[top] = (uint8_t)(d + 1);
        memcpy(hs[top], h, 32);
        top++;
    }
    if (root) memcpy(root, hs[0], 32);
    if (has_root) *has_root = 1;
    if (root_depth) *root_depth = lv[0];
    return INF_OK;
}

/* ---- leaves (provider.rs:218-287) ------------------------------------------------------------ */
int inf_registration_leaves(inf_ctx* ctx, const uint8_t* pk, const uint64_t* ts, uint64_t n, uint8_t* leaves) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t in[4 * 32];
        memset(in, 0, sizeof in);
        memcpy(in, pk + 64 * i, 64);
        in[95] = 1;
        for (int b = 0; b < 8; b++) in[127 - b] = (uint8_t)(ts[i] >> (8 * b));
        oracle_hash(4, in, NULL, leaves + 32 * i, 0);
    }
    return INF_OK;
}
int inf_interaction_leaves(inf_ctx* ctx, const uint8_t* pk, const uint8_t* data, uint64_t n, uint8_t* leaves) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t in[4 * 32];
        oracle_hash(5, data + 320 * i, NULL, in, 0);
        oracle_hash(5, data + 320 * i + 160, NULL, in + 32, 0);
        memcpy(in + 64, pk + 64 * i, 64);
        oracle_hash(4, in, NULL, leaves + 32 * i, 0);
    }
    return INF_OK;
}

/* ---- paths (provider.rs:396-436) and retained trees --------------------------------------------- */
int inf_merkle_roots_from_paths(inf_ctx* ctx, uint32_t arity, uint32_t depth, const uint64_t* indices,
                                const uint8_t* leaves, const uint8_t* paths, uint64_t n, uint8_t* roots) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (depth > 32) return INF_ERR_BAD_DEPTH;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t cur[32];
        memcpy(cur, leaves + 32 * i, 32);
        uint64_t idx = indices[i];
        const uint8_t* p = paths + i * (uint64_t)depth * (arity - 1) * 32;
        for (uint32_t l = 0; l < depth; l++) {
            const uint32_t pos = (uint32_t)(idx % arity);
            uint8_t in[5 * 32];
            for (uint32_t j = 0, s = 0; j < arity; j++) {
                if (j == pos) memcpy(in + 32 * j, cur, 32);
                else memcpy(in + 32 * j, p + (l * (arity - 1) + s++) * 32, 32);
            }
            oracle_hash((int)arity, in, NULL, cur, 0);
            idx /= arity;
        }
        memcpy(roots + 32 * i, cur, 32);
    }
    return INF_OK;
}

int inf_tree_build(inf_ctx* ctx, uint32_t arity, uint32_t depth, int blank, const uint8_t* leaves, uint64_t n,
                   inf_tree** out) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    *out = NULL;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (depth > 32) return INF_ERR_BAD_DEPTH;
    const uint64_t total = n + (blank ? 1 : 0);
    if (total > ipow(arity, depth)) return INF_ERR_TREE_ALREADY_FULL;
    if (total == 0) return INF_ERR_MERGE_FAILED;
    uint8_t z[33][32];
    zero_table(arity, z);
    inf_tree* t = (inf_tree*)calloc(1, sizeof(inf_tree));
    t->arity = arity;
    t->depth = depth;
    t->counts[0] = total;
    t->levels[0] = (uint8_t*)malloc(total * 32);
    if (blank) memcpy(t->levels[0], z[0], 32);
    if (n) memcpy(t->levels[0] + (blank ? 32 : 0), leaves, n * 32);
    for (uint32_t l = 0; l < depth; l++) {
        const uint64_t c = t->counts[l], nx = (c + arity - 1) / arity;
        t->counts[l + 1] = nx;
        t->levels[l + 1] = (uint8_t*)malloc(nx * 32);
        for (uint64_t i = 0; i < nx; i++) {
            uint8_t in[5 * 32];
            for (uint32_t k = 0; k < arity; k++)
                memcpy(in + 32 * k, i * arity + k < c ? t->levels[l] + (i * arity + k) * 32 : z[l], 32);
            oracle_hash((int)arity, in, NULL, t->levels[l + 1] + i * 32, 0);
        }
    }
    *out = t;
    return INF_OK;
}
int inf_tree_root(inf_tree* t, uint8_t root[32]) {
    if (!t || !root) return INF_ERR_NULL_POINTER;
    memcpy(root, t->levels[t->depth], 32);
    return INF_OK;
}
int inf_tree_paths(inf_tree* t, const uint64_t* idx, uint64_t n, uint8_t* paths) {
    if (!t) return INF_ERR_NULL_POINTER;
    uint8_t z[33][32];
    zero_table(t->arity, z);
    const uint32_t a = t->arity;
    for (uint64_t i = 0; i < n; i++) {
        if (idx[i] >= ipow(a, t->depth)) return INF_ERR_BAD_DEPTH;
        uint64_t x = idx[i];
        uint8_t* p = paths + i * (uint64_t)t->depth * (a - 1) * 32;
        for (uint32_t l = 0; l < t->depth; l++) {
            const uint64_t base = x - x % a;
            for (uint32_t j = 0, s = 0; j < a; j++) {
                if (j == x % a) continue;
                memcpy(p + (l * (a - 1) + s++) * 32, base + j < t->counts[l] ? t->levels[l] + (base + j) * 32 : z[l], 32);
            }
            x /= a;
        }
    }
    return INF_OK;
}
void inf_tree_destroy(inf_tree* t) {
    if (!t) return;
    for (int l = 0; l < 34; l++) free(t->levels[l]);
    free(t);
}
