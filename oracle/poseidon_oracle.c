/* Oracle B: plain-C restatement of the reference hot path, for parity at scale
 * and as the timed CPU baseline ("port": the Rust reference cannot be built in
 * this environment — no cargo/rustc, ark-ff 0.4.2 / ark-bn254 0.4.0 not vendored).
 *
 * TEST INFRASTRUCTURE — see oracle/__init__.py.  Never linked into the product.
 *
 * Follows, step for step (paths relative to the reference checkout):
 *   pallet/src/hash/poseidon.rs:123-157   apply_ark / apply_sbox_full / apply_sbox_partial / apply_mds
 *   pallet/src/hash/poseidon.rs:162-208   PoseidonHasher::hash (dense schedule, every round)
 *   pallet/src/poll/state.rs:176-225      insert (frontier stack, cascade)
 *   pallet/src/poll/state.rs:230-281      merge (zero padding, to_depth)
 *   pallet/src/poll/state.rs:284-302      PollStateTree::hash (from_be_bytes_mod_order, to_bytes_be)
 * Arithmetic: 4 x u64 Montgomery with unsigned __int128, the same strategy as
 * ark-ff's MontBackend<FrConfig,4> that the reference delegates to.
 * Constants: oracle/_gen/poseidon_constants.h (gen_constants.py).
 *
 * Two variants of the hasher cost model:
 *   hoisted (default)  parameters converted to Montgomery form once
 *   faithful           parameters rebuilt for EVERY hash, as state.rs:286 ->
 *                      poseidon.rs:322 -> parameters.rs:35 does (t*(8+RP)+t*t
 *                      Montgomery conversions and fresh vectors per hash, and a
 *                      fresh vector per apply_mds)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "_gen/poseidon_constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr;

static const fr MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const uint64_t NINV = 0xc2e1f593efffffffULL;
static const fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const fr ONE = {{1, 0, 0, 0}};

static inline int geq(const fr* a, const fr* b) {
    for (int i = 3; i >= 0; i--) if (a->l[i] != b->l[i]) return a->l[i] > b->l[i];
    return 1;
}
static inline void sub_raw(fr* r, const fr* a, const fr* b) {
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->l[i] - b->l[i] - bw;
        r->l[i] = (uint64_t)t;
        bw = (uint64_t)(t >> 64) & 1;
    }
}
static inline void fr_add(fr* r, const fr* a, const fr* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
    if (geq(r, &MOD)) sub_raw(r, r, &MOD);
}
static inline void fr_mul(fr* r, const fr* a, const fr* b) {   /* Montgomery product */
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NINV;
        c = (u128)m * MOD.l[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * MOD.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fr x = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq(&x, &MOD)) sub_raw(&x, &x, &MOD);
    *r = x;
}
/* Fr::from_be_bytes_mod_order for exactly 32 bytes, then into Montgomery form */
static inline void fr_from_be(fr* r, const uint8_t* b) {
    fr x;
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[(3 - i) * 8 + k];
        x.l[i] = w;
    }
    while (geq(&x, &MOD)) sub_raw(&x, &x, &MOD);      /* < 2^256 <= 5.3 p */
    fr_mul(r, &x, &R2);
}
static inline void fr_to_be(uint8_t* b, const fr* a) {    /* into_bigint().to_bytes_be() */
    fr x;
    fr_mul(&x, a, &ONE);
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) b[(3 - i) * 8 + k] = (uint8_t)(x.l[i] >> (56 - 8 * k));
}

/* ---- parameters ----------------------------------------------------------- */
typedef struct { int t, rp; fr* ark; fr* mds; } params;
static params G[14];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build_params(params* p, int t) {      /* get_poseidon_parameters(t): F::from(BigInteger256) per literal */
    p->t = t; p->rp = ORACLE_RP[t];
    int n_ark = (8 + p->rp) * t;
    p->ark = (fr*)malloc(sizeof(fr) * n_ark);
    p->mds = (fr*)malloc(sizeof(fr) * t * t);
    for (int i = 0; i < n_ark; i++) {
        fr c; memcpy(c.l, ORACLE_CONST[ORACLE_ARK_OFF[t] + i], 32);
        fr_mul(&p->ark[i], &c, &R2);
    }
    for (int i = 0; i < t * t; i++) {
        fr c; memcpy(c.l, ORACLE_CONST[ORACLE_MDS_OFF[t] + i], 32);
        fr_mul(&p->mds[i], &c, &R2);
    }
}
static void init_all(void) { for (int t = 2; t <= 13; t++) build_params(&G[t], t); }

/* ---- the hash (poseidon.rs:162-208) ----------------------------------------- */
static inline void pow5(fr* a) { fr a2, a4; fr_mul(&a2, a, a); fr_mul(&a4, &a2, &a2); fr_mul(a, &a4, a); }

static void hash_with(const params* p, const fr* inputs, const fr* tag, fr* out, int faithful) {
    const int t = p->t, rounds = 8 + p->rp, half = 4;
    fr sbuf[13], nbuf[13];
    fr* state = sbuf;
    fr* fresh = nbuf;
    state[0] = *tag;
    for (int i = 1; i < t; i++) state[i] = inputs[i - 1];
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < t; i++) fr_add(&state[i], &state[i], &p->ark[r * t + i]);      /* apply_ark */
        if (r < half || r >= half + p->rp) { for (int i = 0; i < t; i++) pow5(&state[i]); } /* sbox_full */
        else pow5(&state[0]);                                                                /* sbox_partial */
        if (faithful) fresh = (fr*)malloc(sizeof(fr) * t);                                   /* new Vec per apply_mds */
        for (int i = 0; i < t; i++) {                                                        /* apply_mds */
            fr acc = {{0, 0, 0, 0}}, m;
            for (int j = 0; j < t; j++) { fr_mul(&m, &state[j], &p->mds[i * t + j]); fr_add(&acc, &acc, &m); }
            fresh[i] = acc;
        }
        if (faithful) { memcpy(sbuf, fresh, sizeof(fr) * t); free(fresh); state = sbuf; }
        else { fr* tmp = state; state = fresh; fresh = tmp; }
    }
    *out = state[0];
}

/* PollStateTree::hash on bytes (state.rs:284-302); k inputs of 32 bytes BE */
static void hash_bytes(int k, const uint8_t* in, const uint8_t* tag_be, uint8_t* out, int faithful) {
    pthread_once(&g_once, init_all);
    const int t = k + 1;
    params local;
    const params* p = &G[t];
    if (faithful) { build_params(&local, t); p = &local; }     /* new_circom per hash (state.rs:286) */
    fr ins[12], tag = {{0, 0, 0, 0}}, h;
    for (int i = 0; i < k; i++) fr_from_be(&ins[i], in + 32 * i);
    if (tag_be) fr_from_be(&tag, tag_be);
    hash_with(p, ins, &tag, &h, faithful);
    fr_to_be(out, &h);
    if (faithful) { free(local.ark); free(local.mds); }
}

int oracle_hash(int n_inputs, const uint8_t* in, const uint8_t* tag_be, uint8_t* out, int faithful) {
    if (n_inputs < 1 || n_inputs > 12) return -1;
    hash_bytes(n_inputs, in, tag_be, out, faithful);
    return 0;
}

/* ---- threads ----------------------------------------------------------------- */
typedef struct { int k, faithful; const uint8_t* in; uint8_t* out; uint64_t lo, hi; } job;
static void* batch_worker(void* a) {
    job* j = (job*)a;
    for (uint64_t i = j->lo; i < j->hi; i++)
        hash_bytes(j->k, j->in + i * (uint64_t)j->k * 32, NULL, j->out + i * 32, j->faithful);
    return NULL;
}
/* n independent hashes, split over `threads` threads */
int oracle_hash_batch(int n_inputs, const uint8_t* in, uint64_t n, uint8_t* out, int threads, int faithful) {
    if (n_inputs < 1 || n_inputs > 12) return -1;
    pthread_once(&g_once, init_all);
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    job* jobs = (job*)malloc(sizeof(job) * threads);
    for (int w = 0; w < threads; w++) {
        jobs[w] = (job){n_inputs, faithful, in, out, n * w / threads, n * (w + 1) / threads};
        pthread_create(&th[w], NULL, batch_worker, &jobs[w]);
    }
    for (int w = 0; w < threads; w++) pthread_join(th[w], NULL);
    free(th); free(jobs);
    return 0;
}

/* ---- the tree, faithfully: new + insert*N + merge (state.rs:142-281) ----------- */
typedef struct { uint8_t d; uint8_t h[32]; } entry;
typedef struct {
    int arity, full_depth, depth; uint32_t count; int has_root; uint8_t root[32];
    entry st[512]; int n;   /* frontier stack: <= (arity-1)*depth + arity entries */
} tree;

static const uint8_t (*zeroes_for(int arity))[32] { return arity == 2 ? ORACLE_ZEROES_BINARY : ORACLE_ZEROES_QUINARY; }

static int tree_insert(tree* t, const uint8_t* leaf, int faithful) {
    if (t->has_root) return 1;                                   /* TreeAlreadyFull */
    t->count++;
    t->st[t->n].d = 0; memcpy(t->st[t->n].h, leaf, 32); t->n++;
    const int a = t->arity;
    for (;;) {
        if (t->n < a) break;
        entry* sub = &t->st[t->n - a];
        int d = sub[0].d, same = 1;
        for (int i = 1; i < a; i++) if (sub[i].d != d) { same = 0; break; }
        if (!same) break;
        uint8_t buf[5 * 32], h[32];
        for (int i = 0; i < a; i++) memcpy(buf + 32 * i, sub[i].h, 32);
        hash_bytes(a, buf, NULL, h, faithful);
        t->n -= a;
        t->st[t->n].d = (uint8_t)(d + 1); memcpy(t->st[t->n].h, h, 32); t->n++;
        if (t->depth < d + 1) t->depth = d + 1;
    }
    if (t->n == 1 && t->st[0].d == t->full_depth) { t->has_root = 1; memcpy(t->root, t->st[0].h, 32); t->n = 0; }
    return 0;
}
static int tree_merge(tree* t, int to_depth, int faithful) {
    if (t->has_root) return 2;                                   /* TreeAlreadyMerged */
    const uint8_t (*Z)[32] = zeroes_for(t->arity);
    const int a = t->arity;
    while (t->n > 0) {
        int d = t->st[t->n - 1].d;
        if (t->n == 1 && (!to_depth || d == t->full_depth)) break;
        int size = 0;
        while (size < t->n && t->st[t->n - 1 - size].d == d) size++;
        uint8_t buf[5 * 32], h[32];
        for (int i = 0; i < size; i++) memcpy(buf + 32 * i, t->st[t->n - size + i].h, 32);
        for (int i = size; i < a; i++) memcpy(buf + 32 * i, Z[d], 32);
        hash_bytes(a, buf, NULL, h, faithful);
        t->n -= size;
        t->st[t->n].d = (uint8_t)(d + 1); memcpy(t->st[t->n].h, h, 32); t->n++;
    }
    if (t->n == 1) { t->has_root = 1; memcpy(t->root, t->st[0].h, 32); t->n = 0; }
    return 0;
}

/* Returns 0, or the MerkleTreeError code (1 full, 2 already merged).  out_state:
 * [0]=depth field, [1]=count, [2]=has_root. */
int oracle_tree_insert_merge(int arity, int full_depth, int blank, int to_depth, const uint8_t* leaves,
                             uint64_t n, uint8_t* root, uint32_t* out_state, int faithful) {
    pthread_once(&g_once, init_all);
    tree* t = (tree*)calloc(1, sizeof(tree));
    t->arity = arity; t->full_depth = full_depth;
    if (blank) { t->st[0].d = 0; memcpy(t->st[0].h, zeroes_for(arity)[0], 32); t->n = 1; }
    int rc = 0;
    for (uint64_t i = 0; i < n && !rc; i++) rc = tree_insert(t, leaves + 32 * i, faithful);
    if (!rc) rc = tree_merge(t, to_depth, faithful);
    if (root && t->has_root) memcpy(root, t->root, 32);
    if (out_state) { out_state[0] = (uint32_t)t->depth; out_state[1] = t->count; out_state[2] = (uint32_t)t->has_root; }
    free(t);
    return rc;
}

/* ---- dense level-by-level tree over all host threads (CPU baseline shape) ------- */
typedef struct { int a, faithful; const uint8_t* in; uint64_t n_in; uint8_t* out; uint64_t lo, hi; const uint8_t* zero; } ljob;
static void* level_worker(void* p) {
    ljob* j = (ljob*)p;
    uint8_t buf[5 * 32];
    for (uint64_t i = j->lo; i < j->hi; i++) {
        for (int c = 0; c < j->a; c++) {
            uint64_t idx = i * j->a + c;
            memcpy(buf + 32 * c, idx < j->n_in ? j->in + 32 * idx : j->zero, 32);
        }
        hash_bytes(j->a, buf, NULL, j->out + 32 * i, j->faithful);
    }
    return NULL;
}
/* Root of the dense tree of `depth` levels over `nodes` (already including any
 * blank leaf), zero padded.  Scratch is allocated inside. */
int oracle_dense_tree_root(int arity, int depth, const uint8_t* nodes, uint64_t n, uint8_t* root, int threads,
                           int faithful) {
    pthread_once(&g_once, init_all);
    const uint8_t (*Z)[32] = zeroes_for(arity);
    if (n == 0) { memcpy(root, Z[depth], 32); return 0; }
    if (threads < 1) threads = 1;
    uint8_t* cur = (uint8_t*)malloc(n * 32);
    memcpy(cur, nodes, n * 32);
    uint64_t n_cur = n;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    ljob* jobs = (ljob*)malloc(sizeof(ljob) * threads);
    for (int l = 0; l < depth; l++) {
        uint64_t n_next = (n_cur + arity - 1) / arity;
        uint8_t* nxt = (uint8_t*)malloc(n_next * 32);
        int w_used = (int)(n_next < (uint64_t)threads ? n_next : (uint64_t)threads);
        for (int w = 0; w < w_used; w++) {
            jobs[w] = (ljob){arity, faithful, cur, n_cur, nxt, n_next * w / w_used, n_next * (w + 1) / w_used, Z[l]};
            pthread_create(&th[w], NULL, level_worker, &jobs[w]);
        }
        for (int w = 0; w < w_used; w++) pthread_join(th[w], NULL);
        free(cur); cur = nxt; n_cur = n_next;
    }
    memcpy(root, cur, 32);
    free(cur); free(th); free(jobs);
    return 0;
}
