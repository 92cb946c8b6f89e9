// Host build of the device arithmetic (fr.cuh / poseidon.cuh with the PTX
// blocks replaced by their C bodies) behind a tiny C interface, so that the
// CPU test-suite can exercise the exact kernel logic where there is no GPU.
// TEST SUPPORT: built as libinfimum_hostemu.so, never loaded by the product.
#define INF_HOST_CHECKS 1
#include <cstring>

#include "host_params.h"
#include "poseidon.cuh"

using namespace inf;

template <int T>
static void hash_t(const uint8_t* in, const uint8_t* tag, uint8_t* out, int le, const uint32_t* tbl) {
    uint32_t iw[T - 1][8], ow[8], tw[8];
    memcpy(iw, in, (T - 1) * 32);
    if (tag) memcpy(tw, tag, 32);
    if (le) hash_words<T, true>(ow, iw, tag ? tw : nullptr, tbl);
    else    hash_words<T, false>(ow, iw, tag ? tw : nullptr, tbl);
    memcpy(out, ow, 32);
}

extern "C" {

unsigned long long hostemu_overflow_count() { return host_overflow_count; }

void hostemu_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    uint32_t t[8];
    mont_mul(t, a, b);
    memcpy(r, t, 32);
}
void hostemu_mont_sqr(const uint32_t* a, uint32_t* r) {
    uint32_t t[8];
    mont_sqr(t, a);
    memcpy(r, t, 32);
}
void hostemu_mont_mul_add(const uint32_t* a, const uint32_t* b, const uint32_t* v, uint32_t* r) {
    uint32_t t[8];
    mont_mul_add(t, a, b, v);
    memcpy(r, t, 32);
}
void hostemu_redc(const uint32_t* x, uint32_t* r) {
    uint32_t t[8];
    mont_redc(t, x);
    memcpy(r, t, 32);
}
// n-term lazy dot, n in 1..8: a = n x 8 limbs, b = n x 8 limbs, v = 8 limbs or NULL
int hostemu_dot(int n, const uint32_t* a, const uint32_t* b, const uint32_t* v, uint32_t* r) {
    uint32_t t[8];
    switch (n) {
        case 1: dot<1, 8>(t, a, b, v); break;
        case 2: dot<2, 8>(t, a, b, v); break;
        case 3: dot<3, 8>(t, a, b, v); break;
        case 4: dot<4, 8>(t, a, b, v); break;
        case 5: dot<5, 8>(t, a, b, v); break;
        case 6: dot<6, 8>(t, a, b, v); break;
        case 7: dot<7, 8>(t, a, b, v); break;
        case 8: dot<8, 8>(t, a, b, v); break;
        default: return -1;
    }
    memcpy(r, t, 32);
    return 0;
}
void hostemu_csub2p(uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    csub2p(t);
    memcpy(x, t, 32);
}
void hostemu_csub_p_exact(uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    csub_p_exact(t);
    memcpy(x, t, 32);
}

// One hash through the optimised schedule.  width t in 2..8.  Returns 0, or -1.
int hostemu_hash(int t, const uint8_t* in, const uint8_t* tag, uint8_t* out, int le) {
    static std::vector<uint32_t> tables[14];
    if (t < 2 || t > 8) return -1;
    if (tables[t].empty()) tables[t] = host::build_opt_table(t);
    const uint32_t* tbl = tables[t].data();
    switch (t) {
        case 2: hash_t<2>(in, tag, out, le, tbl); break;
        case 3: hash_t<3>(in, tag, out, le, tbl); break;
        case 4: hash_t<4>(in, tag, out, le, tbl); break;
        case 5: hash_t<5>(in, tag, out, le, tbl); break;
        case 6: hash_t<6>(in, tag, out, le, tbl); break;
        case 7: hash_t<7>(in, tag, out, le, tbl); break;
        case 8: hash_t<8>(in, tag, out, le, tbl); break;
    }
    return 0;
}

// Table access for cross-checks against the Python derivation.
int hostemu_opt_table_words(int t) { return (int)host::build_opt_table(t).size(); }
int hostemu_opt_table(int t, uint32_t* out) {
    std::vector<uint32_t> v = host::build_opt_table(t);
    memcpy(out, v.data(), v.size() * 4);
    return (int)v.size();
}
int hostemu_dense_params(int t, uint32_t* out /* canonical limbs: ark then mds */) {
    const host::DenseParams& d = host::grain_params(t);
    size_t k = 0;
    for (const host::F& x : d.ark) host::to_limbs32(x, out + 8 * k++);
    for (const host::F& x : d.mds) host::to_limbs32(x, out + 8 * k++);
    return (int)k;
}
}
