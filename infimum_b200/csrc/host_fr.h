// Host-side BN254 Fr arithmetic for table derivation at library init.
// Not a hot path: plain 4 x u64 Montgomery with unsigned __int128.
// Values are canonical integers in [0,p) at this API level (NOT Montgomery
// form); mul() converts internally.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace inf {
namespace host {

struct F {
    uint64_t l[4];
    bool operator==(const F& o) const { return !memcmp(l, o.l, sizeof l); }
    bool operator!=(const F& o) const { return !(*this == o); }
    bool is_zero() const { return !(l[0] | l[1] | l[2] | l[3]); }
};

static const F MOD = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                       0x30644e72e131a029ull}};
static const uint64_t NINV64 = 0xc2e1f593efffffffull;  // -p^-1 mod 2^64
// R = 2^256 mod p and R^2 mod p
static const F R1 = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull,
                      0x0e0a77c19a07df2full}};
static const F R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                      0x0216d0b17f4e44a5ull}};

inline F zero() { return F{{0, 0, 0, 0}}; }
inline F one() { return F{{1, 0, 0, 0}}; }
inline F from_u64(uint64_t x) { return F{{x, 0, 0, 0}}; }

inline bool geq(const F& a, const F& b) {
    for (int i = 3; i >= 0; i--)
        if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
    return true;
}
// a - b over 256 bits (wraps)
inline F sub_raw(const F& a, const F& b) {
    F r;
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)a.l[i] - b.l[i] - bw;
        r.l[i] = (uint64_t)t;
        bw = (uint64_t)(t >> 64) & 1;
    }
    return r;
}
// a + b over 256 bits; carry returned
inline F add_raw(const F& a, const F& b, uint64_t* carry) {
    F r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    *carry = (uint64_t)c;
    return r;
}
inline F add(const F& a, const F& b) {
    uint64_t c;
    F r = add_raw(a, b, &c);      // a,b < p < 2^254: no carry
    if (geq(r, MOD)) r = sub_raw(r, MOD);
    return r;
}
inline F sub(const F& a, const F& b) {
    if (geq(a, b)) return sub_raw(a, b);
    uint64_t c;
    return sub_raw(add_raw(a, MOD, &c), b);
}
inline F neg(const F& a) { return a.is_zero() ? a : sub_raw(MOD, a); }

// Montgomery product a*b/R mod p (inputs < p, output < p)
inline F mont(const F& a, const F& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NINV64;
        c = (unsigned __int128)m * MOD.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (unsigned __int128)m * MOD.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    F r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq(r, MOD)) r = sub_raw(r, MOD);
    return r;
}
inline F to_mont(const F& a) { return mont(a, R2); }              // a*R
inline F from_mont(const F& a) { return mont(a, one()); }         // a/R
inline F mul(const F& a, const F& b) { return mont(mont(a, b), R2); }
inline F pow(const F& a, const F& e) {
    F r = one(), base = a;
    for (int i = 0; i < 256; i++) {
        if ((e.l[i / 64] >> (i % 64)) & 1) r = mul(r, base);
        base = mul(base, base);
    }
    return r;
}
inline F inv(const F& a) { return pow(a, sub_raw(MOD, from_u64(2))); }

// any 256-bit value -> canonical
inline F reduce256(F a) {
    while (geq(a, MOD)) a = sub_raw(a, MOD);
    return a;
}

inline void to_limbs32(const F& a, uint32_t* out) {
    for (int i = 0; i < 4; i++) {
        out[2 * i] = (uint32_t)a.l[i];
        out[2 * i + 1] = (uint32_t)(a.l[i] >> 32);
    }
}
inline F from_be_bytes(const uint8_t* b) {   // 32 bytes, big-endian, not reduced
    F r;
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[(3 - i) * 8 + k];
        r.l[i] = w;
    }
    return r;
}
inline void to_be_bytes(const F& a, uint8_t* b) {
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) b[(3 - i) * 8 + k] = (uint8_t)(a.l[i] >> (56 - 8 * k));
}

}  // namespace host
}  // namespace inf
