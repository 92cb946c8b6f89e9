// BN254 scalar field Fr on 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces, on the device, the ark-ff 0.4.2 `Fp<MontBackend<FrConfig,4>,4>`
// operations the reference calls from pallet/src/hash/poseidon.rs:127,135,
// 142,153 (add, pow([5]), mul+add) and pallet/src/poll/state.rs:290,294-296
// (from_be_bytes_mod_order, into_bigint().to_bytes_be()).
//
// One design idea carries everything here: a *lazy* Montgomery dot product
//
//        dot(a_0..a_{n-1}; b_0..b_{n-1}; V) = ( sum_j a_j*b_j + V ) / R   (mod p)
//
// computed as one interleaved (CIOS) pass: for every 32-bit limb i of the b's
// the 8x1 partial products of all n terms are accumulated, then one
// reduction step clears column i.  n products therefore share ONE reduction,
// and an additive constant V (stored pre-multiplied by R) rides along for
// free as the accumulator's initial value.  Plain Montgomery multiplication
// is the n = 1 case.
//
// Carry handling: 64-bit partial products a_k*b_i are accumulated with
// IMAD.WIDE.U32(.X) chains (PTX mad.lo.cc/madc.hi.cc pairs, which ptxas fuses).
// Products whose low column is even go to accumulator z[0], odd to z[1]
// (absolute column indexing, so nothing is ever shifted or swapped); the two
// accumulators are summed once at the end.  A chain of four products covers 8
// columns and drops its carry into a 9th limb that, in this schedule, only
// ever holds a handful of carries ("fresh" limb), so a single addc suffices.
//
// Range discipline: p < R/4 (p ~ 0.189 R), so for a_j < alpha_j p, b_j < beta_j p
//        dot < ( 0.189 * sum_j alpha_j beta_j + 1 + eps ) p
// with NO final subtraction.  Values are kept below ~2p between operations
// (csub2p), never canonical, until the single exact reduction at output.
// The per-width bound tables are in DESIGN.md; tests/test_fr_host.py runs this
// very header on the host with every intermediate checked against 2^256.
//
// The same source compiles for the host (plain C++ emulation of each PTX
// block) so the arithmetic is unit-tested on the CPU box, where there is no
// GPU; only the asm blocks differ between the two builds.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define INF_HD __host__ __device__ __forceinline__
#else
#define INF_HD inline
#endif

namespace inf {

// p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
// (pallet/src/hash/parameters.rs:14), little-endian 32-bit limbs.
#define INF_P0 0xf0000001u
#define INF_P1 0x43e1f593u
#define INF_P2 0x79b97091u
#define INF_P3 0x2833e848u
#define INF_P4 0x8181585du
#define INF_P5 0xb85045b6u
#define INF_P6 0xe131a029u
#define INF_P7 0x30644e72u
// -p^-1 mod 2^32
#define INF_PINV 0xefffffffu
// 2p, little-endian limbs
#define INF_2P0 0xe0000002u
#define INF_2P1 0x87c3eb27u
#define INF_2P2 0xf372e122u
#define INF_2P3 0x5067d090u
#define INF_2P4 0x0302b0bau
#define INF_2P5 0x70a08b6du
#define INF_2P6 0xc2634053u
#define INF_2P7 0x60c89ce5u

struct Fr {
    uint32_t v[8];
};

#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
// Host unit-test builds count every accumulator that spilled past 2^256.
inline unsigned long long host_overflow_count = 0;
#endif

// ---------------------------------------------------------------------------
// Primitive carry chains.  Each has a PTX body (device) and a C body (host).
// ---------------------------------------------------------------------------

// (r1:r0) += a0*b ; (r3:r2) += a1*b ; (r5:r4) += a2*b ; (r7:r6) += a3*b with the
// carry rippling upward; top += final carry.
INF_HD void chain4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4,
                   uint32_t& r5, uint32_t& r6, uint32_t& r7, uint32_t& top, uint32_t a0,
                   uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a2 * b + (((uint64_t)r5 << 32) | r4) + (uint64_t)(t >> 64);
    r4 = (uint32_t)t; r5 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a3 * b + (((uint64_t)r7 << 32) | r6) + (uint64_t)(t >> 64);
    r6 = (uint32_t)t; r7 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}

// Shorter chains (the squaring's triangular rows need 1..3 products).
INF_HD void chain3(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4,
                   uint32_t& r5, uint32_t& top, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32  %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32  %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32  %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32  %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32  %5, %9, %10, %5;\n\t"
        "addc.u32        %6, %6, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a2 * b + (((uint64_t)r5 << 32) | r4) + (uint64_t)(t >> 64);
    r4 = (uint32_t)t; r5 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}
INF_HD void chain2(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& top,
                   uint32_t a0, uint32_t a1, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32  %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32  %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32  %3, %6, %7, %3;\n\t"
        "addc.u32        %4, %4, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(top)
        : "r"(a0), "r"(a1), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}
INF_HD void chain1(uint32_t& r0, uint32_t& r1, uint32_t& top, uint32_t a0, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32  %1, %3, %4, %1;\n\t"
        "addc.u32        %2, %2, 0;"
        : "+r"(r0), "+r"(r1), "+r"(top)
        : "r"(a0), "r"(b));
#else
    unsigned __int128 t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}
// The same chains without the final carry (callers prove it is zero).
INF_HD void chain4_nt(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4,
                   uint32_t& r5, uint32_t& r6, uint32_t& r7, uint32_t a0,
                   uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32  %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32  %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32  %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32  %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32  %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32  %6, %11, %12, %6;\n\t"
        "madc.hi.cc.u32  %7, %11, %12, %7;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a2 * b + (((uint64_t)r5 << 32) | r4) + (uint64_t)(t >> 64);
    r4 = (uint32_t)t; r5 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a3 * b + (((uint64_t)r7 << 32) | r6) + (uint64_t)(t >> 64);
    r6 = (uint32_t)t; r7 = (uint32_t)(t >> 32);
#if defined(INF_HOST_CHECKS)
    if ((uint32_t)(t >> 64)) host_overflow_count++;     // proven zero: see MontAcc::row
#endif
#endif
}

INF_HD void chain3_nt(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4,
                   uint32_t& r5, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %6, %9, %0;\n\t"
        "madc.hi.cc.u32  %1, %6, %9, %1;\n\t"
        "madc.lo.cc.u32  %2, %7, %9, %2;\n\t"
        "madc.hi.cc.u32  %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32  %4, %8, %9, %4;\n\t"
        "madc.hi.cc.u32  %5, %8, %9, %5;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5)
        : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a2 * b + (((uint64_t)r5 << 32) | r4) + (uint64_t)(t >> 64);
    r4 = (uint32_t)t; r5 = (uint32_t)(t >> 32);
#if defined(INF_HOST_CHECKS)
    if ((uint32_t)(t >> 64)) host_overflow_count++;     // proven zero: see MontAcc::row
#endif
#endif
}

INF_HD void chain2_nt(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                   uint32_t a0, uint32_t a1, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32  %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32  %2, %5, %6, %2;\n\t"
        "madc.hi.cc.u32  %3, %5, %6, %3;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)
        : "r"(a0), "r"(a1), "r"(b));
#else
    unsigned __int128 t;
    t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)a1 * b + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
#if defined(INF_HOST_CHECKS)
    if ((uint32_t)(t >> 64)) host_overflow_count++;     // proven zero: see MontAcc::row
#endif
#endif
}

INF_HD void chain1_nt(uint32_t& r0, uint32_t& r1, uint32_t a0, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %2, %3, %0;\n\t"
        "madc.hi.cc.u32  %1, %2, %3, %1;"
        : "+r"(r0), "+r"(r1)
        : "r"(a0), "r"(b));
#else
    unsigned __int128 t = (unsigned __int128)a0 * b + (((uint64_t)r1 << 32) | r0);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
#if defined(INF_HOST_CHECKS)
    if ((uint32_t)(t >> 64)) host_overflow_count++;     // proven zero: see MontAcc::row
#endif
#endif
}

// n products (n = 0..4) into the pairs starting at zc[0], carry into zc[2n].
INF_HD void chain_n(const int n, uint32_t* zc, uint32_t o0, uint32_t o1, uint32_t o2, uint32_t o3,
                    uint32_t b, const bool top = true) {
    if (top) {
        if (n == 4) chain4(zc[0], zc[1], zc[2], zc[3], zc[4], zc[5], zc[6], zc[7], zc[8], o0, o1, o2, o3, b);
        else if (n == 3) chain3(zc[0], zc[1], zc[2], zc[3], zc[4], zc[5], zc[6], o0, o1, o2, b);
        else if (n == 2) chain2(zc[0], zc[1], zc[2], zc[3], zc[4], o0, o1, b);
        else if (n == 1) chain1(zc[0], zc[1], zc[2], o0, b);
    } else {
        if (n == 4) chain4_nt(zc[0], zc[1], zc[2], zc[3], zc[4], zc[5], zc[6], zc[7], o0, o1, o2, o3, b);
        else if (n == 3) chain3_nt(zc[0], zc[1], zc[2], zc[3], zc[4], zc[5], o0, o1, o2, b);
        else if (n == 2) chain2_nt(zc[0], zc[1], zc[2], zc[3], o0, o1, b);
        else if (n == 1) chain1_nt(zc[0], zc[1], o0, b);
    }
}

// Reduction step with the fold of the shared column:
//   m_col += s_col                      (column i lives in both accumulators)
//   m      = m_col * (-p^-1) mod 2^32
//   (r1:r0) += p1*m + carry ; (r3:r2) += p3*m ; (r5:r4) += p5*m ; (r7:r6) += p7*m
//   top    += final carry
// i.e. the odd-limb half of "+= m*p".  Returns m.  mul.lo leaves CC alone.
// (m and the p0 row CAN be formed with shifts and adds, since p0 = 2^32-2^28+1
// and -p^-1 = -(2^28+1); measured on B200 that is 11 % SLOWER — the extra ALU
// instructions cost more issue time than the IMAD + IMAD.HI they replace,
// profiles/r01_alu_reduction_experiment.md — so the multiplies stay.)
INF_HD uint32_t chain4_fold_reduce(uint32_t& m_col, uint32_t s_col, uint32_t& r0, uint32_t& r1,
                                   uint32_t& r2, uint32_t& r3, uint32_t& r4, uint32_t& r5,
                                   uint32_t& r6, uint32_t& r7, uint32_t& top) {
    uint32_t m;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32      %0, %0, %11;\n\t"
        "mul.lo.u32      %10, %0, 0xefffffff;\n\t"
        "madc.lo.cc.u32  %1, %10, 0x43e1f593, %1;\n\t"
        "madc.hi.cc.u32  %2, %10, 0x43e1f593, %2;\n\t"
        "madc.lo.cc.u32  %3, %10, 0x2833e848, %3;\n\t"
        "madc.hi.cc.u32  %4, %10, 0x2833e848, %4;\n\t"
        "madc.lo.cc.u32  %5, %10, 0xb85045b6, %5;\n\t"
        "madc.hi.cc.u32  %6, %10, 0xb85045b6, %6;\n\t"
        "madc.lo.cc.u32  %7, %10, 0x30644e72, %7;\n\t"
        "madc.hi.cc.u32  %8, %10, 0x30644e72, %8;\n\t"
        "addc.u32        %9, %9, 0;"
        : "+r"(m_col), "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6),
          "+r"(r7), "+r"(top), "=&r"(m)
        : "r"(s_col));
#else
    uint64_t c = (uint64_t)m_col + s_col;
    m_col = (uint32_t)c;
    m = m_col * INF_PINV;
    unsigned __int128 t;
    t = (unsigned __int128)INF_P1 * m + (((uint64_t)r1 << 32) | r0) + (c >> 32);
    r0 = (uint32_t)t; r1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)INF_P3 * m + (((uint64_t)r3 << 32) | r2) + (uint64_t)(t >> 64);
    r2 = (uint32_t)t; r3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)INF_P5 * m + (((uint64_t)r5 << 32) | r4) + (uint64_t)(t >> 64);
    r4 = (uint32_t)t; r5 = (uint32_t)(t >> 32);
    t = (unsigned __int128)INF_P7 * m + (((uint64_t)r7 << 32) | r6) + (uint64_t)(t >> 64);
    r6 = (uint32_t)t; r7 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
    return m;
}

// r = x + y over 8 limbs; returns the carry out.
INF_HD uint32_t add8(uint32_t (&r)[8], const uint32_t* x, const uint32_t* y) {
    uint32_t c;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(c)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)x[i] + y[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// r = x - k (k given as 8 immediate-able limbs); returns borrow (1 if x < k).
INF_HD uint32_t sub8(uint32_t (&r)[8], const uint32_t (&x)[8], uint32_t k0, uint32_t k1,
                     uint32_t k2, uint32_t k3, uint32_t k4, uint32_t k5, uint32_t k6,
                     uint32_t k7) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(bw)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(k0), "r"(k1), "r"(k2), "r"(k3), "r"(k4), "r"(k5), "r"(k6), "r"(k7));
    bw &= 1u;   // subc of 0-0-borrow gives 0 or 0xffffffff
#else
    const uint32_t k[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
    int64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (int64_t)x[i] - (int64_t)k[i];
        r[i] = (uint32_t)t;
        t >>= 32;   // arithmetic: 0 or -1
    }
    bw = (uint32_t)(t & 1);
#endif
    return bw;
}

// Exact: x -= p if x >= p.
INF_HD void csub_p_exact(uint32_t (&x)[8]) {
    uint32_t d[8];
    uint32_t bw = sub8(d, x, INF_P0, INF_P1, INF_P2, INF_P3, INF_P4, INF_P5, INF_P6, INF_P7);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = bw ? x[i] : d[i];
}

// Cheap range step: if the top limb says x is certainly >= 2p, subtract 2p.
// Afterwards x < max(2p + 2^224, bound_before - 2p).
INF_HD void csub2p(uint32_t (&x)[8]) {
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .pred q;\n\t"
        "setp.gt.u32      q, %7, 0x60c89ce5;\n\t"
        "@q sub.cc.u32    %0, %0, 0xe0000002;\n\t"
        "@q subc.cc.u32   %1, %1, 0x87c3eb27;\n\t"
        "@q subc.cc.u32   %2, %2, 0xf372e122;\n\t"
        "@q subc.cc.u32   %3, %3, 0x5067d090;\n\t"
        "@q subc.cc.u32   %4, %4, 0x0302b0ba;\n\t"
        "@q subc.cc.u32   %5, %5, 0x70a08b6d;\n\t"
        "@q subc.cc.u32   %6, %6, 0xc2634053;\n\t"
        "@q subc.u32      %7, %7, 0x60c89ce5;\n\t"
        "}"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]),
          "+r"(x[7]));
#else
    if (x[7] > INF_2P7) {
        uint32_t d[8];
        sub8(d, x, INF_2P0, INF_2P1, INF_2P2, INF_2P3, INF_2P4, INF_2P5, INF_2P6, INF_2P7);
        for (int i = 0; i < 8; i++) x[i] = d[i];
    }
#endif
}

// The same step for 4p: afterwards x < max(4p + 2^224, bound_before - 4p).  Any 256-bit
// integer (< 5.29 p) is below 4p + 2^224 after it, and below 2p + 2^224 after csub2p.
INF_HD void csub4p(uint32_t (&x)[8]) {
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .pred q;\n\t"
        "setp.gt.u32      q, %7, 0xc19139cb;\n\t"
        "@q sub.cc.u32    %0, %0, 0xc0000004;\n\t"
        "@q subc.cc.u32   %1, %1, 0x0f87d64f;\n\t"
        "@q subc.cc.u32   %2, %2, 0xe6e5c245;\n\t"
        "@q subc.cc.u32   %3, %3, 0xa0cfa121;\n\t"
        "@q subc.cc.u32   %4, %4, 0x06056174;\n\t"
        "@q subc.cc.u32   %5, %5, 0xe14116da;\n\t"
        "@q subc.cc.u32   %6, %6, 0x84c680a6;\n\t"
        "@q subc.u32      %7, %7, 0xc19139cb;\n\t"
        "}"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]),
          "+r"(x[7]));
#else
    if (x[7] > 0xc19139cbu) {
        uint32_t d[8];
        sub8(d, x, 0xc0000004u, 0x0f87d64fu, 0xe6e5c245u, 0xa0cfa121u, 0x06056174u, 0xe14116dau, 0x84c680a6u,
             0xc19139cbu);
        for (int i = 0; i < 8; i++) x[i] = d[i];
    }
#endif
}

// ---------------------------------------------------------------------------
// The lazy Montgomery accumulator.
// ---------------------------------------------------------------------------
// z[par][c]: limb at absolute column c of the accumulator that holds products
// whose low column has parity par.  Columns 0..17.
struct MontAcc {
    uint32_t z[2][18];

    INF_HD void zero() {
#pragma unroll
        for (int c = 0; c < 18; c++) z[0][c] = z[1][c] = 0;
    }
    // Start from V (8 limbs, columns 0..7): the result will contain V / R.
    INF_HD void init(const uint32_t* v) {
        zero();
#pragma unroll
        for (int c = 0; c < 8; c++) z[0][c] = v[c];
    }

    // Row i of one term: accumulate a * bi * 2^(32 i).  `first`: this is the first
    // term of row i.  Then the odd-limb chain is the first thing to touch the pair
    // (i+7, i+8) of its accumulator in this pass — column i+8 is still zero and
    // column i+7 holds only a few carries — so its carry out is provably zero
    // ((2^32-1)^2 + small < 2^64) and the final addc is dropped.
    INF_HD void row(const int i, const uint32_t* a, const uint32_t bi, const bool first = false) {
        const int A = i & 1, S = A ^ 1;   // aligned / shifted accumulator for this row
        if (first)
            chain4_nt(z[S][i + 1], z[S][i + 2], z[S][i + 3], z[S][i + 4], z[S][i + 5], z[S][i + 6],
                      z[S][i + 7], z[S][i + 8], a[1], a[3], a[5], a[7], bi);
        else
        chain4(z[S][i + 1], z[S][i + 2], z[S][i + 3], z[S][i + 4], z[S][i + 5], z[S][i + 6],
               z[S][i + 7], z[S][i + 8], z[S][i + 9], a[1], a[3], a[5], a[7], bi);
        chain4(z[A][i], z[A][i + 1], z[A][i + 2], z[A][i + 3], z[A][i + 4], z[A][i + 5],
               z[A][i + 6], z[A][i + 7], z[A][i + 8], a[0], a[2], a[4], a[6], bi);
    }

    // Row i of a squaring: only the products a_k * a_i with k >= i, the
    // off-diagonal ones against the doubled operand (36 products instead of
    // 64).  d = limbs of 2a; e[k] = a[k] << 1 is limb k of 2*(a with limbs
    // 0..k-1 cleared), which is what multiplies a_i at k = i + 1 (the bit that
    // d[i+1] inherits from a_i belongs to the diagonal term and must not be
    // counted twice).  Column i is complete after rows 0..i, as for a product:
    // every pair (j, k) with j <= k and j + k = i has j <= i.
    INF_HD void sqr_row(const int i, const uint32_t* a, const uint32_t* d, const uint32_t* e) {
        const int A = i & 1, S = A ^ 1;
        const int ke = (i & 1) ? i + 1 : i;      // first even limb index >= i
        const int ko = (i & 1) ? i : i + 1;      // first odd  limb index >= i
        const int ne = ke <= 6 ? (8 - ke) / 2 : 0, no = ko <= 7 ? (9 - ko) / 2 : 0;
#define INF_SQ_OP(k) ((k) > 7 ? 0u : (k) == i ? a[(k)] : ((k) == i + 1 ? e[(k)] : d[(k)]))
        if (no > 0)   // first (only) chain on the pair (i+7, i+8): no carry out, as in row()
            chain_n(no, &z[S][i + ko], INF_SQ_OP(ko), INF_SQ_OP(ko + 2), INF_SQ_OP(ko + 4), INF_SQ_OP(ko + 6), a[i],
                    false);
        if (ne > 0)
            chain_n(ne, &z[A][i + ke], INF_SQ_OP(ke), INF_SQ_OP(ke + 2), INF_SQ_OP(ke + 4), INF_SQ_OP(ke + 6), a[i]);
#undef INF_SQ_OP
    }

    // Reduction step i (after all terms' row i): fold the column the two
    // accumulators share, then make column i zero by adding m*p*2^(32 i).
    INF_HD void reduce(const int i) {
        const int A = i & 1, S = A ^ 1;
        uint32_t m;
        if (i > 0) {
            m = chain4_fold_reduce(z[A][i], z[S][i], z[S][i + 1], z[S][i + 2], z[S][i + 3],
                                   z[S][i + 4], z[S][i + 5], z[S][i + 6], z[S][i + 7], z[S][i + 8],
                                   z[S][i + 9]);
        } else {
            m = z[A][i] * INF_PINV;
            chain4(z[S][i + 1], z[S][i + 2], z[S][i + 3], z[S][i + 4], z[S][i + 5], z[S][i + 6],
                   z[S][i + 7], z[S][i + 8], z[S][i + 9], INF_P1, INF_P3, INF_P5, INF_P7, m);
        }
        chain4(z[A][i], z[A][i + 1], z[A][i + 2], z[A][i + 3], z[A][i + 4], z[A][i + 5],
               z[A][i + 6], z[A][i + 7], z[A][i + 8], INF_P0, INF_P2, INF_P4, INF_P6, m);
    }

    // Sum the two accumulators over columns 8..15.  (Column 16 is zero by the
    // range discipline; the host build checks it.)
    INF_HD void finish(uint32_t (&r)[8]) {
        uint32_t c = add8(r, &z[0][8], &z[1][8]);
        (void)c;
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
        if ((c | z[0][16] | z[1][16] | z[0][17] | z[1][17]) != 0) host_overflow_count++;
#endif
    }
};

// r = a*b/R
INF_HD void mont_mul(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b) {
    MontAcc t;
    t.zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t.row(i, a, b[i], true);
        t.reduce(i);
    }
    t.finish(r);
}

// r = a*a/R with 36 + 64 wide multiplies instead of 64 + 64.  Needs a < 2^255
// (true for every value the range discipline lets through) so that 2a fits 8 limbs.
INF_HD void mont_sqr(uint32_t (&r)[8], const uint32_t* a) {
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
    if (a[7] >> 31) host_overflow_count++;          // precondition a < 2^255
#endif
    uint32_t d[8], e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        e[k] = a[k] << 1;
        d[k] = k == 0 ? e[k] : (e[k] | (a[k - 1] >> 31));
    }
    MontAcc t;
    t.zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t.sqr_row(i, a, d, e);
        t.reduce(i);
    }
    t.finish(r);
}

// r = (a*b + V)/R
INF_HD void mont_mul_add(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b,
                         const uint32_t* v) {
    MontAcc t;
    t.init(v);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t.row(i, a, b[i], true);
        t.reduce(i);
    }
    t.finish(r);
}

// r = x/R  (leave Montgomery form); result <= p, exact reduction by caller.
INF_HD void mont_redc(uint32_t (&r)[8], const uint32_t* x) {
    MontAcc t;
    t.init(x);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t.reduce(i);
    }
    t.finish(r);
}

}  // namespace inf
