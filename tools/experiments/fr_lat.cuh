// EXPERIMENT — NOT USED BY THE PRODUCT.  Kept with its benchmark
// (tools/experiments/lonewarp.cu) because the measurement decided the design: on
// B200 this split-accumulator form is 2.1 x SLOWER for a lone warp than fr.cuh's
// interleaved pass (mont_mul 1729 vs 813 cycles, sbox 4610 vs 2079): with the
// product rows free of the reduction, ptxas keeps so many carry chains in flight
// that their carries no longer fit the seven predicate registers and are spilled
// to general registers (P2R / LOP3 / ISETP: ~250 extra instructions per product),
// with or without the LEAD throttle below.  The carry-free 29-bit form (fr29.cuh)
// is slower than fr.cuh for a lone warp too (mul 1054, sqr 880 cycles).
// profiles/r02_lone_warp_latency.md has the numbers.
//
// Latency-oriented variants of the field primitives of fr.cuh, for code that runs
// ONE warp per SM sub-partition (the warp-cooperative kernel of the tree levels
// near the root, coop.cuh).
//
// There the multiply pipe is not the limit — a lone warp issues a wide multiply
// every ~5.6 cycles in fr.cuh's interleaved CIOS pass against ~4.6 when several
// warps share the pipe — dependencies are: in MontAcc the a*b rows and the m*p
// rows accumulate into the SAME registers, so row i+1 cannot start before
// reduction step i has written them.  Here the products a*b go to one pair of
// accumulators (a MontAcc that is never reduced) and the m*p rows to another;
// the only thing that ties them is the 32-bit column sum each m is made from.
// All 64 (36 for a square) partial products are then independent of the
// reduction chain, which is a short recurrence of its own:
//
//     u_i  = P0[i] + P1[i] + R0[i] + R1[i] + c_i          (33..35 bits)
//     m_i  = lo32(u_i) * (-p^-1)
//     R   += m_i * p * 2^(32 i)      without the low word of m_i * p_0, which cancels
//                                    lo32(u_i) by construction: its carry is accounted
//     c_{i+1} = hi(u_i) + (lo32(u_i) != 0)                  for here
//
// Same value as fr.cuh's result limb for limb (the sum of the four accumulators is
// the integer (sum a_j b_j + V + sum m_i p 2^(32 i)) either way), so the two forms
// are interchangeable bit for bit; the throughput kernels keep MontAcc, which
// needs 36 registers less and ~100 ALU instructions less per product.
#pragma once
#include "fr.cuh"

namespace inf {

// (r1, r2..r7, top) += hi32(m*p0) 2^0 + m*p2 2^32 + m*p4 2^96 + m*p6 2^160 with the
// carry rippling upward; r1 is the limb one above the column being cleared.
INF_HD void chain_p_even_skiplo(uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4, uint32_t& r5,
                                uint32_t& r6, uint32_t& r7, uint32_t& top, uint32_t m) {
#ifdef __CUDA_ARCH__
    asm("mad.hi.cc.u32   %0, %8, 0xf0000001, %0;\n\t"
        "madc.lo.cc.u32  %1, %8, 0x79b97091, %1;\n\t"
        "madc.hi.cc.u32  %2, %8, 0x79b97091, %2;\n\t"
        "madc.lo.cc.u32  %3, %8, 0x8181585d, %3;\n\t"
        "madc.hi.cc.u32  %4, %8, 0x8181585d, %4;\n\t"
        "madc.lo.cc.u32  %5, %8, 0xe131a029, %5;\n\t"
        "madc.hi.cc.u32  %6, %8, 0xe131a029, %6;\n\t"
        "addc.u32        %7, %7, 0;"
        : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(top)
        : "r"(m));
#else
    uint64_t t = (((uint64_t)INF_P0 * m) >> 32) + r1;
    r1 = (uint32_t)t;
    unsigned __int128 w = (unsigned __int128)INF_P2 * m + (((uint64_t)r3 << 32) | r2) + (t >> 32);
    r2 = (uint32_t)w; r3 = (uint32_t)(w >> 32);
    w = (unsigned __int128)INF_P4 * m + (((uint64_t)r5 << 32) | r4) + (uint64_t)(w >> 64);
    r4 = (uint32_t)w; r5 = (uint32_t)(w >> 32);
    w = (unsigned __int128)INF_P6 * m + (((uint64_t)r7 << 32) | r6) + (uint64_t)(w >> 64);
    r6 = (uint32_t)w; r7 = (uint32_t)(w >> 32);
    top += (uint32_t)(w >> 64);
#endif
}

struct MontAccLat {
    MontAcc p;              // products: row() / sqr_row() only, never reduce()
    uint32_t r[2][18];      // the m*p rows, same absolute column indexing as MontAcc::z
    uint32_t c;             // carry of the cleared columns into the current one

    INF_HD void zero() {
        p.zero();
#pragma unroll
        for (int k = 0; k < 18; k++) r[0][k] = r[1][k] = 0;
        c = 0;
    }
    INF_HD void init(const uint32_t* v) {
        zero();
#pragma unroll
        for (int k = 0; k < 8; k++) p.z[0][k] = v[k];
    }
    INF_HD uint32_t reduce(const int i) {
        const int A = i & 1, S = A ^ 1;
        const uint64_t u = (uint64_t)p.z[0][i] + p.z[1][i] + r[0][i] + r[1][i] + c;
        const uint32_t lo = (uint32_t)u;
        const uint32_t m = lo * INF_PINV;
        chain4(r[S][i + 1], r[S][i + 2], r[S][i + 3], r[S][i + 4], r[S][i + 5], r[S][i + 6], r[S][i + 7],
               r[S][i + 8], r[S][i + 9], INF_P1, INF_P3, INF_P5, INF_P7, m);
        chain_p_even_skiplo(r[A][i + 1], r[A][i + 2], r[A][i + 3], r[A][i + 4], r[A][i + 5], r[A][i + 6],
                            r[A][i + 7], r[A][i + 8], m);
        c = (uint32_t)(u >> 32) + (lo != 0 ? 1u : 0u);
        return m;
    }
    // limbs 8..15 of the sum of the four accumulators, plus the carry of the cleared half
    INF_HD void finish(uint32_t (&out)[8]) {
        uint64_t t = c;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            t += (uint64_t)p.z[0][8 + k] + p.z[1][8 + k] + r[0][8 + k] + r[1][8 + k];
            out[k] = (uint32_t)t;
            t >>= 32;
        }
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
        t += (uint64_t)p.z[0][16] + p.z[1][16] + r[0][16] + r[1][16];
        if (t | p.z[0][17] | p.z[1][17] | r[0][17] | r[1][17]) host_overflow_count++;
#endif
    }
};

// A zero the compiler cannot see through: `x | (m & opaque_zero())` is x, but is
// scheduled after m.  Used to keep product row i+LEAD behind reduction step i, which
// bounds the number of carry chains in flight (ptxas has seven predicates to hold
// their carries in and spills them to registers beyond that).
INF_HD uint32_t opaque_zero() {
#ifdef __CUDA_ARCH__
    uint32_t z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z));
    return z;
#else
    return 0;
#endif
}

// r = a*b/R.  LEAD = 0: product rows are free to run arbitrarily far ahead of the
// reduction; LEAD >= 1: row i + LEAD is tied to reduction step i.
template <int LEAD = 0>
INF_HD void mont_mul_lat(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b) {
    MontAccLat t;
    t.zero();
    if (LEAD == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) t.p.row(i, a, b[i], true);
#pragma unroll
        for (int i = 0; i < 8; i++) t.reduce(i);
    } else {
        const uint32_t z = opaque_zero();
#pragma unroll
        for (int i = 0; i < LEAD && i < 8; i++) t.p.row(i, a, b[i], true);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t m = t.reduce(i);
            if (i + LEAD < 8) t.p.row(i + LEAD, a, b[i + LEAD] | (m & z), true);
        }
    }
    t.finish(r);
}

// r = a*a/R, a < 2^255 (as mont_sqr)
INF_HD void mont_sqr_lat(uint32_t (&r)[8], const uint32_t* a) {
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
    if (a[7] >> 31) host_overflow_count++;
#endif
    uint32_t d[8], e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        e[k] = a[k] << 1;
        d[k] = k == 0 ? e[k] : (e[k] | (a[k - 1] >> 31));
    }
    MontAccLat t;
    t.zero();
#pragma unroll
    for (int i = 0; i < 8; i++) t.p.sqr_row(i, a, d, e);
#pragma unroll
    for (int i = 0; i < 8; i++) t.reduce(i);
    t.finish(r);
}

// r = (a*b + V)/R
INF_HD void mont_mul_add_lat(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b, const uint32_t* v) {
    MontAccLat t;
    t.init(v);
#pragma unroll
    for (int i = 0; i < 8; i++) t.p.row(i, a, b[i], true);
#pragma unroll
    for (int i = 0; i < 8; i++) t.reduce(i);
    t.finish(r);
}

// out = ( sum_j a[j] * b[j] + V ) / R, then (RANGE_STEP) the cheap range step
template <int N, int STRIDE_A, bool RANGE_STEP = true>
INF_HD void dot_lat(uint32_t (&out)[8], const uint32_t* a, const uint32_t* b, const uint32_t* v) {
    MontAccLat acc;
    if (v) acc.init(v); else acc.zero();
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < N; j++) acc.p.row(i, a + j * STRIDE_A, b[j * 8 + i], j == 0);
#pragma unroll
    for (int i = 0; i < 8; i++) acc.reduce(i);
    acc.finish(out);
    if (RANGE_STEP) csub2p(out);
}

// x^5
INF_HD void sbox_lat(uint32_t (&y)[8], const uint32_t (&x)[8]) {
    uint32_t x2[8], x4[8];
    mont_sqr_lat(x2, x);
    mont_sqr_lat(x4, x2);
    mont_mul_lat(y, x4, x);
}

}  // namespace inf
