// EXPERIMENT (not part of the library): one level of Karatsuba under the lazy
// Montgomery dot product.  a = aL + aH 2^128, b = bL + bH 2^128:
//      sum_j a_j b_j = P0 + (P1 - P0 - P2) 2^128 + P2 2^256
//      P0 = sum aL bL,  P2 = sum aH bH,  P1 = sum (aL + aH)(bL + bH)
// 3 x 16 = 48 wide multiplies per term instead of 64; the 64 of the shared
// reduction are unchanged.  What it costs is ALU work (the half sums with their
// carry bits, P1 - P0 - P2, the recombination) and registers (three product
// accumulators live at once).  tools/experiments/mulbench.cu times it against
// MontAcc; every block has a C body so the arithmetic is checked on the host.
#pragma once
#include "fr.cuh"

namespace inf {

// r += x over N limbs, returns the carry out (CGBN-style chained asm statements).
template <int N>
INF_HD uint32_t add_n(uint32_t* r, const uint32_t* x) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r[0]) : "r"(x[0]));
#pragma unroll
    for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
    return c;
#else
    uint64_t t = 0;
    for (int i = 0; i < N; i++) {
        t += (uint64_t)r[i] + x[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    return (uint32_t)t;
#endif
}
// r -= x over N limbs (caller guarantees r >= x)
template <int N>
INF_HD void sub_n(uint32_t* r, const uint32_t* x) {
#ifdef __CUDA_ARCH__
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(r[0]) : "r"(x[0]));
#pragma unroll
    for (int i = 1; i < N - 1; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
    asm volatile("subc.u32 %0, %0, %1;" : "+r"(r[N - 1]) : "r"(x[N - 1]));
#else
    int64_t t = 0;
    for (int i = 0; i < N; i++) {
        t += (int64_t)r[i] - (int64_t)x[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
#endif
}
// r[0..N) += x[0..M) (M <= N), carry rippling through the rest of r
template <int N, int M>
INF_HD void add_ripple(uint32_t* r, const uint32_t* x) {
#ifdef __CUDA_ARCH__
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r[0]) : "r"(x[0]));
#pragma unroll
    for (int i = 1; i < M; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
#pragma unroll
    for (int i = M; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(r[i]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(r[N - 1]));
#else
    uint64_t t = 0;
    for (int i = 0; i < N; i++) {
        t += (uint64_t)r[i] + (i < M ? x[i] : 0u);
        r[i] = (uint32_t)t;
        t >>= 32;
    }
#endif
}

// 4x4-limb product accumulator, carry-save over two registers sets like MontAcc.
struct HalfAcc {
    uint32_t z[2][9];
    INF_HD void zero() {
#pragma unroll
        for (int c = 0; c < 9; c++) z[0][c] = z[1][c] = 0;
    }
    // += a[0..3] * bi * 2^(32 i), i = 0..3
    INF_HD void row(const int i, const uint32_t* a, const uint32_t bi) {
        const int A = i & 1, S = A ^ 1;
        chain2(z[S][i + 1], z[S][i + 2], z[S][i + 3], z[S][i + 4], z[S][i + 5], a[1], a[3], bi);
        chain2(z[A][i], z[A][i + 1], z[A][i + 2], z[A][i + 3], z[A][i + 4], a[0], a[2], bi);
    }
    INF_HD void sum(uint32_t (&r)[9]) {
#pragma unroll
        for (int c = 0; c < 9; c++) r[c] = z[0][c];
        add_n<9>(r, z[1]);
    }
};

// out = (sum_j a_j b_j) / R.  bs[j] = {bL + bH (4 limbs), carry}: for the real
// kernel these would sit in the constant table next to b.
template <int N, int STRIDE_A>
INF_HD void dot_kara(uint32_t (&out)[8], const uint32_t* a, const uint32_t* b, const uint32_t* bs) {
    HalfAcc p0, p1, p2;
    p0.zero(); p1.zero(); p2.zero();
    uint32_t fix[6] = {0, 0, 0, 0, 0, 0};          // carry-bit cross terms of P1, at limb 4
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t* aj = a + j * STRIDE_A;
        const uint32_t* bj = b + j * 8;
        const uint32_t* sj = bs + j * 5;
        uint32_t sa[4] = {aj[0], aj[1], aj[2], aj[3]};
        const uint32_t ca = add_n<4>(sa, aj + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            p0.row(i, aj, bj[i]);
            p2.row(i, aj + 4, bj[4 + i]);
            p1.row(i, sa, sj[i]);
        }
        // (sa + ca 2^128)(sb + cb 2^128) = sa sb + (ca sb + cb sa) 2^128 + ca cb 2^256
        const uint32_t ma = 0u - ca, mb = 0u - sj[4];
        uint32_t t[5] = {ma & sj[0], ma & sj[1], ma & sj[2], ma & sj[3], ca & sj[4]};
        add_ripple<6, 5>(fix, t);
        uint32_t u[4] = {mb & sa[0], mb & sa[1], mb & sa[2], mb & sa[3]};
        add_ripple<6, 4>(fix, u);
    }
    uint32_t P0[9], P1[11], P2[9];
    p0.sum(P0);
    p2.sum(P2);
    {
        uint32_t t[9];
        p1.sum(t);
#pragma unroll
        for (int c = 0; c < 9; c++) P1[c] = t[c];
        P1[9] = P1[10] = 0;
    }
    add_ripple<7, 6>(P1 + 4, fix);
    // mid = P1 - P0 - P2  (>= 0, < N 2^258)
    {
        uint32_t t[11];
#pragma unroll
        for (int c = 0; c < 11; c++) t[c] = c < 9 ? P0[c] : 0u;
        sub_n<11>(P1, t);
#pragma unroll
        for (int c = 0; c < 9; c++) t[c] = P2[c];
        sub_n<11>(P1, t);
    }
    // T = P0 + mid 2^128 + P2 2^256 in z[0]; reduce as usual
    MontAcc acc;
    acc.zero();
#pragma unroll
    for (int c = 0; c < 8; c++) acc.z[0][c] = P0[c];
    {
        // P0's 9th limb (only non-zero for N > 1) lands on column 8
        uint32_t hi[10];
#pragma unroll
        for (int c = 0; c < 9; c++) hi[c] = P2[c];
        hi[9] = 0;
        uint32_t p08[1] = {P0[8]};
        add_ripple<10, 1>(hi, p08);
#pragma unroll
        for (int c = 0; c < 10; c++) acc.z[0][8 + c] = hi[c];
    }
    add_ripple<14, 11>(&acc.z[0][4], P1);
#pragma unroll
    for (int i = 0; i < 8; i++) acc.reduce(i);
    acc.finish(out);
}

INF_HD void half_sums(uint32_t* bs, const uint32_t* b, int n) {
    for (int j = 0; j < n; j++) {
        uint32_t s[4] = {b[8 * j], b[8 * j + 1], b[8 * j + 2], b[8 * j + 3]};
        bs[5 * j + 4] = add_n<4>(s, b + 8 * j + 4);
        for (int k = 0; k < 4; k++) bs[5 * j + k] = s[k];
    }
}

}  // namespace inf
