// BN254 scalar field Fr on 9 x 29-bit limbs, Montgomery form with R = 2^261.
//
// Replaces, on the device, the ark-ff 0.4.2 `Fp<MontBackend<FrConfig,4>,4>`
// operations the reference calls from pallet/src/hash/poseidon.rs:127,135,
// 142,153 (add, pow([5]), mul+add) and pallet/src/poll/state.rs:290,294-296
// (from_be_bytes_mod_order, into_bigint().to_bytes_be()).
//
// EXPERIMENT — NOT USED BY THE PRODUCT.  Kept with its benchmark
// (tools/experiments/mulbench.cu) because the measurement decided the design:
// on B200 this carry-free 29-bit-limb form is ~25 % SLOWER than the 32-bit
// carry-chain form in infimum_b200/csrc/fr.cuh (49 vs 67 G mul/s), because
// IMAD.WIDE.U32 runs at half the IMAD rate with or without the carry
// predicate, so 81+81 wide multiplies lose to 64+64 (profiles/r01_imad_microbench.md).
// The premise below ("full rate without carry") came from a microbenchmark that
// ptxas had strength-reduced into 64-bit adds; it is wrong.
//
// Why 29-bit limbs (the premise that did not hold).  Measured on B200:
// IMAD.WIDE.U32 without carry issues at the full 64 lanes/clk/SM, but the
// carry-linked form (IMAD.WIDE.U32.X, what mad.lo.cc/madc.hi.cc chains compile
// to) and IMAD.HI run at HALF that rate.  With saturated 32-bit limbs every
// partial product needs the carry form.  With 29-bit limbs a product is
// < 2^58, so a 64-bit column accumulator absorbs up to 63 products with no
// carry at all: the whole multiplication becomes carry-free, full-rate
// `acc64 += a*b` (one IMAD.WIDE.U32 each, written as plain C here — nvcc emits
// mad.wide.u32), and the few carries move to the ALU pipe (shift/add), which
// issues in the slots the multiply pipe leaves free.  9 limbs instead of 8
// costs 81 products instead of 64 per 254-bit product, at twice the rate.
//
// The one idea that carries everything is a *lazy* Montgomery dot product
//
//      dot(a_0..a_{n-1}; b_0..b_{n-1}; V; U) = ( sum_j a_j*b_j + V ) / R + U   (mod p)
//
// computed in one interleaved pass over the nine limbs of the b's: n products
// share ONE reduction, an additive constant V (stored pre-multiplied by R)
// rides along as the initial value of the low columns, and an addend U in
// plain Montgomery form as the initial value of the high columns.  Montgomery
// multiplication is the n = 1 case.
//
// Range discipline.  R = 2^261 ~ 169 p, so for a_j < alpha_j p, b_j < beta_j p
//      dot < ( 0.0059 * sum_j alpha_j beta_j + 1 ) p + U
// with no conditional subtraction anywhere: values float up to < 2^261 and
// every multiplication pulls them back to ~p.  The only exact reduction is at
// the output.  Limbs are kept normalised (< 2^29, top limb holds the rest).
// Column bound: a column receives at most 9 products per term plus 9 from the
// reduction, each < (2^29-1)^2, plus a carry < 2^36 and an initial limb:
// (9n + 9) * 2^58 < 2^64 needs n <= 6.
//
// The same source compiles for the host, so tests/test_host_emulation.py runs
// this very code on the CPU box with every bound checked.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define INF_HD __host__ __device__ __forceinline__
#else
#define INF_HD inline
#endif

namespace inf {

constexpr int NL = 9;                       // limbs
constexpr int LB = 29;                      // bits per limb
constexpr uint32_t LMASK = (1u << LB) - 1;

// p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
// (pallet/src/hash/parameters.rs:14) in 29-bit limbs, little-endian.
#define INF_PL0 0x10000001u
#define INF_PL1 0x1f0fac9fu
#define INF_PL2 0x0e5c2450u
#define INF_PL3 0x07d090f3u
#define INF_PL4 0x1585d283u
#define INF_PL5 0x02db40c0u
#define INF_PL6 0x00a6e141u
#define INF_PL7 0x0e5c2634u
#define INF_PL8 0x0030644eu
// -p^-1 mod 2^29
#define INF_NINV29 0x0fffffffu

#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
// Host unit-test builds count every violated bound (column overflow, limb or
// value out of range).
inline unsigned long long host_overflow_count = 0;
#define INF_CHECK(cond) do { if (!(cond)) host_overflow_count++; } while (0)
#else
#define INF_CHECK(cond) do { } while (0)
#endif

INF_HD uint32_t p_limb(int k) {
    switch (k) {
        case 0: return INF_PL0; case 1: return INF_PL1; case 2: return INF_PL2;
        case 3: return INF_PL3; case 4: return INF_PL4; case 5: return INF_PL5;
        case 6: return INF_PL6; case 7: return INF_PL7; default: return INF_PL8;
    }
}

// The lazy Montgomery accumulator: 18 columns of weight 2^(29 c).
struct Acc29 {
    uint64_t c[2 * NL];

    // Start from V in the low columns (result will contain V / R) and U in the
    // high columns (result will contain U).  Either may be null.
    INF_HD void init(const uint32_t* v, const uint32_t* u) {
#pragma unroll
        for (int k = 0; k < NL; k++) {
            c[k] = v ? (uint64_t)v[k] : 0ull;
            c[NL + k] = u ? (uint64_t)u[k] : 0ull;
        }
    }

    // Row i of one term: += a * bi * 2^(29 i).  Carry-free.
    INF_HD void row(const int i, const uint32_t* a, const uint32_t bi) {
#pragma unroll
        for (int k = 0; k < NL; k++) {
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
            INF_CHECK(a[k] <= (k == NL - 1 ? 0xffffffffu : LMASK) && bi <= LMASK);
            INF_CHECK(c[i + k] <= ~0ull - (uint64_t)a[k] * bi);
#endif
            c[i + k] += (uint64_t)a[k] * bi;
        }
    }

    // Reduction step i (after every term's row i): add m*p*2^(29 i) so that
    // column i becomes a multiple of 2^29, and push its carry up.
    INF_HD void reduce(const int i) {
        const uint32_t m = ((uint32_t)c[i] * INF_NINV29) & LMASK;
#pragma unroll
        for (int k = 0; k < NL; k++) {
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
            INF_CHECK(c[i + k] <= ~0ull - (uint64_t)m * p_limb(k));
#endif
            c[i + k] += (uint64_t)m * p_limb(k);
        }
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
        INF_CHECK((c[i] & LMASK) == 0);
        INF_CHECK(c[i + 1] <= ~0ull - (c[i] >> LB));
#endif
        c[i + 1] += c[i] >> LB;
    }

    // Normalise columns 9..17 into limbs.
    INF_HD void finish(uint32_t (&r)[NL]) {
#pragma unroll
        for (int k = 0; k < NL - 1; k++) {
            r[k] = (uint32_t)c[NL + k] & LMASK;
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
            INF_CHECK(c[NL + k + 1] <= ~0ull - (c[NL + k] >> LB));
#endif
            c[NL + k + 1] += c[NL + k] >> LB;
        }
        r[NL - 1] = (uint32_t)c[2 * NL - 1];
        INF_CHECK(c[2 * NL - 1] <= LMASK);          // value < 2^261
    }
};

// out = ( sum_{j<N} a[j*STRIDE_A ..] * b[j*NL ..] + V ) / R + U
template <int N, int STRIDE_A>
INF_HD void dot29(uint32_t (&out)[NL], const uint32_t* a, const uint32_t* b, const uint32_t* v,
                  const uint32_t* u) {
    static_assert(N >= 1 && N <= 6, "column accumulators hold at most 6 terms");
    Acc29 acc;
    acc.init(v, u);
#pragma unroll
    for (int i = 0; i < NL; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) acc.row(i, a + j * STRIDE_A, b[j * NL + i]);
        acc.reduce(i);
    }
    acc.finish(out);
}

// r = a*b/R
INF_HD void mul29(uint32_t (&r)[NL], const uint32_t* a, const uint32_t* b) {
    dot29<1, NL>(r, a, b, nullptr, nullptr);
}

// r = a*a/R.  Symmetric products are formed once against the doubled limbs:
// 45 multiplications instead of 81.  2*a[k] < 2^30 keeps every product below
// 2^59; a column gets at most 5 of those plus 9 reduction products.
INF_HD void sqr29(uint32_t (&r)[NL], const uint32_t* a) {
    Acc29 acc;
    acc.init(nullptr, nullptr);
    uint32_t a2[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) a2[k] = a[k] << 1;
    // Row i: a[i]*a[i] at column 2i, and 2*a[k]*a[i] for k > i at column i+k.
    // Interleaved with the reduction exactly like the general product: column
    // i is complete once rows 0..i have been added, because every product
    // landing in column i has min(index) <= i/2 <= i.
#pragma unroll
    for (int i = 0; i < NL; i++) {
        acc.c[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
        for (int k = i + 1; k < NL; k++) acc.c[i + k] += (uint64_t)a2[k] * a[i];
        acc.reduce(i);
    }
    acc.finish(r);
}

// Plain limb-wise sum with carry propagation (values, not mod p).
INF_HD void add29(uint32_t (&r)[NL], const uint32_t* x, const uint32_t* y) {
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < NL - 1; k++) {
        const uint32_t s = x[k] + y[k] + carry;     // < 2^30 + 1
        r[k] = s & LMASK;
        carry = s >> LB;
    }
    r[NL - 1] = x[NL - 1] + y[NL - 1] + carry;
    INF_CHECK(r[NL - 1] <= LMASK);
}

// 8 x 32-bit little-endian limbs (any 256-bit value) -> 9 x 29-bit limbs.
INF_HD void limbs32_to_29(uint32_t (&r)[NL], const uint32_t (&w)[8]) {
#pragma unroll
    for (int k = 0; k < NL; k++) {
        const int bit = LB * k, j = bit >> 5, sh = bit & 31;
        uint32_t lo = w[j] >> sh;
        if (sh > 32 - LB && j + 1 < 8) lo |= w[j + 1] << (32 - sh);
        r[k] = lo & LMASK;
    }
}

// 9 x 29-bit limbs (value < 2^256) -> 8 x 32-bit little-endian limbs.
INF_HD void limbs29_to_32(uint32_t (&w)[8], const uint32_t (&x)[NL]) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int bit = 32 * j, k = bit / LB, sh = bit - LB * k;     // limb k holds bit `bit` at offset sh
        uint32_t v = x[k] >> sh;
        if (k + 1 < NL) v |= x[k + 1] << (LB - sh);
        if (LB - sh + LB < 32 && k + 2 < NL) v |= x[k + 2] << (2 * LB - sh);
        w[j] = v;
    }
}

// Exact x -= p if x >= p, on 8 x 32-bit limbs (value < 2^256).
INF_HD void csub_p_exact32(uint32_t (&x)[8]) {
    const uint32_t pw[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                            0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t d[8];
    uint32_t borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t t = (uint64_t)x[k] - pw[k] - borrow;
        d[k] = (uint32_t)t;
        borrow = (uint32_t)(t >> 63);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = borrow ? x[k] : d[k];
}

}  // namespace inf
