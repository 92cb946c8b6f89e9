// Poseidon over BN254 Fr, circom parameters, one hash per thread.
//
// Computes exactly what `PoseidonHasher::hash` does
// (pallet/src/hash/poseidon.rs:162-208: state = [tag, inputs...], 4 full
// rounds, RP partial rounds, 4 full rounds of ARK -> x^5 -> MDS, output
// state[0]) through an algebraically identical schedule:
//
//   * round constants folded into the matrix products (they ride along as the
//     initial value of the lazy Montgomery accumulator, fr.cuh);
//   * partial rounds in sparse form: the constants of the partial section are
//     pushed forward until only a scalar on s[0] remains, and the dense MDS is
//     factored so that a partial round costs 2t-1 products instead of t^2;
//   * widths 2..6: that sparse form turned into a recurrence over the S-box
//     inputs and outputs of the last t-1 rounds (Layout::HR below) -- 2(t-1)
//     products and ONE reduction per round, no passive state to update;
//   * round 0 on the inputs as they arrive (no Montgomery conversion: its
//     matrix carries the scale) and, under domain tag 0, with the S-box of the
//     constant state[0] read from the table;
//   * in the last round only row 0 of the MDS is evaluated (the reference
//     discards the rest, poseidon.rs:205), against the non-Montgomery copy of
//     that row, so the dot product lands directly on the canonical value.
//
// The tables are derived at library init from the Grain-LFSR constants
// (host_params.cpp) and are cross-checked in the tests against an independent
// Python derivation (tests/opt_model.py), which in turn is proven equal to
// the dense reference algorithm.
#pragma once
#include "fr.cuh"

// Experiment (profiles/r02_lockstep_experiment.md): -DINF_LOCKSTEP puts a block barrier in every
// iteration of the round loops so that all warps of a block fetch the same instructions at the
// same time (the loop bodies are larger than the instruction cache).  Off in the shipped build.
#if defined(INF_LOCKSTEP) && defined(__CUDA_ARCH__)
#define INF_LOCKSTEP_SYNC() __syncthreads()
#else
#define INF_LOCKSTEP_SYNC() ((void)0)
#endif

namespace inf {

// PARTIAL_ROUNDS[t-2], pallet/src/hash/parameters.rs:17-18
INF_HD constexpr int partial_rounds(int t) {
    return t == 2 ? 56 : t == 3 ? 57 : t == 4 ? 56 : t == 5 ? 60 : t == 6 ? 60 : t == 7 ? 63
         : t == 8 ? 64 : t == 9 ? 63 : t == 10 ? 60 : t == 11 ? 66 : t == 12 ? 60 : 65;
}

// Paired partial rounds (the per-thread kernels of widths 7 and 8, and the table every
// width's warp-cooperative schedule reads): two partial rounds share the update of s[1..]:
//     round A:  z_a = u^5 ;  n = z_a + v'_A . s[1..] + k'_A
//     round B:  z_b = n^5 ;  u = z_b + v'_B . s[1..] + c_B * z_a + k'_B ,   c_B = v'_B . w'_A
//     then      s_i += w'_A[i] * z_a + w'_B[i] * z_b        (one 2-term lazy dot, ONE reduction)
// (u, v', w', k': the unit-leading-coefficient form described at Layout)
// which is the same arithmetic (s_i after round A is s_i + w_A[i] x_a, substituted
// into round B's row) with T-1 fewer reductions and one more product per pair:
// -4.6 % (t=3) .. -10 % (t=6) multiply-pipe instructions per round.  An odd number
// of partial rounds (t = 3, 7) leaves one ordinary round at the end.  (The pair
// doubles the loop body; at 5 resident blocks per SM that cost t=3 more in
// instruction fetch than it saved, at the shipped 3 blocks it costs nothing —
// profiles/r01_occupancy_sweep.md.)  INF_NO_PAIRING builds the unpaired form.
#ifdef INF_NO_PAIRING
INF_HD constexpr bool paired_rounds(int) { return false; }
#else
INF_HD constexpr bool paired_rounds(int t) { return t >= 2; }
#endif

// Table layout, in units of one field element (8 x u32).  All entries are
// Montgomery form (x*R mod p) except the "V" entries, which are x*R^2 mod p
// (they enter an accumulator that is then divided by R), and OUT_ROW, which is
// canonical.
//
// Partial rounds carry s0 as u = s0 / lambda_j (lambda_0 = 1, lambda_{j+1} =
// m00_j lambda_j^5), which makes the coefficient of the fresh S-box output one:
//     z = u^5 ;  u' = z + v'_j . s[1..] + k'_j ;  s_i' = s_i + w'_j[i] z
// (v' = v / lambda_{j+1}, k' = k / lambda_{j+1}, w' = w lambda_j^5; exact field
// identities, tests/opt_model.py).  One product less per partial round than the
// plain sparse form, and nothing but an addition between one S-box and the next.
// The scale comes off in the first round of the second half, whose matrix has
// column 0 multiplied by lambda_RP^5 (TAIL0_M).
//
// Two forms of the partial section live in the table (both exact rewritings of the same map):
//   * the history recurrence (HR, widths 2..6; what the per-thread kernels run),
//   * the paired sparse rounds (widths 7, 8, and every width's warp-cooperative schedule, coop.cuh).

// History recurrence (derive_hr in tests/opt_model.py).  The passive state s[1..] (n = T-1 elements) is
// only ever read through row functionals, and the n equations
//     u_{j+1-i} = z_{j-i} + v'_{j-i} . s[1..]^(j-i) + k'_{j-i}        (i = 1..n)
// determine it from S-box inputs u and outputs z, values that exist anyway.  So a partial round is
//     z_j = u_j^5 ;   u_{j+1} = z_j + sum_{i=1..n} ( al_i u_{j+1-i} + be_i z_{j-i} ) + const :
// 2n products and ONE reduction beside the S-box and nothing else to update, against 2n + 1/2 products
// and (n + 2)/2 reductions per round of the paired sparse form (-12 % multiply-pipe instructions per
// hash5, -11 % per hash2).  The first n rounds read their rows F_j = v'_j . s[1..]^(0) + k'_j straight
// out of the merged round-3 matrix plus the few z that exist by then; after the last round n rows over
// the same history return the plain s[1..] (constants D folded in).  The widest row sums 2n terms:
// bounded by (0.189 (2n + 1.7n) + 1) p = 4.5 p < 2^256 = 5.29 p for n = 5; 5.2 p for n = 6 (no margin), 5.9 p for n = 7.
#ifdef INF_NO_HR
INF_HD constexpr bool hr_rounds(int) { return false; }
#else
INF_HD constexpr bool hr_rounds(int t) { return t >= 2 && t <= 6; }
#endif
#ifndef INF_HR_UNROLL
#define INF_HR_UNROLL 1      // rounds per loop body: 1 shifts the history with register moves, n renames them away
#endif

template <int T>
struct Layout {
    static constexpr int RP = partial_rounds(T);
    static constexpr int N = T - 1;
    // ---- every schedule -------------------------------------------------------------------------
    static constexpr int R2 = 0;                         // R^2 mod p
    static constexpr int IN_V = R2 + 1;                  // [T]   C_0[i] * R^2
    static constexpr int S0 = IN_V + T;                  // C_0[0] * R   (state[0] when tag == 0)
    static constexpr int FULL_M = S0 + 1;                // [T][T] MDS
    static constexpr int TAIL0_M = FULL_M + T * T;       // [T][T] MDS, column 0 times lambda_RP^5
    static constexpr int FULL_V = TAIL0_M + T * T;       // [3][T] C_{r+1}, r = 0..2
    static constexpr int TAIL_V = FULL_V + 3 * T;        // [3][T] C_{4+RP+r+1}, r = 0..2
    static constexpr int OUT_ROW = TAIL_V + 3 * T;       // [T]   MDS row 0, canonical
    static constexpr int OUT_ROW_MONT = OUT_ROW + T;     // [T]   MDS row 0, Montgomery (chaining)
    // (C_0[0])^5: what the first S-box makes of state[0] when the domain tag is zero (every circom
    // hasher, every tree node) -- a constant, so round 0 takes it from here instead of computing it
    static constexpr int X0 = OUT_ROW_MONT + T;
    // Round 0 of the per-thread kernels runs on the inputs as they arrive, x + C_0 as a plain integer
    // (absorb_raw) instead of (x + C_0) R: the S-box then yields s^5 / R^4, and the round's matrix
    // carries the missing R^5 (R0_M = M R^6 against FULL_M = M R), so no input pays a conversion
    // product.  X0 is stored in the same scale, (C_0[0])^5 / R^4.  IN_C = C_0 as canonical integers.
    static constexpr int R0_M = X0 + 1;                  // [T][T]
    static constexpr int IN_C = R0_M + T * T;            // [T]
    static constexpr int COMMON_END = IN_C + T;
    // ---- history recurrence ------------------------------------------------------------------------
    static constexpr bool HR = hr_rounds(T);
    static constexpr int HR_PRE_M = COMMON_END;          // [T][T] row 1 + j of PRE_M mapped to F_j
    static constexpr int HR_PRE_V = HR_PRE_M + T * T;    // [T]    (k_0, k'_0, .., k'_{n-1})
    // rounds j = 1..n-1: c_{j,i} = v'_j . w'_i for i = j-1 .. 0 (newest z first), at HR_BOOT + j(j-1)/2
    static constexpr int HR_BOOT = HR_PRE_V + T;
    // rounds j = n..RP-1: al[n] (u_j, u_{j-1}, ..), be[n] (z_{j-1}, z_{j-2}, ..), const
    static constexpr int HR_PART = HR_BOOT + N * (N - 1) / 2;
    static constexpr int HR_STRIDE = 2 * N + 1;
    static constexpr int HR_ROUNDS = RP - N;
    static constexpr int HR_EXIT = HR_PART + HR_ROUNDS * HR_STRIDE;   // n rows at j = RP: the plain s[1..] (+ D)
    static constexpr int HR_END = HR ? HR_EXIT + N * HR_STRIDE : COMMON_END;
    // ---- paired sparse rounds ----------------------------------------------------------------------
    static constexpr int PRE_M = HR_END;                 // [T][T] MDS with the sparse prefix merged
    static constexpr int PRE_V = PRE_M + T * T;          // [T]   (k_0, 0, ..., 0)
    static constexpr bool PAIRED = paired_rounds(T);
    // single round record [2T-1] : v'[T-1], w'[T-1], k'
    // pair record        [4T-1] : v'_A[T-1], k'_A, v'_B[T-1], c_B, k'_B, (w'_A[i], w'_B[i]) for i = 1..T-1
    // PAIRED: RP/2 pair records, then one single record if RP is odd; else RP single records.
    static constexpr int PART = PRE_V + T;
    static constexpr int PAIR_STRIDE = 4 * T - 1;
    static constexpr int SINGLE_STRIDE = 2 * T - 1;
    static constexpr int N_PAIRS = PAIRED ? RP / 2 : 0;
    static constexpr int N_SINGLES = PAIRED ? RP % 2 : RP;
    static constexpr int SINGLES = PART + N_PAIRS * PAIR_STRIDE;
    static constexpr int LAST_D = SINGLES + N_SINGLES * SINGLE_STRIDE;   // [T-1]  D[1..] * R (added once)
    // c_j = v'_j . w'_{j-1} for every partial round j (c_0 = 0): lets the warp-cooperative
    // kernel (coop.cuh) form v'_j . s[1..] from the state of one round earlier
    static constexpr int COOP_C = LAST_D + (T - 1);      // [RP]
    static constexpr int COUNT = COOP_C + RP;
    static constexpr int WORDS = COUNT * 8;
    // what the per-thread kernels read (a unit that only has those, leaves.cu, keeps just this prefix
    // in its constant bank)
    static constexpr int THREAD_COUNT = HR ? HR_END : COUNT;
    static constexpr int THREAD_WORDS = THREAD_COUNT * 8;
    // offsets inside a pair record
    static constexpr int P_VA = 0, P_KA = T - 1, P_VB = T, P_CB = 2 * T - 1, P_KB = 2 * T, P_W = 2 * T + 1;
    // offsets inside a single record
    static constexpr int S_V = 0, S_W = T - 1, S_K = 2 * T - 2;
};

// out = ( sum_j a[j] * b[j] + V ) / R, then (RANGE_STEP) the cheap range step.
template <int N, int STRIDE_A, bool RANGE_STEP = true>
INF_HD void dot(uint32_t (&out)[8], const uint32_t* a, const uint32_t* b, const uint32_t* v) {
    MontAcc acc;
    if (v) acc.init(v); else acc.zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) acc.row(i, a + j * STRIDE_A, b[j * 8 + i], j == 0);
        acc.reduce(i);
    }
    acc.finish(out);
    if (RANGE_STEP) csub2p(out);
}

// x^5
INF_HD void sbox(uint32_t (&y)[8], const uint32_t (&x)[8]) {
    uint32_t x2[8], x4[8];
    mont_sqr(x2, x);
    mont_sqr(x4, x2);
    mont_mul(y, x4, x);
}

// Range steps of a history-recurrence row (Layout::HR): below (0.7 n + 1) p on entry, below 2p + 2^224 on return.
template <int N>
INF_HD void hr_range(uint32_t (&v)[8]) {
    if (N >= 5) csub4p(v);          // 0.7 n + 1 > 4: one step of 2p would leave up to 2.5 p
    csub2p(v);
}

// Rounds 0..n-1 of the history recurrence: u_{J+1} = z_J + F_J + sum_{i<J} c_{J,i} z_i, with u_{J+1}
// and z_J written where the steady rounds expect them once J = n-1 is through (h[n-1-J], h[2n-1-J]):
// the z that exist so far are then contiguous, newest first, from h[2n-J].
template <int T, int J>
INF_HD void hr_boot(uint32_t (&h)[2 * (T - 1)][8], uint32_t (&s)[T][8], const uint32_t* tbl) {
    using L = Layout<T>;
    constexpr int N = T - 1;
    if constexpr (J < N) {
        uint32_t (&z)[8] = h[2 * N - 1 - J];
        uint32_t (&u)[8] = h[N - 1 - J];
        if constexpr (J == 0) {
            sbox(z, s[0]);
            add8(u, z, s[1]);
        } else {
            uint32_t v[8];
            sbox(z, h[N - J]);                                                   // u_J
            dot<J, 8, false>(v, &h[2 * N - J][0], tbl + (L::HR_BOOT + J * (J - 1) / 2) * 8, nullptr);
            add8(v, v, s[1 + J]);                                                // < 2.3 p + 2p
            csub2p(v);
            add8(u, z, v);
        }
        csub2p(u);
        hr_boot<T, J + 1>(h, s, tbl);
    }
}

// The permutation proper.  On entry s = [tag, inputs...] + C_0 in Montgomery
// form, every element < 2p + eps.  On return `out` holds state[0] after the
// last round: canonical integer in [0, p) if !MONT_OUT, Montgomery form
// (< 2p + eps) if MONT_OUT.
// `tag0`: state[0] is the table's S0 (domain tag zero), so its first S-box output is the table's X0.
template <int T, bool MONT_OUT>
// `raw_in`: s came from absorb_raw (round 0 then uses R0_M, see Layout).
INF_HD void poseidon_rounds(uint32_t (&out)[8], uint32_t (&s)[T][8], const uint32_t* tbl, const bool tag0 = false,
                            const bool raw_in = false) {
    using L = Layout<T>;
    static_assert(T >= 2 && T <= 8, "optimised path covers widths 2..8");
    uint32_t x[T][8];
    // Range step after the rows that feed an S-box.  Needed for every width >= 3:
    // without it the t = 3 bound settles at 2.646 p, a hair above the 2^255 =
    // 2.645 p the squaring needs (tools/bounds.py); only t = 2 could do without.
    constexpr bool RS = T > 2;

    // ---- first half: rounds 0..3 (round 3 uses the merged matrix) ----------
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        constexpr int PM = L::HR ? L::HR_PRE_M : L::PRE_M;
        constexpr int PV = L::HR ? L::HR_PRE_V : L::PRE_V;
        const uint32_t* m = tbl + (r == 0 && raw_in ? L::R0_M : r < 3 ? L::FULL_M : PM) * 8;
        const uint32_t* v = tbl + (r < 3 ? L::FULL_V + r * T : PV) * 8;
        INF_LOCKSTEP_SYNC();
        if (r == 0 && tag0) {
#pragma unroll
            for (int k = 0; k < 8; k++) x[0][k] = tbl[L::X0 * 8 + k];
        } else {
            sbox(x[0], s[0]);
        }
#pragma unroll
        for (int i = 1; i < T; i++) sbox(x[i], s[i]);
#pragma unroll
        for (int i = 0; i < T; i++) dot<T, 8, RS>(s[i], &x[0][0], m + i * T * 8, v + i * 8);
    }

    if constexpr (L::HR) {
        // ---- partial rounds as a recurrence over the last n S-box inputs and outputs (Layout::HR) ----
        // round 3 left s = (u_0, F_0, .., F_{n-1}).  h = (u_j, u_{j-1}, .., u_{j-n+1}; z_{j-1}, .., z_{j-n}).
        // Ranges: u < 2p + eps and z < 1.7 p, so a row is below (0.189 * 3.7 n + 1) p -- 4.5 p for n = 5,
        // under 2^256 = 5.29 p -- and below 2p + eps after its range steps; z + row < 3.7 p before its own.
        constexpr int N = L::N;
        constexpr int UNR = INF_HR_UNROLL;
        uint32_t h[2 * N][8];
        hr_boot<T, 0>(h, s, tbl);
#pragma unroll UNR
        for (int j = 0; j < L::HR_ROUNDS; j++) {
            const uint32_t* pt = tbl + (L::HR_PART + j * L::HR_STRIDE) * 8;
            uint32_t z[8], v[8];
            INF_LOCKSTEP_SYNC();
            sbox(z, h[0]);                                                      // z_j = u_j^5
            dot<2 * N, 8, false>(v, &h[0][0], pt, pt + 2 * N * 8);
            hr_range<N>(v);
#pragma unroll
            for (int i = N - 1; i > 0; i--)
#pragma unroll
                for (int k = 0; k < 8; k++) h[i][k] = h[i - 1][k], h[N + i][k] = h[N + i - 1][k];
#pragma unroll
            for (int k = 0; k < 8; k++) h[N][k] = z[k];
            add8(h[0], z, v);                                                   // u_{j+1}
            csub2p(h[0]);
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t* pt = tbl + (L::HR_EXIT + i * L::HR_STRIDE) * 8;
            dot<2 * N, 8, false>(s[1 + i], &h[0][0], pt, pt + 2 * N * 8);
            hr_range<N>(s[1 + i]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) s[0][k] = h[0][k];
    } else {
    // ---- partial rounds -----------------------------------------------------
    // q[0..T-2] = s[1..T-1], q[T-1] = z_a, q[T] = z_b: contiguous, so that the lazy
    // dots stride over (s[1..], z_a) and (z_a, z_b).  s[0] holds u.
    // Rows over s[1..] get a range step of their own from width 5 on: the sum
    // with z (< 1.66 p) must stay below 2^256 and, after its step, below 2^255.
    constexpr bool RI = T >= 5;
    uint32_t q[T + 1][8];
#pragma unroll
    for (int i = 1; i < T; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) q[i - 1][k] = s[i][k];
    if constexpr (L::N_PAIRS > 0) {
#pragma unroll 1
        for (int j = 0; j < L::N_PAIRS; j++) {
            const uint32_t* pt = tbl + (L::PART + j * L::PAIR_STRIDE) * 8;
            uint32_t n[8];
            INF_LOCKSTEP_SYNC();
            sbox(q[T - 1], s[0]);                                               // round A: z_a = u^5
            dot<T - 1, 8, RI>(n, &q[0][0], pt + L::P_VA * 8, pt + L::P_KA * 8);
            add8(n, n, q[T - 1]);
            csub2p(n);
            sbox(q[T], n);                                                      // round B: z_b = n^5
            dot<T, 8, RI>(s[0], &q[0][0], pt + L::P_VB * 8, pt + L::P_KB * 8);  // v'_B . s[1..] + c_B z_a + k'_B
            add8(s[0], s[0], q[T]);
            csub2p(s[0]);
#pragma unroll
            for (int i = 1; i < T; i++) {                                       // s_i += w'_A z_a + w'_B z_b
                uint32_t w[8];
                // (z_a, z_b < 1.7 p, constants < p: w < 1.65 p, no range step needed before the add)
                dot<2, 8, false>(w, &q[T - 1][0], pt + (L::P_W + 2 * (i - 1)) * 8, nullptr);
                add8(q[i - 1], q[i - 1], w);
                csub2p(q[i - 1]);
            }
        }
    }
#pragma unroll 1
    for (int j = 0; j < L::N_SINGLES; j++) {
        const uint32_t* pt = tbl + (L::SINGLES + j * L::SINGLE_STRIDE) * 8;
        uint32_t n0[8];
        sbox(q[T - 1], s[0]);
        dot<T - 1, 8, RI>(n0, &q[0][0], pt + L::S_V * 8, pt + L::S_K * 8);
        add8(s[0], n0, q[T - 1]);
        csub2p(s[0]);
#pragma unroll
        for (int i = 1; i < T; i++) {
            uint32_t w[8];
            mont_mul(w, q[T - 1], pt + (L::S_W + i - 1) * 8);
            add8(q[i - 1], q[i - 1], w);
            csub2p(q[i - 1]);
        }
    }
#pragma unroll
    for (int i = 1; i < T; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) s[i][k] = q[i - 1][k];
    // remaining constants of the first tail round on elements 1..T-1
#pragma unroll
    for (int i = 1; i < T; i++) {
        add8(s[i], s[i], tbl + (L::LAST_D + i - 1) * 8);
        csub2p(s[i]);
    }
    }

    // ---- second half: 3 full rounds, then the output row --------------------
#pragma unroll 1
    for (int r = 0; r < 3; r++) {
        const uint32_t* m = tbl + (r == 0 ? L::TAIL0_M : L::FULL_M) * 8;   // s[0] still holds u: see Layout
        const uint32_t* v = tbl + (L::TAIL_V + r * T) * 8;
        INF_LOCKSTEP_SYNC();
#pragma unroll
        for (int i = 0; i < T; i++) sbox(x[i], s[i]);
#pragma unroll
        for (int i = 0; i < T; i++) dot<T, 8, RS>(s[i], &x[0][0], m + i * T * 8, v + i * 8);
    }
#pragma unroll
    for (int i = 0; i < T; i++) sbox(x[i], s[i]);
    if (MONT_OUT) {
        dot<T, 8>(out, &x[0][0], tbl + L::OUT_ROW_MONT * 8, nullptr);
    } else {
        dot<T, 8>(out, &x[0][0], tbl + L::OUT_ROW * 8, nullptr);
        csub_p_exact(out);
        csub_p_exact(out);
    }
}

// Raw 256-bit integer (little-endian limbs, any value < 2^256) -> element i of
// the initial state, first round constant included:  (x + C_0[i]) * R mod p.
template <int T>
INF_HD void absorb(uint32_t (&s)[8], const uint32_t (&raw)[8], int i, const uint32_t* tbl) {
    using L = Layout<T>;
    MontAcc acc;
    acc.init(tbl + (L::IN_V + i) * 8);
    const uint32_t* r2 = tbl + L::R2 * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        acc.row(k, raw, r2[k], true);
        acc.reduce(k);
    }
    acc.finish(s);
    // (0.189 * 5.29 + 1) p < 2p: already in range
}

// The same element left as a plain integer: x + C_0[i] mod p, below 2p + 2^224 (what the
// squaring that follows needs is < 2^255 = 2.645 p).  No product: three conditional
// subtractions and one addition.  Round 0 must then run against R0_M (Layout).
template <int T>
INF_HD void absorb_raw(uint32_t (&s)[8], const uint32_t (&raw)[8], int i, const uint32_t* tbl) {
    using L = Layout<T>;
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = raw[k];
    csub4p(x);                                              // any 256-bit integer -> < 4p + 2^224
    csub2p(x);                                              // < 2p + 2^224
    uint32_t c = add8(s, x, tbl + (L::IN_C + i) * 8);       // + C_0[i] (< p): < 3p + 2^224, no carry
    (void)c;
#if defined(INF_HOST_CHECKS) && !defined(__CUDA_ARCH__)
    if (c) host_overflow_count++;
#endif
    csub2p(s);                                              // < 2p + 2^224
}

// 32-byte wire format (pallet/src/poll/poll.rs:9 `HashBytes`, big-endian; the
// little-endian variant serves hash_bytes_le, poseidon.rs:233-250) <-> limbs.
// w[] are the eight 32-bit words as loaded from memory.
INF_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
template <bool LE>
INF_HD void words_to_limbs(uint32_t (&limb)[8], const uint32_t (&w)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) limb[k] = LE ? w[k] : bswap32(w[7 - k]);
}
template <bool LE>
INF_HD void limbs_to_words(uint32_t (&w)[8], const uint32_t (&limb)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = LE ? limb[k] : bswap32(limb[7 - k]);
}

// One complete hash: in_words = (T-1) x 8 words (wire order), tag_words = 8
// words or nullptr (domain tag 0, poseidon.rs:304-307), out_words = 8 words.
template <int T, bool LE>
INF_HD void hash_words(uint32_t (&out_words)[8], const uint32_t (&in_words)[T - 1][8],
                       const uint32_t* tag_words, const uint32_t* tbl) {
    using L = Layout<T>;
    uint32_t s[T][8];
    if (tag_words) {
        uint32_t w[8], raw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = tag_words[k];
        words_to_limbs<LE>(raw, w);
        absorb_raw<T>(s[0], raw, 0, tbl);
    } else {
        // tag 0: round 0 reads X0 instead of s[0]; keep the register contents defined
#pragma unroll
        for (int k = 0; k < 8; k++) s[0][k] = 0;
    }
#pragma unroll
    for (int i = 1; i < T; i++) {
        uint32_t raw[8];
        words_to_limbs<LE>(raw, in_words[i - 1]);
        absorb_raw<T>(s[i], raw, i, tbl);
    }
    uint32_t h[8];
    poseidon_rounds<T, false>(h, s, tbl, tag_words == nullptr, true);
    limbs_to_words<LE>(out_words, h);
}

}  // namespace inf
