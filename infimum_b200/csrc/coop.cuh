// Warp-cooperative Poseidon for the levels near the root: ONE hash spread over
// T + 1 roles so that its critical path — not its instruction count — is what
// a level costs.
//
// A level with fewer nodes than resident threads costs one hash latency whatever
// its size, and a lone warp needs ~800 cycles per *dependent* multiplication
// (profiles/r01_imad_microbench.md).  In the unit-leading-coefficient form of
// the partial rounds (Layout<T>, poseidon.cuh)
//
//     z_j = u_j^5 ;   u_{j+1} = z_j + V_j ;   V_j = k'_j + v'_j . s^(j) ,   s^(j+1) = s^(j) + w'_j z_j
//
// everything except the S-box itself can be formed off the chain, one round
// ahead, because  v'_j . s^(j) = v'_j . s^(j-1) + c_j z_{j-1}  with the scalar
// c_j = v'_j . w'_{j-1}  (Layout::COOP_C):
//
//   role 0      (chain)   z_j = u_j^5, publishes z_j, takes V_j, u_{j+1} = z_j + V_j:
//                         3 dependent multiplications and one addition per round
//   role 1..T-1 (state)   own s_i: on z_j, s_i += w'_j[i] z_j, then publishes
//                         p_i = v'_{j+2}[i] s_i for the round after next
//   role T      (sum)     on z_j: V_{j+1} = c_{j+1} z_j + k'_{j+1} + sum_i p_i, publishes it
//
// Full rounds: role i < T owns state element i — S-box, publish, row i of the MDS.
// Same tables and field values as the per-thread kernel, different association of
// the additions, so results are bit-identical after the final exact reduction.
//
// The schedule is written once against a `Bus` (where published values live and
// how roles wait for each other): shared memory and named barriers on the device
// (poseidon_tu.cuh), arrays and condition variables in the threaded host
// emulation the CPU tests run (hostemu.cpp), which also checks that no barrier
// is ever entered twice in one generation.
#pragma once
#include "poseidon.cuh"

namespace inf {

// Constants of partial round j, wherever the pair / single records keep them.
template <int T>
struct CoopRound {
    const uint32_t* v;   // v'_j[0..T-2], stride 8 words
    const uint32_t* k;   // k'_j (V form)
    const uint32_t* w;   // w'_j[0..T-2], stride `ws` words
    const uint32_t* c;   // c_j
    int ws;
};
template <int T>
INF_HD CoopRound<T> coop_round(const uint32_t* tbl, int j) {
    using L = Layout<T>;
    CoopRound<T> r;
    if (j < 2 * L::N_PAIRS) {
        const uint32_t* pt = tbl + (L::PART + (j >> 1) * L::PAIR_STRIDE) * 8;
        const int odd = j & 1;
        r.v = pt + (odd ? L::P_VB : L::P_VA) * 8;
        r.k = pt + (odd ? L::P_KB : L::P_KA) * 8;
        r.w = pt + (L::P_W + odd) * 8;
        r.ws = 16;
    } else {
        const uint32_t* pt = tbl + (L::SINGLES + (j - 2 * L::N_PAIRS) * L::SINGLE_STRIDE) * 8;
        r.v = pt + L::S_V * 8;
        r.k = pt + L::S_K * 8;
        r.w = pt + L::S_W * 8;
        r.ws = 8;
    }
    r.c = tbl + (L::COOP_C + j) * 8;
    return r;
}

// Bus concept
//   slots:   x(buf, i)  S-box outputs of a full round          buf 0..1, i 0..T-1
//            z(k)       z_j                                    k = j mod 3
//            v(par)     V_j                                    par = j & 1
//            p(k, i)    p_i for round m, k = m mod 3           i 1..T-1
//   void put(slot, const uint32_t (&)[8]);  void get(uint32_t (&)[8], slot);
//   barriers: block()                       every role
//             z_arrive(k) / z_wait(k)       chain -> state roles and sum
//             v_arrive(par) / v_wait(par)   sum -> chain
//             p_arrive(k) / p_wait(k)       state roles -> sum
//             pro_arrive()  / pro_wait()    state roles -> sum, once, before round 0
// `s` on entry: role 0 the first state element (constant included), role i < T
// element i after absorb, role T ignored.  On return role 0 holds the canonical
// hash in `out`; the other roles return nothing.
template <int T, class Bus>
INF_HD void coop_hash(uint32_t (&out)[8], uint32_t (&s)[8], const int role, Bus& bus, const uint32_t* tbl) {
    using L = Layout<T>;
    constexpr int RP = L::RP;
    uint32_t xs[T][8];

    // ---- first half: rounds 0..3 (round 3 with the merged matrix) ----------------
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const uint32_t* m = tbl + (r < 3 ? L::FULL_M : L::PRE_M) * 8;
        const uint32_t* v = tbl + (r < 3 ? L::FULL_V + r * T : L::PRE_V) * 8;
        if (role < T) {
            uint32_t x[8];
            sbox(x, s);
            bus.put(bus.x(r & 1, role), x);
        }
        bus.block();
        if (role < T) {
#pragma unroll
            for (int i = 0; i < T; i++) bus.get(xs[i], bus.x(r & 1, i));
            dot<T, 8>(s, &xs[0][0], m + role * T * 8, v + role * 8);
        }
    }

    // ---- partial rounds -----------------------------------------------------------
    // Buffers: z_j and p^(m) rotate over THREE slots / barriers (index mod 3), V_j over two.
    // The chain can publish z_{j+2} before a state role has consumed z_j (u_{j+2} needs
    // V_{j+1}, which needs only p^(j+1), made from z_{j-1}); z_{j+3} it cannot (V_{j+2} needs
    // p^(j+2), made from z_j).  Likewise a state role can publish p^(j+2) while the sum role
    // is still adding up p^(j) and has not touched p^(j+1), but p^(j+3) needs z_{j+1}, hence
    // V_j delivered.  V_{j+2} needs z_{j+1}, hence V_j consumed: two slots.
    if (role == 0) {
        int j3 = 0;
#pragma unroll 1
        for (int j = 0; j < RP; j++) {
            uint32_t z[8], t[8];
            sbox(z, s);
            bus.put(bus.z(j3), z);
            bus.z_arrive(j3);
            bus.v_wait(j & 1);
            bus.get(t, bus.v(j & 1));
            add8(s, z, t);
            csub2p(s);
            j3 = j3 == 2 ? 0 : j3 + 1;
        }
    } else if (role < T) {
        {   // p for rounds 0 and 1 come from the state as it enters the section
            uint32_t t[8];
            mont_mul(t, s, coop_round<T>(tbl, 0).v + (role - 1) * 8);
            bus.put(bus.p(0, role), t);
            if (RP > 1) {
                mont_mul(t, s, coop_round<T>(tbl, 1).v + (role - 1) * 8);
                bus.put(bus.p(1, role), t);
            }
            bus.pro_arrive();
        }
        int j3 = 0;
#pragma unroll 1
        for (int j = 0; j < RP; j++) {
            const CoopRound<T> rc = coop_round<T>(tbl, j);
            uint32_t z[8], t[8];
            bus.z_wait(j3);
            bus.get(z, bus.z(j3));
            mont_mul(t, z, rc.w + (role - 1) * rc.ws);
            add8(s, s, t);
            csub2p(s);
            const int m3 = j3 == 0 ? 2 : j3 - 1;                // (j + 2) mod 3
            if (j + 2 < RP) {
                mont_mul(t, s, coop_round<T>(tbl, j + 2).v + (role - 1) * 8);
                bus.put(bus.p(m3, role), t);
                bus.p_arrive(m3);
            }
            j3 = j3 == 2 ? 0 : j3 + 1;
        }
        add8(s, s, tbl + (L::LAST_D + role - 1) * 8);       // remaining constants of the first tail round
        csub2p(s);
    } else {
        uint32_t t[8], a[8];
        bus.pro_wait();
        mont_redc(t, coop_round<T>(tbl, 0).k);               // k'_0 R^2 / R
#pragma unroll 1
        for (int i = 1; i < T; i++) {
            bus.get(a, bus.p(0, i));
            add8(t, t, a);
            csub2p(t);
        }
        bus.put(bus.v(0), t);
        bus.v_arrive(0);
        int j3 = 0;
#pragma unroll 1
        for (int j = 0; j < RP; j++) {
            uint32_t z[8];
            bus.z_wait(j3);
            if (j + 1 >= RP) break;
            const CoopRound<T> rc = coop_round<T>(tbl, j + 1);
            bus.get(z, bus.z(j3));
            mont_mul_add(t, z, rc.c, rc.k);                   // c_{j+1} z_j + k'_{j+1}
            const int m3 = j3 == 2 ? 0 : j3 + 1;                // (j + 1) mod 3
            if (j + 1 >= 2) bus.p_wait(m3);
#pragma unroll 1
            for (int i = 1; i < T; i++) {
                bus.get(a, bus.p(m3, i));
                add8(t, t, a);
                csub2p(t);
            }
            bus.put(bus.v((j + 1) & 1), t);
            bus.v_arrive((j + 1) & 1);
            j3 = m3;
        }
    }

    // ---- second half: 3 full rounds (the first takes the scale off u), output row ---
#pragma unroll 1
    for (int r = 0; r < 3; r++) {
        const uint32_t* m = tbl + (r == 0 ? L::TAIL0_M : L::FULL_M) * 8;
        const uint32_t* v = tbl + (L::TAIL_V + r * T) * 8;
        if (role < T) {
            uint32_t x[8];
            sbox(x, s);
            bus.put(bus.x(r & 1, role), x);
        }
        bus.block();
        if (role < T) {
#pragma unroll
            for (int i = 0; i < T; i++) bus.get(xs[i], bus.x(r & 1, i));
            dot<T, 8>(s, &xs[0][0], m + role * T * 8, v + role * 8);
        }
    }
    if (role < T) {
        uint32_t x[8];
        sbox(x, s);
        bus.put(bus.x(1, role), x);                           // rounds 0..2 left buffer 0 last
    }
    bus.block();
    if (role == 0) {
#pragma unroll
        for (int i = 0; i < T; i++) bus.get(xs[i], bus.x(1, i));
        dot<T, 8>(out, &xs[0][0], tbl + L::OUT_ROW * 8, nullptr);
        csub_p_exact(out);
        csub_p_exact(out);
    }
}

}  // namespace inf
