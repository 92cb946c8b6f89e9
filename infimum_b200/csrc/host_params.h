// Poseidon parameter generation and table derivation, done once at library
// init on the host.
//
//  * grain_params(t): the circom/BN254 round constants and MDS matrix for
//    width t, regenerated with the Grain-LFSR procedure of the Poseidon paper
//    (hadeshash generate_parameters_grain.sage 1 0 254 t 8 RP p) — the same
//    numbers the reference stores as literals in
//    pallet/src/hash/parameters.rs:43-43075 and hands out from
//    get_poseidon_parameters (parameters.rs:35).
//  * build_opt_table<T>(): the device table of poseidon.cuh's Layout<T>
//    (folded constants, sparse partial rounds).
//  * build_dense_table(t): (ark, mds) in Montgomery form for the generic
//    dense kernel that serves widths 9..13 and cross-checks the optimised one.
#pragma once
#include <vector>

#include "host_fr.h"

namespace inf {
namespace host {

struct DenseParams {
    int t = 0, rf = 8, rp = 0;
    std::vector<F> ark;   // (rf+rp)*t, indexed round*t + i   (poseidon.rs:126)
    std::vector<F> mds;   // t*t row-major, mds[i*t+j]         (poseidon.rs:153)
};

const DenseParams& grain_params(int t);            // 2 <= t <= 13, cached

// Layout<T> table, T in 2..8; returns empty vector for unsupported T.
std::vector<uint32_t> build_opt_table(int t);

// [ark (rf+rp)*t][mds t*t] each element 8 x u32 Montgomery form.
std::vector<uint32_t> build_dense_table(int t);

// Seeds of the zero tables (pallet/src/poll/zeroes.rs:2 and :38) and the empty
// ballot roots (zeroes.rs:73-79), 32-byte big-endian.  The chains
// Z[l+1] = H(Z[l] x arity) are recomputed on the device at init.
extern const uint8_t BINARY_ZERO_LEAF_BE[32];
extern const uint8_t QUINARY_ZERO_LEAF_BE[32];
extern const uint8_t EMPTY_BALLOT_ROOTS_BE[5][32];

}  // namespace host
}  // namespace inf
