// C++ parity tests through the C++ host mirror (include/infimum_b200.hpp) and
// the C ABI.  Test names and bodies follow the reference's own tests:
//   pallet/src/tests/poseidon.rs      fr_one, bytes_ones_twos, with_domain_tag, fr_one_two,
//                                     random_input, empty_input, circomlibjs_compat_1_to_12_inputs
//   pallet/src/tests/extrinsics.rs    merge_registration_state_success, merge_interaction_state_success,
//                                     process_messages_public_signals, participant_limit_reached
// plus oracle parity for ragged trees and the frontier.  Golden values come
// from tests/golden/reference_vectors.json via the generated golden_vectors.h.
// The oracle (liboracle.so) is linked as the checker only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <string>
#include <tuple>

#include "golden_vectors.h"
#include "infimum_b200.hpp"

extern "C" {
int oracle_hash(int n_inputs, const uint8_t* in, const uint8_t* tag_be, uint8_t* out, int faithful);
int oracle_tree_insert_merge(int arity, int full_depth, int blank, int to_depth, const uint8_t* leaves, uint64_t n,
                             uint8_t* root, uint32_t* out_state, int faithful);
}

using namespace infimum;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                   \
    do {                                                                              \
        g_checks++;                                                                   \
        if (!(cond)) { g_failed++; printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static HashBytes HB(const uint8_t* p) { HashBytes h; memcpy(h.data(), p, 32); return h; }
static HashBytes be32(uint64_t x) { return Fr::from(x).be; }
static HashBytes rev(HashBytes h) { for (int i = 0; i < 16; i++) std::swap(h[i], h[31 - i]); return h; }

// ---- pallet/src/tests/poseidon.rs -------------------------------------------------------------
static void fr_one() {
    auto hasher = Poseidon::new_circom(2).unwrap();
    const uint8_t one1[] = {1}, one2[] = {0, 1}, one3[] = {0, 0, 1};
    for (auto [p, n] : {std::pair<const uint8_t*, size_t>{one1, 1}, {one2, 2}, {one3, 3}}) {
        Fr input = Fr::from_be_bytes_mod_order(p, n);
        Fr hash = hasher.hash({input, input}).unwrap();
        CHECK(hash.to_bytes_be() == HB(G_FR_ONE_EXPECTED_BE));
    }
}

static void bytes_ones_twos() {
    HashBytes ones, twos;
    ones.fill(1); twos.fill(2);
    Fr input1 = Fr::from_be_bytes_mod_order(ones), input2 = Fr::from_be_bytes_mod_order(twos);
    auto hasher = Poseidon::new_circom(2).unwrap();
    CHECK(hasher.hash({input1, input2}).unwrap().to_bytes_be() == HB(G_BYTES_ONES_TWOS_BE));
    CHECK(hasher.hash_bytes_be({{ones.data(), 32}, {twos.data(), 32}}).unwrap() == HB(G_BYTES_ONES_TWOS_BE));
    CHECK(hasher.hash_bytes_le({{ones.data(), 32}, {twos.data(), 32}}).unwrap() == HB(G_BYTES_ONES_TWOS_LE));
}

static void with_domain_tag() {
    HashBytes ones, twos;
    ones.fill(1); twos.fill(2);
    Fr input1 = Fr::from_be_bytes_mod_order(ones), input2 = Fr::from_be_bytes_mod_order(twos);
    auto hasher = Poseidon::with_domain_tag_circom(2, Fr::zero()).unwrap();
    CHECK(hasher.hash({input1, input2}).unwrap().to_bytes_be() == HB(G_WITH_DOMAIN_TAG_ZERO_BE));
    auto tagged = Poseidon::with_domain_tag_circom(2, Fr::one()).unwrap();
    HashBytes got = tagged.hash({input1, input2}).unwrap().to_bytes_be();
    CHECK(got != HB(G_WITH_DOMAIN_TAG_ZERO_BE));
    uint8_t buf[64], exp[32];
    memcpy(buf, input1.be.data(), 32); memcpy(buf + 32, input2.be.data(), 32);
    oracle_hash(2, buf, be32(1).data(), exp, 0);
    CHECK(got == HB(exp));
}

static void fr_one_two() {
    auto hasher = Poseidon::new_circom(2).unwrap();
    const uint8_t a[] = {1}, b[] = {2};
    Fr hash = hasher.hash({Fr::from_be_bytes_mod_order(a, 1), Fr::from_be_bytes_mod_order(b, 1)}).unwrap();
    CHECK(hash.to_bytes_le() == HB(G_FR_ONE_TWO_LE));
}

static void random_input() {      // both inputs exceed the modulus
    Fr input1 = Fr::from_be_bytes_mod_order(G_RANDOM_INPUT_1, 32), input2 = Fr::from_be_bytes_mod_order(G_RANDOM_INPUT_2, 32);
    auto hasher = Poseidon::new_circom(2).unwrap();
    CHECK(hasher.hash({input1, input2}).unwrap().to_bytes_le() == HB(G_RANDOM_INPUT_LE));
    // the byte path reduces on the device
    CHECK(rev(hasher.hash_bytes_be({{G_RANDOM_INPUT_1, 32}, {G_RANDOM_INPUT_2, 32}}).unwrap()) == HB(G_RANDOM_INPUT_LE));
}

static void empty_input() {
    const uint8_t non_empty[32] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    for (size_t nr_inputs = 1; nr_inputs < 12; nr_inputs++) {
        auto hasher = Poseidon::new_circom(nr_inputs).unwrap();
        std::vector<Poseidon::Slice> inputs(nr_inputs, {nullptr, 0});            // all inputs empty
        CHECK(hasher.hash_bytes_be(inputs).unwrap_err() == PoseidonError::empty_input());
        CHECK(hasher.hash_bytes_le(inputs).unwrap_err() == PoseidonError::empty_input());
        std::vector<Poseidon::Slice> one_empty(nr_inputs - 1, {non_empty, 32});  // one empty input
        one_empty.push_back({nullptr, 0});
        CHECK(hasher.hash_bytes_be(one_empty).unwrap_err() == PoseidonError::empty_input());
        CHECK(hasher.hash_bytes_le(one_empty).unwrap_err() == PoseidonError::empty_input());
    }
    auto hasher = Poseidon::new_circom(2).unwrap();
    uint8_t long33[33] = {0};
    CHECK(hasher.hash_bytes_be({{long33, 33}, {non_empty, 32}}).unwrap_err() == PoseidonError::invalid_input_length(33));
    CHECK(hasher.hash_bytes_be({{non_empty, 31}, {non_empty, 32}}).unwrap_err() == PoseidonError::invalid_input_length(31));
    CHECK(hasher.hash({Fr::one()}).unwrap_err() == PoseidonError::invalid_number_of_inputs(1, 3));
    CHECK(Poseidon::new_circom(13).unwrap_err() == PoseidonError::invalid_width_circom(14));
    CHECK(Poseidon::new_circom(0).is_err());
}

static void circomlibjs_compat_1_to_12_inputs() {
    HashBytes one = be32(1), two = be32(2);
    for (size_t i = 1; i < 13; i++) {
        auto hasher = Poseidon::new_circom(i).unwrap();
        std::vector<Poseidon::Slice> ones(i, {one.data(), 32}), twos(i, {two.data(), 32});
        CHECK(hasher.hash_bytes_be(ones).unwrap() == HB(G_CIRCOMLIBJS[i - 1]));
        CHECK(hasher.hash_bytes_be(twos).unwrap() != HB(G_CIRCOMLIBJS[i - 1]));
    }
}

// Poseidon::new(params) with the parameters.rs tables (fetched from the library's own Grain
// regeneration) must be new_circom: pinned by the circomlibjs vectors.
static void custom_parameters_equal_new_circom() {
    HashBytes one = be32(1);
    for (size_t n_inputs : {1u, 2u, 5u, 12u}) {
        const size_t t = n_inputs + 1;
        static const int RP[12] = {56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65};
        const size_t n_ark = (8 + RP[t - 2]) * t, n_el = n_ark + t * t;
        std::vector<uint32_t> w(n_el * 8);
        CHECK(inf_debug_dense_params((uint32_t)t, w.data(), w.size()) == (int)n_el);
        auto el = [&](size_t k) {
            Fr r;
            for (int i = 0; i < 8; i++)
                for (int b = 0; b < 4; b++) r.be[31 - (4 * i + b)] = (uint8_t)(w[8 * k + i] >> (8 * b));
            return r;
        };
        std::vector<Fr> ark;
        for (size_t k = 0; k < n_ark; k++) ark.push_back(el(k));
        std::vector<std::vector<Fr>> mds(t);
        for (size_t i = 0; i < t; i++)
            for (size_t j = 0; j < t; j++) mds[i].push_back(el(n_ark + i * t + j));
        auto hasher = Poseidon::with_params(PoseidonParameters(ark, mds, 8, RP[t - 2], t, 5)).unwrap();
        std::vector<Poseidon::Slice> ones(n_inputs, {one.data(), 32});
        CHECK(hasher.hash_bytes_be(ones).unwrap() == HB(G_CIRCOMLIBJS[n_inputs - 1]));
        CHECK(hasher.hash(std::vector<Fr>(n_inputs, Fr::one())).unwrap().be == HB(G_CIRCOMLIBJS[n_inputs - 1]));
        CHECK(hasher.hash({}).is_err());
    }
}

// ---- pallet/src/poll/zeroes.rs -----------------------------------------------------------------
static void zero_tables() {
    auto b = get_merkle_zeroes(2), q = get_merkle_zeroes(5), other = get_merkle_zeroes(7);
    for (int l = 0; l < 33; l++) {
        CHECK(b[l] == HB(G_BINARY_ZEROES[l]));
        CHECK(q[l] == HB(G_QUINARY_ZEROES[l]));
        CHECK(other[l] == q[l]);
    }
    auto e = empty_ballot_roots();
    for (int i = 0; i < 5; i++) CHECK(e[i] == HB(G_EMPTY_BALLOT_ROOTS[i]));
    CHECK(PollStateTree::hash({b[3], b[3]}).unwrap() == b[4]);
    CHECK(PollStateTree::hash({q[3], q[3], q[3], q[3], q[3]}).unwrap() == q[4]);
}

// ---- pallet/src/tests/extrinsics.rs ------------------------------------------------------------
static HashBytes registration_leaf(const uint8_t* x, const uint8_t* y, uint64_t block) {   // provider.rs:224-233
    auto h = Poseidon::new_circom(4).unwrap();
    return h.hash({Fr::from_be_bytes_mod_order(x, 32), Fr::from_be_bytes_mod_order(y, 32), Fr::from(1), Fr::from(block)})
        .unwrap().to_bytes_be();
}
static PollStateTree registered_tree() {
    PollStateTree t = new_registration_tree(G_REGISTRATION_DEPTH);
    for (int i = 0; i < 3; i++)
        t = std::move(t).insert(registration_leaf(G_PARTICIPANTS[i][0], G_PARTICIPANTS[i][1], G_REGISTRATION_BLOCK)).unwrap();
    return t;
}

static void merge_registration_state_success() {
    auto r = merge_registrations(registered_tree()).unwrap();
    CHECK(r.first.root == std::optional<HashBytes>(HB(G_REGISTRATIONS_ROOT)));
    CHECK(r.second.process == std::make_pair(0u, HB(G_PROCESS_COMMITMENT)));
    CHECK(r.first.hashes.empty());
}

static void merge_interaction_state_success() {
    auto h5 = Poseidon::new_circom(5).unwrap(), h4 = Poseidon::new_circom(4).unwrap();
    std::vector<Fr> left, right;                                   // consume_interaction, provider.rs:249-278
    for (int i = 0; i < 5; i++) left.push_back(Fr::from_be_bytes_mod_order(G_MESSAGE[i], 32));
    for (int i = 5; i < 10; i++) right.push_back(Fr::from_be_bytes_mod_order(G_MESSAGE[i], 32));
    Fr l = h5.hash(left).unwrap(), r = h5.hash(right).unwrap();
    HashBytes leaf = h4.hash({l, r, Fr::from_be_bytes_mod_order(G_SHARED_PK[0], 32), Fr::from_be_bytes_mod_order(G_SHARED_PK[1], 32)})
                         .unwrap().to_bytes_be();
    PollStateTree t = new_interaction_tree(G_INTERACTION_DEPTH);
    t = std::move(t).insert(leaf).unwrap();
    auto m = merge_interactions(std::move(t), 3, G_PROCESS_SUBTREE_DEPTH, G_TALLY_SUBTREE_DEPTH).unwrap();
    CHECK(m.first.root == std::optional<HashBytes>(HB(G_INTERACTIONS_ROOT)));
    CHECK(m.second.expected_process == G_EXPECTED_PROCESS);
    CHECK(m.second.expected_tally == G_EXPECTED_TALLY);
}

static void process_messages_public_signals() {
    auto r = merge_registrations(registered_tree()).unwrap();
    CHECK(r.first.count + 1 == 4);
    CHECK(r.first.depth == G_REGISTRATIONS_DEPTH);
    auto hasher = Poseidon::new_circom(2).unwrap();
    Fr coord = hasher.hash({Fr::from_be_bytes_mod_order(G_COORDINATOR_PK[0], 32), Fr::from_be_bytes_mod_order(G_COORDINATOR_PK[1], 32)}).unwrap();
    CHECK(coord.to_string() == G_COORD_PUB_KEY_HASH_DECIMAL);
    CHECK(r.second.process == std::make_pair(0u, HB(G_PROCESS_COMMITMENT)));
}

static void participant_limit_reached() {     // depth-2 tree: blank + 3 leaves is completed by insert
    PollStateTree t = new_registration_tree(2);
    for (int i = 0; i < 3; i++) t = std::move(t).insert(be32(100 + i)).unwrap();
    CHECK(t.root.has_value() && t.hashes.empty());
    PollStateTree copy = t;
    CHECK(std::move(copy).merge(false).unwrap_err() == MerkleTreeError::TreeAlreadyMerged);
    CHECK(to_u8(std::move(t).insert(be32(7)).unwrap_err()) == 1);
}

// ---- oracle parity: ragged trees and the frontier ------------------------------------------------
static void tree_and_frontier_vs_oracle() {
    std::mt19937_64 rng(0x494E46);
    for (auto [arity, full_depth, blank, to_depth] : {std::tuple<int, int, bool, bool>{2, 12, true, false}, {5, 5, false, true},
                                                     {2, 12, false, true}, {5, 5, true, false}}) {
        for (uint64_t n : {0ull, 1ull, 2ull, 3ull, 4ull, 5ull, 6ull, 24ull, 25ull, 26ull, 63ull, 64ull, 100ull, 124ull, 125ull,
                           126ull, 624ull, 625ull, 1000ull, 2047ull}) {
            if (n + blank >= (uint64_t)std::pow(arity, full_depth)) continue;
            std::vector<uint8_t> leaves(32 * n);
            for (auto& b : leaves) b = (uint8_t)rng();
            uint8_t exp[32];
            uint32_t st[3];
            int rc = oracle_tree_insert_merge(arity, full_depth, blank, to_depth, leaves.data(), n, exp, st, 0);
            auto zero = blank ? std::optional(std::make_pair((uint8_t)0, get_merkle_zeroes(arity)[0])) : std::nullopt;
            PollStateTree t = PollStateTree::new_(arity, full_depth, zero);
            t = std::move(t).extend(leaves.data(), n).unwrap();
            CHECK(t.depth == st[0] && t.count == st[1]);
            // frontier: merging the frontier by hand (state.rs:240-271) must give the same root
            PollStateTree f = std::move(PollStateTree(t)).frontier().unwrap();
            for (size_t i = 1; i < f.hashes.size(); i++) CHECK(f.hashes[i - 1].first >= f.hashes[i].first);
            t = std::move(t).merge(to_depth).unwrap();
            CHECK(rc == 0);
            if (st[2]) CHECK(t.root == std::optional<HashBytes>(HB(exp)));
            else CHECK(!t.root.has_value());
            // replay merge() over the device-computed frontier with single hashes
            auto Z = get_merkle_zeroes(arity);
            auto hs = f.hashes;
            while (!hs.empty()) {
                uint8_t d = hs.back().first;
                if (hs.size() == 1 && (!to_depth || d == full_depth)) break;
                std::vector<HashBytes> run;
                while (!hs.empty() && hs.back().first == d) { run.insert(run.begin(), hs.back().second); hs.pop_back(); }
                while ((int)run.size() < arity) run.push_back(Z[d]);
                hs.push_back({(uint8_t)(d + 1), PollStateTree::hash(run).unwrap()});
            }
            if (st[2]) CHECK(hs.size() == 1 && hs[0].second == HB(exp));
        }
    }
}

// ---- the stored tree: persist {depth, count, hashes, root}, resume, insert more, merge ---------------
// (what every extrinsic does through storage, lib.rs:706-714; state.rs:176-225, 230-281)
static void insert_and_merge_on_a_stored_frontier() {
    std::mt19937_64 rng(0x46524F4E54);
    for (auto [arity, full_depth, blank, to_depth] : {std::tuple<int, int, bool, bool>{2, 11, true, false}, {5, 5, false, true},
                                                     {2, 11, false, true}, {5, 5, true, false}}) {
        const uint64_t cap = (uint64_t)std::pow(arity, full_depth);
        for (uint64_t n : {1ull, 2ull, 5ull, 26ull, 127ull, 700ull, 2046ull}) {
            if (n + blank > cap) continue;
            std::vector<uint8_t> leaves(32 * n);
            for (auto& b : leaves) b = (uint8_t)rng();
            uint8_t exp[32];
            uint32_t st[3];
            int rc = oracle_tree_insert_merge(arity, full_depth, blank, to_depth, leaves.data(), n, exp, st, 0);
            CHECK(rc == 0);
            auto zero = blank ? std::optional(std::make_pair((uint8_t)0, get_merkle_zeroes(arity)[0])) : std::nullopt;
            for (uint64_t k : {(uint64_t)0, n / 3, n / 2, n - 1, n}) {
                // first session: k leaves, state brought up to date and "persisted"
                PollStateTree a = PollStateTree::new_(arity, full_depth, zero);
                a = std::move(a).extend(leaves.data(), k).unwrap();
                a = std::move(a).frontier().unwrap();
                // second session: resumed from the stored fields only
                PollStateTree b = PollStateTree::from_state(a.arity, a.full_depth, a.depth, a.count, a.hashes, a.root);
                if (b.root) {                      // the first k leaves already completed the tree
                    CHECK(k + blank == cap);
                    continue;
                }
                b = std::move(b).extend(leaves.data() + 32 * k, n - k).unwrap();
                CHECK(b.count == st[1]);
                if (!b.root) b = std::move(b).merge(to_depth).unwrap();
                CHECK(b.depth == st[0]);
                if (st[2]) CHECK(b.root == std::optional<HashBytes>(HB(exp)));
                CHECK(b.hashes.empty() == b.root.has_value());
            }
        }
    }
    // a resumed tree that is full rejects the next leaf, a merged one rejects merge (state.rs:178-182, 236)
    PollStateTree t = new_registration_tree(2);
    t = std::move(t).extend(be32(1).data(), 1).unwrap();
    t = std::move(t).frontier().unwrap();
    PollStateTree r = PollStateTree::from_state(t.arity, t.full_depth, t.depth, t.count, t.hashes, t.root);
    std::vector<uint8_t> two(64, 7);
    r = std::move(r).extend(two.data(), 2).unwrap();                 // blank + 3 = 4 = 2^2: completed by insert
    CHECK(r.root.has_value() && r.hashes.empty() && r.depth == 2);
    PollStateTree c = r;
    CHECK(std::move(c).merge(false).unwrap_err() == MerkleTreeError::TreeAlreadyMerged);
    CHECK(to_u8(std::move(r).insert(be32(9)).unwrap_err()) == 1);
}

// ---- leaf hashing through the Poll mirror (provider.rs:218-287) -----------------------------------
static void poll_register_interact_merge() {
    Poll poll(G_REGISTRATION_DEPTH, G_INTERACTION_DEPTH, G_PROCESS_SUBTREE_DEPTH, G_TALLY_SUBTREE_DEPTH);
    for (int i = 0; i < 3; i++) {
        PublicKey pk{HB(G_PARTICIPANTS[i][0]), HB(G_PARTICIPANTS[i][1])};
        CHECK(poll.register_participant(pk, G_REGISTRATION_BLOCK).unwrap() == (uint32_t)(i + 1));
    }
    CHECK(!poll.merge_registrations().has_value());
    CHECK(poll.registrations.root == std::optional<HashBytes>(HB(G_REGISTRATIONS_ROOT)));
    CHECK(poll.commitment.process == std::make_pair(0u, HB(G_PROCESS_COMMITMENT)));
    CHECK(poll.registrations.depth == G_REGISTRATIONS_DEPTH);
    PollInteractionData data;
    for (int i = 0; i < 10; i++) data[i] = HB(G_MESSAGE[i]);
    CHECK(poll.consume_interaction(PublicKey{HB(G_SHARED_PK[0]), HB(G_SHARED_PK[1])}, data).unwrap() == 1);
    CHECK(!poll.merge_interactions().has_value());
    CHECK(poll.interactions.root == std::optional<HashBytes>(HB(G_INTERACTIONS_ROOT)));
    CHECK(poll.commitment.expected_process == G_EXPECTED_PROCESS && poll.commitment.expected_tally == G_EXPECTED_TALLY);
    // the nine public inputs of the first process-messages proof (extrinsics.rs:621-633)
    poll.created_at = G_CREATED_AT; poll.signup_period = G_SIGNUP_PERIOD; poll.voting_period = G_VOTING_PERIOD;
    PublicKey coordinator{HB(G_COORDINATOR_PK[0]), HB(G_COORDINATOR_PK[1])};
    auto pi = poll.prepare_public_inputs(coordinator, HB(G_EXPECTED_PUBLIC_INPUTS[8]));
    CHECK(pi.has_value() && pi->process && pi->inputs.size() == 9);
    if (pi && pi->inputs.size() == 9)
        for (int i = 0; i < 9; i++) CHECK(pi->inputs[i].be == HB(G_EXPECTED_PUBLIC_INPUTS[i]));
    CHECK(pi && pi->commitment.process == std::make_pair(1u, HB(G_EXPECTED_PUBLIC_INPUTS[8])));
    if (pi) poll.commitment = pi->commitment;
    auto ti = poll.prepare_public_inputs(coordinator, HashBytes{});
    CHECK(ti.has_value() && !ti->process && ti->inputs.size() == 5 && ti->inputs[0].be == HB(G_EXPECTED_PUBLIC_INPUTS[8]) &&
          ti->inputs[4].be == be32(4) && ti->commitment.tally.first == 1);
}

// ---- verify_outcome against the reference's scenario fixtures (data.rs:241-275) ---------------------
static void verify_outcome_scenarios() {
    for (int sid = 0; sid < 2; sid++) {
        PollOutcome o;
        for (int i = 0; i < 25; i++) {
            o.tally_results.push_back(G_OUTCOME_TALLY_RESULTS[sid][i]);
            MerklePath p(2, std::vector<HashBytes>(4));
            for (int l = 0; l < 2; l++) for (int k = 0; k < 4; k++) p[l][k] = HB(G_OUTCOME_PROOFS[sid][i][l][k]);
            o.tally_result_proofs.push_back(p);
        }
        o.total_spent = HB(G_OUTCOME_FIELDS[sid][0]); o.total_spent_salt = HB(G_OUTCOME_FIELDS[sid][1]);
        o.tally_result_salt = HB(G_OUTCOME_FIELDS[sid][2]); o.new_results_commitment = HB(G_OUTCOME_FIELDS[sid][3]);
        o.spent_votes_hash = HB(G_OUTCOME_FIELDS[sid][4]);
        const HashBytes commitment = HB(G_OUTCOME_FIELDS[sid][5]);
        CHECK(verify_outcome(2, 25, commitment, o) == std::optional<uint32_t>(G_OUTCOME_EXPECTED[sid]));
        PollOutcome bad = o;
        bad.tally_result_salt[31] ^= 1;
        CHECK(!verify_outcome(2, 25, commitment, bad).has_value());
    }
}

// ---- retained tree + paths round trip ------------------------------------------------------------------
static void retained_tree_paths() {
    std::mt19937_64 rng(7);
    std::vector<uint8_t> leaves(32 * 611);
    for (auto& b : leaves) b = (uint8_t)rng();
    RetainedTree tree(5, 4, leaves.data(), 611);
    uint8_t exp[32];
    uint32_t st[3];
    oracle_tree_insert_merge(5, 4, 0, 1, leaves.data(), 611, exp, st, 0);
    CHECK(tree.root() == HB(exp));
    std::vector<uint64_t> idx = {0, 1, 4, 5, 124, 125, 300, 610};
    auto paths = tree.paths(idx);
    for (size_t k = 0; k < idx.size(); k++) {
        auto r = compute_merkle_root_from_path(4, (uint32_t)idx[k], HB(&leaves[32 * idx[k]]), paths[k]);
        CHECK(r == std::optional<HashBytes>(tree.root()));
    }
}

int main() {
    struct T { const char* name; std::function<void()> fn; };
    const T tests[] = {
        {"fr_one", fr_one}, {"bytes_ones_twos", bytes_ones_twos}, {"with_domain_tag", with_domain_tag},
        {"fr_one_two", fr_one_two}, {"random_input", random_input}, {"empty_input", empty_input},
        {"circomlibjs_compat_1_to_12_inputs", circomlibjs_compat_1_to_12_inputs},
        {"custom_parameters_equal_new_circom", custom_parameters_equal_new_circom}, {"zero_tables", zero_tables},
        {"merge_registration_state_success", merge_registration_state_success},
        {"merge_interaction_state_success", merge_interaction_state_success},
        {"process_messages_public_signals", process_messages_public_signals},
        {"participant_limit_reached", participant_limit_reached},
        {"tree_and_frontier_vs_oracle", tree_and_frontier_vs_oracle},
        {"insert_and_merge_on_a_stored_frontier", insert_and_merge_on_a_stored_frontier},
        {"poll_register_interact_merge", poll_register_interact_merge},
        {"verify_outcome_scenarios", verify_outcome_scenarios},
        {"retained_tree_paths", retained_tree_paths},
    };
    for (const T& t : tests) {
        int before = g_failed;
        try { t.fn(); } catch (const std::exception& e) { g_failed++; printf("  EXCEPTION in %s: %s\n", t.name, e.what()); }
        printf("test %s ... %s\n", t.name, g_failed == before ? "ok" : "FAILED");
    }
    printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
