// infimum_b200.hpp — C++ host-side mirror of the reference's interface for the
// hot path, over the C ABI in infimum_b200.h.
//
// The reference is Rust; no Rust toolchain exists in the build image, so the
// host side above the C ABI is written in C++ with the reference's names,
// argument meaning and error behaviour, so that parity tests read like the
// reference's own (tests/cpp/parity_tests.cpp vs pallet/src/tests/poseidon.rs
// and pallet/src/tests/extrinsics.rs):
//
//   reference (Rust)                                   here (C++)
//   Poseidon::<Fr>::new_circom(n)         poseidon.rs:304   infimum::Poseidon::new_circom(n)
//   Poseidon::with_domain_tag_circom      poseidon.rs:309   infimum::Poseidon::with_domain_tag_circom
//   PoseidonHasher::hash(&[Fr])           poseidon.rs:162   Poseidon::hash(std::vector<Fr>)
//   PoseidonBytesHasher::hash_bytes_be/le poseidon.rs:213   Poseidon::hash_bytes_be / hash_bytes_le
//   PoseidonError                         poseidon.rs:13    infimum::PoseidonError
//   PollStateTree{depth,..,root}          state.rs:70       infimum::PollStateTree (same fields)
//   AmortizedIncrementalMerkleTree::new/insert/merge/hash   state.rs:120   PollStateTree::new_/insert/merge/hash
//   MerkleTreeError -> u8                 state.rs:94-118   infimum::MerkleTreeError (same codes)
//   get_merkle_zeroes, EMPTY_BALLOT_ROOTS zeroes.rs:73-85   infimum::get_merkle_zeroes, empty_ballot_roots
//   PollProvider::merge_registrations     provider.rs:289   infimum::merge_registrations
//   PollProvider::merge_interactions      provider.rs:313   infimum::merge_interactions
//
// `Result<T,E>` is a minimal stand-in for Rust's.  Hashing only ever happens on
// the GPU; the single piece of host arithmetic is Fr::from_be_bytes_mod_order's
// reduction (<= 5 subtractions of p), needed so that Fr values compare equal.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "infimum_b200.h"

namespace infimum {

constexpr size_t HASH_LEN = 32;      // poseidon.rs:9
constexpr size_t MAX_X5_LEN = 13;    // poseidon.rs:10
using HashBytes = std::array<uint8_t, 32>;   // poll.rs:9

// ---- Result ---------------------------------------------------------------------------
template <class T, class E>
class Result {
    std::variant<T, E> v_;
public:
    Result(T t) : v_(std::in_place_index<0>, std::move(t)) {}
    Result(E e) : v_(std::in_place_index<1>, std::move(e)) {}
    bool is_ok() const { return v_.index() == 0; }
    bool is_err() const { return v_.index() == 1; }
    T& unwrap() {
        if (!is_ok()) throw std::runtime_error("called unwrap() on an Err value");
        return std::get<0>(v_);
    }
    const E& unwrap_err() const {
        if (!is_err()) throw std::runtime_error("called unwrap_err() on an Ok value");
        return std::get<1>(v_);
    }
    std::optional<T> ok() && { return is_ok() ? std::optional<T>(std::move(std::get<0>(v_))) : std::nullopt; }
};

// ---- errors ---------------------------------------------------------------------------
struct PoseidonError {                                   // poseidon.rs:13-31
    enum Kind { InvalidNumberOfInputs, EmptyInput, InvalidInputLength, BytesToPrimeFieldElement,
                InputLargerThanModulus, VecToArray, U64ToU8, BytesToBigInt, InvalidWidthCircom } kind;
    size_t inputs = 0, max_limit = 0, width = 0, len = 0, modulus_bytes_len = 0;
    bool operator==(const PoseidonError& o) const {
        return kind == o.kind && inputs == o.inputs && max_limit == o.max_limit && width == o.width &&
               len == o.len && modulus_bytes_len == o.modulus_bytes_len;
    }
    static PoseidonError empty_input() { return PoseidonError{EmptyInput}; }
    static PoseidonError invalid_input_length(size_t len) {
        PoseidonError e{InvalidInputLength};
        e.len = len; e.modulus_bytes_len = HASH_LEN;
        return e;
    }
    static PoseidonError invalid_number_of_inputs(size_t inputs, size_t width) {
        PoseidonError e{InvalidNumberOfInputs};
        e.inputs = inputs; e.max_limit = width - 1; e.width = width;
        return e;
    }
    static PoseidonError invalid_width_circom(size_t width) {
        PoseidonError e{InvalidWidthCircom};
        e.width = width; e.max_limit = MAX_X5_LEN;
        return e;
    }
};

enum class MerkleTreeError : uint8_t {                   // state.rs:94-118, u8 codes
    TreeAlreadyFull = 1, TreeAlreadyMerged = 2, HashFailed = 3, MergeFailed = 4
};
inline uint8_t to_u8(MerkleTreeError e) { return static_cast<uint8_t>(e); }

// ---- device context ---------------------------------------------------------------------
class Context {
    inf_ctx* ctx_ = nullptr;
public:
    explicit Context(int device = 0) {
        int rc = inf_init(device, &ctx_);
        if (rc != INF_OK)
            throw std::runtime_error(std::string("inf_init failed: ") + inf_strerror(rc) +
                                     " (infimum_b200 has no CPU fallback)");
    }
    ~Context() { inf_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    inf_ctx* get() const { return ctx_; }
    static Context& global() {           // one per process, like the shim in INTEGRATION.md
        static Context c(0);
        return c;
    }
};

// ---- Fr -----------------------------------------------------------------------------------
// A field element as its canonical 32-byte big-endian encoding.
struct Fr {
    HashBytes be{};
    static const HashBytes& modulus() {
        static const HashBytes p = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45,
                                    0xb6, 0x81, 0x81, 0x58, 0x5d, 0x28, 0x33, 0xe8, 0x48, 0x79, 0xb9,
                                    0x70, 0x91, 0x43, 0xe1, 0xf5, 0x93, 0xf0, 0x00, 0x00, 0x01};
        return p;
    }
    static Fr zero() { return Fr{}; }
    static Fr one() { return from(1); }
    static Fr from(uint64_t x) {
        Fr r;
        for (int i = 0; i < 8; i++) r.be[31 - i] = (uint8_t)(x >> (8 * i));
        return r;
    }
    // Fr::from_be_bytes_mod_order for up to 32 bytes (state.rs:290, tests/poseidon.rs:26)
    static Fr from_be_bytes_mod_order(const uint8_t* b, size_t n) {
        if (n > 32) throw std::invalid_argument("from_be_bytes_mod_order: more than 32 bytes");
        Fr r;
        memcpy(r.be.data() + (32 - n), b, n);
        const HashBytes& p = modulus();
        while (memcmp(r.be.data(), p.data(), 32) >= 0) {
            int borrow = 0;
            for (int i = 31; i >= 0; i--) {
                int d = (int)r.be[i] - (int)p[i] - borrow;
                borrow = d < 0;
                r.be[i] = (uint8_t)(d + (borrow << 8));
            }
        }
        return r;
    }
    template <class C> static Fr from_be_bytes_mod_order(const C& c) { return from_be_bytes_mod_order(c.data(), c.size()); }
    static Fr from_le_bytes_mod_order(const uint8_t* b, size_t n) {
        std::vector<uint8_t> t(b, b + n);
        for (size_t i = 0; i < n / 2; i++) std::swap(t[i], t[n - 1 - i]);
        return from_be_bytes_mod_order(t.data(), n);
    }
    HashBytes to_bytes_be() const { return be; }         // into_bigint().to_bytes_be()
    HashBytes to_bytes_le() const {
        HashBytes r;
        for (int i = 0; i < 32; i++) r[i] = be[31 - i];
        return r;
    }
    bool is_zero() const { for (uint8_t x : be) if (x) return false; return true; }
    bool operator==(const Fr& o) const { return be == o.be; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    std::string to_string() const {                      // decimal, like into_bigint().to_string()
        std::array<uint8_t, 32> t = be;
        std::string s;
        for (;;) {
            int rem = 0; bool nz = false;
            for (int i = 0; i < 32; i++) {
                int cur = rem * 256 + t[i];
                t[i] = (uint8_t)(cur / 10); rem = cur % 10;
                nz |= t[i] != 0;
            }
            s.insert(s.begin(), (char)('0' + rem));
            if (!nz) break;
        }
        return s;
    }
};

inline PoseidonError poseidon_error_from(int rc, size_t inputs, size_t width) {
    switch (rc) {
        case INF_ERR_INVALID_NUMBER_OF_INPUTS: return PoseidonError::invalid_number_of_inputs(inputs, width);
        case INF_ERR_EMPTY_INPUT: return PoseidonError::empty_input();
        case INF_ERR_INVALID_INPUT_LENGTH: return PoseidonError::invalid_input_length(0);
        case INF_ERR_INVALID_WIDTH_CIRCOM: return PoseidonError::invalid_width_circom(width);
        default: return PoseidonError{PoseidonError::BytesToBigInt};      // device failure
    }
}

// ---- Poseidon -----------------------------------------------------------------------------
// PoseidonParameters (poseidon.rs:33-71): ark indexed round*width + i, mds[i][j] row i column j.
struct PoseidonParameters {
    std::vector<Fr> ark;
    std::vector<std::vector<Fr>> mds;
    size_t full_rounds = 0, partial_rounds = 0, width = 0;
    uint64_t alpha = 5;
    PoseidonParameters(std::vector<Fr> ark_, std::vector<std::vector<Fr>> mds_, size_t full_rounds_,
                       size_t partial_rounds_, size_t width_, uint64_t alpha_)
        : ark(std::move(ark_)), mds(std::move(mds_)), full_rounds(full_rounds_), partial_rounds(partial_rounds_),
          width(width_), alpha(alpha_) {}
};

class Poseidon {
    size_t width_;
    Fr domain_tag_;
    Context* ctx_;
    // set when built from caller-supplied parameters (Poseidon::new(params), poseidon.rs:105-108)
    std::vector<uint8_t> ark_, mds_;
    size_t full_rounds_ = 0, partial_rounds_ = 0;
    uint64_t alpha_ = 5;
    bool custom_ = false;
    Poseidon(size_t width, Fr tag, Context* ctx) : width_(width), domain_tag_(tag), ctx_(ctx) {}
    int run(uint32_t flags, const uint8_t* tag, const uint8_t* rows, uint64_t n, uint8_t* out) const {
        if (custom_)
            return inf_poseidon_hash_batch_params(ctx_->get(), (uint32_t)width_, (uint32_t)full_rounds_,
                                                  (uint32_t)partial_rounds_, alpha_, ark_.data(), mds_.data(), flags,
                                                  tag, rows, n, out);
        return inf_poseidon_hash_batch(ctx_->get(), (uint32_t)(width_ - 1), flags, tag, rows, n, out);
    }
    const uint8_t* tag_ptr(HashBytes& scratch, bool le) const {
        if (domain_tag_.is_zero()) return nullptr;
        scratch = le ? domain_tag_.to_bytes_le() : domain_tag_.be;
        return scratch.data();
    }
public:
    static Result<Poseidon, PoseidonError> new_circom(size_t nr_inputs, Context* ctx = nullptr) {
        return with_domain_tag_circom(nr_inputs, Fr::zero(), ctx);
    }
    static Result<Poseidon, PoseidonError> with_domain_tag_circom(size_t nr_inputs, Fr domain_tag,
                                                                  Context* ctx = nullptr) {
        const size_t width = nr_inputs + 1;
        if (width > MAX_X5_LEN || width < 2) return PoseidonError::invalid_width_circom(width);
        return Poseidon(width, domain_tag, ctx ? ctx : &Context::global());
    }
    // Poseidon::new(params) / with_domain_tag(params, tag)  (poseidon.rs:105-118; `new` is a keyword here)
    static Result<Poseidon, PoseidonError> with_params(const PoseidonParameters& params, Fr domain_tag = Fr::zero(),
                                                       Context* ctx = nullptr) {
        if (params.width > MAX_X5_LEN || params.width < 2) return PoseidonError::invalid_width_circom(params.width);
        if (params.ark.size() != (params.full_rounds + params.partial_rounds) * params.width ||
            params.mds.size() != params.width)
            throw std::invalid_argument("PoseidonParameters: ark / mds sizes do not match the round counts and width");
        Poseidon h(params.width, domain_tag, ctx ? ctx : &Context::global());
        h.custom_ = true;
        h.full_rounds_ = params.full_rounds;
        h.partial_rounds_ = params.partial_rounds;
        h.alpha_ = params.alpha;
        for (const Fr& x : params.ark) h.ark_.insert(h.ark_.end(), x.be.begin(), x.be.end());
        for (const auto& row : params.mds) {
            if (row.size() != params.width) throw std::invalid_argument("PoseidonParameters: mds is not square");
            for (const Fr& x : row) h.mds_.insert(h.mds_.end(), x.be.begin(), x.be.end());
        }
        h.ark_.resize(h.ark_.size() + 32);          // never hand the library a null pointer for zero rounds
        return h;
    }
    size_t width() const { return width_; }

    // PoseidonHasher::hash
    Result<Fr, PoseidonError> hash(const std::vector<Fr>& inputs) {
        if (inputs.size() != width_ - 1) return PoseidonError::invalid_number_of_inputs(inputs.size(), width_);
        std::vector<uint8_t> buf(inputs.size() * 32);
        for (size_t i = 0; i < inputs.size(); i++) memcpy(&buf[32 * i], inputs[i].be.data(), 32);
        Fr out;
        HashBytes tag;
        int rc = run(0, tag_ptr(tag, false), buf.data(), 1, out.be.data());
        if (rc) return poseidon_error_from(rc, inputs.size(), width_);
        return out;
    }

    // n independent hashes over dense big-endian rows
    Result<std::vector<HashBytes>, PoseidonError> hash_batch(const uint8_t* rows, uint64_t n) {
        std::vector<HashBytes> out(n);
        HashBytes tag;
        int rc = run(0, tag_ptr(tag, false), rows, n, n ? out[0].data() : nullptr);
        if (rc) return poseidon_error_from(rc, width_ - 1, width_);
        return out;
    }

    // PoseidonBytesHasher
    using Slice = std::pair<const uint8_t*, size_t>;
    Result<HashBytes, PoseidonError> hash_bytes_be(const std::vector<Slice>& inputs) { return hash_bytes(inputs, 0); }
    Result<HashBytes, PoseidonError> hash_bytes_le(const std::vector<Slice>& inputs) {
        return hash_bytes(inputs, INF_FLAG_LITTLE_ENDIAN);
    }

private:
    Result<HashBytes, PoseidonError> hash_bytes(const std::vector<Slice>& inputs, uint32_t flags) {
        // every input is converted before hash() counts them (poseidon.rs:215-226)
        for (const Slice& s : inputs) {
            if (s.second == 0) return PoseidonError::empty_input();
            if (s.second != HASH_LEN) return PoseidonError::invalid_input_length(s.second);
        }
        if (inputs.size() != width_ - 1) return PoseidonError::invalid_number_of_inputs(inputs.size(), width_);
        HashBytes out, tag;
        if (custom_) {
            std::vector<uint8_t> buf;
            for (const Slice& s : inputs) buf.insert(buf.end(), s.first, s.first + 32);
            int rc = run(flags, tag_ptr(tag, flags != 0), buf.data(), 1, out.data());
            if (rc) return poseidon_error_from(rc, inputs.size(), width_);
            return out;
        }
        std::vector<const uint8_t*> ptrs;
        std::vector<size_t> lens;
        for (const Slice& s : inputs) { ptrs.push_back(s.first); lens.push_back(s.second); }
        int rc = inf_poseidon_hash_bytes(ctx_->get(), flags, tag_ptr(tag, flags != 0), ptrs.data(), lens.data(),
                                         (uint32_t)inputs.size(), out.data());
        if (rc) return poseidon_error_from(rc, inputs.size(), width_);
        return out;
    }
};

// ---- zero tables ------------------------------------------------------------------------------
inline std::array<HashBytes, 33> get_merkle_zeroes(uint8_t arity, Context* ctx = nullptr) {   // zeroes.rs:81-85
    std::array<HashBytes, 33> z;
    inf_merkle_zeroes((ctx ? ctx : &Context::global())->get(), arity, z[0].data());
    return z;
}
inline std::array<HashBytes, 5> empty_ballot_roots() {                                         // zeroes.rs:73-79
    std::array<HashBytes, 5> r;
    inf_empty_ballot_roots(r[0].data());
    return r;
}

// ---- PollStateTree ------------------------------------------------------------------------------
// Same fields as the reference struct (state.rs:70-91).  What has been hashed in so far is
// held the way the pallet stores it — the frontier `hashes` — plus a buffer of leaves not yet
// folded in, so a tree can be resumed from persisted state (from_state; what the pallet reads
// back from storage, lib.rs:706-714) and never needs its leaf history.  The reference hashes
// inside insert(); here insert() buffers the leaf and the device folds the whole buffer into
// the frontier in one go (inf_tree_append) when frontier() asks for it, when the last insert
// completes the tree, or in merge().  A tree that still is what new_() made is reduced by
// merge() in one shot (inf_tree_merge), any other from its frontier (inf_tree_merge_frontier).
// `hashes` is current after frontier(), merge() or a completing insert.
struct PollStateTree {
    uint8_t depth = 0;
    uint8_t full_depth = 0;
    uint8_t arity = 0;
    uint32_t count = 0;
    std::vector<std::pair<uint8_t, HashBytes>> hashes;    // the stored frontier; empty once merged
    std::optional<HashBytes> root;

    static PollStateTree new_(uint8_t arity, uint8_t full_depth,
                              std::optional<std::pair<uint8_t, HashBytes>> zero_hash, Context* ctx = nullptr) {
        PollStateTree t;
        t.arity = arity;
        t.full_depth = full_depth;
        t.ctx_ = ctx ? ctx : &Context::global();
        if (zero_hash) {                                                   // state.rs:150-158
            t.hashes.push_back(*zero_hash);
            // the pallet only ever seeds (0, zeroes[0]) (state.rs:48-52): that case has the one-shot merge
            t.blank_ = zero_hash->first == 0 && zero_hash->second == get_merkle_zeroes(arity, t.ctx_)[0];
            t.fresh_ = t.blank_;
        }
        return t;
    }

    // Resume from the persisted struct: insert / merge continue from the stored frontier.
    static PollStateTree from_state(uint8_t arity, uint8_t full_depth, uint8_t depth, uint32_t count,
                                    std::vector<std::pair<uint8_t, HashBytes>> hashes, std::optional<HashBytes> root,
                                    Context* ctx = nullptr) {
        PollStateTree t;
        t.arity = arity;
        t.full_depth = full_depth;
        t.depth = depth;
        t.count = count;
        t.hashes = std::move(hashes);
        t.root = root;
        t.ctx_ = ctx ? ctx : &Context::global();
        t.fresh_ = false;
        t.depth_flushed_ = depth;
        return t;
    }

    Result<PollStateTree, MerkleTreeError> insert(const HashBytes& leaf) && {
        return std::move(*this).extend(leaf.data(), 1);
    }
    // bulk insert of n leaves (dense 32-byte rows), in insertion order
    Result<PollStateTree, MerkleTreeError> extend(const uint8_t* leaves, uint64_t n) && {
        if (root) return MerkleTreeError::TreeAlreadyFull;                 // state.rs:182
        const unsigned __int128 cap = capacity();
        if ((unsigned __int128)total() + n > cap) return MerkleTreeError::TreeAlreadyFull;
        if (pending() == 0) depth_flushed_ = depth;
        leaves_.insert(leaves_.end(), leaves, leaves + 32 * n);
        count += (uint32_t)n;
        uint8_t d = 0;
        while (d < full_depth && ipow(d + 1) <= total()) d++;
        if (d > depth) depth = d;                                          // state.rs:212-213
        if ((unsigned __int128)total() == cap) {                           // state.rs:218-222
            int rc = fresh_ ? run_merge(true) : flush();
            if (rc != INF_OK && rc != INF_ERR_TREE_ALREADY_MERGED) return tree_error(rc);
        }
        return std::move(*this);
    }
    Result<PollStateTree, MerkleTreeError> merge(bool to_depth) && {       // state.rs:230-281
        if (root) return MerkleTreeError::TreeAlreadyMerged;
        if (fresh_) {
            int rc = run_merge(to_depth);
            if (rc != INF_OK) return tree_error(rc);
            return std::move(*this);
        }
        int rc = flush();
        if (rc != INF_OK && rc != INF_ERR_TREE_ALREADY_MERGED) return tree_error(rc);
        if (root) return std::move(*this);
        std::vector<uint8_t> lv, hs;
        pack(lv, hs);
        HashBytes r;
        int has = 0;
        uint32_t rdepth = 0;
        rc = inf_tree_merge_frontier(ctx_->get(), arity, full_depth, lv.data(), hs.data(), (uint32_t)lv.size(), to_depth,
                                     r.data(), &has, &rdepth);
        if (rc) return tree_error(rc);
        if (has) {
            root = r;
            hashes.clear();
        }
        return std::move(*this);
    }
    // state.rs:284-302
    static Result<HashBytes, PoseidonError> hash(const std::vector<HashBytes>& inputs, Context* ctx = nullptr) {
        auto h = Poseidon::new_circom(inputs.size(), ctx);
        if (h.is_err()) return h.unwrap_err();
        std::vector<Poseidon::Slice> s;
        for (const HashBytes& b : inputs) s.push_back({b.data(), b.size()});
        return h.unwrap().hash_bytes_be(s);
    }
    // Bring `hashes` (the reference's persisted frontier) up to date with every leaf inserted so far.
    Result<PollStateTree, MerkleTreeError> frontier() && {
        int rc = flush();
        if (rc != INF_OK && rc != INF_ERR_TREE_ALREADY_MERGED) return tree_error(rc);
        return std::move(*this);
    }

private:
    Context* ctx_ = nullptr;
    bool blank_ = false;              // seeded with the reference's blank leaf (0, zeroes[0])
    bool fresh_ = true;               // the frontier is still what new_() made: one-shot merge applies
    uint8_t depth_flushed_ = 0;       // `depth` as of the stored frontier (before the buffered leaves)
    std::vector<uint8_t> leaves_;     // leaves not yet folded into `hashes`
    uint64_t pending() const { return leaves_.size() / 32; }
    uint64_t logical() const {        // leaves the frontier stands for (the blank leaf included)
        unsigned __int128 n = 0;
        for (const auto& e : hashes) n += ipow(e.first);
        return n > (unsigned __int128)UINT64_MAX ? UINT64_MAX : (uint64_t)n;
    }
    uint64_t total() const { return logical() + pending(); }
    unsigned __int128 capacity() const {
        unsigned __int128 c = 1;
        for (int i = 0; i < full_depth; i++) c *= arity;
        return c;
    }
    uint64_t ipow(int e) const {
        unsigned __int128 c = 1;
        for (int i = 0; i < e; i++) { c *= arity; if (c > (unsigned __int128)UINT64_MAX) return UINT64_MAX; }
        return (uint64_t)c;
    }
    static MerkleTreeError tree_error(int rc) {
        return (rc >= 1 && rc <= 4) ? static_cast<MerkleTreeError>(rc) : MerkleTreeError::HashFailed;
    }
    void pack(std::vector<uint8_t>& lv, std::vector<uint8_t>& hs) const {
        for (const auto& e : hashes) {
            lv.push_back(e.first);
            hs.insert(hs.end(), e.second.begin(), e.second.end());
        }
    }
    // Fold the buffered leaves into the frontier: insert() x pending on the stored state.
    int flush() {
        if (pending() == 0 || root) return INF_OK;
        std::vector<uint8_t> lv, hs;
        pack(lv, hs);
        uint8_t out_lv[4 * 33];
        std::vector<uint8_t> out_hs(4 * 33 * 32);
        uint32_t n = 0, d = 0;
        int has = 0;
        HashBytes r;
        int rc = inf_tree_append(ctx_->get(), arity, full_depth, lv.data(), hs.data(), (uint32_t)lv.size(), depth_flushed_,
                                 leaves_.data(), pending(), out_lv, out_hs.data(), 4 * 33, &n, &d, &has, r.data());
        if (rc != INF_OK && rc != INF_ERR_TREE_ALREADY_MERGED) return rc;
        hashes.clear();
        for (uint32_t i = 0; i < n; i++) {
            HashBytes h;
            memcpy(h.data(), &out_hs[32 * i], 32);
            hashes.push_back({out_lv[i], h});
        }
        leaves_.clear();
        leaves_.shrink_to_fit();
        fresh_ = false;
        depth = (uint8_t)d;
        depth_flushed_ = depth;
        if (has) root = r;
        return rc;
    }
    // One-shot new + insert x N + merge over the buffered leaves (fresh trees only).
    int run_merge(bool to_depth) {
        HashBytes r;
        uint32_t idepth = 0, rdepth = 0;
        int has = 0;
        int rc = inf_tree_merge(ctx_->get(), arity, full_depth, blank_, to_depth, leaves_.data(), pending(), r.data(),
                                &idepth, &rdepth, &has);
        if (rc == INF_OK || rc == INF_ERR_TREE_ALREADY_MERGED) {
            if (has) {
                root = r;
                hashes.clear();
                leaves_.clear();
                leaves_.shrink_to_fit();
            }
            depth = (uint8_t)idepth;
        }
        return rc;
    }
};

// PollState::new (state.rs:42-67): the two trees of a poll
inline PollStateTree new_registration_tree(uint8_t registration_depth, Context* ctx = nullptr) {
    return PollStateTree::new_(2, registration_depth, std::make_pair((uint8_t)0, get_merkle_zeroes(2, ctx)[0]), ctx);
}
inline PollStateTree new_interaction_tree(uint8_t interaction_depth, Context* ctx = nullptr) {
    return PollStateTree::new_(5, interaction_depth, std::nullopt, ctx);
}

// ---- provider.rs:289-327 ------------------------------------------------------------------------------
struct Commitment {                                      // coordinator.rs (subset used here)
    std::pair<uint32_t, HashBytes> process{0, HashBytes{}};
    std::pair<uint32_t, HashBytes> tally{0, HashBytes{}};
    uint32_t expected_process = 0, expected_tally = 0;
};

// merge_registrations: registrations.merge(false), then
// commitment.process = (0, H3(root, EMPTY_BALLOT_ROOTS[1], 0))
inline Result<std::pair<PollStateTree, Commitment>, MerkleTreeError>
merge_registrations(PollStateTree registrations, Commitment commitment = {}) {
    auto m = std::move(registrations).merge(false);
    if (m.is_err()) return m.unwrap_err();
    PollStateTree t = std::move(m.unwrap());
    if (!t.root) return MerkleTreeError::MergeFailed;
    auto h = PollStateTree::hash({*t.root, empty_ballot_roots()[1], HashBytes{}});
    if (h.is_err()) return MerkleTreeError::HashFailed;
    commitment.process = {0, h.unwrap()};
    return std::make_pair(std::move(t), commitment);
}

// merge_interactions: interactions.merge(true) and the expected proof counts
inline Result<std::pair<PollStateTree, Commitment>, MerkleTreeError>
merge_interactions(PollStateTree interactions, uint32_t registrations_count, uint8_t process_subtree_depth,
                   uint8_t tally_subtree_depth, Commitment commitment = {}) {
    auto m = std::move(interactions).merge(true);
    if (m.is_err()) return m.unwrap_err();
    PollStateTree t = std::move(m.unwrap());
    uint32_t pb = 1, tb = 1;
    for (int i = 0; i < process_subtree_depth; i++) pb *= t.arity;
    for (int i = 0; i < tally_subtree_depth; i++) tb *= 2;
    commitment.expected_process = t.count / pb + ((t.count % pb) ? 1 : 0);
    commitment.expected_tally = 1 + registrations_count / tb;
    return std::make_pair(std::move(t), commitment);
}


// ---- keys and leaf hashing (keys.rs, provider.rs:218-287) -------------------------------------------
struct PublicKey {                                       // keys.rs: PublicKey { x, y }
    HashBytes x{}, y{};
};
using PollInteractionData = std::array<HashBytes, 10>;   // poll.rs

// Bulk forms of the leaf computations, one participant / message per GPU thread.
inline Result<std::vector<HashBytes>, MerkleTreeError>
registration_leaves(const std::vector<PublicKey>& keys, const std::vector<uint64_t>& timestamps, Context* ctx = nullptr) {
    if (keys.size() != timestamps.size()) return MerkleTreeError::HashFailed;
    std::vector<HashBytes> out(keys.size());
    static_assert(sizeof(PublicKey) == 64, "PublicKey must be two packed 32-byte coordinates");
    int rc = inf_registration_leaves((ctx ? ctx : &Context::global())->get(),
                                     keys.empty() ? nullptr : keys[0].x.data(), timestamps.data(), keys.size(),
                                     out.empty() ? nullptr : out[0].data());
    if (rc) return MerkleTreeError::HashFailed;
    return out;
}
inline Result<std::vector<HashBytes>, MerkleTreeError>
interaction_leaves(const std::vector<PublicKey>& keys, const std::vector<PollInteractionData>& data, Context* ctx = nullptr) {
    if (keys.size() != data.size()) return MerkleTreeError::HashFailed;
    std::vector<HashBytes> out(keys.size());
    int rc = inf_interaction_leaves((ctx ? ctx : &Context::global())->get(),
                                    keys.empty() ? nullptr : keys[0].x.data(),
                                    data.empty() ? nullptr : data[0][0].data(), keys.size(),
                                    out.empty() ? nullptr : out[0].data());
    if (rc) return MerkleTreeError::HashFailed;
    return out;
}

// The slice of `Poll` that drives the hot path: the two trees, the commitment,
// and register_participant / consume_interaction / merge_* with the reference's
// signatures (provider.rs:218-327).  Leaves are hashed in bulk at merge time.
// What prepare_public_inputs returns, without the verify key (Groth16 is out of scope):
// which circuit the inputs are for, the inputs, and the commitment the proof would install.
struct PublicInputs {
    bool process = true;                 // false: tally circuit
    std::vector<Fr> inputs;
    Commitment commitment;
};

struct Poll {
    PollStateTree registrations, interactions;
    Commitment commitment;
    uint8_t process_subtree_depth = 0, tally_subtree_depth = 0;
    uint64_t created_at = 0, signup_period = 0, voting_period = 0;      // poll.rs / config.rs, for get_voting_period_end

    Poll(uint8_t registration_depth, uint8_t interaction_depth, uint8_t process_subtree_depth_,
         uint8_t tally_subtree_depth_, Context* ctx = nullptr)
        : registrations(new_registration_tree(registration_depth, ctx)),
          interactions(new_interaction_tree(interaction_depth, ctx)),
          process_subtree_depth(process_subtree_depth_), tally_subtree_depth(tally_subtree_depth_), ctx_(ctx) {}

    Result<uint32_t, MerkleTreeError> register_participant(const PublicKey& pk, uint64_t timestamp) {
        if (registrations.root) return MerkleTreeError::TreeAlreadyFull;
        reg_keys_.push_back(pk);
        reg_ts_.push_back(timestamp);
        return registrations.count + (uint32_t)reg_keys_.size();
    }
    Result<uint32_t, MerkleTreeError> consume_interaction(const PublicKey& pk, const PollInteractionData& data) {
        if (interactions.root) return MerkleTreeError::TreeAlreadyFull;
        msg_keys_.push_back(pk);
        msg_data_.push_back(data);
        return interactions.count + (uint32_t)msg_keys_.size();
    }
    std::optional<MerkleTreeError> merge_registrations() {
        if (auto e = flush()) return e;
        auto r = infimum::merge_registrations(std::move(registrations), commitment);
        if (r.is_err()) return r.unwrap_err();
        registrations = std::move(r.unwrap().first);
        commitment = r.unwrap().second;
        return std::nullopt;
    }
    std::optional<MerkleTreeError> merge_interactions() {
        if (auto e = flush()) return e;
        auto r = infimum::merge_interactions(std::move(interactions), registrations.count, process_subtree_depth,
                                             tally_subtree_depth, commitment);
        if (r.is_err()) return r.unwrap_err();
        interactions = std::move(r.unwrap().first);
        commitment = r.unwrap().second;
        return std::nullopt;
    }

    uint64_t get_voting_period_end() const { return created_at + signup_period + voting_period; }   // provider.rs:355-358

    // prepare_public_inputs (provider.rs:141-215); u32 arithmetic as in the reference
    std::optional<PublicInputs> prepare_public_inputs(const PublicKey& coordinator_public_key,
                                                      const HashBytes& new_commitment) const {
        PublicInputs out;
        uint32_t message_batch_size = 1;
        for (int i = 0; i < process_subtree_depth; i++) message_batch_size *= interactions.arity;
        uint32_t current_batch_index = interactions.count;
        if (current_batch_index > 0) {
            const uint32_t r = interactions.count % message_batch_size;
            current_batch_index -= r == 0 ? message_batch_size : r;
        }
        uint32_t proof_index = commitment.process.first;
        const uint32_t index_offset = proof_index * message_batch_size;
        out.commitment = commitment;
        if (index_offset <= current_batch_index) {
            auto hasher = Poseidon::new_circom(2, ctx_);
            if (hasher.is_err()) return std::nullopt;
            auto h = hasher.unwrap().hash({Fr::from_be_bytes_mod_order(coordinator_public_key.x),
                                           Fr::from_be_bytes_mod_order(coordinator_public_key.y)});
            if (h.is_err() || !interactions.root) return std::nullopt;
            current_batch_index -= index_offset;
            uint32_t end_batch_index = current_batch_index + message_batch_size;
            if (end_batch_index > interactions.count) end_batch_index = interactions.count;
            out.inputs = {Fr::from(registrations.count + 1), Fr::from(get_voting_period_end()),
                          Fr::from_be_bytes_mod_order(*interactions.root), Fr::from(registrations.depth),
                          Fr::from(end_batch_index), Fr::from(current_batch_index), h.unwrap(),
                          Fr::from_be_bytes_mod_order(commitment.process.second),
                          Fr::from_be_bytes_mod_order(new_commitment)};
            out.commitment.process = {proof_index + 1, new_commitment};
            return out;
        }
        out.process = false;
        proof_index = commitment.tally.first;
        uint32_t batch_size = 1;
        for (int i = 0; i < tally_subtree_depth; i++) batch_size *= registrations.arity;
        current_batch_index = proof_index * batch_size;
        if (current_batch_index >= registrations.count + 1) return std::nullopt;
        out.inputs = {Fr::from_be_bytes_mod_order(commitment.process.second),
                      Fr::from_be_bytes_mod_order(commitment.tally.second), Fr::from_be_bytes_mod_order(new_commitment),
                      Fr::from(current_batch_index), Fr::from(registrations.count + 1)};
        out.commitment.tally = {proof_index + 1, new_commitment};
        return out;
    }

private:
    Context* ctx_;
    std::vector<PublicKey> reg_keys_, msg_keys_;
    std::vector<uint64_t> reg_ts_;
    std::vector<PollInteractionData> msg_data_;
    std::optional<MerkleTreeError> flush() {
        if (!reg_keys_.empty()) {
            auto lv = registration_leaves(reg_keys_, reg_ts_, ctx_);
            if (lv.is_err()) return lv.unwrap_err();
            auto t = std::move(registrations).extend(lv.unwrap()[0].data(), lv.unwrap().size());
            if (t.is_err()) return t.unwrap_err();
            registrations = std::move(t.unwrap());
            reg_keys_.clear(); reg_ts_.clear();
        }
        if (!msg_keys_.empty()) {
            auto lv = interaction_leaves(msg_keys_, msg_data_, ctx_);
            if (lv.is_err()) return lv.unwrap_err();
            auto t = std::move(interactions).extend(lv.unwrap()[0].data(), lv.unwrap().size());
            if (t.is_err()) return t.unwrap_err();
            interactions = std::move(t.unwrap());
            msg_keys_.clear(); msg_data_.clear();
        }
        return std::nullopt;
    }
};

// ---- Merkle paths and outcome verification (provider.rs:76-139, 396-436) ------------------------------
using MerklePath = std::vector<std::vector<HashBytes>>;     // [level][arity-1 siblings]

// compute_merkle_root_from_path (arity 5, VOTE_TREE_ARITY)
inline std::optional<HashBytes> compute_merkle_root_from_path(uint8_t depth, uint32_t index, const HashBytes& leaf,
                                                              const MerklePath& path, Context* ctx = nullptr) {
    if (path.size() < depth) return std::nullopt;
    std::vector<uint8_t> flat;
    for (uint8_t l = 0; l < depth; l++) {
        if (path[l].size() < 4) return std::nullopt;
        for (int k = 0; k < 4; k++) flat.insert(flat.end(), path[l][k].begin(), path[l][k].end());
    }
    uint64_t idx = index;
    HashBytes root;
    int rc = inf_merkle_roots_from_paths((ctx ? ctx : &Context::global())->get(), 5, depth, &idx, leaf.data(),
                                         flat.data(), 1, root.data());
    if (rc) return std::nullopt;
    return root;
}

struct PollOutcome {                                        // coordinator.rs:53-77
    std::vector<uint32_t> tally_results;
    std::vector<MerklePath> tally_result_proofs;
    HashBytes total_spent{}, total_spent_salt{}, tally_result_salt{}, new_results_commitment{}, spent_votes_hash{};
};

// verify_outcome without the `is_proven` guard; the per-option work is batched
// (one path-root launch over all vote options, then two hash2 batches).
inline std::optional<uint32_t> verify_outcome(uint8_t vote_option_tree_depth, size_t n_options,
                                              const HashBytes& tally_commitment, const PollOutcome& o,
                                              Context* ctx = nullptr) {
    Context* c = ctx ? ctx : &Context::global();
    if (o.tally_results.size() < n_options || o.tally_result_proofs.size() < n_options) return std::nullopt;
    const uint8_t d = vote_option_tree_depth;
    std::vector<uint64_t> idx(n_options);
    std::vector<uint8_t> leaves(32 * n_options, 0), flat;
    for (size_t i = 0; i < n_options; i++) {
        idx[i] = i;
        const uint32_t r = o.tally_results[i];
        for (int b = 0; b < 4; b++) leaves[32 * i + 28 + b] = (uint8_t)(r >> (24 - 8 * b));
        if (o.tally_result_proofs[i].size() < d) return std::nullopt;
        for (uint8_t l = 0; l < d; l++) {
            if (o.tally_result_proofs[i][l].size() < 4) return std::nullopt;
            for (int k = 0; k < 4; k++)
                flat.insert(flat.end(), o.tally_result_proofs[i][l][k].begin(), o.tally_result_proofs[i][l][k].end());
        }
    }
    std::vector<HashBytes> roots(n_options);
    if (inf_merkle_roots_from_paths(c->get(), 5, d, idx.data(), leaves.data(), flat.data(), n_options, roots[0].data()))
        return std::nullopt;
    auto h2 = Poseidon::new_circom(2, c).unwrap();
    std::vector<uint8_t> rows(64 * n_options);
    for (size_t i = 0; i < n_options; i++) {
        memcpy(&rows[64 * i], roots[i].data(), 32);
        memcpy(&rows[64 * i + 32], o.tally_result_salt.data(), 32);
    }
    auto a = h2.hash_batch(rows.data(), n_options);
    if (a.is_err()) return std::nullopt;
    for (size_t i = 0; i < n_options; i++) {
        memcpy(&rows[64 * i], a.unwrap()[i].data(), 32);
        memcpy(&rows[64 * i + 32], o.spent_votes_hash.data(), 32);
    }
    auto b = h2.hash_batch(rows.data(), n_options);
    if (b.is_err()) return std::nullopt;
    for (size_t i = 0; i < n_options; i++)
        if (b.unwrap()[i] != tally_commitment) return std::nullopt;
    auto t1 = PollStateTree::hash({o.total_spent, o.total_spent_salt}, c);
    if (t1.is_err()) return std::nullopt;
    auto t2 = PollStateTree::hash({o.new_results_commitment, t1.unwrap()}, c);
    if (t2.is_err() || t2.unwrap() != tally_commitment) return std::nullopt;
    uint32_t best = 0, best_val = 0;
    for (size_t i = 0; i < n_options; i++)
        if (o.tally_results[i] > best_val) { best = (uint32_t)i; best_val = o.tally_results[i]; }
    return best;
}

// ---- retained tree: every level on the device, bulk sibling paths ----------------------------------------
class RetainedTree {
    inf_tree* t_ = nullptr;
    uint32_t arity_, depth_;
public:
    RetainedTree(uint8_t arity, uint8_t depth, const uint8_t* leaves, uint64_t n, bool prepend_blank_leaf = false,
                 Context* ctx = nullptr) : arity_(arity), depth_(depth) {
        int rc = inf_tree_build((ctx ? ctx : &Context::global())->get(), arity, depth, prepend_blank_leaf, leaves, n, &t_);
        if (rc) throw std::runtime_error(std::string("inf_tree_build: ") + inf_strerror(rc));
    }
    ~RetainedTree() { inf_tree_destroy(t_); }
    RetainedTree(const RetainedTree&) = delete;
    RetainedTree& operator=(const RetainedTree&) = delete;
    HashBytes root() const { HashBytes r; inf_tree_root(t_, r.data()); return r; }
    std::vector<MerklePath> paths(const std::vector<uint64_t>& leaf_indices) const {
        std::vector<uint8_t> flat(leaf_indices.size() * depth_ * (arity_ - 1) * 32);
        int rc = inf_tree_paths(t_, leaf_indices.data(), leaf_indices.size(), flat.data());
        if (rc) throw std::runtime_error(std::string("inf_tree_paths: ") + inf_strerror(rc));
        std::vector<MerklePath> out(leaf_indices.size(), MerklePath(depth_, std::vector<HashBytes>(arity_ - 1)));
        size_t k = 0;
        for (auto& p : out) for (auto& lvl : p) for (auto& h : lvl) { memcpy(h.data(), &flat[32 * k++], 32); }
        return out;
    }
};

}  // namespace infimum
