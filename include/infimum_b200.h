/* infimum_b200 — C ABI of the B200-native Poseidon-BN254 hasher and poll-tree
 * merge, the drop-in boundary for the one data-parallel hot path of
 * rhysbalevicius/infimum (citations are relative to the reference checkout).
 *
 * Every entry point below is what a `gpu` (std-only) feature of the pallet
 * would bind over FFI in place of the reference function named beside it; the
 * Rust shim a maintainer would add is in INTEGRATION.md.
 *
 * Conventions
 *   - Field elements travel as 32-byte big-endian `HashBytes`
 *     (pallet/src/poll/poll.rs:9).  Inputs may be any 256-bit value and are
 *     reduced mod p exactly like `Fr::from_be_bytes_mod_order`
 *     (pallet/src/poll/state.rs:290); outputs are canonical (< p).
 *   - Arrays are dense row-major: n hashes of k inputs = n*k*32 bytes.
 *   - The caller owns every buffer; the library keeps no pointer after return.
 *   - Calls on one inf_ctx must be serialised by the caller (like `&mut self`,
 *     pallet/src/hash/poseidon.rs:78); distinct contexts are independent.
 *   - Device pointers handed to `_dev` calls must be 16-byte aligned (the kernels
 *     move nodes with 128-bit loads and stores); a misaligned pointer is refused
 *     with INF_ERR_BAD_ALIGNMENT before anything is launched.
 *   - Host-buffer calls are synchronous.  `_dev` calls take device pointers,
 *     enqueue on the given CUDA stream (a cudaStream_t passed as void*; NULL =
 *     the context's own non-blocking stream, so pass cudaStreamLegacy /
 *     cudaStreamPerThread explicitly to mean a default stream) and return
 *     without synchronising unless noted.
 *   - There is no CPU fallback: without a usable CUDA device inf_init fails.
 *
 * Return codes (int): 0 ok.
 *   1..4    MerkleTreeError as u8, pallet/src/poll/state.rs:106-118
 *   16..31  PoseidonError / argument errors, pallet/src/hash/poseidon.rs:13-31
 *   64..    CUDA failure (the Rust side maps these to HashFailed = 3)
 */
#ifndef INFIMUM_B200_H
#define INFIMUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INF_OK 0
/* MerkleTreeError -> u8 (state.rs:106-118) */
#define INF_ERR_TREE_ALREADY_FULL 1
#define INF_ERR_TREE_ALREADY_MERGED 2
#define INF_ERR_HASH_FAILED 3
#define INF_ERR_MERGE_FAILED 4
/* PoseidonError (poseidon.rs:13-31) */
#define INF_ERR_INVALID_NUMBER_OF_INPUTS 16
#define INF_ERR_EMPTY_INPUT 17
#define INF_ERR_INVALID_INPUT_LENGTH 18
#define INF_ERR_INVALID_WIDTH_CIRCOM 19
/* argument errors with no reference counterpart */
#define INF_ERR_NULL_POINTER 24
#define INF_ERR_BAD_ARITY 25
#define INF_ERR_BAD_DEPTH 26
#define INF_ERR_BUFFER_TOO_SMALL 27 /* an output array (frontier entries) has too few slots */
#define INF_ERR_BAD_FRONTIER 28     /* not a state insert() can leave behind (state.rs:176-225) */
#define INF_ERR_BAD_ALIGNMENT 29    /* device pointers must be 16-byte aligned (128-bit loads) */
/* device */
#define INF_ERR_NO_DEVICE 64
#define INF_ERR_CUDA 65
#define INF_ERR_OUT_OF_MEMORY 66
#define INF_ERR_NCCL 67

/* flags */
#define INF_FLAG_LITTLE_ENDIAN 1u /* hash_bytes_le wire order (poseidon.rs:233-250) */
#define INF_MULTI_PEER_COPY 1u    /* inf_multi_init: gather subtree roots with peer copies instead of NCCL */

typedef struct inf_ctx inf_ctx;

/* ---- lifetime -------------------------------------------------------------- */

/* Create a context on CUDA device `device` (ordinal).  Derives the Poseidon
 * tables (what get_poseidon_parameters returns, parameters.rs:35, plus the
 * optimised-schedule tables) and the Merkle zero tables (zeroes.rs:1-71) and
 * uploads them. */
int inf_init(int device, inf_ctx** out);
void inf_destroy(inf_ctx* ctx);
/* Static description of a return code. */
const char* inf_strerror(int code);
/* Text of the last CUDA error seen by this context ("" if none). */
const char* inf_last_cuda_error(const inf_ctx* ctx);
/* Library version / build info. */
const char* inf_version(void);

/* ---- host buffers ---------------------------------------------------------------
 * Host-buffer calls accept any host memory.  Page-locked buffers are copied by
 * DMA straight from / into the caller's memory; pageable buffers (a Vec<u8>, a
 * numpy array) are staged through pinned memory inside the library by helper
 * threads, which costs host memory bandwidth but keeps the pipeline full
 * (bench.py: e2e.hash2_2^22_pageable next to e2e.value).  A caller that keeps its
 * buffers for many calls can take them from inf_host_alloc (cudaHostAlloc,
 * portable) or pin its own with inf_host_register (cudaHostRegister; pinning
 * costs about as much as one copy, so it pays only for buffers that are reused). */
int inf_host_alloc(inf_ctx* ctx, size_t bytes, void** out);
int inf_host_free(inf_ctx* ctx, void* p);
int inf_host_register(inf_ctx* ctx, void* p, size_t bytes);
int inf_host_unregister(inf_ctx* ctx, void* p);

/* ---- Poseidon hasher --------------------------------------------------------
 * Replaces Poseidon::<Fr>::new_circom(n_inputs) / with_domain_tag_circom
 * (poseidon.rs:302-327) followed by PoseidonHasher::hash (poseidon.rs:162-208)
 * or PoseidonBytesHasher::hash_bytes_be / _le (poseidon.rs:213-250), applied
 * to `n` independent input tuples.
 *   n_inputs     1..12 (width t = n_inputs+1 <= 13) else INVALID_WIDTH_CIRCOM
 *   domain_tag   32 bytes in the same wire order as the inputs, or NULL for 0
 *   in           n * n_inputs * 32 bytes
 *   out          n * 32 bytes
 */
int inf_poseidon_hash_batch(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                            const uint8_t* domain_tag, const uint8_t* in, uint64_t n,
                            uint8_t* out);
int inf_poseidon_hash_batch_dev(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                                const uint8_t* domain_tag /* host */, const void* d_in, uint64_t n,
                                void* d_out, void* stream);

/* Poseidon::new(params) (poseidon.rs:47-71, 105-108): the same hash with
 * caller-supplied PoseidonParameters instead of the circom tables.
 *   ark   (full_rounds + partial_rounds) * width elements, index round*width + i
 *         (poseidon.rs:126)
 *   mds   width * width elements, row-major: mds[i][j] at i*width + j (poseidon.rs:153)
 * both 32-byte big-endian (values >= p are reduced).  Rounds 0 .. full_rounds/2 - 1
 * and the last full_rounds - full_rounds/2 are full, the partial_rounds in between
 * apply x^alpha to state[0] only (poseidon.rs:184-203).  width 2..13 (else
 * INVALID_WIDTH_CIRCOM); at most 4096 rounds.  Runs the reference schedule
 * literally (the dense kernel): correct for any parameters, not tuned. */
int inf_poseidon_hash_batch_params(inf_ctx* ctx, uint32_t width, uint32_t full_rounds,
                                   uint32_t partial_rounds, uint64_t alpha, const uint8_t* ark,
                                   const uint8_t* mds, uint32_t flags, const uint8_t* domain_tag,
                                   const uint8_t* in, uint64_t n, uint8_t* out);

/* Byte-slice front end with the reference's length checks, one hash:
 * inputs[i] has lens[i] bytes.  Empty -> EMPTY_INPUT, any other length than 32
 * -> INVALID_INPUT_LENGTH (validate_bytes_length + bytes_to_prime_field_element,
 * poseidon.rs:255-300); n_inputs != width-1 cannot happen here by construction. */
int inf_poseidon_hash_bytes(inf_ctx* ctx, uint32_t flags, const uint8_t* domain_tag,
                            const uint8_t* const* inputs, const size_t* lens, uint32_t n_inputs,
                            uint8_t out[32]);

/* Same hash through the generic dense kernel (reference schedule taken
 * literally, separate tables).  Diagnostic: lets callers cross-check the
 * optimised kernels on the device. */
int inf_poseidon_hash_batch_dense(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                                  const uint8_t* domain_tag, const uint8_t* in, uint64_t n,
                                  uint8_t* out);

/* ---- leaf hashing (the step before the trees) ------------------------------------
 * Registration leaves, PollProvider::register_participant (provider.rs:218-241):
 *     leaf[i] = hash4(pk[i].x, pk[i].y, 1, timestamp[i])
 * Interaction leaves, PollProvider::consume_interaction (provider.rs:243-287):
 *     leaf[i] = hash4(hash5(d[i][0..5]), hash5(d[i][5..10]), pk[i].x, pk[i].y)
 *   public_keys  n * 64 bytes: `PublicKey {x, y}` (keys.rs), 32-byte big-endian each
 *   timestamps   n * u64 (the block number the reference passes)
 *   data         n * 320 bytes: `PollInteractionData = [[u8;32];10]` (poll.rs)
 *   leaves       n * 32 bytes out
 * The `_dev` forms take device pointers and enqueue on `stream`. */
int inf_registration_leaves(inf_ctx* ctx, const uint8_t* public_keys, const uint64_t* timestamps,
                            uint64_t n, uint8_t* leaves);
int inf_interaction_leaves(inf_ctx* ctx, const uint8_t* public_keys, const uint8_t* data, uint64_t n,
                           uint8_t* leaves);
int inf_registration_leaves_dev(inf_ctx* ctx, const void* d_public_keys, const void* d_timestamps,
                                uint64_t n, void* d_leaves, void* stream);
int inf_interaction_leaves_dev(inf_ctx* ctx, const void* d_public_keys, const void* d_data,
                               uint64_t n, void* d_leaves, void* stream);

/* ---- Merkle zero tables -------------------------------------------------------
 * get_merkle_zeroes(arity) (zeroes.rs:81-85): 33 x 32 bytes; arity 2 -> binary
 * table, anything else -> quinary table.  EMPTY_BALLOT_ROOTS: 5 x 32 bytes. */
int inf_merkle_zeroes(inf_ctx* ctx, uint32_t arity, uint8_t out[33 * 32]);
int inf_empty_ballot_roots(uint8_t out[5 * 32]);

/* ---- poll tree ------------------------------------------------------------------
 * One-shot equivalent of
 *     PollStateTree::new(arity, full_depth, zero_hash)        state.rs:142-170
 *       .insert(leaf) for every leaf                          state.rs:176-225
 *       .merge(to_depth)                                      state.rs:230-281
 * over the full leaf array.
 *   arity               2 or 5 (the two trees PollState::new builds, state.rs:42-67)
 *   prepend_blank_leaf  non-zero: tree is seeded with zeroes[0] at leaf index 0,
 *                       as the registration tree is (state.rs:48-52)
 *   to_depth            merge's argument: pad with zero siblings up to full_depth
 *   leaves              n_leaves * 32 bytes
 * Outputs (any may be NULL):
 *   root          32 bytes
 *   insert_depth  PollStateTree.depth after the inserts (public signal
 *                 actualStateTreeDepth, provider.rs:182)
 *   root_depth    number of levels under root
 * Errors: more leaves than arity^full_depth -> TREE_ALREADY_FULL (what insert
 * returns); exactly arity^full_depth leaves -> TREE_ALREADY_MERGED (insert
 * completed the tree, so merge refuses, state.rs:236); in that case `root`
 * is still written so the caller can mirror `root: Some(..)`.  Zero leaves in
 * total (no blank leaf either): root is left untouched, root_depth = 0 and the
 * call returns INF_OK with *has_root = 0 (merge on an empty frontier leaves
 * root = None, state.rs:240-248).
 */
int inf_tree_merge(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                   int to_depth, const uint8_t* leaves, uint64_t n_leaves, uint8_t root[32],
                   uint32_t* insert_depth, uint32_t* root_depth, int* has_root);
/* Same with the leaves already in device memory; synchronises before return. */
int inf_tree_merge_dev(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                       int to_depth, const void* d_leaves, uint64_t n_leaves, uint8_t root[32],
                       uint32_t* insert_depth, uint32_t* root_depth, int* has_root, void* stream);

/* Building block for sharded trees: reduce a run of consecutive nodes of level
 * `level_in` (device memory) by `n_levels` levels, padding the right edge of
 * level l with zeroes[l].  The run is `shift` copies of zeroes[level_in]
 * followed by the n_in nodes at d_in (shift = 1 on the rank that owns leaf 0 of
 * a registration tree — the blank state leaf, state.rs:48-52 — else 0).
 * Writes ceil((shift + n_in) / arity^n_levels) nodes to d_out (device) and that
 * count to *n_out.  Enqueued on `stream`; does not synchronise.  Scratch is
 * owned by the context. */
int inf_tree_reduce_dev(inf_ctx* ctx, uint32_t arity, uint32_t level_in, uint32_t n_levels,
                        uint64_t shift, const void* d_in, uint64_t n_in, void* d_out,
                        uint64_t* n_out, void* stream);

/* State of the tree after the inserts, WITHOUT merging: the frontier the
 * reference persists in `PollStateTree.hashes` (state.rs:85-86) — the (level,
 * hash) pairs left by new(..) + insert(leaf) for every leaf (state.rs:176-225),
 * levels non-increasing towards the tail — plus `depth` and, if the inserts
 * completed the tree (state.rs:218-222), the root (then the frontier is empty).
 *   out_levels   cap bytes, out_hashes cap*32 bytes; cap >= 4*33 always suffices
 *   n_entries    number of frontier entries written
 * Errors: TREE_ALREADY_FULL if more leaves than arity^full_depth; BUFFER_TOO_SMALL
 * if the frontier has more than `cap` entries. */
int inf_tree_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                      const uint8_t* leaves, uint64_t n_leaves, uint8_t* out_levels,
                      uint8_t* out_hashes, uint32_t cap, uint32_t* n_entries,
                      uint32_t* insert_depth, int* has_root, uint8_t root[32]);

/* ---- the stored tree: insert and merge on a persisted frontier -----------------------
 * The pallet persists `PollStateTree {depth, count, hashes, root}` after every
 * extrinsic (pallet/src/lib.rs:706-714) and the next insert / merge starts from it.
 * These two calls are that pair on the stored state, so a node never needs the
 * leaf history:
 *
 * inf_tree_append = `insert(leaf)` for every leaf of a batch (state.rs:176-225)
 * applied to a tree whose frontier is (in_levels[i], in_hashes[i]), i < n_in, and
 * whose `depth` field is depth_in.  A fresh interaction tree is the empty frontier;
 * a fresh registration tree is the one-entry frontier (0, zeroes[0]) that
 * PollStateTree::new seeds (state.rs:150-158, 48-52).  Outputs as inf_tree_frontier:
 * the new frontier (highest level first), the new `depth`, and — if the batch
 * completed the tree (state.rs:218-222) — *has_root = 1 with the root and an empty
 * frontier.  `count` is the caller's: count_in + n_leaves.
 * Errors: TREE_ALREADY_FULL if the leaves do not fit arity^full_depth (nothing is
 * hashed; the reference would fail at the first leaf that does not fit — insert
 * the fitting prefix first to mirror that); BAD_FRONTIER if the input is not a
 * frontier insert() can leave (levels non-increasing, fewer than `arity` entries per
 * level, all below full_depth); BUFFER_TOO_SMALL if cap is short (4*33 always suffices).
 * The `_dev` form takes the leaves from device memory (e.g. straight from
 * inf_interaction_leaves_dev) and orders its work after `stream`'s.
 *
 * inf_tree_merge_frontier = `merge(to_depth)` (state.rs:230-281) on that frontier:
 * trailing runs of equal level are padded with the level's zero and hashed upwards
 * until one entry is left (and, with to_depth, until it sits at full_depth).
 * *has_root = 0 and INF_OK for an empty frontier (root stays None, state.rs:240-248).
 * The work is a chain of at most ~full_depth dependent hashes — latency, not
 * throughput: the bulk of a tree is hashed by inf_tree_append / inf_tree_merge. */
int inf_tree_append(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels,
                    const uint8_t* in_hashes, uint32_t n_in, uint32_t depth_in, const uint8_t* leaves,
                    uint64_t n_leaves, uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap,
                    uint32_t* n_entries, uint32_t* depth_out, int* has_root, uint8_t root[32]);
int inf_tree_append_dev(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels,
                        const uint8_t* in_hashes, uint32_t n_in, uint32_t depth_in, const void* d_leaves,
                        uint64_t n_leaves, uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap,
                        uint32_t* n_entries, uint32_t* depth_out, int* has_root, uint8_t root[32],
                        void* stream);
int inf_tree_merge_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* levels,
                            const uint8_t* hashes, uint32_t n, int to_depth, uint8_t root[32], int* has_root,
                            uint32_t* root_depth);

/* ---- retained trees and Merkle paths ------------------------------------------------
 * inf_tree_build keeps every level of the dense, zero-padded tree of `depth`
 * levels over the leaves on the device (the batch equivalent of the tree the
 * reference only keeps a frontier of), so sibling paths can be served in bulk —
 * what the off-chain coordinator feeds the circuits
 * (circuits/process-messages.circom:57,85).  Path layout = the argument of
 * compute_merkle_root_from_path (provider.rs:396-436): for every level the
 * arity-1 siblings in order, the node's own position skipped;
 * n_idx * depth * (arity-1) * 32 bytes.  Leaf indices count the blank leaf (it
 * is leaf 0 when prepend_blank_leaf is set). */
typedef struct inf_tree inf_tree;
int inf_tree_build(inf_ctx* ctx, uint32_t arity, uint32_t depth, int prepend_blank_leaf,
                   const uint8_t* leaves, uint64_t n_leaves, inf_tree** out);
int inf_tree_root(inf_tree* tree, uint8_t root[32]);
int inf_tree_paths(inf_tree* tree, const uint64_t* leaf_indices, uint64_t n_idx, uint8_t* paths);
/* The same for nodes of any level: paths from node_indices[i] of `level` up to the
 * root, (depth - level) * (arity-1) * 32 bytes each, and the nodes themselves
 * (all-zero subtrees to the right of the stored nodes read as zeroes[level]).
 * With level = process_subtree_depth these are the batch subroots and the
 * msgSubrootPathElements the coordinator feeds processMessages, one batch per index
 * (circuits/process-messages.circom:57,85; cli/src/utils.ts:104-126), all batches
 * in one call. */
int inf_tree_node_paths(inf_tree* tree, uint32_t level, const uint64_t* node_indices, uint64_t n_idx,
                        uint8_t* paths);
int inf_tree_level_nodes(inf_tree* tree, uint32_t level, uint64_t first, uint64_t count, uint8_t* out);
void inf_tree_destroy(inf_tree* tree);

/* compute_merkle_root_from_path (provider.rs:396-436), batched: n paths of
 * `depth` levels, arity 5 as in the reference (VOTE_TREE_ARITY) or 2.
 *   indices  n leaf indices; leaves n*32; paths n*depth*(arity-1)*32; roots n*32 out */
int inf_merkle_roots_from_paths(inf_ctx* ctx, uint32_t arity, uint32_t depth, const uint64_t* indices,
                                const uint8_t* leaves, const uint8_t* paths, uint64_t n,
                                uint8_t* roots);

/* ---- several GPUs from one process ---------------------------------------------------
 * For hosts that are a single process (the Rust shim): one context per device,
 * leaves sharded into contiguous subtrees, ONE ncclAllGather of the subtree
 * roots (libnccl.so.2 is loaded with dlopen on first use; INF_ERR_NCCL if it is
 * missing), top levels finished on the first device.  Same results and error
 * codes as inf_tree_merge.  INF_MULTI_PEER_COPY replaces the collective by
 * cudaMemcpyPeerAsync (also allows the same device to be listed twice, which
 * NCCL refuses — used to test the sharding on one GPU). */
typedef struct inf_multi inf_multi;
int inf_multi_init(const int* devices, int n_devices, uint32_t flags, inf_multi** out);
void inf_multi_destroy(inf_multi* m);
int inf_multi_device_count(const inf_multi* m);
int inf_multi_tree_merge(inf_multi* m, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                         int to_depth, const uint8_t* leaves, uint64_t n_leaves, uint8_t root[32],
                         uint32_t* insert_depth, uint32_t* root_depth, int* has_root);
/* n independent hashes, contiguous slices hashed concurrently on all devices. */
int inf_multi_poseidon_hash_batch(inf_multi* m, uint32_t n_inputs, uint32_t flags,
                                  const uint8_t* domain_tag, const uint8_t* in, uint64_t n,
                                  uint8_t* out);

/* ---- replay: raw registrations / messages -> leaves -> merged tree on the device ------
 * The reference's order of work for a poll is leaf hash -> insert -> merge
 * (provider.rs:218-287, then 289-327).  These calls run that whole chain for a batch
 * without the leaves leaving the device: the raw rows are uploaded in chunks of whole
 * kernel waves, each chunk is leaf-hashed while the next uploads, then the tree runs
 * over the resident leaves (level 0 in one launch: hashing it chunk by chunk behind the
 * leaf kernels was measured 3-5 % slower end to end).
 *   inf_replay_registrations = register_participant x n + merge_registrations:
 *       public_keys n*64, timestamps n*u64 -> registrations root (blank leaf first,
 *       merge(false)), process commitment, `depth`
 *   inf_replay_interactions  = consume_interaction x n + merge_interactions:
 *       public_keys n*64, data n*320 -> interactions root (merge(true)), `depth`,
 *       expected proof counts
 *   leaves_out  optional host array (n*32) that receives the leaves
 *   retained    optional: keeps every level on the device as an inf_tree of root-depth
 *               levels for inf_tree_paths / inf_tree_node_paths; free with
 *               inf_tree_destroy
 * Errors as inf_tree_merge (TREE_ALREADY_FULL before any work if the rows do not
 * fit; TREE_ALREADY_MERGED with the root written when they fill the tree exactly). */
int inf_replay_registrations(inf_ctx* ctx, uint32_t registration_depth, const uint8_t* public_keys,
                             const uint64_t* timestamps, uint64_t n, uint8_t root[32],
                             uint8_t process_commitment[32], uint32_t* insert_depth, uint8_t* leaves_out,
                             inf_tree** retained);
int inf_replay_interactions(inf_ctx* ctx, uint32_t interaction_depth, const uint8_t* public_keys,
                            const uint8_t* data, uint64_t n, uint32_t registrations_count,
                            uint32_t process_subtree_depth, uint32_t tally_subtree_depth, uint8_t root[32],
                            int* has_root, uint32_t* insert_depth, uint32_t* expected_process,
                            uint32_t* expected_tally, uint8_t* leaves_out, inf_tree** retained);

/* merge_registrations (provider.rs:289-311): inf_tree_merge(2, depth, 1, 0, ..)
 * followed by the process commitment H3(root, EMPTY_BALLOT_ROOTS[1], 0). */
int inf_merge_registrations(inf_ctx* ctx, uint32_t registration_depth, const uint8_t* leaves,
                            uint64_t n_leaves, uint8_t root[32], uint8_t process_commitment[32],
                            uint32_t* insert_depth);
/* merge_interactions (provider.rs:313-327): inf_tree_merge(5, depth, 0, 1, ..)
 * plus the expected proof counts. */
int inf_merge_interactions(inf_ctx* ctx, uint32_t interaction_depth, const uint8_t* leaves,
                           uint64_t n_leaves, uint32_t registrations_count,
                           uint32_t process_subtree_depth, uint32_t tally_subtree_depth,
                           uint8_t root[32], int* has_root, uint32_t* expected_process,
                           uint32_t* expected_tally);

/* ---- diagnostics ---------------------------------------------------------------- */
/* The round constants and MDS matrix for width t as canonical little-endian
 * 32-bit limbs, ark ((8+RP)*t elements) then mds (t*t): exactly the numbers in
 * parameters.rs for that width.  Returns the element count, or -1. */
int inf_debug_dense_params(uint32_t t, uint32_t* out, size_t out_words);
/* The optimised device table for width t (2..8).  Returns word count or -1. */
int inf_debug_opt_table(uint32_t t, uint32_t* out, size_t out_words);
/* Sustained 32-bit IMAD issue rate of the device, measured with independent
 * multiply-add chains on every SM sub-partition.  kind 0: IMAD (lo), kind 1:
 * IMAD.WIDE.U32 counted as 2 IMAD-equivalents each, kind 2: the carry-linked
 * chains the field product uses (same unit), kind 3: IMAD.HI.  Result in
 * IMAD/s.  kinds 4-6 probe the FP64 pipe for a double-precision formulation
 * of the product: DFMA alone, DFMA + IMAD.WIDE 1:1 and 2:1; result in
 * instructions/s. */
int inf_measure_imad_peak(inf_ctx* ctx, int kind, double* imad_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* INFIMUM_B200_H */
