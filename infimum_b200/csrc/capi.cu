// C ABI of the library (include/infimum_b200.h): argument checking, the
// reference's error semantics, device memory staging and the level-by-level
// orchestration of the tree kernels.  No arithmetic happens on the host here —
// hashes are only ever computed by the kernels.
#include "infimum_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_params.h"
#include "launch.h"
#include "poseidon.cuh"

namespace inf {
cudaError_t launch_imad_peak(int kind, int sm_count, double* imad_per_s, double* clock_mhz,
                             cudaStream_t st);
}

using namespace inf;

struct inf_tree {
    inf_ctx* ctx = nullptr;
    uint32_t arity = 0, depth = 0;
    uint64_t shift = 0, n_leaves = 0;
    void* d_nodes = nullptr;                   // all levels 0..depth back to back (level 0 = logical leaves)
    std::vector<uint64_t> offsets, counts;     // per level, in nodes
    void* d_level_ptrs = nullptr;              // device copies for the gather kernel
    void* d_level_counts = nullptr;
    void* d_zero_nodes = nullptr;
};

struct inf_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};   // copy/compute overlap for host-buffer calls
    cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
    void* scratch[2] = {nullptr, nullptr};
    size_t scratch_bytes[2] = {0, 0};
    void* io[2] = {nullptr, nullptr};          // staging for host-buffer calls
    size_t io_bytes[2] = {0, 0};
    void* d_front = nullptr;                   // frontier entries on the device: 33 levels x 4 nodes (+ 1 scratch node)
    void* bounce = nullptr;                    // pinned staging for callers with pageable buffers
    size_t bounce_bytes = 0;
    cudaEvent_t bounce_done[3][2] = {};        // per pipeline stream, per half of its staging
    uint32_t* d_dense[14] = {};
    uint8_t zeroes[2][33][32];                 // [0] binary, [1] quinary (zeroes.rs)
    std::string last_cuda_error;
};

namespace {

int cuda_fail(inf_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) {
        ctx->last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
        // Do not return to the caller with copies into or out of its buffers still
        // in flight on the context's streams (the chunked host paths queue several).
        for (int i = 0; i < 3; i++)
            if (ctx->pipe[i]) cudaStreamSynchronize(ctx->pipe[i]);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
    return e == cudaErrorMemoryAllocation ? INF_ERR_OUT_OF_MEMORY : INF_ERR_CUDA;
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

int grow(inf_ctx* ctx, void** p, size_t* have, size_t need) {
    if (*have >= need) return INF_OK;
    if (*p) {
        CU(cudaFree(*p));
        *p = nullptr;
        *have = 0;
    }
    // round up so that repeated slightly-larger calls do not reallocate
    size_t cap = std::max<size_t>(need, 1 << 20);
    cap = (cap + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    CU(cudaMalloc(p, cap));
    *have = cap;
    return INF_OK;
}

struct Bind {
    int dev_prev = -1;
    bool ok = false;
    explicit Bind(inf_ctx* ctx) {
        if (cudaGetDevice(&dev_prev) != cudaSuccess) dev_prev = -1;
        ok = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~Bind() {
        if (dev_prev >= 0) cudaSetDevice(dev_prev);
    }
};

typedef cudaError_t (*upload_fn)(const uint32_t*, size_t);
typedef cudaError_t (*hash_fn)(const void*, void*, uint64_t, const TagArg&, bool, cudaStream_t);

upload_fn uploaders[9] = {nullptr, nullptr, upload_table_t2, upload_table_t3, upload_table_t4,
                          upload_table_t5, upload_table_t6, upload_table_t7, upload_table_t8};
hash_fn hashers[9] = {nullptr, nullptr, launch_hash_batch_t2, launch_hash_batch_t3,
                      launch_hash_batch_t4, launch_hash_batch_t5, launch_hash_batch_t6,
                      launch_hash_batch_t7, launch_hash_batch_t8};

TagArg make_tag(const uint8_t* tag) {
    TagArg t;
    memset(&t, 0, sizeof t);
    if (tag) {
        memcpy(t.w, tag, 32);
        t.has = 1;
    }
    return t;
}

cudaError_t launch_level(uint32_t arity, const void* in, uint64_t shift, uint64_t n_in, void* out,
                         uint64_t n_out, const uint8_t* zero, cudaStream_t st, const void* prefix = nullptr) {
    return arity == 2 ? launch_tree_level_t3(in, prefix, shift, n_in, out, n_out, zero, st)
                      : launch_tree_level_t6(in, prefix, shift, n_in, out, n_out, zero, st);
}

// arity^e saturating at 2^64-1
uint64_t pow_sat(uint64_t a, uint32_t e) {
    unsigned __int128 r = 1;
    for (uint32_t i = 0; i < e; i++) {
        r *= a;
        if (r > (unsigned __int128)UINT64_MAX) return UINT64_MAX;
    }
    return (uint64_t)r;
}

// Caller-supplied PoseidonParameters (Poseidon::new, poseidon.rs:47-71, 105-108), table on the device.
struct CustomParams {
    const uint32_t* d_tbl;
    int full_rounds, partial_rounds;
    uint64_t alpha;
};

int hash_batch_dev(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags, const uint8_t* tag,
                   const void* d_in, uint64_t n, void* d_out, cudaStream_t st, bool dense,
                   const CustomParams* custom = nullptr) {
    const uint32_t t = n_inputs + 1;
    const bool le = flags & INF_FLAG_LITTLE_ENDIAN;
    const TagArg ta = make_tag(tag);
    if (custom) {
        CU(launch_hash_dense_params((int)t, custom->full_rounds, custom->partial_rounds, custom->alpha, custom->d_tbl,
                                    d_in, d_out, n, ta, le, st));
    } else if (!dense && t <= 8) {
        CU(hashers[t](d_in, d_out, n, ta, le, st));
    } else {
        CU(launch_hash_dense((int)t, ctx->d_dense[t], d_in, d_out, n, ta, le, st));
    }
    return INF_OK;
}

int check_hash_args(inf_ctx* ctx, uint32_t n_inputs, const void* in, uint64_t n, void* out) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    // new_circom: width = n_inputs + 1 must be in 2..13 (poseidon.rs:315-320,
    // parameters.rs:38-42)
    if (n_inputs < 1 || n_inputs + 1 > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    if (n && (!in || !out)) return INF_ERR_NULL_POINTER;
    return INF_OK;
}

// Where level 0's input comes from when it is not on the device yet: fill(lo, hi, dst, ps,
// slot) makes leaves [lo, hi) appear at dst (device memory) in stream order on ps, one of the
// three pipeline streams (slot = its index, for per-stream staging).  Level 0 is then cut into
// chunks of `chunk_out` parents that rotate over the pipeline streams, each producing its slice
// of leaves and hashing it, so that uploads (and, for a replay from raw messages, the leaf
// hashing) of one chunk hide behind the hashing of another.
struct LeafFeed {
    std::function<int(uint64_t lo, uint64_t hi, char* dst, cudaStream_t ps, int slot)> fill;
    uint64_t chunk_out = 1ull << 19;         // parents per chunk (level0_per_chunk) or leaves per chunk (otherwise)
    // Hash level 0 chunk by chunk behind each chunk's fill (uploads of host leaves hide behind it), or
    // only fill per chunk and hash level 0 in one launch afterwards (a feed that hashes leaves from raw
    // rows keeps the GPU busy by itself, and level 0 in one piece runs at the full rate: the chunked
    // form cost a replay 3 % — small level kernels, and two large kernels evicting each other's
    // instructions; tools/replay_probe.py).
    bool level0_per_chunk = true;
    uint8_t* leaves_out = nullptr;           // optional host copy of all leaves (n_leaves * 32 bytes)
};

int alloc_tree(inf_ctx* ctx, uint32_t arity, uint32_t depth, uint64_t shift, uint64_t n_leaves, inf_tree** out);

// Leaves in host memory: one asynchronous upload per chunk.
// Chunks of the host pipelines are whole waves of the wide launch shapes (one 384- or 512-thread
// block per SM, launch.h): sms x 1536 threads = 4 waves of 384 = 3 waves of 512.
inline uint64_t wave_unit(const inf_ctx* ctx) { return (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 1536; }

LeafFeed host_leaf_feed(inf_ctx* ctx, const uint8_t* h_leaves) {
    LeafFeed f;
    f.chunk_out = 2 * wave_unit(ctx);
    f.fill = [ctx, h_leaves](uint64_t lo, uint64_t hi, char* dst, cudaStream_t ps, int) -> int {
        CU(cudaMemcpyAsync(dst, h_leaves + lo * 32, (hi - lo) * 32, cudaMemcpyHostToDevice, ps));
        return INF_OK;
    };
    return f;
}

// Levels level_in .. level_in + n_levels - 1 over a run of nodes (`shift` leading zero nodes, then
// n_in nodes at d_in): output of local level l goes to dst_of(l).  With a feed, level 0's input is
// produced chunk by chunk on the pipeline streams (see LeafFeed) and `st` is made to wait for them.
// Everything is enqueued; nothing is synchronised.  n_levels >= 1.
int reduce_levels(inf_ctx* ctx, uint32_t arity, uint32_t level_in, uint32_t n_levels, uint64_t shift,
                  const void* d_in, uint64_t n_in, const LeafFeed* feed,
                  const std::function<char*(uint32_t)>& dst_of, cudaStream_t st, const void** out_ptr,
                  uint64_t* out_n) {
    const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
    const uint64_t n_total = n_in + shift;
    const uint64_t n1 = (n_total + arity - 1) / arity;
    const void* cur = d_in;
    uint64_t n_cur = n_in, sh = shift;
    uint32_t l_first = 0;
    int rc;
    if (feed) {
        // The first chunk is a quarter and the second half the size: nothing can hide the first
        // upload, so it is kept short.
        static const bool ramp = !(getenv("INF_NO_RAMP") && atoi(getenv("INF_NO_RAMP")));
        int k = 0;
        auto step_of = [&](int i) { return (ramp && i < 2 && feed->chunk_out >= 64) ? feed->chunk_out >> (2 - i) : feed->chunk_out; };
        if (feed->level0_per_chunk) {
            for (uint64_t o0 = 0, step = 0; o0 < n1; o0 += step, k++) {
                step = step_of(k);
                const uint64_t o1 = std::min<uint64_t>(o0 + step, n1);
                const uint64_t L0 = o0 * arity, L1 = std::min<uint64_t>(o1 * arity, n_total);
                const uint64_t leaf_lo = L0 >= shift ? L0 - shift : 0, leaf_hi = L1 - shift;
                cudaStream_t ps = ctx->pipe[k % 3];
                char* dl = (char*)d_in + leaf_lo * 32;
                if (leaf_hi > leaf_lo && (rc = feed->fill(leaf_lo, leaf_hi, dl, ps, k % 3))) return rc;
                CU(launch_level(arity, dl, o0 == 0 ? shift : 0, leaf_hi - leaf_lo, dst_of(0) + o0 * 32, o1 - o0,
                                Z[level_in], ps));
            }
        } else {
            for (uint64_t lo = 0; lo < n_in; k++) {
                const uint64_t hi = std::min<uint64_t>(lo + step_of(k), n_in);
                if ((rc = feed->fill(lo, hi, (char*)d_in + lo * 32, ctx->pipe[k % 3], k % 3))) return rc;
                lo = hi;
            }
        }
        for (int i = 0; i < 3; i++) {
            CU(cudaEventRecord(ctx->pipe_done[i], ctx->pipe[i]));
            CU(cudaStreamWaitEvent(st, ctx->pipe_done[i], 0));
        }
        if (feed->leaves_out && n_in) {
            // all leaves exist once every pipeline stream is through: read them back on pipe[0]
            // while `st` hashes the upper levels (the caller synchronises pipe[0])
            for (int i = 1; i < 3; i++) CU(cudaStreamWaitEvent(ctx->pipe[0], ctx->pipe_done[i], 0));
            CU(cudaMemcpyAsync(feed->leaves_out, d_in, (size_t)n_in * 32, cudaMemcpyDeviceToHost, ctx->pipe[0]));
        }
        if (feed->level0_per_chunk) {
            cur = dst_of(0);
            n_cur = n1;
            sh = 0;
            l_first = 1;
        }
    }
    for (uint32_t l = l_first; l < n_levels; l++) {
        const uint64_t n_next = (n_cur + sh + arity - 1) / arity;
        void* dst = dst_of(l);
        CU(launch_level(arity, cur, sh, n_cur, dst, n_next, Z[level_in + l], st));
        cur = dst;
        n_cur = n_next;
        sh = 0;
    }
    if (out_ptr) *out_ptr = cur;
    if (out_n) *out_n = n_cur;
    return INF_OK;
}

// Core of the tree merge.  Leaves the root (if any) in host memory.  `st` is synchronised
// before return.  Without a feed the leaves are already at d_leaves; with one, d_leaves is the
// (uninitialised) device area they are produced into.  With `keep`, every level is retained in
// a new inf_tree of root_depth levels (for Merkle paths) instead of ping-pong scratch; d_leaves
// is then ignored unless there is no feed, in which case the leaves are copied from it.
int tree_merge_dev(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int blank, int to_depth,
                   const void* d_leaves, uint64_t n_leaves, uint8_t* root, uint32_t* insert_depth,
                   uint32_t* root_depth, int* has_root, cudaStream_t st, const LeafFeed* feed = nullptr,
                   inf_tree** keep = nullptr) {
    if (keep) *keep = nullptr;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    if (n_leaves && !d_leaves && !(feed && keep)) return INF_ERR_NULL_POINTER;
    const uint64_t shift = blank ? 1 : 0;
    const uint64_t n_total = n_leaves + shift;
    const uint64_t cap = pow_sat(arity, full_depth);
    if (insert_depth) *insert_depth = 0;
    if (root_depth) *root_depth = 0;
    if (has_root) *has_root = 0;
    if (n_total > cap) return INF_ERR_TREE_ALREADY_FULL;          // insert(): state.rs:182
    if (n_total == 0) return INF_OK;                              // merge() on an empty frontier
    // depth reached by insert(): largest d with arity^d <= n_total (state.rs:212-213)
    uint32_t idepth = 0;
    while (idepth < full_depth && pow_sat(arity, idepth + 1) <= n_total) idepth++;
    // levels under the root
    uint32_t rdepth;
    const bool completed_by_insert = (n_total == cap);           // state.rs:218-222
    if (to_depth || completed_by_insert) {
        rdepth = full_depth;
    } else {
        rdepth = 0;
        while (pow_sat(arity, rdepth) < n_total) rdepth++;
    }
    if (insert_depth) *insert_depth = idepth;
    if (root_depth) *root_depth = rdepth;

    const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
    int rc;
    inf_tree* kt = nullptr;
    LeafFeed copy_feed;
    cudaError_t e0 = cudaSuccess;
    if (keep) {
        if ((rc = alloc_tree(ctx, arity, rdepth, shift, n_leaves, &kt))) return rc;
        if (!feed) {                                               // device leaves: copy them into level 0
            const char* src = (const char*)d_leaves;
            copy_feed.fill = [ctx, src](uint64_t lo, uint64_t hi, char* dst, cudaStream_t ps, int) -> int {
                CU(cudaMemcpyAsync(dst, src + lo * 32, (hi - lo) * 32, cudaMemcpyDeviceToDevice, ps));
                return INF_OK;
            };
            feed = &copy_feed;
        }
        d_leaves = (char*)kt->d_nodes + shift * 32;
        if (shift && (e0 = cudaMemcpyAsync(kt->d_nodes, Z[0], 32, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
            inf_tree_destroy(kt);
            return cuda_fail(ctx, e0, "blank leaf");
        }
    }
    auto bail = [&](int code) {
        if (kt) inf_tree_destroy(kt);
        return code;
    };
    // output of level l (the nodes of level l + 1)
    const std::function<char*(uint32_t)> level_dst = [&](uint32_t l) -> char* {
        return kt ? (char*)kt->d_nodes + kt->offsets[l + 1] * 32 : (char*)ctx->scratch[l & 1];
    };
    uint8_t root_local[32];
    cudaError_t e = cudaSuccess;
#define CUB(call)                                                         \
    do {                                                                  \
        if ((e = (call)) != cudaSuccess) return bail(cuda_fail(ctx, e, #call)); \
    } while (0)
    if (rdepth == 0) {
        // single node: the blank leaf itself, or the only leaf
        if (!blank && feed && (rc = feed->fill(0, 1, (char*)d_leaves, st, 0))) return bail(rc);
        if (blank) memcpy(root_local, Z[0], 32);
        else CUB(cudaMemcpyAsync(root_local, d_leaves, 32, cudaMemcpyDeviceToHost, st));
        if (feed && feed->leaves_out && n_leaves) CUB(cudaMemcpyAsync(feed->leaves_out, d_leaves, 32, cudaMemcpyDeviceToHost, st));
        CUB(cudaStreamSynchronize(st));
    } else {
        const uint64_t n1 = (n_total + arity - 1) / arity;
        const uint64_t n2 = (n1 + arity - 1) / arity;
        if (!kt) {
            if ((rc = grow(ctx, &ctx->scratch[0], &ctx->scratch_bytes[0], n1 * 32))) return rc;
            if ((rc = grow(ctx, &ctx->scratch[1], &ctx->scratch_bytes[1], n2 * 32))) return rc;
        }
        const void* cur = nullptr;
        if ((rc = reduce_levels(ctx, arity, 0, rdepth, shift, d_leaves, n_leaves, feed, level_dst, st, &cur, nullptr)))
            return bail(rc);
        CUB(cudaMemcpyAsync(root_local, cur, 32, cudaMemcpyDeviceToHost, st));
        CUB(cudaStreamSynchronize(st));
        if (feed && feed->leaves_out) CUB(cudaStreamSynchronize(ctx->pipe[0]));
    }
#undef CUB
    if (root) memcpy(root, root_local, 32);
    if (has_root) *has_root = 1;
    if (keep) *keep = kt;
    return completed_by_insert ? INF_ERR_TREE_ALREADY_MERGED : INF_OK;   // merge(): state.rs:236
}

// Generic chunked pipeline for per-row kernels with up to two host inputs:
// rows rotate over the three pipeline streams as H2D -> kernel -> D2H.
template <class Launch>
int rows_pipeline(inf_ctx* ctx, const uint8_t* in0, size_t row0, const uint8_t* in1, size_t row1,
                         uint64_t n, uint8_t* out, Launch launch) {
    const uint64_t super = 1ull << 22, chunk = wave_unit(ctx) / 2;      // two waves of the 384-thread leaf blocks
    const uint64_t n_stage = std::min<uint64_t>(n, super);
    int rc;
    // staging: io[0] holds both inputs back to back, io[1] the output
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], n_stage * (row0 + row1)))) return rc;
    if ((rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], n_stage * 32))) return rc;
    char* s0 = (char*)ctx->io[0];
    char* s1 = s0 + n_stage * row0;
    // The read-back of chunk k is queued two chunks late: with pageable caller
    // buffers a device-to-host copy blocks the calling thread until the chunk's
    // kernel is done, which would serialise the whole loop; this way the thread
    // is staging the upload of chunk k+2 while chunk k is being hashed.
    const uint64_t lag = 2;
    for (uint64_t base = 0; base < n; base += super) {
        const uint64_t m = std::min<uint64_t>(super, n - base);
        const uint64_t n_chunks = (m + chunk - 1) / chunk;
        for (uint64_t k = 0; k < n_chunks + lag; k++) {
            if (k < n_chunks) {
                const uint64_t off = k * chunk, c = std::min<uint64_t>(chunk, m - off);
                cudaStream_t st = ctx->pipe[k % 3];
                CU(cudaMemcpyAsync(s0 + off * row0, in0 + (base + off) * row0, c * row0, cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(s1 + off * row1, in1 + (base + off) * row1, c * row1, cudaMemcpyHostToDevice, st));
                CU(launch(s0 + off * row0, s1 + off * row1, (char*)ctx->io[1] + off * 32, c, st));
            }
            if (k >= lag) {
                const uint64_t off = (k - lag) * chunk, c = std::min<uint64_t>(chunk, m - off);
                CU(cudaMemcpyAsync(out + (base + off) * 32, (char*)ctx->io[1] + off * 32, c * 32, cudaMemcpyDeviceToHost,
                                   ctx->pipe[(k - lag) % 3]));
            }
        }
        for (int i = 0; i < 3; i++) CU(cudaStreamSynchronize(ctx->pipe[i]));
    }
    return INF_OK;
}


// ---- stored frontiers (`PollStateTree.hashes`, state.rs:85-86) ---------------------------------
constexpr uint32_t FRONT_NODES = 33 * 4;      // at most arity - 1 <= 4 entries per level, levels 0..32

inline bool misaligned(const void* p) { return ((uintptr_t)p & 15) != 0; }

int ensure_front(inf_ctx* ctx) {
    if (!ctx->d_front) CU(cudaMalloc(&ctx->d_front, (FRONT_NODES + 1) * 32));
    return INF_OK;
}

// Entries per level (levels must not increase towards the tail, fewer than `arity` per level,
// all below full_depth) and the number of leaves the frontier stands for.
int parse_frontier(uint32_t arity, uint32_t full_depth, const uint8_t* levels, uint32_t n, uint32_t cnt[33],
                   uint64_t* n_leaves) {
    for (int l = 0; l < 33; l++) cnt[l] = 0;
    unsigned __int128 total = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t l = levels[i];
        if (l > 32 || l >= full_depth + (n == 1 ? 1u : 0u) || (i && l > levels[i - 1])) return INF_ERR_BAD_FRONTIER;
        if (++cnt[l] >= arity) return INF_ERR_BAD_FRONTIER;
        total += pow_sat(arity, l);
    }
    if (total > (unsigned __int128)UINT64_MAX) return INF_ERR_BAD_FRONTIER;
    *n_leaves = (uint64_t)total;
    return INF_OK;
}

// The insert cascade of state.rs:176-225 for a whole batch: frontier in + leaves -> frontier out.
// Appending leaves is an addition in base `arity`: at every level the pending nodes are the
// frontier's entries at that level (they open the first, incomplete group) followed by the
// nodes the level below produced; every full group of `arity` becomes a parent, the remainder
// is the new frontier at that level.  Only full groups are hashed (no zero padding): that is
// merge's business.  Leaves come from host memory (h_leaves) or device memory (d_leaves).
int tree_append_core(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels,
                     const uint8_t* in_hashes, uint32_t n_in, uint32_t depth_in, const uint8_t* h_leaves,
                     const void* d_leaves, uint64_t n_leaves, uint8_t* out_levels, uint8_t* out_hashes,
                     uint32_t cap, uint32_t* n_entries, uint32_t* depth_out, int* has_root, uint8_t root[32],
                     cudaStream_t st) {
    *n_entries = 0;
    if (depth_out) *depth_out = depth_in;
    if (has_root) *has_root = 0;
    uint32_t cnt[33];
    uint64_t n_old = 0;
    int rc = parse_frontier(arity, full_depth, in_levels, n_in, cnt, &n_old);
    if (rc) return rc;
    const uint64_t capacity = pow_sat(arity, full_depth);
    if (n_old > capacity || n_leaves > capacity - n_old) return INF_ERR_TREE_ALREADY_FULL;   // insert(): state.rs:182
    // frontier entries by level, in order; `first[l]` = index of level l's first entry
    uint32_t first[34];
    {
        uint32_t at = n_in;
        for (int l = 0; l <= 33; l++) {                 // levels are stored highest first
            first[l] = l < 33 ? at - cnt[l] : 0;
            if (l < 33) at -= cnt[l];
        }
    }
    std::vector<std::vector<uint8_t>> tails(33);        // new frontier per level
    int stop = -1;                                       // levels above `stop` keep their input entries
    uint32_t depth = depth_in;
    bool completed = false;
    uint8_t root_local[32];
    if (n_leaves) {
        if ((rc = ensure_front(ctx))) return rc;
        const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
        // device copy of the input frontier, level l at node 4 l
        std::vector<uint8_t> packed(FRONT_NODES * 32, 0);
        for (int l = 0; l < 33; l++)
            for (uint32_t k = 0; k < cnt[l]; k++) memcpy(&packed[(4 * l + k) * 32], in_hashes + (size_t)(first[l] + k) * 32, 32);
        if (n_in) CU(cudaMemcpyAsync(ctx->d_front, packed.data(), packed.size(), cudaMemcpyHostToDevice, st));
        const char* cur;
        if (h_leaves) {
            if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], (size_t)n_leaves * 32))) return rc;
            CU(cudaMemcpyAsync(ctx->io[0], h_leaves, (size_t)n_leaves * 32, cudaMemcpyHostToDevice, st));
            cur = (const char*)ctx->io[0];
        } else {
            cur = (const char*)d_leaves;
        }
        const uint64_t n1 = (n_leaves + 4) / arity + 1, n2 = n1 / arity + 2;
        if ((rc = grow(ctx, &ctx->scratch[0], &ctx->scratch_bytes[0], n1 * 32))) return rc;
        if ((rc = grow(ctx, &ctx->scratch[1], &ctx->scratch_bytes[1], n2 * 32))) return rc;
        uint64_t n_cur = n_leaves;
        for (uint32_t l = 0;; l++) {
            if (l == full_depth) {                       // the cascade reached the top: n_old + n_leaves == capacity
                CU(cudaMemcpyAsync(root_local, cur, 32, cudaMemcpyDeviceToHost, st));
                completed = true;
                stop = 32;
                break;
            }
            const uint64_t pc = cnt[l], total = pc + n_cur, n_par = total / arity, rem = total % arity;
            tails[l].resize(rem * 32);
            for (uint64_t k = 0; k < rem; k++) {
                const uint64_t idx = total - rem + k;
                if (idx < pc) memcpy(&tails[l][k * 32], in_hashes + (size_t)(first[l] + idx) * 32, 32);
                else CU(cudaMemcpyAsync(&tails[l][k * 32], cur + (idx - pc) * 32, 32, cudaMemcpyDeviceToHost, st));
            }
            stop = (int)l;
            if (n_par == 0) break;
            char* dst = (char*)ctx->scratch[l & 1];
            CU(launch_level(arity, cur, pc, n_par * arity - pc, dst, n_par, Z[l], st,
                            pc ? (const char*)ctx->d_front + (size_t)4 * l * 32 : nullptr));
            depth = std::max(depth, l + 1);              // state.rs:212-213
            cur = dst;
            n_cur = n_par;
        }
        CU(cudaStreamSynchronize(st));
    }
    if (depth_out) *depth_out = depth;
    if (completed) {                                     // state.rs:218-222: root set, frontier cleared
        if (has_root) *has_root = 1;
        if (root) memcpy(root, root_local, 32);
        return INF_OK;
    }
    uint32_t n = 0;
    for (int l = 32; l >= 0; l--) {
        const uint32_t k_n = l > stop ? cnt[l] : (uint32_t)(tails[l].size() / 32);
        for (uint32_t k = 0; k < k_n; k++) {
            if (n >= cap || !out_levels || !out_hashes) return INF_ERR_BUFFER_TOO_SMALL;
            out_levels[n] = (uint8_t)l;
            memcpy(out_hashes + 32 * n, l > stop ? in_hashes + (size_t)(first[l] + k) * 32 : &tails[l][k * 32], 32);
            n++;
        }
    }
    *n_entries = n;
    return INF_OK;
}

// rows [lo, hi) of two host arrays -> per-stream staging -> leaf kernel -> dst
template <class Launch>
LeafFeed raw_rows_feed(inf_ctx* ctx, const uint8_t* in0, size_t row0, const uint8_t* in1, size_t row1,
                       uint64_t chunk_out, size_t slot_rows, Launch launch) {
    LeafFeed f;
    f.chunk_out = chunk_out;
    f.level0_per_chunk = false;
    f.fill = [=](uint64_t lo, uint64_t hi, char* dst, cudaStream_t ps, int slot) -> int {
        char* s0 = (char*)ctx->io[0] + (size_t)slot * slot_rows * (row0 + row1);
        char* s1 = s0 + slot_rows * row0;
        CU(cudaMemcpyAsync(s0, in0 + lo * row0, (hi - lo) * row0, cudaMemcpyHostToDevice, ps));
        CU(cudaMemcpyAsync(s1, in1 + lo * row1, (hi - lo) * row1, cudaMemcpyHostToDevice, ps));
        CU(launch(s0, s1, dst, hi - lo, ps));
        return INF_OK;
    };
    return f;
}

// A retained tree: every level 0..depth of the dense zero-padded tree, back to back on the device.
int alloc_tree(inf_ctx* ctx, uint32_t arity, uint32_t depth, uint64_t shift, uint64_t n_leaves, inf_tree** out) {
    inf_tree* t = new inf_tree();
    t->ctx = ctx; t->arity = arity; t->depth = depth; t->shift = shift; t->n_leaves = n_leaves;
    uint64_t total_nodes = 0, c = n_leaves + shift;
    for (uint32_t l = 0; l <= depth; l++) {
        t->offsets.push_back(total_nodes);
        t->counts.push_back(c);
        total_nodes += c;
        c = (c + arity - 1) / arity;
    }
    const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
    cudaError_t e;
    auto fail = [&](const char* what) {
        inf_tree_destroy(t);
        return cuda_fail(ctx, e, what);
    };
    if ((e = cudaMalloc(&t->d_nodes, total_nodes * 32)) != cudaSuccess) return fail("cudaMalloc tree levels");
    if ((e = cudaMalloc(&t->d_level_ptrs, (depth + 1) * sizeof(void*))) != cudaSuccess) return fail("cudaMalloc");
    if ((e = cudaMalloc(&t->d_level_counts, (depth + 1) * 8)) != cudaSuccess) return fail("cudaMalloc");
    if ((e = cudaMalloc(&t->d_zero_nodes, 33 * 32)) != cudaSuccess) return fail("cudaMalloc");
    std::vector<void*> ptrs;
    for (uint32_t l = 0; l <= depth; l++) ptrs.push_back((char*)t->d_nodes + t->offsets[l] * 32);
    // synchronous copies of a few hundred bytes: the host arrays above do not outlive this call
    if ((e = cudaMemcpy(t->d_level_ptrs, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy");
    if ((e = cudaMemcpy(t->d_level_counts, t->counts.data(), t->counts.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy");
    if ((e = cudaMemcpy(t->d_zero_nodes, Z, 33 * 32, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy");
    *out = t;
    return INF_OK;
}

std::mutex g_init_mu;

}  // namespace

extern "C" {

const char* inf_version(void) { return "infimum_b200 0.1 (sm_100a)"; }

const char* inf_strerror(int code) {
    switch (code) {
        case INF_OK: return "ok";
        case INF_ERR_TREE_ALREADY_FULL: return "MerkleTreeError::TreeAlreadyFull";
        case INF_ERR_TREE_ALREADY_MERGED: return "MerkleTreeError::TreeAlreadyMerged";
        case INF_ERR_HASH_FAILED: return "MerkleTreeError::HashFailed";
        case INF_ERR_MERGE_FAILED: return "MerkleTreeError::MergeFailed";
        case INF_ERR_INVALID_NUMBER_OF_INPUTS: return "PoseidonError::InvalidNumberOfInputs";
        case INF_ERR_EMPTY_INPUT: return "PoseidonError::EmptyInput";
        case INF_ERR_INVALID_INPUT_LENGTH: return "PoseidonError::InvalidInputLength";
        case INF_ERR_INVALID_WIDTH_CIRCOM: return "PoseidonError::InvalidWidthCircom";
        case INF_ERR_NULL_POINTER: return "null pointer argument";
        case INF_ERR_BAD_ARITY: return "arity must be 2 or 5";
        case INF_ERR_BAD_DEPTH: return "tree depth exceeds the 33-level zero table";
        case INF_ERR_BUFFER_TOO_SMALL: return "output buffer too small";
        case INF_ERR_BAD_FRONTIER: return "frontier is not a state insert() can leave behind";
        case INF_ERR_BAD_ALIGNMENT: return "device pointer must be 16-byte aligned";
        case INF_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case INF_ERR_CUDA: return "CUDA error (see inf_last_cuda_error)";
        case INF_ERR_OUT_OF_MEMORY: return "device out of memory";
        case INF_ERR_NCCL: return "NCCL unavailable or failed";
        default: return "unknown error code";
    }
}

const char* inf_last_cuda_error(const inf_ctx* ctx) { return ctx ? ctx->last_cuda_error.c_str() : ""; }

int inf_init(int device, inf_ctx** out) {
    if (!out) return INF_ERR_NULL_POINTER;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return INF_ERR_NO_DEVICE;
    }
    std::lock_guard<std::mutex> lock(g_init_mu);
    inf_ctx* ctx = new inf_ctx();
    ctx->device = device;
    Bind bind(ctx);
    auto fail = [&](int rc) {
        inf_destroy(ctx);
        return rc;
    };
    if (!bind.ok) return fail(INF_ERR_NO_DEVICE);
    cudaError_t e;
    if ((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return fail(cuda_fail(nullptr, e, "cudaDeviceGetAttribute"));
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(cuda_fail(nullptr, e, "cudaStreamCreate"));
    for (int i = 0; i < 3; i++) {
        if ((e = cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking)) != cudaSuccess)
            return fail(cuda_fail(nullptr, e, "cudaStreamCreate"));
        if ((e = cudaEventCreateWithFlags(&ctx->pipe_done[i], cudaEventDisableTiming)) != cudaSuccess)
            return fail(cuda_fail(nullptr, e, "cudaEventCreate"));
    }
    try {
        for (int t = 2; t <= 8; t++) {
            std::vector<uint32_t> tbl = host::build_opt_table(t);
            if ((e = uploaders[t](tbl.data(), tbl.size())) != cudaSuccess)
                return fail(cuda_fail(nullptr, e, "upload optimised table"));
        }
        {
            std::vector<uint32_t> t5 = host::build_opt_table(5), t6 = host::build_opt_table(6);
            if ((e = upload_leaf_tables(t5.data(), t5.size(), t6.data(), t6.size())) != cudaSuccess)
                return fail(cuda_fail(nullptr, e, "upload leaf tables"));
        }
        for (int t = 2; t <= 13; t++) {
            std::vector<uint32_t> tbl = host::build_dense_table(t);
            if ((e = cudaMalloc((void**)&ctx->d_dense[t], tbl.size() * 4)) != cudaSuccess)
                return fail(cuda_fail(nullptr, e, "cudaMalloc dense table"));
            if ((e = cudaMemcpy(ctx->d_dense[t], tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
                return fail(cuda_fail(nullptr, e, "cudaMemcpy dense table"));
        }
    } catch (const std::exception&) {
        return fail(INF_ERR_HASH_FAILED);
    }
    // Zero tables: Z[l+1] = H(Z[l] x arity), 32 links each, hashed on the device: one launch per
    // arity (a loop over the links inside one block of the cooperative kernel), the two chains
    // side by side on two streams.
    {
        void* d = nullptr;
        if ((e = cudaMalloc(&d, 2 * 33 * 32)) != cudaSuccess) return fail(cuda_fail(nullptr, e, "cudaMalloc"));
        cudaStream_t sts[2] = {ctx->stream, ctx->pipe[0]};
        for (int a = 0; a < 2 && e == cudaSuccess; a++) {
            char* da = (char*)d + a * 33 * 32;
            memcpy(ctx->zeroes[a][0], a == 0 ? host::BINARY_ZERO_LEAF_BE : host::QUINARY_ZERO_LEAF_BE, 32);
            e = cudaMemcpyAsync(da, ctx->zeroes[a][0], 32, cudaMemcpyHostToDevice, sts[a]);
            if (e == cudaSuccess) e = a == 0 ? launch_hash_chain_t3(da, 32, sts[a]) : launch_hash_chain_t6(da, 32, sts[a]);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->zeroes[a], da, 33 * 32, cudaMemcpyDeviceToHost, sts[a]);
        }
        for (int a = 0; a < 2; a++) {
            const cudaError_t es = cudaStreamSynchronize(sts[a]);
            if (e == cudaSuccess) e = es;
        }
        cudaFree(d);
        if (e != cudaSuccess) return fail(cuda_fail(nullptr, e, "zero-table chain"));
    }
    *out = ctx;
    return INF_OK;
}

void inf_destroy(inf_ctx* ctx) {
    if (!ctx) return;
    {
        Bind bind(ctx);
        for (int i = 0; i < 2; i++) {
            if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
            if (ctx->io[i]) cudaFree(ctx->io[i]);
        }
        if (ctx->d_front) cudaFree(ctx->d_front);
        if (ctx->bounce) cudaFreeHost(ctx->bounce);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 2; j++)
                if (ctx->bounce_done[i][j]) cudaEventDestroy(ctx->bounce_done[i][j]);
        for (int t = 0; t < 14; t++)
            if (ctx->d_dense[t]) cudaFree(ctx->d_dense[t]);
        for (int i = 0; i < 3; i++) {
            if (ctx->pipe[i]) cudaStreamDestroy(ctx->pipe[i]);
            if (ctx->pipe_done[i]) cudaEventDestroy(ctx->pipe_done[i]);
        }
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

int inf_poseidon_hash_batch_dev(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                                const uint8_t* domain_tag, const void* d_in, uint64_t n,
                                void* d_out, void* stream) {
    int rc = check_hash_args(ctx, n_inputs, d_in, n, d_out);
    if (rc) return rc;
    if (misaligned(d_in) || misaligned(d_out)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return hash_batch_dev(ctx, n_inputs, flags, domain_tag, d_in, n, d_out,
                          stream ? (cudaStream_t)stream : ctx->stream, false);
}

// Is this host pointer ordinary pageable memory (neither cudaHostAlloc'ed nor cudaHostRegister'ed)?
static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// The same batch for callers whose buffers are pageable (a Rust Vec<u8>, a numpy array).  A copy
// from or to pageable memory is staged by the calling thread and blocks it, which serialises
// upload, hashing and read-back (measured -12 %: 124.6 vs 139.9 M hash2/s in round 1).  Here each
// of the three pipeline streams gets a host thread of its own and two halves of pinned staging:
// the thread copies chunk j into pinned memory while chunk j-1 of its stream is being hashed and
// copies results out of pinned memory one chunk late, so the three threads together move the
// ~14 GB/s the GPU consumes and nothing in the loop waits for the device except on a chunk that
// is already one behind.
static int hash_batch_pageable(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags, const uint8_t* tag,
                               const uint8_t* in, uint64_t n, uint8_t* out, bool dense, const CustomParams* custom) {
    // whole waves of 384-thread blocks, at most four, and at most ~24 MB of pinned staging per slot
    // (six slots: wide inputs would otherwise pin half a gigabyte)
    const size_t in_row = (size_t)n_inputs * 32;
    const uint64_t per_wave = wave_unit(ctx) / 4;
    const uint64_t chunk = per_wave * std::max<uint64_t>(1, std::min<uint64_t>(4, (24ull << 20) / ((in_row + 32) * per_wave)));
    const size_t slot = chunk * (in_row + 32);
    int rc;
    if (ctx->bounce_bytes < 6 * slot) {
        if (ctx->bounce) CU(cudaFreeHost(ctx->bounce));
        ctx->bounce = nullptr;
        ctx->bounce_bytes = 0;
        CU(cudaHostAlloc(&ctx->bounce, 6 * slot, cudaHostAllocDefault));
        ctx->bounce_bytes = 6 * slot;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++)
            if (!ctx->bounce_done[i][j]) CU(cudaEventCreateWithFlags(&ctx->bounce_done[i][j], cudaEventDisableTiming));
    // device staging: every chunk in flight has its own region (3 streams x 2 halves)
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], 6 * chunk * in_row))) return rc;
    if ((rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], 6 * chunk * 32))) return rc;
    // Chunk boundaries: full chunks in the middle, smaller ones at both ends — the first upload
    // cannot overlap with anything and neither can the last copy-out, so both are kept short.
    std::vector<uint64_t> cut{0};
    {
        const uint64_t ramp[6] = {chunk / 8, chunk / 8, chunk / 8, chunk / 2, chunk / 2, chunk / 2};
        uint64_t tail = 0;
        for (int i = 0; i < 6; i++) tail += ramp[i];
        uint64_t pos = 0;
        for (int i = 0; i < 6 && pos + ramp[i] + tail < n; i++) cut.push_back(pos += ramp[i]);
        while (pos + chunk + tail < n) cut.push_back(pos += chunk);
        for (int i = 5; i >= 0 && pos < n; i--) cut.push_back(pos = std::min<uint64_t>(n, pos + std::max<uint64_t>(ramp[i], (n - pos) / (i + 1))));
        if (pos < n) cut.push_back(n);
    }
    const uint64_t n_chunks = cut.size() - 1;
    int rcs[3] = {INF_OK, INF_OK, INF_OK};
    cudaError_t errs[3] = {cudaSuccess, cudaSuccess, cudaSuccess};
    auto worker = [&](int i) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) {
            rcs[i] = INF_ERR_NO_DEVICE;
            return;
        }
        cudaStream_t st = ctx->pipe[i];
        cudaError_t e = cudaSuccess;
        auto region = [&](int half, char** h_in, char** h_out, char** d_in, char** d_out) {
            const int r = 2 * i + half;
            *h_in = (char*)ctx->bounce + (size_t)r * slot;
            *h_out = *h_in + chunk * in_row;
            *d_in = (char*)ctx->io[0] + (size_t)r * chunk * in_row;
            *d_out = (char*)ctx->io[1] + (size_t)r * chunk * 32;
        };
        uint64_t prev = UINT64_MAX;                             // this thread's previous chunk, results not yet copied out
        int j = 0;
        for (uint64_t k = i; k < n_chunks + 3 && e == cudaSuccess && !rcs[i]; k += 3, j++) {
            char *h_in, *h_out, *d_in, *d_out;
            if (k < n_chunks) {
                const uint64_t off = cut[k], c = cut[k + 1] - off;
                region(j & 1, &h_in, &h_out, &d_in, &d_out);
                // this half was last used by chunk j-2, whose results were copied out in iteration j-1
                memcpy(h_in, in + off * in_row, c * in_row);
                if ((e = cudaMemcpyAsync(d_in, h_in, c * in_row, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
                if ((rcs[i] = hash_batch_dev(ctx, n_inputs, flags, tag, d_in, c, d_out, st, dense, custom))) break;
                if ((e = cudaMemcpyAsync(h_out, d_out, c * 32, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
                if ((e = cudaEventRecord(ctx->bounce_done[i][j & 1], st)) != cudaSuccess) break;
            }
            if (prev != UINT64_MAX) {
                const uint64_t off = cut[prev], c = cut[prev + 1] - off;
                region((j - 1) & 1, &h_in, &h_out, &d_in, &d_out);
                if ((e = cudaEventSynchronize(ctx->bounce_done[i][(j - 1) & 1])) != cudaSuccess) break;
                memcpy(out + off * 32, h_out, c * 32);
            }
            prev = k < n_chunks ? k : UINT64_MAX;
        }
        errs[i] = e;
        cudaStreamSynchronize(st);                               // nothing of this call left in flight on return
    };
    {
        std::thread t1(worker, 1), t2(worker, 2);
        worker(0);
        t1.join();
        t2.join();
    }
    for (int i = 0; i < 3; i++) {
        if (rcs[i]) return rcs[i];
        if (errs[i] != cudaSuccess) return cuda_fail(ctx, errs[i], "pageable batch pipeline");
    }
    return INF_OK;
}

// Host-buffer batch: the batch is cut into chunks that rotate over three
// streams, each doing H2D -> kernel -> D2H for its chunk, so that with pinned
// host buffers the copies of one chunk hide behind the hashing of another (the
// path is compute-bound: ~8 ns of hashing per 96 bytes moved).
static int hash_batch_host(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags, const uint8_t* tag,
                           const uint8_t* in, uint64_t n, uint8_t* out, bool dense,
                           const CustomParams* custom = nullptr) {
    int rc = check_hash_args(ctx, n_inputs, in, n, out);
    if (rc) return rc;
    if (n == 0) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    if (n >= (1ull << 20) && (is_pageable(in) || is_pageable(out)))
        return hash_batch_pageable(ctx, n_inputs, flags, tag, in, n, out, dense, custom);
    const uint64_t super = 1ull << 24;                 // device staging is sized for at most 2^24 hashes
    const uint64_t chunk = 2 * wave_unit(ctx);
    const size_t in_row = (size_t)n_inputs * 32;
    const uint64_t n_stage = std::min<uint64_t>(n, super);
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], n_stage * in_row))) return rc;
    if ((rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], n_stage * 32))) return rc;
    const uint64_t lag = 2;     // read-back queued two chunks late, see rows_pipeline
    for (uint64_t base = 0; base < n; base += super) {
        const uint64_t m = std::min<uint64_t>(super, n - base);
        const uint64_t n_chunks = (m + chunk - 1) / chunk;
        const bool single = n_chunks == 1;
        for (uint64_t k = 0; k < n_chunks + lag; k++) {
            if (k < n_chunks) {
                const uint64_t off = k * chunk, c = std::min<uint64_t>(chunk, m - off);
                cudaStream_t st = single ? ctx->stream : ctx->pipe[k % 3];
                char* d_in = (char*)ctx->io[0] + off * in_row;
                CU(cudaMemcpyAsync(d_in, in + (base + off) * in_row, c * in_row, cudaMemcpyHostToDevice, st));
                if ((rc = hash_batch_dev(ctx, n_inputs, flags, tag, d_in, c, (char*)ctx->io[1] + off * 32, st, dense,
                                         custom)))
                    return rc;
            }
            if (k >= lag) {
                const uint64_t off = (k - lag) * chunk, c = std::min<uint64_t>(chunk, m - off);
                CU(cudaMemcpyAsync(out + (base + off) * 32, (char*)ctx->io[1] + off * 32, c * 32, cudaMemcpyDeviceToHost,
                                   single ? ctx->stream : ctx->pipe[(k - lag) % 3]));
            }
        }
        if (single) {
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            for (int i = 0; i < 3; i++) CU(cudaStreamSynchronize(ctx->pipe[i]));
        }
    }
    return INF_OK;
}

int inf_poseidon_hash_batch(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                            const uint8_t* domain_tag, const uint8_t* in, uint64_t n,
                            uint8_t* out) {
    return hash_batch_host(ctx, n_inputs, flags, domain_tag, in, n, out, false);
}

int inf_poseidon_hash_batch_dense(inf_ctx* ctx, uint32_t n_inputs, uint32_t flags,
                                  const uint8_t* domain_tag, const uint8_t* in, uint64_t n,
                                  uint8_t* out) {
    return hash_batch_host(ctx, n_inputs, flags, domain_tag, in, n, out, true);
}

int inf_poseidon_hash_batch_params(inf_ctx* ctx, uint32_t width, uint32_t full_rounds,
                                   uint32_t partial_rounds, uint64_t alpha, const uint8_t* ark,
                                   const uint8_t* mds, uint32_t flags, const uint8_t* domain_tag,
                                   const uint8_t* in, uint64_t n, uint8_t* out) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (width < 2 || width > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    const uint64_t rounds = (uint64_t)full_rounds + partial_rounds;
    if (rounds > 4096) return INF_ERR_BAD_DEPTH;
    if (!mds || (rounds && !ark)) return INF_ERR_NULL_POINTER;
    int rc = check_hash_args(ctx, width - 1, in, n, out);
    if (rc) return rc;
    if (n == 0) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    // the table the dense kernel reads: [ark rounds*width][mds width*width], Montgomery form
    const size_t n_ark = (size_t)rounds * width, n_el = n_ark + (size_t)width * width;
    std::vector<uint32_t> tbl(n_el * 8);
    for (size_t k = 0; k < n_el; k++) {
        const uint8_t* src = k < n_ark ? ark + 32 * k : mds + 32 * (k - n_ark);
        host::to_limbs32(host::to_mont(host::reduce256(host::from_be_bytes(src))), &tbl[8 * k]);
    }
    uint32_t* d_tbl = nullptr;
    CU(cudaMalloc((void**)&d_tbl, tbl.size() * 4));
    cudaError_t e = cudaMemcpy(d_tbl, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d_tbl);
        return cuda_fail(ctx, e, "upload custom parameters");
    }
    const CustomParams cp = {d_tbl, (int)full_rounds, (int)partial_rounds, alpha};
    rc = hash_batch_host(ctx, width - 1, flags, domain_tag, in, n, out, true, &cp);
    cudaFree(d_tbl);
    return rc;
}

int inf_poseidon_hash_bytes(inf_ctx* ctx, uint32_t flags, const uint8_t* domain_tag,
                            const uint8_t* const* inputs, const size_t* lens, uint32_t n_inputs,
                            uint8_t out[32]) {
    if (!ctx || !out || (n_inputs && (!inputs || !lens))) return INF_ERR_NULL_POINTER;
    if (n_inputs < 1 || n_inputs + 1 > 13) return INF_ERR_INVALID_WIDTH_CIRCOM;
    uint8_t buf[12 * 32];
    for (uint32_t i = 0; i < n_inputs; i++) {
        // validate_bytes_length (poseidon.rs:255-273) then the exact-length
        // check of bytes_to_prime_field_element (poseidon.rs:282-288), in the
        // order the iterator applies them: first failing input wins.
        if (lens[i] == 0) return INF_ERR_EMPTY_INPUT;
        if (lens[i] != 32) return INF_ERR_INVALID_INPUT_LENGTH;
        if (!inputs[i]) return INF_ERR_NULL_POINTER;
        memcpy(buf + 32 * i, inputs[i], 32);
    }
    return hash_batch_host(ctx, n_inputs, flags, domain_tag, buf, 1, out, false);
}

int inf_registration_leaves(inf_ctx* ctx, const uint8_t* public_keys, const uint64_t* timestamps,
                            uint64_t n, uint8_t* leaves) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!public_keys || !timestamps || !leaves)) return INF_ERR_NULL_POINTER;
    if (n == 0) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return rows_pipeline(ctx, public_keys, 64, (const uint8_t*)timestamps, 8, n, leaves,
                         [](const void* a, const void* b, void* o, uint64_t c, cudaStream_t st) {
                             return launch_registration_leaves(a, b, o, c, st);
                         });
}

int inf_interaction_leaves(inf_ctx* ctx, const uint8_t* public_keys, const uint8_t* data, uint64_t n,
                           uint8_t* leaves) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!public_keys || !data || !leaves)) return INF_ERR_NULL_POINTER;
    if (n == 0) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return rows_pipeline(ctx, public_keys, 64, data, 320, n, leaves,
                         [](const void* a, const void* b, void* o, uint64_t c, cudaStream_t st) {
                             return launch_interaction_leaves(a, b, o, c, st);
                         });
}

int inf_registration_leaves_dev(inf_ctx* ctx, const void* d_public_keys, const void* d_timestamps,
                                uint64_t n, void* d_leaves, void* stream) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!d_public_keys || !d_timestamps || !d_leaves)) return INF_ERR_NULL_POINTER;
    if (misaligned(d_public_keys) || misaligned(d_leaves) || ((uintptr_t)d_timestamps & 7)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(launch_registration_leaves(d_public_keys, d_timestamps, d_leaves, n,
                                  stream ? (cudaStream_t)stream : ctx->stream));
    return INF_OK;
}

int inf_interaction_leaves_dev(inf_ctx* ctx, const void* d_public_keys, const void* d_data,
                               uint64_t n, void* d_leaves, void* stream) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!d_public_keys || !d_data || !d_leaves)) return INF_ERR_NULL_POINTER;
    if (misaligned(d_public_keys) || misaligned(d_data) || misaligned(d_leaves)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(launch_interaction_leaves(d_public_keys, d_data, d_leaves, n, stream ? (cudaStream_t)stream : ctx->stream));
    return INF_OK;
}

int inf_host_alloc(inf_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    *out = nullptr;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return INF_OK;
}

int inf_host_free(inf_ctx* ctx, void* p) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (!p) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(cudaFreeHost(p));
    return INF_OK;
}

int inf_host_register(inf_ctx* ctx, void* p, size_t bytes) {
    if (!ctx || !p) return INF_ERR_NULL_POINTER;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return INF_OK;
}

int inf_host_unregister(inf_ctx* ctx, void* p) {
    if (!ctx || !p) return INF_ERR_NULL_POINTER;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(cudaHostUnregister(p));
    return INF_OK;
}

int inf_merkle_zeroes(inf_ctx* ctx, uint32_t arity, uint8_t out[33 * 32]) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    memcpy(out, ctx->zeroes[arity == 2 ? 0 : 1], 33 * 32);
    return INF_OK;
}

int inf_empty_ballot_roots(uint8_t out[5 * 32]) {
    if (!out) return INF_ERR_NULL_POINTER;
    memcpy(out, host::EMPTY_BALLOT_ROOTS_BE, 5 * 32);
    return INF_OK;
}

int inf_tree_merge_dev(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                       int to_depth, const void* d_leaves, uint64_t n_leaves, uint8_t root[32],
                       uint32_t* insert_depth, uint32_t* root_depth, int* has_root, void* stream) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (misaligned(d_leaves)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return tree_merge_dev(ctx, arity, full_depth, prepend_blank_leaf, to_depth, d_leaves, n_leaves,
                          root, insert_depth, root_depth, has_root,
                          stream ? (cudaStream_t)stream : ctx->stream);
}

int inf_tree_merge(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                   int to_depth, const uint8_t* leaves, uint64_t n_leaves, uint8_t root[32],
                   uint32_t* insert_depth, uint32_t* root_depth, int* has_root) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n_leaves && !leaves) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    // capacity check before moving any data (insert would have failed first)
    if (n_leaves + (prepend_blank_leaf ? 1 : 0) > pow_sat(arity, full_depth)) {
        if (insert_depth) *insert_depth = 0;
        if (root_depth) *root_depth = 0;
        if (has_root) *has_root = 0;
        return INF_ERR_TREE_ALREADY_FULL;
    }
    if (n_leaves) {
        int rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], (size_t)n_leaves * 32);
        if (rc) return rc;
    }
    const LeafFeed feed = host_leaf_feed(ctx, leaves);
    return tree_merge_dev(ctx, arity, full_depth, prepend_blank_leaf, to_depth, ctx->io[0], n_leaves,
                          root, insert_depth, root_depth, has_root, ctx->stream, &feed);
}

int inf_tree_reduce_dev(inf_ctx* ctx, uint32_t arity, uint32_t level_in, uint32_t n_levels,
                        uint64_t shift, const void* d_in, uint64_t n_in, void* d_out,
                        uint64_t* n_out, void* stream) {
    if (!ctx || !d_out) return INF_ERR_NULL_POINTER;
    if (n_in && !d_in) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (level_in + n_levels > 32) return INF_ERR_BAD_DEPTH;
    if (misaligned(d_in) || misaligned(d_out)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
    const uint64_t n_total = n_in + shift;
    if (n_total == 0) {
        if (n_out) *n_out = 0;
        return INF_OK;
    }
    if (n_levels == 0) {
        for (uint64_t i = 0; i < shift; i++)
            CU(cudaMemcpyAsync((char*)d_out + 32 * i, Z[level_in], 32, cudaMemcpyHostToDevice, st));
        if (n_in) CU(cudaMemcpyAsync((char*)d_out + 32 * shift, d_in, (size_t)n_in * 32, cudaMemcpyDeviceToDevice, st));
        if (n_out) *n_out = n_total;
        return INF_OK;
    }
    const uint64_t n1 = (n_total + arity - 1) / arity;
    const uint64_t n2 = (n1 + arity - 1) / arity;
    int rc;
    if (n_levels > 1 && (rc = grow(ctx, &ctx->scratch[0], &ctx->scratch_bytes[0], n1 * 32))) return rc;
    if (n_levels > 2 && (rc = grow(ctx, &ctx->scratch[1], &ctx->scratch_bytes[1], n2 * 32))) return rc;
    const std::function<char*(uint32_t)> dst_of = [&](uint32_t l) -> char* {
        return (char*)((l + 1 == n_levels) ? d_out : ctx->scratch[l & 1]);
    };
    uint64_t n_res = 0;
    if ((rc = reduce_levels(ctx, arity, level_in, n_levels, shift, d_in, n_in, nullptr, dst_of, st, nullptr, &n_res))) return rc;
    if (n_out) *n_out = n_res;
    return INF_OK;
}

int inf_tree_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                      const uint8_t* leaves, uint64_t n_leaves, uint8_t* out_levels,
                      uint8_t* out_hashes, uint32_t cap, uint32_t* n_entries,
                      uint32_t* insert_depth, int* has_root, uint8_t root[32]) {
    if (!ctx || !n_entries) return INF_ERR_NULL_POINTER;
    if (n_leaves && !leaves) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    // PollStateTree::new seeds `hashes` with the blank leaf (state.rs:150-158): a frontier of one entry
    const uint8_t lvl0 = 0;
    const uint8_t* z0 = ctx->zeroes[arity == 2 ? 0 : 1][0];
    return tree_append_core(ctx, arity, full_depth, &lvl0, z0, prepend_blank_leaf ? 1 : 0, 0, leaves, nullptr,
                            n_leaves, out_levels, out_hashes, cap, n_entries, insert_depth, has_root, root,
                            ctx->stream);
}

int inf_tree_append(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels,
                    const uint8_t* in_hashes, uint32_t n_in, uint32_t depth_in, const uint8_t* leaves,
                    uint64_t n_leaves, uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap,
                    uint32_t* n_entries, uint32_t* depth_out, int* has_root, uint8_t root[32]) {
    if (!ctx || !n_entries) return INF_ERR_NULL_POINTER;
    if ((n_leaves && !leaves) || (n_in && (!in_levels || !in_hashes))) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return tree_append_core(ctx, arity, full_depth, in_levels, in_hashes, n_in, depth_in, leaves, nullptr, n_leaves,
                            out_levels, out_hashes, cap, n_entries, depth_out, has_root, root, ctx->stream);
}

int inf_tree_append_dev(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* in_levels,
                        const uint8_t* in_hashes, uint32_t n_in, uint32_t depth_in, const void* d_leaves,
                        uint64_t n_leaves, uint8_t* out_levels, uint8_t* out_hashes, uint32_t cap,
                        uint32_t* n_entries, uint32_t* depth_out, int* has_root, uint8_t root[32],
                        void* stream) {
    if (!ctx || !n_entries) return INF_ERR_NULL_POINTER;
    if ((n_leaves && !d_leaves) || (n_in && (!in_levels || !in_hashes))) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    if (misaligned(d_leaves)) return INF_ERR_BAD_ALIGNMENT;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    return tree_append_core(ctx, arity, full_depth, in_levels, in_hashes, n_in, depth_in, nullptr, d_leaves, n_leaves,
                            out_levels, out_hashes, cap, n_entries, depth_out, has_root, root,
                            stream ? (cudaStream_t)stream : ctx->stream);
}

int inf_tree_merge_frontier(inf_ctx* ctx, uint32_t arity, uint32_t full_depth, const uint8_t* levels,
                            const uint8_t* hashes, uint32_t n, int to_depth, uint8_t root[32], int* has_root,
                            uint32_t* root_depth) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!levels || !hashes)) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    if (has_root) *has_root = 0;
    if (root_depth) *root_depth = 0;
    uint64_t n_old = 0;
    uint32_t cnt[33];
    int rc = parse_frontier(arity, full_depth, levels, n, cnt, &n_old);
    if (rc) return rc;
    if (n == 0) return INF_OK;                                     // merge on an empty frontier: root stays None
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    if ((rc = ensure_front(ctx))) return rc;
    cudaStream_t st = ctx->stream;
    const uint8_t(*Z)[32] = ctx->zeroes[arity == 2 ? 0 : 1];
    char* E = (char*)ctx->d_front;
    char* tmp = E + FRONT_NODES * 32;
    CU(cudaMemcpyAsync(E, hashes, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    // PollStateTree::merge (state.rs:240-271): hash the trailing run of equal-level entries, padded
    // with that level's zero, into one entry a level up, until a single entry is left (and, with
    // to_depth, until it sits at full_depth).  Each step consumes the previous one's output: a
    // chain of at most ~full_depth single hashes, enqueued back to back without host round trips.
    std::vector<uint8_t> lv(levels, levels + n);
    for (;;) {
        const uint8_t d = lv.back();
        if (lv.size() == 1 && (!to_depth || d == full_depth)) break;
        if (d >= 32 || d >= full_depth) return INF_ERR_BAD_FRONTIER;
        size_t size = 0;
        while (size < lv.size() && lv[lv.size() - 1 - size] == d) size++;
        const size_t i = lv.size() - size;
        CU(launch_level(arity, E + i * 32, 0, size, tmp, 1, Z[d], st));
        CU(cudaMemcpyAsync(E + i * 32, tmp, 32, cudaMemcpyDeviceToDevice, st));
        lv.resize(i);
        lv.push_back((uint8_t)(d + 1));
    }
    uint8_t r[32];
    CU(cudaMemcpyAsync(r, E, 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (root) memcpy(root, r, 32);
    if (has_root) *has_root = 1;
    if (root_depth) *root_depth = lv[0];
    return INF_OK;
}

int inf_tree_build(inf_ctx* ctx, uint32_t arity, uint32_t depth, int prepend_blank_leaf,
                   const uint8_t* leaves, uint64_t n_leaves, inf_tree** out) {
    if (!ctx || !out) return INF_ERR_NULL_POINTER;
    *out = nullptr;
    if (n_leaves && !leaves) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (depth > 32) return INF_ERR_BAD_DEPTH;
    const uint64_t shift = prepend_blank_leaf ? 1 : 0, n_total = n_leaves + shift;
    if (n_total > pow_sat(arity, depth)) return INF_ERR_TREE_ALREADY_FULL;
    if (n_total == 0) return INF_ERR_MERGE_FAILED;             // nothing to build a tree over
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const LeafFeed feed = host_leaf_feed(ctx, leaves);
    int has = 0;
    int rc = tree_merge_dev(ctx, arity, depth, prepend_blank_leaf, 1, nullptr, n_leaves, nullptr, nullptr, nullptr, &has,
                            ctx->stream, &feed, out);
    return rc == INF_ERR_TREE_ALREADY_MERGED ? INF_OK : rc;    // exactly arity^depth leaves is a full tree, not an error
}

int inf_tree_root(inf_tree* tree, uint8_t root[32]) {
    if (!tree || !root) return INF_ERR_NULL_POINTER;
    inf_ctx* ctx = tree->ctx;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    CU(cudaMemcpy(root, (char*)tree->d_nodes + tree->offsets[tree->depth] * 32, 32, cudaMemcpyDeviceToHost));
    return INF_OK;
}

int inf_tree_paths(inf_tree* tree, const uint64_t* leaf_indices, uint64_t n_idx, uint8_t* paths) {
    return inf_tree_node_paths(tree, 0, leaf_indices, n_idx, paths);
}

int inf_tree_node_paths(inf_tree* tree, uint32_t level, const uint64_t* node_indices, uint64_t n_idx, uint8_t* paths) {
    if (!tree) return INF_ERR_NULL_POINTER;
    if (level > tree->depth) return INF_ERR_BAD_DEPTH;
    if (n_idx && (!node_indices || !paths)) return INF_ERR_NULL_POINTER;
    const uint32_t levels_up = tree->depth - level;
    if (n_idx == 0 || levels_up == 0) return INF_OK;
    inf_ctx* ctx = tree->ctx;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const uint64_t cap = pow_sat(tree->arity, levels_up);
    for (uint64_t i = 0; i < n_idx; i++)
        if (node_indices[i] >= cap) return INF_ERR_BAD_DEPTH;
    const size_t out_bytes = (size_t)n_idx * levels_up * (tree->arity - 1) * 32;
    int rc;
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], n_idx * 8))) return rc;
    if ((rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], out_bytes))) return rc;
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->io[0], node_indices, n_idx * 8, cudaMemcpyHostToDevice, st));
    CU(launch_gather_paths((const void* const*)tree->d_level_ptrs + level, (const uint64_t*)tree->d_level_counts + level,
                           (const char*)tree->d_zero_nodes + (size_t)level * 32, tree->arity, levels_up, ctx->io[0], n_idx,
                           ctx->io[1], st));
    CU(cudaMemcpyAsync(paths, ctx->io[1], out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return INF_OK;
}

int inf_tree_level_nodes(inf_tree* tree, uint32_t level, uint64_t first, uint64_t count, uint8_t* out) {
    if (!tree) return INF_ERR_NULL_POINTER;
    if (level > tree->depth) return INF_ERR_BAD_DEPTH;
    if (count && !out) return INF_ERR_NULL_POINTER;
    if (count == 0) return INF_OK;
    inf_ctx* ctx = tree->ctx;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const uint64_t cap = pow_sat(tree->arity, tree->depth - level), have = tree->counts[level];
    if (first > cap || count > cap - first) return INF_ERR_BAD_DEPTH;
    // stored nodes first, then the level's zero value for the all-zero subtrees to their right
    const uint64_t n_real = first < have ? std::min<uint64_t>(count, have - first) : 0;
    if (n_real)
        CU(cudaMemcpy(out, (const char*)tree->d_nodes + (tree->offsets[level] + first) * 32, (size_t)n_real * 32,
                      cudaMemcpyDeviceToHost));
    const uint8_t* z = ctx->zeroes[tree->arity == 2 ? 0 : 1][level];
    for (uint64_t i = n_real; i < count; i++) memcpy(out + 32 * i, z, 32);
    return INF_OK;
}

void inf_tree_destroy(inf_tree* tree) {
    if (!tree) return;
    {
        Bind bind(tree->ctx);
        if (tree->d_nodes) cudaFree(tree->d_nodes);
        if (tree->d_level_ptrs) cudaFree(tree->d_level_ptrs);
        if (tree->d_level_counts) cudaFree(tree->d_level_counts);
        if (tree->d_zero_nodes) cudaFree(tree->d_zero_nodes);
    }
    delete tree;
}

int inf_merkle_roots_from_paths(inf_ctx* ctx, uint32_t arity, uint32_t depth, const uint64_t* indices,
                                const uint8_t* leaves, const uint8_t* paths, uint64_t n,
                                uint8_t* roots) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (depth > 32) return INF_ERR_BAD_DEPTH;
    if (n && (!indices || !leaves || !roots || (depth && !paths))) return INF_ERR_NULL_POINTER;
    if (n == 0) return INF_OK;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const size_t path_bytes = (size_t)n * depth * (arity - 1) * 32;
    const size_t idx_bytes = (n * 8 + 31) & ~(size_t)31;          // keep the nodes 32-byte aligned
    int rc;
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], idx_bytes + n * 32 + path_bytes))) return rc;
    if ((rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], n * 32))) return rc;
    char* d_idx = (char*)ctx->io[0];
    char* d_leaves = d_idx + idx_bytes;
    char* d_paths = d_leaves + n * 32;
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_idx, indices, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_leaves, leaves, n * 32, cudaMemcpyHostToDevice, st));
    if (path_bytes) CU(cudaMemcpyAsync(d_paths, paths, path_bytes, cudaMemcpyHostToDevice, st));
    CU(arity == 2 ? launch_path_root_t3(d_idx, d_leaves, d_paths, depth, ctx->io[1], n, st)
                  : launch_path_root_t6(d_idx, d_leaves, d_paths, depth, ctx->io[1], n, st));
    CU(cudaMemcpyAsync(roots, ctx->io[1], n * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return INF_OK;
}

// ---- replay from raw inputs: leaf hashing chained into the tree on the device ------------------
int inf_replay_registrations(inf_ctx* ctx, uint32_t registration_depth, const uint8_t* public_keys,
                             const uint64_t* timestamps, uint64_t n, uint8_t root[32],
                             uint8_t process_commitment[32], uint32_t* insert_depth, uint8_t* leaves_out,
                             inf_tree** retained) {
    if (!ctx || !root || !process_commitment) return INF_ERR_NULL_POINTER;
    if (n && (!public_keys || !timestamps)) return INF_ERR_NULL_POINTER;
    if (retained) *retained = nullptr;
    if (registration_depth > 32) return INF_ERR_BAD_DEPTH;
    if (n + 1 > pow_sat(2, registration_depth)) return INF_ERR_TREE_ALREADY_FULL;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const uint64_t chunk_out = 2 * wave_unit(ctx);     // participants per chunk: whole waves of the leaf kernel
    const size_t slot_rows = ((size_t)std::min<uint64_t>(chunk_out, n + 2) + 3) & ~(size_t)3;   // rows one pipeline stream stages at a time (keeps the slots 16-byte aligned)
    int rc;
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], 3 * slot_rows * 72))) return rc;
    if (!retained && (rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], std::max<uint64_t>(n, 1) * 32))) return rc;
    LeafFeed feed = raw_rows_feed(ctx, public_keys, 64, (const uint8_t*)timestamps, 8, chunk_out, slot_rows,
                                  [](const void* a, const void* b, void* o, uint64_t c, cudaStream_t st) {
                                      return launch_registration_leaves(a, b, o, c, st);
                                  });
    feed.leaves_out = leaves_out;
    int has = 0;
    uint32_t rdepth = 0;
    // register_participant x n (provider.rs:218-241), then merge_registrations (provider.rs:289-311)
    rc = tree_merge_dev(ctx, 2, registration_depth, 1, 0, ctx->io[1], n, root, insert_depth, &rdepth, &has, ctx->stream, &feed,
                        retained);
    if (rc) {
        if (retained && *retained) { inf_tree_destroy(*retained); *retained = nullptr; }
        return rc;
    }
    if (!has) return INF_ERR_MERGE_FAILED;
    uint8_t in[3 * 32];
    memcpy(in, root, 32);
    memcpy(in + 32, host::EMPTY_BALLOT_ROOTS_BE[1], 32);
    memset(in + 64, 0, 32);
    rc = inf_poseidon_hash_batch(ctx, 3, 0, nullptr, in, 1, process_commitment);
    return rc ? INF_ERR_HASH_FAILED : INF_OK;
}

int inf_replay_interactions(inf_ctx* ctx, uint32_t interaction_depth, const uint8_t* public_keys,
                            const uint8_t* data, uint64_t n, uint32_t registrations_count,
                            uint32_t process_subtree_depth, uint32_t tally_subtree_depth, uint8_t root[32],
                            int* has_root, uint32_t* insert_depth, uint32_t* expected_process,
                            uint32_t* expected_tally, uint8_t* leaves_out, inf_tree** retained) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    if (n && (!public_keys || !data)) return INF_ERR_NULL_POINTER;
    if (retained) *retained = nullptr;
    if (has_root) *has_root = 0;
    if (interaction_depth > 32) return INF_ERR_BAD_DEPTH;
    if (n > pow_sat(5, interaction_depth)) return INF_ERR_TREE_ALREADY_FULL;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    static const int waves = getenv("INF_REPLAY_WAVES") ? atoi(getenv("INF_REPLAY_WAVES")) : 4;
    const uint64_t chunk_out = (uint64_t)(waves > 0 ? waves : 4) * (wave_unit(ctx) / 4);   // messages per chunk: whole waves of 384-thread blocks
    const size_t slot_rows = ((size_t)std::min<uint64_t>(chunk_out, n + 5) + 3) & ~(size_t)3;
    int rc;
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], 3 * slot_rows * 384))) return rc;
    if (!retained && (rc = grow(ctx, &ctx->io[1], &ctx->io_bytes[1], std::max<uint64_t>(n, 1) * 32))) return rc;
    LeafFeed feed = raw_rows_feed(ctx, public_keys, 64, data, 320, chunk_out, slot_rows,
                                  [](const void* a, const void* b, void* o, uint64_t c, cudaStream_t st) {
                                      return launch_interaction_leaves(a, b, o, c, st);
                                  });
    feed.leaves_out = leaves_out;
    uint32_t rdepth = 0;
    // consume_interaction x n (provider.rs:243-287), then merge_interactions (provider.rs:313-327)
    rc = tree_merge_dev(ctx, 5, interaction_depth, 0, 1, ctx->io[1], n, root, insert_depth, &rdepth, has_root, ctx->stream,
                        &feed, retained);
    if (rc) {
        if (retained && *retained) { inf_tree_destroy(*retained); *retained = nullptr; }
        return rc;
    }
    const uint32_t count = (uint32_t)n;
    const uint32_t pb = (uint32_t)pow_sat(5, process_subtree_depth);
    const uint32_t tb = (uint32_t)pow_sat(2, tally_subtree_depth);
    if (expected_process) *expected_process = pb ? count / pb + ((count % pb) ? 1u : 0u) : 0u;
    if (expected_tally) *expected_tally = tb ? 1u + registrations_count / tb : 0u;
    return INF_OK;
}

int inf_merge_registrations(inf_ctx* ctx, uint32_t registration_depth, const uint8_t* leaves,
                            uint64_t n_leaves, uint8_t root[32], uint8_t process_commitment[32],
                            uint32_t* insert_depth) {
    if (!ctx || !root || !process_commitment) return INF_ERR_NULL_POINTER;
    int has = 0;
    uint32_t rdepth = 0;
    // registrations.merge(false)  (provider.rs:293)
    int rc = inf_tree_merge(ctx, 2, registration_depth, 1, 0, leaves, n_leaves, root, insert_depth,
                            &rdepth, &has);
    if (rc) return rc;
    if (!has) return INF_ERR_MERGE_FAILED;                         // provider.rs:295
    // commitment.process = (0, H3(root, EMPTY_BALLOT_ROOTS[1], 0))  (provider.rs:296-308)
    uint8_t in[3 * 32];
    memcpy(in, root, 32);
    memcpy(in + 32, host::EMPTY_BALLOT_ROOTS_BE[1], 32);
    memset(in + 64, 0, 32);
    rc = inf_poseidon_hash_batch(ctx, 3, 0, nullptr, in, 1, process_commitment);
    return rc ? INF_ERR_HASH_FAILED : INF_OK;
}

int inf_merge_interactions(inf_ctx* ctx, uint32_t interaction_depth, const uint8_t* leaves,
                           uint64_t n_leaves, uint32_t registrations_count,
                           uint32_t process_subtree_depth, uint32_t tally_subtree_depth,
                           uint8_t root[32], int* has_root, uint32_t* expected_process,
                           uint32_t* expected_tally) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    uint32_t idepth = 0, rdepth = 0;
    // interactions.merge(true)  (provider.rs:317)
    int rc = inf_tree_merge(ctx, 5, interaction_depth, 0, 1, leaves, n_leaves, root, &idepth, &rdepth,
                            has_root);
    if (rc) return rc;
    // provider.rs:319-324 (u32 arithmetic as in the reference)
    const uint32_t count = (uint32_t)n_leaves;
    const uint32_t pb = (uint32_t)pow_sat(5, process_subtree_depth);
    const uint32_t tb = (uint32_t)pow_sat(2, tally_subtree_depth);
    if (expected_process) *expected_process = pb ? count / pb + ((count % pb) ? 1u : 0u) : 0u;
    if (expected_tally) *expected_tally = tb ? 1u + registrations_count / tb : 0u;
    return INF_OK;
}

// ---- internal (multi.cu) -----------------------------------------------------------------
int inf_internal_stream(inf_ctx* ctx, void** stream) {
    if (!ctx || !stream) return INF_ERR_NULL_POINTER;
    *stream = (void*)ctx->stream;
    return INF_OK;
}
// Reduce a run of level-0 nodes that is still in host memory by n_levels (>= 1) levels into d_out:
// chunked upload overlapped with level-0 hashing on the context's pipeline streams, the rest on its
// own stream.  Enqueues only; the caller synchronises the context's stream (inf_internal_drain).
int inf_internal_tree_reduce_host(inf_ctx* ctx, uint32_t arity, uint32_t n_levels, uint64_t shift,
                                  const uint8_t* h_in, uint64_t n_in, void* d_out, uint64_t* n_out) {
    if (!ctx || !d_out || (n_in && !h_in)) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (n_levels < 1 || n_levels > 32) return INF_ERR_BAD_DEPTH;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    const uint64_t n_total = n_in + shift;
    if (n_total == 0) {
        if (n_out) *n_out = 0;
        return INF_OK;
    }
    const uint64_t n1 = (n_total + arity - 1) / arity, n2 = (n1 + arity - 1) / arity;
    int rc;
    if ((rc = grow(ctx, &ctx->io[0], &ctx->io_bytes[0], std::max<uint64_t>(n_in, 1) * 32))) return rc;
    if (n_levels > 1 && (rc = grow(ctx, &ctx->scratch[0], &ctx->scratch_bytes[0], n1 * 32))) return rc;
    if (n_levels > 2 && (rc = grow(ctx, &ctx->scratch[1], &ctx->scratch_bytes[1], n2 * 32))) return rc;
    const LeafFeed feed = host_leaf_feed(ctx, h_in);
    const std::function<char*(uint32_t)> dst_of = [&](uint32_t l) -> char* {
        return (char*)((l + 1 == n_levels) ? d_out : ctx->scratch[l & 1]);
    };
    return reduce_levels(ctx, arity, 0, n_levels, shift, ctx->io[0], n_in, &feed, dst_of, ctx->stream, nullptr, n_out);
}
// Wait for everything queued on the context's streams (also after an error).
int inf_internal_drain(inf_ctx* ctx) {
    if (!ctx) return INF_ERR_NULL_POINTER;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    cudaError_t e = cudaSuccess, e2;
    for (int i = 0; i < 3; i++)
        if ((e2 = cudaStreamSynchronize(ctx->pipe[i])) != cudaSuccess) e = e2;
    if ((e2 = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) e = e2;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "drain");
    return INF_OK;
}
int inf_internal_grow_io(inf_ctx* ctx, int which, size_t bytes, void** ptr) {
    if (!ctx || !ptr || which < 0 || which > 1) return INF_ERR_NULL_POINTER;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    int rc = grow(ctx, &ctx->io[which], &ctx->io_bytes[which], bytes);
    *ptr = ctx->io[which];
    return rc;
}

int inf_debug_dense_params(uint32_t t, uint32_t* out, size_t out_words) {
    if (t < 2 || t > 13 || !out) return -1;
    const host::DenseParams& d = host::grain_params((int)t);
    const size_t n = d.ark.size() + d.mds.size();
    if (out_words < n * 8) return -1;
    size_t k = 0;
    for (const host::F& x : d.ark) host::to_limbs32(x, out + 8 * k++);
    for (const host::F& x : d.mds) host::to_limbs32(x, out + 8 * k++);
    return (int)n;
}

int inf_debug_opt_table(uint32_t t, uint32_t* out, size_t out_words) {
    if (t < 2 || t > 8 || !out) return -1;
    std::vector<uint32_t> v = host::build_opt_table((int)t);
    if (out_words < v.size()) return -1;
    memcpy(out, v.data(), v.size() * 4);
    return (int)v.size();
}

int inf_measure_imad_peak(inf_ctx* ctx, int kind, double* imad_per_s, double* sm_clock_mhz) {
    if (!ctx || !imad_per_s) return INF_ERR_NULL_POINTER;
    Bind bind(ctx);
    if (!bind.ok) return INF_ERR_NO_DEVICE;
    double clk = 0;
    CU(launch_imad_peak(kind, ctx->sm_count, imad_per_s, &clk, ctx->stream));
    if (sm_clock_mhz) *sm_clock_mhz = clk;
    return INF_OK;
}

}  // extern "C"
