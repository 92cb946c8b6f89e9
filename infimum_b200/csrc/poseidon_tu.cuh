// Per-width translation unit body.  Each poseidon_tN.cu defines INF_T and
// includes this file, so that every width has its own 64 KB __constant__ bank
// for its table (Layout<T>) and the widths compile in parallel.
//
// Kernels (one hash per thread; the path is bound by the integer-multiply
// pipe, ~2300 IMAD per byte moved, so memory layout only has to be sane:
// each thread moves whole 32-byte sectors with 128-bit loads/stores):
//   hash_batch_kernel   n x (T-1) x 32 B  ->  n x 32 B        K1 in SURVEY.md
//   tree_level_kernel   one tree level, arity T-1, zero padded  K2
#include <cuda_runtime.h>

#include <cstdlib>

#include "coop.cuh"
#include "launch.h"
#include "poseidon.cuh"

#ifndef INF_T
#error "define INF_T before including poseidon_tu.cuh"
#endif

namespace inf {
namespace {

constexpr int T = INF_T;
__constant__ uint32_t c_tbl[Layout<T>::WORDS];

struct Node32 {
    uint32_t w[8];
};

__device__ __forceinline__ void load_node(uint32_t (&w)[8], const uint4* p) {
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void store_node(uint4* p, const uint32_t (&w)[8]) {
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <bool LE>
__global__ void __launch_bounds__(INF_MAX_BLOCK(INF_T), 1)
hash_batch_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n, TagArg tag) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
#ifdef INF_LOCKSTEP
    const bool live = idx < n;          // every thread runs along to the barriers
    if (!live) idx = n - 1;
#else
    constexpr bool live = true;
    if (idx >= n) return;
#endif
    uint32_t iw[T - 1][8], ow[8];
    const uint4* p = in + idx * (uint64_t)(2 * (T - 1));
#pragma unroll
    for (int i = 0; i < T - 1; i++) load_node(iw[i], p + 2 * i);
    hash_words<T, LE>(ow, iw, tag.has ? tag.w : nullptr, c_tbl);
    if (live) store_node(out + 2 * idx, ow);
}

// One level of an arity-(T-1) Merkle tree over big-endian 32-byte nodes.
// The level's logical node array is `shift` leading nodes followed by
// in[0..n_in).  The leading nodes are prefix[0..shift) if `prefix` is given —
// the entries a stored frontier holds at this level, which open the first,
// incomplete group when leaves are appended to it (state.rs:176-225) — else
// copies of the zero value: shift = 1 at level 0 of the registration tree,
// whose leaf 0 is the blank state leaf = zeroes[0] (state.rs:48-52).
// out[i] = H(node[A*i], ..., node[A*i+A-1]); nodes past the end are the
// level's zero value (PollStateTree::merge's right padding, state.rs:262-266).
__global__ void __launch_bounds__(INF_MAX_BLOCK(INF_T), 1)
tree_level_kernel(const uint4* __restrict__ in, const uint4* __restrict__ prefix, uint64_t shift, uint64_t n_in,
                  uint4* __restrict__ out, uint64_t n_out, Node32 zero) {
    constexpr int A = T - 1;
    griddep_launch_dependents();
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out) return;
    uint32_t iw[A][8], ow[8];
    const uint64_t first = idx * A;
    griddep_wait();                         // the level below is complete and visible
#pragma unroll
    for (int i = 0; i < A; i++) {
        const uint64_t j = first + i;
        if (j >= shift && j - shift < n_in) {
            load_node(iw[i], in + 2 * (j - shift));
        } else if (j < shift && prefix) {
            load_node(iw[i], prefix + 2 * j);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) iw[i][k] = zero.w[k];
        }
    }
    hash_words<T, false>(ow, iw, nullptr, c_tbl);
    store_node(out + 2 * idx, ow);
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative tree level, for the levels near the root (schedule: coop.cuh).
//
// T + 1 warps share 32 hashes (lane = hash): warp roles are the roles of
// coop.cuh, published values live in shared memory as [limb][lane] (consecutive
// lanes hit consecutive banks), and roles wait for each other on named barriers
// (producer bar.arrive, consumer bar.sync), so that the chain warp never waits
// for anything but the value it needs next.  Block-wide barriers only in code
// every warp runs through.
constexpr int COOP_WARPS = T + 1;
// The chain warp should have a sub-partition to itself: warps map to sub-partitions
// by index mod 4, so with 5..7 warps that is warp 3.
constexpr int COOP_CHAIN_WARP = (COOP_WARPS >= 5 && COOP_WARPS <= 7) ? 3 : 0;

struct CoopSmem {
    uint32_t x[2][T][8][32];
    uint32_t z[3][8][32];
    uint32_t v[2][8][32];
    uint32_t p[3][T][8][32];
};

struct CoopBus {
    typedef uint32_t (*Slot)[32];
    CoopSmem& sm;
    const int lane;
    enum { BAR_Z = 1, BAR_V = 4, BAR_P = 6, BAR_PRO = 9 };
    __device__ __forceinline__ Slot x(int buf, int i) const { return sm.x[buf][i]; }
    __device__ __forceinline__ Slot z(int par) const { return sm.z[par]; }
    __device__ __forceinline__ Slot v(int par) const { return sm.v[par]; }
    __device__ __forceinline__ Slot p(int par, int i) const { return sm.p[par][i]; }
    __device__ __forceinline__ void put(Slot d, const uint32_t (&val)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; k++) d[k][lane] = val[k];
    }
    __device__ __forceinline__ void get(uint32_t (&val)[8], Slot s) const {
#pragma unroll
        for (int k = 0; k < 8; k++) val[k] = s[k][lane];
    }
    static __device__ __forceinline__ void arrive(int id, int count) {
        asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
    }
    static __device__ __forceinline__ void wait(int id, int count) {
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    }
    __device__ __forceinline__ void block() const { __syncthreads(); }
    __device__ __forceinline__ void z_arrive(int par) const { arrive(BAR_Z + par, 32 * (T + 1)); }
    __device__ __forceinline__ void z_wait(int par) const { wait(BAR_Z + par, 32 * (T + 1)); }
    __device__ __forceinline__ void v_arrive(int par) const { arrive(BAR_V + par, 64); }
    __device__ __forceinline__ void v_wait(int par) const { wait(BAR_V + par, 64); }
    __device__ __forceinline__ void p_arrive(int par) const { arrive(BAR_P + par, 32 * T); }
    __device__ __forceinline__ void p_wait(int par) const { wait(BAR_P + par, 32 * T); }
    __device__ __forceinline__ void pro_arrive() const { arrive(BAR_PRO, 32 * T); }
    __device__ __forceinline__ void pro_wait() const { wait(BAR_PRO, 32 * T); }
};

// Role of this warp.  Blocks that share an SM all put their chain warp on the same sub-partition
// (warps map to sub-partitions by index mod 4) and that is the better layout: chain warps are
// latency-bound and interleave with each other almost for free, while next to the helper warps
// of another block they are delayed by work that is never critical.  Rotating the layout by the
// order in which blocks arrive on their SM was measured slower (8 192 parents 116 -> 127 us,
// 16 384 parents 203 -> 232 us; profiles/r02_tree_levels.md).
__device__ __forceinline__ int coop_role() {
    const int w = threadIdx.x >> 5;
    return w == COOP_CHAIN_WARP ? 0 : (w == 0 ? COOP_CHAIN_WARP : w);
}

__device__ __forceinline__ void load_node_coherent(uint32_t (&w)[8], const uint4* p) {
    const uint4 a = __ldcg(p), b = __ldcg(p + 1);          // L2, never the non-coherent path
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

__global__ void __launch_bounds__(32 * COOP_WARPS, 1)
tree_level_coop_kernel(const uint4* __restrict__ in, const uint4* __restrict__ prefix, uint64_t shift, uint64_t n_in,
                       uint4* __restrict__ out, uint64_t n_out, Node32 zero) {
    using L = Layout<T>;
    constexpr int A = T - 1;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    CoopBus bus{*reinterpret_cast<CoopSmem*>(coop_raw), (int)(threadIdx.x & 31)};
    griddep_launch_dependents();
    const int role = coop_role();
    const uint64_t h = (uint64_t)blockIdx.x * 32 + bus.lane;
    const bool live = h < n_out;          // dead lanes run along on zeros (barriers count whole warps)
    const uint32_t* tbl = c_tbl;

    // ---- absorb: role i in 1..T-1 takes child i-1 of its hash ------------------
    uint32_t s[8];
    if (role == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = tbl[L::S0 * 8 + k];
    } else if (role < T) {
        uint32_t wd[8], raw[8];
        const uint64_t j = h * A + (role - 1);
        griddep_wait();                     // the level below is complete and visible
        if (live && j >= shift && j - shift < n_in) {
            load_node(wd, in + 2 * (j - shift));
        } else if (live && j < shift && prefix) {
            load_node(wd, prefix + 2 * j);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) wd[k] = zero.w[k];
        }
        words_to_limbs<false>(raw, wd);
        absorb<T>(s, raw, role, tbl);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = 0;
    }
    uint32_t hsh[8];
    coop_hash<T>(hsh, s, role, bus, tbl);
    if (role == 0 && live) {
        uint32_t ow[8];
        limbs_to_words<false>(ow, hsh);
        store_node(out + 2 * h, ow);
    }
}

// The same formulation for small batches (PoseidonHasher::hash called for a handful of
// inputs: the commitments, the coordinator-key hash, verify_outcome): one hash costs the
// critical path of the cooperative schedule instead of a whole single-thread hash.
template <bool LE>
__global__ void __launch_bounds__(32 * COOP_WARPS, 1)
hash_batch_coop_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n, TagArg tag) {
    using L = Layout<T>;
    constexpr int A = T - 1;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    CoopBus bus{*reinterpret_cast<CoopSmem*>(coop_raw), (int)(threadIdx.x & 31)};
    const int role = coop_role();
    const uint64_t h = (uint64_t)blockIdx.x * 32 + bus.lane;
    const bool live = h < n;
    const uint32_t* tbl = c_tbl;
    uint32_t s[8];
    if (role < T && (role > 0 || tag.has)) {
        uint32_t wd[8], raw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) wd[k] = role == 0 ? tag.w[k] : 0u;
        if (role > 0 && live) load_node(wd, in + 2 * (h * A + (role - 1)));
        words_to_limbs<LE>(raw, wd);
        absorb<T>(s, raw, role, tbl);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = role == 0 ? tbl[L::S0 * 8 + k] : 0u;
    }
    uint32_t hsh[8];
    coop_hash<T>(hsh, s, role, bus, tbl);
    if (role == 0 && live) {
        uint32_t ow[8];
        limbs_to_words<LE>(ow, hsh);
        store_node(out + 2 * h, ow);
    }
}

// nodes[l + 1] = H(nodes[l] x A) for l = 0 .. n_links - 1, one block: the zero tables of
// inf_init (zeroes.rs) as ONE launch per arity instead of 32 launch + copy round trips.
__global__ void __launch_bounds__(32 * COOP_WARPS, 1)
hash_chain_coop_kernel(uint4* nodes, int n_links) {
    using L = Layout<T>;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    CoopBus bus{*reinterpret_cast<CoopSmem*>(coop_raw), (int)(threadIdx.x & 31)};
    const int role = coop_role();
    const uint32_t* tbl = c_tbl;
#pragma unroll 1
    for (int l = 0; l < n_links; l++) {
        uint32_t s[8];
        if (role == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = tbl[L::S0 * 8 + k];
        } else if (role < T) {
            uint32_t wd[8], raw[8];
            load_node_coherent(wd, nodes + 2 * l);
            words_to_limbs<false>(raw, wd);
            absorb<T>(s, raw, role, tbl);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = 0;
        }
        uint32_t hsh[8];
        coop_hash<T>(hsh, s, role, bus, tbl);
        if (role == 0 && bus.lane == 0) {
            uint32_t ow[8];
            limbs_to_words<false>(ow, hsh);
            store_node(nodes + 2 * (l + 1), ow);
            __threadfence_block();
        }
        __syncthreads();
    }
}

// Batched compute_merkle_root_from_path (pallet/src/poll/provider.rs:396-436):
// one path per thread.  At every level the node sits at position idx % A among
// its A-1 siblings (which are stored in order, skipping that position), the A
// values are hashed, idx /= A.  paths: n x depth x (A-1) x 32 bytes.
__global__ void __launch_bounds__(INF_MAX_BLOCK(INF_T), 1)
path_root_kernel(const uint64_t* __restrict__ indices, const uint4* __restrict__ leaves,
                 const uint4* __restrict__ paths, uint32_t depth, uint4* __restrict__ roots,
                 uint64_t n) {
    constexpr int A = T - 1;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t idx = indices[t];
    uint32_t cur[8];
    load_node(cur, leaves + 2 * t);
    const uint4* p = paths + t * (uint64_t)depth * (A - 1) * 2;
#pragma unroll 1
    for (uint32_t l = 0; l < depth; l++) {
        const uint32_t pos = (uint32_t)(idx % A);
        uint32_t iw[A][8];
#pragma unroll
        for (int j = 0; j < A; j++) {
            // sibling slot for position j: j-1 above the node, j below it; the
            // slot read when j == pos is in range and discarded
            int k = (uint32_t)j > pos ? j - 1 : j;
            if (k > A - 2) k = A - 2;
            uint32_t sib[8];
            load_node(sib, p + ((uint64_t)l * (A - 1) + k) * 2);
#pragma unroll
            for (int w = 0; w < 8; w++) iw[j][w] = ((uint32_t)j == pos) ? cur[w] : sib[w];
        }
        hash_words<T, false>(cur, iw, nullptr, c_tbl);
        idx /= A;
    }
    store_node(roots + 2 * t, cur);
}

}  // namespace

// ---- host-side launchers for this width (C++ linkage, used by capi.cu) ------
#define INF_CAT2(a, b) a##b
#define INF_CAT(a, b) INF_CAT2(a, b)

// Resident blocks per SM.  The multiply pipe saturates at 12 warps per SM; more
// resident warps only add instruction-cache misses (the partial-round body is
// 21-35 KB and every warp sits at a different place in it): measured for t=3,
// 5 blocks/SM 135.3, 4 blocks 136.3, 3 blocks 138.8, 2 blocks 135.4 M hash2/s
// (profiles/r01_occupancy_sweep.md).  Widths >= 4 are held at 3 blocks by their
// register count; widths 2 and 3 (94 registers) are held there by reserving
// dynamic shared memory they do not use.  INF_SMEM_PAD overrides (experiments).
static int occupancy_pad() {
    static int pad = -1;
    static bool set_on[64] = {};          // the opt-in above 48 KB is a per-device function attribute
    if (pad < 0) {
        const char* e = getenv("INF_SMEM_PAD");
        pad = e ? atoi(e) : (T <= 3 ? 74 * 1024 : 0);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !set_on[dev]) {
        const int most = 200 * 1024;      // any pad the geometry helpers may ask for
        cudaFuncSetAttribute(hash_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        cudaFuncSetAttribute(hash_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        cudaFuncSetAttribute(tree_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        cudaFuncSetAttribute(path_root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        set_on[dev] = true;
    }
    return pad;
}

static unsigned sm_count() {
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return (unsigned)sms[dev];
}

// Block shape of a launch of n threads.  Wide blocks run one per SM, so a launch takes a whole
// number of waves of sms x block threads, at a rate that depends on the shape (measured,
// profiles/r02_lockstep_experiment.md, relative to one 384-thread block per SM):
//   widths <= 3:  512 threads 0.996, three 128-thread blocks 0.985 (256 x 1 does not saturate the pipe)
//   widths >= 4:  256 threads 0.958, three 128-thread blocks 0.935
// The shape with the smallest estimated time (waves x threads per wave / rate) is taken; launches
// of less than one wave of 384 keep 128-thread blocks, which spread over all SMs (latency regime).
// INF_WIDE_BLOCK=128 gives the round-1 geometry everywhere, 256 / 384 / 512 force one shape.
struct Geom {
    unsigned block;
    int pad;
    bool wide;
};
static Geom pick_geom(uint64_t n) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("INF_WIDE_BLOCK");
        forced = e ? atoi(e) : -1;
        if (forced > INF_MAX_BLOCK(T)) forced = INF_MAX_BLOCK(T);
    }
    const uint64_t sms = sm_count();
    unsigned best = 128;
    if (forced == 256 || forced == 384 || forced == 512) {
        best = (unsigned)forced;
    } else if (forced != 128 && n >= sms * 384) {
        auto est = [&](unsigned b, double rate) {
            const uint64_t per_wave = sms * b;
            return (double)((n + per_wave - 1) / per_wave) * (double)b / rate;
        };
        double tb = est(128, T <= 3 ? 0.985 : 0.935);
        const unsigned shapes[3] = {384, 512, 256};
        const double rates[3] = {1.0, 0.996, T <= 3 ? 0.0 : 0.958};
        for (int i = 0; i < 3; i++) {
            if (shapes[i] > (unsigned)INF_MAX_BLOCK(T) || rates[i] <= 0.0) continue;
            const double t = est(shapes[i], rates[i]);
            if (t < tb) {
                tb = t;
                best = shapes[i];
            }
        }
    }
    if (best == 128) return {INF_BLOCK, occupancy_pad(), false};
    return {best, best == 256 ? 120 * 1024 : 0, true};
}

// A tree level whose grid is a little more than 3 blocks per SM would run its last
// few blocks alone, at single-warp latency, after everything else has finished; if
// 4 or 5 blocks per SM hold the whole grid, let them (94 registers allow 5).
// (Cutting such levels into 64-thread blocks changes nothing — 65 536 parents 509 vs 520 us:
// what quantises a level of a few warps per sub-partition is the number of warps on the
// fullest sub-partition, 4 against an average 3.46, and no block shape changes that.)
static int level_pad(unsigned grid) {
    const int pad = occupancy_pad();
    if (T > 3 || pad != 74 * 1024) return pad;
    const unsigned n = sm_count();
    if (grid <= 3 * n || grid > 5 * n) return pad;
    return grid <= 4 * n ? 55 * 1024 : 44 * 1024;
}

// Levels with at most this many parents go to the warp-cooperative kernel.  Measured crossover
// with the per-thread kernel (profiles/r02_tree_levels.md, tools/level_probe.py): for hash2 the
// cooperative kernel takes 116 / 172 / 203 us at 8 192 / 12 288 / 16 384 parents against 178 us for
// one thread's hash (196 us before the history recurrence, when the threshold was 16 384); for
// hash5 218 us at 8 192 and ~320 at 12 288 against 338.  INF_COOP_MAX overrides, 0 disables.
static uint64_t coop_max() {
    static long long v = -1;
    static bool set_on[64] = {};
    if (v < 0) {
        const char* e = getenv("INF_COOP_MAX");
        v = e ? atoll(e) : (T >= 5 ? 8192 : 12288);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (sizeof(CoopSmem) > 48 * 1024 && dev >= 0 && dev < 64 && !set_on[dev]) {
        cudaFuncSetAttribute(tree_level_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
        cudaFuncSetAttribute(hash_batch_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
        cudaFuncSetAttribute(hash_batch_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
        cudaFuncSetAttribute(hash_chain_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
        set_on[dev] = true;
    }
    return (uint64_t)v;
}

cudaError_t INF_CAT(launch_hash_batch_t, INF_T)(const void* d_in, void* d_out, uint64_t n,
                                                const TagArg& tag, bool le, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n <= coop_max()) {                  // small batch: one hash costs the cooperative critical path
        const unsigned grid = (unsigned)((n + 31) / 32);
        if (le)
            hash_batch_coop_kernel<true><<<grid, 32 * COOP_WARPS, sizeof(CoopSmem), st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
        else
            hash_batch_coop_kernel<false><<<grid, 32 * COOP_WARPS, sizeof(CoopSmem), st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
        return cudaGetLastError();
    }
    const Geom g = pick_geom(n);
    const unsigned grid = (unsigned)((n + g.block - 1) / g.block);
    if (le)
        hash_batch_kernel<true><<<grid, g.block, g.pad, st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
    else
        hash_batch_kernel<false><<<grid, g.block, g.pad, st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
    return cudaGetLastError();
}

// d_nodes[0] given; d_nodes[l + 1] = H(d_nodes[l] x (T-1)) for l < n_links, one launch.
cudaError_t INF_CAT(launch_hash_chain_t, INF_T)(void* d_nodes, int n_links, cudaStream_t st) {
    coop_max();
    hash_chain_coop_kernel<<<1, 32 * COOP_WARPS, sizeof(CoopSmem), st>>>((uint4*)d_nodes, n_links);
    return cudaGetLastError();
}

// Called by inf_init for every width on the context's device, under the init
// lock: also the place where the per-device function attributes (and the
// statics above) are settled, before any concurrent launches can happen.
cudaError_t INF_CAT(upload_table_t, INF_T)(const uint32_t* host_tbl, size_t words) {
    if (words != (size_t)Layout<T>::WORDS) return cudaErrorInvalidValue;
    occupancy_pad();
    coop_max();
    sm_count();
    pick_geom(1);
    return cudaMemcpyToSymbol(c_tbl, host_tbl, words * sizeof(uint32_t));
}

cudaError_t INF_CAT(launch_tree_level_t, INF_T)(const void* d_in, const void* d_prefix, uint64_t shift, uint64_t n_in,
                                                void* d_out, uint64_t n_out,
                                                const uint8_t* zero_be, cudaStream_t st) {
    if (n_out == 0) return cudaSuccess;
    Node32 z;
    memcpy(z.w, zero_be, 32);
    if (n_out <= coop_max()) {
        const unsigned grid = (unsigned)((n_out + 31) / 32);
        return launch_chained(tree_level_coop_kernel, grid, 32 * COOP_WARPS, sizeof(CoopSmem), st, grid <= 128,
                              (const uint4*)d_in,
                              (const uint4*)d_prefix, shift, n_in, (uint4*)d_out, n_out, z);
    }
    const Geom g = pick_geom(n_out);
    if (g.wide)
        return launch_chained(tree_level_kernel, (unsigned)((n_out + g.block - 1) / g.block), g.block, g.pad, st, false,
                              (const uint4*)d_in, (const uint4*)d_prefix, shift, n_in, (uint4*)d_out, n_out, z);
    const unsigned grid = (unsigned)((n_out + INF_BLOCK - 1) / INF_BLOCK);
    return launch_chained(tree_level_kernel, grid, INF_BLOCK, level_pad(grid), st, false, (const uint4*)d_in,
                          (const uint4*)d_prefix, shift, n_in, (uint4*)d_out, n_out, z);
}

cudaError_t INF_CAT(launch_path_root_t, INF_T)(const void* d_idx, const void* d_leaves,
                                               const void* d_paths, uint32_t depth, void* d_roots,
                                               uint64_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + INF_BLOCK - 1) / INF_BLOCK);
    path_root_kernel<<<grid, INF_BLOCK, occupancy_pad(), st>>>((const uint64_t*)d_idx, (const uint4*)d_leaves,
                                                 (const uint4*)d_paths, depth, (uint4*)d_roots, n);
    return cudaGetLastError();
}

}  // namespace inf
