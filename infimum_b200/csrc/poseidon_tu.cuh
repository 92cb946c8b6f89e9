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
__global__ void __launch_bounds__(INF_BLOCK, INF_MIN_BLOCKS(INF_T))
hash_batch_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n, TagArg tag) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    uint32_t iw[T - 1][8], ow[8];
    const uint4* p = in + idx * (uint64_t)(2 * (T - 1));
#pragma unroll
    for (int i = 0; i < T - 1; i++) load_node(iw[i], p + 2 * i);
    hash_words<T, LE>(ow, iw, tag.has ? tag.w : nullptr, c_tbl);
    store_node(out + 2 * idx, ow);
}

// One level of an arity-(T-1) Merkle tree over big-endian 32-byte nodes.
// The level's logical node array is `shift` copies of the zero value followed
// by in[0..n_in): shift = 1 at level 0 of the registration tree, whose leaf 0
// is the blank state leaf = zeroes[0] (state.rs:48-52), else 0.
// out[i] = H(node[A*i], ..., node[A*i+A-1]); nodes past the end are the
// level's zero value (PollStateTree::merge's right padding, state.rs:262-266).
__global__ void __launch_bounds__(INF_BLOCK, INF_MIN_BLOCKS(INF_T))
tree_level_kernel(const uint4* __restrict__ in, uint64_t shift, uint64_t n_in,
                  uint4* __restrict__ out, uint64_t n_out, Node32 zero) {
    constexpr int A = T - 1;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out) return;
    uint32_t iw[A][8], ow[8];
    const uint64_t first = idx * A;
#pragma unroll
    for (int i = 0; i < A; i++) {
        const uint64_t j = first + i;
        if (j >= shift && j - shift < n_in) {
            load_node(iw[i], in + 2 * (j - shift));
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) iw[i][k] = zero.w[k];
        }
    }
    hash_words<T, false>(ow, iw, nullptr, c_tbl);
    store_node(out + 2 * idx, ow);
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative tree level, for the levels near the root.
//
// A level with fewer nodes than resident threads costs one single-thread hash
// latency (~190 us: a lone warp needs 810 cycles per dependent multiplication,
// profiles/r01_imad_microbench.md) whatever its size.  Here T warps share 32
// hashes: warp w owns state element w of all 32 (lane = hash), the warps run on
// different sub-partitions, and the state is staged through shared memory:
//   full round     every warp: own S-box, publish it, barrier, own MDS row
//                  (1 S-box + 1 row instead of T + T on the critical path)
//   partial pair   warp 0 carries the chain S-box, row A, S-box, row B; warps
//                  1..T-1 take their  s_w += w_A x_a + w_B x_b  off it
// Same tables, same arithmetic, so results are bit-identical to the per-thread
// kernel; the critical path drops ~1.5x (t=3) / ~2x (t=6).
// Shared layout [element][limb][lane]: consecutive lanes hit consecutive banks.
struct CoopSmem {
    uint32_t x[2][T][8][32];      // S-box outputs of a full round, double buffered
    uint32_t s[T][8][32];         // s[1..T-1] as of the start of the current pair
    uint32_t xa[8][32], xb[8][32];
};

__device__ __forceinline__ void sm_put(uint32_t (*dst)[32], const uint32_t (&v)[8], int lane) {
#pragma unroll
    for (int k = 0; k < 8; k++) dst[k][lane] = v[k];
}
__device__ __forceinline__ void sm_get(uint32_t* v, const uint32_t (*src)[32], int lane) {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = src[k][lane];
}

__global__ void __launch_bounds__(32 * T, 1)
tree_level_coop_kernel(const uint4* __restrict__ in, uint64_t shift, uint64_t n_in,
                       uint4* __restrict__ out, uint64_t n_out, Node32 zero) {
    using L = Layout<T>;
    constexpr int A = T - 1;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    CoopSmem& sm = *reinterpret_cast<CoopSmem*>(coop_raw);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t h = (uint64_t)blockIdx.x * 32 + lane;
    const bool live = h < n_out;          // dead lanes run along on zeros (barriers are block wide)
    const uint32_t* tbl = c_tbl;

    // ---- absorb: warp w >= 1 takes child w-1 of its hash ---------------------
    uint32_t s[8];
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = tbl[L::S0 * 8 + k];
    } else {
        uint32_t wd[8], raw[8];
        const uint64_t j = h * A + (w - 1);
        if (live && j >= shift && j - shift < n_in) {
            load_node(wd, in + 2 * (j - shift));
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) wd[k] = zero.w[k];
        }
        words_to_limbs<false>(raw, wd);
        absorb<T>(s, raw, w, tbl);
    }

    uint32_t xs[T + 1][8];
    // ---- first half: rounds 0..3 -----------------------------------------------
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const uint32_t* m = tbl + (r < 3 ? L::FULL_M : L::PRE_M) * 8;
        const uint32_t* v = tbl + (r < 3 ? L::FULL_V + r * T : L::PRE_V) * 8;
        uint32_t x[8];
        sbox(x, s);
        sm_put(sm.x[r & 1][w], x, lane);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < T; i++) sm_get(xs[i], sm.x[r & 1][i], lane);
        dot<T, 8>(s, &xs[0][0], m + w * T * 8, v + w * 8);
    }

    // ---- partial rounds ------------------------------------------------------------
    if (w > 0) sm_put(sm.s[w], s, lane);
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 1; i < T; i++) sm_get(xs[i], sm.s[i], lane);
    }
#pragma unroll 1
    for (int j = 0; j < L::N_PAIRS; j++) {
        const uint32_t* pt = tbl + (L::PART + j * L::PAIR_STRIDE) * 8;
        if (w == 0) {
            uint32_t n[8];
            sbox(xs[0], s);                                   // round A
            sm_put(sm.xa, xs[0], lane);
#pragma unroll
            for (int k = 0; k < 8; k++) xs[T][k] = xs[0][k];
            __syncthreads();                                  // (A) x_a published
            dot<T, 8>(n, &xs[0][0], pt, pt + T * 8);
            sbox(xs[0], n);                                   // round B
            sm_put(sm.xb, xs[0], lane);
            __syncthreads();                                  // (B) x_b published
            dot<T + 1, 8>(s, &xs[0][0], pt + (T + 1) * 8, pt + (2 * T + 2) * 8);
            __syncthreads();                                  // (C) s[1..] updated by the others
#pragma unroll
            for (int i = 1; i < T; i++) sm_get(xs[i], sm.s[i], lane);
        } else {
            uint32_t ab[2][8], d[8];
            __syncthreads();                                  // (A)
            sm_get(ab[0], sm.xa, lane);
            __syncthreads();                                  // (B)
            sm_get(ab[1], sm.xb, lane);
            dot<2, 8, false>(d, &ab[0][0], pt + (2 * T + 3 + 2 * (w - 1)) * 8, nullptr);
            add8(s, s, d);
            csub2p(s);
            sm_put(sm.s[w], s, lane);
            __syncthreads();                                  // (C)
        }
    }
#pragma unroll 1
    for (int j = 0; j < L::N_SINGLES; j++) {                  // the odd round out
        const uint32_t* pt = tbl + (L::SINGLES + j * L::SINGLE_STRIDE) * 8;
        if (w == 0) {
            uint32_t n[8];
            sbox(xs[0], s);
            sm_put(sm.xa, xs[0], lane);
            __syncthreads();                                  // (A)
            dot<T, 8>(n, &xs[0][0], pt, pt + (2 * T - 1) * 8);
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = n[k];
            __syncthreads();                                  // (C)
#pragma unroll
            for (int i = 1; i < T; i++) sm_get(xs[i], sm.s[i], lane);
        } else {
            uint32_t xa[8], d[8];
            __syncthreads();                                  // (A)
            sm_get(xa, sm.xa, lane);
            mont_mul(d, xa, pt + (T + w - 1) * 8);
            add8(s, s, d);
            csub2p(s);
            sm_put(sm.s[w], s, lane);
            __syncthreads();                                  // (C)
        }
    }
    if (w > 0) {                                              // remaining constants of the first tail round
        add8(s, s, tbl + (L::LAST_D + w - 1) * 8);
        csub2p(s);
    }

    // ---- second half: 3 full rounds, then the output row -------------------------
#pragma unroll 1
    for (int r = 0; r < 3; r++) {
        const uint32_t* m = tbl + L::FULL_M * 8;
        const uint32_t* v = tbl + (L::TAIL_V + r * T) * 8;
        uint32_t x[8];
        sbox(x, s);
        sm_put(sm.x[r & 1][w], x, lane);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < T; i++) sm_get(xs[i], sm.x[r & 1][i], lane);
        dot<T, 8>(s, &xs[0][0], m + w * T * 8, v + w * 8);
    }
    {
        uint32_t x[8];
        sbox(x, s);
        sm_put(sm.x[1][w], x, lane);                          // rounds 0..2 left buffer 0 last
        __syncthreads();
    }
    if (w == 0 && live) {
#pragma unroll
        for (int i = 0; i < T; i++) sm_get(xs[i], sm.x[1][i], lane);
        uint32_t hsh[8], ow[8];
        dot<T, 8>(hsh, &xs[0][0], tbl + L::OUT_ROW * 8, nullptr);
        csub_p_exact(hsh);
        csub_p_exact(hsh);
        limbs_to_words<false>(ow, hsh);
        store_node(out + 2 * h, ow);
    }
}

// Batched compute_merkle_root_from_path (pallet/src/poll/provider.rs:396-436):
// one path per thread.  At every level the node sits at position idx % A among
// its A-1 siblings (which are stored in order, skipping that position), the A
// values are hashed, idx /= A.  paths: n x depth x (A-1) x 32 bytes.
__global__ void __launch_bounds__(INF_BLOCK, INF_MIN_BLOCKS(INF_T))
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
    if (pad < 0) {
        const char* e = getenv("INF_SMEM_PAD");
        pad = e ? atoi(e) : (T <= 3 ? 74 * 1024 : 0);
        if (pad > 48 * 1024) {
            cudaFuncSetAttribute(hash_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
            cudaFuncSetAttribute(hash_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
            cudaFuncSetAttribute(tree_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
            cudaFuncSetAttribute(path_root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
        }
    }
    return pad;
}

cudaError_t INF_CAT(upload_table_t, INF_T)(const uint32_t* host_tbl, size_t words) {
    if (words != (size_t)Layout<T>::WORDS) return cudaErrorInvalidValue;
    return cudaMemcpyToSymbol(c_tbl, host_tbl, words * sizeof(uint32_t));
}

cudaError_t INF_CAT(launch_hash_batch_t, INF_T)(const void* d_in, void* d_out, uint64_t n,
                                                const TagArg& tag, bool le, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + INF_BLOCK - 1) / INF_BLOCK);
    const int pad = occupancy_pad();
    if (le)
        hash_batch_kernel<true><<<grid, INF_BLOCK, pad, st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
    else
        hash_batch_kernel<false><<<grid, INF_BLOCK, pad, st>>>((const uint4*)d_in, (uint4*)d_out, n, tag);
    return cudaGetLastError();
}

// Levels with at most this many parents go to the warp-cooperative kernel (they
// are latency-bound either way; INF_COOP_MAX overrides, 0 disables).
static uint64_t coop_max() {
    static long long v = -1;
    if (v < 0) {
        const char* e = getenv("INF_COOP_MAX");
        v = e ? atoll(e) : 8192;
        cudaFuncSetAttribute(tree_level_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
    }
    return (uint64_t)v;
}

cudaError_t INF_CAT(launch_tree_level_t, INF_T)(const void* d_in, uint64_t shift, uint64_t n_in,
                                                void* d_out, uint64_t n_out,
                                                const uint8_t* zero_be, cudaStream_t st) {
    if (n_out == 0) return cudaSuccess;
    Node32 z;
    memcpy(z.w, zero_be, 32);
    if (n_out <= coop_max()) {
        const unsigned grid = (unsigned)((n_out + 31) / 32);
        tree_level_coop_kernel<<<grid, 32 * T, sizeof(CoopSmem), st>>>((const uint4*)d_in, shift, n_in,
                                                                       (uint4*)d_out, n_out, z);
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((n_out + INF_BLOCK - 1) / INF_BLOCK);
    tree_level_kernel<<<grid, INF_BLOCK, occupancy_pad(), st>>>((const uint4*)d_in, shift, n_in, (uint4*)d_out, n_out, z);
    return cudaGetLastError();
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
