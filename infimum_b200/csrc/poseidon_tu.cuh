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
//   partial pair   warp 0 carries only the chain  S-box, x_a*m00, S-box, x_b*m00':
//                  the products of the two rows that do not involve the fresh
//                  S-box output (s_w*rowA[w], s_w*rowB[w], c_B*x_a) and the updates
//                  s_w += w_A x_a + w_B x_b are formed by warps 1..T-1 meanwhile
//                  and only added in (separately reduced, so the sums agree mod p)
// Same tables, same field values, so results are bit-identical to the per-thread
// kernel; per pair the critical path is 944 multiply-pipe instructions instead
// of 1 664 (t=3) / 2 648 (t=6).
// Shared layout [element][limb][lane]: consecutive lanes hit consecutive banks.
struct CoopSmem {
    uint32_t x[2][T][8][32];      // S-box outputs of a full round, double buffered
    // partial rounds, everything double buffered by pair parity:
    uint32_t xa[2][8][32], xb[2][8][32];   // S-box outputs of rounds A and B (warp 0)
    uint32_t pa[2][T][8][32];              // s_w * rowA[w] / R   (warp w >= 1)
    uint32_t pb[2][T][8][32];              // s_w * rowB[w] / R   (warp w >= 1)
    uint32_t cx[2][8][32];                 // c_B * x_a / R       (warp 1)
    uint32_t s[T][8][32];                  // s[1..T-1] for the odd round out
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
    // Pair j, buffers of parity j & 1.  Two barriers per pair:
    //   before (A): warp 0 publishes x_a; warp w >= 1 publishes s_w*rowA[w] and s_w*rowB[w]
    //   before (B): warp 0 forms n = x_a*m00 + k_A + sum pa, publishes x_b = n^5;
    //               warp 1 publishes c_B*x_a
    //   after  (B): warp 0 forms s_0 = x_b*m00' + k_B + sum pb + cx;
    //               warp w >= 1 updates s_w += w_A x_a + w_B x_b  and runs ahead into pair j+1
    // A buffer of parity p is rewritten in pair j+2 only by warps that have passed
    // barrier (B) of pair j+1, which its readers of pair j reach after reading it.
#pragma unroll 1
    for (int j = 0; j < L::N_PAIRS; j++) {
        const uint32_t* pt = tbl + (L::PART + j * L::PAIR_STRIDE) * 8;
        const int p = j & 1;
        if (w == 0) {
            uint32_t xa[8], xb[8], n[8], t[8];
            sbox(xa, s);                                              // round A
            sm_put(sm.xa[p], xa, lane);
            __syncthreads();                                          // (A)
            mont_mul_add(n, xa, pt, pt + T * 8);                      // x_a*m00 + k_A
#pragma unroll
            for (int i = 1; i < T; i++) {
                sm_get(t, sm.pa[p][i], lane);
                add8(n, n, t);
                csub2p(n);
            }
            sbox(xb, n);                                              // round B
            sm_put(sm.xb[p], xb, lane);
            __syncthreads();                                          // (B)
            mont_mul_add(s, xb, pt + (T + 1) * 8, pt + (2 * T + 2) * 8);   // x_b*m00' + k_B
#pragma unroll
            for (int i = 1; i < T; i++) {
                sm_get(t, sm.pb[p][i], lane);
                add8(s, s, t);
                csub2p(s);
            }
            sm_get(t, sm.cx[p], lane);
            add8(s, s, t);
            csub2p(s);
        } else {
            uint32_t t[8], ab[2][8];
            mont_mul(t, s, pt + w * 8);                               // s_w * rowA[w]
            sm_put(sm.pa[p][w], t, lane);
            mont_mul(t, s, pt + (T + 1 + w) * 8);                     // s_w * rowB[w]
            sm_put(sm.pb[p][w], t, lane);
            __syncthreads();                                          // (A)
            sm_get(ab[0], sm.xa[p], lane);
            if (w == 1) {
                mont_mul(t, ab[0], pt + (2 * T + 1) * 8);             // c_B * x_a
                sm_put(sm.cx[p], t, lane);
            }
            __syncthreads();                                          // (B)
            sm_get(ab[1], sm.xb[p], lane);
            dot<2, 8, false>(t, &ab[0][0], pt + (2 * T + 3 + 2 * (w - 1)) * 8, nullptr);
            add8(s, s, t);
            csub2p(s);
        }
    }
    if (L::N_SINGLES > 0) {                                           // the odd round out, in the plain form
        if (w > 0) sm_put(sm.s[w], s, lane);
        __syncthreads();
        const uint32_t* pt = tbl + L::SINGLES * 8;
        if (w == 0) {
            uint32_t n[8];
#pragma unroll
            for (int i = 1; i < T; i++) sm_get(xs[i], sm.s[i], lane);
            sbox(xs[0], s);
            sm_put(sm.xa[0], xs[0], lane);
            __syncthreads();
            dot<T, 8>(n, &xs[0][0], pt, pt + (2 * T - 1) * 8);
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = n[k];
        } else {
            uint32_t xa[8], d[8];
            __syncthreads();
            sm_get(xa, sm.xa[0], lane);
            mont_mul(d, xa, pt + (T + w - 1) * 8);
            add8(s, s, d);
            csub2p(s);
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
    static bool set_on[64] = {};          // the opt-in above 48 KB is a per-device function attribute
    if (pad < 0) {
        const char* e = getenv("INF_SMEM_PAD");
        pad = e ? atoi(e) : (T <= 3 ? 74 * 1024 : 0);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (pad > 48 * 1024 && dev >= 0 && dev < 64 && !set_on[dev]) {
        cudaFuncSetAttribute(hash_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
        cudaFuncSetAttribute(hash_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
        cudaFuncSetAttribute(tree_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
        cudaFuncSetAttribute(path_root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
        set_on[dev] = true;
    }
    return pad;
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
// are latency-bound either way; measured best 16 384..32 768 on B200, tools/tree_probe.py;
// INF_COOP_MAX overrides, 0 disables).
static uint64_t coop_max() {
    static long long v = -1;
    static bool set_on[64] = {};
    if (v < 0) {
        const char* e = getenv("INF_COOP_MAX");
        v = e ? atoll(e) : 16384;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (sizeof(CoopSmem) > 48 * 1024 && dev >= 0 && dev < 64 && !set_on[dev]) {
        cudaFuncSetAttribute(tree_level_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem));
        set_on[dev] = true;
    }
    return (uint64_t)v;
}

// Called by inf_init for every width on the context's device, under the init
// lock: also the place where the per-device function attributes (and the
// statics above) are settled, before any concurrent launches can happen.
cudaError_t INF_CAT(upload_table_t, INF_T)(const uint32_t* host_tbl, size_t words) {
    if (words != (size_t)Layout<T>::WORDS) return cudaErrorInvalidValue;
    occupancy_pad();
    coop_max();
    return cudaMemcpyToSymbol(c_tbl, host_tbl, words * sizeof(uint32_t));
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
