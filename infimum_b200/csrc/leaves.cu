// Leaf hashing: the step immediately before the trees.
//
//   registration leaf  = hash4(pk.x, pk.y, 1, timestamp)
//        PollProvider::register_participant, pallet/src/poll/provider.rs:218-241
//   interaction leaf   = hash4(hash5(d[0..5]), hash5(d[5..10]), pk.x, pk.y)
//        PollProvider::consume_interaction,  pallet/src/poll/provider.rs:243-287
//        (mirrors MessageHasher, circuits/utils/hashers.circom:39-78)
//
// One participant / one message per thread; the three hashes of a message are
// fused in one kernel (2 x t=6 then t=5), intermediates never leave registers.
// This translation unit owns its own copies of the t=5 and t=6 tables (22 KB +
// 27 KB of the 64 KB constant bank).
#include <cuda_runtime.h>

#include "launch.h"
#include "poseidon.cuh"

namespace inf {
namespace {

__constant__ uint32_t c_tbl5[Layout<5>::THREAD_WORDS];   // the per-thread kernels' prefix of the table
__constant__ uint32_t c_tbl6[Layout<6>::THREAD_WORDS];

__device__ __forceinline__ void load_node(uint32_t (&w)[8], const uint4* p) {
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void store_node(uint4* p, const uint32_t (&w)[8]) {
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Launch bound: one block of 384 threads per SM (cap 168 registers, no spills), which is also
// the shape large launches use: twelve warps that start together share the instruction stream
// (27.2 M messages/s), three independent 128-thread blocks drift apart and thrash the
// instruction cache (22.2 M/s); block sizes that are not a multiple of 128 leave sub-partitions
// a warp short (320: 20.3, 352: 22.3 M/s).  A launch whose last wave of 384-thread blocks would be
// mostly empty takes one 256-thread block per SM instead (26.9 M/s).  INF_LEAF_BLOCK forces a shape
// (profiles/r02_lockstep_experiment.md).
#ifndef INF_LEAF_MAX_BLOCK
#define INF_LEAF_MAX_BLOCK 384
#endif
static int leaf_sms() {
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}
static unsigned leaf_block(uint64_t n) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("INF_LEAF_BLOCK");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 128 && forced <= INF_LEAF_MAX_BLOCK && !(forced & 31)) return (unsigned)forced;
    const uint64_t sms = (uint64_t)leaf_sms();
    if (n < sms * 256) return 128;                   // latency regime: spread over all SMs
    // estimated time = waves x threads per wave / measured rate (27.2, 26.9, 21.1 M messages/s)
    auto est = [&](unsigned b, double rate) {
        const uint64_t per_wave = sms * b;
        return (double)((n + per_wave - 1) / per_wave) * (double)b / rate;
    };
    const double t384 = 0.99 * est(384, 27.2), t256 = est(256, 26.9), t128 = est(128, 21.1);   // ties go to 384
    return t384 <= t256 && t384 <= t128 ? 384 : (t256 <= t128 ? 256 : 128);
}

// pk: n x (x, y) 32-byte big-endian; data: n x 10 x 32 bytes; out: n x 32 bytes
__global__ void __launch_bounds__(INF_LEAF_MAX_BLOCK, 1)
interaction_leaf_kernel(const uint4* __restrict__ pk, const uint4* __restrict__ data,
                        uint4* __restrict__ out, uint64_t n) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    uint32_t in4[4][8];
    {
        uint32_t half[5][8];
        const uint4* d = data + idx * 20;
#pragma unroll
        for (int i = 0; i < 5; i++) load_node(half[i], d + 2 * i);
        hash_words<6, false>(in4[0], half, nullptr, c_tbl6);          // hash5(d[0..5])
#pragma unroll
        for (int i = 0; i < 5; i++) load_node(half[i], d + 10 + 2 * i);
        hash_words<6, false>(in4[1], half, nullptr, c_tbl6);          // hash5(d[5..10])
    }
    load_node(in4[2], pk + idx * 4);
    load_node(in4[3], pk + idx * 4 + 2);
    uint32_t leaf[8];
    hash_words<5, false>(leaf, in4, nullptr, c_tbl5);                 // hash4(left, right, pk.x, pk.y)
    store_node(out + 2 * idx, leaf);
}

// pk: n x (x, y); timestamps: n x u64 (block numbers); out: n x 32 bytes
__global__ void __launch_bounds__(INF_LEAF_MAX_BLOCK, 1)
registration_leaf_kernel(const uint4* __restrict__ pk, const unsigned long long* __restrict__ ts,
                         uint4* __restrict__ out, uint64_t n) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    uint32_t in4[4][8];
    load_node(in4[0], pk + idx * 4);
    load_node(in4[1], pk + idx * 4 + 2);
    const unsigned long long t = ts[idx];
#pragma unroll
    for (int k = 0; k < 8; k++) in4[2][k] = in4[3][k] = 0;
    in4[2][7] = 0x01000000u;                                           // Fr::from(1), big-endian bytes
    in4[3][6] = bswap32((uint32_t)(t >> 32));                          // Fr::from(timestamp: u64)
    in4[3][7] = bswap32((uint32_t)t);
    uint32_t leaf[8];
    hash_words<5, false>(leaf, in4, nullptr, c_tbl5);
    store_node(out + 2 * idx, leaf);
}

}  // namespace

cudaError_t upload_leaf_tables(const uint32_t* t5, size_t w5, const uint32_t* t6, size_t w6) {
    if (w5 != (size_t)Layout<5>::WORDS || w6 != (size_t)Layout<6>::WORDS) return cudaErrorInvalidValue;
    leaf_block(1);          // settle the lazily read overrides and the per-device SM count here, under inf_init's lock
    leaf_sms();
    cudaError_t e = cudaMemcpyToSymbol(c_tbl5, t5, (size_t)Layout<5>::THREAD_WORDS * 4);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_tbl6, t6, (size_t)Layout<6>::THREAD_WORDS * 4);
}

cudaError_t launch_interaction_leaves(const void* d_pk, const void* d_data, void* d_out, uint64_t n,
                                      cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned block = leaf_block(n);
    const unsigned grid = (unsigned)((n + block - 1) / block);
    interaction_leaf_kernel<<<grid, block, 0, st>>>((const uint4*)d_pk, (const uint4*)d_data, (uint4*)d_out, n);
    return cudaGetLastError();
}

cudaError_t launch_registration_leaves(const void* d_pk, const void* d_ts, void* d_out, uint64_t n,
                                       cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned block = leaf_block(n);
    const unsigned grid = (unsigned)((n + block - 1) / block);
    registration_leaf_kernel<<<grid, block, 0, st>>>((const uint4*)d_pk, (const unsigned long long*)d_ts,
                                                       (uint4*)d_out, n);
    return cudaGetLastError();
}

}  // namespace inf
