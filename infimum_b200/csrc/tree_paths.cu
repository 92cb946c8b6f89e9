// Merkle sibling paths in bulk from retained tree levels (SURVEY.md 8f rank 4:
// what the off-chain coordinator needs as msgSubrootPathElements /
// currentStateLeavesPathElements, circuits/process-messages.circom:57,85, in
// the layout compute_merkle_root_from_path consumes, provider.rs:396-436).
//
// Pure gather: byte movement, HBM-bound.  One thread per (path, level,
// sibling) moves one 32-byte node with two 128-bit accesses; consecutive
// threads write consecutive 32-byte slots, so stores are fully coalesced and
// the reads of one level's siblings of one path are contiguous.
#include <cuda_runtime.h>

#include "launch.h"

namespace inf {
namespace {

__global__ void __launch_bounds__(256)
gather_paths_kernel(const uint4* const* __restrict__ levels, const uint64_t* __restrict__ counts,
                    const uint4* __restrict__ zeros, uint32_t arity, uint32_t depth,
                    const uint64_t* __restrict__ indices, uint64_t n, uint4* __restrict__ out) {
    const uint64_t per_path = (uint64_t)depth * (arity - 1);
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * per_path) return;
    const uint64_t t = g / per_path;
    const uint32_t r = (uint32_t)(g - t * per_path);
    const uint32_t l = r / (arity - 1), k = r - l * (arity - 1);
    uint64_t idx = indices[t];
    for (uint32_t i = 0; i < l; i++) idx /= arity;            // node index at level l
    const uint32_t pos = (uint32_t)(idx % arity);
    const uint64_t sib = idx - pos + (k >= pos ? k + 1 : k);   // k-th sibling, skipping the node itself
    uint4 a, b;
    if (sib < counts[l]) {
        const uint4* p = levels[l] + 2 * sib;
        a = __ldg(p);
        b = __ldg(p + 1);
    } else {                                                   // right of the last node: the level's zero value
        a = zeros[2 * l];
        b = zeros[2 * l + 1];
    }
    out[2 * g] = a;
    out[2 * g + 1] = b;
}

}  // namespace

cudaError_t launch_gather_paths(const void* const* d_level_ptrs, const uint64_t* d_level_counts,
                                const void* d_zero_nodes, uint32_t arity, uint32_t depth,
                                const void* d_indices, uint64_t n, void* d_out, cudaStream_t st) {
    const uint64_t total = n * (uint64_t)depth * (arity - 1);
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 255) / 256);
    gather_paths_kernel<<<grid, 256, 0, st>>>((const uint4* const*)d_level_ptrs, d_level_counts,
                                              (const uint4*)d_zero_nodes, arity, depth,
                                              (const uint64_t*)d_indices, n, (uint4*)d_out);
    return cudaGetLastError();
}

}  // namespace inf
