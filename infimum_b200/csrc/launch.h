// Launch geometry and the launcher prototypes shared by the per-width
// translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

// 128 threads x 4 resident blocks = 16 warps/SM at <= 128 registers/thread.
#ifndef INF_BLOCK
#define INF_BLOCK 128
#endif
#ifndef INF_MIN_BLOCKS
#define INF_MIN_BLOCKS 4
#endif

namespace inf {

struct TagArg {          // domain tag in wire order (poseidon.rs:110-120); has == 0 -> tag 0
    uint32_t w[8];
    int has;
};

#define INF_DECLARE_WIDTH(N)                                                                        \
    cudaError_t upload_table_t##N(const uint32_t* host_tbl, size_t words);                          \
    cudaError_t launch_hash_batch_t##N(const void* d_in, void* d_out, uint64_t n, const TagArg& tag, \
                                       bool le, cudaStream_t st);                                   \
    cudaError_t launch_tree_level_t##N(const void* d_in, uint64_t shift, uint64_t n_in, void* d_out, \
                                       uint64_t n_out, const uint8_t* zero_be, cudaStream_t st);
INF_DECLARE_WIDTH(2)
INF_DECLARE_WIDTH(3)
INF_DECLARE_WIDTH(4)
INF_DECLARE_WIDTH(5)
INF_DECLARE_WIDTH(6)
INF_DECLARE_WIDTH(7)
INF_DECLARE_WIDTH(8)
#undef INF_DECLARE_WIDTH

// Generic dense kernel (any width 2..13), dense_generic.cu
cudaError_t launch_hash_dense(int t, const uint32_t* d_tbl, const void* d_in, void* d_out,
                              uint64_t n, const TagArg& tag, bool le, cudaStream_t st);

}  // namespace inf
