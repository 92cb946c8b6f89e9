// Launch geometry and the launcher prototypes shared by the per-width
// translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

// Block shapes.  Register caps, measured on B200 (tools/variant_probe.py,
// profiles/r01_occupancy_sweep.md): widths 2 and 3 fit 94 registers; from width 4 on, a cap
// of 128 registers spills and 168 is 5-21 % faster — the multiply pipe is the bottleneck,
// 12 warps per SM already saturate it, spills only add traffic.  The caps are expressed as
// launch bounds of ONE block of 512 threads (65536 / 512 = 128 registers, widths <= 3) or 384
// threads (170 -> 168 registers, widths >= 4) per SM, which is also the shape large launches
// use (INF_WIDE_BLOCK): the twelve warps of a 384-thread block start together and stay
// within a few instructions of each other, so they share the instruction stream (the round
// loops are 37-104 KB against a ~32 KB instruction cache), where three independent 128-thread
// blocks drift apart — measured hash5 +13.5 %, hash3 +6.5 %, hash4 +2.4 %, hash2 +0.8 %
// (profiles/r02_lockstep_experiment.md).  Small launches (tree levels of fewer than a few
// waves) keep 128-thread blocks, which spread over all SMs.
#ifndef INF_BLOCK
#define INF_BLOCK 128
#endif
#ifndef INF_MAX_BLOCK
#define INF_MAX_BLOCK(T) ((T) >= 4 ? 384 : 512)
#endif
#ifndef INF_WIDE_BLOCK
#define INF_WIDE_BLOCK 384
#endif

namespace inf {

// Programmatic dependent launch for the level-after-level kernels of a tree: the next
// level's grid is set up while the previous one drains (its blocks wait in
// griddep_wait() before they touch the previous level's output), which takes the
// launch gap (~3 us) off a latency-bound level.  Only for grids of at most one block
// per SM (`early`): blocks of a larger grid that become resident early sit wherever a
// slot happened to be free, and the level then runs on a lopsided placement — measured
// on B200, chaining the wide levels made a 2^20-leaf tree 0.44 ms SLOWER, chaining the
// levels near the root only makes a 2^10-leaf tree 0.03 ms faster
// (profiles/r02_tree_levels.md).  INF_NO_PDL=1 disables it.
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                  cudaStream_t st, bool early, Args... args) {
    static const bool pdl_on = !(getenv("INF_NO_PDL") && atoi(getenv("INF_NO_PDL")));
    const bool pdl = pdl_on && early;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

struct TagArg {          // domain tag in wire order (poseidon.rs:110-120); has == 0 -> tag 0
    uint32_t w[8];
    int has;
};

#define INF_DECLARE_WIDTH(N)                                                                        \
    cudaError_t upload_table_t##N(const uint32_t* host_tbl, size_t words);                          \
    cudaError_t launch_hash_batch_t##N(const void* d_in, void* d_out, uint64_t n, const TagArg& tag, \
                                       bool le, cudaStream_t st);                                   \
    cudaError_t launch_tree_level_t##N(const void* d_in, const void* d_prefix, uint64_t shift,      \
                                       uint64_t n_in, void* d_out, uint64_t n_out,                  \
                                       const uint8_t* zero_be, cudaStream_t st);                    \
    cudaError_t launch_path_root_t##N(const void* d_idx, const void* d_leaves, const void* d_paths,  \
                                      uint32_t depth, void* d_roots, uint64_t n, cudaStream_t st);  \
    cudaError_t launch_hash_chain_t##N(void* d_nodes, int n_links, cudaStream_t st);
INF_DECLARE_WIDTH(2)
INF_DECLARE_WIDTH(3)
INF_DECLARE_WIDTH(4)
INF_DECLARE_WIDTH(5)
INF_DECLARE_WIDTH(6)
INF_DECLARE_WIDTH(7)
INF_DECLARE_WIDTH(8)
#undef INF_DECLARE_WIDTH

// Leaf hashing (leaves.cu)
cudaError_t upload_leaf_tables(const uint32_t* t5, size_t w5, const uint32_t* t6, size_t w6);
cudaError_t launch_interaction_leaves(const void* d_pk, const void* d_data, void* d_out, uint64_t n,
                                      cudaStream_t st);
cudaError_t launch_registration_leaves(const void* d_pk, const void* d_ts, void* d_out, uint64_t n,
                                       cudaStream_t st);

// Sibling-path gather over retained tree levels (tree_paths.cu)
cudaError_t launch_gather_paths(const void* const* d_level_ptrs /* device array [depth] */,
                                const uint64_t* d_level_counts /* device array [depth] */,
                                const void* d_zero_nodes /* device, depth x 32 B */, uint32_t arity,
                                uint32_t depth, const void* d_indices, uint64_t n, void* d_out,
                                cudaStream_t st);

// Generic dense kernel (any width 2..13), dense_generic.cu
cudaError_t launch_hash_dense_params(int t, int full_rounds, int rp, uint64_t alpha, const uint32_t* d_tbl,
                                     const void* d_in, void* d_out, uint64_t n, const TagArg& tag, bool le,
                                     cudaStream_t st);
cudaError_t launch_hash_dense(int t, const uint32_t* d_tbl, const void* d_in, void* d_out,
                              uint64_t n, const TagArg& tag, bool le, cudaStream_t st);

}  // namespace inf
