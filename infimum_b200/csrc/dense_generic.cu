// Generic dense Poseidon kernel: any width 2..13, the reference's schedule
// taken literally (pallet/src/hash/poseidon.rs:184-203: ARK, x^5, dense MDS,
// every round), table of Montgomery-form (ark, mds) in global memory, state in
// local memory with run-time indexing.  It is slow by design and serves two
// purposes: widths 9..13, which no caller on the hot path uses but
// `Poseidon::new_circom` accepts (poseidon.rs:315), and an on-device
// cross-check of the optimised kernels that shares none of their tables.
#include <cuda_runtime.h>

#include "launch.h"
#include "poseidon.cuh"

namespace inf {
namespace {

__device__ __forceinline__ void add_mod2p(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b) {
    add8(r, a, b);   // a, b < 2p + eps  =>  sum < 4p + eps < 2^256
    csub2p(r);
}

template <bool LE>
__global__ void __launch_bounds__(128)
hash_dense_kernel(int t, int rp, const uint32_t* __restrict__ tbl, const uint32_t* __restrict__ in,
                  uint32_t* __restrict__ out, uint64_t n, TagArg tag) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t* ark = tbl;
    const uint32_t* mds = tbl + (size_t)(8 + rp) * t * 8;
    // R^2 mod p, to enter Montgomery form
    const uint32_t r2[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                            0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    uint32_t s[13][8], nx[13][8];
    for (int i = 0; i < t; i++) {
        uint32_t w[8], raw[8];
        if (i == 0) {
            for (int k = 0; k < 8; k++) w[k] = tag.has ? tag.w[k] : 0u;
        } else {
            const uint32_t* p = in + (idx * (uint64_t)(t - 1) + (i - 1)) * 8;
            for (int k = 0; k < 8; k++) w[k] = p[k];
        }
        words_to_limbs<LE>(raw, w);
        mont_mul(s[i], raw, r2);
    }
    const int rounds = 8 + rp;
    for (int r = 0; r < rounds; r++) {
        const bool full = r < 4 || r >= 4 + rp;
        for (int i = 0; i < t; i++) {
            uint32_t a[8];
            add_mod2p(a, s[i], ark + ((size_t)r * t + i) * 8);
            if (full || i == 0) sbox(s[i], a);
            else
                for (int k = 0; k < 8; k++) s[i][k] = a[k];
        }
        for (int i = 0; i < t; i++) {
            uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < t; j++) {
                uint32_t pr[8], sum[8];
                mont_mul(pr, s[j], mds + ((size_t)i * t + j) * 8);
                add_mod2p(sum, acc, pr);
                for (int k = 0; k < 8; k++) acc[k] = sum[k];
            }
            for (int k = 0; k < 8; k++) nx[i][k] = acc[k];
        }
        for (int i = 0; i < t; i++)
            for (int k = 0; k < 8; k++) s[i][k] = nx[i][k];
    }
    uint32_t h[8], w[8];
    mont_redc(h, s[0]);
    csub_p_exact(h);
    limbs_to_words<LE>(w, h);
    for (int k = 0; k < 8; k++) out[idx * 8 + k] = w[k];
}

}  // namespace

cudaError_t launch_hash_dense(int t, const uint32_t* d_tbl, const void* d_in, void* d_out,
                              uint64_t n, const TagArg& tag, bool le, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    const int rp = partial_rounds(t);
    if (le)
        hash_dense_kernel<true><<<grid, 128, 0, st>>>(t, rp, d_tbl, (const uint32_t*)d_in, (uint32_t*)d_out, n, tag);
    else
        hash_dense_kernel<false><<<grid, 128, 0, st>>>(t, rp, d_tbl, (const uint32_t*)d_in, (uint32_t*)d_out, n, tag);
    return cudaGetLastError();
}

}  // namespace inf
