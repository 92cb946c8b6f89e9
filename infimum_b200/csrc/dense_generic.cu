// Generic dense Poseidon kernel: any width 2..13, the reference's schedule
// taken literally (pallet/src/hash/poseidon.rs:184-203: ARK, x^5, dense MDS,
// every round), table of Montgomery-form (ark, mds) in global memory, state in
// local memory with run-time indexing.  It is slow by design and serves two
// purposes: widths 9..13, which no caller on the hot path uses but
// `Poseidon::new_circom` accepts (poseidon.rs:315), and an on-device
// cross-check of the optimised kernels that shares none of their tables.
// It is also what runs `Poseidon::new(params)` (poseidon.rs:105-108) with
// caller-supplied PoseidonParameters: any round counts and any S-box exponent.
#include <cuda_runtime.h>

#include "launch.h"
#include "poseidon.cuh"

namespace inf {
namespace {

__device__ __forceinline__ void add_mod2p(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b) {
    add8(r, a, b);   // a, b < 2p + eps  =>  sum < 4p + eps < 2^256
    csub2p(r);
}

// y = x^alpha (a.pow([alpha]), poseidon.rs:135,142), square and multiply from the low bit
__device__ void pow_alpha(uint32_t (&y)[8], const uint32_t (&x)[8], uint64_t alpha) {
    if (alpha == 5) {
        sbox(y, x);
        return;
    }
    // R mod p: one in Montgomery form
    uint32_t r[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                     0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    uint32_t base[8], tmp[8];
    for (int k = 0; k < 8; k++) base[k] = x[k];
    for (uint64_t e = alpha; e; e >>= 1) {
        if (e & 1) {
            mont_mul(tmp, r, base);
            for (int k = 0; k < 8; k++) r[k] = tmp[k];
        }
        if (e >> 1) {
            mont_sqr(tmp, base);
            for (int k = 0; k < 8; k++) base[k] = tmp[k];
        }
    }
    for (int k = 0; k < 8; k++) y[k] = r[k];
}

// Rounds 0..half-1 and half+rp..rounds-1 are full, the rp in between partial
// (poseidon.rs:184-203; half = full_rounds / 2, rounds = full_rounds + partial_rounds).
template <bool LE>
__global__ void __launch_bounds__(128)
hash_dense_kernel(int t, int half, int rp, int rounds, uint64_t alpha, const uint32_t* __restrict__ tbl,
                  const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n, TagArg tag) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t* ark = tbl;
    const uint32_t* mds = tbl + (size_t)rounds * t * 8;
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
    for (int r = 0; r < rounds; r++) {
        const bool full = r < half || r >= half + rp;
        for (int i = 0; i < t; i++) {
            uint32_t a[8];
            add_mod2p(a, s[i], ark + ((size_t)r * t + i) * 8);
            if (full || i == 0) pow_alpha(s[i], a, alpha);
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

cudaError_t launch_hash_dense_params(int t, int full_rounds, int rp, uint64_t alpha, const uint32_t* d_tbl,
                                     const void* d_in, void* d_out, uint64_t n, const TagArg& tag, bool le,
                                     cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    const int half = full_rounds / 2, rounds = full_rounds + rp;
    if (le)
        hash_dense_kernel<true><<<grid, 128, 0, st>>>(t, half, rp, rounds, alpha, d_tbl, (const uint32_t*)d_in,
                                                       (uint32_t*)d_out, n, tag);
    else
        hash_dense_kernel<false><<<grid, 128, 0, st>>>(t, half, rp, rounds, alpha, d_tbl, (const uint32_t*)d_in,
                                                        (uint32_t*)d_out, n, tag);
    return cudaGetLastError();
}

cudaError_t launch_hash_dense(int t, const uint32_t* d_tbl, const void* d_in, void* d_out,
                              uint64_t n, const TagArg& tag, bool le, cudaStream_t st) {
    return launch_hash_dense_params(t, 8, partial_rounds(t), 5, d_tbl, d_in, d_out, n, tag, le, st);
}

}  // namespace inf
