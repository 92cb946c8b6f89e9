// Lone-warp issue rate of IMAD.WIDE.U32 as a function of the number of independent
// dependency chains, and the latency of the field primitives when one warp has a
// sub-partition to itself (the regime of the tree levels near the root).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I infimum_b200/csrc -I tools/experiments tools/experiments/lonewarp.cu -o tools/_bin/lonewarp
#include <cstdio>
#include <cuda_runtime.h>
#include "fr_lat.cuh"
#include "poseidon.cuh"
#include "fr29.cuh"
using namespace inf;

template <int K>
__global__ void __launch_bounds__(128) k_chains(unsigned long long* out, unsigned b, long long* cyc) {
    unsigned long long acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < K; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((unsigned)acc[k]), "r"(b));
    }
    long long t1 = clock64();
    unsigned long long x = 0;
#pragma unroll
    for (int k = 0; k < K; k++) x ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// accumulate-only chains: the multiplicands do not depend on the accumulators
template <int K>
__global__ void __launch_bounds__(128) k_acc(unsigned long long* out, unsigned b, long long* cyc) {
    unsigned long long acc[K];
    unsigned a[K];
#pragma unroll
    for (int k = 0; k < K; k++) { acc[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k; a[k] = b * (k + 3) + threadIdx.x; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < K; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b));
#pragma unroll
        for (int k = 0; k < K; k++) a[k] += 0x9e37u;
    }
    long long t1 = clock64();
    unsigned long long x = 0;
#pragma unroll
    for (int k = 0; k < K; k++) x ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// the 29-bit-limb, carry-free representation (tools/experiments/fr29.cuh): lone-warp latency
template <int OP>
__global__ void __launch_bounds__(128) k_prim29(uint32_t* out, const uint32_t* in, long long* cyc) {
    uint32_t x[3][NL], y[3][NL];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < NL; i++) { x[j][i] = (in[i] + threadIdx.x + j) & LMASK; y[j][i] = in[8 + i + j] & LMASK; }
    for (int j = 0; j < 3; j++) { x[j][NL - 1] &= 0xfffff; y[j][NL - 1] &= 0xfffff; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1000; it++) {
        uint32_t t[NL];
        if (OP == 0) mul29(t, x[0], y[0]);
        if (OP == 1) sqr29(t, x[0]);
        if (OP == 2) { uint32_t x2[NL], x4[NL]; sqr29(x2, x[0]); sqr29(x4, x2); mul29(t, x4, x[0]); }
        if (OP == 3) dot29<3, NL>(t, &x[0][0], &y[0][0], y[1], nullptr);
        if (OP == 4) { uint32_t x2[NL], x4[NL]; sqr29(x2, x[0]); sqr29(x4, x2); dot29<1, NL>(t, x4, x[0], nullptr, x[1]); }
        for (int i = 0; i < NL; i++) x[0][i] = t[i];
        x[0][NL - 1] &= 0xfffff;
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < NL; i++) acc ^= x[0][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(128) k_prim(uint32_t* out, const uint32_t* in, long long* cyc) {
    uint32_t x[3][8], y[8];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 8; i++) x[j][i] = in[i] + threadIdx.x + j;
    for (int i = 0; i < 8; i++) y[i] = in[8 + i];
    for (int j = 0; j < 3; j++) x[j][7] &= 0x1fffffff;
    y[7] &= 0x1fffffff;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1000; it++) {
        uint32_t t[8];
        if (OP == 0) mont_mul(t, x[0], y);
        if (OP == 1) mont_sqr(t, x[0]);
        if (OP == 2) sbox(t, x[0]);
        if (OP == 3) dot<3, 8>(t, &x[0][0], in + 16, in + 8);
        if (OP == 4) dot<2, 8>(t, &x[0][0], in + 16, in + 8);
        if (OP == 5) { mont_mul(t, x[0], y); add8(t, t, x[1]); csub2p(t); }
        if (OP == 6) mont_mul_lat(t, x[0], y);
        if (OP == 7) mont_sqr_lat(t, x[0]);
        if (OP == 8) sbox_lat(t, x[0]);
        if (OP == 9) dot_lat<3, 8>(t, &x[0][0], in + 16, in + 8);
        if (OP == 10) dot_lat<2, 8>(t, &x[0][0], in + 16, in + 8);
        if (OP == 11) mont_mul_lat<1>(t, x[0], y);
        if (OP == 12) mont_mul_lat<2>(t, x[0], y);
        if (OP == 13) mont_mul_lat<3>(t, x[0], y);
        for (int i = 0; i < 8; i++) x[0][i] = t[i];
        x[0][7] &= 0x3fffffff;
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < 8; i++) acc ^= x[0][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    uint32_t h[64];
    for (int i = 0; i < 64; i++) h[i] = 0x9e3779b9u * (i + 1);
    for (int j = 0; j < 8; j++) h[7 + 8 * j] &= 0x1fffffff;
    uint32_t *in, *out;
    unsigned long long* out64;
    long long* cyc;
    cudaMalloc(&in, sizeof h);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 148 * 128 * 4);
    cudaMalloc(&out64, 148 * 128 * 8);
    cudaMalloc(&cyc, 8);
    long long c;
#define CH(K)                                                                                      \
    for (int rep = 0; rep < 2; rep++) k_chains<K><<<148, 128>>>(out64, 0x12345677u, cyc);          \
    cudaDeviceSynchronize();                                                                       \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                                \
    printf("IMAD.WIDE, 1 warp per SMSP, %d independent chains: %.2f cycles per instruction\n", K, c / (4096.0 * 4 * K));
    CH(1) CH(2) CH(3) CH(4) CH(6) CH(8)
#define AC(K)                                                                                      \
    for (int rep = 0; rep < 2; rep++) k_acc<K><<<148, 128>>>(out64, 0x12345677u, cyc);             \
    cudaDeviceSynchronize();                                                                       \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                                \
    printf("IMAD.WIDE accumulate-only, 1 warp per SMSP, %d independent chains: %.2f cycles per instruction\n", K, c / (4096.0 * 4 * K));
    AC(1) AC(2) AC(3) AC(4) AC(6) AC(8) AC(12)
    const char* names[] = {"mont_mul", "mont_sqr", "sbox (sqr, sqr, mul)", "dot<3> + csub2p", "dot<2> + csub2p", "mont_mul + add8 + csub2p",
                           "mont_mul_lat", "mont_sqr_lat", "sbox_lat", "dot_lat<3> + csub2p", "dot_lat<2> + csub2p",
                           "mont_mul_lat<LEAD=1>", "mont_mul_lat<LEAD=2>", "mont_mul_lat<LEAD=3>"};
#define PR(OP)                                                                     \
    for (int rep = 0; rep < 2; rep++) k_prim<OP><<<148, 128>>>(out, in, cyc);      \
    cudaDeviceSynchronize();                                                       \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                \
    printf("lone warp, dependent %s: %.0f cycles\n", names[OP], c / 1000.0);
    PR(0) PR(1) PR(2) PR(3) PR(4) PR(5) PR(6) PR(7) PR(8) PR(9) PR(10) PR(11) PR(12) PR(13)
    const char* names29[] = {"mul29", "sqr29", "sbox29 (sqr, sqr, mul)", "dot29<3>", "sbox29 + addend (one partial round of the chain)"};
#define PR29(OP)                                                                   \
    for (int rep = 0; rep < 2; rep++) k_prim29<OP><<<148, 128>>>(out, in, cyc);    \
    cudaDeviceSynchronize();                                                       \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                \
    printf("lone warp, dependent %s: %.0f cycles\n", names29[OP], c / 1000.0);
    PR29(0) PR29(1) PR29(2) PR29(3) PR29(4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
