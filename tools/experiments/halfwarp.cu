// Does a warp with only 16 (or 8) active lanes issue wide multiplies faster?
// One warp per SM sub-partition, dependent carry chains, vary the active lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I infimum_b200/csrc tools/experiments/halfwarp.cu -o tools/_bin/halfwarp
#include <cstdio>
#include <cuda_runtime.h>
#include "fr.cuh"
using namespace inf;

__global__ void __launch_bounds__(128) k(uint32_t* out, const uint32_t* in, int active, long long* cyc) {
    const int lane = threadIdx.x & 31;
    if (lane >= active) return;
    uint32_t x[8], y[8];
    for (int i = 0; i < 8; i++) { x[i] = in[i] + threadIdx.x; y[i] = in[8 + i]; }
    x[7] &= 0x1fffffff; y[7] &= 0x1fffffff;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 2000; it++) {
        uint32_t t[8];
        mont_mul(t, x, y);
        for (int i = 0; i < 8; i++) x[i] = t[i];
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < 8; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    uint32_t h[16]; for (int i = 0; i < 16; i++) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t *in, *out; long long* cyc;
    cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 8);
    for (int active : {32, 16, 8, 1}) {
        for (int rep = 0; rep < 2; rep++) k<<<148, 128>>>(out, in, active, cyc);   // 4 warps per SM = 1 per SMSP
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("active lanes %2d: %.1f cycles per dependent mont_mul (single warp per SMSP)\n", active, c / 2000.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
