// Measured denominator for the integer-multiply roofline: the
// sustained issue rate of 32-bit IMAD on this device, from independent
// multiply-add chains on every SM sub-partition.  kind 0 issues plain IMAD
// (mad.lo.u32), kind 1 IMAD.WIDE.U32 (mad.wide.u32, 32x32+64->64) counted as
// two IMAD-equivalents each, the unit W(t) in BASELINE.md is expressed in.
#include <cuda_runtime.h>

#include <cstdint>

namespace inf {
namespace {

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) imad_lo_kernel(uint32_t* sink, uint32_t a, uint32_t b,
                                                       long long* cycles) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = threadIdx.x + k;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k];
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void __launch_bounds__(256) imad_wide_kernel(uint32_t* sink, uint32_t a, uint32_t b,
                                                         long long* cycles) {
    unsigned long long x[CHAINS];
    uint32_t m[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) {
        x[k] = threadIdx.x + k;
        m[k] = a + k;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(m[k]), "r"(b));
        }
    }
    const long long t1 = clock64();
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k];
    if (acc == 0x12345ull) sink[0] = (uint32_t)acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

}  // namespace

cudaError_t launch_imad_peak(int kind, int sm_count, double* imad_per_s, double* clock_mhz,
                             cudaStream_t st) {
    uint32_t* sink = nullptr;
    long long* cyc = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&sink, 4)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&cyc, 8)) != cudaSuccess) { cudaFree(sink); return e; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sm_count * 8, threads = 256;
    float best_ms = 1e30f;
    long long best_cycles = 0;
    for (int rep = 0; rep < 5 && e == cudaSuccess; rep++) {
        cudaEventRecord(e0, st);
        if (kind == 0) imad_lo_kernel<<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else imad_wide_kernel<<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        e = cudaGetLastError();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) {
            best_ms = ms;
            cudaMemcpy(&best_cycles, cyc, 8, cudaMemcpyDeviceToHost);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) return e;
    const double ops = (double)blocks * threads * (double)ITERS * 4 * CHAINS * (kind == 0 ? 1.0 : 2.0);
    *imad_per_s = ops / (best_ms * 1e-3);
    // one block's clock64 span vs. wall time of the whole grid (8 blocks per SM
    // run concurrently, 2048 threads), so this is the SM clock under this load
    *clock_mhz = best_cycles > 0 ? (double)best_cycles / (best_ms * 1e-3) / 1e6 : 0.0;
    return cudaSuccess;
}

}  // namespace inf
