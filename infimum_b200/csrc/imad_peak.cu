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
    // x = lo32(x) * b + x : the multiplicand depends on the accumulator, so the
    // product cannot be hoisted out of the loop (a loop-invariant mad.wide is
    // strength-reduced by ptxas into 64-bit adds, which measures the ALU).
    unsigned long long x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = ((unsigned long long)(a + k) << 32) | (threadIdx.x + k);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) x[k] = (unsigned long long)(uint32_t)x[k] * b + x[k];
        }
    }
    const long long t1 = clock64();
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k];
    if (acc == 0x12345ull) sink[0] = (uint32_t)acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// kind 2: the carry-chain shape the field arithmetic actually uses — chains of
// four IMAD.WIDE.U32 linked by the carry predicate (mad.lo.cc / madc.hi.cc
// pairs), four independent chains per thread.  Counted as 2 IMAD-eq per wide.
__global__ void __launch_bounds__(256) imad_chain_kernel(uint32_t* sink, uint32_t a, uint32_t b,
                                                          long long* cycles) {
    uint32_t r[4][9], m[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int k = 0; k < 9; k++) r[c][k] = threadIdx.x + k + c;
#pragma unroll
        for (int k = 0; k < 4; k++) m[c][k] = a + k * 77 + c;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                asm volatile(
                    "mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
                    "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
                    "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
                    "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
                    "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
                    "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
                    "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
                    "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
                    "addc.u32        %8, %8, 0;"
                    : "+r"(r[c][0]), "+r"(r[c][1]), "+r"(r[c][2]), "+r"(r[c][3]), "+r"(r[c][4]),
                      "+r"(r[c][5]), "+r"(r[c][6]), "+r"(r[c][7]), "+r"(r[c][8])
                    : "r"(m[c][0]), "r"(m[c][1]), "r"(m[c][2]), "r"(m[c][3]), "r"(b));
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 9; k++) acc ^= r[c][k];
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// kind 3: mad.hi.u32 alone (IMAD.HI), 8 independent chains.
__global__ void __launch_bounds__(256) imad_hi_kernel(uint32_t* sink, uint32_t a, uint32_t b,
                                                       long long* cycles) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = threadIdx.x * 0x01010101u + k;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++)
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k];
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// kind 4: fma.rn.f64 alone (DFMA), 8 independent chains -- the FP64 pipe's rate, for
// judging a 52-bit-limb double-precision formulation of the field product.
// kind 5: DFMA and IMAD.WIDE.U32 interleaved 1:1 (16 + 16 per iteration), kind 6: 2:1
// (the ratio a hi/lo split product needs): do the two pipes overlap?
template <int DF, int WIDE>
__global__ void __launch_bounds__(256) dfma_mix_kernel(uint32_t* sink, uint32_t a, uint32_t b,
                                                        long long* cycles) {
    double d[CHAINS];
    unsigned long long x[CHAINS];
    const double fa = 0.99999 + 1e-9 * (double)(a & 0xff), fb = 1e-3 + 1e-9 * (double)(b & 0xff);
#pragma unroll
    for (int k = 0; k < CHAINS; k++) {
        d[k] = 1.0 + threadIdx.x + k;
        x[k] = ((unsigned long long)(a + k) << 32) | (threadIdx.x + k);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 32 / (DF + WIDE); u++) {
#pragma unroll
            for (int k = 0; k < DF; k++)
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k % CHAINS]) : "d"(fa), "d"(fb));
#pragma unroll
            for (int k = 0; k < WIDE; k++) x[k % CHAINS] = (unsigned long long)(uint32_t)x[k % CHAINS] * b + x[k % CHAINS];
        }
    }
    const long long t1 = clock64();
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc ^= x[k] ^ (unsigned long long)__double_as_longlong(d[k]);
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
        else if (kind == 1) imad_wide_kernel<<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else if (kind == 2) imad_chain_kernel<<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else if (kind == 3) imad_hi_kernel<<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else if (kind == 4) dfma_mix_kernel<8, 0><<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else if (kind == 5) dfma_mix_kernel<8, 8><<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
        else dfma_mix_kernel<8, 4><<<blocks, threads, 0, st>>>(sink, 0x9e3779b1u, 0x7f4a7c15u, cyc);
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
    // instructions per thread per iteration: 32 (4 x 8 chains, or 2 x 4 chains x 4 wide), 24 for kind 6
    // (2 x (8 DFMA + 4 wide)); kinds 1, 2 count a wide multiply as two IMAD-equivalents, kinds 4-6
    // report plain instructions per second (DFMA and wide multiplies alike)
    const double ops = (double)blocks * threads * (double)ITERS * (kind == 6 ? 24 : 32) *
                       ((kind == 1 || kind == 2) ? 2.0 : 1.0);
    *imad_per_s = ops / (best_ms * 1e-3);
    // SM cycles one resident block spent in the loop over the wall time of the
    // launch: the SM clock under this load, to first order (the grid is exactly
    // one wave of 8 x 256 threads per SM)
    *clock_mhz = best_cycles > 0 ? (double)best_cycles / (best_ms * 1e-3) / 1e6 : 0.0;
    return cudaSuccess;
}

}  // namespace inf
