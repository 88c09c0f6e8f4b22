// int_peak.cuh -- issue-rate microbenchmark for the integer pipe the DP fill
// is bound by (SURVEY.md 8d asks for the measured denominator of the roofline).
// Eight independent dependency chains per thread of
//   (a) VIADDMNMX.S16x2  (__viaddmax_s16x2: add + max on two int16 lanes)
//   (b) VIADDMNMX (s32)  (__viaddmax_s32)
// Reported as lane-instructions per second: threads * instructions * lanes / s,
// i.e. one "op" per int16 (resp. int32) lane per issued instruction, the same
// unit as SURVEY's  N_SM x 64 lanes/clk x f x 2.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace lb2 {

template <bool PACKED>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8], b = seed | 0x00010001u, c = seed * 3u;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 8u + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (PACKED) a[k] = __viaddmax_s16x2(a[k], b, c);
                else a[k] = (uint32_t)__viaddmax_s32((int)a[k], (int)b, (int)c);
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) x ^= a[k];
    if (x == 0x12345678u) out[0] = x;      // keep the chains alive
}

// returns 0 on success; g16 / g32 in 1e9 lane-instructions per second
static inline int int_peak_measure(cudaStream_t s, int sm_count, double* g16, double* g32) {
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = sm_count * 8, block = 256;
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0, s);
            if (mode == 0) int_peak_kernel<true><<<grid, block, 0, s>>>(d, 12345u + rep, iters);
            else int_peak_kernel<false><<<grid, block, 0, s>>>(d, 12345u + rep, iters);
            cudaEventRecord(e1, s);
            if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return 1; }
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double instr = (double)grid * block * iters * 64.0;
        res[mode] = instr * (mode == 0 ? 2.0 : 1.0) / (best * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *g16 = res[0]; *g32 = res[1];
    return 0;
}

}  // namespace lb2
