// Register-resident FP64 peak probes: what the chip can issue when nothing but the FP64 pipe is busy.
// bench.py reports them beside cuBLAS DGEMM so the Cholesky's roofline fraction has a measured denominator
// (MEASURED_PEAKS.json carries no FP64 figure).
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

__global__ void __launch_bounds__(512) dmma_peak_kernel(double *out, int iters)
{
    double c[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
    const double a = 1.0 + 1e-9 * threadIdx.x, bb = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(bb));
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(512) dfma_peak_kernel(double *out, int iters)
{
    double c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 1e-3 * j;
    const double a = 1.0 + 1e-9 * threadIdx.x, bb = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = fma(c[j], a, bb);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j];
    if (s == 123.456) out[0] = s;
}

// DMMA with NACC independent accumulators per warp: the issue rate as a function of the instruction-level
// parallelism tells the accumulate latency of DMMA.8x8x4 (needed to size warp tiles).
template <int NACC>
__global__ void __launch_bounds__(1024) dmma_ilp_kernel(double *out, int iters)
{
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
    const double a = 1.0 + 1e-9 * threadIdx.x, bb = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(bb));
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

int run_dmma_ilp(int nacc, int warps_per_sm, int iters, double *tflops)
{
    int dev = 0, sms = 0;
    GPMC_CUDA_CHECK(cudaGetDevice(&dev));
    GPMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr;
    GPMC_CUDA_CHECK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    GPMC_CUDA_CHECK(cudaEventCreate(&e0));
    GPMC_CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    const int threads = warps_per_sm * 32;
    for (int rep = 0; rep < 3; ++rep) {
        GPMC_CUDA_CHECK(cudaEventRecord(e0));
        switch (nacc) {
            case 1: dmma_ilp_kernel<1><<<sms, threads>>>(out, iters); break;
            case 2: dmma_ilp_kernel<2><<<sms, threads>>>(out, iters); break;
            case 4: dmma_ilp_kernel<4><<<sms, threads>>>(out, iters); break;
            case 8: dmma_ilp_kernel<8><<<sms, threads>>>(out, iters); break;
            case 16: dmma_ilp_kernel<16><<<sms, threads>>>(out, iters); break;
            case 32: dmma_ilp_kernel<32><<<sms, threads>>>(out, iters); break;
            default: set_error("dmma_ilp: nacc must be 1,2,4,8,16,32"); return GPMC_EINVAL;
        }
        GPMC_CUDA_CHECK(cudaEventRecord(e1));
        GPMC_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        GPMC_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    GPMC_LAUNCH_CHECK();
    *tflops = (double)sms * warps_per_sm * (double)iters * nacc * 512.0 / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

int run_fp64_peak(int which, int iters, double *tflops, double *ms_out)
{
    int dev = 0, sms = 0;
    GPMC_CUDA_CHECK(cudaGetDevice(&dev));
    GPMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr;
    GPMC_CUDA_CHECK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    GPMC_CUDA_CHECK(cudaEventCreate(&e0));
    GPMC_CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = sms * 2, threads = 512;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        GPMC_CUDA_CHECK(cudaEventRecord(e0));
        if (which == 0) dmma_peak_kernel<<<blocks, threads>>>(out, iters);
        else dfma_peak_kernel<<<blocks, threads>>>(out, iters);
        GPMC_CUDA_CHECK(cudaEventRecord(e1));
        GPMC_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        GPMC_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    GPMC_LAUNCH_CHECK();
    const double warps = (double)blocks * threads / 32.0;
    const double flops = which == 0 ? warps * (double)iters * 16.0 * 512.0
                                    : (double)blocks * threads * (double)iters * 16.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

}  // namespace gpmc
