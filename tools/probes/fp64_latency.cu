#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void probe(double *out, long long *cyc, double seed, int n)
{
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = seed;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, c0 = 0.0, c1 = 0.0;
    long long t0, t1;
    // 0: dependent DFMA
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); cyc[0] = t1 - t0;
    // 1: dependent DMMA (accumulator chain)
    t0 = clock64();
    for (int i = 0; i < n; ++i) dmma(c0, c1, x, y);
    t1 = clock64(); cyc[1] = t1 - t0;
    // 2: DMMA chain through the A operand (result feeds next A)
    double a = x;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { double d0 = 0.0, d1 = 0.0; dmma(d0, d1, a, y); a = d0; }
    t1 = clock64(); cyc[2] = t1 - t0;
    // 3: dependent 64-bit shuffle
    double s = x;
    t0 = clock64();
    for (int i = 0; i < n; ++i) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
    t1 = clock64(); cyc[3] = t1 - t0;
    // 4: MUFU.RSQ64H approx chain
    double r = fabs(x) + 1.0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { double q; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + 1.0; }
    t1 = clock64(); cyc[4] = t1 - t0;   // includes one DADD
    // 5: library rsqrt chain
    double r2 = fabs(x) + 1.0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) r2 = rsqrt(r2) + 1.0;
    t1 = clock64(); cyc[5] = t1 - t0;
    // 6: dependent LDS (pointer chase through shared memory values)
    int idx = threadIdx.x & 63; double acc = 0.0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { double v = sm[idx]; acc += v; idx = (idx + (int)v + 1) & 63; }
    t1 = clock64(); cyc[6] = t1 - t0;
    // 7: dependent DMUL
    double m = x;
    t0 = clock64();
    for (int i = 0; i < n; ++i) m = m * y;
    t1 = clock64(); cyc[7] = t1 - t0;
    // 8: __syncthreads cost with 8 warps
    t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64(); cyc[8] = t1 - t0;
    out[threadIdx.x] = x + c0 + c1 + a + s + r + r2 + acc + m;
}
int main() {
    double *out; long long *cyc, h[9];
    cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 9 * 8);
    const int n = 1000;
    for (int threads : {32, 256}) {
        probe<<<1, threads>>>(out, cyc, 0.0, n);
        cudaDeviceSynchronize();
        probe<<<1, threads>>>(out, cyc, 0.0, n);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const char *names[9] = {"DFMA dep", "DMMA acc-chain", "DMMA A-chain", "SHFL64 dep", "MUFU.RSQ64H+DADD", "rsqrt()+DADD", "LDS chase (+I2F..)", "DMUL dep", "__syncthreads"};
        for (int k = 0; k < 9; ++k) printf("threads=%d %-22s %.1f cycles/op\n", threads, names[k], (double)h[k] / n);
    }
    return 0;
}
