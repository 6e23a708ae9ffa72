// Truncated-Gaussian log-likelihood (likK.TruncatedGauss2.evaluate(y=, mu=), ASSUMPTION-1 of the oracle shim),
// shared by the sweep's control kernels and the stand-alone entry point.
#pragma once

namespace gpmc {

// log(Phi(b) - Phi(a)), a < b, three-branch form shared with the oracle's TruncatedGauss2 (ASSUMPTION-1)
__device__ __forceinline__ double tg2_log_mass(double a, double b)
{
    const double SQRT2 = 1.4142135623730951;
    if (a > 0.0) return log(0.5 * (erfc(a / SQRT2) - erfc(b / SQRT2)));
    if (b < 0.0) return log(0.5 * (erfc(-b / SQRT2) - erfc(-a / SQRT2)));
    return log(0.5 * (erf(b / SQRT2) - erf(a / SQRT2)));
}

// sum_i log TN(y_i - my; mu_i, sn, [lower, upper])  -- likK.TruncatedGauss2.evaluate(y=, mu=), sliceSample.py:118,143
__device__ __forceinline__ double tg2_loglik_block(const double *__restrict__ y, double my, const double *__restrict__ mu, int n, double sn,
                                   double lower, double upper, double *red)
{
    const double HALF_LOG_2PI = 0.9189385332046727;
    const double logsn = log(sn);
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double yi = y[i] - my, mi = mu[i];
        const double r = (yi - mi) / sn;
        const double a = (lower - mi) / sn, b = (upper - mi) / sn;
        acc += -0.5 * r * r - HALF_LOG_2PI - logsn - tg2_log_mass(a, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    return tot;
}

}  // namespace gpmc
