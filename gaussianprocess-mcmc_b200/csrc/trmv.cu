// Triangular matrix-vector products of the SDS proposal (HBM-read bound, 8*n*(n+1)/2 bytes per item):
//   mode 0, lower:  f' = C eta + m                         sliceSample.py:140  (np.dot(chol_R_theta, ita) + m_theta_g)
//   mode 2, lower:  nu = L z                               sliceSample.py:41   (the Cholesky draw of elliptical_slice)
//   mode 1, upper:  m  = g - S * (U z),  U = L^-T, z = L^-1 g   i.e.  m = R S^-1 g = g - S (K+S)^-1 g
//                                                          sliceSample.py:204  (np.dot(np.dot(R_theta, inv(S)), g))
// One warp per row, lanes stride the row with double2 loads (512 contiguous bytes per warp instruction);
// a CTA takes 32 consecutive rows, the grid covers (row blocks, batch items).
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int TRMV_ROWS = 32;
constexpr int TRMV_THREADS = 256;

__global__ void __launch_bounds__(TRMV_THREADS)
trmv_kernel(BatchView T, int n, int upper, int mode, const double *__restrict__ x, const double *__restrict__ add,
            const double *__restrict__ svec, int ldv, double *__restrict__ out)
{
    const int b = blockIdx.y;
    if (T.count && b >= *T.count) return;
    const int m = batch_item(T, b);
    const double *Tb = T.base + (size_t)m * T.stride;
    const double *xv = x + (size_t)m * ldv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < TRMV_ROWS; rr += TRMV_THREADS / 32) {
        const int r = blockIdx.x * TRMV_ROWS + rr;
        if (r >= n) break;
        const double *row = Tb + (size_t)r * T.ld;
        // contraction range [k_lo, k_hi): lower -> [0, r], upper -> [r, n)
        const int k_lo = upper ? r : 0, k_hi = upper ? n : r + 1;
        double acc = 0.0;
        const int ka = (k_lo + 1) & ~1;                 // first even index >= k_lo
        const int kb = k_hi & ~1;                       // even end of the vector part
        if (k_lo < ka && k_lo < k_hi && lane == 0) acc = fma(row[k_lo], xv[k_lo], acc);
#pragma unroll 8
        for (int k = ka + 2 * lane; k < kb; k += 64) {
            const double2 t = *reinterpret_cast<const double2 *>(row + k);
            const double2 v = *reinterpret_cast<const double2 *>(xv + k);
            acc = fma(t.x, v.x, acc);
            acc = fma(t.y, v.y, acc);
        }
        if (kb < k_hi && kb >= ka && lane == 1) acc = fma(row[kb], xv[kb], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const size_t o = (size_t)m * ldv + r;
            out[o] = (mode == 0) ? acc + add[o] : (mode == 1 ? add[o] - svec[o] * acc : acc);
        }
    }
}

int launch_trmv(BatchView T, int n, int upper, int mode, const double *x, const double *add, const double *svec,
                int ldv, double *out, int B, cudaStream_t s)
{
    if (B <= 0) return 0;
    if ((ldv & 1) || (T.ld & 1)) { set_error("trmv: ldv=%d / ld=%d must be even", ldv, T.ld); return GPMC_EALIGN; }
    dim3 grid((n + TRMV_ROWS - 1) / TRMV_ROWS, B);
    prof_begin(KC_VEC, s);
    trmv_kernel<<<grid, TRMV_THREADS, 0, s>>>(T, n, upper, mode, x, add, svec, ldv, out);
    prof_end(KC_VEC, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
