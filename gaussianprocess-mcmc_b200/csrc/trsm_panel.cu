// Panel triangular solve of the blocked Cholesky:  X L11^T = A21  (in place), i.e. the rows of L below a
// factored 128x128 diagonal block -- the dtrsm inside LAPACK dpotrf (kcGP.tools.jitchol, sliceSample.py:196,205).
//
// It is a true substitution, not a multiplication by an explicit inverse: chol(R + 1e-11 I) at sliceSample.py:205
// has condition ~1e11, and  A21 * inv(L11)  would lose cond(L11) * eps (measured: spurious "not positive definite"
// pivots); substitution keeps the residual at eps * |X| |L11|.
//
// One CTA owns 128 rows of the panel in shared memory.  The 128 columns are processed in four sub-blocks of 32:
//   solve : one row per thread, its 32 entries in registers, L entries broadcast from shared memory (row_trsv32)
//   update: the not-yet-solved columns get  R[:, later] -= R[:, sb] * L[later, sb]^T  on FP64 DMMA fragments.
// L11 is kept as its ten lower 32x32 blocks (stride 36: conflict-free fragments), the row tile with stride 132.
#include "common.cuh"
#include "tri_solve.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int TP_ROWS = 128;
constexpr int TP_T = 132;                     // row-tile stride
constexpr int TP_B = 36;                      // L block stride
constexpr int TP_THREADS = 256;
constexpr int TP_LBLK = 32 * TP_B;            // doubles per 32x32 L block
constexpr int TP_SMEM = (TP_ROWS * TP_T + 10 * TP_LBLK + NB) * (int)sizeof(double);    // 165,376 B

__device__ __forceinline__ int lblk_index(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }     // bi >= bj

__device__ __forceinline__ void dmma884_t(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(TP_THREADS, 1)
trsm_panel_kernel(BatchView A, int n, int j0)
{
    extern __shared__ __align__(16) double sm[];
    double *R = sm;                               // [128][132] rows of the panel
    double *Lb = sm + TP_ROWS * TP_T;             // 10 lower blocks of L11
    double *dinv = Lb + 10 * TP_LBLK;             // [128]
    const int b = blockIdx.y;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride;
    const int ld = A.ld;
    const int row0 = j0 + NB + blockIdx.x * TP_ROWS;
    const int rows_valid = min(TP_ROWS, n - row0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;

    // L11 (rows/cols j0 .. j0+127), lower 32-blocks; the strict upper part of diagonal blocks is never read
    for (int e = tid; e < 10 * 32 * 16; e += TP_THREADS) {        // 16 double2 per block row
        const int blk = e / (32 * 16), rem = e - blk * 32 * 16;
        const int r = rem / 16, c2 = (rem - r * 16) * 2;
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
        const int bj = blk - bi * (bi + 1) / 2;
        const double2 v = *reinterpret_cast<const double2 *>(Ab + (size_t)(j0 + bi * 32 + r) * ld + j0 + bj * 32 + c2);
        *reinterpret_cast<double2 *>(&Lb[blk * TP_LBLK + r * TP_B + c2]) = v;
    }
    // panel rows (zero beyond the matrix)
    for (int e = tid; e < TP_ROWS * 64; e += TP_THREADS) {
        const int r = e >> 6, c2 = (e & 63) * 2;
        double2 v = make_double2(0.0, 0.0);
        if (r < rows_valid) v = *reinterpret_cast<const double2 *>(Ab + (size_t)(row0 + r) * ld + j0 + c2);
        *reinterpret_cast<double2 *>(&R[r * TP_T + c2]) = v;
    }
    __syncthreads();
    if (tid < NB) {
        const int bi = tid >> 5, r = tid & 31;
        dinv[tid] = 1.0 / Lb[lblk_index(bi, bi) * TP_LBLK + r * TP_B + r];
    }
    __syncthreads();

    for (int sb = 0; sb < 4; ++sb) {
        // ---- solve 32 columns: one row per thread
        if (tid < TP_ROWS) {
            double x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = R[tid * TP_T + sb * 32 + c];
            row_trsv32(x, Lb + lblk_index(sb, sb) * TP_LBLK, TP_B, dinv + sb * 32);
#pragma unroll
            for (int c = 0; c < 32; ++c) R[tid * TP_T + sb * 32 + c] = x[c];
        }
        __syncthreads();
        // ---- update the later column blocks with DMMA:  R[:, cb2] -= R[:, sb] * L[cb2, sb]^T
        if (sb < 3) {
            int u = 0;
            for (int cb2 = sb + 1; cb2 < 4; ++cb2) {
                const double *Lq = Lb + lblk_index(cb2, sb) * TP_LBLK;
                for (int rb = 0; rb < TP_ROWS / 8; ++rb) {
                    for (int c8 = 0; c8 < 4; ++c8, ++u) {
                        if ((u & 7) != warp) continue;
                        double2 *cp = reinterpret_cast<double2 *>(&R[(rb * 8 + fr) * TP_T + cb2 * 32 + c8 * 8 + 2 * fk]);
                        double2 cv = *cp;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const double av = R[(rb * 8 + fr) * TP_T + sb * 32 + ks * 4 + fk];
                            const double bv = Lq[(c8 * 8 + fr) * TP_B + ks * 4 + fk];
                            dmma884_t(cv.x, cv.y, -av, bv);
                        }
                        *cp = cv;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < TP_ROWS * 64; e += TP_THREADS) {
        const int r = e >> 6, c2 = (e & 63) * 2;
        if (r < rows_valid)
            *reinterpret_cast<double2 *>(Ab + (size_t)(row0 + r) * ld + j0 + c2) = *reinterpret_cast<const double2 *>(&R[r * TP_T + c2]);
    }
}

int launch_trsm_panel(BatchView A, int n, int j0, int B, cudaStream_t s)
{
    const int rows = n - j0 - NB;
    if (B <= 0 || rows <= 0) return 0;
    if ((A.ld & 1) || (j0 & 1)) { set_error("trsm_panel: ld=%d j0=%d must be even", A.ld, j0); return GPMC_EALIGN; }
    static bool attr_set = false;
    if (!attr_set) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
        attr_set = true;
    }
    dim3 grid((rows + TP_ROWS - 1) / TP_ROWS, B);
    prof_begin(KC_TRSM, s);
    trsm_panel_kernel<<<grid, TP_THREADS, TP_SMEM, s>>>(A, n, j0);
    prof_end(KC_TRSM, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
