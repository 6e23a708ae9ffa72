// Panel kernel of the blocked Cholesky: factor one 128x128 diagonal block in shared memory and
// produce its inverse in the same sweep.
//
// Part of kcGP.tools.jitchol -> LAPACK dpotrf(lower=1) (sliceSample.py:196,205): the unblocked
// dpotf2 step on the diagonal block, with LAPACK's failure convention (info = index of the first
// non-positive / NaN pivot, 1-based, sliceSample's jitchol turns it into a jitter retry).
//
// The block is held as one 128x129 array T in shared memory: the lower triangle becomes L11, and the
// strict upper triangle accumulates X = L11^-T by running the same column operations on an appended
// identity (rows of [A11; I] are updated alike).  W = L11^-1 is written out dense so the panel TRSM
// below the block is a DMMA GEMM (mode 1 of gemm_dmma.cu).
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int PT = 129;                       // smem row stride (doubles), odd -> conflict-free columns
constexpr int POTF2_THREADS = 256;
constexpr int POTF2_SMEM = (NB * PT + NB) * (int)sizeof(double);

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(BatchView A, int n, int j0, double *__restrict__ W, long long strideW, int *__restrict__ info,
                 int zero_upper)
{
    extern __shared__ __align__(16) double sm[];
    double *T = sm;                           // [NB][PT]
    double *dinv = sm + NB * PT;              // 1 / L_kk
    const int b = blockIdx.x;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride + (size_t)j0 * A.ld + j0;
    const int ld = A.ld;
    const int nv = min(NB, n - j0);           // valid order of this block
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // load lower triangle; strict upper = 0 (identity's off-diagonal); padding = identity
    for (int r = warp; r < NB; r += POTF2_THREADS / 32) {
        for (int c = lane; c < NB; c += 32) {
            double v = 0.0;
            if (r < nv && c <= r) v = Ab[(size_t)r * ld + c];
            else if (r >= nv && c == r) v = 1.0;
            T[r * PT + c] = v;
        }
    }
    __syncthreads();

    int fail = 0;
    for (int k = 0; k < NB; ++k) {
        const double akk = T[k * PT + k];
        if (!(akk > 0.0) && fail == 0) fail = j0 + k + 1;          // dpotf2: ajj <= 0 or NaN
        const double d = sqrt(akk);
        const double rinv = 1.0 / d;
        __syncthreads();                                            // everyone has read akk
        // scale column k: L part below the diagonal (divide), X part above it (x / d)
        if (tid < NB) {
            const int t = tid;
            if (t > k)      T[t * PT + k] = T[t * PT + k] / d;
            else if (t < k) T[t * PT + k] = T[t * PT + k] * rinv;
            else          { T[k * PT + k] = d; dinv[k] = rinv; }
        }
        __syncthreads();
        // trailing update of columns j > k:  T[t][j] -= colk(t) * L[j][k]
        //   rows t <= k  : X part (colk(k) = 1/d), all j > k
        //   rows t >  k  : L part, k < j <= t
        double ljk[4];
        int jj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            jj[q] = k + 1 + lane + 32 * q;
            ljk[q] = (jj[q] < NB) ? T[jj[q] * PT + k] : 0.0;
        }
        for (int t = warp; t < NB; t += POTF2_THREADS / 32) {
            const double ck = (t == k) ? rinv : T[t * PT + k];
            const int jend = (t <= k) ? NB - 1 : t;                 // last column touched in this row
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (jj[q] <= jend) T[t * PT + jj[q]] -= ck * ljk[q];
            }
        }
        // no barrier needed here: the next iteration's first barrier orders these writes before
        // column k+1 is scaled; akk of the next step is read after ... the writes of this step -> sync
        __syncthreads();
    }

    if (fail != 0 && tid == 0) {
        if (info[m] == 0) info[m] = fail;
    }

    // write L11 back (lower incl. diagonal); optionally zero the strict upper triangle
    for (int r = warp; r < nv; r += POTF2_THREADS / 32) {
        for (int c = lane; c < nv; c += 32) {
            if (c <= r) Ab[(size_t)r * ld + c] = T[r * PT + c];
            else if (zero_upper) Ab[(size_t)r * ld + c] = 0.0;
        }
    }
    // W = L11^-1, dense row-major [NB][NB]:  W[c][k] = X[k][c] for k < c, 1/L_cc on the diagonal
    double *Wb = W + (size_t)m * strideW;
    for (int c = warp; c < NB; c += POTF2_THREADS / 32) {
        for (int k = lane; k < NB; k += 32) {
            double v = 0.0;
            if (k < c) v = T[k * PT + c];
            else if (k == c) v = dinv[c];
            Wb[c * NB + k] = v;
        }
    }
}

int launch_potf2(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                 int B, cudaStream_t s)
{
    if (B <= 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
        attr_set = true;
    }
    prof_begin(KC_POTF2, s);
    potf2_inv_kernel<<<B, POTF2_THREADS, POTF2_SMEM, s>>>(A, n, j0, W, strideW, info, zero_upper);
    prof_end(KC_POTF2, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
