// Panel triangular solve of the blocked Cholesky:  X L11^T = A21  (in place), i.e. the rows of L below a
// factored 128x128 diagonal block -- the dtrsm inside LAPACK dpotrf (kcGP.tools.jitchol, sliceSample.py:196,205).
//
// It is a true substitution, not a multiplication by an explicit inverse: chol(R + 1e-11 I) at sliceSample.py:205
// has condition ~1e11, and  A21 * inv(L11)  would lose cond(L11) * eps (measured: spurious "not positive definite"
// pivots); substitution keeps the residual at eps * |X| |L11|.
//
// One CTA owns 64 rows of the panel (two CTAs per SM).  Each of its 8 warps keeps 8 rows x 128 columns as FP64 DMMA accumulator
// fragments in registers for the whole kernel and, after the initial load of L11, never meets a block barrier.
// The 128 columns are processed in four sub-blocks of 32:
//   solve : substitution in the fragment layout itself -- column c of a row lives in one lane of the row's quad;
//           the solved value is broadcast with a quad shuffle and each lane updates its own later columns with L
//           entries read from shared memory (no explicit inverse anywhere)
//   update: the later columns get  acc[:, later] -= X[:, sb] * L[later, sb]^T  on DMMA; the solved sub-block is
//           staged through the warp's own shared-memory rows to become A fragments (and goes out to global).
// L11 is kept as its lower 32x32 blocks, staged with cp.async.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int TP_ROWS = 64;                   // rows per CTA; two CTAs per SM overlap each other's load/store phases
constexpr int TP_B = 34;                      // stride of L blocks and of the staged sub-block: the substitution's
                                              // L[j][c] reads (j = 2*fk + ..) fall in 4 distinct bank groups
constexpr int TP_THREADS = 256;               // 8 warps x 8 rows, 2 CTAs per SM: the kernel is latency bound, warps hide it
constexpr int TP_LBLK = 32 * TP_B;            // doubles per 32x32 L block
constexpr int TP_NSB = NB / 32;               // 32-column sub-blocks of the panel
constexpr int TP_NLB = TP_NSB * (TP_NSB + 1) / 2;   // lower 32x32 blocks of L11
constexpr int TP_NC8 = NB / 8;                // 8-column fragments per row
constexpr int TP_SMEM = (TP_ROWS * TP_B + TP_NLB * TP_LBLK + NB) * (int)sizeof(double);

__device__ __forceinline__ int lblk_index(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }     // bi >= bj

__device__ __forceinline__ void dmma884_t(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(TP_THREADS, NB == 64 ? 3 : 2)
trsm_panel_kernel(BatchView A, int n, int j0)
{
    extern __shared__ __align__(16) double sm[];
    double *R = sm;                               // [128][36] the sub-block being solved
    double *Lb = sm + TP_ROWS * TP_B;             // 10 lower blocks of L11
    double *dinv = Lb + TP_NLB * TP_LBLK;             // [128]
    const int b = blockIdx.y;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride;
    const int ld = A.ld;
    const int row0 = j0 + NB + blockIdx.x * TP_ROWS;
    const int rows_valid = min(TP_ROWS, n - row0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;

    // L11 (rows/cols j0 .. j0+127), lower 32-blocks, 16-byte cp.async pieces (all in flight at once);
    // the strict upper part of diagonal blocks is never read
    for (int e = tid; e < TP_NLB * 32 * 16; e += TP_THREADS) {        // 16 pieces per block row
        const int blk = e / (32 * 16), rem = e - blk * 32 * 16;
        const int r = rem / 16, c2 = (rem - r * 16) * 2;
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
        const int bj = blk - bi * (bi + 1) / 2;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&Lb[blk * TP_LBLK + r * TP_B + c2]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ab + (size_t)(j0 + bi * 32 + r) * ld + j0 + bj * 32 + c2));
    }
    asm volatile("cp.async.commit_group;\n" ::);
    // this warp's 8 rows as accumulator fragments: acc[cb8] = row warp*8 + fr, cols cb8*8 + 2fk, +1
    double acc[TP_NC8][2];
    {
        const int r = warp * 8 + fr;
        const double *src = Ab + (size_t)(row0 + min(r, rows_valid - 1)) * ld + j0 + 2 * fk;
#pragma unroll
        for (int cb8 = 0; cb8 < TP_NC8; ++cb8) {
            double2 v = make_double2(0.0, 0.0);
            if (r < rows_valid) v = *reinterpret_cast<const double2 *>(src + cb8 * 8);
            acc[cb8][0] = v.x;
            acc[cb8][1] = v.y;
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (tid < NB) {
        const int bi = tid >> 5, r = tid & 31;
        dinv[tid] = 1.0 / Lb[lblk_index(bi, bi) * TP_LBLK + r * TP_B + r];
    }
    __syncthreads();
    // From here on every warp works on its own 16 rows only (L and dinv are read-only): no block barriers.
    double *Rw = R + warp * 8 * TP_B;             // this warp's staging rows
#pragma unroll
    for (int sb = 0; sb < TP_NSB; ++sb) {
        const double *Ld = Lb + lblk_index(sb, sb) * TP_LBLK;
        // ---- solve the sub-block's 32 columns in fragment layout.  Column c lives in lane fk == (c%8)/2 of each
        //      row's quad; the solved value is broadcast inside the quad and every lane updates its own later columns.
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int CB = c >> 3, OW = (c & 7) >> 1, S = c & 1;
            const double xs = __shfl_sync(0xffffffffu, acc[sb * 4 + CB][S] * dinv[sb * 32 + c], (lane & ~3) | OW);
            if (fk == OW) acc[sb * 4 + CB][S] = xs;
#pragma unroll
            for (int CB2 = CB; CB2 < 4; ++CB2) {
#pragma unroll
                for (int S2 = 0; S2 < 2; ++S2) {
                    const int j = CB2 * 8 + 2 * fk + S2;              // this lane's column
                    if (CB2 > CB || j > c) acc[sb * 4 + CB2][S2] = fma(-xs, Ld[j * TP_B + c], acc[sb * 4 + CB2][S2]);
                }
            }
        }
        // ---- stage the solved columns (this warp's rows): final values -> global, and A fragments for the update
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2 *>(&Rw[fr * TP_B + q * 8 + 2 * fk]) = make_double2(acc[sb * 4 + q][0], acc[sb * 4 + q][1]);
        __syncwarp();
        for (int e = lane; e < 8 * 16; e += 32) {
            const int r = e >> 4, c2 = (e & 15) * 2;
            if (warp * 8 + r < rows_valid)
                *reinterpret_cast<double2 *>(Ab + (size_t)(row0 + warp * 8 + r) * ld + j0 + sb * 32 + c2) =
                    *reinterpret_cast<const double2 *>(&Rw[r * TP_B + c2]);
        }
        // ---- update the later columns:  acc[:, cb8] -= X[:, sb] * L[cb8 rows, sb cols]^T
        if (sb < TP_NSB - 1) {
            double af[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) af[ks] = -Rw[fr * TP_B + ks * 4 + fk];
#pragma unroll
            for (int cb8 = (sb + 1) * 4; cb8 < TP_NC8; ++cb8) {
                const double *Lq = Lb + lblk_index(cb8 >> 2, sb) * TP_LBLK + ((cb8 & 3) * 8 + fr) * TP_B + fk;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) dmma884_t(acc[cb8][0], acc[cb8][1], af[ks], Lq[ks * 4]);
            }
        }
        __syncwarp();
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
