// Panel triangular solve of the blocked Cholesky:  X L11^T = A21  (in place), i.e. the rows of L below a
// factored 128x128 diagonal block -- the dtrsm inside LAPACK dpotrf (kcGP.tools.jitchol, sliceSample.py:196,205).
//
// Each 32-column sub-block is solved as  X0 = A W_d^T,  r = A - X0 L_d^T,  X = X0 + r W_d^T  (W_d = L_d^-1 from the
// panel kernel): an inverse-multiply followed by ONE step of iterative refinement, all on FP64 DMMA.  The refinement
// is what makes it safe: chol(R + 1e-11 I) at sliceSample.py:205 has condition ~1e11 and a plain  A21 * inv(L11)
// (128-wide, what batched GPU Cholesky codes commonly do) loses cond(L11) * eps -- measured: spurious "not positive
// definite" pivots on a golden fixture.  With the residual taken against the 32x32 block L_d itself the backward error
// is back at eps * |X| |L_d| (the level of substitution) as long as cond(L_d) * eps << 1, and the dependent chain of a
// substitution (128 steps of shuffle -> multiply -> fma per row) is gone.
//
// One CTA owns 128 rows of the panel; each of its 16 warps keeps 8 rows x 128 columns as DMMA accumulator fragments
// in registers and never meets a block barrier after the initial load.  Per sub-block: three small DMMA products (the
// triangular zero parts are skipped), then the later columns get  acc[:, later] -= X[:, sb] * L[later, sb]^T.
// Fragment-layout changes (accumulator -> A operand) go through the warp's own shared-memory rows.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int TP_ROWS = 128;                  // rows per CTA
constexpr int TP_B = 36;                      // stride of L / W blocks and of the staged sub-block (4 mod 16: conflict-free
                                              // DMMA fragments)
constexpr int TP_THREADS = 544;               // 16 warps x 8 rows + one warp for a border row that follows a full last CTA
constexpr int TP_SROWS = TP_ROWS + 8;         // staging rows
constexpr int TP_LBLK = 32 * TP_B;            // doubles per 32x32 L block
constexpr int TP_NSB = NB / 32;               // 32-column sub-blocks of the panel
constexpr int TP_NLB = TP_NSB * (TP_NSB + 1) / 2;   // lower 32x32 blocks of L11
constexpr int TP_NC8 = NB / 8;                // 8-column fragments per row
constexpr int TP_SMEM = (TP_SROWS * TP_B + (TP_NLB + TP_NSB) * TP_LBLK) * (int)sizeof(double);

__device__ __forceinline__ int lblk_index(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }     // bi >= bj

__device__ __forceinline__ int stage_rot(int row) { return ((0x0310 >> ((row & 3) * 4)) & 0xf) * 4; }

__device__ __forceinline__ void dmma884_t(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(TP_THREADS, 1)
trsm_panel_kernel(BatchView A, int n_rows, int j0, const double *__restrict__ W, long long strideW)
{
    extern __shared__ __align__(16) double sm[];
    double *R = sm;                               // [128][36] the sub-block being solved
    double *Lb = sm + TP_SROWS * TP_B;            // 10 lower blocks of L11
    double *Wd = Lb + TP_NLB * TP_LBLK;           // the NB/32 diagonal 32x32 blocks of W = L11^-1
    const int b = blockIdx.y;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride;
    const int ld = A.ld;
    const int row0 = j0 + NB + blockIdx.x * TP_ROWS;
    // n_rows counts the border rows too; the last CTA takes every row that is left (at most 128 + 1 border row)
    const int rows_valid = (blockIdx.x == gridDim.x - 1) ? min(TP_SROWS, n_rows - row0) : TP_ROWS;
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
    const double *Wb = W + (size_t)m * strideW;
    for (int e = tid; e < TP_NSB * 32 * 16; e += TP_THREADS) {
        const int d = e / (32 * 16), rem = e - d * 32 * 16;
        const int r = rem / 16, c2 = (rem - r * 16) * 2;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&Wd[d * TP_LBLK + r * TP_B + c2]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Wb + (size_t)(d * 32 + r) * NB + d * 32 + c2));
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
    // the strict upper triangle of the diagonal L blocks is garbage in global memory: the residual product needs zeros
    for (int e = tid; e < TP_NSB * 32 * 32; e += TP_THREADS) {
        const int d = e >> 10, r = (e >> 5) & 31, c = e & 31;
        if (c > r) Lb[lblk_index(d, d) * TP_LBLK + r * TP_B + c] = 0.0;
    }
    __syncthreads();
    // From here on every warp works on its own 8 rows only (L and W are read-only): no block barriers.
    if (warp * 8 >= rows_valid) return;           // ragged last CTA (e.g. only a border row): nothing to solve
    double *Rw = R + warp * 8 * TP_B;             // this warp's staging rows
    // staging rows are rotated by {0, 4, 12, 0}[row & 3] columns: with stride 36 alone the 16-byte fragment stores of rows
    // fr and fr+1 share bank groups (2-way conflict, ncu: 10 M conflicts per launch); with the rotation both the 16-byte
    // stores (quarter-warp: rows fr, fr+1) and the 8-byte fragment loads (half-warp: rows fr..fr+3) are conflict free
    const int rot = stage_rot(fr);
#pragma unroll
    for (int sb = 0; sb < TP_NSB; ++sb) {
        const double *Ld = Lb + lblk_index(sb, sb) * TP_LBLK;
        const double *Wq = Wd + sb * TP_LBLK;
        // ---- X0 = A W_d^T : accumulator fragments -> A fragments through the warp's shared rows
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2 *>(&Rw[fr * TP_B + ((q * 8 + 2 * fk + rot) & 31)]) = make_double2(acc[sb * 4 + q][0], acc[sb * 4 + q][1]);
        __syncwarp();
        double fa[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) fa[ks] = Rw[fr * TP_B + ((ks * 4 + fk + rot) & 31)];
        double x0[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            x0[q][0] = x0[q][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks <= 2 * q + 1; ++ks)                       // W_d[c][k] = 0 for k > c
                dmma884_t(x0[q][0], x0[q][1], fa[ks], Wq[(q * 8 + fr) * TP_B + ks * 4 + fk]);
        }
        __syncwarp();
        // ---- r = A - X0 L_d^T
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2 *>(&Rw[fr * TP_B + ((q * 8 + 2 * fk + rot) & 31)]) = make_double2(x0[q][0], x0[q][1]);
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) fa[ks] = -Rw[fr * TP_B + ((ks * 4 + fk + rot) & 31)];
        double rr[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            rr[q][0] = acc[sb * 4 + q][0]; rr[q][1] = acc[sb * 4 + q][1];
#pragma unroll
            for (int ks = 0; ks <= 2 * q + 1; ++ks)                       // L_d[c][k] = 0 for k > c
                dmma884_t(rr[q][0], rr[q][1], fa[ks], Ld[(q * 8 + fr) * TP_B + ks * 4 + fk]);
        }
        __syncwarp();
        // ---- X = X0 + r W_d^T   (one step of iterative refinement)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2 *>(&Rw[fr * TP_B + ((q * 8 + 2 * fk + rot) & 31)]) = make_double2(rr[q][0], rr[q][1]);
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) fa[ks] = Rw[fr * TP_B + ((ks * 4 + fk + rot) & 31)];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int ks = 0; ks <= 2 * q + 1; ++ks)
                dmma884_t(x0[q][0], x0[q][1], fa[ks], Wq[(q * 8 + fr) * TP_B + ks * 4 + fk]);
            acc[sb * 4 + q][0] = x0[q][0]; acc[sb * 4 + q][1] = x0[q][1];
        }
        __syncwarp();
        // ---- stage the solved columns (this warp's rows): final values -> global, and A fragments for the update
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2 *>(&Rw[fr * TP_B + ((q * 8 + 2 * fk + rot) & 31)]) = make_double2(acc[sb * 4 + q][0], acc[sb * 4 + q][1]);
        __syncwarp();
        for (int e = lane; e < 8 * 16; e += 32) {
            const int r = e >> 4, c2 = (e & 15) * 2;
            if (warp * 8 + r < rows_valid)
                *reinterpret_cast<double2 *>(Ab + (size_t)(row0 + warp * 8 + r) * ld + j0 + sb * 32 + c2) =
                    *reinterpret_cast<const double2 *>(&Rw[r * TP_B + ((c2 + stage_rot(r)) & 31)]);
        }
        // ---- update the later columns:  acc[:, cb8] -= X[:, sb] * L[cb8 rows, sb cols]^T
        if (sb < TP_NSB - 1) {
            double af[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) af[ks] = -Rw[fr * TP_B + ((ks * 4 + fk + rot) & 31)];
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

// n_rows = matrix order + border rows (0 or 1); a border row that follows a full last CTA is taken by that CTA's 17th warp
int launch_trsm_panel(BatchView A, int n_rows, int j0, const double *W, long long strideW, int B, cudaStream_t s)
{
    const int n = n_rows;
    int rows = n_rows - j0 - NB;
    if (B <= 0 || rows <= 0) return 0;
    if (rows > 1 && rows % TP_ROWS == 1) rows -= 1;         // 128 k + 1: the extra row rides in the last CTA
    if ((A.ld & 1) || (j0 & 1)) { set_error("trsm_panel: ld=%d j0=%d must be even", A.ld, j0); return GPMC_EALIGN; }
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
    }
    dim3 grid((rows + TP_ROWS - 1) / TP_ROWS, B);
    prof_begin(KC_TRSM, s);
    trsm_panel_kernel<<<grid, TP_THREADS, TP_SMEM, s>>>(A, n, j0, W, strideW);
    prof_end(KC_TRSM, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
