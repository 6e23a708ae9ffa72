// Panel kernel of the blocked Cholesky: factor one 128x128 diagonal block in shared memory and produce its
// inverse in the same sweep.
//
// Part of kcGP.tools.jitchol -> LAPACK dpotrf(lower=1) (sliceSample.py:196,205): the dpotf2 step on the diagonal
// block, with LAPACK's failure convention (info = 1-based index of the first non-positive / NaN pivot; jitchol
// turns it into a jitter retry).
//
// The block lives in shared memory as ONE 128x132 array T: its lower triangle becomes L11 and its strict upper
// triangle accumulates X = L11^-T, obtained by running the same block operations on an appended identity (the
// rows of [A11; I] are updated alike).  Inner blocking is 32:
//   * the 32x32 diagonal sub-block is factored by ONE WARP, one matrix row (and one identity row) per lane in
//     registers, pivots and multipliers broadcast with warp shuffles -- no block barrier inside;
//   * the sub-panel rows of L below are solved against the sub-block (inverse-multiply + one refinement step, so the
//     factor keeps the backward error of a substitution), the rows of X above are multiplied by its inverse, and the
//     trailing sub-blocks are updated, all with FP64 DMMA (mma.sync.m8n8k4) on fragments read from T.
// W = L11^-1 is written out dense so the panel TRSM below the block is a DMMA GEMM (gemm_dmma.cu).
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int PT = NB + 4;                    // smem row stride (doubles): 4 mod 16 -> conflict-free DMMA fragments
constexpr int PB = 32;                        // inner block
constexpr int PC = 36;                        // row stride of the clean 32x32 inverse
constexpr int POTF2_THREADS = NB == 64 ? 128 : 256;
constexpr int POTF2_WARPS = POTF2_THREADS / 32;
constexpr int POTF2_CTAS = NB == 64 ? 2 : 1;
constexpr int POTF2_SMEM = (NB * PT + PB * PC + NB) * (int)sizeof(double);

__device__ __forceinline__ void dmma884_p(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(POTF2_THREADS, POTF2_CTAS)
potf2_inv_kernel(BatchView A, int n, int j0, double *__restrict__ W, long long strideW, int *__restrict__ info,
                 int zero_upper)
{
    extern __shared__ __align__(16) double sm[];
    double *T = sm;                           // [NB][PT]   lower: L, strict upper: X = L^-T
    double *Lc = sm + NB * PT;                // [PB][PC]   clean inverse of the current 32x32 diagonal sub-block
    double *dinv = Lc + PB * PC;              // [NB]       1 / L_kk
    __shared__ int s_fail;
    const int b = blockIdx.x;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride + (size_t)j0 * A.ld + j0;
    const int ld = A.ld;
    const int nv = min(NB, n - j0);           // valid order of this block
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;  // DMMA fragment coordinates

    // stage the block with 16-byte cp.async pieces (all in flight at once), then fix up in shared memory:
    // strict upper = 0 (the identity's off-diagonal), rows/cols beyond the matrix = identity
    for (int e = tid; e < NB * (NB / 2); e += POTF2_THREADS) {
        const int r = e / (NB / 2), c2 = (e % (NB / 2)) * 2;
        if (r < nv && c2 <= r) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&T[r * PT + c2]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ab + (size_t)r * ld + c2));
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    for (int r = warp; r < NB; r += POTF2_THREADS / 32) {
        for (int c = lane; c < NB; c += 32) {
            if (r >= nv) T[r * PT + c] = (c == r) ? 1.0 : 0.0;
            else if (c > r) T[r * PT + c] = 0.0;
        }
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int kb = 0; kb < NB / PB; ++kb) {
        const int k0 = kb * PB;
        // ---------------------------------------------------------------- 32x32 diagonal sub-block, one warp
        if (warp == 0) {
            double a[PB], x[PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                a[c] = (c <= lane) ? T[(k0 + lane) * PT + k0 + c] : 0.0;
                x[c] = (c == lane) ? 1.0 : 0.0;
            }
            int fail = 0;
            // micro-blocks of 8 columns: inside a micro-block the pivot chain (shuffle -> rsqrt -> scale -> shuffle ->
            // fma) only touches the micro-block's own columns; the remaining columns are updated afterwards in one
            // sweep whose (column, multiplier) pairs are independent, so the critical path is 4 x 8 short steps
#pragma unroll
            for (int q = 0; q < PB / 8; ++q) {
#pragma unroll
                for (int c = q * 8; c < q * 8 + 8; ++c) {
                    const double piv = __shfl_sync(0xffffffffu, a[c], c);
                    if (!(piv > 0.0) && fail == 0) fail = j0 + k0 + c + 1;          // dpotf2: ajj <= 0 or NaN
                    const double rinv = rsqrt(piv);
                    const double d = piv * rinv;
                    a[c] = (lane == c) ? d : a[c] * rinv;                          // column c of L
                    x[c] = x[c] * rinv;                                            // column c of X = L^-T (rows <= c)
#pragma unroll
                    for (int j = c + 1; j < q * 8 + 8; ++j) {
                        const double ljc = __shfl_sync(0xffffffffu, a[c], j);
                        a[j] = fma(-a[c], ljc, a[j]);
                        x[j] = fma(-x[c], ljc, x[j]);
                    }
                }
#pragma unroll
                for (int j = q * 8 + 8; j < PB; ++j) {
#pragma unroll
                    for (int c = q * 8; c < q * 8 + 8; ++c) {
                        const double ljc = __shfl_sync(0xffffffffu, a[c], j);
                        a[j] = fma(-a[c], ljc, a[j]);
                        x[j] = fma(-x[c], ljc, x[j]);
                    }
                }
            }
            // write back: L (lower incl. diagonal) and X (strict upper) into T; clean inverse Lc[c][k] = X[k][c]
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                T[(k0 + lane) * PT + k0 + c] = (c <= lane) ? a[c] : x[c];
                Lc[c * PC + lane] = (lane <= c) ? x[c] : 0.0;                   // row c of L_d^-1, entry k = lane
            }
            double xd = 0.0;
#pragma unroll
            for (int c = 0; c < PB; ++c) xd = (c == lane) ? x[c] : xd;           // x[lane] without dynamic indexing
            dinv[k0 + lane] = xd;
            if (fail != 0 && lane == 0 && s_fail == 0) s_fail = fail;
        }
        __syncthreads();
        // ---------------------------------------------------------------- sub-panel (12 row blocks of 8)
        //   rows below (L):  solve  x L_d^T = a  as inverse-multiply + one step of iterative refinement on DMMA
        //   rows above (X):  T[r][k0 + c] = sum_k T[r][k0 + k] * Lc[c][k]   -- this IS the inverse being built
        {
            const int nblk_below = (NB - k0 - PB) / 8, nblk_above = k0 / 8;
            for (int u = warp; u < nblk_below + nblk_above; u += POTF2_THREADS / 32) {
                if (u < nblk_below) {
                    // X0 = A Lc^T ; r = A - X0 L_d^T ; X = X0 + r Lc^T  (inverse-multiply + one refinement step: backward
                    // error at the level of substitution, no dependent 32-step chain; see trsm_panel.cu)
                    const int r0 = k0 + PB + u * 8;
                    double fa[8], a0[4][2], x0[4][2], rr[4][2];
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) fa[ks] = T[(r0 + fr) * PT + k0 + ks * 4 + fk];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(&T[(r0 + fr) * PT + k0 + q * 8 + 2 * fk]);
                        a0[q][0] = t2.x; a0[q][1] = t2.y;
                        x0[q][0] = x0[q][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks <= 2 * q + 1; ++ks)
                            dmma884_p(x0[q][0], x0[q][1], fa[ks], Lc[(q * 8 + fr) * PC + ks * 4 + fk]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<double2 *>(&T[(r0 + fr) * PT + k0 + q * 8 + 2 * fk]) = make_double2(x0[q][0], x0[q][1]);
                    __syncwarp();
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) fa[ks] = -T[(r0 + fr) * PT + k0 + ks * 4 + fk];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        rr[q][0] = a0[q][0]; rr[q][1] = a0[q][1];
#pragma unroll
                        for (int ks = 0; ks <= 2 * q + 1; ++ks) {
                            // L_d[c][k], k <= c: the diagonal sub-block's lower triangle (its strict upper part holds X_d)
                            const int c = q * 8 + fr, k = ks * 4 + fk;
                            dmma884_p(rr[q][0], rr[q][1], fa[ks], (k <= c) ? T[(k0 + c) * PT + k0 + k] : 0.0);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<double2 *>(&T[(r0 + fr) * PT + k0 + q * 8 + 2 * fk]) = make_double2(rr[q][0], rr[q][1]);
                    __syncwarp();
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) fa[ks] = T[(r0 + fr) * PT + k0 + ks * 4 + fk];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int ks = 0; ks <= 2 * q + 1; ++ks)
                            dmma884_p(x0[q][0], x0[q][1], fa[ks], Lc[(q * 8 + fr) * PC + ks * 4 + fk]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<double2 *>(&T[(r0 + fr) * PT + k0 + q * 8 + 2 * fk]) = make_double2(x0[q][0], x0[q][1]);
                } else {
                    const int r0 = (u - nblk_below) * 8;
                    double af[8];
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) af[ks] = T[(r0 + fr) * PT + k0 + ks * 4 + fk];
                    double out[4][2];
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        out[cb][0] = out[cb][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            dmma884_p(out[cb][0], out[cb][1], af[ks], Lc[(cb * 8 + fr) * PC + ks * 4 + fk]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb)
                        *reinterpret_cast<double2 *>(&T[(r0 + fr) * PT + k0 + cb * 8 + 2 * fk]) = make_double2(out[cb][0], out[cb][1]);
                }
            }
        }
        __syncthreads();
        // ---------------------------------------------------------------- trailing sub-blocks (column blocks cb > kb)
        //   T[r][c] -= sum_k P[r][k] * T[c][k0 + k],   P = the sub-panel just computed (rows of the diagonal
        //   sub-block itself contribute X_d = Lc^T).  Units of 8x8 outputs are dealt round-robin to the warps.
        {
            int u = 0;
            for (int cbk = kb + 1; cbk < NB / PB; ++cbk) {
                const int c0 = cbk * PB;
                const int rows_x = k0 + PB;                   // X part: rows [0, k0+32)
                const int nrb = rows_x / 8 + (NB - c0) / 8;   // + L part: rows [c0, 128)
                for (int rb = 0; rb < nrb; ++rb, ++u) {
                    if ((u % POTF2_WARPS) != warp) continue;
                    const int r0 = (rb < rows_x / 8) ? rb * 8 : c0 + (rb - rows_x / 8) * 8;
                    const bool diag_rows = (r0 >= k0) && (r0 < k0 + PB);
                    double af[8];
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const int k = ks * 4 + fk;
                        af[ks] = -(diag_rows ? Lc[k * PC + (r0 - k0 + fr)] : T[(r0 + fr) * PT + k0 + k]);
                    }
                    double2 cv[4];
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8)
                        cv[c8] = *reinterpret_cast<const double2 *>(&T[(r0 + fr) * PT + c0 + c8 * 8 + 2 * fk]);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8)
                            dmma884_p(cv[c8].x, cv[c8].y, af[ks], T[(c0 + c8 * 8 + fr) * PT + k0 + ks * 4 + fk]);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8)
                        *reinterpret_cast<double2 *>(&T[(r0 + fr) * PT + c0 + c8 * 8 + 2 * fk]) = cv[c8];
                }
            }
        }
        __syncthreads();
    }

    if (tid == 0 && s_fail != 0) {
        if (info[m] == 0) info[m] = s_fail;
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
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    }
    prof_begin(KC_POTF2, s);
    potf2_inv_kernel<<<B, POTF2_THREADS, POTF2_SMEM, s>>>(A, n, j0, W, strideW, info, zero_upper);
    prof_end(KC_POTF2, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
