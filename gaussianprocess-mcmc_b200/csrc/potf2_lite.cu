// Panel kernel without the block inverse: factor one 128x128 diagonal block and emit only the inverses of its sixteen
// 8x8 diagonal sub-blocks (all that trsm_panel8.cu needs).  Used wherever the caller does not go on to invert the
// factor: the log-likelihood unit (sliceSample.py:196,147) and chol(R + 1e-11 I) (sliceSample.py:205).
//
// Same scheme as potf2.cu -- ONE warp factors each 32x32 diagonal sub-block with a matrix row per lane in registers and
// warp-shuffle broadcasts of pivots and multipliers -- with three differences:
//  * the identity rows that ride along (the inverse) are updated only inside their own 8-column micro-block: the 8x8
//    diagonal inverses are all anyone needs, and the full 32x32 inverse costs as many DFMAs as the factor itself;
//  * sub-panel rows are solved and trailing sub-blocks are updated on FP64 DMMA with the contraction index permuted
//    (k-step 0 takes k = 2 fk, k-step 1 takes k = 2 fk + 1): accumulator pairs are the next A operands and every
//    operand is one 16-byte shared-memory access -- no fragment-layout changes, no warp synchronisation; the solve is
//    the 8-column inverse-multiply + one refinement step of trsm_panel8.cu;
//  * only the lower triangle is kept in shared memory (block rows of growing length: 90 KB instead of 135 KB) and the
//    CTA has 4 warps, so TWO CTAs share an SM: the kernel is latency bound (one CTA per matrix), and a second resident
//    CTA nearly doubles the batched throughput.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int QB = 32;                        // inner block
constexpr int LITE_THREADS = 128;
constexpr int LITE_WARPS = LITE_THREADS / 32;
constexpr int LITE_NBLK = NB / QB;            // block rows (4 for a 128 panel)
// lower-triangular storage: block row bi keeps 32*(bi+1) columns, row stride 32*(bi+1) + 8 (8 mod 16: conflict-free
// 16-byte DMMA fragment accesses); doubles before block row bi = 512*bi*(bi+1) + 256*bi
constexpr int LITE_T_ELEMS = 512 * LITE_NBLK * (LITE_NBLK + 1) + 256 * LITE_NBLK;
constexpr int LITE_W8_ELEMS = (QB / 8) * 8 * 8;      // the four 8x8 diagonal inverses of the current 32x32 sub-block
constexpr int LITE_SMEM = (LITE_T_ELEMS + LITE_W8_ELEMS) * (int)sizeof(double);

__device__ __forceinline__ int tix(int r, int c)
{
    const int bi = r >> 5;
    return 512 * bi * (bi + 1) + 256 * bi + (r & 31) * (32 * (bi + 1) + 8) + c;
}

__device__ __forceinline__ void dmma884_l(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(LITE_THREADS, 2)
potf2_lite_kernel(BatchView A, int n, int j0, double *__restrict__ W, long long strideW, int *__restrict__ info, int zero_upper)
{
    extern __shared__ __align__(16) double sm[];
    double *T = sm;                           // lower triangle, block rows
    double *W8 = sm + LITE_T_ELEMS;           // [4][8][8] diagonal 8x8 inverses of the current sub-block, [blk][c][k]
    __shared__ int s_fail;
    const int b = blockIdx.x;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride + (size_t)j0 * A.ld + j0;
    const int ld = A.ld;
    const int nv = min(NB, n - j0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;

    // stage the lower blocks with 16-byte cp.async pieces, then fix up: strict upper of the diagonal sub-blocks = 0,
    // rows/cols beyond the matrix = identity
    for (int e = tid; e < NB * (NB / 2); e += LITE_THREADS) {
        const int r = e / (NB / 2), c2 = (e % (NB / 2)) * 2;
        if (r < nv && c2 <= r) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&T[tix(r, c2)]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ab + (size_t)r * ld + c2));
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    for (int r = warp; r < NB; r += LITE_WARPS) {
        const int cend = 32 * ((r >> 5) + 1);
        for (int c = lane; c < cend; c += 32) {
            if (r >= nv) T[tix(r, c)] = (c == r) ? 1.0 : 0.0;
            else if (c > r) T[tix(r, c)] = 0.0;
        }
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();

    double *Wb = W + (size_t)m * strideW;
#pragma unroll 1
    for (int kb = 0; kb < LITE_NBLK; ++kb) {
        const int k0 = kb * QB;
        // ---------------------------------------------------------------- 32x32 diagonal sub-block, one warp
        if (warp == 0) {
            double a[QB], x[QB];
            {
                const double *row = &T[tix(k0 + lane, k0)];
#pragma unroll
                for (int c = 0; c < QB; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + c);
                    a[c] = (c <= lane) ? v.x : 0.0;
                    a[c + 1] = (c + 1 <= lane) ? v.y : 0.0;
                }
            }
#pragma unroll
            for (int c = 0; c < QB; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
            int fail = 0;
#pragma unroll
            for (int q = 0; q < QB / 8; ++q) {
#pragma unroll
                for (int c = q * 8; c < q * 8 + 8; ++c) {
                    const double piv = __shfl_sync(0xffffffffu, a[c], c);
                    if (!(piv > 0.0) && fail == 0) fail = j0 + k0 + c + 1;          // dpotf2: ajj <= 0 or NaN
                    const double rinv = rsqrt(piv);
                    const double d = piv * rinv;
                    a[c] = (lane == c) ? d : a[c] * rinv;
                    x[c] = x[c] * rinv;
#pragma unroll
                    for (int j = c + 1; j < q * 8 + 8; ++j) {
                        const double ljc = __shfl_sync(0xffffffffu, a[c], j);
                        a[j] = fma(-a[c], ljc, a[j]);
                        x[j] = fma(-x[c], ljc, x[j]);          // stays inside the micro-block: the 8x8 diagonal inverse
                    }
                }
#pragma unroll
                for (int j = q * 8 + 8; j < QB; ++j) {
#pragma unroll
                    for (int c = q * 8; c < q * 8 + 8; ++c) {
                        const double ljc = __shfl_sync(0xffffffffu, a[c], j);
                        a[j] = fma(-a[c], ljc, a[j]);
                    }
                }
            }
            // L_d (lower, zeros above) back into T; the 8x8 diagonal inverses W8[blk][c][k] = x_k[c] (k <= c, same
            // micro-block) to shared memory and to the diagonal of W in global memory
            {
                double *row = &T[tix(k0 + lane, k0)];
#pragma unroll
                for (int c = 0; c < QB; c += 2)
                    *reinterpret_cast<double2 *>(row + c) = make_double2((c <= lane) ? a[c] : 0.0, (c + 1 <= lane) ? a[c + 1] : 0.0);
            }
            const int qb = lane >> 3, kl = lane & 7;
#pragma unroll
            for (int c = 0; c < QB; ++c) {
                if ((c >> 3) == qb) {
                    const double w = (kl <= (c & 7)) ? x[c] : 0.0;
                    W8[(qb * 8 + (c & 7)) * 8 + kl] = w;
                    Wb[(k0 + c) * NB + k0 + qb * 8 + kl] = w;
                }
            }
            if (fail != 0 && lane == 0 && s_fail == 0) s_fail = fail;
        }
        __syncthreads();
        // ---------------------------------------------------------------- sub-panel rows below, 8 rows x 32 columns
        //   per unit: per 8-column block  X0 = A W8^T, r = A - X0 L8^T, X = X0 + r W8^T, later blocks -= X L^T
        {
            const int nunits = (NB - k0 - QB) / 8;
            for (int u = warp; u < nunits; u += LITE_WARPS) {
                double *row = &T[tix(k0 + QB + u * 8 + fr, k0 + 2 * fk)];
                double acc[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + q * 8);
                    acc[q][0] = v.x; acc[q][1] = v.y;
                }
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const double2 w = *reinterpret_cast<const double2 *>(&W8[(bb * 8 + fr) * 8 + 2 * fk]);
                    const double2 l = *reinterpret_cast<const double2 *>(&T[tix(k0 + bb * 8 + fr, k0 + bb * 8 + 2 * fk)]);
                    double x0 = 0.0, x1 = 0.0;
                    dmma884_l(x0, x1, acc[bb][0], w.x);
                    dmma884_l(x0, x1, acc[bb][1], w.y);
                    double r0 = acc[bb][0], r1 = acc[bb][1];
                    dmma884_l(r0, r1, -x0, l.x);
                    dmma884_l(r0, r1, -x1, l.y);
                    dmma884_l(x0, x1, r0, w.x);
                    dmma884_l(x0, x1, r1, w.y);
                    acc[bb][0] = x0; acc[bb][1] = x1;
#pragma unroll
                    for (int q = bb + 1; q < 4; ++q) {
                        const double2 lp = *reinterpret_cast<const double2 *>(&T[tix(k0 + q * 8 + fr, k0 + bb * 8 + 2 * fk)]);
                        dmma884_l(acc[q][0], acc[q][1], -x0, lp.x);
                        dmma884_l(acc[q][0], acc[q][1], -x1, lp.y);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<double2 *>(row + q * 8) = make_double2(acc[q][0], acc[q][1]);
            }
        }
        __syncthreads();
        // ---------------------------------------------------------------- trailing sub-blocks: for column blocks
        //   cbk > kb and rows r >= 32*cbk:  T[r][c] -= sum_k T[r][k0 + k] * T[c][k0 + k]
        {
            int u = 0;
            for (int cbk = kb + 1; cbk < LITE_NBLK; ++cbk) {
                const int c0 = cbk * QB;
                for (int rb = 0; rb < (NB - c0) / 8; ++rb, ++u) {
                    if ((u % LITE_WARPS) != warp) continue;
                    const int r0 = c0 + rb * 8;
                    double2 af[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        af[s] = *reinterpret_cast<const double2 *>(&T[tix(r0 + fr, k0 + s * 8 + 2 * fk)]);
                        af[s].x = -af[s].x; af[s].y = -af[s].y;
                    }
                    double *crow = &T[tix(r0 + fr, c0 + 2 * fk)];
                    double2 cv[4];
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) cv[c8] = *reinterpret_cast<const double2 *>(crow + c8 * 8);
#pragma unroll
                    for (int s = 0; s < 4; ++s)
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            const double2 bf = *reinterpret_cast<const double2 *>(&T[tix(c0 + c8 * 8 + fr, k0 + s * 8 + 2 * fk)]);
                            dmma884_l(cv[c8].x, cv[c8].y, af[s].x, bf.x);
                            dmma884_l(cv[c8].x, cv[c8].y, af[s].y, bf.y);
                        }
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<double2 *>(crow + c8 * 8) = cv[c8];
                }
            }
        }
        __syncthreads();
    }

    if (tid == 0 && s_fail != 0) {
        if (info[m] == 0) info[m] = s_fail;
    }
    // write L11 back (lower incl. diagonal); optionally zero the strict upper triangle
    for (int r = warp; r < nv; r += LITE_WARPS) {
        for (int c = lane; c < nv; c += 32) {
            if (c <= r) Ab[(size_t)r * ld + c] = T[tix(r, c)];
            else if (zero_upper) Ab[(size_t)r * ld + c] = 0.0;
        }
    }
}

int launch_potf2_lite(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper, int B, cudaStream_t s)
{
    if (B <= 0) return 0;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_lite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LITE_SMEM));
    }
    prof_begin(KC_POTF2, s);
    potf2_lite_kernel<<<B, LITE_THREADS, LITE_SMEM, s>>>(A, n, j0, W, strideW, info, zero_upper);
    prof_end(KC_POTF2, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
