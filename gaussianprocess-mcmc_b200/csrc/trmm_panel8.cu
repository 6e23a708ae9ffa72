// Panel multiply of the triangular inverse:  U[0:i0, i] = -(Y W_i^T)  in place, Y = U[0:i0, 0:i0] L[i, 0:i0]^T already
// sitting in the block column, W_i = L_ii^-1 (128 x 128 lower triangular, from the panel factor kernel).
//
// Same skeleton as trsm_panel8.cu: a warp owns 8 rows x 128 columns.  With the contraction index of m8n8k4 permuted
// (k-step 0 takes k = 2 fk, k-step 1 takes k = 2 fk + 1) the two doubles a lane LOADS of an 8-wide block of Y (row fr,
// columns 2 fk, 2 fk + 1 -- one 16-byte load) are its A-operand fragments, the matching B fragments of W_i are adjacent
// in shared memory, and the accumulator pair of an output block is what the lane stores (one 16-byte store).  Every
// warp reads its rows completely before it writes them, so the product is safely in place; the zero blocks of the
// triangular W_i are skipped (272 DMMAs per 8 rows x 128 columns, half of the dense product); two CTAs of 64 rows per SM.
#include "common.cuh"
#include "../../include/gpmc.h"

#include <algorithm>

namespace gpmc {

constexpr int M8_ROWS = 64;
constexpr int M8_THREADS = M8_ROWS / 8 * 32;
constexpr int M8_B = 40;                      // row stride of a 32x32 W block (8 mod 16: conflict-free 16-byte fragment loads)
constexpr int M8_LBLK = 32 * M8_B;
constexpr int M8_NSB = NB / 32;
constexpr int M8_NLB = M8_NSB * (M8_NSB + 1) / 2;
constexpr int M8_NB8 = NB / 8;
constexpr int M8_SMEM = M8_NLB * M8_LBLK * (int)sizeof(double);

__device__ __forceinline__ void dmma884_m(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Y: rows [0, rows) x columns [c0, c0 + 128) of every item; only columns below `cols_ok` (relative) exist in memory
__global__ void __launch_bounds__(M8_THREADS, 2)
trmm_panel8_kernel(BatchView A, int rows, int c0, int cols_ok, const double *__restrict__ W, long long strideW)
{
    extern __shared__ __align__(16) double Wb[];      // 10 lower 32x32 blocks of W_i
    const int b = blockIdx.y;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);
    double *Ab = A.base + (size_t)m * A.stride;
    const int ld = A.ld;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;
    const double *Wg = W + (size_t)m * strideW;
    for (int e = tid; e < M8_NLB * 32 * 16; e += M8_THREADS) {        // 16 16-byte pieces per block row
        const int blk = e / (32 * 16), rem = e - blk * 32 * 16;
        const int r = rem / 16, c2 = (rem - r * 16) * 2;
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
        const int bj = blk - bi * (bi + 1) / 2;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&Wb[blk * M8_LBLK + r * M8_B + c2]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Wg + (size_t)(bi * 32 + r) * NB + bj * 32 + c2));
    }
    asm volatile("cp.async.commit_group;\n" ::);
    const int r_glob = blockIdx.x * M8_ROWS + warp * 8 + fr;
    const bool live = r_glob < rows;
    double *grow = Ab + (size_t)(live ? r_glob : 0) * ld + c0 + 2 * fk;
    double y[M8_NB8][2];
#pragma unroll
    for (int b8 = 0; b8 < M8_NB8; ++b8) {
        double2 v = make_double2(0.0, 0.0);
        if (live && b8 * 8 + 2 * fk < cols_ok) v = *reinterpret_cast<const double2 *>(grow + b8 * 8);
        y[b8][0] = v.x;
        y[b8][1] = v.y;
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (blockIdx.x * M8_ROWS + warp * 8 >= rows) return;
    // out[:, b] = - sum_{kb <= b} Y[:, kb] W[b][kb]^T     (W[b][kb] = 0 for kb > b; the dense W from the panel kernel
    // carries explicit zeros above its diagonal, so diagonal blocks need no masking)
#pragma unroll
    for (int b8 = 0; b8 < M8_NB8; ++b8) {
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int kb = 0; kb <= b8; ++kb) {
            const int bi = b8 >> 2, bj = kb >> 2;
            const double2 w = *reinterpret_cast<const double2 *>(Wb + (bi * (bi + 1) / 2 + bj) * M8_LBLK + ((b8 & 3) * 8 + fr) * M8_B + (kb & 3) * 8 + 2 * fk);
            dmma884_m(o0, o1, y[kb][0], w.x);
            dmma884_m(o0, o1, y[kb][1], w.y);
        }
        if (live && b8 * 8 + 2 * fk < cols_ok) *reinterpret_cast<double2 *>(grow + b8 * 8) = make_double2(-o0, -o1);
    }
}

int launch_trmm_panel8(BatchView A, int rows, int c0, int width, const double *W, long long strideW, int B, cudaStream_t s)
{
    if (B <= 0 || rows <= 0) return 0;
    if ((A.ld & 1) || (c0 & 1)) { set_error("trmm_panel: ld=%d c0=%d must be even", A.ld, c0); return GPMC_EALIGN; }
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(trmm_panel8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M8_SMEM));
    }
    // column pairs starting at or beyond the padded width do not exist in memory (the last block column of a ragged N);
    // the pad columns inside ld are zero and W_i is identity padded, so they are rewritten with zeros
    const int cols_ok = std::min(NB, std::min((width + 1) & ~1, A.ld - c0));
    dim3 grid((rows + M8_ROWS - 1) / M8_ROWS, B);
    prof_begin(KC_INV, s);
    trmm_panel8_kernel<<<grid, M8_THREADS, M8_SMEM, s>>>(A, rows, c0, cols_ok, W, strideW);
    prof_end(KC_INV, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
