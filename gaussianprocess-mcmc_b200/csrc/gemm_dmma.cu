// FP64 tensor-core (DMMA) tile kernel of the blocked Cholesky.
//
// One kernel serves both dense contractions of the factorisation of A = L L^T (row-major, lower):
//   mode 0 (update):  C[i][j] -= sum_k L[i][k0+k] * L[j][k0+k]   the SYRK/GEMM update of a block column
//                                                                 (left-looking) or of the trailing matrix
//                                                                 (right-looking), "NT" form: both operands
//                                                                 are rows of the same matrix, K contiguous.
//   mode 1 (panel):   X[i][c] = sum_k A[i][c0+k] * W[c][k]        the panel TRSM  X = A21 * L11^-T  written as
//                                                                 a GEMM with W = L11^-1 from the potf2 kernel;
//                                                                 in place (a CTA owns whole rows of the panel).
// This is the O(N^3) part of kcGP.tools.jitchol -> LAPACK dpotrf at sliceSample.py:196,205.
//
// sm_100a has no FP64 kind in tcgen05; FP64 tensor cores are reached with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// CTA tile 128x128, 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA fragments (64 accumulator doubles per
// thread).  Operands are staged global -> shared with 16-byte cp.async in a 4-stage ring of K=16 chunks;
// shared rows are padded to 20 doubles so the 8-byte fragment loads of a half-warp (4 rows x 4 k) hit 16
// distinct bank pairs.  Rows beyond the matrix are zero-filled by cp.async's src-size operand.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int SROW = BK + 4;                    // padded smem row (doubles)
constexpr int OPER_ELEMS = BM * SROW;           // one operand, one stage
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = STAGES * 2 * OPER_ELEMS * (int)sizeof(double);   // 81,920 B

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stage one K=16 chunk of a 128-row operand: 128 rows x 8 16-byte pieces = 1024 pieces, 4 per thread.
__device__ __forceinline__ void load_operand(double *sdst, const double *gsrc, int ld, int rows_valid, int tid)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int piece = tid + i * GEMM_THREADS;
        const int row = piece >> 3, kc = piece & 7;
        const bool ok = row < rows_valid;
        const double *src = gsrc + (size_t)(ok ? row : 0) * ld + kc * 2;
        cp_async16(sdst + row * SROW + kc * 2, src, ok ? 16 : 0);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_dmma_kernel(GemmArgs p)
{
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.y;
    if (p.A.count && b >= *p.A.count) return;
    const int m = batch_item(p.A, b);

    // tile decode
    const int tiles_n = (p.cols + BN - 1) / BN;
    int tm, tn;
    if (p.lower_only) {
        // output rows and cols start at the same diagonal position: enumerate tm >= tn
        const int t = blockIdx.x;
        tm = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
        while (tm * (tm + 1) / 2 > t) --tm;
        tn = t - tm * (tm + 1) / 2;
    } else {
        tm = blockIdx.x / tiles_n;
        tn = blockIdx.x - tm * tiles_n;
    }
    const int row0 = p.r0 + tm * BM;            // first output row (global)
    const int col0 = p.c0 + tn * BN;            // first output col (global)
    const int rows_valid = min(BM, p.r0 + p.rows - row0);
    const int cols_valid = min(BN, p.c0 + p.cols - col0);

    double *Ab = p.A.base + (size_t)m * p.A.stride;
    const int ld = p.A.ld;
    const double *gA;       // operand with the output rows, K contiguous
    const double *gB;       // operand with the output cols, K contiguous
    int ldb, b_rows_valid;
    if (p.mode == 0) {
        gA = Ab + (size_t)row0 * ld + p.k0;
        gB = Ab + (size_t)col0 * ld + p.k0;
        ldb = ld;
        b_rows_valid = cols_valid;
    } else {
        gA = Ab + (size_t)row0 * ld + p.c0;
        gB = p.W + (size_t)m * p.strideW;
        ldb = NB;
        b_rows_valid = BN;
    }
    const int nk = p.klen / BK;
    const int tid = threadIdx.x;

    auto stage_a = [&](int s) { return smem + (size_t)s * 2 * OPER_ELEMS; };
    auto stage_b = [&](int s) { return smem + (size_t)s * 2 * OPER_ELEMS + OPER_ELEMS; };

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_operand(stage_a(s), gA + s * BK, ld, rows_valid, tid);
            load_operand(stage_b(s), gB + s * BK, ldb, b_rows_valid, tid);
        }
        cp_async_commit();
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;        // 2 x 4 warps
    const int frow = lane >> 2, fk = lane & 3;      // fragment coordinates

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            const int nxt = kc + STAGES - 1;
            if (nxt < nk) {
                const int s = nxt % STAGES;
                load_operand(stage_a(s), gA + nxt * BK, ld, rows_valid, tid);
                load_operand(stage_b(s), gB + nxt * BK, ldb, b_rows_valid, tid);
            }
            cp_async_commit();
        }
        const double *sa = stage_a(kc % STAGES) + (wm * 64 + frow) * SROW + fk;
        const double *sb = stage_b(kc % STAGES) + (wn * 32 + frow) * SROW + fk;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = sa[i * 8 * SROW + ks * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sb[j * 8 * SROW + ks * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns, per fragment, row = lane/4 and two adjacent columns 2*(lane%4)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = wm * 64 + i * 8 + frow;
        if (rl >= rows_valid) continue;
        double *crow = Ab + (size_t)(row0 + rl) * ld + col0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cl = wn * 32 + j * 8 + fk * 2;
            if (cl >= cols_valid) continue;
            double2 *dst = reinterpret_cast<double2 *>(crow + cl);
            if (cl + 1 < cols_valid) {
                if (p.mode == 0) {
                    double2 c = *dst;
                    c.x -= acc[i][j][0];
                    c.y -= acc[i][j][1];
                    *dst = c;
                } else {
                    *dst = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            } else {
                if (p.mode == 0) crow[cl] -= acc[i][j][0]; else crow[cl] = acc[i][j][0];
            }
        }
    }
}

int launch_gemm(const GemmArgs &a, int B, cudaStream_t s)
{
    if (B <= 0 || a.rows <= 0 || a.cols <= 0) return 0;
    if (a.klen % BK != 0 || (a.A.ld & 1) || (a.k0 & 1) || (a.c0 & 1)) {
        set_error("gemm: klen=%d k0=%d c0=%d ld=%d violate the 16-byte staging alignment", a.klen, a.k0, a.c0, a.A.ld);
        return GPMC_EALIGN;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        attr_set = true;
    }
    const int tiles_m = (a.rows + BM - 1) / BM;
    const int tiles_n = (a.cols + BN - 1) / BN;
    const int tiles = a.lower_only ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    dim3 grid(tiles, B);
    const int kc = a.mode == 0 ? KC_GEMM : KC_TRSM;
    prof_begin(kc, s);
    gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(a);
    prof_end(kc, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
