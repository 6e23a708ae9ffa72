// FP64 tensor-core (DMMA) tile kernel: every dense contraction of the hot path is an "NT" product
//     C[i][j]  (op)=  sum_k  A[i][k] * B[j][k]
// of two row-major operands whose contraction index is contiguous in memory:
//   * Cholesky block-column / trailing update   C -= L[i,:k] L[j,:k]^T          (kcGP.tools.jitchol -> dpotrf,
//   * panel TRSM as a GEMM                      X  = A21 * (L11^-1)^T            sliceSample.py:196,205)
//   * triangular inverse U = L^-T, block column  Y  = U[:i,:i] L[i,:i]^T, then  U[:i,i] = -Y (L_ii^-1)^T
//   * posterior covariance                      R  = S - S (U U^T) S (+1e-11 I)  (sliceSample.py:197-198,205 in the
//                                                                                 algebraically reduced form)
// sm_100a has no FP64 kind in tcgen05; FP64 tensor cores are reached with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// CTA tile 128x128; operands are staged global -> shared with 16-byte cp.async in a 4-stage ring of K=16 chunks;
// shared rows are padded to 20 doubles so the 8-byte fragment loads of a half-warp (4 rows x 4 k) hit 16 distinct
// bank pairs (ncu: 0 shared bank conflicts).  Rows beyond the matrix are zero-filled by cp.async's src-size operand.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int SROW = BK + 4;                    // padded smem row (doubles)
constexpr int OPER_ELEMS = BM * SROW;           // one operand, one stage
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

// Stage one K=16 chunk of a 128-row operand: 128 rows x 8 16-byte pieces = 1024 pieces.
template <int THREADS>
__device__ __forceinline__ void load_operand(double *sdst, const double *gsrc, int ld, int rows_valid, int tid)
{
#pragma unroll
    for (int i = 0; i < 1024 / THREADS; ++i) {
        const int piece = tid + i * THREADS;
        const int row = piece >> 3, kc = piece & 7;
        const bool ok = row < rows_valid;
        const double *src = gsrc + (size_t)(ok ? row : 0) * ld + kc * 2;
        cp_async16(sdst + row * SROW + kc * 2, src, ok ? 16 : 0);
    }
}

// WARPS_M x WARPS_N warps; each warp owns (128/WARPS_M) x (128/WARPS_N) of the tile as 8x8 DMMA fragments.
template <int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, 1)
gemm_dmma_kernel(GemmArgs p)
{
    constexpr int THREADS = WARPS_M * WARPS_N * 32;
    constexpr int FM = BM / WARPS_M / 8;        // fragments per warp along M
    constexpr int FN = BN / WARPS_N / 8;
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.y;
    if (p.C.count && b >= *p.C.count) return;
    const int m = batch_item(p.C, b);

    // tile decode
    int tm, tn;
    if (p.lower_only) {
        const int t = blockIdx.x;
        tm = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
        while (tm * (tm + 1) / 2 > t) --tm;
        tn = t - tm * (tm + 1) / 2;
    } else {
        const int tiles_n = (p.cols + BN - 1) / BN;
        tm = blockIdx.x / tiles_n;
        tn = blockIdx.x - tm * tiles_n;
    }
    const int rows_valid = min(BM, p.rows - tm * BM);
    const int cols_valid = min(BN, p.cols - tn * BN);

    // contraction range of this tile (triangular operands skip their zero part)
    int koff = 0;
    if (p.k_follow_row) koff = max(0, (p.ar0 + tm * BM) - p.k0) & ~(BK - 1);
    const int klen = p.klen - koff;
    const int nk = klen > 0 ? (klen + BK - 1) / BK : 0;

    const double *gA = p.A.base + (size_t)m * p.A.stride + (size_t)(p.ar0 + tm * BM) * p.A.ld + p.k0 + koff;
    const double *gB = p.B.base + (size_t)m * p.B.stride + (size_t)(p.br0 + tn * BN) * p.B.ld + p.bk0 + koff;
    const int lda = p.A.ld, ldb = p.B.ld;
    const int tid = threadIdx.x;

    auto stage_a = [&](int s) { return smem + (size_t)s * 2 * OPER_ELEMS; };
    auto stage_b = [&](int s) { return smem + (size_t)s * 2 * OPER_ELEMS + OPER_ELEMS; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_operand<THREADS>(stage_a(s), gA + s * BK, lda, rows_valid, tid);
            load_operand<THREADS>(stage_b(s), gB + s * BK, ldb, cols_valid, tid);
        }
        cp_async_commit();
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int frow = lane >> 2, fk = lane & 3;      // fragment coordinates

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            const int nxt = kc + STAGES - 1;
            if (nxt < nk) {
                const int s = nxt % STAGES;
                load_operand<THREADS>(stage_a(s), gA + nxt * BK, lda, rows_valid, tid);
                load_operand<THREADS>(stage_b(s), gB + nxt * BK, ldb, cols_valid, tid);
            }
            cp_async_commit();
        }
        const double *sa = stage_a(kc % STAGES) + (wm * FM * 8 + frow) * SROW + fk;
        const double *sb = stage_b(kc % STAGES) + (wn * FN * 8 + frow) * SROW + fk;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) af[i] = sa[i * 8 * SROW + ks * 4];
#pragma unroll
            for (int j = 0; j < FN; ++j) bf[j] = sb[j * 8 * SROW + ks * 4];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns, per fragment, row = lane/4 and two adjacent columns 2*(lane%4)
    double *Cb = p.C.base + (size_t)m * p.C.stride;
    const int ldc = p.C.ld;
    const double *sv = (p.epi == EPI_R) ? p.svec + (size_t)m * p.stride_s : nullptr;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int rl = wm * FM * 8 + i * 8 + frow;
        if (rl >= rows_valid) continue;
        const int gr = p.cr0 + tm * BM + rl;
        double *crow = Cb + (size_t)gr * ldc + p.cc0 + tn * BN;
        const double s_r = sv ? sv[gr] : 0.0;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int cl = wn * FN * 8 + j * 8 + fk * 2;
            if (cl >= cols_valid) continue;
            const bool two = (cl + 1 < cols_valid);
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            if (p.epi == EPI_SUB) {
                if (two) { const double2 c = *reinterpret_cast<const double2 *>(crow + cl); v0 = c.x - v0; v1 = c.y - v1; }
                else v0 = crow[cl] - v0;
            } else if (p.epi == EPI_NEGSET) {
                v0 = -v0; v1 = -v1;
            } else if (p.epi == EPI_R) {
                // R = S - S P S  (+ 1e-11 on the diagonal, sliceSample.py:205)
                const int gc = p.cc0 + tn * BN + cl;
                const double s_c0 = sv[gc], s_c1 = two ? sv[gc + 1] : 0.0;
                v0 = -(s_r * v0 * s_c0);
                v1 = -(s_r * v1 * s_c1);
                if (gr == gc) v0 = (s_r + v0) + 1e-11;
                if (gr == gc + 1) v1 = (s_r + v1) + 1e-11;
            }
            if (two) *reinterpret_cast<double2 *>(crow + cl) = make_double2(v0, v1);
            else crow[cl] = v0;
        }
    }
}

static int g_gemm_cfg = 1;      // 0: 8 warps (2x4, warp tile 64x32); 1: 16 warps (4x4, warp tile 32x32)
void set_gemm_config(int cfg) { g_gemm_cfg = cfg; }

int launch_gemm(const GemmArgs &a, int B, int kclass, cudaStream_t s)
{
    if (B <= 0 || a.rows <= 0 || a.cols <= 0) return 0;
    if ((a.A.ld & 1) || (a.B.ld & 1) || (a.C.ld & 1) || (a.k0 & 1) || (a.bk0 & 1) || (a.cc0 & 1)) {
        set_error("gemm: k0=%d bk0=%d cc0=%d lda=%d ldb=%d ldc=%d violate the 16-byte staging alignment",
                  a.k0, a.bk0, a.cc0, a.A.ld, a.B.ld, a.C.ld);
        return GPMC_EALIGN;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        attr_set = true;
    }
    const int tiles_m = (a.rows + BM - 1) / BM;
    const int tiles_n = (a.cols + BN - 1) / BN;
    const int tiles = a.lower_only ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    dim3 grid(tiles, B);
    prof_begin(kclass, s);
    if (g_gemm_cfg == 0) gemm_dmma_kernel<2, 4><<<grid, 256, GEMM_SMEM, s>>>(a);
    else gemm_dmma_kernel<4, 4><<<grid, 512, GEMM_SMEM, s>>>(a);
    prof_end(kclass, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
