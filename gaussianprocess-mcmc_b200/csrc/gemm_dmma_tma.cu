// TMA-staged variant of the FP64 DMMA tile kernel (same contraction and epilogues as gemm_dmma.cu).
//
// Operand tiles are fetched by the Tensor Memory Accelerator: one elected thread issues
// cp.async.bulk.tensor.3d (box = 16 doubles x {128|64} rows x 1 matrix) per operand and stage, completion is
// signalled on an mbarrier (expect_tx / complete_tx), the 256 consumer threads wait on the barrier's phase bit.
// Rows beyond the operand's extent are zero-filled by the TMA unit itself (tensor-map bounds), so there is no
// per-thread predication, address arithmetic or cp.async issue slot in the main loop.
//
// Shared-memory layout: dense 128-byte rows (16 doubles) with the hardware 128B swizzle: the 16-byte piece p of
// row r lands at piece p ^ (r & 7).  A DMMA m8n8k4 fragment wants, per half-warp, 4 rows x 4 k: to keep that
// conflict-free under the swizzle the kernel permutes the contraction index inside a K=16 chunk (legal: both
// operands use the same permutation): step s of a chunk uses k in {2s, 2s+1, 8+2s, 9+2s}, lane (row, fk) reads
// piece ((fk>>1)*4 + s) ^ row, half (fk & 1) -> the 16 lanes of a half-warp touch 16 distinct 8-byte slots.
#include "common.cuh"
#include "gemm_common.cuh"
#include "../../include/gpmc.h"

#include <cuda.h>
#include <type_traits>
#include <cudaTypedefs.h>

namespace gpmc {

constexpr int TBM = 128, TBN = 64, TBK = 16;
constexpr int TSTAGES = 4;
constexpr int TTHREADS = 256;
constexpr int A_BYTES = TBM * TBK * 8;                 // 16 KiB
constexpr int B_BYTES = TBN * TBK * 8;                 //  8 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMA_SMEM = TSTAGES * STAGE_BYTES + 1024 + 128;  // + alignment slack + barriers

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    // bounded: a TMA that never completes (bad descriptor) must not hang the GPU
    for (unsigned it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// FREE_RUNNING = false: one __syncthreads per K chunk recycles the stages (all warps in lock step).
// FREE_RUNNING = true : full/empty mbarrier pairs per stage -- every warp arrives on empty[s] when it has read stage s,
//                       the producer thread refills a stage once its empty barrier completed, and no block barrier
//                       is left in the main loop, so the eight warps drift apart instead of stalling together.
template <bool FREE_RUNNING>
__global__ void __launch_bounds__(TTHREADS, 2)
gemm_dmma_tma_kernel(GemmArgs p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB)
{
    constexpr int WARPS_N = 2, FM = 4, FN = 4;          // 4 x 2 warps, warp tile 32 x 32
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024 B
    uint64_t *full = (uint64_t *)(smem + TSTAGES * STAGE_BYTES);
    uint64_t *empty = full + TSTAGES;
    const int b = blockIdx.y;
    if (p.C.count && b >= *p.C.count) return;
    const int m = batch_item(p.C, b);
    int tm, tn;
    gemm_tile_decode<TBM, TBN>(p, blockIdx.x, tm, tn);
    if (p.skip_upper && p.cr0 + tm * TBM + TBM - 1 < p.cc0 + tn * TBN) return;     // whole tile above the diagonal
    const int rows_valid = min(TBM, p.rows - tm * TBM);
    const int cols_valid = min(TBN, p.cols - tn * TBN);
    int koff = 0;
    if (p.k_follow_row) koff = max(0, (p.ar0 + tm * TBM) - p.k0) & ~(TBK - 1);
    const int klen = p.klen - koff;
    const int nk = klen > 0 ? (klen + TBK - 1) / TBK : 0;
    const int tid = threadIdx.x;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA));
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB));
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TTHREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();

    const int arow = p.ar0 + tm * TBM, brow = p.br0 + tn * TBN;
    const int ak = p.k0 + koff, bk = p.bk0 + koff;
    auto issue = [&](int chunk) {
        const int s = chunk % TSTAGES;
        unsigned char *st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(st, &tmA, &full[s], ak + chunk * TBK, arow, m);
        tma_load_3d(st + A_BYTES, &tmB, &full[s], bk + chunk * TBK, brow, m);
    };
    if (tid == 0)      // lock step: one stage stays free for the refill; free running: all stages start full
        for (int s = 0; s < (FREE_RUNNING ? TSTAGES : TSTAGES - 1) && s < nk; ++s) issue(s);

    // the epilogue subtracts from the C tile: ask L2 for its 512 lines now (two per thread), so that the loads at the
    // end of the K loop do not wait for HBM -- matters for the short contractions only (K = 128 parts of the look-ahead,
    // N = 512 matrices: +1 ... 1.5 %; long contractions hide the epilogue behind the co-resident CTA anyway)
    if (p.epi == EPI_SUB && klen <= 512) {
        const double *Cb = p.C.base + (size_t)m * p.C.stride + (size_t)(p.cr0 + tm * TBM) * p.C.ld + p.cc0 + tn * TBN;
#pragma unroll
        for (int e = tid; e < TBM * (TBN / 16); e += TTHREADS) {
            const int r = e / (TBN / 16), c = (e % (TBN / 16)) * 16;
            if (r < rows_valid && c < cols_valid) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Cb + (size_t)r * p.C.ld + c));
        }
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int frow = lane >> 2, fk = lane & 3;
    // swizzled fragment offsets (bytes) inside a tile row: piece ((fk>>1)*4 + s) ^ frow, half (fk & 1)
    int foff[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) foff[s] = ((((fk >> 1) * 4 + s) ^ frow) << 4) + ((fk & 1) << 3);

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // diagonal tiles of a symmetric update: a warp whose 32x32 sub-tile is strictly above the diagonal has nothing to
    // contribute (6 of the 16 sub-tiles of a 128x128 diagonal block) -- it only keeps the stage hand-shake going and
    // leaves its DMMA issue slots to the other resident warps
    // ... and so has a warp whose rows all lie beyond the output (ragged last row tile, e.g. a single border row)
    const bool skip = (wm * FM * 8 >= rows_valid) ||
                      (p.skip_upper && (p.cr0 + tm * TBM + (wm + 1) * FM * 8 - 1 < p.cc0 + tn * TBN + wn * FN * 8));
    const bool on_diag = p.skip_upper && (p.cr0 + tm * TBM + wm * FM * 8 == p.cc0 + tn * TBN + wn * FN * 8);
    if (FREE_RUNNING && skip) {
        // Border duty: in a tile on the diagonal the warp (wm, wn) = (0, 1) is always one of the skippers.  It carries
        // the right-hand side stored as row p.border_row through the same update for the tile's 64 columns:
        // one real row in the 8-row A fragment (read straight from global memory, one chunk ahead), B fragments
        // from the staged tile -- 8 DMMAs per k step instead of none.
        const bool duty = p.border_row > 0 && wm == 0 && wn == 1 && (p.cr0 + tm * TBM == p.cc0 + (tn * TBN) / TBM * TBM);
        const double *zrow = p.A.base + (size_t)m * p.A.stride + (size_t)p.border_row * p.A.ld + ak;
        const int kq = (fk >> 1) * 8 + (fk & 1);        // this lane's k inside a chunk at step ks: kq + 2 ks (the permuted order)
        double bacc[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) bacc[j][0] = bacc[j][1] = 0.0;
        double zn[4] = {0.0, 0.0, 0.0, 0.0};
        auto zload = [&](int chunk) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int k = chunk * TBK + kq + 2 * ks;
                zn[ks] = (frow == 0 && k < klen) ? zrow[k] : 0.0;
            }
        };
        if (duty && nk > 0) zload(0);
        for (int kc = 0; kc < nk; ++kc) {
            const int s = kc % TSTAGES;
            double zc[4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) zc[ks] = zn[ks];
            if (duty && kc + 1 < nk) zload(kc + 1);
            mbar_wait(&full[s], (kc / TSTAGES) & 1);
            if (duty) {
                const unsigned char *sb = smem + s * STAGE_BYTES + A_BYTES + frow * 128;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dmma884(bacc[j][0], bacc[j][1], zc[ks], *reinterpret_cast<const double *>(sb + j * 8 * 128 + foff[ks]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (tid == 0 && kc >= 1 && kc + TSTAGES - 1 < nk) {
                const int prev = kc - 1;
                mbar_wait(&empty[prev % TSTAGES], (prev / TSTAGES) & 1);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                issue(kc + TSTAGES - 1);
            }
        }
        if (duty && frow == 0) {
            double *crow = p.C.base + (size_t)m * p.C.stride + (size_t)p.border_row * p.C.ld + p.cc0 + tn * TBN;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cl = j * 8 + fk * 2;
                if (cl >= cols_valid) continue;
                if (cl + 1 < cols_valid) {
                    double2 c = *reinterpret_cast<const double2 *>(crow + cl);
                    c.x -= bacc[j][0]; c.y -= bacc[j][1];
                    *reinterpret_cast<double2 *>(crow + cl) = c;
                } else crow[cl] -= bacc[j][0];
            }
        }
        return;
    }

    // the K loop, in two compiled shapes: every 8x8 block of the warp's sub-tile, or (sub-tile ON the diagonal of a
    // skip_upper update) only the blocks with j <= i -- chosen once per warp, outside the loop
    auto main_loop = [&](auto diag_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value;
        for (int kc = 0; kc < nk; ++kc) {
            const int s = kc % TSTAGES;
            mbar_wait(&full[s], (kc / TSTAGES) & 1);
            if (!FREE_RUNNING) {
                __syncthreads();                              // everyone is done with the stage refilled below
                if (tid == 0 && kc + TSTAGES - 1 < nk) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    issue(kc + TSTAGES - 1);
                }
            }
            const unsigned char *sa = smem + s * STAGE_BYTES + (wm * FM * 8 + frow) * 128;
            const unsigned char *sb = smem + s * STAGE_BYTES + A_BYTES + (wn * FN * 8 + frow) * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double af[FM], bf[FN];
#pragma unroll
                for (int i = 0; i < FM; ++i) af[i] = *reinterpret_cast<const double *>(sa + i * 8 * 128 + foff[ks]);
#pragma unroll
                for (int j = 0; j < FN; ++j) bf[j] = *reinterpret_cast<const double *>(sb + j * 8 * 128 + foff[ks]);
                if (FREE_RUNNING && ks == 3) {
                    // the fragments of this stage are in registers: release it, and let the producer refill the stage
                    // that every warp released one chunk ago
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                    if (tid == 0 && kc >= 1 && kc + TSTAGES - 1 < nk) {
                        const int prev = kc - 1;
                        mbar_wait(&empty[prev % TSTAGES], (prev / TSTAGES) & 1);
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        issue(kc + TSTAGES - 1);
                    }
                }
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j)
                        if (!DIAG || j <= i) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
    };
    // lock step keeps every warp on one path (block barriers inside the loop)
    if (FREE_RUNNING && on_diag) main_loop(std::true_type{});
    else main_loop(std::false_type{});
    gemm_epilogue<FM, FN, TBM, TBN>(p, m, tm, tn, wm, wn, frow, fk, rows_valid, cols_valid, acc);
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encoder()
{
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
        return GPMC_EINVAL;
    }
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return 0;
}

// 3-D map over matrices stored as [batch][row][ld]: dims (inner -> outer) = {ld, rows_extent, nbatch}
static int make_map(CUtensorMap *map, const Operand &op, int rows_extent, int box_rows)
{
    const cuuint64_t dims[3] = {(cuuint64_t)op.ld, (cuuint64_t)rows_extent, (cuuint64_t)1 << 20};
    const cuuint64_t strides[2] = {(cuuint64_t)op.ld * 8, (cuuint64_t)op.stride * 8};
    const cuuint32_t box[3] = {(cuuint32_t)TBK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)op.base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): base=%p ld=%d stride=%lld rows=%d", (int)r, (const void *)op.base, op.ld,
                  op.stride, rows_extent);
        return GPMC_EINVAL;
    }
    return 0;
}

bool gemm_tma_supported(const GemmArgs &a)
{
    // TMA needs 16-byte aligned bases and strides; a zero batch stride (shared operand) is not mapped
    return a.A.stride > 0 && a.B.stride > 0 && ((uintptr_t)a.A.base % 16 == 0) && ((uintptr_t)a.B.base % 16 == 0) &&
           (a.A.ld % 2 == 0) && (a.B.ld % 2 == 0) && (a.A.stride % 2 == 0) && (a.B.stride % 2 == 0) &&
           a.A.ld >= TBK && a.B.ld >= TBK;
}

int launch_gemm_tma(const GemmArgs &a, int B, int kclass, bool free_running, cudaStream_t s)
{
    int rc = get_encoder();
    if (rc) return rc;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
    }
    CUtensorMap tmA, tmB;
    // rows at or beyond (origin + extent) are out of bounds for the TMA unit -> zero fill
    if ((rc = make_map(&tmA, a.A, a.ar0 + a.rows, TBM))) return rc;
    if ((rc = make_map(&tmB, a.B, a.br0 + a.cols, TBN))) return rc;
    const int tiles_m = (a.rows + TBM - 1) / TBM;
    const int tiles_n = (a.cols + TBN - 1) / TBN;
    const int tiles = a.lower_only ? (TBM / TBN) * tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    dim3 grid(tiles, B);
    prof_begin(kclass, s);
    if (free_running) gemm_dmma_tma_kernel<true><<<grid, TTHREADS, TMA_SMEM, s>>>(a, tmA, tmB);
    else gemm_dmma_tma_kernel<false><<<grid, TTHREADS, TMA_SMEM, s>>>(a, tmA, tmB);
    prof_end(kclass, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
