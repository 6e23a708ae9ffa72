// Shared declarations of libgpmc's kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace gpmc {

#ifndef GPMC_NB
#define GPMC_NB 128
#endif
constexpr int NB = GPMC_NB;        // panel width of the blocked Cholesky (64 or 128)
static_assert(NB == 64 || NB == 128, "panel width must be 64 or 128");
constexpr int MAX_ELL = 8;         // max number of length-scales (ARD input dimension)
constexpr int MAX_BATCH_ITEMS = 32768;   // batch items of one launch ride in grid.y / grid.z (limit 65535)

// kernel classes for the profiling hooks (gpmc_profile_read)
enum KernelClass { KC_ASSEMBLE = 0, KC_GEMM = 1, KC_POTF2 = 2, KC_TRSM = 3, KC_SOLVE = 4, KC_INV = 5, KC_SYRK_R = 6, KC_VEC = 7, KC_COUNT = 8 };

void set_error(const char *fmt, ...);
// Every computing entry point of the C ABI takes this (recursive) lock: the library keeps per-process state (side streams,
// staging buffers, tuning switches), so calls from several host threads are SERIALISED rather than racing.
struct ApiLock { ApiLock(); ~ApiLock(); };
#define GPMC_API_LOCK() gpmc::ApiLock _gpmc_api_lock
void prof_begin(int kc, cudaStream_t s);
void prof_end(int kc, cudaStream_t s);

#define GPMC_CUDA_CHECK(expr)                                                                 \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            gpmc::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                            cudaGetErrorString(_e));                                          \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

#define GPMC_LAUNCH_CHECK()                                                                   \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            gpmc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                            cudaGetErrorString(_e));                                          \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

// Per-device "do this once" latch: function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) belong to the
// device that was current when they were set, so a process that drives several GPUs must set them once PER DEVICE.
struct DeviceOnce {
    unsigned long long mask[4] = {0, 0, 0, 0};
    bool first()
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return true;
        unsigned long long &w = mask[(dev >> 6) & 3];
        const unsigned long long bit = 1ull << (dev & 63);
        if (w & bit) return false;
        w |= bit;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Batched matrix view.  Item b of a launch is matrix `map ? map[b] : b` of the allocation, so a
// compacted list of active chains / failed items can be processed without moving data.
struct BatchView {
    double *base;          // first matrix
    long long stride;      // elements between matrices
    int ld;                // leading dimension (even; multiple of 16 for internal buffers)
    const int *map;        // optional indirection (device), may be nullptr
    const int *count;      // optional device-side item count; CTAs with blockIdx.y >= *count exit
};

__device__ __forceinline__ int batch_item(const BatchView &v, int b) { return v.map ? v.map[b] : b; }

// ---------------------------------------------------------------------------------------------
// launchers (definitions in the .cu files)

// cov_assemble.cu
int launch_cov_assemble(const double *x, int N, int D, const double *hyp, int P, int n_ell, int flags,
                        const double *jitter, BatchView A, int B, cudaStream_t s);

// gemm_dmma.cu :  C[i][j] (op)= sum_k A[i][k] * B[j][k]   ("NT": both operands K-contiguous, row-major)
struct Operand {
    const double *base;    // first matrix
    long long stride;      // elements between matrices of consecutive batch items
    int ld;
};
enum Epilogue {
    EPI_SUB = 0,           // C = C - acc            (Cholesky update)
    EPI_SET = 1,           // C = acc                (panel TRSM via inverse; Y = U L^T)
    EPI_NEGSET = 2,        // C = -acc               (block column of U = L^-T)
    EPI_R = 3,             // C = S - S acc S (+1e-11 on the diagonal), S = diag(svec)   (posterior covariance R)
    EPI_ADD = 4            // C = C + acc            (windowed triangular inverse: Y accumulated window by window)
};
struct GemmArgs {
    BatchView C;           // output matrices; its map/count select the batch items of the launch
    Operand A, B;          // A rows <-> output rows, B rows <-> output cols (indexed by the same mapped item)
    int cr0, cc0;          // output origin (row, col) in C
    int rows, cols;        // output extent
    int ar0, br0;          // operand rows corresponding to cr0 / cc0
    int k0, bk0;           // first contraction column in A / in B
    int klen;              // contraction length (the tail beyond a multiple of 16 must be zero-padded in memory)
    int lower_only;        // enumerate only tiles with tile-row >= tile-col (square output, SYRK use)
    int k_follow_row;      // A (and B) are upper triangular in (row, k): start k at the tile's first A row
    int epi;               // Epilogue
    int skip_upper;        // C's strict upper triangle (global row < global col) is never read by the caller: warps whose
                           // whole sub-tile lies there may skip their contraction and leave C untouched
    int border_row;        // > 0 (EPI_SUB, skip_upper, cr0 == cc0, A == C buffers): global row of a right-hand side carried as a
                           // ROW below the matrix; it receives the same update, C[border_row, cols] -= A[border_row, k] B[cols, k]^T
    const double *svec;    // EPI_R: per-item diagonal of S, [N]
    long long stride_s;
};
int launch_gemm(const GemmArgs &a, int B, int kclass, cudaStream_t s);
void set_gemm_config(int cfg);

// potf2.cu : factor the NB x NB diagonal block at (j0, j0) in place, write its inverse to W
int launch_potf2(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                 int B, cudaStream_t s);
// potf2_lite.cu : same factor, but only the inverses of the sixteen 8x8 diagonal sub-blocks are written to W (all that
// launch_trsm_panel8 reads); 2 CTAs per SM.  Not for callers that go on to inverse_sequence or use launch_trsm_panel.
int launch_potf2_lite(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                      int B, cudaStream_t s);
// potf2_reg.cu : same contract as launch_potf2_lite; the block lives in registers as DMMA fragments over 8 warps and is
// factored in 16 steps of 8 columns (the default panel factor kernel)
int launch_potf2_reg(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                     int B, cudaStream_t s);
// potf2_flow.cu : the register-resident panel factor kernel as a dataflow program (flags in shared memory instead of block
// barriers between the 16 steps; measured slower than potf2_reg.cu, kept for A/B: gpmc_set_tuning(1, 3)), and its fused form
int launch_potf2_flow(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                      int B, cudaStream_t s);
int launch_panel_fused_flow(BatchView A, int n, int n_rows, int j0, double *W, long long strideW, int *info, int zero_upper,
                            int B, cudaStream_t s);
// potf2_reg.cu : panel factor AND panel solve of block column j0 in one launch (n_rows = n + border rows; in the last block
// column the border rows are solved too) -- for many small matrices in flight
int launch_panel_fused(BatchView A, int n, int n_rows, int j0, double *W, long long strideW, int *info, int zero_upper,
                       int B, cudaStream_t s);
void set_lookahead_mode(int mode); // potrf_sequence: 0 auto (few matrices in flight), 1 off, 2 on
void set_potrf_window(int w);      // override the window of the windowed schedule (multiple of NB; 0 = default)
void set_potf2_mode(int mode);     // 0: register-resident kernel (default), 3: the same as a dataflow program (no block barriers),
                                   // 2: the shared-memory lite kernel, 1: always the full-inverse one

// trsm_panel.cu : rows below the factored diagonal block at (j0, j0):  X L11^T = A21  (W = L11^-1 from launch_potf2)
int launch_trsm_panel(BatchView A, int n, int j0, const double *W, long long strideW, int B, cudaStream_t s);
// trsm_panel8.cu : the same solve with 8-column sub-blocks, shuffle-based fragment conversion, 2 CTAs per SM (default)
//   row_start >= 0: solve only rows row_start .. n-1 (border rows of the last block column); n_mat >= 0: order of the matrix
//   (rows / columns of the diagonal block at or beyond it read as identity)
int launch_trsm_panel8(BatchView A, int n, int j0, const double *W, long long strideW, int B, cudaStream_t s,
                       int row_start = -1, int n_mat = -1);
// trmm_panel8.cu : in place  A[0:rows, c0:c0+width] = -(A[0:rows, c0:c0+128] W^T)  for a lower triangular 128x128 W
//   (the second product of the triangular inverse, inverse_sequence)
int launch_trmm_panel8(BatchView A, int rows, int c0, int width, const double *W, long long strideW, int B, cudaStream_t s);
// W_i = L_ii^-1 for every diagonal block of a factored matrix, from the 8x8 diagonal inverses potf2_lite left in W
//   (W: nt blocks of NB x NB per item, item stride strideW)
int launch_inv_blocks8(BatchView A, int n, double *W, long long strideW, int B, cudaStream_t s);
void set_trsm_mode(int mode);      // 0: trsm_panel8 (default), 1: trsm_panel (32-column sub-blocks)
void set_trsm_blocks_per_cta(int v);   // 0: auto (1, 2 or 4 by launch size), else forced

// solve_reduce.cu : z = L^-1 (a - b), optional outputs: z, loglik = -(0.5 z.z + sum log L_ii + 0.5 n log 2pi)
//   a, b, zout are per-item vectors with row stride ldv (b, zout, loglik may be nullptr)
int launch_solve_reduce(BatchView L, int n, const double *a, const double *b, int ldv, double *zout,
                        double *loglik, const int *info, int B, cudaStream_t s);
// loglik[m] = -(0.5 z.z + sum_i log L_ii + 0.5 n log 2pi) from a finished z (row stride ldv per item)
int launch_quad_logdet(BatchView L, int n, const double *z, int ldv, double *loglik, const int *info, int B, cudaStream_t s);
// trmv.cu : out = T x (+ add) for a triangular row-major T;  upper=0: lower triangle, upper=1: upper triangle.
//   mode 0: out = T x + add            (f' = C eta + m, sliceSample.py:140)
//   mode 1: out = add - svec * (T x)   (m = g - S (K+S)^-1 g with T = L^-T, x = L^-1 g; sliceSample.py:204)
//   mode 2: out = T x                  (nu = chol(K) z, the draw of elliptical_slice; sliceSample.py:41)
int launch_trmv(BatchView T, int n, int upper, int mode, const double *x, const double *add, const double *svec,
                int ldv, double *out, int B, cudaStream_t s);

// sweep.cu
void set_sds_mode(int mode);       // 0: resident loop (default), 1: wave loop
void set_panel_fuse(int mode);
void set_lookahead_split(int mode);
void set_inverse_window(int w);     // fused panel factor + solve launches: 0 auto (many small matrices), 1 never, 2 whenever possible
void set_sds_literal(int v);       // 1: R = K - V^T V as the reference writes it (parity), 0: reduced form (default)
void set_sds_runahead(int r);      // rounds queued ahead of the last status word seen (0: auto)
void set_sds_pin_schedule(int v);  // 1: factorisation schedules of the resident loop chosen from a constant (bit-reproducible runs)
void sds_loop_stats(long long *rounds, long long *idle_rounds, long long *ladders);

// microbench.cu
int run_fp64_peak(int which, int iters, double *tflops, double *ms);
int run_dmma_ilp(int nacc, int warps_per_sm, int iters, double *tflops);

}  // namespace gpmc
