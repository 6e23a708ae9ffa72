// Shared declarations of libgpmc's kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace gpmc {

constexpr int NB = 128;            // panel width / tile edge of the blocked Cholesky
constexpr int MAX_ELL = 8;         // max number of length-scales (ARD input dimension)

// kernel classes for the profiling hooks (gpmc_profile_read)
enum KernelClass { KC_ASSEMBLE = 0, KC_GEMM = 1, KC_POTF2 = 2, KC_TRSM = 3, KC_SOLVE = 4, KC_COUNT = 5 };

void set_error(const char *fmt, ...);
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

// gemm_dmma.cu
//   mode 0: C[r0+i][c0+j] -= sum_k A[r0+i][k0+k] * A[c0+j][k0+k]      (left-looking / trailing update)
//   mode 1: C[r0+i][c0+j]  = sum_k A[r0+i][c0+k] * W[j][k]            (panel TRSM as GEMM with W = L11^-1)
struct GemmArgs {
    BatchView A;           // the matrix being factorised
    const double *W;       // mode 1: per-item [NB][NB] inverse of the diagonal block (row-major, dense)
    long long strideW;
    int n;                 // matrix order (rows/cols valid)
    int r0, rows;          // output rows [r0, r0+rows)
    int c0, cols;          // output cols [c0, c0+cols)   (cols <= NB in mode 1)
    int k0, klen;          // contraction range (multiple of 16)
    int lower_only;        // skip tiles strictly above the diagonal (r-tile < c-tile), SYRK use
    int mode;
};
int launch_gemm(const GemmArgs &a, int B, cudaStream_t s);

// potf2.cu : factor the NB x NB diagonal block at (j0, j0) in place, write its inverse to W
int launch_potf2(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper,
                 int B, cudaStream_t s);

// solve_reduce.cu : z = L^-1 g, loglik = -(0.5 z.z + sum log L_ii + 0.5 n log 2pi)
int launch_solve_reduce(BatchView L, int n, const double *g, int ldg, double *loglik, const int *info,
                        double *zbuf, int B, cudaStream_t s);

// microbench.cu
int run_fp64_peak(int which, int iters, double *tflops, double *ms);

}  // namespace gpmc
