// Host-side sequencing of the batched kernels (definitions in sequences.cu).
#pragma once
#include "common.cuh"

namespace gpmc {

// Left-looking blocked Cholesky of every item of A (lower, in place).  The inverse of diagonal block j is
// written to W + item*strideW + j*w_step (w_step = 0: one scratch block per item, overwritten every step;
// w_step = NB*NB: all blocks kept, as inverse_sequence needs them).
int potrf_sequence(BatchView A, int n, int B, int *info, double *W, long long strideW, long long w_step,
                   int zero_upper, cudaStream_t s);

// In place L (lower) -> U = L^-T (upper triangle incl. diagonal blocks; the strict lower block part keeps L).
// Needs the saved diagonal-block inverses of potrf_sequence (w_step = NB*NB).
int inverse_sequence(BatchView A, int n, int B, const double *W, long long strideW, cudaStream_t s);

// R = S - S (U U^T) S + 1e-11 I, lower tiles only, from U (upper, in A) into Rm.  svec = diag(S) per item.
int r_sequence(BatchView Rm, BatchView U, int n, int B, const double *svec, long long stride_s, cudaStream_t s);

// small helpers
int fill_int(int *p, int v, int n, cudaStream_t s);
int fill_int_mapped(int *p, int v, const int *map, const int *count, int nmax, cudaStream_t s);
int add_diag(BatchView A, int n, const double *jitter, int B, cudaStream_t s);
int diag_stats(BatchView A, int n, double *mean_out, int *nonpos_out, int B, cudaStream_t s);
int add_diag_vec(BatchView A, int n, const double *v, int ldv, int B, cudaStream_t s);
int zero_upper(BatchView A, int n, int B, cudaStream_t s);
int copy_rows(double *dst, int ldd, const double *src, int lds, int n, int rows, cudaStream_t s);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ld_for(int n) { return (n + 15) / 16 * 16; }

}  // namespace gpmc
