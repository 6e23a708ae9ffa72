// Row-wise triangular solve against a 32x32 lower block held in shared memory (backward-stable substitution).
#pragma once

namespace gpmc {

// x <- x * Ld^-T, i.e. solve  x Ld^T = a  for one row a of 32 entries kept in registers.
// Ld: shared, lower triangular incl. diagonal, row stride lds; dinv[c] = 1 / Ld[c][c].
// Every thread of a warp reads the same Ld entry (broadcast), the row itself never leaves registers.
__device__ __forceinline__ void row_trsv32(double (&x)[32], const double *Ld, int lds, const double *dinv)
{
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        x[c] = x[c] * dinv[c];
#pragma unroll
        for (int j = c + 1; j < 32; ++j) x[j] = fma(-x[c], Ld[j * lds + c], x[j]);
    }
}

}  // namespace gpmc
