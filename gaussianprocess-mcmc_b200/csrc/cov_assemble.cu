// Fused covariance assembly: SE (iso / ARD) kernel + S diagonal + jitter + symmetric fill.
//
// Replaces, per batch item, the reference's
//   Kc = covK.RBF(np.log(hyp[0]), np.log(hyp[1])); K = Kc.getCovMatrix(x=x, mode='train')   sliceSample.py:104-105,136-137
//   S_ii = 1/((1/sn^2 + 1/K_ii) - 1/K_ii); S = max(S, 0); K+S                               sliceSample.py:183-190,196
// (pyGPs 1.3.4 cov.RBF: K = sf2 * exp(-0.5 * cdist(x/ell, x/ell, 'sqeuclidean')), ell = exp(log_ell),
//  sf2 = exp(2 log_sigma)); the operation order of those expressions is kept so the matrix matches
//  numpy's to the last bits of exp().
//
// Layout: one CTA = one 64x64 tile on or below the diagonal.  Each lane owns two adjacent columns
// (double2, 16 B stores; a warp writes 512 contiguous bytes of one row), the tile is mirrored
// through shared memory so the transposed tile is written with the same coalesced double2 stores
// and every exp() is evaluated once.  HBM-write bound: 8*N^2 bytes per matrix (8*N*(N+64)/2 with
// GPMC_ASM_LOWER_ONLY).
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

constexpr int AT = 64;          // tile edge
constexpr int AT_PAD = 66;      // smem row stride (doubles): keeps double2 stores 16 B aligned
constexpr int ASM_TPC = 1;      // tiles per CTA (4 was measured: N=4096 x 256 4.17 -> 4.69 ms, slower -- the per-CTA scalar prologue is not the limiter)

// DT > 0: input dimension known at compile time (the distance loop unrolls and the column values live in registers);
// DT == 0: generic D.  The kernel is instruction-issue bound (ncu: 79 % issue slots, FP64 pipe 47 %, DRAM 33 %), so the
// fast path strips everything that is not the exp itself: no bounds tests on interior tiles, no diagonal tests off
// the diagonal.
// PRED: the K / sn^2 + I variant of inf_mcmc (GPMC_ASM_PRED) -- a template parameter, not a run-time test: a uniform
// `if (pred)` in the fast path cost the hot (non-PRED) instantiation 20 % (3.5 vs 4.2 TB/s at N=4096, round 2).
template <int DT, bool PRED>
__global__ void __launch_bounds__(256)
cov_assemble_kernel(const double *__restrict__ x, int N, int Drt, const double *__restrict__ hyp, int P, int n_ell,
                    int flags, const double *__restrict__ jitter, BatchView A)
{
    const int D = DT > 0 ? DT : Drt;
    const int b = blockIdx.y;
    if (A.count && b >= *A.count) return;
    const int m = batch_item(A, b);

    __shared__ double s_ell[MAX_ELL];
    __shared__ double s_ui[MAX_ELL][AT];
    __shared__ double s_uj[MAX_ELL][AT];
    __shared__ __align__(16) double s_tile[AT][AT_PAD];
    __shared__ double s_scal[4];       // sf2, Sii (or 1 with GPMC_ASM_PRED), jitter, sn2 (GPMC_ASM_PRED)

    const double *h = hyp + (size_t)m * P;
    const int tid = threadIdx.x;
    if (tid < n_ell) {
        // covK.RBF(np.log(ll), ...): ell = exp(log(ll))   (log/exp round trip kept)
        s_ell[tid] = exp(log(h[tid]));
    }
    if (tid == 32) {
        const double sf = h[n_ell], sn = h[n_ell + 1];
        const double sf2 = exp(2.0 * log(sf));                 // sf2 = exp(2*log_sigma)
        const double Kii = sf2;                                // sf2 * exp(-0.5*0) == sf2
        // sliceSample.py:185-187,190
        const double K_ii_inv = 1.0 / Kii;
        const double v_1 = 1.0 / (sn * sn) + K_ii_inv;
        double Sii = 1.0 / (v_1 - K_ii_inv);
        Sii = (Sii < 0.0) ? 0.0 : Sii;                         // np.maximum(S, 0) (NaN propagates)
        s_scal[0] = sf2;
        s_scal[1] = (flags & GPMC_ASM_ADD_S) ? Sii : 0.0;
        s_scal[2] = jitter ? jitter[m] : 0.0;
        if (PRED) {
            // inf_mcmc (sliceSample.py:256-257): sn2 = likfunc.sn**2 with likfunc.sn = exp(log_sigma); K/sn2 + eye(n)
            const double snl = exp(log(sn));
            s_scal[3] = snl * snl;
            s_scal[1] = 1.0;
        } else s_scal[3] = 1.0;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const double sf2 = s_scal[0];
    const double diag_add = s_scal[1], jit = s_scal[2], sn2 = s_scal[3];
    constexpr bool pred = PRED;
    const bool has_jit = (jitter != nullptr);
    double *Ab = A.base + (size_t)m * A.stride;
    const int cl = 2 * lane;
    const int nt1 = (N + AT - 1) / AT, ntiles = nt1 * (nt1 + 1) / 2;
    // A CTA takes ASM_TPC consecutive tiles of its matrix (experiment: amortise the per-item scalars above -- two
    // exp(log(.)) round trips and the S_ii expression on one or two threads while the rest of the CTA waits; ncu r02f shows
    // barrier stalls of 2.6 per issue -- over several tiles.  Measured slower with 4, so the default is 1.)
    for (int tt = 0; tt < ASM_TPC; ++tt) {
    // tile decode: t -> (tm >= tn)
    const int t = blockIdx.x * ASM_TPC + tt;
    if (t >= ntiles) break;
    int tm = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
    while (tm * (tm + 1) / 2 > t) --tm;
    const int tn = t - tm * (tm + 1) / 2;
    const int r0 = tm * AT, c0 = tn * AT;
    if (tt > 0) __syncthreads();                 // the previous tile's u vectors / mirrored tile are still being read
    // u = x / ell for the tile's rows and columns
    for (int e = tid; e < 2 * AT * D; e += 256) {
        const int which = e / (AT * D);
        const int rem = e - which * AT * D;
        const int i = rem / D, d = rem - i * D;
        const int gi = (which ? c0 : r0) + i;
        const double ell = s_ell[n_ell == 1 ? 0 : d];
        const double v = (gi < N) ? x[(size_t)gi * D + d] / ell : 0.0;
        if (which) s_uj[d][i] = v; else s_ui[d][i] = v;
    }
    __syncthreads();

    const int gc = c0 + cl;
    const bool mirror = (tm != tn) && !(flags & GPMC_ASM_LOWER_ONLY);

    const bool interior = (r0 + AT <= N) && (c0 + AT <= N) && (tm != tn);
    if (DT > 0 && interior) {
        // fast path: full tile strictly below the diagonal
        double uj0[DT > 0 ? DT : 1], uj1[DT > 0 ? DT : 1];
#pragma unroll
        for (int d = 0; d < DT; ++d) { uj0[d] = s_uj[d][cl]; uj1[d] = s_uj[d][cl + 1]; }
        double *dst = Ab + (size_t)(r0 + warp * 8) * A.ld + gc;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const int rl = warp * 8 + rr;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int d = 0; d < DT; ++d) {
                const double ui = s_ui[d][rl];
                const double d0 = ui - uj0[d], d1 = ui - uj1[d];
                s0 = __dadd_rn(s0, __dmul_rn(d0, d0));
                s1 = __dadd_rn(s1, __dmul_rn(d1, d1));
            }
            double k0 = sf2 * exp(-0.5 * s0);
            double k1 = sf2 * exp(-0.5 * s1);
            if (pred) { k0 = k0 / sn2; k1 = k1 / sn2; }
            if (mirror) *reinterpret_cast<double2 *>(&s_tile[rl][cl]) = make_double2(k0, k1);
            *reinterpret_cast<double2 *>(dst + (size_t)rr * A.ld) = make_double2(k0, k1);
        }
    } else {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int rl = warp * 8 + rr;
        const int gr = r0 + rl;
        double s0 = 0.0, s1 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double ui = s_ui[d][rl];
            const double d0 = ui - s_uj[d][cl];
            const double d1 = ui - s_uj[d][cl + 1];
            s0 = __dadd_rn(s0, __dmul_rn(d0, d0));            // cdist 'sqeuclidean': s += d*d, no FMA
            s1 = __dadd_rn(s1, __dmul_rn(d1, d1));
        }
        double k0 = sf2 * exp(-0.5 * s0);
        double k1 = sf2 * exp(-0.5 * s1);
        if (pred) { k0 = k0 / sn2; k1 = k1 / sn2; }
        if (gr == gc)     { k0 = k0 + diag_add; if (has_jit) k0 = k0 + jit; }
        if (gr == gc + 1) { k1 = k1 + diag_add; if (has_jit) k1 = k1 + jit; }
        if (mirror) *reinterpret_cast<double2 *>(&s_tile[rl][cl]) = make_double2(k0, k1);
        if (gr < N) {
            double *dst = Ab + (size_t)gr * A.ld + gc;
            if (gc + 1 < N)      *reinterpret_cast<double2 *>(dst) = make_double2(k0, k1);
            else if (gc < N)     dst[0] = k0;
        }
    }
    }
    if (!mirror) continue;
    __syncthreads();
    // mirrored tile: rows c0.., cols r0..
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int cm = warp * 8 + rr;            // row of the mirrored tile (a column of this tile)
        const int gr = c0 + cm;
        const int gcol = r0 + cl;
        const double k0 = s_tile[cl][cm];
        const double k1 = s_tile[cl + 1][cm];
        if (gr < N) {
            double *dst = Ab + (size_t)gr * A.ld + gcol;
            if (gcol + 1 < N)    *reinterpret_cast<double2 *>(dst) = make_double2(k0, k1);
            else if (gcol < N)   dst[0] = k0;
        }
    }
    }
}

int launch_cov_assemble(const double *x, int N, int D, const double *hyp, int P, int n_ell, int flags,
                        const double *jitter, BatchView A, int B, cudaStream_t s)
{
    if (D > MAX_ELL || n_ell > MAX_ELL) { set_error("D=%d exceeds MAX_ELL=%d", D, MAX_ELL); return GPMC_EINVAL; }
    if (B <= 0) return 0;
    const int nt = (N + AT - 1) / AT;
    dim3 grid((nt * (nt + 1) / 2 + ASM_TPC - 1) / ASM_TPC, B);
    prof_begin(KC_ASSEMBLE, s);
#define GPMC_ASM_LAUNCH(DT_)                                                                                                     \
    do {                                                                                                                          \
        if (flags & GPMC_ASM_PRED) cov_assemble_kernel<DT_, true><<<grid, 256, 0, s>>>(x, N, D, hyp, P, n_ell, flags, jitter, A);  \
        else cov_assemble_kernel<DT_, false><<<grid, 256, 0, s>>>(x, N, D, hyp, P, n_ell, flags, jitter, A);                       \
    } while (0)
    switch (D) {
        case 1: GPMC_ASM_LAUNCH(1); break;
        case 2: GPMC_ASM_LAUNCH(2); break;
        case 3: GPMC_ASM_LAUNCH(3); break;
        case 4: GPMC_ASM_LAUNCH(4); break;
        default: GPMC_ASM_LAUNCH(0); break;
    }
#undef GPMC_ASM_LAUNCH
    prof_end(KC_ASSEMBLE, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
