// Panel factor kernel, fourth form: potf2_reg.cu's register-resident fragments, with the single-warp factor step taken off
// the other warps' critical path by SPLIT named barriers (bar.arrive / bar.sync) instead of two block barriers per step.
//
// In potf2_reg.cu every step is  A (one warp factors the 8x8 diagonal fragment) | barrier | B (all warps solve their
// fragments of the block column) | barrier | C (all warps update their live fragments); phase A alone is 49 % of the wall
// time (ncu, profiles/r02k) and nothing overlaps it.  The true dependency chain is much shorter:
//     A(b)  ->  solve of ONE fragment, (b+1, b)  ->  update of ONE fragment, (b+1, b+1)  ->  A(b+1)
// Here the warp that owns block row b+1 does exactly that and nothing else before it factors; the other seven warps run
// their solves and updates of step b meanwhile.  Synchronisation per step (barrier ids alternate with the parity of b):
//     P_b  "L8 / W8 of step b are published"      owner(b): bar.arrive      the other warps: bar.sync
//     X_b  "every X(r, b) is published"           owner(b+1): bar.arrive    the other warps: bar.sync
//     Q_b  "... and visible to owner(b+1)"        the other warps: bar.arrive (right after X_b)   owner(b+1): bar.sync,
//                                                  after it has factored -- by then the others are long past it
// Every solved fragment X(r, c) and every L8 / W8 is written ONCE to its own place in shared memory (the lower 8x8 blocks
// of L11, 76 KB -- the layout the fused panel solve needs anyway), so there are no write-after-read hazards.
// Results are bit-identical to potf2_reg.cu (same operations on the same operands, only their interleaving changes).
//
// MEASURED, and therefore NOT the default (gpmc_set_tuning(1, 3) selects it): 39.3 us per block / 579 us per 4096 blocks
// against 35.1 / 512 for potf2_reg.cu; a first version synchronised with flags in shared memory (st.release, spinning
// ld.acquire) gave 38.6 / 585.  Overlapping the single-warp factor step with the other warps' work does not pay here:
// the stall samples of potf2_reg.cu are spread evenly over its ~550 straight-line instructions per factor step at ~3.7
// cycles each, shortening the pivot chain changed nothing, and every variant that makes more warps execute different
// code at the same time (this one, or walking the update suffix twice) got slower -- the signature of a kernel that is
// bound by instruction issue / fetch of long unrolled code, not by the dependency chain the restructuring shortens.
#include "common.cuh"
#include "../../include/gpmc.h"

namespace gpmc {

static_assert(NB == 128, "potf2_flow is written for a panel width of 128 (16 x 16 fragments, 8 warps)");
constexpr int PW_NF = NB / 8;
constexpr int PW_WARPS = PW_NF / 2;
constexpr int PW_THREADS = PW_WARPS * 32;
constexpr int PW_SLOTS = PW_NF + 1;
constexpr int PW_LC_ELEMS = (PW_NF * (PW_NF + 1) / 2) * 64;      // lower 8x8 blocks of L11: block (R, C) at R (R + 1) / 2 + C
constexpr int PW_SMEM = (PW_LC_ELEMS + PW_NF * 64) * (int)sizeof(double);

__device__ __forceinline__ void dmma884_w(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double pivot_rsqrt_w(double a)
{
    // one third-order step from the hardware approximation (MUFU.RSQ64H, ~2^-22):  e = 1/2 - (a/2) y^2,
    // y <- y + y e (1 + 3/2 e)  (error ~ e^3): 4 dependent FP64 operations
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    const double e = fma(-(h * y), y, 0.5);
    return fma(y * e, fma(1.5, e, 1.0), y);
}
// 1/a for the pivot CHAIN (the next pivot is a_{c+1,c+1} - a_{c+1,c}^2 / a_cc and must not wait for the square root):
// MUFU.RCP64H + one third-order step,  e = 1 - a y,  y <- y + y (e + e^2): 3 dependent FP64 operations
__device__ __forceinline__ double pivot_rcp_w(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

__device__ __forceinline__ void bar_sync_n(int id)
{
    asm volatile("barrier.sync %0, %1;\n" ::"r"(id), "r"(PW_THREADS) : "memory");
}
__device__ __forceinline__ void bar_arrive_n(int id)
{
    asm volatile("barrier.arrive %0, %1;\n" ::"r"(id), "r"(PW_THREADS) : "memory");
}

__device__ __forceinline__ void slot_rc_w(int s, int w, int &R, int &C, bool &isA)
{
    if (s < 2 * (w + 1)) { C = s >> 1; isA = (s & 1) == 0; }
    else { C = s - w - 1; isA = false; }
    R = isA ? w : PW_NF - 1 - w;
}

// phase C for one slot: acc -= X_R X_C^T with X_C = block (C, b) of the published factor
template <int S>
__device__ __forceinline__ void slot_update_w(double (&acc)[PW_SLOTS][2], int w, int b, int skip, double xa0, double xa1, double xb0, double xb1,
                                              const double *Lc, int frag_off)
{
    if (S == skip) return;
    int R, C; bool isA;
    slot_rc_w(S, w, R, C, isA);
    const double n0 = isA ? -xa0 : -xb0, n1 = isA ? -xa1 : -xb1;
    double2 xc;
    if (C == R) xc = make_double2(-n0, -n1);                              // diagonal fragment: X_C is this warp's own pair
    else xc = *reinterpret_cast<const double2 *>(&Lc[(C * (C + 1) / 2 + b) * 64 + frag_off]);
    dmma884_w(acc[S][0], acc[S][1], n0, xc.x);
    dmma884_w(acc[S][0], acc[S][1], n1, xc.y);
}

#define PW_CASE(S) case S: slot_update_w<S>(acc, w, b, skip, xa0, xa1, xb0, xb1, Lc, frag_off);
#define PW_ONE(S) case S: slot_update_w<S>(acc, w, b, -1, xa0, xa1, xb0, xb1, Lc, frag_off); break;
#define PW_PICK(S) case S: p0 = acc[S][0]; p1 = acc[S][1]; break;
#define PW_ALL(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16)

__device__ __forceinline__ void pick_slot_w(const double (&acc)[PW_SLOTS][2], int slot, double &p0, double &p1)
{
    p0 = p1 = 0.0;
    switch (slot) { PW_ALL(PW_PICK) default: break; }
}

template <bool FUSED>
__global__ void __launch_bounds__(PW_THREADS, 2)
potf2_flow_kernel(BatchView A, int n, int j0, double *__restrict__ W, long long strideW, int *__restrict__ info, int zero_upper, int n_rows)
{
    extern __shared__ __align__(16) double pw_dyn[];
    double *Lc = pw_dyn;                                 // every published 8x8 block of L11 (written once)
    double *W8c = pw_dyn + PW_LC_ELEMS;                  // the 16 diagonal 8x8 inverses
    __shared__ __align__(16) double sD[64];              // diagonal fragment on its way to the row-per-lane layout
    __shared__ int s_fail;
    const int item = blockIdx.x;
    if (A.count && item >= *A.count) return;
    const int m = batch_item(A, item);
    double *Ab = A.base + (size_t)m * A.stride + (size_t)j0 * A.ld + j0;
    double *Wb = W + (size_t)m * strideW;
    const int ld = A.ld;
    const int nv = min(NB, n - j0);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;
    const int frag_off = fr * 8 + 2 * fk;
    const int rowA = w, rowB = PW_NF - 1 - w;

    if (tid == 0) s_fail = 0;
    // ---- the block's lower fragments straight into registers (rows / columns beyond the matrix: identity)
    double acc[PW_SLOTS][2];
#pragma unroll
    for (int s = 0; s < PW_SLOTS; ++s) {
        int R, C; bool isA;
        slot_rc_w(s, w, R, C, isA);
        const int gr = R * 8 + fr, gc = C * 8 + 2 * fk;
        double2 v = make_double2(0.0, 0.0);
        if (gr < nv && gc < nv) v = *reinterpret_cast<const double2 *>(Ab + (size_t)gr * ld + gc);
        if (gr >= nv || gc >= nv) v.x = (gr == gc) ? 1.0 : 0.0;
        if (gr >= nv || gc + 1 >= nv) v.y = (gr == gc + 1) ? 1.0 : 0.0;
        acc[s][0] = v.x;
        acc[s][1] = v.y;
    }
    double xa0 = 0.0, xa1 = 0.0, xb0 = 0.0, xb1 = 0.0;   // solved fragments of rows rowA / rowB in the current block column
    __syncthreads();

    // ---- phase A of step b (the warp that owns fragment (b, b)): factor the 8x8 block with its inverse, publish
    auto factor_diag = [&](int b) {
        double d0, d1;
        pick_slot_w(acc, (b < PW_WARPS) ? 2 * b : PW_SLOTS - 1, d0, d1);
        *reinterpret_cast<double2 *>(&sD[frag_off]) = make_double2(d0, d1);
        __syncwarp();
        // lanes 0-7: row (lane) of the fragment; lanes 8-15: row (lane - 8) of an identity -- eliminating [A; I] gives
        // [L; W^T] column by column: lane 8 + k ends with column k of W8 = L8^-1.  Lanes 16-31 repeat lanes 0-15.
        const int l16 = lane & 15, r8 = lane & 7;
        double v[8];
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(&sD[r8 * 8 + c]);
            v[c] = (l16 < 8) ? ((c <= r8) ? t.x : 0.0) : ((c == r8) ? 1.0 : 0.0);
            v[c + 1] = (l16 < 8) ? ((c + 1 <= r8) ? t.y : 0.0) : ((c + 1 == r8) ? 1.0 : 0.0);
        }
        int fail = 0;
        double piv = __shfl_sync(0xffffffffu, v[0], 0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (!(piv > 0.0) && fail == 0) fail = j0 + b * 8 + c + 1;               // dpotf2: ajj <= 0 or NaN
            const double pr = pivot_rcp_w(piv);                                     // (see potf2_reg.cu: reciprocal on the chain,
            const double rinv = pivot_rsqrt_w(piv);                                 //  square root off it)
            const double pnext = (c < 7) ? __shfl_sync(0xffffffffu, fma(-(v[c] * v[c]), pr, v[c + 1]), c + 1) : 0.0;
            v[c] = (l16 == c) ? piv * rinv : v[c] * rinv;                           // the pivot row: sqrt(piv) from the pivot the chain used
#pragma unroll
            for (int j = c + 1; j < 8; ++j) {
                const double ljc = __shfl_sync(0xffffffffu, v[c], j);               // l_{j,c} from row j of the fragment
                v[j] = fma(-v[c], ljc, v[j]);
            }
            piv = pnext;
        }
        if (lane < 8) {
            const int gr = b * 8 + r8;
#pragma unroll
            for (int c = 0; c < 8; c += 2)
                *reinterpret_cast<double2 *>(&Lc[(b * (b + 1) / 2 + b) * 64 + r8 * 8 + c]) =
                    make_double2((c <= r8) ? v[c] : 0.0, (c + 1 <= r8) ? v[c + 1] : 0.0);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c <= r8 && gr < nv) Ab[(size_t)gr * ld + b * 8 + c] = v[c];
            if (fail != 0 && lane == 0 && s_fail == 0) s_fail = fail;               // the FIRST failing pivot: the steps are ordered
        } else if (lane < 16) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                W8c[b * 64 + c * 8 + r8] = v[c];                                    // W8[c][k = r8]; zero for c < k
                Wb[(size_t)(b * 8 + c) * NB + b * 8 + r8] = v[c];
            }
        }
    };

    if (w == 0) { factor_diag(0); bar_arrive_n(1); }

#pragma unroll 1
    for (int b = 0; b < PW_NF; ++b) {
        const int par = b & 1;
        const bool own = (b < PW_WARPS) ? (w == b) : (w == PW_NF - 1 - b);
        if (!own) bar_sync_n(1 + par);                                              // P_b (its owner arrived when it published)
        // slot of (rowA, b): 2b while b < w;  slot of (rowB, b): 2b+1 while b <= w, then w+1+b while b < rowB
        const int slotA = (b < w) ? 2 * b : -1;
        const int slotB = (b <= w) ? 2 * b + 1 : ((b < rowB) ? w + 1 + b : -1);
        // ------------------------------------------------------------------ B: my fragments of block column b
        if (slotB >= 0) {
            const double2 wv = *reinterpret_cast<const double2 *>(&W8c[b * 64 + frag_off]);
            const double2 lv = *reinterpret_cast<const double2 *>(&Lc[(b * (b + 1) / 2 + b) * 64 + frag_off]);
            double a0, a1, b0, b1;
            pick_slot_w(acc, slotA, a0, a1);                                        // zeros when the warp has no row-A fragment here
            pick_slot_w(acc, slotB, b0, b1);
            double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
            dmma884_w(x0, x1, a0, wv.x);                                            // X0 = A W8^T   (two independent chains)
            dmma884_w(y0, y1, b0, wv.x);
            dmma884_w(x0, x1, a1, wv.y);
            dmma884_w(y0, y1, b1, wv.y);
            double r0 = a0, r1 = a1, q0 = b0, q1 = b1;
            dmma884_w(r0, r1, -x0, lv.x);                                           // r = A - X0 L8^T
            dmma884_w(q0, q1, -y0, lv.x);
            dmma884_w(r0, r1, -x1, lv.y);
            dmma884_w(q0, q1, -y1, lv.y);
            dmma884_w(x0, x1, r0, wv.x);                                            // X = X0 + r W8^T
            dmma884_w(y0, y1, q0, wv.x);
            dmma884_w(x0, x1, r1, wv.y);
            dmma884_w(y0, y1, q1, wv.y);
            xb0 = y0; xb1 = y1;
            *reinterpret_cast<double2 *>(&Lc[(rowB * (rowB + 1) / 2 + b) * 64 + frag_off]) = make_double2(y0, y1);
            if (slotA >= 0) {
                xa0 = x0; xa1 = x1;
                *reinterpret_cast<double2 *>(&Lc[(rowA * (rowA + 1) / 2 + b) * 64 + frag_off]) = make_double2(x0, x1);
            }
        }
        if (b + 1 == PW_NF) break;
        // ------------------------------------------------------------------ the critical path: the owner of (b+1, b+1) updates
        // that fragment (its own X on both sides) and factors it before anything else; everybody else goes on to phase C
        int skip = -1;
        const bool next_own = (b + 1 < PW_WARPS) ? (w == b + 1) : (w == PW_NF - 2 - b);
        if (next_own) {
            bar_arrive_n(3 + par);                                                  // X_b: my X(., b) are published
            skip = (b + 1 < PW_WARPS) ? 2 * (b + 1) : PW_SLOTS - 1;
            switch (skip) { PW_ALL(PW_ONE) default: break; }
            factor_diag(b + 1);
            bar_arrive_n(1 + (par ^ 1));                                            // P_{b+1}
            bar_sync_n(5 + par);                                                    // Q_b: the others' X(., b) are visible to me
        } else {
            bar_sync_n(3 + par);                                                    // X_b
            bar_arrive_n(5 + par);                                                  // Q_b
        }
        // the finished fragments of block column b leave for global memory off the critical path
        if (slotB >= 0) {
            {
                const int gr = rowB * 8 + fr, gc = b * 8 + 2 * fk;
                if (gr < nv) {
                    if (gc + 1 < nv) *reinterpret_cast<double2 *>(Ab + (size_t)gr * ld + gc) = make_double2(xb0, xb1);
                    else if (gc < nv) Ab[(size_t)gr * ld + gc] = xb0;
                }
            }
            if (slotA >= 0) {
                const int gr = rowA * 8 + fr, gc = b * 8 + 2 * fk;
                if (gr < nv) {
                    if (gc + 1 < nv) *reinterpret_cast<double2 *>(Ab + (size_t)gr * ld + gc) = make_double2(xa0, xa1);
                    else if (gc < nv) Ab[(size_t)gr * ld + gc] = xa0;
                }
            }
        }
        // ------------------------------------------------------------------ C: my other live fragments (c > b): a suffix of the slots
        if (slotB >= 0) {
            const int start = (b + 1 <= w) ? 2 * (b + 1) : w + b + 2;
            switch (start) { PW_ALL(PW_CASE) default: break; }
        }
    }

    __syncthreads();
    if (tid == 0 && s_fail != 0) {
        if (info[m] == 0) info[m] = s_fail;
    }
    if (FUSED) {
        // ---------------------------------------------------------------- panel solve: X L11^T = A21 for the rows below
        // (trsm_panel8.cu's chain; a warp owns 8 rows x 128 columns as accumulator fragments, 64 rows per pass of the CTA)
        const int row_start = (j0 + NB < n) ? NB : n - j0;            // relative to the block's first row; last column: border rows only
        const int rows_end = n_rows - j0;
#pragma unroll 1
        for (int row0 = row_start; row0 < rows_end; row0 += 8 * PW_WARPS) {
            if (row0 + w * 8 >= rows_end) continue;                     // warp-uniform
            const int r = row0 + w * 8 + fr;
            const bool rv = r < rows_end;
            double *grow = Ab + (size_t)min(r, rows_end - 1) * ld + 2 * fk;
            double t[PW_NF][2];
#pragma unroll
            for (int b8 = 0; b8 < PW_NF; ++b8) {
                double2 v = make_double2(0.0, 0.0);
                if (rv && b8 * 8 + 2 * fk < nv) v = *reinterpret_cast<const double2 *>(grow + b8 * 8);
                t[b8][0] = v.x;
                t[b8][1] = v.y;
            }
#pragma unroll
            for (int b8 = 0; b8 < PW_NF; ++b8) {
                const double2 wv = *reinterpret_cast<const double2 *>(&W8c[b8 * 64 + frag_off]);
                const double2 lv = *reinterpret_cast<const double2 *>(&Lc[(b8 * (b8 + 1) / 2 + b8) * 64 + frag_off]);
                double x0 = 0.0, x1 = 0.0;
                dmma884_w(x0, x1, t[b8][0], wv.x);                                  // X0 = A W8^T
                dmma884_w(x0, x1, t[b8][1], wv.y);
                double r0 = t[b8][0], r1 = t[b8][1];
                dmma884_w(r0, r1, -x0, lv.x);                                       // r = A - X0 L8^T
                dmma884_w(r0, r1, -x1, lv.y);
                dmma884_w(x0, x1, r0, wv.x);                                        // X = X0 + r W8^T
                dmma884_w(x0, x1, r1, wv.y);
                const double nx0 = -x0, nx1 = -x1;
#pragma unroll
                for (int bp = b8 + 1; bp < PW_NF; ++bp) {
                    const double2 lp = *reinterpret_cast<const double2 *>(&Lc[(bp * (bp + 1) / 2 + b8) * 64 + frag_off]);
                    dmma884_w(t[bp][0], t[bp][1], nx0, lp.x);
                    dmma884_w(t[bp][0], t[bp][1], nx1, lp.y);
                }
                if (rv) {
                    const int c = b8 * 8 + 2 * fk;
                    if (c + 1 < nv) *reinterpret_cast<double2 *>(grow + b8 * 8) = make_double2(x0, x1);
                    else if (c < nv) grow[b8 * 8] = x0;
                }
            }
        }
    }
    if (zero_upper) {
        for (int e = tid; e < nv * nv; e += PW_THREADS) {
            const int r = e / nv, c = e - r * nv;
            if (c > r) Ab[(size_t)r * ld + c] = 0.0;
        }
    }
}

int launch_potf2_flow(BatchView A, int n, int j0, double *W, long long strideW, int *info, int zero_upper, int B, cudaStream_t s)
{
    if (B <= 0) return 0;
    if ((A.ld & 1) || (j0 & 1)) { set_error("potf2: ld=%d j0=%d must be even", A.ld, j0); return GPMC_EALIGN; }
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_flow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM));
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM));
    }
    prof_begin(KC_POTF2, s);
    potf2_flow_kernel<false><<<B, PW_THREADS, PW_SMEM, s>>>(A, n, j0, W, strideW, info, zero_upper, n);
    prof_end(KC_POTF2, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

// panel factor + panel solve of block column j0 in one launch; n_rows = n + border rows
int launch_panel_fused_flow(BatchView A, int n, int n_rows, int j0, double *W, long long strideW, int *info, int zero_upper, int B, cudaStream_t s)
{
    if (B <= 0) return 0;
    if ((A.ld & 1) || (j0 & 1)) { set_error("panel: ld=%d j0=%d must be even", A.ld, j0); return GPMC_EALIGN; }
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_flow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM));
        GPMC_CUDA_CHECK(cudaFuncSetAttribute(potf2_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM));
    }
    prof_begin(KC_POTF2, s);
    potf2_flow_kernel<true><<<B, PW_THREADS, PW_SMEM, s>>>(A, n, j0, W, strideW, info, zero_upper, n_rows);
    prof_end(KC_POTF2, s);
    GPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace gpmc
