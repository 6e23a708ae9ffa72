"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden vectors.  B200 only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_rows

pytestmark = pytest.mark.gpu

RTOL_LOGLIK = 1e-10      # north_star: FP64 log-likelihood within 1e-10 relative of the reference


@pytest.fixture(scope='module')
def so():
    from oracle import sds_oracle
    return sds_oracle


def _series(gp, n):
    return gp.synthetic.ih45_series(n)


# ------------------------------------------------------------------ K1/K2 assembly
@pytest.mark.parametrize('n', [1, 8, 63, 200, 455, 512, 1000])
def test_cov_assemble_iso_matches_cdist_form(gp, so, n):
    x = np.arange(n, dtype=np.float64).reshape(n, 1) * 0.7 + 0.1
    H = np.array([[1., 10., 1.2], [0.35, 2.0, 0.2], [7.8868277, 9.69728866, 1.2], [12.5, 18., 0.05]])
    A = gp.ops.cov_assemble(x, H).cpu().numpy()[:, :, :n]
    for b in range(H.shape[0]):
        K = so.cov_matrix(x, H[b])
        # identical argument arithmetic; exp() is <= 1 ulp on both sides
        np.testing.assert_allclose(A[b], K, rtol=1e-15 * 4, atol=0)
        assert np.array_equal(A[b], A[b].T)
    AS = gp.ops.cov_assemble(x, H, add_S=True).cpu().numpy()[:, :, :n]
    for b in range(H.shape[0]):
        K = so.cov_matrix(x, H[b])
        KS = K + np.diag(so.s_diagonal(np.diagonal(K), H[b, 2]))
        np.testing.assert_allclose(AS[b], KS, rtol=4e-15, atol=0)


@pytest.mark.parametrize('n,d', [(32, 2), (300, 4), (512, 4), (150, 5), (70, 8)])
def test_cov_assemble_ard(gp, so, n, d):
    rs = np.random.RandomState(5)
    x = rs.uniform(0, 10, size=(n, d))
    H = np.column_stack([rs.uniform(0.5, 6, size=(3, d)), rs.uniform(1, 12, size=3), rs.uniform(0.3, 3, size=3)])
    A = gp.ops.cov_assemble(x, H, add_S=True).cpu().numpy()[:, :, :n]
    for b in range(3):
        K = so.cov_matrix(x, H[b])
        KS = K + np.diag(so.s_diagonal(np.diagonal(K), H[b, -1]))
        # D > 1: scipy's cdist may sum the D squares in another order (1 ulp of the squared distance s),
        # which exp() turns into a relative 1.1e-16 * s/2 of K; s <= ~1400 before K underflows.
        np.testing.assert_allclose(A[b], KS, rtol=2e-13, atol=1e-300)


def test_cov_assemble_lower_only_and_jitter(gp, so):
    n = 200
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[3., 5., 1.0]])
    jit = np.array([0.125])
    A = gp.ops.cov_assemble(x, H, add_S=True, lower_only=True, jitter=jit).cpu().numpy()[0, :, :n]
    K = so.cov_matrix(x, H[0])
    KS = K + np.diag(so.s_diagonal(np.diagonal(K), 1.0)) + 0.125 * np.eye(n)
    err = np.abs(np.tril(A) - np.tril(KS))
    print('lower_only/jitter max abs err', err.max(), 'at', np.unravel_index(err.argmax(), err.shape))
    # ell = exp(log(3.0)) is evaluated by CUDA's libm here and by the host libm in the oracle (each <= 1 ulp);
    # one ulp of ell is a relative 2.2e-16 * |arg| of K, |arg| <= 745 before K underflows
    np.testing.assert_allclose(np.tril(A), np.tril(KS), rtol=1e-12, atol=1e-300)


# ------------------------------------------------------------------ K3 batched Cholesky
@pytest.mark.parametrize('n', [1, 5, 128, 129, 200, 257, 384, 385, 455, 1000])
def test_potrf_batched_matches_lapack(gp, so, n):
    import torch
    import scipy.linalg
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[1., 10., 1.2], [5., 4., 2.5], [0.35, 2., 0.2]])
    A = gp.ops.cov_assemble(x, H, add_S=True)
    A0 = A.cpu().numpy()[:, :, :n].copy()
    info = gp.ops.potrf_batched(A, n=n)
    assert np.all(info.cpu().numpy() == 0)
    L = A.cpu().numpy()[:, :, :n]
    for b in range(3):
        assert np.all(np.triu(L[b], 1) == 0)
        ref = scipy.linalg.cholesky(A0[b], lower=True)
        resid = np.linalg.norm(L[b] @ L[b].T - A0[b]) / np.linalg.norm(A0[b])
        assert resid < 5e-15
        np.testing.assert_allclose(np.diag(L[b]), np.diag(ref), rtol=1e-12)
        np.testing.assert_allclose(np.log(np.diag(L[b])).sum(), np.log(np.diag(ref)).sum(), rtol=1e-12)


@pytest.mark.parametrize('n,ld', [(300, 302), (129, 130), (257, 272), (1000, 1000), (455, 456)])
def test_potrf_with_caller_leading_dimension_and_dirty_pads(gp, n, ld):
    """gpmc_potrf_batched takes any even ld >= N; whatever sits in the pad columns (NaN here) and in the strict upper
    triangle (garbage here) must not reach the factor (LAPACK dpotrf reads the lower triangle only)."""
    import torch
    import scipy.linalg
    rs = np.random.RandomState(n)
    B = 3
    A = np.full((B, n, ld), np.nan)
    ref = []
    for b in range(B):
        M = rs.standard_normal((n, n))
        S = M @ M.T / n + (1.0 + b) * np.eye(n)
        ref.append(scipy.linalg.cholesky(S, lower=True))
        A[b, :, :n] = np.tril(S) + np.triu(rs.standard_normal((n, n)) * 1e3, 1)
    T = torch.tensor(A, device='cuda')
    info = gp.ops.potrf_batched(T, n=n, jitter_policy=gp.JITTER_NONE, zero_upper=True).cpu().numpy()
    assert np.all(info == 0)
    L = T.cpu().numpy()[:, :, :n]
    for b in range(B):
        assert np.all(np.triu(L[b], 1) == 0)
        np.testing.assert_allclose(L[b], ref[b], rtol=0, atol=1e-12 * np.abs(ref[b]).max())


def test_potrf_reports_lapack_info(gp):
    import torch
    import scipy.linalg
    n = 300
    rs = np.random.RandomState(0)
    M = rs.standard_normal((n, n))
    A = M @ M.T + n * np.eye(n)
    A[170, 170] = -1.0                      # leading minor 171 is not positive definite
    _, ref_info = scipy.linalg.lapack.dpotrf(A, lower=1)
    assert ref_info == 171
    T = torch.tensor(np.stack([A, M @ M.T + n * np.eye(n)]), device='cuda')
    info = gp.ops.potrf_batched(T, jitter_policy=gp.JITTER_NONE).cpu().numpy()
    assert info[0] == 171 and info[1] == 0


@pytest.mark.parametrize('n', [37, 128, 300, 1000])
def test_both_panel_factor_kernels_agree(gp, so, n):
    """potf2_reg.cu (register-resident fragments, default = mode 0), potf2_flow.cu (the same as a dataflow program without
    block barriers, mode 3; must be BIT-identical to mode 0), potf2.cu (full block inverse, mode 1) and potf2_lite.cu
    (shared-memory block, mode 2): force each in turn over the same matrices, incl. a non-PD one (LAPACK info must match)."""
    import torch
    import scipy.linalg
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[1., 10., 1.2], [5., 4., 2.5], [0.35, 2., 0.2], [2., 3., 0.5]])
    A0 = gp.ops.cov_assemble(x, H, add_S=True)
    bad = min(n - 1, 170)
    out = {}
    try:
        for mode in (0, 1, 2, 3):
            gp.ops.set_tuning(1, mode)
            A = A0.clone()
            A[3, bad, bad] = -1.0
            info = gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE).cpu().numpy()
            out[mode] = (A.cpu().numpy()[:, :, :n], info)
    finally:
        gp.ops.set_tuning(1, 0)
    (L1, i1) = out[1]
    Ah = A0.cpu().numpy()[:, :, :n]
    assert np.array_equal(out[0][0][:3], out[3][0][:3]) and np.array_equal(out[0][1], out[3][1])
    for mode in (0, 2):
        L2, i2 = out[mode]
        assert np.array_equal(i1, i2) and np.all(i1[:3] == 0) and i1[3] == bad + 1
        for b in range(3):
            assert np.all(np.triu(L2[b], 1) == 0)
            resid = np.linalg.norm(L2[b] @ L2[b].T - Ah[b]) / np.linalg.norm(Ah[b])
            assert resid < 5e-15, (mode, b, resid)
            ref = scipy.linalg.cholesky(Ah[b], lower=True)
            np.testing.assert_allclose(np.diag(L2[b]), np.diag(ref), rtol=1e-12)
            np.testing.assert_allclose(L2[b], L1[b], rtol=0, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize('n', [129, 193, 257, 300, 1000, 1153])
def test_both_panel_solve_kernels_agree(gp, so, n):
    """trsm_panel8.cu (8-column sub-blocks, 2 CTAs/SM, default) and trsm_panel.cu (32-column sub-blocks) solve the same
    rows with the same inverse-multiply + refinement scheme; sizes put the last CTA at 1, 64, 65 and ragged row counts."""
    import scipy.linalg
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[1., 10., 1.2], [5., 4., 2.5], [0.35, 2., 0.2]])
    A0 = gp.ops.cov_assemble(x, H, add_S=True)
    Ah = A0.cpu().numpy()[:, :, :n]
    G = np.random.RandomState(n).standard_normal((3, n))
    out = {}
    try:
        for mode in (0, 1):
            gp.ops.set_tuning(4, mode)
            A = A0.clone()
            info = gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE).cpu().numpy()
            assert np.all(info == 0)
            ll, _ = gp.ops.loglik_host(x, G, H)           # exercises the border row in the last CTA
            out[mode] = (A.cpu().numpy()[:, :, :n], ll)
    finally:
        gp.ops.set_tuning(4, 0)
    for b in range(3):
        ref = scipy.linalg.cholesky(Ah[b], lower=True)
        for mode in (0, 1):
            L = out[mode][0][b]
            assert np.linalg.norm(L @ L.T - Ah[b]) / np.linalg.norm(Ah[b]) < 5e-15
            np.testing.assert_allclose(L, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
            want = so.loglik_unit(x, G[b], H[b], form='chol')
            assert abs(out[mode][1][b] - want) <= RTOL_LOGLIK * abs(want)


def test_potrf_jitter_ladder_matches_jitchol(gp):
    import torch
    from oracle import kcgp_shim
    n = 200
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    # long length-scale, no noise: numerically singular -> dpotrf fails -> pyGPs jitter ladder
    K = kcgp_shim.RBF(np.log(30.), np.log(3.)).getCovMatrix(x=x, mode='train')
    import scipy.linalg
    assert scipy.linalg.lapack.dpotrf(K, lower=1)[1] != 0
    ref = kcgp_shim.jitchol(K)
    T = torch.tensor(K[None].copy(), device='cuda')
    info = gp.ops.potrf_batched(T, jitter_policy=gp.JITTER_PYGPS).cpu().numpy()
    assert info[0] == 0
    L = T.cpu().numpy()[0]
    # same jitter level reached: L L^T - K is jitter * I with the ladder's value
    jit_ref = np.mean(np.diag(ref @ ref.T - K))
    jit_got = np.mean(np.diag(L @ L.T - K))
    assert jit_got == pytest.approx(jit_ref, rel=1e-3)
    Bad = torch.tensor((-np.eye(n))[None].copy(), device='cuda')
    assert gp.ops.potrf_batched(Bad, jitter_policy=gp.JITTER_PYGPS).cpu().numpy()[0] == -1


# ------------------------------------------------------------------ the metric's unit
def _check_rows(gp, rows, ard=False):
    worst = 0.0
    for r in rows:
        x = r['x'] if ard else r['x'].reshape(-1, 1)
        ll, info = gp.ops.loglik_host(x, r['g'][None], r['hyp'][None])
        assert info[0] == 0
        ref_c, ref_i, cond = float(r['ll_chol']), float(r['ll_inv']), float(r['cond'])
        rel_c = abs(ll[0] - ref_c) / abs(ref_c)
        rel_i = abs(ll[0] - ref_i) / abs(ref_i)
        # Cholesky form (sliceSample.py:145-146): 1e-10 relative while cond(K+S) <= ~1e6; beyond that no two
        # FP64 evaluation orders agree to 1e-10 (the reference's own two forms differ by 2.4e-9 at cond 3e8),
        # so the bound becomes the conditioning limit 0.5 * eps * cond.
        assert rel_c < max(RTOL_LOGLIK, 1.1e-16 * cond), (r['hyp'], cond, rel_c)
        # literal inv form (:147): 1e-10 up to cond ~1e7; beyond that the reference's own two forms
        # disagree by more than 1e-10 (recorded in the fixture), so allow that gap
        gap = abs(ref_i - ref_c) / abs(ref_c)
        assert rel_i < max(RTOL_LOGLIK, gap + 1.1e-16 * cond), (r['hyp'], cond, rel_i, gap)
        worst = max(worst, rel_c)
    return worst


def test_loglik_golden_iso(gp):
    rows = load_rows(os.path.join(GOLDEN, 'loglik_iso.npz'))
    worst = _check_rows(gp, rows)
    print('worst rel err vs chol form: %.2e over %d cases' % (worst, len(rows)))


def test_loglik_golden_ard(gp):
    _check_rows(gp, load_rows(os.path.join(GOLDEN, 'loglik_ard.npz')), ard=True)


def test_loglik_matches_reference_curG_in_sds_fixtures(gp):
    import glob
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'sds_N*.npz'))):
        z = np.load(path)
        ll, info = gp.ops.loglik_host(z['x'], z['ref_g'][None], z['hyp'][None])
        assert info[0] == 0
        assert abs(ll[0] - float(z['ref_curG'])) <= RTOL_LOGLIK * abs(float(z['ref_curG'])), path
        # every proposal the reference evaluated in its shrink loop (sliceSample.py:147)
        T = z['trace_prop_hyp'].shape[0]
        Gs = np.repeat(z['ref_g'][None], T, axis=0)
        ll, info = gp.ops.loglik_host(z['x'], Gs, z['trace_prop_hyp'])
        ok = np.isfinite(z['trace_propG'])
        np.testing.assert_allclose(ll[ok], z['trace_propG'][ok], rtol=RTOL_LOGLIK)


# sizes around the panel width: the border row that carries g (sequences.cu) lands in every position of the last row
# tile / of the panel solve's last CTA (N = 128k, 128k + 1, 128k - 1, ragged)
@pytest.mark.parametrize('n,B', [(1000, 5), (2048, 3), (127, 3), (128, 3), (129, 3), (256, 2), (257, 3), (385, 2), (641, 200)])
def test_loglik_batched_vs_oracle_seeded(gp, so, n, B):
    x, _ = _series(gp, n)
    G, H = gp.synthetic.loglik_batch(B, n)
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    for b in range(B):
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(ll[b] - ref) <= RTOL_LOGLIK * abs(ref)


def test_loglik_device_and_host_paths_agree_and_are_deterministic(gp):
    import torch
    n, B = 640, 9
    x, _ = _series(gp, n)
    G, H = gp.synthetic.loglik_batch(B, n)
    a, _ = gp.ops.loglik_host(x, G, H)
    b, _ = gp.ops.loglik_batched(torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda())
    c, _ = gp.ops.loglik_batched(torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda(), max_wave=2)
    assert np.array_equal(a, b.cpu().numpy()) and np.array_equal(a, c.cpu().numpy())
    perm = np.random.RandomState(1).permutation(B)
    d, _ = gp.ops.loglik_host(x, G[perm], H[perm])
    assert np.array_equal(d, a[perm])


def test_loglik_full_size_properties(gp, so):
    """BASELINE size N=4096: quadratic form is homogeneous of degree 2 in g, log-det term is g-free,
    and one item is checked against the oracle directly."""
    n, B = 4096, 4
    x, _ = _series(gp, n)
    G, H = gp.synthetic.loglik_batch(B, n)
    H[:] = H[0]
    G[1] = 0.0
    G[2] = 3.0 * G[0]
    G[3] = -G[0]
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    q0 = ll[0] - ll[1]
    assert abs((ll[2] - ll[1]) - 9.0 * q0) <= 1e-11 * abs(ll[2])
    assert ll[3] == ll[0]
    ref = so.loglik_unit(x, G[0], H[0], form='chol')
    assert abs(ll[0] - ref) <= RTOL_LOGLIK * abs(ref)
    ref0 = so.loglik_unit(x, G[1], H[1], form='chol')
    assert abs(ll[1] - ref0) <= RTOL_LOGLIK * abs(ref0)


def test_loglik_edge_cases(gp, so):
    n = 200
    x, _ = _series(gp, n)
    g = np.random.RandomState(3).standard_normal(n)
    # empty batch
    ll, info = gp.ops.loglik_host(x, np.zeros((0, n)), np.zeros((0, 3)))
    assert ll.shape == (0,) and info.shape == (0,)
    # sn = 0 (bracket clamp, sliceSample.py:111): S = 0, K singular -> jitchol ladder, same answer as the oracle
    H = np.array([[30., 3., 0.0], [1., 10., 1.2], [2., 0.0, 1.0]])
    G = np.stack([g, g, g])
    ll, info = gp.ops.loglik_host(x, G, H)
    ref0 = so.loglik_unit(x, g, H[0], form='chol')
    assert info[0] == 0 and abs(ll[0] - ref0) <= 1e-6 * abs(ref0)      # cond ~1e6/jitter: looser, stated
    ref1 = so.loglik_unit(x, g, H[1], form='chol')
    assert info[1] == 0 and abs(ll[1] - ref1) <= RTOL_LOGLIK * abs(ref1)
    # sf = 0: K = 0, S = NaN -> the reference's jitchol raises LinAlgError; here info = -1 and NaN (rejected)
    assert info[2] == -1 and np.isnan(ll[2])
    # (the oracle either raises LinAlgError, as LAPACK's dpotrf flags a NaN pivot, or -- with OpenBLAS's own
    #  potrf, which does not -- returns NaN; both are a rejected proposal at sliceSample.py:154)
    try:
        assert np.isnan(so.loglik_unit(x, g, H[2]))
    except np.linalg.LinAlgError:
        pass
    # N = 1
    ll, info = gp.ops.loglik_host(np.zeros((1, 1)), np.array([[0.7]]), np.array([[1., 2., 0.5]]))
    assert abs(ll[0] - so.loglik_unit(np.zeros((1, 1)), np.array([0.7]), np.array([1., 2., 0.5]))) < 1e-13


def test_loglik_large_single_matrix(gp, so):
    """BASELINE config 4 shape: one N=16384 matrix (2 GiB): g-homogeneity property plus a direct oracle check."""
    n = 16384
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    rs = np.random.RandomState(16384)
    H = np.array([[5.0, 4.0, 2.5]] * 3)
    g = 2.5 * rs.standard_normal(n)
    G = np.stack([g, np.zeros(n), -2.0 * g])
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    q = ll[0] - ll[1]
    assert abs((ll[2] - ll[1]) - 4.0 * q) <= 1e-11 * abs(ll[2])
    ref = so.loglik_unit(x, g, H[0], form='trsv')
    assert abs(ll[0] - ref) <= RTOL_LOGLIK * abs(ref)


@pytest.mark.parametrize('n,B', [(1000, 3), (2500, 1), (3000, 2), (600, 40)])
def test_lookahead_schedule_equals_sequential_schedule(gp, so, n, B):
    """The blocked Cholesky overlaps panel kernels and update GEMMs on side streams when few matrices are in flight
    (sequences.cu).  Forced on and forced off must agree (same kernels, the update's contraction is only split in two),
    match the oracle, and be reproducible run to run (a missing dependency would show as run-to-run differences)."""
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(B, n)
    out = {}
    try:
        for mode in (1, 2):
            gp.ops.set_tuning(2, mode)
            runs = [gp.ops.loglik_host(x, G, H) for _ in range(4 if mode == 2 else 1)]
            for ll, info in runs:
                assert np.all(info == 0)
                assert np.array_equal(ll, runs[0][0])
            out[mode] = runs[0][0]
    finally:
        gp.ops.set_tuning(2, 0)
    np.testing.assert_allclose(out[2], out[1], rtol=1e-12)
    for b in range(min(B, 2)):
        ref = so.loglik_unit(x, G[b], H[b], form='trsv')
        assert abs(out[2][b] - ref) <= RTOL_LOGLIK * abs(ref)


def test_loglik_random_shapes_against_oracle(gp, so):
    """Seeded sweep over ragged shapes: every (N, B) picks its own mix of schedules (left-looking / windowed, look-ahead
    on or off, either panel factor kernel, border row in any position of the last tile) -- all must match the oracle."""
    rs = np.random.RandomState(20261018)
    worst = 0.0
    for case in range(28):
        n = int(rs.choice([rs.randint(1, 130), rs.randint(130, 400), rs.randint(400, 900)]))
        B = int(rs.choice([1, 2, 3, 7, 33, 160]))
        d = int(rs.choice([1, 1, 2, 3]))
        ard = d > 1 and bool(rs.randint(2))
        x = np.sort(rs.uniform(0, n, size=(n, d)), axis=0) if d > 1 else np.arange(n, dtype=np.float64).reshape(n, 1)
        G, H = gp.synthetic.loglik_batch(B, n, n_ell=d if ard else 1)
        ll, info = gp.ops.loglik_host(x, G, H)
        assert np.all(info == 0), (n, B, d, ard)
        for b in sorted(set([0, B - 1, B // 2])):
            ref = so.loglik_unit(x, G[b], H[b], form='chol')
            rel = abs(ll[b] - ref) / abs(ref)
            K = so.cov_matrix(x, H[b])
            cond = np.linalg.cond(K + np.diag(so.s_diagonal(np.diagonal(K), H[b, -1])))
            tol = max(RTOL_LOGLIK, 1.1e-16 * cond)          # the conditioning limit of the quantity itself
            worst = max(worst, rel / tol)
            assert rel <= tol, (n, B, d, ard, b, rel, cond)
    print('random shapes: worst error / tolerance', worst)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5])
def test_every_tile_kernel_variant_factors_correctly(gp, so, cfg):
    """The DMMA tile kernel exists in several variants (cp.async 128x128 with 8 or 16 warps, cp.async 128x64 with two CTAs
    per SM, TMA-staged 128x64 in lock step / free running / persistent with in-order tile hand-out); all must give the
    same factor."""
    import scipy.linalg
    n = 700
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[2.0, 7.0, 0.9], [6.0, 3.0, 2.0]])
    try:
        gp.ops.set_tuning(0, cfg)
        A = gp.ops.cov_assemble(x, H, add_S=True)
        A0 = A.cpu().numpy()[:, :, :n].copy()
        info = gp.ops.potrf_batched(A, n=n)
        G = np.random.RandomState(cfg).standard_normal((2, n))
        ll, _ = gp.ops.loglik_host(x, G, H)
    finally:
        gp.ops.set_tuning(0, 4)
    assert np.all(info.cpu().numpy() == 0)
    L = A.cpu().numpy()[:, :, :n]
    for b in range(2):
        ref = scipy.linalg.cholesky(A0[b], lower=True)
        np.testing.assert_allclose(L[b], ref, rtol=1e-10, atol=1e-12)
        want = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(ll[b] - want) <= RTOL_LOGLIK * abs(want)


@pytest.mark.parametrize('n,d', [(129, 2), (257, 3)])
def test_iso_kernel_with_multidimensional_inputs(gp, so, n, d):
    """covK.RBF on D > 1 inputs (one shared length-scale, P = 3): assembly and the fused unit."""
    rs = np.random.RandomState(n)
    x = rs.uniform(0, 20, size=(n, d))
    H = np.array([[3.0, 6.0, 1.1], [0.8, 2.0, 0.4]])
    A = gp.ops.cov_assemble(x, H).cpu().numpy()[:, :, :n]
    G = rs.standard_normal((2, n))
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    for b in range(2):
        np.testing.assert_allclose(A[b], so.cov_matrix(x, H[b]), rtol=2e-13, atol=1e-300)
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(ll[b] - ref) <= RTOL_LOGLIK * abs(ref)


def test_odd_and_unaligned_sizes_through_the_fused_unit(gp, so):
    """N that is odd / not a multiple of 16, 64 or 128: padded leading dimension, partial tiles and blocks."""
    rs = np.random.RandomState(11)
    for n in (3, 17, 127, 129, 191, 1001, 2047):
        x = np.arange(n, dtype=np.float64).reshape(n, 1) * 0.9
        H = np.array([[4.0, 5.0, 1.5], [1.2, 9.0, 0.7]])
        G = rs.standard_normal((2, n))
        ll, info = gp.ops.loglik_host(x, G, H)
        assert np.all(info == 0)
        for b in range(2):
            ref = so.loglik_unit(x, G[b], H[b], form='chol')
            assert abs(ll[b] - ref) <= RTOL_LOGLIK * abs(ref), (n, b)


@pytest.mark.parametrize('n', [64, 127, 128, 129, 200, 257, 385, 512, 641, 1000])
def test_fused_panel_launches_agree_with_separate_ones(gp, so, n):
    """gpmc_set_tuning(9, .): panel factor + panel solve of a block column in one launch (one CTA per matrix, chosen
    automatically for many small matrices) against the separate potf2 / trsm launches.  The factor is the same arithmetic
    in the same order -> bit-identical L; the log marginal differs only in how the border row's LAST block is solved
    (panel-solve chain vs substitution kernel) -> 1e-13, and both match the oracle to 1e-10."""
    import torch
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.array([[1., 10., 1.2], [5., 4., 2.5], [0.35, 2., 0.2], [2., 3., 0.5], [7.0, 6.0, 1.0]])
    G = np.random.RandomState(n).standard_normal((5, n)) * H[:, 2:3]
    A0 = gp.ops.cov_assemble(x, H, add_S=True)
    out = {}
    try:
        for mode in (1, 2, 3):
            gp.ops.set_tuning(9, min(mode, 2))
            gp.ops.set_tuning(1, 3 if mode == 3 else 0)          # mode 3: fused launches of the dataflow kernel
            A = A0.clone()
            info = gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE).cpu().numpy()
            ll, info2 = gp.ops.loglik_host(x, G, H)
            out[mode] = (A.cpu().numpy()[:, :, :n], info, ll, info2)
    finally:
        gp.ops.set_tuning(9, 0); gp.ops.set_tuning(1, 0)
    assert np.array_equal(out[3][0], out[2][0]) and np.array_equal(out[3][2], out[2][2])
    assert np.all(out[1][1] == 0) and np.all(out[2][1] == 0) and np.all(out[2][3] == 0)
    assert np.array_equal(out[1][0], out[2][0])
    np.testing.assert_allclose(out[2][2], out[1][2], rtol=1e-13)
    for b in range(5):
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(out[2][2][b] - ref) <= RTOL_LOGLIK * abs(ref)


def test_fused_panel_launches_in_the_sds_sweep(gp):
    """The same switch inside the sampler (bordered chol(K+S), chol(R + 1e-11 I), and the literal form's n more border
    rows): theta', trips and log N(g) exact against the separate-launch path, f' within its resolution."""
    import torch
    from gpmc_b200 import ops
    n, B = 200, 7
    x, y = gp.synthetic.ih45_series(n)
    scale = np.array(gp.synthetic.SCALE)
    F0, H0 = gp.synthetic.chain_states(B, n)
    res = {}
    try:
        for lit in (0, 1):
            ops.set_tuning(8, lit)
            for mode in (1, 2):
                ops.set_tuning(9, mode)
                F = torch.tensor(F0.copy()).cuda(); H = torch.tensor(H0.copy()).cuda()
                nt, ll, st = ops.sds_sweep(x, y, F, H, scale, 600, seed=3, workspace=ops.Workspace())
                res[(lit, mode)] = (F.cpu().numpy(), H.cpu().numpy(), nt.cpu().numpy(), ll.cpu().numpy(), st.cpu().numpy())
    finally:
        ops.set_tuning(9, 0); ops.set_tuning(8, 0)
    for lit in (0, 1):
        a, b = res[(lit, 1)], res[(lit, 2)]
        assert np.all(a[4] == 0) and np.all(b[4] == 0)
        assert np.array_equal(a[2], b[2])
        np.testing.assert_allclose(b[1], a[1], rtol=1e-12)
        np.testing.assert_allclose(b[3], a[3], rtol=1e-10)
        assert np.abs(a[0] - b[0]).max() < 5e-2


def test_two_host_threads_are_serialised_not_racing(gp, so):
    """The library keeps per-process state (side streams, staging buffers): every computing entry point of the C ABI takes a
    recursive lock, so two host threads (ctypes drops the GIL during the calls) get correct, identical results."""
    import threading
    n = 300
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(24, n)
    want, _ = gp.ops.loglik_host(x, G, H)
    out = {}

    def work(k):
        res = []
        for _ in range(6):
            ll, info = gp.ops.loglik_host(x, G[k::2], H[k::2])
            res.append(ll)
        out[k] = res
    ts = [threading.Thread(target=work, args=(k,)) for k in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in (0, 1):
        for ll in out[k]:
            assert np.array_equal(ll, want[k::2])


@pytest.mark.parametrize('n,B', [(300, 9), (512, 700), (1000, 3), (2500, 1), (129, 40)])
def test_persistent_tile_kernel_is_bit_identical(gp, n, B):
    """gemm_dmma_tma_persistent_kernel (tuning key 0 = 5: persistent CTAs, tiles drawn in order from a global counter,
    next tile's first chunks fetched during the current tile's tail) computes every tile exactly as the default kernel
    does, whatever order the tiles are handed out in -- including under the look-ahead streams (few matrices) and with
    far more tiles than CTAs (many matrices); repeated launches reuse the scheduler ring."""
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(B, n)
    ref, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    try:
        gp.ops.set_tuning(0, 5)
        for _ in range(3):
            ll, info = gp.ops.loglik_host(x, G, H)
            assert np.all(info == 0)
            assert np.array_equal(ll, ref)
    finally:
        gp.ops.set_tuning(0, 4)


@pytest.mark.parametrize('n', [1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 33])
def test_tiny_matrices(gp, so, n):
    """Orders below and around one 8x8 fragment / one 16-column pad: a single ragged panel with identity padding."""
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(3, n)
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0)
    for b in range(3):
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(ll[b] - ref) <= RTOL_LOGLIK * abs(ref)
