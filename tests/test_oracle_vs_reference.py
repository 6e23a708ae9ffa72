"""Pin the restatement to the reference file itself (build container only: needs /root/reference)."""
import numpy as np
import pytest

from oracle import reference_loader as rl, sds_oracle as so, make_golden as mg

pytestmark = pytest.mark.skipif(not rl.available(), reason='reference tree not present on this machine')


@pytest.mark.parametrize('n,it,seed', [(24, 0, 1), (24, 777, 2), (96, 499, 3), (96, 500, 4)])
def test_restatement_is_bit_identical_to_literal(n, it, seed):
    x, y = mg._series(n)
    f = np.zeros(n); hyp = np.array([1., 10., 1.2]); scale = np.array([10., 10., 5.])
    tape = rl.Tape.from_seed(seed, n)
    pf, ph, trips = rl.run_literal_with_tape(f, x, y, hyp, scale, it, tape)
    tr = so.SweepTrace()
    of, oh = so.surrogate_slice_sampling(f, x, y, hyp, scale, it, tape, trace=tr)
    assert trips == tr.n_trips and np.array_equal(ph, oh) and np.array_equal(pf, of)


def test_literal_module_functions():
    mod = rl.load_literal()
    hyp = np.array([0.7, 3.0, 1.9])
    a, b = mod.log_gamma(hyp, so.PRIOR_K3, so.PRIOR_THETA3, True)
    c, d = so.log_gamma(hyp, so.PRIOR_K3, so.PRIOR_THETA3, True)
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_multivariate_normal_draw_is_f_plus_sd_z():
    # sliceSample.py:194 goes through an SVD of the diagonal S; with this numpy it is f + sqrt(S_ii) z
    n = 50
    f = np.linspace(-1, 1, n)
    np.random.seed(7)
    g = np.random.multivariate_normal(f, np.diag(np.full(n, 1.44)), 1).reshape(n)
    np.random.seed(7)
    z = np.random.standard_normal(n)
    np.testing.assert_allclose(g, f + 1.2 * z, rtol=0, atol=1e-15)


@pytest.mark.parametrize('n,seed', [(40, 5), (120, 6)])
def test_elliptical_slice_restatement_is_bit_identical_to_literal(n, seed):
    x, y = mg._series(n)
    rs = np.random.RandomState(seed)
    hyp = np.array([4., 5., 1.8])
    f = 0.7 * (y - y.mean())
    tape = rl.EssTape(so.ess_nu_from_z(x, hyp, rs.standard_normal(n)), rs.random_sample(), rs.random_sample(64))
    pf, trips = rl.run_literal_ess_with_tape(f, x, y, hyp, tape)
    of, otrips = so.elliptical_slice(f, x, y, hyp, tape)
    assert trips == otrips and np.array_equal(pf, of)
