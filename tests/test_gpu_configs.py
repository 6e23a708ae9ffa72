"""Parity at the FULL batch sizes of the BASELINE.json configs (the shapes bench.py reports as sub-records):
log-lik unit at config 2 (N=2048 x 64) and config 3 (N=512 x 4096, ARD D=4) against the oracle on sampled items plus
size-independent properties over the whole batch; one tape-driven SDS sweep at the config 3 shape against the
tape-driven oracle on sampled chains (theta' and trip counts exact).  B200 only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_LOGLIK = 1e-10          # north_star: FP64 log-likelihood within 1e-10 relative


@pytest.fixture(scope='module')
def so():
    from oracle import sds_oracle
    return sds_oracle


def _sampled(B, k):
    return np.unique(np.linspace(0, B - 1, k).astype(int))


def test_config2_loglik_full_batch(gp, so):
    """BASELINE config 2: N=2048, 64 chains, SE+noise on the IH45-shaped grid."""
    n, B = 2048, 64
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(B, n)
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0) and np.all(np.isfinite(ll))
    for b in _sampled(B, 10):
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        assert abs(ll[b] - ref) <= RTOL_LOGLIK * abs(ref), (b, ll[b], ref)
    # batch-order independence over the whole batch (bit-exact): every item is computed from its own inputs only
    perm = np.random.RandomState(2).permutation(B)
    ll2, _ = gp.ops.loglik_host(x, G[perm], H[perm])
    assert np.array_equal(ll2, ll[perm])


def test_config3_loglik_full_batch(gp, so):
    """BASELINE config 3: N=512, 4096 chains, ARD kernel with D=4 (P=6)."""
    n, B, D = 512, 4096, 4
    x, _ = gp.synthetic.ard_inputs(n, D)
    G, H = gp.synthetic.loglik_batch(B, n, n_ell=D)
    ll, info = gp.ops.loglik_host(x, G, H)
    assert np.all(info == 0) and np.all(np.isfinite(ll))
    worst = 0.0
    for b in _sampled(B, 12):
        ref = so.loglik_unit(x, G[b], H[b], form='chol')
        worst = max(worst, abs(ll[b] - ref) / abs(ref))
    print('config 3 full batch: worst rel err on 12 sampled items %.2e' % worst)
    assert worst <= RTOL_LOGLIK
    # properties over the WHOLE batch: g -> -g leaves the value unchanged (bit-exact: the quadratic form is even), and
    # the second half of the batch evaluated alone (different wave position) reproduces its values bit for bit
    ll_neg, _ = gp.ops.loglik_host(x, -G, H)
    assert np.array_equal(ll_neg, ll)
    ll_half, _ = gp.ops.loglik_host(x, G[B // 2:], H[B // 2:])
    assert np.array_equal(ll_half, ll[B // 2:])


def test_config3_sds_sweep_full_batch_vs_oracle(gp, so):
    """One SDS transition (sliceSample.py:76-163) for 4096 chains at N=512, ARD D=4, every chain on its own explicit
    tape; sampled chains are compared with the tape-driven oracle: trip counts and theta' exact (theta' to rounding),
    log N(g) to 1e-10, f' at the reference's own resolution (see test_gpu_sds.py)."""
    import torch
    from gpmc_b200 import ops
    from oracle.reference_loader import Tape
    n, B, D = 512, 4096, 4
    P = D + 2
    x, y = gp.synthetic.ard_inputs(n, D)
    scale = np.array([gp.synthetic.SCALE[0]] * D + list(gp.synthetic.SCALE[1:]))
    F0, H0 = gp.synthetic.chain_states(B, n, n_ell=D)
    T = 40
    rs = np.random.RandomState(512)
    z = rs.standard_normal((B, n)); v = rs.random_sample((B, P)); u0 = rs.random_sample(B); U = rs.random_sample((B, T, P))
    F = torch.tensor(F0).cuda(); Hd = torch.tensor(H0).cuda()
    nt, ll, st = ops.sds_sweep(x, y, F, Hd, scale, 0, tape=ops.Tape(z, v, u0, U), max_trips=T)
    nt, ll, st = nt.cpu().numpy(), ll.cpu().numpy(), st.cpu().numpy()
    Fn, Hn = F.cpu().numpy(), Hd.cpu().numpy()
    assert int((st != 0).sum()) <= 2, 'more than 2 of 4096 chains used up %d proposals' % T
    assert np.all(Hn > 0) and np.all(np.isfinite(Fn))
    # the chain with the most trips, the one with the fewest, and a spread in between
    order = np.argsort(nt, kind='stable')
    picks = sorted(set([int(order[0]), int(order[-1])] + [int(i) for i in _sampled(B, 5)]))
    for c in picks:
        if st[c] != 0:
            continue
        tr = so.SweepTrace()
        of, oh = so.surrogate_slice_sampling(F0[c], x, y, H0[c], scale, 0, Tape(z[c], v[c], u0[c], U[c]), trace=tr, r_form='reduced')
        assert nt[c] == tr.n_trips, (c, nt[c], tr.n_trips)
        np.testing.assert_allclose(Hn[c], oh, rtol=1e-12, atol=0)
        ref = tr.propG_chol[-1]
        assert abs(ll[c] - ref) <= RTOL_LOGLIK * abs(ref)
        assert np.abs(Fn[c] - of).max() < 5e-2
    print('config 3 SDS sweep: trips mean %.2f max %d, %d sampled chains match the oracle' % (nt.mean(), nt.max(), len(picks)))


def test_tape_driven_waves_match_single_wave(gp):
    """Forcing several waves (c0 > 0: tape / chain-id offsets, write-back of every wave) must reproduce the single-wave
    result bit for bit, with explicit tapes as well as with Philox."""
    import torch
    from gpmc_b200 import ops
    n, B = 96, 11
    x, y = gp.synthetic.ih45_series(n)
    scale = np.array(gp.synthetic.SCALE)
    F0, H0 = gp.synthetic.chain_states(B, n)
    rs = np.random.RandomState(9)
    tape = ops.Tape(rs.standard_normal((B, n)), rs.random_sample((B, 3)), rs.random_sample(B), rs.random_sample((B, 48, 3)))

    def run(wave, tp):
        F = torch.tensor(F0.copy()).cuda(); H = torch.tensor(H0.copy()).cuda()
        nt, ll, st = ops.sds_sweep(x, y, F, H, scale, 3, tape=tp, seed=77, max_trips=48, chains_per_wave=wave,
                                   workspace=ops.Workspace())
        return F.cpu().numpy(), H.cpu().numpy(), nt.cpu().numpy(), ll.cpu().numpy(), st.cpu().numpy()
    for tp in (tape, None):
        a = run(None, tp)
        for wave in (3, 4):
            b = run(wave, tp)
            for u, w in zip(a, b):
                assert np.array_equal(u, w)
