"""``kcMCMC.sliceSample`` with the reference's function signatures, computed on the B200.

Mirrors ``/root/reference/kcMCMC/sliceSample.py``:

* ``surrogate_slice_sampling(f, x, y, hyp, scale, iter=0) -> (prop_f, prop_hyp)``   (``:76-163``)
* ``log_gamma(x, k, theta, invG) -> (logG, gradG)``                                  (``:209-232``)

numpy arrays in, fresh numpy arrays out; inputs are never mutated.  Randomness comes from the global
``numpy.random`` stream in the reference's draw order (N normals for ``g``, P uniforms for the bracket, one for
the threshold, P per shrink-loop trip), so a seeded caller sees the same stream consumption as with the reference.
Batched multi-chain sampling lives in :mod:`..chains`.
"""
import numpy as np
import scipy.special

from .. import ops

MAX_TRIPS = 256


def log_gamma(x, k, theta, invG):
    """Log pdf of the Gamma (and, for the last entry, inverse-Gamma) hyper-priors and their gradients.

    O(1) host arithmetic, same expressions as ``sliceSample.py:224-230`` (the device copy used inside the sweep is
    ``log_prior_entry`` in ``csrc/sds.cu``)."""
    x = np.asarray(x, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    logG = (k - 1) * np.log(x) - x / theta - k * np.log(theta) - np.log(scipy.special.gamma(k))
    gradG = (k - 1) * (1 / x) - 1 / theta
    if invG:
        j = x.shape[0] - 1 if x.shape[0] != 3 else 2
        logG[j] = np.log(theta[j] ** k[j]) - np.log(scipy.special.gamma(k[j])) + (-k[j] - 1) * np.log(x[j]) + (-theta[j] / x[j])
        gradG[j] = (-k[j] - 1) / x[j] + theta[j] / (x[j] ** 2)
    return logG, gradG


def surrogate_slice_sampling(f, x, y, hyp, scale, iter=0):
    """One surrogate-data slice-sampling update of ``(f, hyp)`` -- same contract as ``sliceSample.py:76``."""
    import torch
    f = np.array(f, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    hyp = np.array(hyp, dtype=np.float64).reshape(-1)
    scale = np.asarray(scale, dtype=np.float64).reshape(-1)
    n, P = f.shape[0], hyp.shape[0]
    # the reference's draw order: g (:194), bracket (:110), threshold (:127), then P uniforms per trip (:132)
    z = np.random.standard_normal(n)
    v = np.random.random_sample(P)
    u0 = np.random.random_sample()
    state = np.random.get_state()
    U = np.random.random_sample((MAX_TRIPS, P))
    F = torch.tensor(f[None]).cuda()
    H = torch.tensor(hyp[None]).cuda()
    ntrips, _, status = ops.sds_sweep(x, y, F, H, scale, iter, tape=ops.Tape(z[None], v[None], [u0], U[None]), max_trips=MAX_TRIPS)
    trips = int(ntrips.item())
    # leave the global stream where the reference would have left it: exactly `trips` proposal draws consumed
    np.random.set_state(state)
    if trips:
        np.random.random_sample((trips, P))
    if int(status.item()) != 0:
        raise RuntimeError('slice did not close within %d proposals' % MAX_TRIPS)
    return F.cpu().numpy()[0], H.cpu().numpy()[0]


def aux_var_model(f, K, sn, g=None):
    """Auxiliary-variable model of the SDS sampler for a caller-supplied covariance ``K`` -- same contract as
    ``sliceSample.py:165``: returns ``(g, K+S, m_theta_g, chol_R_theta, L)``.

    ``S`` is the O(N) diagonal expression of ``:183-190`` (host); ``g`` is drawn as ``f + sqrt(S_ii) z`` from the global
    numpy stream when not given (``:194`` consumes the same N normals through ``multivariate_normal``); the two
    Cholesky factorisations, ``R`` and ``m`` run on the device (``gpmc_aux_var_model``).  A factorisation that fails is
    reported as ``numpy.linalg.LinAlgError`` -- what the reference's ``jitchol`` raises when its ladder gives up."""
    f = np.asarray(f, dtype=np.float64).reshape(-1)
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    Kii = np.diagonal(K)
    with np.errstate(divide='ignore', invalid='ignore'):
        K_ii_inv = 1. / Kii
        v_1 = (sn ** 2) ** (-1) + K_ii_inv
        Sii = np.maximum(1. / (v_1 - K_ii_inv), 0.)
    if g is None:
        g = f + np.sqrt(Sii) * np.random.standard_normal(n)
    L, m, C, info = ops.aux_var_model_device(K, Sii, g)
    info = info.cpu().numpy()
    if info[0] != 0 or info[1] != 0:
        raise np.linalg.LinAlgError('not positive definite (chol(K+S) info=%d, chol(R+1e-11 I) info=%d)' % (info[0], info[1]))
    return g, K + np.diag(Sii), m.cpu().numpy(), C.cpu().numpy(), L.cpu().numpy()


def elliptical_slice(f, x, y, hyp):
    """Elliptical slice sampling update of ``f`` -- same contract as ``sliceSample.py:15-74`` (dead code in the reference:
    both call sites are commented out).  ``nu ~ N(0, K)`` is drawn as ``chol(K) z`` on the device; the reference draws it
    through numpy's SVD route (``:41``), so the two agree in distribution, not sample by sample."""
    from ..kcGP import covK, likK, tools
    f = np.asarray(f, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    nobs = f.shape[0]
    my = np.mean(y)
    K = covK.RBF(np.log(hyp[0]), np.log(hyp[1])).getCovMatrix(x=x, mode='train')
    nu = np.dot(tools.jitchol(K), np.random.standard_normal(nobs))
    lik_func = likK.TruncatedGauss2(upper=100 - my, lower=0 - my, log_sigma=np.log(hyp[2]))
    cur_llk = lik_func.evaluate(y=y - my, mu=f) + np.log(np.random.uniform())
    theta = np.random.uniform(high=2. * np.pi)
    theta_min, theta_max = theta - 2. * np.pi, theta
    while True:
        prop_f = f * np.cos(theta) + nu * np.sin(theta)
        prop_llk = lik_func.evaluate(y=y - my, mu=prop_f)
        if prop_llk > cur_llk and np.isfinite(prop_llk):
            return prop_f
        if theta >= 0:
            theta_max = theta
        else:
            theta_min = theta
        theta = np.random.uniform(low=theta_min, high=theta_max)


def inf_mcmc(f, model, ys=0):
    """Predictive inference ``fs | f`` from stored MCMC samples -- same contract as ``sliceSample.py:234-284``
    (``model`` has ``x, y, xs, meanfunc, covfunc, likfunc``); every O(N^3)/O(N^2) step goes through the GPU-backed
    ``kcGP`` primitives (``jitchol``, ``solve_chol``, ``getCovMatrix``), the rest is the reference's own array glue."""
    from ..kcGP import tools
    x, y, xs = model.x, model.y, model.xs
    my = np.mean(y)
    n_samples = f.shape[1]
    ns = xs.shape[0]
    n, D = x.shape
    m = np.tile(model.meanfunc.getMean(x), (1, n_samples))
    K = model.covfunc.getCovMatrix(x=x, mode='train')
    sn2 = model.likfunc.sn ** 2.
    L = tools.jitchol(K / sn2 + np.eye(n)).T                        # upper
    alpha = tools.solve_chol(L, f - m) / sn2
    sW = np.ones((n, 1)) / np.sqrt(sn2)
    kss = model.covfunc.getCovMatrix(z=xs, mode='self_test')
    Ks = model.covfunc.getCovMatrix(x=x, z=xs, mode='cross')
    ms = model.meanfunc.getMean(xs)
    Fmu = np.tile(ms, (1, n_samples)) + np.dot(Ks.T, alpha)
    V = ops.trsv_lower(np.ascontiguousarray(L.T), np.ascontiguousarray((np.tile(sW, (1, ns)) * Ks).T)).cpu().numpy().T
    fs2 = kss - np.array([(V * V).sum(axis=0)]).T
    Fs2 = np.maximum(fs2, 0)
    Fmu = np.mean(Fmu, axis=1, keepdims=True)
    Ymu, Lower, Upper = model.likfunc.evaluate(mu=Fmu, s2=Fs2)
    ym = np.reshape(np.mean(Ymu, axis=1), (ns, 1)) + my
    ys_lw = np.reshape(np.mean(Lower, axis=1), (ns, 1)) + my
    ys_up = np.reshape(np.mean(Upper, axis=1), (ns, 1)) + my
    return ym, ys_lw, ys_up, Fs2
