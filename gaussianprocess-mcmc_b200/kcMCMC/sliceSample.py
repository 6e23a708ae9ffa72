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
    both call sites are commented out).  The whole update runs on the device (``gpmc_ess_sweep``): ``nu = chol(K) z`` and
    the bracket-shrinking loop on the ellipse.  The reference draws ``nu ~ N(0, K)`` through numpy's SVD route (``:41``),
    so the two agree in distribution, not sample by sample; the global numpy stream is consumed like the reference does
    (N normals, then one uniform for the slice level, one for the angle, one per rejected proposal)."""
    import torch
    f = np.array(f, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    hyp = np.asarray(hyp, dtype=np.float64).reshape(-1)
    z = np.random.standard_normal(f.shape[0])
    u = np.random.random_sample()
    state = np.random.get_state()
    th = np.random.random_sample(MAX_TRIPS)
    F = torch.tensor(f[None]).cuda()
    ntrips, status, info = ops.ess_sweep(x, y, F, hyp[None], tape=ops.EssTape([u], th[None], z=z[None]), max_trips=MAX_TRIPS)
    np.random.set_state(state)
    np.random.random_sample(int(ntrips.item()))          # the angle + one redraw per rejected proposal
    st = int(status.item())
    if st == 2:
        raise np.linalg.LinAlgError('not positive definite, even with jitter.')
    if st != 0:
        raise RuntimeError('slice did not close within %d proposals' % MAX_TRIPS)
    return F.cpu().numpy()[0]


def inf_mcmc_batched(F, Hyp, x, y, xs, lik_factory, mean_x=None, mean_xs=None):
    """``inf_mcmc`` (``sliceSample.py:234-284``) for S stored samples with their own hyper-parameters in ONE device pass
    (``gpmc_predict_batched``) -- what the loops at ``framework.py:223-243`` / ``plotResult.py:98-121`` do one sample at a
    time.  ``F[N, S]`` latent samples, ``Hyp[S, P]``; ``lik_factory(sn) -> likfunc`` builds the predictive likelihood of a
    sample (``likK.TruncatedGauss2``); ``mean_x[N,1]`` / ``mean_xs[ns,1]`` are the mean function's values (zero when
    omitted).  Returns arrays ``(ym[S, ns, 1], ys_lw[S, ns, 1], ys_up[S, ns, 1], Fs2[S, ns, 1])``."""
    F = np.asarray(F, dtype=np.float64)
    Hyp = np.atleast_2d(np.asarray(Hyp, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64)
    my = np.mean(y)
    S, ns = Hyp.shape[0], np.asarray(xs).shape[0]
    m = np.zeros((F.shape[0], 1)) if mean_x is None else np.asarray(mean_x, dtype=np.float64).reshape(-1, 1)
    ms = np.zeros((ns, 1)) if mean_xs is None else np.asarray(mean_xs, dtype=np.float64).reshape(-1, 1)
    fmu, fs2, info = ops.predict_batched(x, xs, np.ascontiguousarray((F - m).T), Hyp)
    info = info.cpu().numpy()
    if np.any(info != 0):
        raise np.linalg.LinAlgError('not positive definite, even with jitter (sample %d)' % int(np.flatnonzero(info)[0]))
    fmu, fs2 = fmu.cpu().numpy(), fs2.cpu().numpy()
    out = [np.zeros((S, ns, 1)) for _ in range(4)]
    for s in range(S):
        Fmu = ms + fmu[s].reshape(ns, 1)                                       # :266
        Fs2 = np.maximum(fs2[s].reshape(ns, 1), 0)                             # :273
        Ymu, Lower, Upper = lik_factory(Hyp[s, -1]).evaluate(mu=Fmu, s2=Fs2)   # :279
        out[0][s] = np.reshape(np.mean(Ymu, axis=1), (ns, 1)) + my
        out[1][s] = np.reshape(np.mean(Lower, axis=1), (ns, 1)) + my
        out[2][s] = np.reshape(np.mean(Upper, axis=1), (ns, 1)) + my
        out[3][s] = Fs2
    return tuple(out)


def inf_mcmc(f, model, ys=0):
    """Predictive inference ``fs | f`` from stored MCMC samples -- same contract as ``sliceSample.py:234-284``
    (``model`` has ``x, y, xs, meanfunc, covfunc, likfunc``; ``f[N, n_samples]`` share the model's hyper-parameters).
    The prediction is linear in ``f`` and the reference averages it over the samples (``:275``), so one right-hand side
    -- the sample mean -- goes through the batched device path."""
    f = np.asarray(f, dtype=np.float64)
    if f.ndim == 1:
        f = f.reshape(-1, 1)
    cov_hyp = np.exp(np.asarray(model.covfunc.hyp, dtype=np.float64))          # (ell.., sf) natural scale
    hyp = np.concatenate([cov_hyp, [model.likfunc.sn]])
    fbar = np.mean(f, axis=1, keepdims=True)
    likfunc = model.likfunc
    ym, lw, up, Fs2 = inf_mcmc_batched(fbar, hyp[None], model.x, model.y, model.xs, lambda sn: likfunc,
                                       mean_x=model.meanfunc.getMean(model.x), mean_xs=model.meanfunc.getMean(model.xs))
    return ym[0], lw[0], up[0], Fs2[0]
