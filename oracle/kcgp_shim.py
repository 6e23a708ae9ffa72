"""``kcGP`` shim: the GP primitives ``kcMCMC/sliceSample.py:13`` imports but the
reference tree does not ship (``.gitignore:12`` hides ``kcGP/``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED at this
boundary: ``kcGP`` is a private fork of pyGPs 1.3.4 (``requirements.txt:10``,
numpy 1.12.1 / scipy 0.19.0 era).  pyGPs is absent from this image, so the
published pyGPs-1.3.4 algorithms are restated here:

* ``covK.RBF``          <- pyGPs/Core/cov.py  class RBF  (``getCovMatrix``)
* ``covK.RBFard``       <- pyGPs/Core/cov.py  class RBFard (used for the ARD config,
                           which the reference itself never calls)
* ``tools.jitchol``     <- pyGPs/Core/tools.py jitchol   (GPy-derived jitter ladder)
* ``tools.solve_chol``  <- pyGPs/Core/tools.py solve_chol
* ``likK.Gauss``        <- pyGPs/Core/lik.py  class Gauss (log-density only)
* ``likK.TruncatedGauss2`` <- author-private, no published source: ASSUMPTION-1.

Call sites that fix the API shape: ``sliceSample.py:38-39,104-105,136-137``
(RBF), ``:196,205,257`` (jitchol), ``:258`` (solve_chol), ``:47,117-118,142-143,279``
(TruncatedGauss2).
"""
import sys
import types

import numpy as np
import scipy.linalg
import scipy.special
from scipy.spatial.distance import cdist

_SQRT2 = np.sqrt(2.0)
_HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


# --------------------------------------------------------------------------- covK
class RBF(object):
    """Squared-exponential kernel, isotropic length-scale (pyGPs 1.3.4 ``cov.RBF``).

    ``hyp = [log_ell, log_sigma]``; ``K = sf2 * exp(-0.5 * sqdist(x/ell, z/ell))``
    with ``ell = exp(hyp[0])``, ``sf2 = exp(2*hyp[1])``.  The reference builds it as
    ``covK.RBF(np.log(hyp[0]), np.log(hyp[1]))`` (``sliceSample.py:104,136``), so
    ``sf2 = hyp[1]**2`` up to the exp/log round trip, which is kept.
    """

    def __init__(self, log_ell=0., log_sigma=0.):
        self.hyp = [log_ell, log_sigma]

    def getCovMatrix(self, x=None, z=None, mode=None):
        ell = np.exp(self.hyp[0])
        sf2 = np.exp(2. * self.hyp[1])
        if mode == 'self_test':
            nn = z.shape[0]
            A = np.zeros((nn, 1))
        elif mode == 'train':
            xs = np.asarray(x, dtype=np.float64) / ell
            if xs.ndim == 1:
                xs = xs.reshape(-1, 1)
            A = cdist(xs, xs, 'sqeuclidean')
        elif mode == 'cross':
            xs = np.asarray(x, dtype=np.float64) / ell
            zs = np.asarray(z, dtype=np.float64) / ell
            A = cdist(xs, zs, 'sqeuclidean')
        else:
            raise ValueError("mode must be 'train', 'cross' or 'self_test'")
        return sf2 * np.exp(-0.5 * A)


class RBFard(object):
    """SE kernel with one length-scale per input dimension (pyGPs 1.3.4 ``cov.RBFard``).

    ``hyp = log_ell_list + [log_sigma]``; columns of ``x`` are divided by ``ell_d``
    before the squared distance.  Not called by the reference; it is the oracle for
    BASELINE config 3 (ARD, D=4)."""

    def __init__(self, D=None, log_ell_list=None, log_sigma=0.):
        if log_ell_list is None:
            log_ell_list = [0. for _ in range(D)]
        self.hyp = list(log_ell_list) + [log_sigma]

    def getCovMatrix(self, x=None, z=None, mode=None):
        ell = np.exp(np.asarray(self.hyp[:-1], dtype=np.float64))
        sf2 = np.exp(2. * self.hyp[-1])
        if mode == 'self_test':
            A = np.zeros((z.shape[0], 1))
        elif mode == 'train':
            xs = np.asarray(x, dtype=np.float64) / ell
            A = cdist(xs, xs, 'sqeuclidean')
        elif mode == 'cross':
            A = cdist(np.asarray(x, dtype=np.float64) / ell,
                      np.asarray(z, dtype=np.float64) / ell, 'sqeuclidean')
        else:
            raise ValueError("mode must be 'train', 'cross' or 'self_test'")
        return sf2 * np.exp(-0.5 * A)


# -------------------------------------------------------------------------- tools
def jitchol(A, maxtries=5):
    """Lower Cholesky with the pyGPs/GPy jitter ladder.

    ``dpotrf(lower=1)``; if it fails and every diagonal entry is positive, retry
    with ``A + jitter*I``, ``jitter = mean(diag A)*1e-6`` growing tenfold, at most
    ``maxtries`` times; then raise ``LinAlgError``.  (Documented intent of
    pyGPs 1.3.4 ``tools.jitchol``; the retry line of that release passes a kwarg
    numpy 1.12 rejected, so the failure path of the original is itself unpinned.)
    Called at ``sliceSample.py:196,205,257``.
    """
    A = np.ascontiguousarray(A)
    L, info = scipy.linalg.lapack.dpotrf(A, lower=1)
    if info == 0:
        return np.tril(L)
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        L, info = scipy.linalg.lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return np.tril(L)
        jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def solve_chol(L, B):
    """``(L^T L)^{-1} B`` for UPPER ``L`` (pyGPs ``tools.solve_chol``; ``sliceSample.py:258``)."""
    return np.linalg.solve(L, np.linalg.solve(L.T, B))


# --------------------------------------------------------------------------- likK
def _log_trunc_mass(a, b):
    """log(Phi(b) - Phi(a)) for a < b, written so that both tails keep precision.

    The product's CUDA device function ``tg2_log_mass`` uses the same three-branch
    form (erfc on the side where both limits lie, erf across the origin)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.empty(np.broadcast(a, b).shape, dtype=np.float64)
    a, b = np.broadcast_arrays(a, b)
    pos = a > 0.
    neg = b < 0.
    mid = ~(pos | neg)
    with np.errstate(divide='ignore', invalid='ignore'):
        out[pos] = np.log(0.5 * (scipy.special.erfc(a[pos] / _SQRT2) - scipy.special.erfc(b[pos] / _SQRT2)))
        out[neg] = np.log(0.5 * (scipy.special.erfc(-b[neg] / _SQRT2) - scipy.special.erfc(-a[neg] / _SQRT2)))
        out[mid] = np.log(0.5 * (scipy.special.erf(b[mid] / _SQRT2) - scipy.special.erf(a[mid] / _SQRT2)))
    return out


class TruncatedGauss2(object):
    """ASSUMPTION-1 -- author-private likelihood, no source anywhere.

    Inferred from the call sites: constructed with ``upper=, lower=, log_sigma=``
    (``sliceSample.py:117``), has a mutable natural-scale ``.sn`` (``:142``) and
    ``.upper/.lower`` (``framework.py:241-242``); ``evaluate(y=, mu=)`` is compared
    and ``isfinite``-tested as a scalar (``:64,154``), so it is the summed
    log-density of ``y_i`` under ``N(mu_i, sn^2)`` truncated to ``[lower, upper]``::

        sum_i [ -0.5*((y_i-mu_i)/sn)^2 - 0.5*log(2*pi) - log(sn)
                - log(Phi((upper-mu_i)/sn) - Phi((lower-mu_i)/sn)) ]

    ``evaluate(mu=, s2=)`` (predictive mode, ``:279``) returns ``(Ymu, Lower, Upper)``:
    the mean of the truncated predictive ``N(mu, s2+sn^2)`` and its central 95% band.
    """

    def __init__(self, upper=1., lower=0., log_sigma=np.log(0.1)):
        self.upper = upper
        self.lower = lower
        self.sn = np.exp(log_sigma)

    def evaluate(self, y=None, mu=None, s2=None):
        if y is not None and s2 is None:
            y = np.asarray(y, dtype=np.float64).reshape(-1)
            mu = np.asarray(mu, dtype=np.float64).reshape(-1)
            sn = self.sn
            with np.errstate(divide='ignore', invalid='ignore'):
                r = (y - mu) / sn
                a = (self.lower - mu) / sn
                b = (self.upper - mu) / sn
                lp = -0.5 * r * r - _HALF_LOG_2PI - np.log(sn) - _log_trunc_mass(a, b)
            return float(np.sum(lp))
        # predictive mode
        mu = np.asarray(mu, dtype=np.float64)
        s = np.sqrt(np.asarray(s2, dtype=np.float64) + self.sn ** 2)
        a = (self.lower - mu) / s
        b = (self.upper - mu) / s
        Z = scipy.special.ndtr(b) - scipy.special.ndtr(a)
        pdf = lambda t: np.exp(-0.5 * t * t) / np.sqrt(2. * np.pi)
        Ymu = mu + s * (pdf(a) - pdf(b)) / Z
        Fa = scipy.special.ndtr(a)
        Lower = mu + s * scipy.special.ndtri(Fa + 0.025 * Z)
        Upper = mu + s * scipy.special.ndtri(Fa + 0.975 * Z)
        if y is not None:
            y = np.asarray(y, dtype=np.float64).reshape(mu.shape)
            r = (y - mu) / s
            return float(np.sum(-0.5 * r * r - _HALF_LOG_2PI - np.log(s) - np.log(Z)))
        return Ymu, Lower, Upper


class Gauss(object):
    """pyGPs 1.3.4 ``lik.Gauss`` log-density (``framework.py:263``; commented alternative at ``sliceSample.py:48``)."""

    def __init__(self, log_sigma=np.log(0.1)):
        self.sn = np.exp(log_sigma)

    def evaluate(self, y=None, mu=None, s2=None):
        sn2 = self.sn ** 2 + (0. if s2 is None else np.asarray(s2))
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        mu = np.asarray(mu, dtype=np.float64).reshape(-1)
        return float(np.sum(-(y - mu) ** 2 / (2. * sn2) - 0.5 * np.log(2. * np.pi * sn2)))


# ------------------------------------------------------------------ module wiring
def make_modules():
    """Build ``kcGP``, ``kcGP.covK``, ``kcGP.likK``, ``kcGP.tools`` module objects."""
    kcGP = types.ModuleType('kcGP')
    covK = types.ModuleType('kcGP.covK')
    likK = types.ModuleType('kcGP.likK')
    tools = types.ModuleType('kcGP.tools')
    covK.RBF, covK.RBFard = RBF, RBFard
    likK.TruncatedGauss2, likK.Gauss = TruncatedGauss2, Gauss
    tools.jitchol, tools.solve_chol = jitchol, solve_chol
    kcGP.covK, kcGP.likK, kcGP.tools = covK, likK, tools
    return {'kcGP': kcGP, 'kcGP.covK': covK, 'kcGP.likK': likK, 'kcGP.tools': tools}


def install():
    """Inject the shim into ``sys.modules`` so ``from kcGP import covK, likK, tools`` resolves."""
    mods = make_modules()
    for name, mod in mods.items():
        sys.modules.setdefault(name, mod)
    return sys.modules['kcGP']
