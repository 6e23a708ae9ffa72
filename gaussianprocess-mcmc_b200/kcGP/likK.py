"""``kcGP.likK``: the truncated-Gaussian likelihood used by the samplers (``sliceSample.py:47,117-118,142-143,279``).

``TruncatedGauss2`` is private to the reference's author; its definition here is ASSUMPTION-1 of ``oracle/kcgp_shim.py``
(summed log density of ``y`` under ``N(mu, sn^2)`` truncated to ``[lower, upper]``).  The ``(y, mu)`` mode -- the one on
the sampler's path -- runs on the device; the predictive ``(mu, s2)`` mode is O(ns) host arithmetic."""
import numpy as np
import scipy.special

from .. import ops


class TruncatedGauss2(object):
    def __init__(self, upper=1., lower=0., log_sigma=np.log(0.1)):
        self.upper = upper
        self.lower = lower
        self.sn = np.exp(log_sigma)

    def evaluate(self, y=None, mu=None, s2=None):
        if y is not None and s2 is None:
            y = np.asarray(y, dtype=np.float64).reshape(-1)
            mu = np.asarray(mu, dtype=np.float64).reshape(-1)
            return float(ops.tg2_loglik(y, mu[None], self.sn, self.lower, self.upper).item())
        mu = np.asarray(mu, dtype=np.float64)
        s = np.sqrt(np.asarray(s2, dtype=np.float64) + self.sn ** 2)
        a, b = (self.lower - mu) / s, (self.upper - mu) / s
        Z = scipy.special.ndtr(b) - scipy.special.ndtr(a)
        pdf = lambda t: np.exp(-0.5 * t * t) / np.sqrt(2. * np.pi)
        if y is not None:
            yy = np.asarray(y, dtype=np.float64).reshape(mu.shape)
            r = (yy - mu) / s
            return float(np.sum(-0.5 * r * r - 0.5 * np.log(2. * np.pi) - np.log(s) - np.log(Z)))
        Fa = scipy.special.ndtr(a)
        return (mu + s * (pdf(a) - pdf(b)) / Z, mu + s * scipy.special.ndtri(Fa + 0.025 * Z),
                mu + s * scipy.special.ndtri(Fa + 0.975 * Z))


class Gauss(object):
    def __init__(self, log_sigma=np.log(0.1)):
        self.sn = np.exp(log_sigma)

    def evaluate(self, y=None, mu=None, s2=None):
        sn2 = self.sn ** 2 + (0. if s2 is None else np.asarray(s2))
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        mu = np.asarray(mu, dtype=np.float64).reshape(-1)
        return float(np.sum(-(y - mu) ** 2 / (2. * sn2) - 0.5 * np.log(2. * np.pi * sn2)))
