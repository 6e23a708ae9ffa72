"""Synthetic inputs shaped like the reference's data (host-side numpy; data preparation only).

The reference ships no data (``houston/``, ``bryan/``, ``output/`` are git-ignored), so the
only template is its synthetic generator, ``demoRegression.py:117-136``: seed 124, unit-spaced
1-D grid, ``y = chol(K + sn^2 I) z + 91.1538461538`` with ``(ll, sf, sn) = (5, 20, 2.5)``.
With ``sf = 20`` that series leaves the [0, 100] support of the truncated likelihood
(``sliceSample.py:114-117``; PMIS condition scores live in [0, 100]), so this generator uses
``sf = 4`` and clips to [0, 100] -- the "IH45-traffic-shaped" series of BASELINE.json.
"""
import numpy as np

IH45_MEAN = 91.1538461538          # demoRegression.py:129
TRUE_HYP = (5.0, 4.0, 2.5)         # (ll, sf, sn); demoRegression.py:119-121 with sf 20 -> 4
HYP0 = (1.0, 10.0, 1.2)            # framework.py:63
SCALE = (10.0, 10.0, 5.0)          # framework.py:69 / demoRegression.py:25


def ih45_series(n, seed=124, hyp=TRUE_HYP):
    """Return ``(x[n,1], y[n])``: unit-spaced grid and a clipped GP draw around the IH45 mean."""
    rs = np.random.RandomState(seed)
    ll, sf, sn = hyp
    x = np.arange(0, n, dtype=np.float64).reshape(n, 1)
    z = rs.normal(size=(n,))
    # the reference's recipe: lower Cholesky of K + sn^2 I applied to z
    d = (x / ll - (x / ll).T) ** 2
    K = sf * sf * np.exp(-0.5 * d)
    K[np.diag_indices(n)] += sn * sn
    L = np.linalg.cholesky(K)
    y = np.clip(L.dot(z) + IH45_MEAN, 0.0, 100.0)
    return x, y


def ard_inputs(n, d=4, seed=124):
    """BASELINE config 3: ``x ~ U(0,10)^{n x d}``, ``y`` a clipped draw from an ARD GP."""
    rs = np.random.RandomState(seed)
    x = rs.uniform(0.0, 10.0, size=(n, d))
    ell = np.array([2.0, 3.0, 4.0, 5.0][:d] + [3.0] * max(0, d - 4))
    u = x / ell
    sq = ((u[:, None, :] - u[None, :, :]) ** 2).sum(-1)
    K = 16.0 * np.exp(-0.5 * sq)
    K[np.diag_indices(n)] += 2.5 ** 2
    y = np.clip(np.linalg.cholesky(K).dot(rs.normal(size=(n,))) + IH45_MEAN, 0.0, 100.0)
    return x, y


def chain_states(n_chains, n, n_ell=1, seed=1000, first_chain=0):
    """Initial per-chain state: ``f0 = 0`` and ``hyp0 = HYP0 * U(0.5, 2)`` from seed ``1000 + c``
    (keyed by GLOBAL chain id so the result does not depend on how chains are sharded)."""
    F = np.zeros((n_chains, n))
    base = np.array([HYP0[0]] * n_ell + [HYP0[1], HYP0[2]])
    H = np.empty((n_chains, n_ell + 2))
    for c in range(n_chains):
        rs = np.random.RandomState(seed + first_chain + c)
        H[c] = base * rs.uniform(0.5, 2.0, size=n_ell + 2)
    return F, H


def loglik_batch(n_chains, n, n_ell=1, seed=2000, first_chain=0):
    """Inputs for pure log-lik throughput: per chain a surrogate vector ``g = sn*z`` and a
    proposal ``theta ~ U(bracket)`` with the bracket of ``sliceSample.py:110-112`` around HYP0."""
    P = n_ell + 2
    scale = np.array([SCALE[0]] * n_ell + [SCALE[1], SCALE[2]])
    base = np.array([HYP0[0]] * n_ell + [HYP0[1], HYP0[2]])
    G = np.empty((n_chains, n))
    H = np.empty((n_chains, P))
    for c in range(n_chains):
        rs = np.random.RandomState(seed + first_chain + c)
        v = rs.uniform(0.0, scale)
        lo = np.maximum(base - v, 0.0)
        H[c] = np.maximum(lo + scale * rs.uniform(size=P), 0.05)
        G[c] = H[c, P - 1] * rs.standard_normal(n)
    return G, H
