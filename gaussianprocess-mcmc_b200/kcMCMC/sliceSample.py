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
