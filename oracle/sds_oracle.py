"""CPU restatement (numpy/scipy, FP64) of the reference hot path, driven by explicit randomness.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the checker the CUDA path is compared
with, and the CPU arm ``bench.py`` times beside it.  Never imported by the product.

Every function cites the reference lines it follows (paths relative to the reference
root).  The restatement keeps the reference's expression order wherever rounding is
visible at the 1e-10 parity target (``S_ii`` round trip ``sliceSample.py:184-187``,
``+1e-11*I`` ``:205``, dense-``inv`` quadratic form ``:147``) and is pinned to the
literal file by ``tests/test_oracle_vs_reference.py`` + ``tests/golden/*.npz``.

The one generalisation: the reference hard-codes P=3 hyper-parameters ``(ll, sf, sn)``
(``:124-125,159``); here ``hyp = (ell_1..ell_D', sf, sn)`` with ``D' = 1`` (isotropic,
the reference) or ``D' = D`` (ARD, BASELINE config 3), identical for P=3.
"""
import numpy as np
import scipy.linalg
import scipy.special

from . import kcgp_shim

LOG_2PI = np.log(2 * np.pi)
BURN_IN = 500                      # sliceSample.py:128,133,151
PRIOR_K3 = np.asarray([1., 3., 3.])        # sliceSample.py:124
PRIOR_THETA3 = np.asarray([1., 1.5, 3.])   # sliceSample.py:125


def prior_constants(n_ell):
    """Prior (k, theta) for P = n_ell + 2: every length-scale gets the reference's Gamma(1,1),
    sf Gamma(3,1.5), sn inverse-Gamma(3,3) (``sliceSample.py:124-125``)."""
    k = np.concatenate([np.full(n_ell, PRIOR_K3[0]), PRIOR_K3[1:]])
    th = np.concatenate([np.full(n_ell, PRIOR_THETA3[0]), PRIOR_THETA3[1:]])
    return k, th


# ------------------------------------------------------------------ a2: covariance
def cov_matrix(x, hyp):
    """``covK.RBF(np.log(ll), np.log(sf)).getCovMatrix(x, mode='train')`` (``sliceSample.py:104-105,136-137``).

    ``hyp`` natural scale; the log/exp round trip of the reference call is kept."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    hyp = np.asarray(hyp, dtype=np.float64)
    n_ell = hyp.shape[0] - 2
    with np.errstate(divide='ignore', invalid='ignore'):
        if n_ell == 1:
            return kcgp_shim.RBF(np.log(hyp[0]), np.log(hyp[1])).getCovMatrix(x=x, mode='train')
        return kcgp_shim.RBFard(log_ell_list=list(np.log(hyp[:n_ell])), log_sigma=np.log(hyp[n_ell])).getCovMatrix(x=x, mode='train')


def s_diagonal(Kii, sn):
    """``S_ii = 1/((1/sn^2 + 1/K_ii) - 1/K_ii)``, clamped at 0 (``sliceSample.py:184-190``)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        K_ii_inv = 1. / Kii
        v_1 = (sn ** 2) ** (-1) + K_ii_inv
        Sii = 1. / (v_1 - K_ii_inv)
    return np.maximum(Sii, 0.)


# ------------------------------------------------------------- a3: aux_var_model
def aux_var_model(f, K, sn, g=None, z=None, r_form='literal'):
    """``aux_var_model`` (``sliceSample.py:165-207``) with the draw of ``g`` taken from ``z``.

    ``r_form='literal'``: ``R = K - V^T V`` with ``V = solve(L, K)`` exactly as ``:197-198``.
    ``r_form='reduced'``: the same matrix as ``S - S (K+S)^-1 S`` and ``m = g - S (K+S)^-1 g`` (what the CUDA
    path evaluates).  ``chol(R + 1e-11 I)`` has condition ~1e11, so the two forms give ``C`` (hence ``f'``) that
    differ far above rounding; the difference between them is the resolution to which the reference itself
    defines ``f'`` (see tests/test_gpu_sds.py).

    ``:194`` draws ``g ~ N(f, S)`` through ``multivariate_normal`` (an SVD of the diagonal
    S); that equals ``f + sqrt(S_ii) z_i`` with ``z`` in draw order (pinned in make_golden).
    Returns ``(g, K+S, m_theta_g, chol_R_theta, L)`` like the reference."""
    n = K.shape[0]
    Sii = s_diagonal(np.diagonal(K), sn)                       # :184-187,190
    S = np.zeros_like(K)
    np.fill_diagonal(S, Sii)                                   # :188-189
    if g is None:
        g = f + np.sqrt(Sii) * z                               # :194
    L = kcgp_shim.jitchol(K + S)                               # :196
    if r_form == 'literal':
        V = np.linalg.solve(L, K)                              # :197
        R_theta = K - np.dot(V.T, V)                           # :198
        with np.errstate(divide='ignore', invalid='ignore'):
            m_theta_g = np.dot(np.dot(R_theta, np.linalg.inv(S)), g)   # :204
    else:
        U = scipy.linalg.solve_triangular(L, np.eye(n), lower=True).T      # U = L^-T
        P = np.dot(U, U.T)                                     # (K+S)^-1
        R_theta = np.diag(Sii) - Sii[:, None] * P * Sii[None, :]
        zz = scipy.linalg.solve_triangular(L, g, lower=True)
        m_theta_g = g - Sii * np.dot(U, zz)
    chol_R_theta = kcgp_shim.jitchol(R_theta + np.eye(n) * 1e-11)  # :205
    return g, K + S, m_theta_g, chol_R_theta, L


# ------------------------------------------------------------- a5: log marginal
def log_marginal_inv_form(g, K_S, L_ks):
    """``propG`` exactly as written at ``sliceSample.py:147`` (dense ``inv``)."""
    return -(np.dot(np.dot(g.T, np.linalg.inv(K_S)), g) / 2. + np.log(np.diag(L_ks.T)).sum() + g.shape[0] * np.log(2 * np.pi) / 2.)


def log_marginal_chol_form(g, L_ks):
    """The author's commented alternative (``sliceSample.py:145-146``): alpha via ``solve_chol``."""
    alpha = kcgp_shim.solve_chol(L_ks.T, g)
    return -(np.dot(g.T, alpha) / 2. + np.log(np.diag(L_ks.T)).sum() + g.shape[0] * np.log(2 * np.pi) / 2.)


def loglik_unit(x, g, hyp, form='chol'):
    """The metric's unit, "one GP log-lik eval" (SURVEY 8a rows a2 + S-diag + a4 + a5):
    assemble ``K+S`` from ``(x, hyp)``, Cholesky, quadratic form, log-det -> scalar.

    ``form='inv'``  : quadratic form as the reference writes it (``:147``)
    ``form='chol'`` : ``solve_chol`` form (``:145-146``)
    ``form='trsv'`` : one forward substitution ``|L^-1 g|^2`` (what the CUDA path does)
    Raises ``LinAlgError`` when ``jitchol`` gives up."""
    K = cov_matrix(x, hyp)
    Sii = s_diagonal(np.diagonal(K), hyp[-1])
    K_S = K + np.diag(Sii)
    L = kcgp_shim.jitchol(K_S)
    if form == 'inv':
        return float(log_marginal_inv_form(g, K_S, L))
    if form == 'chol':
        return float(log_marginal_chol_form(g, L))
    w = scipy.linalg.solve_triangular(L, g, lower=True)
    return float(-(np.dot(w, w) / 2. + np.log(np.diag(L)).sum() + g.shape[0] * LOG_2PI / 2.))


# ------------------------------------------------------------------ a7: priors
def log_gamma(x, k, theta, invG):
    """``log_gamma`` (``sliceSample.py:209-232``): Gamma log-pdf, last entry inverse-Gamma.

    The reference indexes the inverse-Gamma entry as ``[2]``; for P != 3 it is the last."""
    x = np.asarray(x, dtype=np.float64)
    j = x.shape[0] - 1
    with np.errstate(divide='ignore', invalid='ignore'):
        logG = (k - 1) * np.log(x) - x / theta - k * np.log(theta) - np.log(scipy.special.gamma(k))   # :224
        gradG = (k - 1) * (1 / x) - 1 / theta                                                      # :225
        if invG:
            logG[j] = np.log(theta[j] ** k[j]) - np.log(scipy.special.gamma(k[j])) + (-k[j] - 1) * np.log(x[j]) + (-theta[j] / x[j])   # :229
            gradG[j] = (-k[j] - 1) / x[j] + theta[j] / (x[j] ** 2)                                    # :230
    return logG, gradG


# ---------------------------------------------------------- a6: TruncatedGauss2
def trunc_gauss2_loglik(y_centered, mu, sn, lower, upper):
    """``lik_func.evaluate(y=y-my, mu=f)`` (``sliceSample.py:118,143``) under ASSUMPTION-1."""
    lik = kcgp_shim.TruncatedGauss2(upper=upper, lower=lower, log_sigma=0.)
    lik.sn = sn                                                # :142 sets natural-scale sn
    return lik.evaluate(y=y_centered, mu=mu)


# ------------------------------------------------ a1/a8/a9: one SDS transition

def _sum_density(head, prior, G, n_ell):
    """``head + prior[sf] + prior[ell] + G`` in the reference's left-to-right order
    (``sliceSample.py:127,150``: ``... + prior[1] + prior[0] + curG``); with ARD the extra
    length-scale priors follow ``prior[0]``."""
    acc = head + prior[n_ell]
    for d in range(n_ell):
        acc = acc + prior[d]
    return acc + G

class SweepTrace(object):
    """Diagnostics of one restated transition (what the parity tests compare)."""

    def __init__(self):
        self.g = None
        self.cur_llk = self.curG = self.threshold = None
        self.hyp_min0 = self.hyp_max0 = None
        self.prop_hyp = []      # per trip
        self.propG = []         # per trip (inv form, :147)
        self.propG_chol = []    # per trip (chol form, :145-146)
        self.prop_llk = []
        self.proposal = []
        self.n_trips = 0


def surrogate_slice_sampling(f, x, y, hyp, scale, it, tape, max_trips=None, trace=None,
                             log_marginal='inv', r_form='literal'):
    """One surrogate-data slice-sampling transition, ``sliceSample.py:76-163``, on a tape.

    ``tape`` has ``z[N]``, ``v[P]``, ``u0``, ``U[T,P]`` (see ``reference_loader.Tape``).
    ``np.log(hyp[2])``/``exp`` round trip at ``:117`` is kept for the initial ``sn``.
    Returns ``(prop_f, prop_hyp)``."""
    f = np.asarray(f, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    hyp = np.asarray(hyp, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    P = hyp.shape[0]
    n_ell = P - 2
    my = np.mean(y)                                                          # :102
    K = cov_matrix(x, hyp)                                                   # :104-105
    g, K_S, m_theta_g, chol_R_theta, L_ks = aux_var_model(f, K, hyp[P - 1], z=tape.z, r_form=r_form)   # :107
    ita = np.linalg.solve(chol_R_theta, f - m_theta_g)                       # :108

    v = 0. + (scale - 0.) * tape.v                                           # :110
    hyp_min = np.maximum(hyp - v, 0)                                         # :111
    hyp_max = hyp_min + scale                                                # :112
    upper = 100 - my                                                         # :114
    lower = 0 - my                                                           # :115
    sn0 = np.exp(np.log(hyp[P - 1]))                                         # :117 log_sigma round trip
    cur_llk = trunc_gauss2_loglik(y - my, f, sn0, lower, upper)              # :118
    curG = log_marginal_inv_form(g, K_S, L_ks) if log_marginal == 'inv' else log_marginal_chol_form(g, L_ks)   # :122

    k, theta = prior_constants(n_ell)                                        # :124-125
    prior, _ = log_gamma(hyp, k, theta, True)                                # :126
    threshold = _sum_density(np.log(tape.u0) + cur_llk, prior, curG, n_ell)   # :127
    if it >= BURN_IN:
        threshold += prior[P - 1]                                            # :128-129
    if trace is not None:
        trace.g, trace.cur_llk, trace.curG, trace.threshold = g.copy(), cur_llk, curG, threshold
        trace.hyp_min0, trace.hyp_max0 = hyp_min.copy(), hyp_max.copy()

    trip = 0
    while True:                                                              # :131
        if trip >= tape.U.shape[0] or (max_trips is not None and trip >= max_trips):
            raise RuntimeError('tape exhausted after %d trips' % trip)
        prop_hyp = hyp_min + (hyp_max - hyp_min) * tape.U[trip]              # :132
        trip += 1
        if it < BURN_IN:
            prop_hyp[P - 1] = hyp[P - 1]                                     # :133-134
        nK = cov_matrix(x, prop_hyp)                                         # :136-137
        proposal = -np.inf
        propG = propG_c = prop_llk = np.nan
        prop_f = None
        try:
            g, K_S, m_theta_g, chol_R_theta, L_ks = aux_var_model(f, nK, prop_hyp[P - 1], g=g, r_form=r_form)   # :139
            prop_f = np.dot(chol_R_theta, ita) + m_theta_g                   # :140
            prop_llk = trunc_gauss2_loglik(y - my, prop_f, prop_hyp[P - 1], lower, upper)        # :142-143
            propG = log_marginal_inv_form(g, K_S, L_ks)                      # :147
            propG_c = log_marginal_chol_form(g, L_ks)                        # :145-146
            propPrior, _ = log_gamma(prop_hyp, k, theta, True)               # :149
            pg = propG if log_marginal == 'inv' else propG_c
            proposal = _sum_density(prop_llk, propPrior, pg, n_ell)         # :150
            if it >= BURN_IN:
                proposal += propPrior[P - 1]                                 # :151-152
        except (np.linalg.LinAlgError, ValueError):
            # jitchol gave up / non-finite matrix: the reference would abort (uncaught LinAlgError);
            # the restatement (and the CUDA path) treat it as a non-finite => rejected proposal.
            proposal = np.nan
        if trace is not None:
            trace.prop_hyp.append(prop_hyp.copy())
            trace.propG.append(propG)
            trace.propG_chol.append(propG_c)
            trace.prop_llk.append(prop_llk)
            trace.proposal.append(proposal)
            trace.n_trips = trip
        if proposal > threshold and np.isfinite(proposal):                   # :154
            return prop_f, prop_hyp                                          # :156
        for i in range(P):                                                   # :159 (0..2 in the reference)
            if prop_hyp[i] < hyp[i]:
                hyp_min[i] = prop_hyp[i]                                     # :160-161
            else:
                hyp_max[i] = prop_hyp[i]                                     # :162-163


def run_chain(x, y, hyp0, scale, iters, seed, f0=None, start_iter=0, max_trips=64, r_form='literal'):
    """The caller loop of ``framework.py:59-77`` / ``demoRegression.py:15-32`` on per-iteration tapes
    (``Tape.from_seed(seed + it)``).  Returns ``(histF[N,iters], histHyp[P,iters], trips[iters])``."""
    from .reference_loader import Tape
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = y.shape[0]
    propHyp = np.asarray(hyp0, dtype=np.float64).copy()
    propF = np.zeros_like(y) if f0 is None else np.asarray(f0, dtype=np.float64).copy()
    histF = np.zeros((n, iters))
    histHyp = np.zeros((propHyp.shape[0], iters))
    trips = np.zeros(iters, dtype=np.int64)
    for i in range(iters):
        tr = SweepTrace()
        tape = Tape.from_seed(seed + i, n, p=propHyp.shape[0], max_trips=max_trips)
        propF, propHyp = surrogate_slice_sampling(propF, x, y, propHyp, scale, start_iter + i, tape, trace=tr, r_form=r_form)
        histF[:, i] = propF
        histHyp[:, i] = propHyp
        trips[i] = tr.n_trips
    return histF, histHyp, trips


# ------------------------------------------------------------ f4: elliptical slice sampling
def elliptical_slice(f, x, y, hyp, tape, max_trips=None):
    """``elliptical_slice`` (``sliceSample.py:15-74``) on a tape (``reference_loader.EssTape``: ``nu``, ``u``, ``theta[T]``).
    Returns ``(prop_f, n_proposals)``."""
    f = np.asarray(f, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    hyp = np.asarray(hyp, dtype=np.float64)
    my = np.mean(y)                                                          # :36
    nu = np.asarray(tape.nu, dtype=np.float64)                               # :38-43 (the N(0, K) draw is on the tape)
    upper = 100 - my                                                         # :45
    lower = 0 - my                                                           # :46
    sn = np.exp(np.log(hyp[-1]))                                             # :47 log_sigma round trip
    cur_llk = trunc_gauss2_loglik(y - my, f, sn, lower, upper)               # :50
    cur_llk = cur_llk + np.log(tape.u)                                       # :51
    theta = 0. + (2. * np.pi - 0.) * tape.theta[0]                           # :54
    theta_min = theta - 2. * np.pi                                           # :55
    theta_max = theta                                                        # :56
    trip = 0
    while True:                                                              # :59
        trip += 1
        prop_f = f * np.cos(theta) + nu * np.sin(theta)                      # :60
        prop_llk = trunc_gauss2_loglik(y - my, prop_f, sn, lower, upper)     # :62
        if prop_llk > cur_llk and np.isfinite(prop_llk):                     # :64
            return prop_f, trip                                              # :66
        if theta >= 0:
            theta_max = theta                                                # :69-70
        else:
            theta_min = theta                                                # :71-72
        if trip >= tape.theta.shape[0] or (max_trips is not None and trip >= max_trips):
            raise RuntimeError('tape exhausted after %d proposals' % trip)
        theta = theta_min + (theta_max - theta_min) * tape.theta[trip]       # :74


def ess_nu_from_z(x, hyp, z):
    """``nu = jitchol(K) z`` with ``K = covK.RBF(log ll, log sf).getCovMatrix(x, 'train')`` (``:38-39``): the Cholesky
    form of the ``N(0, K)`` draw the CUDA path uses (the reference's ``:41`` goes through numpy's SVD: same law)."""
    K = cov_matrix(x, np.concatenate([np.asarray(hyp, dtype=np.float64)[:-1], [1.0]]))
    return np.dot(kcgp_shim.jitchol(K), np.asarray(z, dtype=np.float64))


# ------------------------------------------------------------ f2: predictive inference
def inf_mcmc_unit(f, x, y, xs, hyp, lower=None, upper=None):
    """``inf_mcmc`` (``sliceSample.py:234-284``) for ONE stored sample ``f[N]`` with hyper-parameters ``hyp`` and a zero
    mean function, restated line by line.  Returns ``(ym, ys_lw, ys_up, Fs2, Fmu)`` (``Fmu`` before the likelihood)."""
    x = np.asarray(x, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    hyp = np.asarray(hyp, dtype=np.float64)
    n_ell = hyp.shape[0] - 2
    my = np.mean(y)                                                          # :249
    n = x.shape[0]
    ns = xs.shape[0]
    f = np.asarray(f, dtype=np.float64).reshape(n, 1)
    if n_ell == 1:
        cov = kcgp_shim.RBF(np.log(hyp[0]), np.log(hyp[1]))
    else:
        cov = kcgp_shim.RBFard(log_ell_list=list(np.log(hyp[:n_ell])), log_sigma=np.log(hyp[n_ell]))
    lik = kcgp_shim.TruncatedGauss2(upper=(100 - my) if upper is None else upper, lower=(0 - my) if lower is None else lower,
                                    log_sigma=np.log(hyp[-1]))
    m = np.zeros((n, 1))                                                     # :254 (zero mean)
    K = cov.getCovMatrix(x=x, mode='train')                                  # :255
    sn2 = lik.sn ** 2.                                                       # :256
    L = kcgp_shim.jitchol(K / sn2 + np.eye(n)).T                             # :257
    alpha = kcgp_shim.solve_chol(L, f - m) / sn2                             # :258
    sW = np.ones((n, 1)) / np.sqrt(sn2)                                      # :259
    kss = cov.getCovMatrix(z=xs, mode='self_test')                           # :262
    Ks = cov.getCovMatrix(x=x, z=xs, mode='cross')                           # :263
    Fmu = np.zeros((ns, 1)) + np.dot(Ks.T, alpha)                            # :265-266
    V = np.linalg.solve(L.T, np.tile(sW, (1, ns)) * Ks)                      # :269
    fs2 = kss - np.array([(V * V).sum(axis=0)]).T                            # :270
    Fs2 = np.maximum(fs2, 0)                                                 # :275
    Fmu = np.mean(Fmu, axis=1, keepdims=True)                                # :277
    Ymu, Lower, Upper = lik.evaluate(mu=Fmu, s2=Fs2)                         # :279
    return (np.reshape(np.mean(Ymu, axis=1), (ns, 1)) + my, np.reshape(np.mean(Lower, axis=1), (ns, 1)) + my,
            np.reshape(np.mean(Upper, axis=1), (ns, 1)) + my, Fs2, Fmu)
