"""One-off: larger statistical comparison of Philox device chains vs numpy oracle chains (not part of the test suite)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
from oracle import sds_oracle as so

n, B, iters, burn = 48, 240, 60, 20
SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # shifts every chain's seed: an independent replication
if len(sys.argv) > 2:
    B = int(sys.argv[2])
MODE = sys.argv[3] if len(sys.argv) > 3 else 'sds'


def ess_check():
    """Philox-driven device elliptical-slice chains (gpmc_ess_sweep) vs numpy-driven oracle chains: fixed
    hyper-parameters, `iters` updates of f per chain, statistics of the last state (mean of f, truncated-Gaussian
    log-likelihood) and the proposal counts."""
    import torch
    from oracle.reference_loader import EssTape
    x, y = gp.synthetic.ih45_series(n)
    hyp = np.array([4.0, 5.0, 1.8])
    F0 = np.tile(0.7 * (y - y.mean()), (B, 1))
    H = np.tile(hyp, (B, 1))
    F, Hd = torch.tensor(F0).cuda(), torch.tensor(H).cuda()
    tr_dev = 0
    for it in range(iters):
        trips, status, info = gp.ops.ess_sweep(x, y, F, Hd, it=it, seed=777 + 7919 * SEED)
        tr_dev += float(trips.double().mean().item())
    Fd = F.cpu().numpy()
    my = y.mean()
    Fo = np.zeros_like(Fd); tr_ora = 0.0
    for c in range(B):
        rs = np.random.RandomState(500000 + c + 7919 * SEED)
        f = F0[c].copy()
        for it in range(iters):
            tape = EssTape(so.ess_nu_from_z(x, hyp, rs.standard_normal(n)), rs.random_sample(), rs.random_sample(256))
            f, t = so.elliptical_slice(f, x, y, hyp, tape)
            tr_ora += t / B
        Fo[c] = f
    stats = {'mean_f': (Fd.mean(axis=1), Fo.mean(axis=1)),
             'f_mid': (Fd[:, n // 2], Fo[:, n // 2]),
             'loglik': (np.array([so.trunc_gauss2_loglik(y - my, f, hyp[-1], 0 - my, 100 - my) for f in Fd]),
                        np.array([so.trunc_gauss2_loglik(y - my, f, hyp[-1], 0 - my, 100 - my) for f in Fo]))}
    for k, (a, b) in stats.items():
        se = np.sqrt(a.var(ddof=1) / B + b.var(ddof=1) / B)
        print(json.dumps({'mode': 'ess', 'stat': k, 'device': float(a.mean()), 'oracle': float(b.mean()), 'z': float((a.mean() - b.mean()) / se),
                          'trips_dev': tr_dev / iters, 'trips_ora': tr_ora / iters}))


if MODE == 'ess':
    iters = 30
    ess_check()
    sys.exit(0)
x, y = gp.synthetic.ih45_series(n)
scale = np.array(gp.synthetic.SCALE)
F0, H0 = gp.synthetic.chain_states(B, n)
for start in (0, 500):
    ens = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=4242 + start + 7919 * SEED)
    hist, ll, trips = ens.run(iters, start_iter=start)
    dev = np.log(hist[:, :, burn:]).mean(axis=2)
    ora = np.zeros((B, 3)); otr = []
    for c in range(B):
        _, hh, tt = so.run_chain(x, y, H0[c], scale, iters, seed=900000 + 1000 * c + start + 7919 * SEED, start_iter=start)
        ora[c] = np.log(hh[:, burn:]).mean(axis=1); otr.append(tt.mean())
    for d in range(3 if start else 2):
        se = np.sqrt(dev[:, d].var(ddof=1) / B + ora[:, d].var(ddof=1) / B)
        print(json.dumps({'start': start, 'dim': d, 'device': float(dev[:, d].mean()), 'oracle': float(ora[:, d].mean()),
                          'z': float((dev[:, d].mean() - ora[:, d].mean()) / se), 'trips_dev': float(trips.mean()), 'trips_ora': float(np.mean(otr))}))
