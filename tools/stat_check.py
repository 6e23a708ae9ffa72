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
