"""CPU side of BASELINE.md section 4: the oracle port timed on the box's host cores at several N.

  unit_inv   : assemble + jitchol + quadratic form through a dense inverse, as sliceSample.py:136-147 writes it
  unit_trsv  : the restated unit (cdist+exp, dpotrf, one triangular solve, log-diag)
  proposal   : one full reference proposal (getCovMatrix + aux_var_model(g=g) + propG, sliceSample.py:136-147 incl. :197-205)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sds_oracle as so


def med(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    rs = np.random.RandomState(0)
    out = {'cores': os.cpu_count(), 'rows': []}
    for n, reps in ((200, 5), (512, 5), (2048, 3), (4096, 3)):
        x = np.arange(n, dtype=np.float64).reshape(n, 1)
        hyp = np.array([1., 10., 1.2])
        g = 1.2 * rs.standard_normal(n)
        f = np.zeros(n)
        row = {'n': n,
               'unit_inv_s': med(lambda: so.loglik_unit(x, g, hyp, form='inv'), reps),
               'unit_trsv_s': med(lambda: so.loglik_unit(x, g, hyp, form='trsv'), reps)}
        if n <= 2048:
            def proposal():
                K = so.cov_matrix(x, hyp)
                gg, KS, m, C, L = so.aux_var_model(f, K, hyp[2], g=g)
                so.log_marginal_inv_form(gg, KS, L)
            row['proposal_s'] = med(proposal, 2 if n == 2048 else 3)
        out['rows'].append(row)
    r4096 = out['rows'][-1]
    out['extrapolated_n16384_unit_trsv_s'] = r4096['unit_trsv_s'] * 64.0       # N^3 scaling from N=4096 (stated)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
