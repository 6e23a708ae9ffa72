"""Experiment: forced window sizes of the blocked Cholesky at BASELINE config 2 (N=2048 x 64) and config 4."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
from tools.config_table import time_loglik
for n, B, reps in ((2048, 64, 10), (16384, 1, 2)):
    for win in (0, 256, 512, 1024, 2048, 100000):
        for la in (0, 1, 2):
            if n == 16384 and (win in (256, 100000) or la == 1):
                continue
            gp.ops.set_tuning(3, win)
            gp.ops.set_tuning(2, la)
            r = time_loglik(n, B, reps=reps)
            print(json.dumps({'n': n, 'B': B, 'window': win, 'lookahead': la, 'ms': round(r['ms_per_pass'], 3), 'evals_per_s': round(r['evals_per_s'], 1)}))
gp.ops.set_tuning(3, 0); gp.ops.set_tuning(2, 0)
