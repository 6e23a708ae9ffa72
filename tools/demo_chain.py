"""BASELINE config 1 (demoRegression.py: N=200, one chain): the drop-in sampler on the GPU and the CPU oracle consume the
same global numpy stream; report how long the two trajectories make identical decisions, and the time per iteration."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
from oracle import sds_oracle as so
from oracle.reference_loader import Tape

n = int(os.environ.get('DEMO_N', '200')); iters = int(os.environ.get('DEMO_ITERS', '300')); start = int(os.environ.get('DEMO_START', '350'))
x, y = gp.synthetic.ih45_series(n)
scale = np.array([10., 10., 5.])
sds = gp.kcMCMC.sliceSample
# GPU: the reference's caller loop (demoRegression.py:23-30) on the drop-in function
np.random.seed(124)
f, h = np.zeros(n), np.array([0.35, 2.0, 0.2])          # demoRegression.py:15
sds.surrogate_slice_sampling(f, x, y, h, scale, iter=start)   # warm-up (does not touch the comparison: re-seeded below)
np.random.seed(124)
t0 = time.perf_counter(); gh = []
for i in range(iters):
    f, h = sds.surrogate_slice_sampling(f, x, y, h, scale, iter=start + i)
    gh.append(h.copy())
t_gpu = (time.perf_counter() - t0) / iters
gh = np.array(gh)
# CPU oracle on the same stream
rs = np.random.RandomState(124)
f, h = np.zeros(n), np.array([0.35, 2.0, 0.2]); oh = []; trips = []
t0 = time.perf_counter()
for i in range(iters):
    z, v, u0 = rs.standard_normal(n), rs.random_sample(3), rs.random_sample()
    st = rs.get_state(); U = rs.random_sample((256, 3))
    tr = so.SweepTrace()
    f, h = so.surrogate_slice_sampling(f, x, y, h, scale, start + i, Tape(z, v, u0, U), trace=tr)
    rs.set_state(st); rs.random_sample((tr.n_trips, 3))
    oh.append(h.copy()); trips.append(tr.n_trips)
t_cpu = (time.perf_counter() - t0) / iters
oh = np.array(oh)
rel = np.abs(gh - oh).max(axis=1) / np.abs(oh).max(axis=1)
same = int(np.argmax(rel > 1e-6)) if np.any(rel > 1e-6) else iters
print(json.dumps({'n': n, 'iters': iters, 'start_iter': start, 'identical_decisions_for_first_iterations': same,
                  'gpu_s_per_iter': t_gpu, 'cpu_oracle_s_per_iter': t_cpu, 'mean_trips': float(np.mean(trips)),
                  'gpu_mean_log_hyp_last_half': np.log(gh[iters // 2:]).mean(axis=0).round(3).tolist(),
                  'cpu_mean_log_hyp_last_half': np.log(oh[iters // 2:]).mean(axis=0).round(3).tolist()}))
