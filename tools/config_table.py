"""Time the log-lik unit (and SDS sweeps) on every BASELINE.json config shape that fits one GPU; prints JSON lines."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch


def time_loglik(n, B, ard=0, reps=3):
    if ard:
        x, _ = gp.synthetic.ard_inputs(n, ard)
        n_ell = ard
    else:
        x = np.arange(n, dtype=np.float64).reshape(n, 1)
        n_ell = 1
    G, H = gp.synthetic.loglik_batch(B, n, n_ell=n_ell)
    xd, Gd, Hd = torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda()
    for _ in range(2):
        ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    gp.ops.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = gp.ops.profile_read()
    gp.ops.profile(False)
    chol_ms = (prof['gemm_update'][0] + prof['potf2'][0] + prof['panel_trsm'][0]) / reps
    return {'config': 'loglik N=%d B=%d%s' % (n, B, ' ARD D=%d' % ard if ard else ''), 'ms_per_pass': ms,
            'evals_per_s': B / (ms * 1e-3), 'cholesky_tflops': B * n ** 3 / 3.0 / (chol_ms * 1e-3) / 1e12,
            'kernel_ms': {k: round(v[0] / reps, 3) for k, v in prof.items() if v[0] > 0}, 'failed': int((info != 0).sum().item())}


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'c2':
        rows = [time_loglik(2048, 64, reps=10), time_loglik(200, 1, reps=20)]
    elif len(sys.argv) > 1 and sys.argv[1] == 'quick':
        rows = [time_loglik(512, 4096, ard=4), time_loglik(16384, 1, reps=2)]
    else:
        rows = [time_loglik(200, 1), time_loglik(2048, 64), time_loglik(512, 4096, ard=4), time_loglik(16384, 1, reps=2),
                time_loglik(4096, 256)]
    for r in rows:
        print(json.dumps(r))
