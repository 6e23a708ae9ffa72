"""Latency / throughput of the panel kernels alone: one N=128 factorisation is exactly one panel-factor launch, and an
N=256 one adds one panel solve.  Prints JSON lines for the panel factor kernels (modes: 0 register-resident, 2 shared-memory
lite, 1 full inverse) at B=1 (latency) and B=4096 (throughput), and for the panel solve with 1/2/4 row blocks per CTA."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch


def time_potrf(n, B, reps=20):
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    H = np.tile(np.array([[5.0, 4.0, 2.5]]), (B, 1))
    A0 = gp.ops.cov_assemble(x, H, add_S=True, ld=(n + 15) // 16 * 16)
    A = A0.clone()
    for _ in range(3):
        A.copy_(A0)
        gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE, zero_upper=False)
    torch.cuda.synchronize()
    gp.ops.profile(True)
    for _ in range(reps):
        A.copy_(A0)
        info = gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE, zero_upper=False)
    torch.cuda.synchronize()
    prof = gp.ops.profile_read()
    gp.ops.profile(False)
    assert int((info != 0).sum().item()) == 0
    return {k: (v[0] / reps * 1e3, v[1] // reps) for k, v in prof.items() if v[1] > 0}


if __name__ == '__main__':
    for mode in (0, 2, 1):
        gp.ops.set_tuning(1, mode)
        for n, B in ((128, 1), (128, 296), (128, 4096), (512, 4096)):
            r = time_potrf(n, B)
            print(json.dumps({'potf2_mode': mode, 'n': n, 'B': B, 'us_per_call': {k: round(v[0], 1) for k, v in r.items()},
                              'launches': {k: v[1] for k, v in r.items()}}))
    gp.ops.set_tuning(1, 0)
    for per in (1, 2, 4, 0):
        gp.ops.set_tuning(5, per)
        for n, B in ((4096, 64), (512, 4096)):
            r = time_potrf(n, B, reps=3)
            print(json.dumps({'trsm_blocks_per_cta': per, 'n': n, 'B': B, 'us_per_call': {k: round(v[0], 1) for k, v in r.items()}}))
    gp.ops.set_tuning(5, 0)
