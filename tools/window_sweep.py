"""Time the single large matrix (BASELINE config 4) for several window widths / look-ahead settings."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch


def run(n, B, reps=3):
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(B, n)
    xd, Gd, Hd = torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda()
    for _ in range(2):
        ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, float(ll[0].item())


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for la in (1, 2):
        for win in (256, 512, 1024, 2048):
            gp.ops.set_tuning(2, la)
            gp.ops.set_tuning(3, win)
            ms, ll = run(n, B)
            print(json.dumps({'n': n, 'B': B, 'lookahead': la == 2, 'window': win, 'ms': round(ms, 3), 'll0': ll,
                              'tflops_whole_pass': round(B * n ** 3 / 3.0 / ms / 1e9, 2)}))
