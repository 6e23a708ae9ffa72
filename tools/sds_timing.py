"""Time device-resident SDS sweeps (BASELINE configs 2/3 style) and print the per-kernel-class breakdown."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nobs', type=int, default=2048)
    ap.add_argument('--chains', type=int, default=64)
    ap.add_argument('--sweeps', type=int, default=3)
    ap.add_argument('--ard', type=int, default=0, help='input dimension D for the ARD kernel (0: 1-D iso)')
    ap.add_argument('--start-iter', type=int, default=0)
    args = ap.parse_args()
    import torch
    n, B = args.nobs, args.chains
    if args.ard:
        x, y = gp.synthetic.ard_inputs(n, args.ard)
        n_ell = args.ard
    else:
        x, y = gp.synthetic.ih45_series(n)
        n_ell = 1
    F0, H0 = gp.synthetic.chain_states(B, n, n_ell=n_ell)
    scale = np.array([gp.synthetic.SCALE[0]] * n_ell + list(gp.synthetic.SCALE[1:]))
    F = torch.tensor(F0).cuda()
    H = torch.tensor(H0).cuda()
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    my = float(np.mean(y))
    gp.ops.sds_sweep(xd, yd, F, H, scale, args.start_iter, my=my, seed=1)          # warm-up
    torch.cuda.synchronize()
    gp.ops.profile(True)
    t0 = time.perf_counter()
    trips = []
    for i in range(args.sweeps):
        nt, ll, st = gp.ops.sds_sweep(xd, yd, F, H, scale, args.start_iter + 1 + i, my=my, seed=1)
        trips.append(nt.cpu().numpy())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = gp.ops.profile_read()
    gp.ops.profile(False)
    trips = np.stack(trips)
    evals = int((trips + 1).sum())                 # aux-model evaluations: one at theta + one per trip
    flops = evals * (4.0 / 3.0) * n ** 3
    print(json.dumps({
        'n': n, 'chains': B, 'sweeps': args.sweeps, 's_per_sweep': dt / args.sweeps,
        'chain_sweeps_per_s': B * args.sweeps / dt, 'mean_trips': float(trips.mean()), 'max_trips': int(trips.max()),
        'aux_evals': evals, 'model_tflops(4/3 N^3 per eval)': flops / dt / 1e12,
        'kernel_ms': {k: round(v[0], 2) for k, v in prof.items()}, 'launches': int(sum(v[1] for v in prof.values())),
        'status_nonzero': int((st != 0).sum().item()), 'hyp_mean': H.mean(dim=0).cpu().numpy().round(3).tolist()}))


if __name__ == '__main__':
    main()
