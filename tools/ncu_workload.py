"""Small workload for `ncu --set full` captures: ONE log-lik pass (assemble -> blocked Cholesky with the border row ->
reductions) for a few N=4096 matrices, bracketed by cudaProfilerStart/Stop so that `--profile-from-start off` captures
exactly the launches of one pass (warm-up passes stay outside).

    python tools/ncu_workload.py [--nobs 4096] [--chains 16] [--sds]
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_full python tools/ncu_workload.py
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nobs', type=int, default=4096)
    ap.add_argument('--chains', type=int, default=16)
    ap.add_argument('--ard', type=int, default=0)
    args = ap.parse_args()
    n, B = args.nobs, args.chains
    if args.ard:
        x, _ = gp.synthetic.ard_inputs(n, args.ard)
        n_ell = args.ard
    else:
        x = np.arange(n, dtype=np.float64).reshape(n, 1)
        n_ell = 1
    G, H = gp.synthetic.loglik_batch(B, n, n_ell=n_ell)
    xd, Gd, Hd = torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda()
    for _ in range(2):
        ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ll, info = gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    assert int((info != 0).sum().item()) == 0
    print('ncu workload ok: N=%d B=%d loglik[0]=%.6f' % (n, B, float(ll[0].item())))


if __name__ == '__main__':
    main()
