"""Small-size pass through every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck /
synccheck) on a B200:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py --small

(Where compute-sanitizer is not available the script still serves as a one-minute pass over every entry point;
tests/test_gpu_next_rows.py runs it.)  Sizes hit the ragged paths: a single ragged block (37), one full block + one row (129), the fused panel path (<= 640),
separate panel kernels with the look-ahead streams (700, 1000 x 2), many matrices (left-looking, no look-ahead).
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gp = importlib.import_module('gaussianprocess-mcmc_b200')


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--small', action='store_true', help='sizes for the slow tools (racecheck)')
    a = ap.parse_args(argv)
    sizes = [(37, 3), (129, 3), (200, 2), (300, 5)] if a.small else [(37, 3), (129, 3), (200, 2), (300, 5), (700, 3), (1000, 2), (257, 300)]
    for n, B in sizes:
        x = np.arange(n, dtype=np.float64).reshape(n, 1)
        G, H = gp.synthetic.loglik_batch(B, n)
        ll, info = gp.ops.loglik_host(x, G, H)
        assert np.all(np.isfinite(ll)) and np.all(info == 0), (n, B, ll, info)
        print('loglik', n, B, 'ok', flush=True)
    # every panel kernel variant on one ragged size
    n = 300
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(3, n)
    ref = gp.ops.loglik_host(x, G, H)[0]
    for key, val in ((1, 1), (1, 2), (1, 3), (9, 1), (4, 1)):
        gp.ops.set_tuning(key, val)
        ll = gp.ops.loglik_host(x, G, H)[0]
        gp.ops.set_tuning(key, 0)
        assert np.allclose(ll, ref, rtol=1e-11), (key, val)
        print('variant', key, val, 'ok', flush=True)
    # ARD assembly + potrf with the ladder
    n = 160
    xa, _ = gp.synthetic.ard_inputs(n, 4)
    G, H = gp.synthetic.loglik_batch(4, n, n_ell=4)
    ll, info = gp.ops.loglik_host(xa, G, H)
    assert np.all(np.isfinite(ll))
    print('ard ok', flush=True)
    # SDS: resident loop, wave loop, run mode, literal R
    n, B = (48, 5) if a.small else (96, 7)
    x, y = gp.synthetic.ih45_series(n)
    F0, H0 = gp.synthetic.chain_states(B, n)
    scale = np.array([10., 10., 5.])
    outs = {}
    for mode in (0, 1):
        gp.ops.set_tuning(6, mode)
        F, H = torch.tensor(F0).cuda(), torch.tensor(H0).cuda()
        trips, ll, status = gp.ops.sds_sweep(x, y, F, H, scale, 3, seed=5)
        outs[mode] = (F.cpu().numpy(), H.cpu().numpy())
    gp.ops.set_tuning(6, 0)
    assert np.array_equal(outs[0][1], outs[1][1])
    print('sds sweep ok', flush=True)
    F, H = torch.tensor(F0).cuda(), torch.tensor(H0).cuda()
    hist = gp.ops.sds_run(x, y, F, H, scale, 0, 3, seed=5, keep_f_every=1)
    assert np.all(np.isfinite(hist[0].cpu().numpy()))
    print('sds run ok', flush=True)
    gp.ops.set_tuning(8, 1)
    F, H = torch.tensor(F0).cuda(), torch.tensor(H0).cuda()
    gp.ops.sds_sweep(x, y, F, H, scale, 3, seed=5)
    gp.ops.set_tuning(8, 0)
    assert np.all(np.isfinite(H.cpu().numpy()))
    print('sds literal ok', flush=True)
    # predictive path and elliptical slice
    xs = np.linspace(0.5, n + 3.5, 11).reshape(-1, 1)
    fm = torch.tensor(F0 - F0.mean(axis=1, keepdims=True)).cuda()
    fmu, fs2, info = gp.ops.predict_batched(x.reshape(-1, 1), xs, fm, torch.tensor(H0).cuda())
    assert np.all(np.isfinite(fmu.cpu().numpy()))
    print('predict ok', flush=True)
    F, H = torch.tensor(F0).cuda(), torch.tensor(H0).cuda()
    trips, status, info = gp.ops.ess_sweep(x, y, F, H, it=1, seed=9)
    assert np.all(np.isfinite(F.cpu().numpy()))
    print('ess ok', flush=True)
    # single-matrix auxiliary model
    K = gp.ops.cov_assemble(x.reshape(-1, 1), H0[:1])[0, :n, :n].contiguous()
    S = torch.full((n,), 1.44, dtype=torch.float64, device='cuda')
    g = torch.tensor(y - y.mean()).cuda()
    L, m, C, info = gp.ops.aux_var_model_device(K, S, g)
    assert int(info.abs().sum().item()) == 0
    print('aux ok', flush=True)
    torch.cuda.synchronize()
    print('ALL OK')


if __name__ == '__main__':
    main()
