"""Launch timeline of ONE log-lik pass (library kernels, CUDA events on their own streams): where the wall time of the
look-ahead configurations goes.  For each kernel class: launches, summed duration, union of its intervals; then the time
no update GEMM was running, and the panel chain (potf2 / panel solve durations by position in the factorisation).

    python tools/timeline.py --n 16384 --batch 1
    python tools/timeline.py --n 2048 --batch 64 [--tune 7=...]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gp = importlib.import_module('gaussianprocess-mcmc_b200')


def union(iv):
    if len(iv) == 0:
        return 0.0, []
    iv = iv[np.argsort(iv[:, 0])]
    out = [list(iv[0])]
    for a, b in iv[1:]:
        if a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return sum(b - a for a, b in out), out


def sds_timeline(a):
    import time
    n, B = a.n, a.batch
    x, y = gp.synthetic.ih45_series(n)
    F0, H0 = gp.synthetic.chain_states(B, n)
    scale = np.array([10., 10., 5.])
    F, H = torch.tensor(F0).cuda(), torch.tensor(H0).cuda()
    for it in range(3):
        gp.ops.sds_sweep(x, y, F, H, scale, it, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trips = 0
    for it in range(a.sds):
        trips += int(gp.ops.sds_sweep(x, y, F, H, scale, 3 + it, seed=1)[0].sum().item())
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.sds
    gp.ops.profile(True)
    tr = int(gp.ops.sds_sweep(x, y, F, H, scale, 3 + a.sds, seed=1)[0].sum().item())
    torch.cuda.synchronize()
    tl = gp.ops.profile_timeline('vector_control')
    gp.ops.profile(False)
    iv = np.concatenate([v for v in tl.values() if len(v)])
    end = iv[:, 1].max()
    u, _ = union(iv)
    print('N=%d x %d chains: %.3f ms per transition (wall, %.1f proposals per transition)' % (n, B, 1e3 * wall, trips / a.sds / B))
    print('profiled transition: %d proposals, %d launches, first launch -> last end %.3f ms, kernels running %.3f ms (%.0f %%), median kernel %.1f us'
          % (tr, len(iv), end, u, 100 * u / end, 1e3 * np.median(iv[:, 1] - iv[:, 0])))
    for k, v in tl.items():
        if len(v):
            d = v[:, 1] - v[:, 0]
            print('  %-16s launches %4d  sum %8.3f ms  median %6.1f us' % (k, len(v), d.sum(), 1e3 * np.median(d)))
    if a.dump:
        rows = sorted((v0, v1, k) for k, v in tl.items() for v0, v1 in v)
        for v0, v1, k in rows:
            print('%9.1f %9.1f  %7.1f us  %s' % (1e3 * v0, 1e3 * v1, 1e3 * (v1 - v0), k))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=16384)
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--tune', action='append', default=[], help='key=value for gpmc_set_tuning')
    ap.add_argument('--json', default=None)
    ap.add_argument('--dump', action='store_true', help='every launch, sorted by start')
    ap.add_argument('--sds', type=int, default=0, help='time this many SDS transitions of --batch chains instead of a log-lik pass')
    a = ap.parse_args()
    for kv in a.tune:
        k, v = kv.split('=')
        gp.ops.set_tuning(int(k), int(v))
    n, B = a.n, a.batch
    if a.sds:
        return sds_timeline(a)
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(B, n, n_ell=1)
    xd, Gd, Hd = torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda()
    for _ in range(3):
        gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gp.ops.loglik_batched(xd, Gd, Hd)
    e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / 3
    gp.ops.profile(True)
    gp.ops.loglik_batched(xd, Gd, Hd)
    torch.cuda.synchronize()
    tl = gp.ops.profile_timeline('assemble')
    gp.ops.profile(False)
    end = max(v[:, 1].max() for v in tl.values() if len(v))
    rec = {'n': n, 'batch': B, 'pass_ms_plain': plain_ms, 'pass_ms_profiled': float(end), 'classes': {}}
    print('pass: %.3f ms plain, %.3f ms with the event pairs' % (plain_ms, end))
    for k, v in tl.items():
        if len(v) == 0:
            continue
        u, _ = union(v)
        d = v[:, 1] - v[:, 0]
        rec['classes'][k] = {'launches': len(v), 'sum_ms': float(d.sum()), 'union_ms': float(u)}
        print('%-16s launches %4d  sum %8.3f ms  union %8.3f ms  median %7.1f us  max %7.1f us' % (k, len(v), d.sum(), u, 1e3 * np.median(d), 1e3 * d.max()))
    g = tl['gemm_update']
    ug, merged = union(g)
    print('no update GEMM running: %.3f ms of %.3f' % (end - ug, end))
    rec['no_gemm_ms'] = float(end - ug)
    # panel chain by position (deciles of the launch order)
    for k in ('potf2', 'panel_trsm'):
        v = tl[k]
        if len(v) < 10:
            continue
        d = 1e3 * (v[:, 1] - v[:, 0])
        parts = np.array_split(d, 8)
        print('%-12s mean us by eighth of the factorisation: %s' % (k, ' '.join('%6.1f' % p.mean() for p in parts)))
        rec['classes'][k]['mean_us_by_eighth'] = [float(p.mean()) for p in parts]
    # the chain: from one potf2 start to the next (column period), by eighth
    p = tl['potf2']
    if len(p) > 10:
        per = 1e3 * np.diff(p[:, 0])
        parts = np.array_split(per, 8)
        print('column period   mean us by eighth: %s' % ' '.join('%6.1f' % q.mean() for q in parts))
        rec['column_period_us_by_eighth'] = [float(q.mean()) for q in parts]
        # gap between the end of the panel solve of column j and the start of potf2 of column j+1
        t = tl['panel_trsm']
        m = min(len(t), len(p) - 1)
        gap = 1e3 * (p[1:m + 1, 0] - t[:m, 1])
        parts = np.array_split(gap, 8)
        print('solve(j) end -> potf2(j+1) start, mean us by eighth: %s' % ' '.join('%6.1f' % q.mean() for q in parts))
        rec['solve_to_next_potf2_us_by_eighth'] = [float(q.mean()) for q in parts]
    if a.dump:
        rows = sorted((v0, v1, k) for k, v in tl.items() for v0, v1 in v)
        for v0, v1, k in rows:
            print('%9.1f %9.1f  %7.1f us  %s' % (1e3 * v0, 1e3 * v1, 1e3 * (v1 - v0), k))
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(rec, f)


if __name__ == '__main__':
    main()
