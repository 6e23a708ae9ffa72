"""Chain-parallel SDS ensemble under torchrun (NCCL): every rank sweeps its shard, one all-gather per sweep;
rank 0 re-runs all chains alone and checks that sharding changed nothing (RNG is keyed by global chain id)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n, B, iters = int(os.environ.get('ENS_N', '512')), int(os.environ.get('ENS_B', '64')), 4
    x, y = gp.synthetic.ih45_series(n)
    F0, H0 = gp.synthetic.chain_states(B, n)
    scale = np.array(gp.synthetic.SCALE)
    ens = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=11)
    ens.sweep(0)                                            # warm-up (allocations, NCCL)
    ens = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=11)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    hist, ll, trips = ens.run(iters)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    ok = None
    if rank == 0:
        solo = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=11, distributed=False)
        h2, l2, t2 = solo.run(iters)
        ok = bool(np.array_equal(h2, hist) and np.array_equal(t2, trips) and np.array_equal(l2, ll))
        print(json.dumps({'world': world, 'n': n, 'chains': B, 'iters': iters, 'shard_independent': ok,
                          's_per_sweep': dt / iters, 'mean_trips': float(trips.mean())}))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
