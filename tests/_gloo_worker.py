"""Worker of tests/test_distributed_gloo.py: world_size-2 run of the chain ensemble's host logic on CPU (gloo).

The device sweeper is replaced by a test double that advances chains with the CPU oracle, so what is exercised is
exactly the multi-rank plumbing of chains.py: contiguous sharding with uneven shards, global-chain-id keyed
randomness, and the single all-gather of [theta, loglik, ntrips] per sweep."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gpmc_b200 as gp                                   # noqa: E402
from oracle import sds_oracle as so                      # noqa: E402
from oracle.reference_loader import Tape                 # noqa: E402


class OracleSweeper(object):
    """Test double for chains._DeviceSweeper: same interface, CPU oracle inside, RNG keyed by global chain id."""

    def __init__(self, x, y, scale, seed, lo):
        self.x, self.y, self.scale, self.seed, self.lo = x, y, scale, seed, lo

    def load(self, F0, Hyp0):
        self.F, self.H = F0.copy(), Hyp0.copy()

    def sweep(self, it):
        nt = np.zeros(self.H.shape[0], dtype=np.int64)
        ll = np.zeros(self.H.shape[0])
        for c in range(self.H.shape[0]):
            tr = so.SweepTrace()
            tape = Tape.from_seed(self.seed + 100000 * (self.lo + c) + it, self.F.shape[1])
            self.F[c], self.H[c] = so.surrogate_slice_sampling(self.F[c], self.x, self.y, self.H[c], self.scale, it, tape, trace=tr)
            nt[c], ll[c] = tr.n_trips, tr.propG[-1]
        return self.H, ll, nt

    def state(self):
        return self.F.copy(), self.H.copy()


class OracleRunSweeper(OracleSweeper):
    """The same double with the `run` method of the device sweeper (many iterations per call, history as tensors), so
    that ChainEnsemble.run takes its block path: one all-gather of [n_local, k, P + 2] per block of iterations."""

    def run(self, it0, n_iters, thin_f=0, thin_phase=0):
        hh = np.zeros((self.H.shape[0], n_iters, self.H.shape[1]))
        ll = np.zeros((self.H.shape[0], n_iters))
        nt = np.zeros((self.H.shape[0], n_iters), dtype=np.int64)
        for k in range(n_iters):
            H, l, t = self.sweep(it0 + k)
            hh[:, k, :], ll[:, k], nt[:, k] = H, l, t
        return torch.tensor(hh), torch.tensor(ll), torch.tensor(nt), None, 0


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    n, B, iters = 24, 5, 3
    x, y = gp.synthetic.ih45_series(n)
    scale = np.array(gp.synthetic.SCALE)
    F0, H0 = gp.synthetic.chain_states(B, n)
    lo, hi = gp.chains.shard_bounds(B, world, rank)
    ens = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=7, sweeper=OracleSweeper(x, y, scale, 7, lo))
    assert (ens.lo, ens.hi) == (lo, hi) and ens.n_chains == B and ens.world == world
    hist, ll, trips = ens.run(iters)
    # every rank holds the full gathered history; compare with one process doing all chains
    solo = OracleSweeper(x, y, scale, 7, 0)
    solo.load(F0, H0)
    for i in range(iters):
        H, l, t = solo.sweep(i)
        assert np.array_equal(hist[:, :, i], H), (rank, i)
        assert np.array_equal(ll[:, i], l) and np.array_equal(trips[:, i], t)
    # sharded_input path: ranks pass only their shard
    ens2 = gp.chains.ChainEnsemble(x, y, F0[lo:hi], H0[lo:hi], scale, seed=7, sharded_input=True,
                                   sweeper=OracleSweeper(x, y, scale, 7, lo))
    assert ens2.n_chains == B and (ens2.lo, ens2.hi) == (lo, hi)
    h2, _, _ = ens2.sweep(0)
    assert np.array_equal(h2, hist[:, :, 0])
    # block path of ChainEnsemble.run (what the device sweeper uses): history gathered per block of iterations
    ens3 = gp.chains.ChainEnsemble(x, y, F0, H0, scale, seed=7, sweeper=OracleRunSweeper(x, y, scale, 7, lo))
    h3, l3, t3 = ens3.run(iters, gather_every=2)
    assert np.array_equal(h3, hist) and np.array_equal(l3, ll) and np.array_equal(t3, trips)
    dist.barrier()
    if rank == 0:
        print('GLOO_OK shards=%s' % [gp.chains.shard_bounds(B, world, r) for r in range(world)])
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
