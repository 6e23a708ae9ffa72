"""Caller loop and on-disk output of the reference's experiment driver, for one chain or an ensemble.

Mirrors ``framework.py:12-122`` of the reference: ``Framework(data)`` takes ``data[:, 0] = y`` (condition score) and
``data[:, 1:] = x``; ``runSimulMCMC(iters)`` is the loop of ``:59-77`` (``propHyp = [1, 10, 1.2]``, ``propF = 0``,
``scale = [10, 10, 5]``) and returns ``(histF[N, iters], histHyp[3, iters])``; ``output`` writes ``hypGap<g>.csv`` /
``fGap<g>.csv`` with the reference's headers (``:93-110``; the column it calls ``sf2`` holds ``sf``).  With ``chains > 1``
the same loop runs as a device-resident ensemble (``chains.ChainEnsemble``) and returns one more leading axis."""
import csv
import os

import numpy as np

from . import chains
from .kcMCMC import sdsK

HYP0 = (1., 10., 1.2)            # framework.py:63
SCALE = (10., 10., 5.)           # framework.py:69


class Framework(object):
    def __init__(self, data, window=None, gap=None):
        if data is None:
            raise Exception('No data is given.')                    # framework.py:25-26
        data = np.asarray(data, dtype=np.float64)
        self.x = data[:, 1:]
        self.y = np.reshape(data[:, 0], (np.shape(data)[0], 1))

    def runSimulMCMC(self, iters, chains_=1, seed=0, verbose=False):
        y = self.y.reshape((self.y.shape[0],))
        if chains_ == 1:
            propHyp = np.asarray(HYP0)
            propF = np.zeros_like(y)
            histF = np.zeros((y.shape[0], iters))
            histHyp = np.zeros((propHyp.shape[0], iters))
            for i in range(iters):
                propF, propHyp = sdsK.surrogate_slice_sampling(propF, self.x, y, propHyp, scale=np.asarray(SCALE), iter=i)
                if verbose:
                    print('Iteration: %r: ll=%.3f, sf=%.3f, sn=%.3f' % (i + 1, propHyp[0], propHyp[1], propHyp[2]))
                histF[:, i] = propF
                histHyp[:, i] = propHyp
            return histF, histHyp
        F0 = np.zeros((chains_, y.shape[0]))
        H0 = np.tile(np.asarray(HYP0), (chains_, 1))
        ens = chains.ChainEnsemble(self.x, y, F0, H0, np.asarray(SCALE), seed=seed)
        histHyp, _, _, histF = ens.run(iters, thin_f=1)
        return histF, histHyp                                        # [chains_local, N, iters], [chains, 3, iters]

    def output(self, gap=0, histHyp=None, histF=None, llk=None, out_dir='./output'):
        os.makedirs(out_dir, exist_ok=True)
        if histHyp is not None:
            with open(os.path.join(out_dir, 'hypGap' + str(gap) + '.csv'), 'w', newline='') as h:
                writer = csv.writer(h)
                writer.writerow(["ll", "sf2", "sn"])
                writer.writerows(histHyp)
        if histF is not None:
            with open(os.path.join(out_dir, 'fGap' + str(gap) + '.csv'), 'w', newline='') as f:
                first_row = list(range(1, histF.shape[1] + 1)) + ["x", "y"]
                writer = csv.writer(f)
                writer.writerow(first_row)
                x = self.x.reshape((self.x.shape[0], -1))[:, :1]
                y = self.y.reshape((self.y.shape[0], 1))
                writer.writerows(np.hstack((histF, np.hstack((x, y)))))
        if llk is not None:
            with open(os.path.join(out_dir, 'llkGap' + str(gap) + '.csv'), 'w', newline='') as k:
                writer = csv.writer(k)
                writer.writerow(['gap'] + [str(i) for i in range(len(llk))])
                writer.writerow([gap] + list(llk))
        return 0


class singleRun(Framework):
    def execute(self, updOpt=None, iterMCMC=1000, out_dir='./output'):
        assert updOpt == 'mcmcSml', 'only the MCMC path is part of this package'
        histF, histHyp = self.runSimulMCMC(iterMCMC)                 # framework.py:164
        self.output(histHyp=histHyp.T, histF=histF, out_dir=out_dir)  # framework.py:165
        return histF, histHyp
