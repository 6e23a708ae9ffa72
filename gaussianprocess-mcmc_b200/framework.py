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


class crossValid(Framework):
    """Cross-validated runs: mirrors ``framework.py:177-248`` -- for every gap and fold hold out ``gap`` points out of
    every ``gap + window``, run the sampler on the rest, predict the held-out points from every 10th sample of the last
    tenth of the chain with ``inf_mcmc`` and score them with the truncated-Gaussian predictive likelihood."""

    def __init__(self, data, window, gapArray):
        super(crossValid, self).__init__(data, window, gapArray)
        self.windowSize = window
        self.gapArray = gapArray

    def getFoldData(self, fold, gap, window):
        """``framework.py:124-147`` (integer division made explicit for Python 3)."""
        test_id = []
        for i in range(self.x.shape[0] // (gap + window)):
            test_id = np.append(test_id, fold + np.arange(gap) + (gap + window) * i)
        test_id = np.asarray(test_id)
        test_id = test_id[test_id < self.x.shape[0]].astype('int')
        train_id = np.delete(np.arange(self.x.shape[0]), test_id)
        return self.x[train_id, :], self.y[train_id, 0], self.x[test_id, :], self.y[test_id, 0], test_id

    def execute(self, updOpt='mcmcSml', iterMCMC=1000, out_dir='./output'):
        from .kcGP import likK
        from .kcMCMC import sliceSample
        assert updOpt == 'mcmcSml', 'only the MCMC path is part of this package'
        originalX, originalY = self.x[:], self.y[:]
        results = {}
        for gap in self.gapArray:
            gapLLK = []
            for fold in range(gap + self.windowSize):
                self.x, self.y = originalX, originalY
                trX, trY, valX, valY, _ = self.getFoldData(fold, gap, window=self.windowSize)
                self.x, self.y = trX, trY.reshape(-1, 1)
                foldF, foldHyp = self.runSimulMCMC(iterMCMC)
                upper, lower = 100. - np.mean(self.y), 0. - np.mean(self.y)
                # every 10th sample of the last tenth of the chain (framework.py:223), predicted in ONE batched device pass
                sel = list(range(iterMCMC * 9 // 10 - 1, iterMCMC, 10))
                liks = {}

                def lik_for(sn):
                    liks[sn] = likK.TruncatedGauss2(upper=upper, lower=lower, log_sigma=np.log(sn))
                    return liks[sn]
                ys_all, _, _, fs2_all = sliceSample.inf_mcmc_batched(foldF[:, sel], foldHyp[:, sel].T, self.x, self.y, valX, lik_for)
                foldLLK = []
                for k, i in enumerate(sel):
                    trunclik = liks[foldHyp[2, i]]
                    trunclik.upper, trunclik.lower = 100., 0.                           # framework.py:241-242
                    foldLLK.append(trunclik.evaluate(y=ys_all[k], mu=valY.reshape(-1, 1), s2=fs2_all[k]) / ys_all[k].shape[0])
                gapLLK.append(float(np.mean(foldLLK)))
            self.output(gap, foldHyp.T, foldF, gapLLK, out_dir=out_dir)                 # framework.py:248 (last fold's chain and data)
            self.x, self.y = originalX, originalY
            results[gap] = gapLLK
        return results
