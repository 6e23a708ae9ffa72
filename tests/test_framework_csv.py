"""CSV writer of the Framework mirror (framework.py:79-122 format); no GPU needed."""
import os

import numpy as np


def test_output_files_have_reference_format(gp, tmp_path):
    n, iters = 7, 3
    data = np.column_stack([np.linspace(80, 95, n), np.arange(n, dtype=float)])
    fw = gp.framework.Framework(data)
    histHyp = np.arange(3 * iters, dtype=float).reshape(3, iters)
    histF = np.arange(n * iters, dtype=float).reshape(n, iters)
    assert fw.output(gap=1, histHyp=histHyp.T, histF=histF, llk=[-1.5, -2.5], out_dir=str(tmp_path)) == 0
    hyp = open(os.path.join(str(tmp_path), 'hypGap1.csv')).read().splitlines()
    assert hyp[0] == 'll,sf2,sn' and len(hyp) == 1 + iters and hyp[1] == '0.0,3.0,6.0'
    f = open(os.path.join(str(tmp_path), 'fGap1.csv')).read().splitlines()
    assert f[0] == '1,2,3,x,y' and len(f) == 1 + n
    assert [float(v) for v in f[1].split(',')] == [0.0, 1.0, 2.0, 0.0, 80.0]
    llk = open(os.path.join(str(tmp_path), 'llkGap1.csv')).read().splitlines()
    assert llk[0] == 'gap,0,1' and llk[1] == '1,-1.5,-2.5'
    import pytest
    with pytest.raises(Exception):
        gp.framework.Framework(None)
