import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden_files(pattern):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load_rows(path):
    """Unpack the ``key_i`` layout written by oracle/make_golden.py into a list of dicts."""
    z = np.load(path)
    n = int(z['n_cases'])
    keys = sorted({k.rsplit('_', 1)[0] for k in z.files if k != 'n_cases'})
    return [{k: z['%s_%d' % (k, i)] for k in keys} for i in range(n)]


@pytest.fixture(scope='session')
def gp():
    """The product package (directory name has a hyphen, so it is imported through its alias)."""
    import gpmc_b200
    return gpmc_b200
