"""One-off: a whole SDS transition at a LARGE N (default 8192, 2 chains; the look-ahead schedule, windows and 64-bit
indexing of the auxiliary-model sequence) against the CPU oracle on the same tape.  Prints bench.py's SDS sub-record."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
gp = importlib.import_module('gaussianprocess-mcmc_b200')

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rec = bench.sub_sds(gp, torch, np, 'large_n_sds', n, B, 0, 1, 0, 35.5, 1, 'N=%d x %d chains, one transition, tape-driven oracle check' % (n, B))
print(json.dumps(rec))
