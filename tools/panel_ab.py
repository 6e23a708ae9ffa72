import json, os, sys
sys.path.insert(0, '/root/repo')
sys.argv=['x']
import gpmc_b200 as gp
from tools.panel_timing import time_potrf
for mode in (0, 3):
    gp.ops.set_tuning(1, mode)
    for n, B in ((128, 1), (128, 296), (128, 4096), (512, 4096)):
        r = time_potrf(n, B)
        print(json.dumps({'potf2_mode': mode, 'n': n, 'B': B, 'us': {k: round(v[0], 1) for k, v in r.items()}}))
gp.ops.set_tuning(1, 0)
