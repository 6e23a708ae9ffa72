"""ncu target for the panel factor kernel alone: N=128 factorisations (one panel-factor launch each), B matrices."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
x = np.arange(n, dtype=np.float64).reshape(n, 1)
H = np.tile(np.array([[5.0, 4.0, 2.5]]), (B, 1))
A0 = gp.ops.cov_assemble(x, H, add_S=True, ld=(n + 15) // 16 * 16)
for _ in range(3):
    A = A0.clone()
    info = gp.ops.potrf_batched(A, n=n, jitter_policy=gp.JITTER_NONE, zero_upper=False)
torch.cuda.synchronize()
assert int((info != 0).sum().item()) == 0
print('ok')
