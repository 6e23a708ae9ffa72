"""``kcGP.tools``: ``jitchol`` (``sliceSample.py:196,205,257``) and ``solve_chol`` (``:258``) on the B200."""
import numpy as np

from .. import ops, _lib


def jitchol(A, maxtries=5):
    """Lower Cholesky factor with the pyGPs jitter ladder; raises ``numpy.linalg.LinAlgError`` when it gives up."""
    import torch
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    ld = n + (n & 1)
    T = torch.zeros((1, n, ld), dtype=torch.float64, device='cuda')
    T[0, :, :n] = torch.as_tensor(A)
    info = ops.potrf_batched(T, jitter_policy=_lib.JITTER_PYGPS, zero_upper=True, n=n)
    code = int(info.item())
    if code != 0:
        raise np.linalg.LinAlgError('not positive definite, even with jitter.' if code == -1 else
                                    'leading minor %d not positive definite' % code)
    return T[0, :, :n].cpu().numpy()


def solve_chol(L, B):
    """``(L^T L)^-1 B`` for UPPER triangular ``L``: two triangular solves (forward with ``L^T``; the backward solve is the
    same forward kernel on the index-reversed matrix)."""
    import torch
    L = torch.as_tensor(np.asarray(L, dtype=np.float64)).cuda()
    Bm = np.asarray(B, dtype=np.float64)
    vec = Bm.ndim == 1
    rhs = torch.as_tensor(Bm.reshape(Bm.shape[0], -1).T.copy()).cuda()            # one right-hand side per row
    z = ops.trsv_lower(L.T.contiguous(), rhs)                                     # L^T z = b
    Lrev = torch.flip(L, dims=(0, 1)).contiguous()                                # J L J is lower triangular
    xr = ops.trsv_lower(Lrev, torch.flip(z, dims=(1,)).contiguous())              # (J L J)(J x) = J z
    x = torch.flip(xr, dims=(1,)).T.cpu().numpy()
    return x.reshape(-1) if vec else x
