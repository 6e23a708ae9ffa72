import os, sys, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['GPMC_DEBUG'] = '1'
import torch
import gpmc_b200 as gp
from gpmc_b200 import ops
from oracle import sds_oracle as so, kcgp_shim
from oracle.reference_loader import Tape
for path in sorted(glob.glob('tests/golden/sds_N*.npz')):
    z = np.load(path)
    F = torch.tensor(z['f'][None].copy()).cuda(); H = torch.tensor(z['hyp'][None].copy()).cuda()
    tape = ops.Tape(z['z'][None], z['v'][None], [float(z['u0'])], z['U'][None])
    nt, ll, st = ops.sds_sweep(z['x'], z['y'], F, H, z['scale'], int(z['it']), tape=tape)
    f = F.cpu().numpy()[0]
    print(os.path.basename(path), 'trips', int(nt.item()), int(z['ref_trips']), 'df %.3e' % np.abs(f - z['ref_prop_f']).max(), flush=True)
    # how close to indefinite is R + 1e-11 I in the oracle at the accepted theta?
    K = so.cov_matrix(z['x'], z['ref_prop_hyp'])
    g, KS, m, C, L = so.aux_var_model(z['f'], K, z['ref_prop_hyp'][2], g=z['ref_g'], r_form='reduced')
    print('   min diag C (reduced form oracle) %.3e' % np.diag(C).min(), flush=True)
