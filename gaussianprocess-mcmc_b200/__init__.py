"""gaussianprocess-mcmc_b200: the GP log-marginal-likelihood hot path of
t-kychen/GaussianProcess-MCMC (``kcMCMC/sliceSample.py``) as hand-written sm_100a CUDA behind the
reference's own Python interface.  Import as ``import gpmc_b200`` (alias module at the repo root)."""
from . import _lib, ops, synthetic, chains, kcMCMC, kcGP, framework      # noqa: F401
from ._lib import GpmcError, JITTER_NONE, JITTER_PYGPS  # noqa: F401

__all__ = ['ops', 'synthetic', 'chains', 'kcMCMC', 'kcGP', 'framework', 'GpmcError', 'JITTER_NONE', 'JITTER_PYGPS']
