"""Import alias: the package directory is ``gaussianprocess-mcmc_b200`` (a hyphen cannot appear in an
``import`` statement), so ``import gpmc_b200`` loads it by name and re-exports it."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module('gaussianprocess-mcmc_b200')
sys.modules[__name__] = _pkg
