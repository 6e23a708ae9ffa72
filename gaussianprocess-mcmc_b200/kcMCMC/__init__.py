"""Drop-in for the reference's ``kcMCMC`` package (``kcMCMC/__init__.py:1-3``): ``sliceSample`` plus the ``sdsK``
alias that ``framework.py:10`` imports (a stale name of the same module in the reference)."""
from . import sliceSample
from . import sliceSample as sdsK

__all__ = ["sliceSample", "sdsK"]
