"""CPU oracle for the GP log-marginal-likelihood / surrogate-data slice-sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or the
CPU arm that is timed beside the GPU path), never as the thing shipped.

Parity status
-------------
* ``kcMCMC/sliceSample.py`` (the reference's hot path) is PINNED: the restatement
  in :mod:`oracle.sds_oracle` is checked against the reference file itself, run
  unmodified in the build container (``oracle/make_golden.py`` ->
  ``tests/golden/*.npz``).
* The ``kcGP`` primitives the reference imports (``covK.RBF``, ``tools.jitchol``,
  ``tools.solve_chol``, ``likK.TruncatedGauss2``) are NOT in the reference tree
  (git-ignored private fork of pyGPs 1.3.4, ``requirements.txt:10``).  They are
  restated in :mod:`oracle.kcgp_shim` from the published pyGPs 1.3.4 algorithm;
  that boundary is "parity unpinned" (no reference test or golden vector exists
  for it), and ``TruncatedGauss2`` is an author-private class whose definition is
  an explicit assumption (ASSUMPTION-1 in the shim).
"""
