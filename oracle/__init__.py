"""CPU oracle for the FDI / P1 assemble -> solve -> evaluate path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``loopstructural_b200`` (the product)
imports this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
only as the checker / timed CPU baseline.

The functions here restate, in plain numpy/scipy, the algorithm of the
reference (LoopStructural 1.8.0, ``/root/reference``) for the hot path; every
function cites the reference file:line it follows.  The restatement is pinned
against outputs of the reference itself (``tests/golden/make_golden.py`` runs
the reference in the build container and freezes its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle
against them).
"""
