"""CPU oracle for the DiffOpt.jl sensitivity hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU arm.  The product path (``diffopt.jl_b200``) never routes through
this package and fails loudly when the CUDA library is missing.

The oracle is a numpy/scipy restatement of the reference's Julia algorithm
(andrewrosemberg/DiffOpt.jl v0.5.0).  The reference itself cannot run in the
build image (no Julia), and the arithmetic at the bottom of the path lives in
un-vendored Julia packages (``Project.toml:6-27``; no Manifest, so only compat
ranges are pinned):

* SparseArrays ``\\`` (UMFPACK LU)        -> numpy/LAPACK ``getrf/getrs`` or
                                             ``scipy.sparse.linalg.splu`` (SuperLU)
* IterativeSolvers 0.9 ``lsqr``           -> ``oracle.lsqr.lsqr`` (Paige-Saunders,
                                             cross-checked against scipy's lsqr)
* MathOptSetDistances 0.2.9 projections   -> ``oracle.cones``
* BlockDiagonals 0.1                      -> dense blocks, ``scipy.linalg.block_diag``

Parity pinning: every module is checked in ``tests/test_oracle_kat.py`` against
the known-answer literals and on-disk fixtures of the reference's own tests
(``test/quadratic_program.jl``, ``test/linear_program.jl``,
``test/conic_program.jl``, ``test/data/*.txt``), committed as
``tests/golden/*.json`` by ``tests/golden/make_golden.py``.  What stays
parity-UNPINNED (no reference test reaches it): the exact IterativeSolvers
stop-test arithmetic/defaults, the Nonnegatives gradient at exactly v == 0, the
``1e-4`` eigenvalue threshold edge of the PSD gradient, and every size above
n = 10.
"""
