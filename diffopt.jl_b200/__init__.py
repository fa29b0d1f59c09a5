"""diffopt.jl_b200 -- B200-native (sm_100a) sensitivity hot path of DiffOpt.jl.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/diffopt_b200.h``), ``lib/`` (the built ``libdiffopt_b200.so``), and a thin host-side
mirror of the reference's backend interface (``qp``, ``conic``, ``lsqr``).  The directory name
contains a dot, so import it through the repo-root shim: ``import diffopt_b200``.
"""
from . import _capi  # noqa: F401
from ._capi import Context, DiffOptB200Error, SingularException, load  # noqa: F401
