"""lamsa_b200 -- B200 (sm_100a) implementation of LAMSA's banded-DP hot path.

The product is ``liblamsa_b200.so`` (C ABI in ``include/lamsa_b200.h``): the
reference's ``ksw_*`` entry points plus a batch interface, all backed by the
CUDA kernels under ``lamsa_b200/csrc``.  This package is a thin ctypes binding
used by the tests and by ``bench.py``; it has no CPU implementation of the DP
and raises when the shared library or a GPU is missing.
"""
from ._lib import (LIB_PATH, LibraryMissing, load_library, TASK_DTYPE, RESULT_DTYPE,
                   AlnPara, KIND_GLOBAL, KIND_EXTEND, FLAG_CIGAR, FLAG_TARGET_PAC, FLAG_TARGET_REV)
from .ksw import (Context, Batch, make_tasks, default_matrix, pinned_pool,
                  ksw_global2, ksw_global, ksw_extend2, ksw_extend, ksw_extend_core,
                  ksw_extend_c, ksw_extend_r, ksw_bi_extend)

__all__ = [
    "LIB_PATH", "LibraryMissing", "load_library", "TASK_DTYPE", "RESULT_DTYPE", "AlnPara",
    "KIND_GLOBAL", "KIND_EXTEND", "FLAG_CIGAR", "FLAG_TARGET_PAC", "FLAG_TARGET_REV", "Context", "Batch", "make_tasks", "default_matrix", "pinned_pool",
    "ksw_global2", "ksw_global", "ksw_extend2", "ksw_extend", "ksw_extend_core",
    "ksw_extend_c", "ksw_extend_r", "ksw_bi_extend",
]
