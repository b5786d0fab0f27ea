"""parsy_bench_b200 — B200-native executor for ParSy's numeric hot path.

Host side = thin ctypes mirror of the reference's executor entry points
(cholesky/parallel_PB_Cholesky_05.h:27, triangularSolve/Triangular_BCSC.h:14-238,
triangularSolve/Triangular_CSC.h:14-76) over the C ABI of ``libparsy_cuda.so``
(include/parsy_cuda.h).  There is no CPU fallback: importing :mod:`executor` without the built
library raises.
"""
from . import matrices  # noqa: F401

__all__ = ["matrices", "executor", "inspector"]
