"""Minimal driver for ncu: own inspector -> one (or a few) resident factorization(s) [+ solves] through the C ABI."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsy_bench_b200 import executor as ex, inspector, matrices  # noqa: E402

kind, N = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
solves = len(sys.argv) > 4 and sys.argv[4] == "solve"
n, Ap, Ai, Ax = matrices.laplacian(kind, N)
S = inspector.analyze(n, Ap, Ai, Ax, 592, 1, 4)
H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels, S.levelPtr,
              S.parPtr, S.partition)
H.set_values(S.A2_x)
for _ in range(reps):
    H.factor()
    assert H.sync()
    print("factor", H.factor_times())
if solves:
    H.set_rhs(np.ones(n))
    H.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
    H.sync()
print("stats", H.stats())
