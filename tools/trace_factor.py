"""Per-step timeline of one un-graphed, single-stream factorization: for every dependency step the device time of each
kernel class (CUDA events around every launch, parsy_cuda_factor_trace).  Usage: trace_factor.py <2d5|3d7|3d27> <N>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsy_bench_b200 import executor as ex, inspector, matrices  # noqa: E402

kind, N = sys.argv[1], int(sys.argv[2])
n, Ap, Ai, Ax = matrices.laplacian(kind, N)
S = inspector.analyze(n, Ap, Ai, Ax, 592, 1, 4)
H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels, S.levelPtr,
              S.parPtr, S.partition)
H.set_values(S.A2_x)
H.factor(); H.sync()
H.factor_trace()
step, cls, ms = H.factor_trace()
names = [k[:12] for k in ex.Solver.KERNEL_CLASSES]
print(f"# {kind} {N}: {len(ms)} launches, {ms.sum():.3f} ms serialised; per step: us per kernel class")
print("step " + " ".join(f"{k:>12s}" for k in names) + "        total")
for s in range(int(step.max()) + 1):
    m = step == s
    row = [1e3 * float(ms[m & (cls == c)].sum()) for c in range(6)]
    print(f"{s:4d} " + " ".join(f"{v:12.1f}" for v in row) + f" {sum(row):12.1f}")
H.factor(); H.sync()
print("graph factor times", H.factor_times(), "stats", {k: v for k, v in H.stats().items() if "launch" in k or "step" in k})
