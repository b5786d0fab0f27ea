"""Developer check (not collected by pytest; run by hand on a GPU box: python tests/dev_check.py 2d5 200): factor + solves
on one synthetic case through the C ABI, compared with the compiled reference.  Lives under tests/ because it executes
oracle/_ref, which only the tests, smoke() and bench.py's CPU legs may do."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refdump import ref_case, rel_err  # noqa: E402
from parsy_bench_b200 import executor as ex  # noqa: E402

kind, N = sys.argv[1], int(sys.argv[2])
cost = int(sys.argv[3]) if len(sys.argv) > 3 else 8
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.time()
R = ref_case(kind, N, cost=cost, threads=int(os.environ.get("REF_THREADS", "1")))
m = R.meta
print("ref:", {k: m[k] for k in ("n", "nsuper", "xsize", "nLevels", "nParts", "flops", "t_factor", "t_inspector")},
      f"({time.time() - t0:.1f}s)")
n, ns = m["n"], m["nsuper"]
t0 = time.time()
S = ex.Solver(n, R.A2_p, R.A2_i, R.p, R.s, R.i_ptr, R.super, ns, R.sParent, R.col2Sup, m["nLevels"], R.levelPtr,
              R.parPtr, R.partition, block_cols=nb)
print("create: %.3fs" % (time.time() - t0), S.stats())
S.set_values(R.A2_x)
FACTOR_ONLY = os.environ.get("FACTOR_ONLY") == "1"
for it in range(1 if FACTOR_ONLY else 3):
    S.factor()
    ok = S.sync()
    print("factor ok", ok, S.factor_times())
Lx = S.get_factor()
ref = R.valL
print("factor rel err:", rel_err(Lx, ref), " fro2:", float(Lx @ Lx), "nan:", int(np.isnan(Lx).sum()))
bad = np.argmax(np.abs(Lx - ref))
print("worst abs idx", bad, Lx[bad], ref[bad])
if FACTOR_ONLY:
    sys.exit(0)
# solves on our own factor
b = R.b_L1.copy()
S.set_rhs(b)
S.solve(ex.SOLVE_FWD)
x = S.get_rhs()
print("fwd (b=L*1): max|x-1| =", float(np.max(np.abs(x - 1))))
ramp = 1.0 + np.arange(n) / n
S.set_rhs(ramp)
S.solve(ex.SOLVE_FWD)
y = S.get_rhs()
print("fwd ramp rel err vs ref:", rel_err(y, R.y_ramp))
S.solve(ex.SOLVE_BWD)
xb = S.get_rhs()
# check L L' x = ramp via A2 (tril(PAP')) residual
import scipy.sparse as sp  # noqa: E402
A2 = sp.csc_matrix((R.A2_x, R.A2_i, R.A2_p), shape=(n, n))
Afull = A2 + sp.tril(A2, -1).T
res = np.linalg.norm(Afull @ xb - ramp) / np.linalg.norm(ramp)
print("full solve residual:", res)
