"""Dev tool (one GPU): the column-solve dataflow kernel on small structured matrices against scipy."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import scipy.sparse as sp, scipy.sparse.linalg as spl
from parsy_bench_b200 import executor as ex
def run(name, L, order=None):
    L = sp.csc_matrix(L); L.sort_indices()
    n = L.shape[0]
    Lp, Li, Lx = L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data.astype(np.float64)
    b = 1.0 + np.arange(n) / n
    ref = spl.spsolve_triangular(sp.csr_matrix(L), b, lower=True)
    H = ex.CscSolver(n, Lp, Li, order=order); H.set_values(Lx)
    x = b.copy(); H.solve(x)
    err = np.abs(x - ref) / np.maximum(np.abs(ref), 1e-300)
    bad = np.flatnonzero(err > 1e-10)
    print(name, "n", n, "nnz", len(Li), "max rel err", err.max(), "first bad", bad[:5], "count", len(bad))
    if len(bad):
        r = bad[0]
        Ld = L.toarray()
        res = Ld[r, :r] @ x[:r] + Ld[r, r] * x[r] - b[r]
        terms = Ld[r, :r] * x[:r]
        k = np.argmin(np.abs(terms + res)), np.argmin(np.abs(terms - res))
        print("   row", r, "residual", res, "closest -term col", k[0], terms[k[0]], "closest +term col", k[1], terms[k[1]])
    H.close()
rng = np.random.default_rng(0)
for n in (33, 34, 35, 40, 64, 66, 100):
    run(f"dense{n}", np.tril(rng.uniform(-0.5, 0.5, (n, n)) / n) + 2 * np.eye(n))
n = 64
M = np.tril(rng.uniform(-0.5, 0.5, (n, n)) / n) + 2 * np.eye(n)
M[1:, 0] = 0; M[40, 0] = 0.3        # only one long-range entry
run("one_far_entry", M)
M = 2 * np.eye(n); M[1:, 0] = 0.01   # a single dense column
run("single_dense_column", M)
