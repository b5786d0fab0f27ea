"""CPU: pins oracle/parsy_oracle.c (the restatement the GPU parity tests trust) against the golden vectors that the
compiled reference produced, and against the reference run live where oracle/_ref exists."""
import os
import sys

import numpy as np
import pytest

from common import CPU_CASES, FULL_CASES, load_golden, rel_err, parse_case, View
from refdump import have_ref, ref_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import parsy_oracle as orc  # noqa: E402

# the reference's BLAS sums in a different order than the scalar restatement: 1e-12 leaves 3 orders of headroom
# to the 1e-9 gate of the GPU tests
TOL_FACTOR = 1e-12


@pytest.mark.parametrize("name", CPU_CASES)
def test_factor_matches_reference_golden(name):
    G = load_golden(name)
    lv = orc.cholesky_left_par_05(G)
    assert lv is not None
    assert rel_err(lv, G.valL) < TOL_FACTOR
    lv2 = orc.cholesky_left_sn(G)   # serial twin (PB_Cholesky.h:16): same arithmetic, different traversal
    assert rel_err(lv2, G.valL) < TOL_FACTOR
    # ||L||_F^2 == trace(A) (exact identity of the Cholesky factor)
    assert abs(float(lv @ lv) - G.meta["trace"]) < 1e-10 * G.meta["trace"]
    # entries the reference never writes stay exactly zero
    assert np.array_equal(lv == 0.0, G.valL == 0.0)


@pytest.mark.parametrize("name", CPU_CASES)
def test_solves_match_reference_golden(name):
    G = load_golden(name)
    b = orc.rhs_init_blocked(G, G.valL)
    assert np.array_equal(b, G.b_L1)                      # rhsInitBlocked is a plain sum: bit-exact
    x = orc.blockedLsolve(G, G.valL, b)
    assert np.max(np.abs(x - G.x_blocked)) < 1e-13
    assert orc.test_triangular(x)
    x2 = orc.H2LeveledBlockedLsolve(G, G.valL, b)
    assert np.max(np.abs(x2 - G.x_h2)) < 1e-13
    n = G.n
    ramp = 1.0 + np.arange(n) / n
    assert rel_err(orc.blockedLsolve(G, G.valL, ramp), G.y_ramp) < 1e-13
    yc = orc.lsolve(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, ramp)
    assert np.array_equal(yc, G.y_ramp_csc)                # same loop, same order: bit-exact


@pytest.mark.parametrize("name", CPU_CASES)
def test_backward_sweep_residual(name):
    """The reference has no L' solve: the restated backward sweep is pinned by the residual of L L' x = b."""
    import scipy.sparse as sp
    G = load_golden(name)
    n = G.n
    ramp = 1.0 + np.arange(n) / n
    y = orc.blockedLsolve(G, G.valL, ramp)
    x = orc.blockedLtsolve(G, G.valL, y)
    A2 = sp.csc_matrix((G.A2_x, G.A2_i, G.A2_p), shape=(n, n))
    A = A2 + sp.tril(A2, -1).T
    assert np.linalg.norm(A @ x - ramp) / np.linalg.norm(ramp) < 1e-13


def test_ereach_counts_match_pairs():
    G = load_golden("2d5_N30_c8_l1_d2")
    total = sum(len(orc.ereach_sn(G, s)) for s in range(G.nsuper))
    # every descendant has rows inside the target: the pair count is the number of maximal row runs
    runs = 0
    for d in range(G.nsuper):
        c0, c1 = G.super[d], G.super[d + 1]
        rows = G.s[int(G.i_ptr[c0]) + (c1 - c0): int(G.i_ptr[c1 - 1 + 1]) if c1 < G.n else len(G.s)]
        rows = G.s[int(G.pi[d]) + (c1 - c0): int(G.pi[d + 1])]
        sup = G.col2Sup[rows]
        runs += int(np.count_nonzero(np.diff(sup)) + (1 if len(sup) else 0))
    assert total == runs


def test_not_spd_reports_failure():
    G = load_golden("2d5_N12_c8_l1_d2")
    vals = G.A2_x.copy()
    vals[G.A2_p[G.n // 2]] = -4.0      # negative diagonal entry
    assert orc.cholesky_left_par_05(G, vals) is None


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", [("2d5", 100, 8, 1, 2), ("3d7", 16, 8, 1, 2), ("3d27", 12, 16, 0, 2)])
def test_oracle_vs_live_reference(case):
    R = ref_case(case[0], case[1], cost=case[2], level=case[3], div=case[4])
    S = View(R)
    S["nsuper"] = R.meta["nsuper"]
    lv = orc.cholesky_left_par_05(S)
    assert rel_err(lv, R.valL) < TOL_FACTOR
    b = orc.rhs_init_blocked(S, R.valL)
    assert np.array_equal(b, R.b_L1)
    assert np.max(np.abs(orc.blockedLsolve(S, R.valL, b) - R.x_blocked)) < 1e-12


def test_oracle_solve_system_solves_the_original_system():
    """oracle.solve_system (the checker of parsy_cuda_solve_system): permutation + restated sweeps + refinement."""
    import scipy.sparse as sp
    from parsy_bench_b200 import inspector, matrices
    n, Ap, Ai, Ax = matrices.laplacian("3d7", 9)
    S = inspector.analyze(n, Ap, Ai, Ax, 8, 1, 2)
    Lo = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    A = Lo + sp.tril(Lo, -1).T
    b = 1.0 + np.arange(n) / n                      # examples/choleskyTest01.cpp:428-432
    lv = orc.cholesky_left_par_05(S)
    x, rel = orc.solve_system(S, lv, b, refine_steps=2)
    true = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    assert true < 1e-14 and len(rel) == 3
    assert abs(rel[-1] - true) <= 0.5 * true + 1e-17
    # the residual helper is b - (PAP')x
    A2 = sp.csc_matrix((S.A2_x, S.A2_i, S.A2_p), shape=(n, n))
    A2 = A2 + sp.tril(A2, -1).T
    v = np.cos(np.arange(n))
    assert np.allclose(orc.residual_sym_lower(S, v, b), b - A2 @ v, rtol=0, atol=1e-12)
