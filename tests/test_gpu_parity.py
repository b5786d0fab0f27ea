"""GPU (-m gpu): parity of the CUDA executor, called through the C ABI, against
  * the golden vectors the compiled reference produced (tests/golden/),
  * the oracle's CPU restatement on seeded inputs of moderate size,
  * size-independent properties at BASELINE.json's full sizes (||L||_F^2 = trace(A), L*1 round trip, residual).
Tolerances: factor relative 1e-9 elementwise (north star) with the absolute floor 1e-6*max|L| for near-zero
entries; solves: residual within 10x of the oracle's."""
import os
import sys

import numpy as np
import pytest

from common import FULL_CASES, load_golden, rel_err, parse_case, full_matrix, View, as_view
from refdump import have_ref, ref_case
from parsy_bench_b200 import executor as ex, inspector, matrices, _lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import parsy_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-9


def dropin_factor(S, values=None, timing=None):
    n = len(S.col2Sup)
    lv = np.full(int(np.asarray(S.p)[n]), np.nan)   # the executor must not depend on the caller zeroing
    ok = ex.cholesky_left_par_05(n, S.A2_p, S.A2_i, S.A2_x if values is None else values, S.p, S.s, S.i_ptr, lv,
                                 S.super, S.nsuper, timing, S.sParent, S.A1_p, S.A1_i, S.col2Sup,
                                 len(S.levelPtr) - 1, S.levelPtr, None, 0, S.parPtr, S.partition, 1, 1, 0, 0)
    if not ok:
        print('cholesky_left_par_05 failed:', _lib.last_error())
    return ok, lv


def analyze(kind, N, c=8, l=1, d=2):
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    return inspector.analyze(n, Ap, Ai, Ax, c, l, d)


@pytest.mark.parametrize("name", FULL_CASES)
def test_factor_vs_reference_golden(name):
    G = load_golden(name)
    timing = np.zeros(4)
    ok, lv = dropin_factor(G, timing=timing)
    assert ok
    assert rel_err(lv, G.valL) < TOL
    assert np.array_equal(lv == 0.0, G.valL == 0.0)     # never-written entries stay exactly 0.0
    assert timing[0] >= 0 and timing[1] > 0


@pytest.mark.parametrize("case", [("2d5", 100, 8, 1, 2), ("2d5", 100, 592, 1, 4), ("3d7", 20, 8, 1, 2),
                                  ("3d27", 16, 8, 1, 2), ("3d27", 20, 148, 0, 4), ("2d5", 257, 37, 2, 2)])
def test_factor_vs_oracle(case):
    S = analyze(*case)
    ok, lv = dropin_factor(S)
    assert ok
    ref = orc.cholesky_left_par_05(S)
    assert rel_err(lv, ref) < TOL
    assert abs(float(lv @ lv) - float(S.A2_x[S.A2_p[:-1]].sum())) < 1e-10 * S.n * 26


def test_serial_twin_with_prune_set():
    S = analyze("2d5", 40)
    ptr = [0]
    pset = []
    for s in range(S.nsuper):
        pset.extend(S.ereach_sn(s).tolist())
        ptr.append(len(pset))
    lv = np.zeros(S.xsize)
    ok = ex.cholesky_left_sn_07(S.n, S.A2_p, S.A2_i, S.A2_x, S.p, S.s, S.i_ptr, lv, S.super, S.nsuper, None,
                                np.array(ptr, np.int32), np.array(pset, np.int32))
    assert ok and rel_err(lv, orc.cholesky_left_sn(S)) < TOL
    bad = np.array(ptr, np.int32)
    bad[-1] += 1
    assert not ex.cholesky_left_sn_07(S.n, S.A2_p, S.A2_i, S.A2_x, S.p, S.s, S.i_ptr, lv, S.super, S.nsuper, None,
                                      bad, np.array(pset + [0], np.int32))


@pytest.mark.parametrize("nb", [32, 64, 128])
def test_resident_handle_and_block_sizes(nb):
    S = analyze("3d27", 14, 16, 0, 2)
    ref = orc.cholesky_left_par_05(S)
    H = ex.Solver(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition, block_cols=nb)
    H.set_values(S.A2_x)
    for _ in range(2):                     # re-factoring on the same handle re-zeroes and re-assembles
        H.factor()
        assert H.sync()
        assert rel_err(H.get_factor(), ref) < TOL
    t = H.factor_times()
    st = H.stats()
    assert t["last_level"] > 0 and st["launches_factor"] > 0 and st["n_pairs"] == st["n_pairs_small"] + st["n_pairs_tiled"]
    # no-graph path gives the same factor
    H2 = ex.Solver(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                   S.levelPtr, S.parPtr, S.partition, block_cols=nb, use_graph=False, ignore_hlevels=True)
    H2.set_values(S.A2_x)
    H2.factor()
    assert H2.sync() and rel_err(H2.get_factor(), ref) < TOL
    H.close()
    H2.close()


@pytest.mark.parametrize("case", [("2d5", 100, 8, 1, 2), ("3d27", 14, 16, 2, 2)])
def test_factor_and_solve_on_the_dag_based_schedule(case):
    """The DAG-based LBC schedule over the factor's blocks (restated getCoarseLevelSet_DAG_BCSC02,
    cholesky/Inspection_DAG_02.h:15 — what analyze_DAG hands to cholesky_left_par_05 and H2LeveledBlockedLsolve in
    examples/triangularTest_DAG.cpp:100-125,281) drives the CUDA executor like the tree-based one."""
    S = analyze(*case)
    nl, lp, pp, part = inspector.dag_lbc_bcsc(S, case[2], case[3], case[4])
    V = View(as_view(S))
    V["levelPtr"], V["parPtr"], V["partition"] = lp, pp, part
    ok, lv = dropin_factor(V)
    assert ok
    ref = orc.cholesky_left_par_05(S)
    assert rel_err(lv, ref) < TOL
    b = orc.rhs_init_blocked(S, ref)
    x = b.copy()
    assert ex.H2LeveledBlockedLsolve(S.n, S.p, S.s, ref, int(S.xsize), S.i_ptr, S.col2Sup, S.super, S.nsuper, x, nl, lp,
                                     None, 0, pp, part, 1) == 1
    assert np.max(np.abs(x - 1.0)) < 1e-9


def test_run_to_run_spread_is_bounded():
    """Every update lands through red.global.add.f64, so the summation order — and with it the last bits of L — changes
    from run to run.  The spread between factorizations of the same matrix on the same handle stays at rounding level,
    three orders of magnitude inside the 1e-9 parity tolerance (wide separators: 3D 27-point, 22^3)."""
    S = analyze("3d27", 22, 148, 1, 4)
    H = ex.Solver(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    runs = []
    for _ in range(4):
        H.factor()
        assert H.sync()
        runs.append(H.get_factor())
    H.close()
    spread = max(rel_err(r, runs[0]) for r in runs[1:])
    print(f"run-to-run spread of the factor: {spread:.2e}")
    assert spread < 1e-12
    assert all(np.array_equal(r == 0.0, runs[0] == 0.0) for r in runs[1:])


def test_not_positive_definite_returns_false():
    G = load_golden("2d5_N30_c8_l1_d2")
    vals = G.A2_x.copy()
    vals[G.A2_p[G.n // 2]] = -4.0
    ok, _ = dropin_factor(G, values=vals)
    assert ok is False
    assert orc.cholesky_left_par_05(G, vals) is None     # same verdict as the reference semantics


def test_illegal_and_degenerate_schedules():
    G = load_golden("2d5_N30_c8_l1_d2")
    bad = View(G)
    bad["partition"] = G.partition[::-1].copy()
    ok, _ = dropin_factor(bad)
    assert ok is False
    # empty w-partitions are legal (InspectionLevel_06.h:302-319 can emit them)
    e = View(G)
    par = G.parPtr.tolist()
    e["parPtr"] = np.array([0, 0] + par[1:], np.int32)       # an empty partition in front
    lp = G.levelPtr.copy()
    lp[1:] += 1
    e["levelPtr"] = lp
    ok, lv = dropin_factor(e)
    assert ok and rel_err(lv, G.valL) < TOL
    # one H-level, one partition, supernode order (what cholesky_left_sn_07 runs)
    one = View(G)
    one["levelPtr"] = np.array([0, 1], np.int32)
    one["parPtr"] = np.array([0, G.nsuper], np.int32)
    one["partition"] = np.arange(G.nsuper, dtype=np.int32)
    ok, lv = dropin_factor(one)
    assert ok and rel_err(lv, G.valL) < TOL


@pytest.mark.parametrize("name", FULL_CASES)
def test_forward_solves_vs_reference_golden(name):
    G = load_golden(name)
    n, ns = G.n, G.nsuper
    nnz = int(G.meta["xsize"])
    for fn, extra in (
        (ex.blockedLsolve, ()),
        (ex.leveledBlockedLsolve, (len(G.etree_levelPtr) - 1, G.etree_levelPtr, G.etree_levelSet, 1)),
        (ex.H2LeveledBlockedLsolve, (len(G.levelPtr) - 1, G.levelPtr, None, 0, G.parPtr, G.partition, 1)),
        (ex.H2LeveledBlockedLsolve_Peeled, (len(G.levelPtr) - 1, G.levelPtr, None, 0, G.parPtr, G.partition, 1, 1)),
    ):
        x = G.b_L1.copy()
        assert fn(n, G.p, G.s, G.valL, nnz, G.i_ptr, G.col2Sup, G.super, ns, x, *extra) == 1
        assert np.max(np.abs(x - 1.0)) < 1e-10 and orc.test_triangular(x)      # known answer x == 1
        y = 1.0 + np.arange(n) / n
        assert fn(n, G.p, G.s, G.valL, nnz, G.i_ptr, G.col2Sup, G.super, ns, y, *extra) == 1
        assert rel_err(y, G.y_ramp) < 1e-11
    assert ex.blockedLsolve(n, None, G.s, G.valL, nnz, G.i_ptr, G.col2Sup, G.super, ns, G.b_L1.copy()) == 0


@pytest.mark.parametrize("name", FULL_CASES)
def test_csc_solves_vs_reference_golden(name):
    G = load_golden(name)
    n = G.n
    ramp = 1.0 + np.arange(n) / n
    x = ramp.copy()
    assert ex.lsolve(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, x) == 1
    assert rel_err(x, G.y_ramp_csc) < 1e-11
    # column level sets for lsolvePar: level(i) = 1 + max level of the columns that update i
    lev = np.zeros(n, np.int64)
    for j in range(n):
        rows = G.Lcsc_i[G.Lcsc_p[j] + 1:G.Lcsc_p[j + 1]]
        if len(rows):
            lev[rows] = np.maximum(lev[rows], lev[j] + 1)
    order = np.argsort(lev, kind="stable").astype(np.int32)
    lptr = np.concatenate([[0], np.cumsum(np.bincount(lev))]).astype(np.int32)
    x = ramp.copy()
    assert ex.lsolvePar(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, x, len(lptr) - 1, lptr, order, 1) == 1
    assert rel_err(x, G.y_ramp_csc) < 1e-11
    # lsolveParH2 (Triangular_CSC.h:76): H-levels x w-partitions over COLUMNS — the LBC schedule of the factorization
    # with every supernode expanded into its columns (sequential inside a w-partition, as the reference runs them)
    col_part, col_parptr = [], [0]
    for j1 in range(len(G.parPtr) - 1):
        for sn in G.partition[G.parPtr[j1]:G.parPtr[j1 + 1]]:
            col_part.extend(range(G.super[sn], G.super[sn + 1]))
        col_parptr.append(len(col_part))
    col_part, col_parptr = np.array(col_part, np.int32), np.array(col_parptr, np.int32)
    x = ramp.copy()
    assert ex.lsolveParH2(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, x, len(G.levelPtr) - 1, G.levelPtr, None, 0, col_parptr,
                          col_part, 1) == 1
    assert rel_err(x, G.y_ramp_csc) < 1e-11
    # a schedule that runs a column before one that updates it is refused (it would hang a dataflow kernel)
    x = ramp.copy()
    assert ex.lsolveParH2(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, x, len(G.levelPtr) - 1, G.levelPtr, None, 0, col_parptr,
                          col_part[::-1].copy(), 1) == 0
    assert ex.lsolveParH2(n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, x, 0, None, None, 0, None, None, 1) == 0
    assert ex.lsolve(n, None, G.Lcsc_i, G.Lcsc_x, x) == 0
    # resident form: structure once, several right-hand sides
    H = ex.CscSolver(n, G.Lcsc_p, G.Lcsc_i, order=col_part)
    H.set_values(G.Lcsc_x)
    for scale in (1.0, -2.5):
        x = ramp * scale
        ms = H.solve(x)
        assert ms > 0 and rel_err(x, G.y_ramp_csc * scale) < 1e-11
    H.close()


@pytest.mark.parametrize("case", [("2d5", 100, 8, 1, 2), ("3d27", 16, 8, 1, 2), ("3d7", 24, 148, 1, 4)])
def test_full_solve_residual_vs_oracle(case):
    """forward + NEW backward sweep: ||Ax-b||/||b|| within 10x of the oracle's residual on its own factor"""
    S = analyze(*case)
    n = S.n
    A = full_matrix(S)
    b = 1.0 + np.arange(n) / n
    Lref = orc.cholesky_left_par_05(S)
    xr = orc.blockedLtsolve(S, Lref, orc.blockedLsolve(S, Lref, b))
    res_ref = np.linalg.norm(A @ xr - b) / np.linalg.norm(b)
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync()
    H.set_rhs(b)
    H.solve(ex.SOLVE_FWD)
    y = H.get_rhs()
    assert rel_err(y, orc.blockedLsolve(S, Lref, b)) < 1e-9
    H.solve(ex.SOLVE_BWD)
    x = H.get_rhs()
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    assert res <= 10 * max(res_ref, np.finfo(float).eps), (res, res_ref)
    assert rel_err(x, xr) < 1e-8
    # drop-in backward sweep on the reference-layout arrays
    z = y.copy()
    assert ex.blockedLtsolve(n, S.p, S.s, Lref, int(S.xsize), S.i_ptr, S.col2Sup, S.super, S.nsuper, z) == 1
    assert rel_err(z, xr) < 1e-8
    H.close()


def test_per_step_sweeps_and_single_stream_factorization():
    """the alternative code paths kept behind options: one launch per dependency step for the sweeps
    (options.reserved[1]) and the single-stream factorization without look-ahead (options.reserved[0])"""
    S = analyze("3d27", 16, 8, 1, 2)
    n = S.n
    Lref = orc.cholesky_left_par_05(S)
    b = 1.0 + np.arange(n) / n
    yr = orc.blockedLsolve(S, Lref, b)
    xr = orc.blockedLtsolve(S, Lref, yr)
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition, lookahead=False, dataflow_sweeps=False)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync() and rel_err(H.get_factor(), Lref) < TOL
    H.set_rhs(b)
    H.solve(ex.SOLVE_FWD)
    assert rel_err(H.get_rhs(), yr) < 1e-9
    H.solve(ex.SOLVE_BWD)
    assert rel_err(H.get_rhs(), xr) < 1e-8
    st = H.stats()
    assert st["launches_fwd"] > 1 and st["launches_bwd"] > 1
    H.close()


@pytest.mark.parametrize("case", [("2d5", 120, 8, 1, 2), ("3d27", 16, 8, 1, 2), ("2d5", 64, 64, 0, 4)])
def test_narrow_sweep_kernels_agree_with_the_general_dataflow_kernel(case):
    """The leaf region of both sweeps runs on the light narrow-only kernels by default (two launches per sweep);
    options.reserved[5] = 1 keeps everything on the general kernel (one launch).  Both against the oracle."""
    S = analyze(*case)
    n = S.n
    Lref = orc.cholesky_left_par_05(S)
    b = 1.0 + np.arange(n) / n
    yr = orc.blockedLsolve(S, Lref, b)
    xr = orc.blockedLtsolve(S, Lref, yr)
    for narrow in (True, False):
        H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                      S.levelPtr, S.parPtr, S.partition, narrow_sweeps=narrow)
        H.set_factor(Lref)
        for _ in range(2):                               # twice: the counters are re-armed by every sweep
            H.set_rhs(b)
            H.solve(ex.SOLVE_FWD)
            assert rel_err(H.get_rhs(), yr) < 1e-9
            H.solve(ex.SOLVE_BWD)
            assert rel_err(H.get_rhs(), xr) < 1e-8
        st = H.stats()
        assert st["launches_fwd"] == st["launches_bwd"] <= (2 if narrow else 1)
        H.close()


@pytest.mark.parametrize("case", [("2d5", 1000, 8, 1, 2), ("3d27", 64, 8, 1, 2)])
def test_full_size_properties(case):
    """BASELINE.json configs 2 and 4 at full size: identities that need no CPU factorization."""
    S = analyze(*case)
    n = S.n
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync()
    L = H.get_factor()
    trace = float(S.A2_x[S.A2_p[:-1]].sum())
    assert abs(float(L @ L) - trace) < 1e-10 * trace                 # ||L||_F^2 = trace(A)
    b = orc.rhs_init_blocked(S, L)                                    # b = L*1 (Util.h:277)
    H.set_rhs(b)
    H.solve(ex.SOLVE_FWD)
    x = H.get_rhs()
    assert orc.test_triangular(x) and np.max(np.abs(x - 1.0)) < 1e-8  # Util.h:294, much tighter
    A = full_matrix(S)
    rhs = 1.0 + np.arange(n) / n
    H.set_rhs(rhs)
    H.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
    sol = H.get_rhs()
    assert np.linalg.norm(A @ sol - rhs) / np.linalg.norm(rhs) < 1e-9
    del L
    H.close()


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


FULL_SIZE = [pytest.param(("2d5", 1000, 592, 1, 4), id="cfg2"), pytest.param(("3d27", 64, 592, 1, 4), id="cfg4"),
             pytest.param(("3d7", 100, 592, 1, 4), id="cfg3-slow",
                          marks=pytest.mark.skipif(os.environ.get("PARSY_TEST_SLOW") != "1",
                                                   reason="cfg3: ~10 min of single-threaded reference + 35 GB of host "
                                                          "RAM; set PARSY_TEST_SLOW=1"))]


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", FULL_SIZE)
def test_full_size_factor_and_solves_vs_compiled_reference(case):
    """BASELINE.json configs 2 and 4 (3 behind PARSY_TEST_SLOW) at FULL size, elementwise against the reference itself:
    oracle/_ref/parsy_ref runs analyze_p2 + cholesky_left_par_05 + the forward solves with ONE OpenMP thread (the `top`
    race, parallel_PB_Cholesky_05.h:43,69,115; the sequential last level may use threaded BLAS, :271) on the same
    inspector triple; the GPU factor must agree within relative 1e-9 with an identical zero pattern, and all four
    supernodal forward solves (Triangular_BCSC.h:14,115,171,238) with the reference's x."""
    kind, N, c, l, d = case
    R = ref_case(kind, N, cost=c, level=l, div=d, threads=1, blas_threads=_host_threads(), csc=False, cache=False)
    assert R.meta["factor_ok"] == 1 and R.meta["solve_ok"] == 1
    S = analyze(kind, N, c, l, d)
    n = S.n
    for k in ("super", "p", "i_ptr", "partition", "parPtr", "levelPtr", "A2_i"):       # same structure on both sides
        assert np.array_equal(getattr(S, k), R[k][:len(getattr(S, k))]), k
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync()
    L = H.get_factor()
    ref = R.valL
    # chunked compare (cfg3: 1.03e9 entries): relative error with the absolute floor of common.rel_err, zero pattern
    floor = 1e-6 * float(np.max(np.abs(ref)))
    worst = 0.0
    for b0 in range(0, L.size, 1 << 26):
        a, b = L[b0:b0 + (1 << 26)], ref[b0:b0 + (1 << 26)]
        assert np.array_equal(a == 0.0, b == 0.0), "zero pattern differs"
        den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
        worst = max(worst, float(np.max(np.abs(a - b) / den)))
    print(f"{kind} N={N}: max rel err vs reference {worst:.3e} over {L.size} entries")
    assert worst < TOL
    # forward sweeps on the GPU's own factor: b = L*1 as the reference builds it (bit-equal b up to the factor's error)
    b = R.b_L1
    H.set_rhs(b)
    H.solve(ex.SOLVE_FWD)
    x = H.get_rhs()
    assert rel_err(x, R.x_blocked) < TOL and np.max(np.abs(x - 1.0)) < 1e-8
    ramp = 1.0 + np.arange(n) / n
    H.set_rhs(ramp)
    H.solve(ex.SOLVE_FWD)
    assert rel_err(H.get_rhs(), R.y_ramp) < TOL
    H.close()
    del L
    # the four drop-in forward solves on the REFERENCE's factor (host arrays in, host arrays out)
    nl, lp, ls = S.etree_level_set()
    common = (n, S.p, S.s, ref, int(S.xsize) & 0x7FFFFFFF, S.i_ptr, S.col2Sup, S.super, S.nsuper)
    runs = {
        "blockedLsolve": lambda x_: ex.blockedLsolve(*common, x_),
        "leveledBlockedLsolve": lambda x_: ex.leveledBlockedLsolve(*common, x_, nl, lp, ls, 1),
        "H2LeveledBlockedLsolve": lambda x_: ex.H2LeveledBlockedLsolve(*common, x_, S.nLevels, S.levelPtr, None, 0,
                                                                        S.parPtr, S.partition, 1),
        "H2LeveledBlockedLsolve_Peeled": lambda x_: ex.H2LeveledBlockedLsolve_Peeled(*common, x_, S.nLevels, S.levelPtr,
                                                                                      None, 0, S.parPtr, S.partition, 1, 1),
    }
    for name, f in runs.items():
        if kind == "3d7" and name != "H2LeveledBlockedLsolve":
            continue            # cfg3: one 8 GB upload instead of four
        x = b.copy()
        assert f(x) == 1, name
        assert rel_err(x, R.x_h2 if name.startswith("H2") else R.x_blocked) < TOL, name
        assert orc.test_triangular(x), name


def shifted_laplacian(kind, N, frac):
    """A - frac * lambda_min(A) * I for the Dirichlet Laplacian: still SPD, condition number 1/(1-frac) times larger."""
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    dims = 2 if kind == "2d5" else 3
    lam = dims * (2.0 - 2.0 * np.cos(np.pi / (N + 1)))       # smallest eigenvalue of the 5- / 7-point operator
    Ax = Ax.copy()
    Ax[Ap[:-1]] -= frac * lam
    return n, Ap, Ai, Ax, lam


@pytest.mark.parametrize("case", [("2d5", 90, 0.9999), ("3d7", 20, 0.999), ("2d5", 128, 0.99999)])
def test_ill_conditioned_factor_and_solve(case):
    """TRSM and the diagonal solves of the sweeps go through explicit inverses of the <= 128-wide diagonal blocks
    (k_potrf_block); a nearly singular SPD matrix — the Dirichlet Laplacian shifted by frac*lambda_min, condition number
    1e6-1e8 — is where that formulation would lose accuracy against the reference's substitution-based dtrsm /
    dlsolve_blas_nonUnit (MyBLAS.h:27-35, BLAS.h:8-103).  Factor: relative 1e-9 against the oracle; full solve:
    residual within 10x of the oracle's."""
    kind, N, frac = case
    n, Ap, Ai, Ax, lam = shifted_laplacian(kind, N, frac)
    S = inspector.analyze(n, Ap, Ai, Ax, 16, 1, 2)
    assert max(np.diff(S.super)) > 64          # block columns (and their inverse blocks) are on the path
    ref = orc.cholesky_left_par_05(S)
    assert ref is not None
    ok, lv = dropin_factor(S)
    assert ok
    err = rel_err(lv, ref)
    print(f"{kind} N={N} shift {frac}: cond ~ {(4 if kind == '2d5' else 6) * 2 / ((1 - frac) * lam):.1e}, factor rel err {err:.2e}")
    assert err < TOL
    rng = np.random.default_rng(3)
    bvec = rng.standard_normal(n)
    xo, ro = orc.solve_system(S, ref, bvec, refine_steps=0)
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync()
    H.set_permutation(S.Perm)
    xg, rg = H.solve_system(bvec, refine_steps=0, residuals=True)
    assert rg[0] <= 10 * max(ro[0], 1e-16), (rg, ro)
    xo1, ro1 = orc.solve_system(S, ref, bvec, refine_steps=1)
    xg1, rg1 = H.solve_system(bvec, refine_steps=1, residuals=True)
    assert rg1[-1] <= 10 * max(ro1[-1], 1e-16), (rg1, ro1)
    H.close()


@pytest.mark.parametrize("world,top,dist_top,case", [(2, 1, True, ("3d27", 14, 64, 1, 2)), (3, 2, True, ("3d27", 14, 64, 1, 2)),
                                                     (2, 2, False, ("3d27", 14, 64, 1, 2)), (4, 1, True, ("2d5", 120, 148, 1, 4)),
                                                     (8, 1, True, ("3d7", 22, 592, 1, 4))])
def test_sharded_factorization_and_solve_emulated_on_one_gpu(world, top, dist_top, case):
    """DESIGN.md §8 through parsy_cuda_sharded with every rank emulated on this one device (the NCCL collectives become
    device copies / a summing kernel, the plans, kernels and ownership rules are the ones the multi-GPU runs use):
    phase 1 (owned subtrees + fan-in of their updates into the top) -> sum of the top panels -> distributed top
    (owner factors, panel broadcast, owners update) or replicated top; then the sharded forward / backward sweeps.
    Factor against the oracle at 1e-9, solution against the oracle's restated sweeps."""
    S = analyze(*case)
    ref = orc.cholesky_left_par_05(S)
    sh = ex.Sharded(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                    S.levelPtr, S.parPtr, S.partition, 0, world, None, top_levels=top, top_distributed=dist_top)
    sh.set_values(S.A2_x)
    for _ in range(2):                       # a second factorization on the same handle re-zeroes what it must
        sh.factor()
        assert sh.sync()
    lv = sh.get_factor(np.full(S.xsize, np.nan))
    assert not np.isnan(lv).any()            # the ranks' subtrees and the top cover the whole factor
    assert rel_err(lv, ref) < TOL
    assert np.array_equal(lv == 0.0, ref == 0.0)
    owned = np.zeros(S.xsize, np.int32)
    for r in range(world):
        for b0, e0 in sh.plan(1, r).owned_ranges(r):
            owned[b0:e0] += 1
    for b0, e0 in sh.plan(1, 0).owned_ranges(-1):
        owned[b0:e0] += 1
    assert owned.min() == 1 and owned.max() == 1          # ownership is a partition of the panels
    st = sh.stats()
    assert st["top_supernodes"] > 0 and st["nccl_allreduces"] > 0 and (st["nccl_broadcasts"] > 0) == dist_top
    # sharded solve: L L' x = b in the factor's ordering
    rng = np.random.default_rng(11)
    bvec = rng.standard_normal(S.n)
    sh.set_rhs(bvec)
    sh.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
    x = sh.get_rhs()
    xr = orc.blockedLtsolve(S, ref, orc.blockedLsolve(S, ref, bvec))
    assert rel_err(x, xr) < 1e-8
    A = full_matrix(S)
    assert np.linalg.norm(A @ x - bvec) / np.linalg.norm(bvec) < 1e-10
    sh.close()


# ---- full system A x = b and Matrix-Market input (row 3) ---------------------------------
def make_solver(S):
    H = ex.Solver(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    H.set_values(S.A2_x)
    H.factor()
    assert H.sync()
    return H


def original_matrix(n, Ap, Ai, Ax):
    import scipy.sparse as sp
    Lo = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    return Lo + sp.tril(Lo, -1).T


@pytest.mark.parametrize("case,refine", [(("2d5", 60, 8, 1, 2), 0), (("3d27", 12, 8, 1, 2), 2), (("3d7", 16, 16, 1, 2), 1)])
def test_solve_system_vs_oracle(case, refine):
    """x = P'(LL')^-1 P b in the caller's ordering, right-hand side of the reference driver (choleskyTest01.cpp:428-432);
    solution against the oracle's restated sweeps, residual within 10x of the oracle's (north star)."""
    kind, N = case[0], case[1]
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    S = inspector.analyze(n, Ap, Ai, Ax, *case[2:])
    H = make_solver(S)
    H.set_permutation(S.Perm)
    b = 1.0 + np.arange(n) / n
    x, rel = H.solve_system(b, refine_steps=refine, residuals=True)
    xo, relo = orc.solve_system(S, orc.cholesky_left_par_05(S), b, refine_steps=refine)
    assert rel.shape == (refine + 1,)
    assert np.max(np.abs(x - xo)) <= 1e-9 * np.max(np.abs(xo))
    A = original_matrix(n, Ap, Ai, Ax)
    true_res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    oracle_res = float(np.linalg.norm(A @ xo - b) / np.linalg.norm(b))
    assert true_res <= 10 * max(oracle_res, 1e-16)
    # the device-side residual is the residual
    assert abs(rel[-1] - true_res) <= 0.5 * true_res + 1e-17
    assert rel[-1] <= 10 * max(relo[-1], 1e-16)
    H.close()


def test_solve_system_many_right_hand_sides_and_identity_permutation():
    S = analyze("2d5", 40, 8, 1, 2)
    H = make_solver(S)
    rng = np.random.default_rng(5)
    B = rng.normal(size=(5, S.n))
    # without set_permutation the system is the permuted one, P A P'
    A2 = full_matrix(S)
    X = H.solve_system(B)
    assert X.shape == B.shape
    for j in range(5):
        assert np.linalg.norm(A2 @ X[j] - B[j]) <= 1e-12 * np.linalg.norm(B[j]) * 40
    # same columns one at a time, and in the original ordering
    H.set_permutation(S.Perm)
    X2, rel = H.solve_system(B, refine_steps=1, residuals=True)
    assert rel.shape == (5, 2) and np.all(rel[:, 1] <= rel[:, 0] * 1.0000001 + 1e-16)
    for j in range(5):
        xo, _ = orc.solve_system(S, orc.cholesky_left_par_05(S), B[j], refine_steps=1)
        assert np.max(np.abs(X2[j] - xo)) <= 1e-9 * np.max(np.abs(xo))
    # one column alone gives the same answer (up to the summation order of the atomics)
    assert np.allclose(H.solve_system(B[0].copy()), X2[0], rtol=1e-10, atol=1e-12)
    H.close()


def test_solve_system_argument_errors():
    S = analyze("2d5", 12, 8, 1, 2)
    H = ex.Solver(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                  S.levelPtr, S.parPtr, S.partition)
    with pytest.raises(ex.ParsyCudaError) as e:
        H.solve_system(np.ones(S.n))                      # before factor
    assert e.value.code == ex.ERR_STATE
    bad = S.Perm.copy()
    bad[0] = bad[1]
    with pytest.raises(ex.ParsyCudaError) as e:
        H.set_permutation(bad)
    assert e.value.code == ex.ERR_BAD_ARG
    # a handle that only received a factor has no A: sweeps work, residuals do not
    H.set_factor(orc.cholesky_left_par_05(S))
    x = H.solve_system(np.ones(S.n))
    assert np.linalg.norm(full_matrix(S) @ x - 1.0) < 1e-10 * np.sqrt(S.n)
    with pytest.raises(ex.ParsyCudaError) as e:
        H.solve_system(np.ones(S.n), refine_steps=1)
    assert e.value.code == ex.ERR_STATE
    H.close()


def test_matrix_market_file_through_inspector_and_executor(tmp_path):
    """A non-stencil SPD matrix written as a Matrix-Market lower half, read back with the restated readMatrix
    (common/Util.h:77), analysed, factored on the GPU and compared with the oracle; then A x = b."""
    from test_mmio import random_spd_lower
    n, Ap, Ai, Ax = random_spd_lower(700, 3, seed=11)
    f = tmp_path / "rand.mtx"
    matrices.write_mtx(f, n, Ap, Ai, Ax)
    n2, Bp, Bi, Bx = inspector.read_matrix(f)
    assert n2 == n and np.array_equal(Bx, Ax)
    S = inspector.analyze(n2, Bp, Bi, Bx, 16, 1, 2)
    ok, lv = dropin_factor(S)
    assert ok
    assert rel_err(lv, orc.cholesky_left_par_05(S)) < TOL
    H = make_solver(S)
    H.set_permutation(S.Perm)
    b = 1.0 + np.arange(n) / n
    x = H.solve_system(b)
    A = original_matrix(n, Ap, Ai, Ax)
    assert np.linalg.norm(A @ x - b) <= 1e-12 * np.linalg.norm(b)
    H.close()


@pytest.mark.parametrize("n,per_col,seed", [(3000, 4, 21), (500, 40, 22)])
def test_csc_solves_of_a_general_triangular_matrix(tmp_path, n, per_col, seed):
    """Non-chordal input (examples/triangularTest_DAG.cpp:171-175): a random lower-triangular CSC matrix that is not a
    Cholesky factor; level sets from the restated buildLevelSet_CSC, column solves lsolve / lsolvePar on the GPU against
    scipy's triangular solve and, where the compiled reference is present, against its own lsolve / lsolvePar."""
    import subprocess
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve_triangular
    import refdump
    from test_mmio import random_lower_triangular
    n, Lp, Li, Lx = random_lower_triangular(n, per_col, seed)
    b = 1.0 + np.arange(n) / n
    want = spsolve_triangular(sp.csc_matrix((Lx, Li, Lp), shape=(n, n)).tocsr(), b, lower=True)
    x = b.copy()
    assert ex.lsolve(n, Lp, Li, Lx, x) == 1
    assert rel_err(x, want) < 1e-11
    levels, lptr, lset = inspector.build_level_set_csc(n, Lp, Li)
    assert levels > 1
    x = b.copy()
    assert ex.lsolvePar(n, Lp, Li, Lx, x, levels, lptr, lset, 1) == 1
    assert rel_err(x, want) < 1e-11
    # lsolveParH2 on the LBC schedule of the column DAG (restated getCoarseLevelSet_DAG_CSC03,
    # examples/triangularTest_DAG_nonChordal.cpp:357,405)
    nl, hlp, hpp, hpart = inspector.dag_lbc_csc(n, Lp, Li, 8, 2, 2)
    xh = b.copy()
    assert ex.lsolveParH2(n, Lp, Li, Lx, xh, nl, hlp, None, 0, hpp, hpart, 1) == 1
    assert rel_err(xh, want) < 1e-11
    if refdump.have_ref():
        f = tmp_path / "tri.mtx"
        matrices.write_mtx(f, n, Lp, Li, Lx, symmetric=False)
        d = tmp_path / "dump"
        d.mkdir()
        subprocess.run([refdump.REF_BIN, "--mtx", str(f), "--tri-only", "--dump", str(d), "--cost", "8", "--level", "2",
                        "--div", "2"], check=True, capture_output=True)
        assert rel_err(x, np.fromfile(d / "tri_x.f64", np.float64)) < 1e-11
        assert rel_err(x, np.fromfile(d / "tri_x_par.f64", np.float64)) < 1e-11
        assert np.array_equal(hpart, np.fromfile(d / "dag_partition.i32", np.int32))      # same schedule ...
        assert rel_err(xh, np.fromfile(d / "tri_x_h2.f64", np.float64)) < 1e-11           # ... same solution
