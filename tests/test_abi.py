"""CPU: the C-ABI libraries load and export every symbol their headers declare; host-only logic of the executor
library (planner, schedule validation, error paths) behaves without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import load_golden
from parsy_bench_b200 import _lib, executor as ex, inspector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared("parsy_cuda.h", "parsy_cuda_")
    assert len(names) >= 30
    for nme in names:
        assert hasattr(L, nme), nme


def test_inspector_library_exports_every_declared_symbol():
    L = inspector.lib()
    for nme in declared("parsy_inspector.h", "parsy_"):
        assert hasattr(L, nme), nme


def test_version_and_error_string():
    L = _lib.lib()
    assert L.parsy_cuda_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_plan_check_counts_pairs():
    G = load_golden("2d5_N30_c8_l1_d2")
    rc, st = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, len(G.levelPtr) - 1, G.levelPtr,
                           G.parPtr, G.partition)
    assert rc == ex.OK
    assert st["nsuper"] == G.nsuper and st["xsize"] == G.meta["xsize"] and st["ssize"] == G.meta["ssize"]
    assert st["n_pairs"] == st["n_pairs_small"] + st["n_pairs_tiled"] > 0
    # executed supernodal flops bound the simplicial count from above (relaxed supernodes store explicit zeros)
    assert st["flops_potrf"] + st["flops_trsm"] + st["flops_update"] >= G.meta["flops"] * 0.99


def test_plan_check_rejects_illegal_schedules():
    G = load_golden("2d5_N30_c8_l1_d2")
    nl = len(G.levelPtr) - 1
    rev = G.partition[::-1].copy()                    # parents before children
    rc, _ = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, nl, G.levelPtr, G.parPtr, rev)
    assert rc == ex.ERR_BAD_SCHEDULE
    dup = G.partition.copy()
    dup[0] = dup[1]                                    # a supernode listed twice
    rc, _ = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, nl, G.levelPtr, G.parPtr, dup)
    assert rc == ex.ERR_BAD_SCHEDULE
    rc, _ = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, nl, G.levelPtr, G.parPtr,
                          G.partition, block_cols=100)  # not a multiple of 8
    assert rc == ex.ERR_BAD_ARG


def test_plan_without_schedule_and_block_sizes():
    G = load_golden("3d27_N6_c4_l0_d2")
    for nb in (0, 32, 64, 128):
        rc, st = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, 0, None, None, None,
                               block_cols=nb)
        assert rc == ex.OK and st["n_steps"] >= 1


def test_etree_levels_are_a_legal_schedule():
    """leveledBlockedLsolve's schedule = getLevelSet waves with one supernode per w-partition"""
    G = load_golden("3d7_N7_c8_l1_d2")
    parPtr = np.arange(G.nsuper + 1, dtype=np.int32)
    rc, st = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, len(G.etree_levelPtr) - 1,
                           G.etree_levelPtr, parPtr, G.etree_levelSet)
    assert rc == ex.OK


@pytest.mark.skipif(ex.device_count() > 0, reason="checks the no-device error path")
def test_no_device_is_an_error_not_a_fallback():
    G = load_golden("2d5_N12_c8_l1_d2")
    with pytest.raises(ex.ParsyCudaError) as e:
        ex.Solver(G.n, G.A2_p, G.A2_i, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.sParent, G.col2Sup,
                  len(G.levelPtr) - 1, G.levelPtr, G.parPtr, G.partition)
    assert e.value.code == ex.ERR_NO_DEVICE
    lv = np.zeros(G.meta["xsize"])
    ok = ex.cholesky_left_par_05(G.n, G.A2_p, G.A2_i, G.A2_x, G.p, G.s, G.i_ptr, lv, G.super, G.nsuper, None,
                                 G.sParent, None, None, G.col2Sup, len(G.levelPtr) - 1, G.levelPtr, None, 0,
                                 G.parPtr, G.partition)
    assert ok is False and not lv.any()
    x = np.ones(G.n)
    assert ex.blockedLsolve(G.n, G.p, G.s, G.valL, 0, G.i_ptr, G.col2Sup, G.super, G.nsuper, x) == 0
    assert ex.blockedLsolve(G.n, None, G.s, G.valL, 0, G.i_ptr, G.col2Sup, G.super, G.nsuper, x) == 0  # NULL Lp
    # the round-2 entry points fail just as loudly: sharded handle (emulated ranks need a device too), column solves
    with pytest.raises(ex.ParsyCudaError) as e:
        ex.Sharded(G.n, G.A2_p, G.A2_i, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.sParent, G.col2Sup,
                   len(G.levelPtr) - 1, G.levelPtr, G.parPtr, G.partition, 0, 2, None)
    assert e.value.code == ex.ERR_NO_DEVICE
    with pytest.raises(ex.ParsyCudaError) as e:
        ex.CscSolver(G.n, G.Lcsc_p, G.Lcsc_i)
    assert e.value.code == ex.ERR_NO_DEVICE
    y = np.ones(G.n)
    assert ex.lsolve(G.n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x, y) == 0 and np.all(y == 1.0)


def test_sharded_create_rejects_bad_arguments_before_touching_a_device():
    G = load_golden("2d5_N12_c8_l1_d2")
    args = (G.n, G.A2_p, G.A2_i, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.sParent, G.col2Sup, len(G.levelPtr) - 1,
            G.levelPtr, G.parPtr, G.partition)
    for rank, world in ((0, 1), (2, 2), (-1, 4)):
        with pytest.raises(ex.ParsyCudaError) as e:
            ex.Sharded(*args, rank, world, None)
        assert e.value.code == ex.ERR_BAD_ARG
    with pytest.raises(ValueError):
        ex.Sharded(*args, 0, 2, b"too short")


@pytest.mark.parametrize("name", ["2d5_N30_c8_l1_d2", "2d5_N30_c64_l0_d4", "3d7_N7_c8_l1_d2", "3d27_N6_c4_l0_d2"])
def test_sweep_task_lists_are_deadlock_free(name):
    """The sweep kernels spin on counters, so their task lists must be topological orders (producers strictly before
    consumers, leaf region closed under descendants): validated on the host by the planner itself
    (plan.cpp: sweep_order_violations) and reported through parsy_cuda_plan_check."""
    G = load_golden(name)
    for nb in (0, 32):
        rc, st = ex.plan_check(G.n, G.p, G.s, G.i_ptr, G.super, G.nsuper, G.col2Sup, len(G.levelPtr) - 1, G.levelPtr,
                               G.parPtr, G.partition, block_cols=nb)
        assert rc == ex.OK
        ctas, leaf, bad = st["reserved"][5], st["reserved"][6], st["reserved"][7]
        assert bad == 0
        assert 0 < leaf <= ctas


def test_sweep_task_lists_on_a_larger_grid():
    from parsy_bench_b200 import matrices
    n, Ap, Ai, Ax = matrices.laplacian("2d5", 150)
    S = inspector.analyze(n, Ap, Ai, Ax, 64, 1, 4)
    rc, st = ex.plan_check(n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
    assert rc == ex.OK and st["reserved"][7] == 0
    assert st["n_block_cols"] > 0 and 0 < st["reserved"][6] < st["reserved"][5]   # block columns exist: two launches per sweep


def _trailing_flops(S, nb=128):
    """Flops of the updates INSIDE wide supernodes, computed independently of the planner: block column b of width K
    applies (2 M N - N^2) K flops to the N columns right of it over the M rows below it (the count k_gemm_tiles is
    charged with, additive over any split of the target columns)."""
    total = 0.0
    sup, iptr = np.asarray(S.super, np.int64), np.asarray(S.i_ptr, np.int64)
    for s in range(S.nsuper):
        w = int(sup[s + 1] - sup[s])
        r = int(iptr[sup[s + 1]] - iptr[sup[s]])
        if w <= 32 and r <= 1024 and (r - w) * w * w <= 100000:
            continue                      # narrow supernode: factored by one warp, no block columns (plan.h SMALL_*)
        for j0 in range(0, w, nb):
            k = min(nb, w - j0)
            N, M = w - j0 - k, r - j0 - k
            total += (2.0 * M * N - float(N) * N) * k
    return total


@pytest.mark.parametrize("case", [("2d5", 150, 64, 1, 2), ("3d27", 18, 16, 0, 2), ("3d7", 26, 592, 1, 4)])
def test_planner_update_flops_are_conserved(case):
    """Every (supernode, descendant) pair and every panel -> trailing-columns product is planned exactly once, however
    the planner groups them (next-column updates, runs of four block columns applied late, K-splits, sharding by rank
    and phase): the flops of the planned GEMM-shaped tasks equal the pair flops plus the trailing flops recomputed here."""
    from parsy_bench_b200 import matrices
    kind, N, c, l, d = case
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    S = inspector.analyze(n, Ap, Ai, Ax, c, l, d)
    args = (n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
    rc, st = ex.plan_check(*args)
    assert rc == ex.OK
    want = st["flops_update"] + _trailing_flops(S)
    assert max(np.diff(S.super)) > 128                                   # block columns and runs exist
    assert abs(st["reserved"][4] - want) <= 1e-9 * want + 8
    for world in (2, 3, 8):
        tot = 0
        for rank in range(world):
            for phase in (1, 2):
                rc, sp = ex.plan_check(*args, rank=rank, world=world, phase=phase)
                assert rc == ex.OK
                tot += sp["reserved"][4]
        assert abs(tot - want) <= 1e-9 * want + 8 * world * 2


def test_plan_digests_match_the_committed_table():
    """tests/golden/plan_digests.json (tools/plan_digests.py): digests of the plans the GPU parity runs of round 2 were
    made with.  A planner change that alters what the device executes — task order, operands, tile classes, sweep plan,
    ownership — shows up here without a GPU; regenerate the table only together with a GPU parity run."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("plan_digests", os.path.join(root, "tools", "plan_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(root, "tests", "golden", "plan_digests.json")) as f:
        want = json.load(f)
    got = mod.table(mod.SMALL)
    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, bad[:5]


def test_plan_does_not_depend_on_the_number_of_planner_threads(tmp_path):
    """The planner splits its long lists over PARSY_PLAN_THREADS threads by index ranges whose order is fixed by prefix
    sums: the digests of two medium-size problems (several ranges per list) must agree between 1, 3 and 8 threads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tables = []
    for nt in (1, 3, 8):
        out = tmp_path / f"d{nt}.json"
        env = dict(os.environ, PARSY_PLAN_THREADS=str(nt))
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "plan_digests.py"), "--medium", "--json", str(out)],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        with open(out) as f:
            tables.append(json.load(f))
    assert len(tables[0]) >= 20 and not any(v.startswith("error") for v in tables[0].values())
    assert tables[0] == tables[1] == tables[2]



def test_serial_twin_rejects_broken_supernode_partitions_before_touching_them():
    """parsy_cuda_cholesky_left_sn_07 rebuilds col2sup from blockSet on the host (PB_Cholesky.h:16 has no such argument):
    NULL structure arrays and partitions that do not cover 0..n in increasing order are refused, not indexed."""
    L = _lib.lib()
    f = L.parsy_cuda_cholesky_left_sn_07
    f.restype = ctypes.c_int
    n = 4
    I = lambda a: np.asarray(a, np.int32)     # noqa: E731
    U = lambda a: np.asarray(a, np.uint64)    # noqa: E731
    c, r, v = I([0, 1, 2, 3, 4]), I([0, 1, 2, 3]), np.ones(4)
    lC, lR, Lip, lv = U([0, 1, 2, 3, 4]), I([0, 1, 2, 3]), U([0, 1, 2, 3, 4]), np.zeros(4)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 8 + [ctypes.c_int] + [ctypes.c_void_p] * 5
    for bad in (I([0, 1, 2, 3, 5]), I([1, 2, 3, 4, 4]), I([0, 3, 2, 3, 4]), I([0, 1, 1, 3, 4])):
        assert f(n, P(c), P(r), P(v), P(lC), P(lR), P(Lip), P(lv), P(bad), 4, None, None, None, None, None) == 0
        assert "blockSet" in ex.last_error() or "supernode" in ex.last_error()
    good = I([0, 1, 2, 3, 4])
    assert f(n, P(c), P(r), P(v), None, P(lR), P(Lip), P(lv), P(good), 4, None, None, None, None, None) == 0
    assert "NULL" in ex.last_error()


def test_planner_refuses_row_indices_outside_the_matrix():
    from parsy_bench_b200 import matrices
    n, Ap, Ai, Ax = matrices.laplacian("2d5", 12)
    S = inspector.analyze(n, Ap, Ai, Ax)
    for where, val in ((int(S.i_ptr[n]) - 1, n + 7), (int(S.i_ptr[n]) // 2, -1), (0, 2 ** 31 - 1)):
        s2 = np.array(S.s, np.int32, copy=True)
        s2[where] = val
        rc, _ = ex.plan_check(n, S.p, s2, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
        assert rc == ex.ERR_BAD_ARG and "row index" in ex.last_error()


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both public headers (and the forwarding header under C++) must compile on their own."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "h.c"
    src.write_text('#include "parsy_cuda.h"\n#include "parsy_inspector.h"\nint main(void) { return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src),
                        "-o", str(tmp_path / "h.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if shutil.which("g++"):
        src2 = tmp_path / "h.cpp"
        src2.write_text('#include <cstddef>\n#include "parsy_cuda_dropin.h"\nint main() { return 0; }\n')
        r = subprocess.run(["g++", "-std=c++11", "-Wall", "-I", os.path.join(root, "include"), "-c", str(src2), "-o",
                            str(tmp_path / "h2.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
