"""CPU: the from-scratch host inspector must reproduce the reference's analyze_p2 / ptranspose output bit for bit."""
import hashlib
import os
import sys

import numpy as np
import pytest

from common import CPU_CASES, FULL_CASES, INT_ARRAYS, load_golden, digests, digests_large, parse_case, View
from refdump import have_ref, ref_case
from parsy_bench_b200 import inspector, matrices

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import parsy_oracle as orc  # noqa: E402


def run_inspector(name):
    kind, N, c, l, d = parse_case(name)
    if kind == "mtx":      # committed Matrix-Market file (N is its path): restated readMatrix, then the inspector
        n, Ap, Ai, Ax = inspector.read_matrix(N)
    else:
        n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    return inspector.analyze(n, Ap, Ai, Ax, c, l, d)


@pytest.mark.parametrize("name", CPU_CASES)
def test_bit_exact_vs_golden(name):
    G = load_golden(name)
    S = run_inspector(name)
    for k in INT_ARRAYS:
        assert np.array_equal(getattr(S, k), G[k]), k
    assert np.array_equal(S.A2_x, G.A2_x)
    assert (S.nsuper, S.xsize, S.ssize) == (G.meta["nsuper"], G.meta["xsize"], G.meta["ssize"])
    assert (S.maxSupWid, S.maxCol) == (G.meta["maxSupWid"], G.meta["maxCol"])
    assert S.flops == G.meta["flops"]
    nl, lp, ls = S.etree_level_set()
    assert nl == G.meta["etree_levels"]
    assert np.array_equal(lp, G.etree_levelPtr) and np.array_equal(ls, G.etree_levelSet)


@pytest.mark.parametrize("name", sorted(digests().keys()))
def test_bit_exact_vs_digest(name):
    D = digests()[name]
    S = run_inspector(name)
    for k in INT_ARRAYS:
        h = hashlib.sha256(np.ascontiguousarray(getattr(S, k)).tobytes()).hexdigest()
        assert h == D["sha256"][k], k
    assert S.flops == D["flops"] and S.xsize == D["xsize"] and S.nLevels == D["nLevels"] and S.nParts == D["nParts"]


@pytest.mark.parametrize("name", sorted(digests_large().keys()))
def test_bit_exact_vs_digest_full_size_configs(name):
    """BASELINE.json configs 2, 3, 4 at full size (n = 1e6 / 1e6 / 262144): every integer array of the inspector against
    the SHA-256 the compiled reference produced (analyze_p2, cholesky/LSparsity.h:256) — the sizes at which the
    int/double zero counts of the relaxed amalgamation (Inspection_BlockC.h:430-469) and the int cost accumulator
    (InspectionLevel_06.h:197) leave the small-case regime."""
    D = digests_large()[name]
    S = run_inspector(name)
    assert (S.n, S.nsuper, S.xsize, S.ssize) == (D["n"], D["nsuper"], D["xsize"], D["ssize"])
    assert (S.nLevels, S.nParts, S.maxSupWid, S.maxCol) == (D["nLevels"], D["nParts"], D["maxSupWid"], D["maxCol"])
    assert S.flops == D["flops"]
    for k in INT_ARRAYS:
        h = hashlib.sha256(np.ascontiguousarray(getattr(S, k)).tobytes()).hexdigest()
        assert h == D["sha256"][k], k


@pytest.mark.parametrize("name", FULL_CASES[:2] + FULL_CASES[3:] + CPU_CASES[len(FULL_CASES):])
def test_ereach_sn_order(name):
    """parsy_ereach_sn returns ereach_sn's stack (common/Reach.h:112) in the reference's order."""
    S = run_inspector(name)
    for s in range(S.nsuper):
        assert np.array_equal(S.ereach_sn(s), orc.ereach_sn(S, s))


def test_bcsc2csc_matches_reference():
    G = load_golden("3d7_N7_c8_l1_d2")
    S = run_inspector("3d7_N7_c8_l1_d2")
    Cp, Ci, Cx = S.bcsc2csc(G.valL)
    assert np.array_equal(Cp, G.Lcsc_p) and np.array_equal(Ci, G.Lcsc_i) and np.array_equal(Cx, G.Lcsc_x)


def test_schedule_is_legal_and_complete():
    S = run_inspector("2d5_N100_c592_l1_d4")
    assert sorted(S.partition.tolist()) == list(range(S.nsuper))
    pos = np.empty(S.nsuper, np.int64)
    lvl = np.empty(S.nsuper, np.int64)
    part = np.empty(S.nsuper, np.int64)
    for H in range(S.nLevels):
        for j in range(S.levelPtr[H], S.levelPtr[H + 1]):
            for k in range(S.parPtr[j], S.parPtr[j + 1]):
                s = S.partition[k]
                pos[s], lvl[s], part[s] = k, H, j
    for s in range(S.nsuper):
        p = S.sParent[s]
        if p >= 0:
            assert lvl[s] < lvl[p] or (part[s] == part[p] and pos[s] < pos[p])


def test_user_permutation_and_errors():
    n, Ap, Ai, Ax = matrices.laplacian("2d5", 8)
    S = inspector.analyze(n, Ap, Ai, Ax, perm=np.arange(n)[::-1].copy())
    assert sorted(S.Perm.tolist()) == list(range(n))
    with pytest.raises(RuntimeError):
        inspector.analyze(n, Ap, Ai, Ax, perm=np.zeros(n, np.int32))
    with pytest.raises(RuntimeError):
        inspector.analyze(n, Ap, Ai, Ax, divRate=1)


def test_permute_values_roundtrip():
    n, Ap, Ai, Ax = matrices.laplacian("3d7", 5)
    S = inspector.analyze(n, Ap, Ai, Ax)
    new = Ax * 2.5
    assert np.array_equal(S.permute_values(new), S.A2_x * 2.5)


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", [("2d5", 64, 16, 2, 2), ("2d5", 150, 592, 1, 4), ("3d7", 18, 8, -1, 4),
                                  ("3d27", 14, 37, 0, 3), ("2d5", 200, 8, -2, 2), ("3d27", 20, 148, 1, 2)])
def test_bit_exact_vs_live_reference(case):
    kind, N, c, l, d = case
    R = ref_case(kind, N, cost=c, level=l, div=d, factor=False, solve=False)
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    S = inspector.analyze(n, Ap, Ai, Ax, c, l, d)
    for k in INT_ARRAYS:
        ref = R[k][:S.ssize] if k == "s" else R[k]
        assert np.array_equal(getattr(S, k), ref), k


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", [("2d5", 30, 8, 1, 2), ("2d5", 64, 16, 2, 2), ("3d7", 12, 8, 1, 2), ("3d27", 10, 4, 0, 2),
                                  ("2d5", 150, 592, 1, 4), ("3d7", 20, 37, 2, 3), ("3d27", 16, 8, -1, 4)])
def test_dag_lbc_over_blocks_matches_the_reference(case):
    """getCoarseLevelSet_DAG_BCSC02 (cholesky/Inspection_DAG_02.h:15, called as analyze_DAG does, LSparsity.h:1412, with
    width x rows as node cost): the DAG-based LBC schedule over the factor's blocks, bit for bit; and the executor's
    planner accepts it (every descendant in an earlier H-level or earlier in the same w-partition)."""
    from parsy_bench_b200 import executor as ex
    kind, N, c, l, d = case
    R = ref_case(kind, N, cost=c, level=l, div=d, factor=False, solve=False)
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    S = inspector.analyze(n, Ap, Ai, Ax, c, l, d)
    nl, lp, pp, part = inspector.dag_lbc_bcsc(S, c, l, d)
    assert np.array_equal(lp, R["dagb_levelPtr"]) and np.array_equal(pp, R["dagb_parPtr"])
    assert np.array_equal(part, R["dagb_partition"])
    assert sorted(part.tolist()) == list(range(S.nsuper))
    rc, _ = ex.plan_check(n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, nl, lp, pp, part)
    assert rc == ex.OK
