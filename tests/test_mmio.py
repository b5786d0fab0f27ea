"""CPU: Matrix-Market input — parsy_read_matrix against the reference's readMatrix
(common/Util.h:77, through oracle/_ref/parsy_ref --mtx), parsy_make_lower_half against the compiled
examples/MakingLowerHalf.cpp and against committed fixtures, and a non-stencil SPD matrix read from a file through the
whole inspector, bit for bit against the reference's inspector."""
import os
import subprocess

import numpy as np
import pytest

import refdump
from parsy_bench_b200 import inspector, matrices

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MLH_BIN = os.path.join(refdump.ROOT, "oracle", "_ref", "making_lower_half")


def random_spd_lower(n, extra_per_col, seed):
    """Lower half (diagonal first, rows ascending) of a symmetric, strictly diagonally dominant matrix with a random
    sparsity pattern plus a path (so that the graph is connected)."""
    rng = np.random.default_rng(seed)
    cols = [set() for _ in range(n)]
    for j in range(n - 1):
        cols[j].add(j + 1)
        for i in rng.integers(j + 1, n, size=extra_per_col):
            cols[j].add(int(i))
    rowsum = np.zeros(n)
    Ap, Ai, Ax = [0], [], []
    offs = []
    for j in range(n):
        rows = sorted(cols[j])
        vals = -rng.uniform(0.1, 1.0, size=len(rows))
        offs.append((rows, vals))
        for i, v in zip(rows, vals):
            rowsum[i] += abs(v)
            rowsum[j] += abs(v)
    for j in range(n):
        rows, vals = offs[j]
        Ai += [j] + rows
        Ax += [rowsum[j] + 1.0 + 0.01 * j] + list(vals)
        Ap.append(len(Ai))
    return n, np.array(Ap, np.int32), np.array(Ai, np.int32), np.array(Ax)


def test_round_trip_of_a_laplacian(tmp_path):
    n, Ap, Ai, Ax = matrices.laplacian("2d5", 12)
    f = tmp_path / "lap.mtx"
    matrices.write_mtx(f, n, Ap, Ai, Ax, comment="2D 5-point Laplacian 12x12, lower half")
    n2, Bp, Bi, Bx = inspector.read_matrix(f)
    assert n2 == n
    assert np.array_equal(Ap, Bp) and np.array_equal(Ai, Bi) and np.array_equal(Ax, Bx)


def test_upper_case_banner_and_comments(tmp_path):
    f = tmp_path / "m.mtx"
    f.write_text("%%MATRIXMARKET Matrix Coordinate REAL Symmetric\n% a comment\n%another\n2 2 3\n1 1 2.5\n2 1 -1\n2 2 3e0\n")
    n, Ap, Ai, Ax = inspector.read_matrix(f)
    assert n == 2 and Ap.tolist() == [0, 2, 3] and Ai.tolist() == [0, 1, 1] and Ax.tolist() == [2.5, -1.0, 3.0]


@pytest.mark.parametrize("text,code", [
    ("", 1),                                                                        # missing file content
    ("%%MatrixMarket matrix coordinate real\n1 1 1\n1 1 1\n", 1),                  # four tokens
    ("%MatrixMarket matrix coordinate real symmetric\n1 1 1\n1 1 1\n", 2),
    ("%%MatrixMarket vector coordinate real symmetric\n1 1 1\n1 1 1\n", 3),
    ("%%MatrixMarket matrix array real symmetric\n1 1 1\n1 1 1\n", 4),
    ("%%MatrixMarket matrix coordinate complex hermitian\n1 1 1\n1 1 1 0\n", 5),
    ("%%MatrixMarket matrix coordinate pattern symmetric\n1 1 1\n1 1\n", 5),
    ("%%MatrixMarket matrix coordinate integer symmetric\n1 1 1\n1 1 1\n", 5),
    ("%%MatrixMarket matrix coordinate real symmetric\n% only comments\n", 6),
    ("%%MatrixMarket matrix coordinate real symmetric\n0 0 0\n", 7),
    ("%%MatrixMarket matrix coordinate real symmetric\n2 2 3\n1 1 1\n2 1 1\n2 3 1\n", 8),   # column 3 of 2
    ("%%MatrixMarket matrix coordinate real symmetric\n2 2 3\n1 1 1\n3 1 1\n2 2 1\n", 8),   # row 3 of 2
    ("%%MatrixMarket matrix coordinate real symmetric\n3 3 3\n1 1 1\n3 3 1\n2 2 1\n", 9),   # column skipped / unordered
    ("%%MatrixMarket matrix coordinate real symmetric\n3 3 2\n1 1 1\n2 2 1\n", 9),          # last column empty
    ("%%MatrixMarket matrix coordinate real symmetric\n2 2 3\n1 1 1\n2 1 1\n", 10),         # short file
])
def test_rejections(tmp_path, text, code):
    f = tmp_path / "bad.mtx"
    f.write_text(text)
    with pytest.raises(inspector.MatrixMarketError) as e:
        inspector.read_matrix(f)
    assert e.value.code == code


def test_missing_file_is_an_invalid_header(tmp_path):
    with pytest.raises(inspector.MatrixMarketError) as e:
        inspector.read_matrix(tmp_path / "does_not_exist.mtx")
    assert e.value.code == 1    # the reference's getline on a closed stream yields an empty header line


@pytest.mark.skipif(not refdump.have_ref(), reason="compiled reference (oracle/_ref) not built")
def test_reader_and_inspector_match_the_reference_on_a_file(tmp_path):
    """A matrix that is NOT a stencil: read by the reference's readMatrix and by ours, then through both inspectors."""
    n, Ap, Ai, Ax = random_spd_lower(400, 2, seed=7)
    f = tmp_path / "rand.mtx"
    matrices.write_mtx(f, n, Ap, Ai, Ax)
    d = tmp_path / "dump"
    d.mkdir()
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
    subprocess.run([refdump.REF_BIN, "--mtx", str(f), "--cost", "8", "--level", "1", "--div", "2", "--threads", "1",
                    "--dump", str(d), "--no-factor", "--no-solve"], check=True, capture_output=True, env=env)
    ld = lambda name, dt: np.fromfile(d / name, dtype=dt)  # noqa: E731
    n2, Bp, Bi, Bx = inspector.read_matrix(f)
    assert n2 == n
    assert np.array_equal(ld("A_p.i32", np.int32), Bp)
    assert np.array_equal(ld("A_i.i32", np.int32), Bi)
    assert np.array_equal(ld("A_x.f64", np.float64), Bx)
    assert np.array_equal(Bx, Ax)          # 17 digits round-trip
    S = inspector.analyze(n2, Bp, Bi, Bx, 8, 1, 2)
    for name, dt in [("Perm", np.int32), ("ColCount", np.int32), ("super", np.int32), ("sParent", np.int32),
                     ("col2Sup", np.int32), ("s", np.int32), ("p", np.uint64), ("i_ptr", np.uint64),
                     ("levelPtr", np.int32), ("parPtr", np.int32), ("partition", np.int32), ("A2_p", np.int32),
                     ("A2_i", np.int32), ("A1_p", np.int32), ("A1_i", np.int32)]:
        ext = "u64" if dt == np.uint64 else "i32"
        assert np.array_equal(ld(f"{name}.{ext}", dt), getattr(S, name)), name
    assert np.array_equal(ld("A2_x.f64", np.float64), S.A2_x)


def _full_symmetric_file(path):
    """Full storage (both triangles) of a small symmetric matrix with a negative and a zero diagonal entry."""
    rng = np.random.default_rng(3)
    n = 9
    M = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            if rng.random() < 0.4:
                M[i, j] = M[j, i] = round(float(rng.normal()), 9)
    M[np.diag_indices(n)] = [4, -3.25, 0, 1e-7, 123456.789, 2, 7.5, 1, 3]
    ent = [(i + 1, j + 1, M[i, j]) for j in range(n) for i in range(n) if M[i, j] != 0 or i == j]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% full storage\n")
        f.write(f"{n} {n} {len(ent)}\n")
        for i, j, v in ent:
            f.write(f"{i} {j} {v:.12g}\n")


def test_make_lower_half_matches_committed_fixture(tmp_path):
    out = tmp_path / "lower.mtx"
    inspector.make_lower_half(os.path.join(GOLDEN, "mm_full.mtx"), out)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "mm_lower_expected.mtx"), "rb").read()
    # and the result is something read_matrix accepts
    n, Ap, Ai, Ax = inspector.read_matrix(out)
    assert n == 9 and Ap[-1] == len(Ai) and all(Ai[Ap[j]] == j for j in range(n))


@pytest.mark.skipif(not os.path.exists(MLH_BIN), reason="compiled MakingLowerHalf (oracle/_ref) not built")
def test_make_lower_half_matches_the_reference_program(tmp_path):
    src = tmp_path / "full.mtx"
    _full_symmetric_file(src)
    ref = subprocess.run([MLH_BIN, str(src)], check=True, capture_output=True).stdout
    out = tmp_path / "lower.mtx"
    inspector.make_lower_half(src, out)
    assert out.read_bytes() == ref
    # same generator as the committed fixture
    assert src.read_bytes() == open(os.path.join(GOLDEN, "mm_full.mtx"), "rb").read()


def test_make_lower_half_propagates_header_errors(tmp_path):
    src = tmp_path / "bad.mtx"
    src.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(inspector.MatrixMarketError) as e:
        inspector.make_lower_half(src, tmp_path / "o.mtx")
    assert e.value.code == 4


# ---- level sets of a general lower-triangular CSC matrix (the inspector of lsolvePar, non-chordal inputs) --------------
def random_lower_triangular(n, per_col, seed):
    """Random lower-triangular pattern (diagonal first, rows ascending), well conditioned values; not a Cholesky factor:
    its graph has no fill closure (a "non-chordal" input in the reference's terms)."""
    rng = np.random.default_rng(seed)
    Ap, Ai, Ax = [0], [], []
    for j in range(n):
        k = min(n - 1 - j, int(rng.integers(0, per_col + 1)))
        rows = sorted(set(int(i) for i in rng.integers(j + 1, n, size=k))) if k else []
        Ai += [j] + rows
        Ax += [2.0 + rng.random()] + list(rng.uniform(-0.5, 0.5, size=len(rows)))
        Ap.append(len(Ai))
    return n, np.array(Ap, np.int32), np.array(Ai, np.int32), np.array(Ax)


def check_level_sets(n, Lp, Li, levels, lptr, lset):
    assert lptr[0] == 0 and lptr[-1] == n and len(lptr) == levels + 1 and sorted(lset.tolist()) == list(range(n))
    lev = np.empty(n, np.int64)
    for l in range(levels):
        cols = lset[lptr[l]:lptr[l + 1]]
        assert len(cols) > 0 and np.all(np.diff(cols) > 0)        # increasing inside a level, no empty level
        lev[cols] = l
    want = np.zeros(n, np.int64)
    for j in range(n):                                             # level = 1 + latest level it depends on
        rows = Li[Lp[j] + 1:Lp[j + 1]]
        want[rows] = np.maximum(want[rows], want[j] + 1)
    assert np.array_equal(lev, want)


@pytest.mark.parametrize("n,per_col,seed", [(1, 0, 0), (50, 0, 1), (300, 2, 2), (2000, 4, 3), (500, 40, 4)])
def test_level_sets_of_a_triangular_matrix(n, per_col, seed):
    n, Lp, Li, _ = random_lower_triangular(n, per_col, seed)
    levels, lptr, lset = inspector.build_level_set_csc(n, Lp, Li)
    check_level_sets(n, Lp, Li, levels, lptr, lset)


def test_level_sets_reject_bad_input():
    with pytest.raises(ValueError):
        inspector.build_level_set_csc(2, [0, 1, 2], [1, 1])         # column 0 does not start with its diagonal
    with pytest.raises(ValueError):
        inspector.build_level_set_csc(2, [0, 1, 3], [0, 1, 0])      # entry above the diagonal


@pytest.mark.skipif(not refdump.have_ref(), reason="compiled reference (oracle/_ref) not built")
@pytest.mark.parametrize("n,per_col,seed", [(400, 3, 11), (1500, 6, 12), (64, 20, 13)])
def test_level_sets_match_the_reference(tmp_path, n, per_col, seed):
    """buildLevelSet_CSC of the reference (through oracle/_ref/parsy_ref --mtx ... --tri-only) on a triangular matrix
    read from a Matrix-Market file: same levelPtr and levelSet, bit for bit."""
    n, Lp, Li, Lx = random_lower_triangular(n, per_col, seed)
    f = tmp_path / "tri.mtx"
    matrices.write_mtx(f, n, Lp, Li, Lx, symmetric=False)
    d = tmp_path / "dump"
    d.mkdir()
    subprocess.run([refdump.REF_BIN, "--mtx", str(f), "--tri-only", "--dump", str(d)], check=True, capture_output=True)
    levels, lptr, lset = inspector.build_level_set_csc(n, Lp, Li)
    assert np.array_equal(np.fromfile(d / "tri_levelPtr.i32", np.int32), lptr)
    assert np.array_equal(np.fromfile(d / "tri_levelSet.i32", np.int32), lset)


# ---- LBC on the DAG of a general lower-triangular matrix (the inspector of lsolveParH2, InspectionDAG_03.h) ---------
def check_h2_schedule(n, Lp, Li, nl, lp, pp, part):
    """every column once; a column's producers sit in an earlier H-level or earlier in the same w-partition"""
    assert lp[0] == 0 and len(lp) == nl + 1 and pp[0] == 0 and len(pp) == lp[-1] + 1 and pp[-1] == n
    assert sorted(part.tolist()) == list(range(n))
    pos, lev, par = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64)
    for H in range(nl):
        for j1 in range(lp[H], lp[H + 1]):
            for k in range(pp[j1], pp[j1 + 1]):
                c = part[k]
                pos[c], lev[c], par[c] = k, H, j1
    for j in range(n):
        for i in Li[Lp[j] + 1:Lp[j + 1]]:
            assert lev[j] < lev[i] or (par[j] == par[i] and pos[j] < pos[i])


@pytest.mark.parametrize("n,per_col,seed,params", [(1, 0, 0, (4, 2, 2)), (50, 0, 1, (4, 2, 2)), (300, 2, 2, (8, 2, 2)),
                                                    (2000, 4, 3, (16, 1, 3)), (500, 40, 4, (3, 5, 2))])
def test_dag_lbc_schedule_is_legal(n, per_col, seed, params):
    n, Lp, Li, _ = random_lower_triangular(n, per_col, seed)
    nl, lp, pp, part = inspector.dag_lbc_csc(n, Lp, Li, *params)
    check_h2_schedule(n, Lp, Li, nl, lp, pp, part)


def test_dag_lbc_rejects_bad_input():
    with pytest.raises(ValueError):
        inspector.dag_lbc_csc(2, [0, 1, 2], [1, 1], 4, 2, 2)           # column 0 does not start with its diagonal
    with pytest.raises(ValueError):
        inspector.dag_lbc_csc(2, [0, 1, 2], [0, 1], 0, 2, 2)           # no bins
    with pytest.raises(ValueError):
        inspector.dag_lbc_csc(2, [0, 1, 2], [0, 1], 4, 2, 0)           # the level cut would never advance


def _ref_dag(tmp_path, n, Lp, Li, Lx, cost, level, div):
    f = tmp_path / "tri.mtx"
    matrices.write_mtx(f, n, Lp, Li, Lx, symmetric=False)
    d = tmp_path / "dump"
    d.mkdir(exist_ok=True)
    subprocess.run([refdump.REF_BIN, "--mtx", str(f), "--tri-only", "--dump", str(d), "--cost", str(cost), "--level",
                    str(level), "--div", str(div)], check=True, capture_output=True)
    return {k: np.fromfile(d / f"dag_{k}.i32", np.int32) for k in ("levelPtr", "parPtr", "partition")}, d


@pytest.mark.skipif(not refdump.have_ref(), reason="compiled reference (oracle/_ref) not built")
@pytest.mark.parametrize("n,per_col,seed,params", [(400, 3, 11, (8, 2, 2)), (1500, 6, 12, (16, 1, 2)), (64, 20, 13, (4, 2, 3)),
                                                    (3000, 2, 5, (8, 3, 4)), (800, 10, 3, (2, 1, 1)), (6000, 2, 8, (148, 2, 2)),
                                                    (300, 2, 2, (8, 40, 2)), (2500, 4, 9, (64, 2, 3))])
def test_dag_lbc_matches_the_reference(tmp_path, n, per_col, seed, params):
    """getCoarseLevelSet_DAG_CSC03 of the reference (cholesky/InspectionDAG_03.h:14, run through oracle/_ref/parsy_ref
    --tri-only with unit node costs as examples/triangularTest_DAG_nonChordal.cpp:343-360 does) on random triangular
    matrices that are not Cholesky factors: same levelPtr, parPtr and partition, bit for bit."""
    n, Lp, Li, Lx = random_lower_triangular(n, per_col, seed)
    R, _ = _ref_dag(tmp_path, n, Lp, Li, Lx, *params)
    nl, lp, pp, part = inspector.dag_lbc_csc(n, Lp, Li, *params)
    assert np.array_equal(lp, R["levelPtr"]) and np.array_equal(pp, R["parPtr"]) and np.array_equal(part, R["partition"])


@pytest.mark.skipif(not refdump.have_ref(), reason="compiled reference (oracle/_ref) not built")
def test_dag_lbc_matches_the_reference_on_a_cholesky_factor(tmp_path):
    """... and on the column form of a supernodal Cholesky factor (dense chains inside the supernodes, components that
    meet and are merged): the golden 2D 30x30 case."""
    from common import load_golden
    G = load_golden("2d5_N30_c8_l1_d2")
    n, Lp, Li, Lx = G.n, G.Lcsc_p, G.Lcsc_i, G.Lcsc_x
    for params in ((8, 2, 2), (4, 3, 5), (16, 1, 7)):
        R, _ = _ref_dag(tmp_path, n, Lp, Li, Lx, *params)
        nl, lp, pp, part = inspector.dag_lbc_csc(n, Lp, Li, *params)
        assert np.array_equal(lp, R["levelPtr"]) and np.array_equal(pp, R["parPtr"]) and np.array_equal(part, R["partition"])
        check_h2_schedule(n, Lp, Li, nl, lp, pp, part)
