"""Shared helpers for the test-suite: golden fixtures, symbolic views."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
FULL_CASES = ["2d5_N12_c8_l1_d2", "2d5_N30_c8_l1_d2", "2d5_N30_c64_l0_d4", "3d7_N7_c8_l1_d2", "3d27_N6_c4_l0_d2"]
# a matrix that is not a stencil: committed as a Matrix-Market file, golden arrays from the reference run on that file
MTX_CASES = {"mtx_rand150_c8_l1_d2": "rand150.mtx"}
CPU_CASES = FULL_CASES + sorted(MTX_CASES)
INT_ARRAYS = ["Perm", "ColCount", "super", "sParent", "col2Sup", "pi", "s", "p", "i_ptr", "levelPtr", "parPtr",
              "partition", "A2_p", "A2_i", "A1_p", "A1_i"]


class View(dict):
    """dict with attribute access: quacks like parsy_bench_b200.inspector.Symbolic for the oracle helpers"""
    __getattr__ = dict.__getitem__


def parse_case(name):
    kind, N, c, l, d = name.split("_")
    if kind == "mtx":
        return kind, os.path.join(GOLDEN, MTX_CASES[name]), int(c[1:]), int(l[1:].replace("m", "-")), int(d[1:])
    return kind, int(N[1:]), int(c[1:]), int(l[1:].replace("m", "-")), int(d[1:])


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    v = View({k: z[k] for k in z.files if k != "meta"})
    v["meta"] = json.loads(str(z["meta"]))
    v["nsuper"] = v["meta"]["nsuper"]
    v["n"] = v["meta"]["n"]
    return v


def digests():
    with open(os.path.join(GOLDEN, "digests.json")) as f:
        return json.load(f)


def digests_large():
    """BASELINE.json configs 2-4 at full size (inspector arrays only; tests/golden/make_golden.py --large)."""
    p = os.path.join(GOLDEN, "digests_large.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def as_view(S):
    """inspector.Symbolic -> View (plain arrays)"""
    keys = INT_ARRAYS + ["A2_x", "nsuper", "n", "xsize", "nLevels"]
    return View({k: getattr(S, k) for k in keys})


def rel_err(a, b, floor_scale=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = floor_scale * float(np.max(np.abs(b)))
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return float(np.max(np.abs(a - b) / den))


def full_matrix(S):
    import scipy.sparse as sp
    n = len(S.col2Sup)
    A2 = sp.csc_matrix((S.A2_x, S.A2_i, S.A2_p), shape=(n, n))
    return A2 + sp.tril(A2, -1).T
