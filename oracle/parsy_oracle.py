"""ctypes access to oracle/libparsy_oracle.so (plain-C restatement of the reference's hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
the product package.  `S` arguments are inspector results (parsy_bench_b200.inspector.Symbolic) or any object with
the same attribute names (e.g. a reference dump)."""
import ctypes
import os
from ctypes import c_int, c_void_p, c_size_t

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libparsy_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, dt):
    b = np.ascontiguousarray(a, dtype=dt)
    return b, b.ctypes.data_as(c_void_p)


def ereach_sn(S, s):
    """common/Reach.h:112 for supernode s (0-based): descendant list in the reference's order."""
    L = lib()
    ns = int(S.nsuper)
    keep = [_p(S.A1_p, np.int32), _p(S.A1_i, np.int32), _p(S.col2Sup, np.int32), _p(S.sParent, np.int32)]
    xi = np.zeros(2 * ns + 2, np.int32)
    f = L.oracle_ereach_sn
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    top = f(ns, keep[0][1], keep[1][1], int(S.super[s]), int(S.super[s + 1]), keep[2][1], keep[3][1],
            xi.ctypes.data_as(c_void_p), xi[ns:].ctypes.data_as(c_void_p))
    return xi[top:ns].copy()


def _factor_args(S, values):
    keep = [_p(S.A2_p, np.int32), _p(S.A2_i, np.int32), _p(S.A2_x if values is None else values, np.float64),
            _p(S.p, np.uint64), _p(S.s, np.int32), _p(S.i_ptr, np.uint64), _p(S.super, np.int32),
            _p(S.sParent, np.int32), _p(S.A1_p, np.int32), _p(S.A1_i, np.int32), _p(S.col2Sup, np.int32)]
    return keep


def _dims(S):
    sup = np.asarray(S.super)
    iptr = np.asarray(S.i_ptr)
    wid = int(np.max(np.diff(sup)))
    rows = int(np.max(iptr[sup[1:]] - iptr[sup[:-1]]))
    return wid + 1, rows + 1


def cholesky_left_par_05(S, values=None):
    """parallel_PB_Cholesky_05.h:27 on one thread in schedule order; returns lValues or None if not SPD."""
    L = lib()
    k = _factor_args(S, values)
    sch = [_p(S.levelPtr, np.int32), _p(S.parPtr, np.int32), _p(S.partition, np.int32)]
    n = len(S.col2Sup)
    xsize = int(np.asarray(S.p)[n])
    lv = np.zeros(xsize)
    sm, cm = _dims(S)
    f = L.oracle_cholesky_left_par_05
    f.restype = c_int
    f.argtypes = [c_int] + [c_void_p] * 8 + [c_int] + [c_void_p] * 4 + [c_int] + [c_void_p] * 3 + [c_int, c_int]
    ok = f(n, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], k[5][1], lv.ctypes.data_as(c_void_p), k[6][1],
           int(S.nsuper), k[7][1], k[8][1], k[9][1], k[10][1], len(S.levelPtr) - 1, sch[0][1], sch[1][1], sch[2][1],
           sm, cm)
    return lv if ok else None


def cholesky_left_sn(S, values=None):
    """PB_Cholesky.h:16 serial order 0..supNo-1; returns lValues or None if not SPD."""
    L = lib()
    k = _factor_args(S, values)
    n = len(S.col2Sup)
    xsize = int(np.asarray(S.p)[n])
    lv = np.zeros(xsize)
    sm, cm = _dims(S)
    f = L.oracle_cholesky_left_sn
    f.restype = c_int
    f.argtypes = [c_int] + [c_void_p] * 8 + [c_int] + [c_void_p] * 4 + [c_int, c_int]
    ok = f(n, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], k[5][1], lv.ctypes.data_as(c_void_p), k[6][1],
           int(S.nsuper), k[7][1], k[8][1], k[9][1], k[10][1], sm, cm)
    return lv if ok else None


def _solve(name, S, Lx, x, schedule=False):
    L = lib()
    n = len(S.col2Sup)
    keep = [_p(S.p, np.uint64), _p(S.s, np.int32), _p(Lx, np.float64), _p(S.i_ptr, np.uint64), _p(S.super, np.int32)]
    y = np.array(x, dtype=np.float64, copy=True)
    f = getattr(L, name)
    f.restype = c_int
    if schedule:
        sch = [_p(S.levelPtr, np.int32), _p(S.parPtr, np.int32), _p(S.partition, np.int32)]
        f.argtypes = [c_int] + [c_void_p] * 5 + [c_int, c_void_p, c_int] + [c_void_p] * 3
        rc = f(n, keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1], int(S.nsuper),
               y.ctypes.data_as(c_void_p), len(S.levelPtr) - 1, sch[0][1], sch[1][1], sch[2][1])
    else:
        f.argtypes = [c_int] + [c_void_p] * 5 + [c_int, c_void_p]
        rc = f(n, keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1], int(S.nsuper),
               y.ctypes.data_as(c_void_p))
    assert rc == 1
    return y


def blockedLsolve(S, Lx, x):
    return _solve("oracle_blockedLsolve", S, Lx, x)


def H2LeveledBlockedLsolve(S, Lx, x):
    return _solve("oracle_H2LeveledBlockedLsolve", S, Lx, x, schedule=True)


def blockedLtsolve(S, Lx, x):
    return _solve("oracle_blockedLtsolve", S, Lx, x)


def lsolve(n, Lp, Li, Lx, x):
    L = lib()
    k = [_p(Lp, np.int32), _p(Li, np.int32), _p(Lx, np.float64)]
    y = np.array(x, dtype=np.float64, copy=True)
    f = L.oracle_lsolve
    f.restype = c_int
    f.argtypes = [c_int] + [c_void_p] * 4
    assert f(int(n), k[0][1], k[1][1], k[2][1], y.ctypes.data_as(c_void_p)) == 1
    return y


def rhs_init_blocked(S, Lx):
    L = lib()
    n = len(S.col2Sup)
    k = [_p(S.p, np.uint64), _p(S.s, np.int32), _p(S.i_ptr, np.uint64), _p(Lx, np.float64)]
    b = np.zeros(n)
    f = L.oracle_rhsInitBlocked
    f.restype = None
    f.argtypes = [c_size_t] + [c_void_p] * 5
    f(n, k[0][1], k[1][1], k[2][1], k[3][1], b.ctypes.data_as(c_void_p))
    return b


def test_triangular(x):
    L = lib()
    a, p = _p(x, np.float64)
    f = L.oracle_testTriangular
    f.restype = c_int
    f.argtypes = [c_size_t, c_void_p]
    return bool(f(a.size, p))


def residual_sym_lower(S, x, b, values=None):
    """b - (P A P') x with the symmetric matrix given by its lower half (S.A2_p, S.A2_i, values)."""
    L = lib()
    n = len(S.col2Sup)
    k = [_p(S.A2_p, np.int32), _p(S.A2_i, np.int32), _p(S.A2_x if values is None else values, np.float64),
         _p(x, np.float64), _p(b, np.float64)]
    res = np.zeros(n)
    f = L.oracle_residual_sym_lower
    f.restype = None
    f.argtypes = [c_int] + [c_void_p] * 6
    f(n, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], res.ctypes.data_as(c_void_p))
    return res


def solve_system(S, Lx, b, refine_steps=0, values=None):
    """x = P' (L L')^{-1} P b with `refine_steps` rounds of iterative refinement; restated sweeps only
    (blockedLsolve Triangular_BCSC.h:14 + the reverse sweep).  Returns (x, rel) with rel[k] = ||y - (PAP')w|| / ||y||
    before refinement step k (last entry: final) — the oracle for parsy_cuda_solve_system."""
    perm = np.asarray(S.Perm)
    y = np.asarray(b, np.float64)[perm]
    w = blockedLtsolve(S, Lx, blockedLsolve(S, Lx, y))
    rel = []
    for it in range(refine_steps + 1):
        r = residual_sym_lower(S, w, y, values)
        rel.append(float(np.linalg.norm(r) / np.linalg.norm(y)))
        if it == refine_steps:
            break
        w = w + blockedLtsolve(S, Lx, blockedLsolve(S, Lx, r))
    x = np.empty_like(w)
    x[perm] = w
    return x, np.array(rel)
