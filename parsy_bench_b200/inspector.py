"""ctypes mirror of the host inspector (include/parsy_inspector.h) — the re-statement of the reference's
``analyze_p2`` (cholesky/LSparsity.h:256) plus the two ``ptranspose`` calls of the drivers
(examples/choleskyTest01.cpp:190-191).  Arrays are returned as numpy copies with the reference's names."""
import ctypes
import os
from ctypes import c_int, c_int64, c_double, c_void_p, c_size_t, POINTER, byref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libparsy_inspector.so")
_lib = None


class _Symbolic(ctypes.Structure):
    _fields_ = [("n", c_int), ("nsuper", c_int), ("nnzA", c_int64), ("xsize", c_int64), ("ssize", c_int64),
                ("maxSupWid", c_int), ("maxCol", c_int), ("flops", c_double), ("t_ordering", c_double),
                ("t_total", c_double),
                ("Perm", POINTER(c_int)), ("ColCount", POINTER(c_int)), ("Parent", POINTER(c_int)),
                ("super", POINTER(c_int)), ("sParent", POINTER(c_int)), ("col2Sup", POINTER(c_int)),
                ("pi", POINTER(c_size_t)), ("s", POINTER(c_int)), ("p", POINTER(c_size_t)),
                ("i_ptr", POINTER(c_size_t)), ("nLevels", c_int), ("nParts", c_int), ("levelPtr", POINTER(c_int)),
                ("parPtr", POINTER(c_int)), ("partition", POINTER(c_int)), ("A1_p", POINTER(c_int)),
                ("A1_i", POINTER(c_int)), ("A2_p", POINTER(c_int)), ("A2_i", POINTER(c_int)),
                ("A2_x", POINTER(c_double)), ("A2_src", POINTER(c_int64))]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = ctypes.CDLL(LIB_PATH)
        L.parsy_inspect.restype = c_int
        L.parsy_inspect.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    POINTER(POINTER(_Symbolic))]
        L.parsy_symbolic_free.restype = None
        L.parsy_symbolic_free.argtypes = [POINTER(_Symbolic)]
        L.parsy_inspector_last_error.restype = ctypes.c_char_p
        L.parsy_ereach_sn.restype = c_int
        L.parsy_ereach_sn.argtypes = [POINTER(_Symbolic), c_int, c_void_p]
        L.parsy_etree_level_set.restype = c_int
        L.parsy_etree_level_set.argtypes = [c_int, c_void_p, c_void_p, c_void_p]
        L.parsy_bcsc2csc.restype = c_int64
        L.parsy_bcsc2csc.argtypes = [POINTER(_Symbolic), c_void_p, c_void_p, c_void_p, c_void_p]
        L.parsy_build_level_set_csc.restype = c_int
        L.parsy_build_level_set_csc.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        L.parsy_read_matrix.restype = c_int
        L.parsy_read_matrix.argtypes = [ctypes.c_char_p, POINTER(c_int), POINTER(c_int64), POINTER(POINTER(c_int)),
                                        POINTER(POINTER(c_int)), POINTER(POINTER(c_double))]
        L.parsy_matrix_free.restype = None
        L.parsy_matrix_free.argtypes = [POINTER(c_int), POINTER(c_int), POINTER(c_double)]
        L.parsy_make_lower_half.restype = c_int
        L.parsy_make_lower_half.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_double]
        _lib = L
    return _lib


class Symbolic:
    """Result of :func:`analyze`; attribute names follow the reference (BCSC fields + schedule + A1/A2)."""

    def __init__(self, raw):
        self._raw = raw
        c = raw.contents
        n, ns = c.n, c.nsuper
        self.n, self.nsuper, self.nnzA, self.xsize, self.ssize = n, ns, c.nnzA, c.xsize, c.ssize
        self.maxSupWid, self.maxCol, self.flops = c.maxSupWid, c.maxCol, c.flops
        self.t_ordering, self.t_total = c.t_ordering, c.t_total
        self.nLevels, self.nParts = c.nLevels, c.nParts

        def arr(ptr, cnt, dt):
            return np.ctypeslib.as_array(ptr, shape=(max(int(cnt), 1),))[:int(cnt)].astype(dt, copy=True)

        self.Perm = arr(c.Perm, n, np.int32)
        self.ColCount = arr(c.ColCount, n, np.int32)
        self.Parent = arr(c.Parent, n, np.int32)
        self.super = arr(c.super, ns + 1, np.int32)
        self.sParent = arr(c.sParent, ns, np.int32)
        self.col2Sup = arr(c.col2Sup, n, np.int32)
        self.pi = arr(c.pi, ns + 1, np.uint64)
        self.s = arr(c.s, c.ssize, np.int32)
        self.p = arr(c.p, n + 1, np.uint64)
        self.i_ptr = arr(c.i_ptr, n + 1, np.uint64)
        self.levelPtr = arr(c.levelPtr, c.nLevels + 1, np.int32)
        self.parPtr = arr(c.parPtr, c.nParts + 1, np.int32)
        self.partition = arr(c.partition, ns, np.int32)
        self.A1_p = arr(c.A1_p, n + 1, np.int32)
        self.A1_i = arr(c.A1_i, c.nnzA, np.int32)
        self.A2_p = arr(c.A2_p, n + 1, np.int32)
        self.A2_i = arr(c.A2_i, c.nnzA, np.int32)
        self.A2_x = arr(c.A2_x, c.nnzA, np.float64)
        self.A2_src = arr(c.A2_src, c.nnzA, np.int64)

    def ereach_sn(self, s):
        out = np.empty(self.nsuper, np.int32)
        cnt = lib().parsy_ereach_sn(self._raw, int(s), out.ctypes.data_as(c_void_p))
        if cnt < 0:
            raise ValueError("bad supernode")
        return out[:cnt].copy()

    def etree_level_set(self):
        lp = np.zeros(self.nsuper + 1, np.int32)
        ls = np.zeros(max(self.nsuper, 1), np.int32)
        nl = lib().parsy_etree_level_set(self.nsuper, self.sParent.ctypes.data_as(c_void_p),
                                         lp.ctypes.data_as(c_void_p), ls.ctypes.data_as(c_void_p))
        return nl, lp[:nl + 1].copy(), ls[:self.nsuper].copy()

    def bcsc2csc(self, Lx):
        Lx = np.ascontiguousarray(Lx, np.float64)
        Cp = np.zeros(self.n + 1, np.int32)
        nz = lib().parsy_bcsc2csc(self._raw, None, Cp.ctypes.data_as(c_void_p), None, None)
        Ci = np.empty(nz, np.int32)
        Cx = np.empty(nz, np.float64)
        lib().parsy_bcsc2csc(self._raw, Lx.ctypes.data_as(c_void_p), Cp.ctypes.data_as(c_void_p),
                             Ci.ctypes.data_as(c_void_p), Cx.ctypes.data_as(c_void_p))
        return Cp, Ci, Cx

    def permute_values(self, Ax):
        """Values of tril(P A P') for a new numeric A with the same pattern."""
        return np.ascontiguousarray(Ax, np.float64)[self.A2_src]

    def close(self):
        if self._raw is not None:
            lib().parsy_symbolic_free(self._raw)
            self._raw = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def analyze(n, Ap, Ai, Ax, costParam=8, levelParam=1, divRate=2, perm=None) -> Symbolic:
    """analyze_p2 + ptranspose x2: ordering, symbolic factorization, supernodes, LBC schedule, P A P'."""
    Ap = np.ascontiguousarray(Ap, np.int32)
    Ai = np.ascontiguousarray(Ai, np.int32)
    Ax = None if Ax is None else np.ascontiguousarray(Ax, np.float64)
    pp = None if perm is None else np.ascontiguousarray(perm, np.int32)
    out = POINTER(_Symbolic)()
    rc = lib().parsy_inspect(int(n), Ap.ctypes.data_as(c_void_p), Ai.ctypes.data_as(c_void_p),
                             None if Ax is None else Ax.ctypes.data_as(c_void_p), int(costParam), int(levelParam),
                             int(divRate), None if pp is None else pp.ctypes.data_as(c_void_p), byref(out))
    if rc != 0:
        raise RuntimeError(f"parsy_inspect failed ({rc}): {lib().parsy_inspector_last_error().decode()}")
    return Symbolic(out)


class MatrixMarketError(ValueError):
    def __init__(self, code, msg):
        super().__init__(f"Matrix-Market input rejected ({code}): {msg}")
        self.code = code


def read_matrix(path):
    """``readMatrix`` (common/Util.h:77-179): lower-half, column-ordered coordinate file -> ``(n, col, row, val)``
    0-based CSC, row order inside a column as written.  Raises MatrixMarketError where the reference returns false."""
    L = lib()
    n, nnz = c_int(), c_int64()
    col, row, val = POINTER(c_int)(), POINTER(c_int)(), POINTER(c_double)()
    rc = L.parsy_read_matrix(os.fsencode(path), byref(n), byref(nnz), byref(col), byref(row), byref(val))
    if rc != 0:
        raise MatrixMarketError(rc, L.parsy_inspector_last_error().decode())
    try:
        Ap = np.ctypeslib.as_array(col, shape=(n.value + 1,)).copy()
        Ai = np.ctypeslib.as_array(row, shape=(nnz.value,)).copy()
        Ax = np.ctypeslib.as_array(val, shape=(nnz.value,)).copy()
    finally:
        L.parsy_matrix_free(col, row, val)
    return n.value, Ap, Ai, Ax


def make_lower_half(in_path, out_path, tol=0.1):
    """``printLower`` of examples/MakingLowerHalf.cpp:10-100: full symmetric coordinate file -> lower-half file with
    the diagonal shifted by ``tol`` (the reference's constant is 0.1)."""
    L = lib()
    rc = L.parsy_make_lower_half(os.fsencode(in_path), os.fsencode(out_path), float(tol))
    if rc != 0:
        raise MatrixMarketError(rc, L.parsy_inspector_last_error().decode())


def dag_lbc_csc(n, Lp, Li, innerParts, minLevelDist, divRate, nodeCost=None):
    """``getCoarseLevelSet_DAG_CSC03`` (cholesky/InspectionDAG_03.h:14): LBC schedule over the COLUMNS of a general
    lower-triangular CSC matrix, the input of ``lsolveParH2``.  Returns ``(nLevels, levelPtr, parPtr, partition)``."""
    Lp = np.ascontiguousarray(Lp, np.int32)
    Li = np.ascontiguousarray(Li, np.int32)
    n = int(n)
    lp, pp, part = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32), np.zeros(n, np.int32)
    nl = ctypes.c_int(0)
    cost = None if nodeCost is None else np.ascontiguousarray(nodeCost, np.float64)
    f = lib().parsy_dag_lbc_csc
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p,
                  ctypes.POINTER(ctypes.c_int), c_void_p, c_void_p, c_void_p]
    rc = f(n, Lp.ctypes.data_as(c_void_p), Li.ctypes.data_as(c_void_p), int(innerParts), int(minLevelDist), int(divRate),
           None if cost is None else cost.ctypes.data_as(c_void_p), ctypes.byref(nl), lp.ctypes.data_as(c_void_p),
           pp.ctypes.data_as(c_void_p), part.ctypes.data_as(c_void_p))
    if rc != 0:
        raise ValueError(f"parsy_dag_lbc_csc: {lib().parsy_inspector_last_error().decode()}")
    nl = int(nl.value)
    nparts = int(lp[nl])
    return nl, lp[:nl + 1].copy(), pp[:nparts + 1].copy(), part


def dag_lbc_bcsc(S, innerParts, minLevelDist, divRate, nodeCost="width_x_rows"):
    """``getCoarseLevelSet_DAG_BCSC02`` (cholesky/Inspection_DAG_02.h:15): LBC on the DAG of the blocks of the factor
    described by ``S`` (a Symbolic or anything with super / col2Sup / s / i_ptr / nsuper).  ``nodeCost``: array, None for
    unit costs, or "width_x_rows" (what the reference's analyze_DAG passes).  Returns (nLevels, levelPtr, parPtr, partition)."""
    sup = np.ascontiguousarray(S.super, np.int32)
    c2s = np.ascontiguousarray(S.col2Sup, np.int32)
    lR = np.ascontiguousarray(S.s, np.int32)
    iptr = np.ascontiguousarray(S.i_ptr, np.uint64)
    nb = int(S.nsuper)
    if isinstance(nodeCost, str):
        ip = iptr.astype(np.int64)
        nodeCost = (np.diff(sup.astype(np.int64)) * (ip[sup[1:]] - ip[sup[:-1]])).astype(np.float64)
    cost = None if nodeCost is None else np.ascontiguousarray(nodeCost, np.float64)
    lp, pp, part = np.zeros(nb + 1, np.int32), np.zeros(nb + 1, np.int32), np.zeros(nb, np.int32)
    nl = ctypes.c_int(0)
    f = lib().parsy_dag_lbc_bcsc
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p,
                  ctypes.POINTER(ctypes.c_int), c_void_p, c_void_p, c_void_p]
    rc = f(nb, iptr.ctypes.data_as(c_void_p), lR.ctypes.data_as(c_void_p), sup.ctypes.data_as(c_void_p),
           c2s.ctypes.data_as(c_void_p), int(innerParts), int(minLevelDist), int(divRate),
           None if cost is None else cost.ctypes.data_as(c_void_p), ctypes.byref(nl), lp.ctypes.data_as(c_void_p),
           pp.ctypes.data_as(c_void_p), part.ctypes.data_as(c_void_p))
    if rc != 0:
        raise ValueError(f"parsy_dag_lbc_bcsc: {lib().parsy_inspector_last_error().decode()}")
    nl = int(nl.value)
    return nl, lp[:nl + 1].copy(), pp[:int(lp[nl]) + 1].copy(), part


def build_level_set_csc(n, Lp, Li):
    """``buildLevelSet_CSC`` (triangularSolve/Inspection_Level.h:12): wavefront level sets of a lower-triangular CSC
    matrix (diagonal first per column) for ``lsolvePar``.  Returns ``(levels, levelPtr[levels+1], levelSet[n])``."""
    Lp = np.ascontiguousarray(Lp, np.int32)
    Li = np.ascontiguousarray(Li, np.int32)
    lp = np.zeros(int(n) + 1, np.int32)
    ls = np.zeros(int(n), np.int32)
    levels = lib().parsy_build_level_set_csc(int(n), Lp.ctypes.data_as(c_void_p), Li.ctypes.data_as(c_void_p),
                                             lp.ctypes.data_as(c_void_p), ls.ctypes.data_as(c_void_p))
    if levels < 0:
        raise ValueError(f"parsy_build_level_set_csc: {lib().parsy_inspector_last_error().decode()}")
    return levels, lp[:levels + 1].copy(), ls
