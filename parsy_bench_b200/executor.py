"""Host-side mirror of the reference's executor interface over libparsy_cuda (ctypes).

The free functions keep the reference's names, argument order and return conventions
(``cholesky_left_par_05`` cholesky/parallel_PB_Cholesky_05.h:27-39; ``blockedLsolve`` …
triangularSolve/Triangular_BCSC.h:14,115,171,238; ``lsolve`` … triangularSolve/Triangular_CSC.h:14,50,76)
and take numpy arrays where the reference takes raw pointers.  :class:`Solver` wraps the resident handle API.
Nothing here computes on the CPU: every call goes through the C ABI into CUDA kernels.
"""
import ctypes
from ctypes import c_int, c_void_p, c_double, c_size_t, POINTER, byref

import numpy as np

from ._lib import lib, last_error, Options, Stats

SOLVE_FWD, SOLVE_BWD = 1, 2
OK, ERR_NOT_SPD, ERR_BAD_ARG, ERR_BAD_SCHEDULE, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_NO_MEMORY = range(8)


class ParsyCudaError(RuntimeError):
    def __init__(self, code, where):
        super().__init__(f"{where}: error {code}: {last_error()}")
        self.code = code


def _arr(a, dtype, name, allow_none=False):
    if a is None:
        if allow_none:
            return None, None
        raise ValueError(f"{name} is None")
    b = np.ascontiguousarray(a, dtype=dtype)
    return b, b.ctypes.data_as(c_void_p)


def _i32(a, name, allow_none=False):
    return _arr(a, np.int32, name, allow_none)


def _u64(a, name, allow_none=False):
    return _arr(a, np.uint64, name, allow_none)


def _f64_inplace(a, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable):
        raise ValueError(f"{name} must be a writable contiguous float64 array (it is updated in place)")
    return a.ctypes.data_as(c_void_p)


def device_count() -> int:
    return int(lib().parsy_cuda_device_count())


# ---------------------------------------------------------------------------------------------------------
# drop-in free functions (host arrays in, host arrays out)
# ---------------------------------------------------------------------------------------------------------
def cholesky_left_par_05(n, c, r, values, lC, lR, Li_ptr, lValues, blockSet, supNo, timing, aTree, cT, rT, col2Sup,
                         nLevels, levelPtr, levelSet, nPar, parPtr, partition, chunk=1, threads=1, super_max=0,
                         col_max=0, nodCost=None) -> bool:
    """LBC-scheduled supernodal Cholesky; fills ``lValues`` (xsize float64) in place. Returns True/False like the
    reference (False iff a diagonal block is not positive definite, parallel_PB_Cholesky_05.h:206-207)."""
    L = lib()
    keep = []

    def P(x):
        keep.append(x[0])
        return x[1]

    f = L.parsy_cuda_cholesky_left_par_05
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                  c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    tptr = None if timing is None else _f64_inplace(timing, "timing")
    rc = f(int(n), P(_i32(c, "c")), P(_i32(r, "r")), P(_arr(values, np.float64, "values")), P(_u64(lC, "lC")),
           P(_i32(lR, "lR")), P(_u64(Li_ptr, "Li_ptr")), _f64_inplace(lValues, "lValues"), P(_i32(blockSet, "blockSet")),
           int(supNo), tptr, P(_i32(aTree, "aTree", True)), P(_i32(cT, "cT", True)), P(_i32(rT, "rT", True)),
           P(_i32(col2Sup, "col2Sup")), int(nLevels), P(_i32(levelPtr, "levelPtr")), None, int(nPar),
           P(_i32(parPtr, "parPtr")), P(_i32(partition, "partition")), int(chunk), int(threads), int(super_max),
           int(col_max), None)
    return bool(rc)


def cholesky_left_sn_07(n, c, r, values, lC, lR, Li_ptr, lValues, blockSet, supNo, timing, prunePtr, pruneSet,
                        map=None, contribs=None) -> bool:  # noqa: A002 - reference argument name
    """Serial twin driven by a prune set (cholesky/PB_Cholesky.h:16-19)."""
    L = lib()
    keep = []

    def P(x):
        keep.append(x[0])
        return x[1]

    f = L.parsy_cuda_cholesky_left_sn_07
    f.restype = c_int
    f.argtypes = [c_int] + [c_void_p] * 8 + [c_int] + [c_void_p] * 5
    tptr = None if timing is None else _f64_inplace(timing, "timing")
    rc = f(int(n), P(_i32(c, "c")), P(_i32(r, "r")), P(_arr(values, np.float64, "values")), P(_u64(lC, "lC")),
           P(_i32(lR, "lR")), P(_u64(Li_ptr, "Li_ptr")), _f64_inplace(lValues, "lValues"), P(_i32(blockSet, "blockSet")),
           int(supNo), tptr, P(_i32(prunePtr, "prunePtr", True)), P(_i32(pruneSet, "pruneSet", True)), None, None)
    return bool(rc)


def _solve_common(fname, n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x, extra_types=(), extra=()):
    L = lib()
    if Lp is None or Li is None or x is None:
        return 0  # Triangular_BCSC.h:24
    keep = []

    def P(v):
        keep.append(v[0])
        return v[1]

    f = getattr(L, fname)
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p] + \
        list(extra_types)
    args = [int(n), P(_u64(Lp, "Lp")), P(_i32(Li, "Li")), P(_arr(Lx, np.float64, "Lx")), int(NNZ) & 0x7FFFFFFF,
            P(_u64(Li_ptr, "Li_ptr")), P(_i32(col2sup, "col2sup")), P(_i32(sup2col, "sup2col")), int(supNo),
            _f64_inplace(x, "x")]
    for v in extra:
        if isinstance(v, tuple):
            args.append(P(v))
        else:
            args.append(v)
    return int(f(*args))


def blockedLsolve(n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x) -> int:
    """Supernodal forward solve L x = b in place on ``x`` (Triangular_BCSC.h:14)."""
    return _solve_common("parsy_cuda_blockedLsolve", n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x)


def blockedLtsolve(n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x) -> int:
    """NEW: supernodal backward solve L' x = b in place (the reference has none)."""
    return _solve_common("parsy_cuda_blockedLtsolve", n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x)


def leveledBlockedLsolve(n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x, levels, levelPtr, levelSet,
                         chunk=1) -> int:
    """Level-set supernodal forward solve (Triangular_BCSC.h:115)."""
    return _solve_common("parsy_cuda_leveledBlockedLsolve", n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x,
                         [c_int, c_void_p, c_void_p, c_int],
                         [int(levels), _i32(levelPtr, "levelPtr"), _i32(levelSet, "levelSet"), int(chunk)])


def H2LeveledBlockedLsolve(n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x, levels, levelPtr, levelSet, parts,
                           parPtr, partition, chunk=1) -> int:
    """LBC-scheduled supernodal forward solve (Triangular_BCSC.h:171)."""
    return _solve_common("parsy_cuda_H2LeveledBlockedLsolve", n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x,
                         [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int],
                         [int(levels), _i32(levelPtr, "levelPtr"), None, int(parts), _i32(parPtr, "parPtr"),
                          _i32(partition, "partition"), int(chunk)])


def H2LeveledBlockedLsolve_Peeled(n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col, supNo, x, levels, levelPtr, levelSet,
                                  parts, parPtr, partition, chunk=1, threads=1) -> int:
    """LBC-scheduled forward solve with the last level peeled (Triangular_BCSC.h:238)."""
    return _solve_common("parsy_cuda_H2LeveledBlockedLsolve_Peeled", n, Lp, Li, Lx, NNZ, Li_ptr, col2sup, sup2col,
                         supNo, x, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int],
                         [int(levels), _i32(levelPtr, "levelPtr"), None, int(parts), _i32(parPtr, "parPtr"),
                          _i32(partition, "partition"), int(chunk), int(threads)])


def _csc_common(fname, n, Lp, Li, Lx, x, extra_types=(), extra=()):
    L = lib()
    if Lp is None or Li is None or x is None:
        return 0  # Triangular_CSC.h:16
    keep = []

    def P(v):
        keep.append(v[0])
        return v[1]

    f = getattr(L, fname)
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p] + list(extra_types)
    args = [int(n), P(_i32(Lp, "Lp")), P(_i32(Li, "Li")), P(_arr(Lx, np.float64, "Lx")), _f64_inplace(x, "x")]
    for v in extra:
        args.append(P(v) if isinstance(v, tuple) else v)
    return int(f(*args))


def lsolve(n, Lp, Li, Lx, x) -> int:
    """Column forward solve on a CSC lower-triangular matrix, diagonal first (Triangular_CSC.h:14)."""
    return _csc_common("parsy_cuda_lsolve", n, Lp, Li, Lx, x)


def lsolvePar(n, Lp, Li, Lx, x, levels, levelPtr, levelSet, chunk=1) -> int:
    """Level-set column forward solve (Triangular_CSC.h:50)."""
    return _csc_common("parsy_cuda_lsolvePar", n, Lp, Li, Lx, x, [c_int, c_void_p, c_void_p, c_int],
                       [int(levels), _i32(levelPtr, "levelPtr"), _i32(levelSet, "levelSet"), int(chunk)])


def lsolveParH2(n, Lp, Li, Lx, x, levels, levelPtr, levelSet, parts, parPtr, partition, chunk=1) -> int:
    """LBC column forward solve (Triangular_CSC.h:76)."""
    return _csc_common("parsy_cuda_lsolveParH2", n, Lp, Li, Lx, x,
                       [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int],
                       [int(levels), _i32(levelPtr, "levelPtr", True), None, int(parts), _i32(parPtr, "parPtr", True),
                        _i32(partition, "partition", True), int(chunk)])


class CscSolver:
    """Resident column (CSC) forward solve: structure and column order once, values / right-hand sides per call
    (parsy_cuda_csc_*).  ``order``: the schedule flattened to a column order, None = 0..n-1."""

    def __init__(self, n, Lp, Li, order=None, device=0):
        self._L = lib()
        self._h = c_void_p()
        keep = [_i32(Lp, "Lp"), _i32(Li, "Li"), _i32(order, "order", True)]
        f = self._L.parsy_cuda_csc_create
        f.restype = c_int
        f.argtypes = [POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p, c_int]
        rc = f(byref(self._h), int(n), keep[0][1], keep[1][1], keep[2][1], int(device))
        if rc != OK:
            self._h = c_void_p()
            raise ParsyCudaError(rc, "parsy_cuda_csc_create")
        self.n, self.nnz = int(n), int(np.asarray(Lp)[n])

    def set_values(self, Lx):
        v = np.ascontiguousarray(Lx, dtype=np.float64)
        if v.size != self.nnz:
            raise ValueError("Lx has the wrong length")
        self._v = v          # the copy is asynchronous: keep the buffer alive until the next solve has synchronised
        f = self._L.parsy_cuda_csc_set_values
        f.restype = c_int
        f.argtypes = [c_void_p, c_void_p]
        rc = f(self._h, v.ctypes.data_as(c_void_p))
        if rc != OK:
            raise ParsyCudaError(rc, "parsy_cuda_csc_set_values")

    def solve(self, x) -> float:
        """Solves L x = b in place on ``x``; returns the device time of the sweep in ms."""
        ms = c_double(0.0)
        f = self._L.parsy_cuda_csc_solve
        f.restype = c_int
        f.argtypes = [c_void_p, c_void_p, POINTER(c_double)]
        rc = f(self._h, _f64_inplace(x, "x"), byref(ms))
        if rc != OK:
            raise ParsyCudaError(rc, "parsy_cuda_csc_solve")
        return float(ms.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            f = self._L.parsy_cuda_csc_destroy
            f.restype = None
            f.argtypes = [c_void_p]
            f(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_check(n, lC, lR, Li_ptr, blockSet, supNo, col2Sup, nLevels, levelPtr, parPtr, partition, block_cols=0,
               ignore_hlevels=False, rank=0, world=1, phase=0, top_levels=1):
    """Host-only planner run (no device needed): returns (status code, stats dict)."""
    L = lib()
    opt = Options()
    opt.block_cols, opt.ignore_hlevels, opt.use_graph = int(block_cols), int(ignore_hlevels), 1
    opt.rank, opt.world, opt.reserved[2], opt.reserved[3] = int(rank), int(world), int(phase), int(top_levels)
    st = Stats()
    f = L.parsy_cuda_plan_check
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                  POINTER(Options), POINTER(Stats)]
    keep = [_u64(lC, "lC"), _i32(lR, "lR"), _u64(Li_ptr, "Li_ptr"), _i32(blockSet, "blockSet"),
            _i32(col2Sup, "col2Sup"), _i32(levelPtr, "levelPtr", True), _i32(parPtr, "parPtr", True),
            _i32(partition, "partition", True)]
    p = [k[1] for k in keep]
    rc = f(int(n), p[0], p[1], p[2], p[3], int(supNo), p[4], int(nLevels), p[5], p[6], p[7], byref(opt), byref(st))
    return rc, st.as_dict()


def plan_digest(n, lC, lR, Li_ptr, blockSet, supNo, col2Sup, nLevels, levelPtr, parPtr, partition, block_cols=0,
                ignore_hlevels=False, rank=0, world=1, phase=0, top_levels=1, top_chunk=0):
    """Host-only: 64-bit digest of the plan the executor would run (parsy_cuda_plan_digest)."""
    L = lib()
    opt = Options()
    opt.block_cols, opt.ignore_hlevels, opt.use_graph = int(block_cols), int(ignore_hlevels), 1
    opt.rank, opt.world, opt.reserved[2], opt.reserved[3] = int(rank), int(world), int(phase), int(top_levels)
    opt.reserved[7] = int(top_chunk)
    out = ctypes.c_uint64(0)
    f = L.parsy_cuda_plan_digest
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                  POINTER(Options), POINTER(ctypes.c_uint64)]
    keep = [_u64(lC, "lC"), _i32(lR, "lR"), _u64(Li_ptr, "Li_ptr"), _i32(blockSet, "blockSet"),
            _i32(col2Sup, "col2Sup"), _i32(levelPtr, "levelPtr", True), _i32(parPtr, "parPtr", True),
            _i32(partition, "partition", True)]
    p = [k[1] for k in keep]
    rc = f(int(n), p[0], p[1], p[2], p[3], int(supNo), p[4], int(nLevels), p[5], p[6], p[7], byref(opt), byref(out))
    if rc != OK:
        raise ParsyCudaError(rc, "parsy_cuda_plan_digest")
    return int(out.value)


def plan_owned_ranges(n, lC, lR, Li_ptr, blockSet, supNo, col2Sup, nLevels, levelPtr, parPtr, partition, world,
                      for_rank, top_levels=1):
    """Host-only: contiguous [begin, end) runs of lValues owned by `for_rank` when sharding over `world` ranks."""
    L = lib()
    f = L.parsy_cuda_plan_owned_ranges
    f.restype = c_int
    f.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                  c_int, c_int, c_int, c_void_p, c_int]
    keep = [_u64(lC, "lC"), _i32(lR, "lR"), _u64(Li_ptr, "Li_ptr"), _i32(blockSet, "blockSet"),
            _i32(col2Sup, "col2Sup"), _i32(levelPtr, "levelPtr"), _i32(parPtr, "parPtr"), _i32(partition, "partition")]
    p = [k[1] for k in keep]
    cnt = f(int(n), p[0], p[1], p[2], p[3], int(supNo), p[4], int(nLevels), p[5], p[6], p[7], int(world),
            int(top_levels), int(for_rank), None, 0)
    if cnt < 0:
        raise ParsyCudaError(ERR_BAD_ARG, "parsy_cuda_plan_owned_ranges")
    out = np.zeros(2 * max(cnt, 1), np.int64)
    f(int(n), p[0], p[1], p[2], p[3], int(supNo), p[4], int(nLevels), p[5], p[6], p[7], int(world), int(top_levels),
      int(for_rank), out.ctypes.data_as(c_void_p), cnt)
    return out[:2 * cnt].reshape(-1, 2)


# ---------------------------------------------------------------------------------------------------------
# resident handle
# ---------------------------------------------------------------------------------------------------------
class Solver:
    """Device-resident symbolic state + factor. Arrays are the inspector's outputs."""

    def __init__(self, n, c, r, lC, lR, Li_ptr, blockSet, supNo, aTree, col2Sup, nLevels, levelPtr, parPtr, partition,
                 device=0, block_cols=0, use_graph=True, ignore_hlevels=False, lookahead=True, dataflow_sweeps=True,
                 narrow_sweeps=True, fan_out=True):
        L = lib()
        self._L = L
        self._h = c_void_p()
        opt = Options()
        opt.device, opt.block_cols, opt.use_graph, opt.ignore_hlevels = int(device), int(block_cols), int(use_graph), \
            int(ignore_hlevels)
        opt.reserved[0], opt.reserved[1] = int(not lookahead), int(not dataflow_sweeps)
        opt.reserved[5], opt.reserved[6] = int(not narrow_sweeps), int(not fan_out)
        f = L.parsy_cuda_create
        f.restype = c_int
        f.argtypes = [POINTER(c_void_p), c_int] + [c_void_p] * 6 + [c_int, c_void_p, c_void_p, c_int] + \
            [c_void_p] * 3 + [POINTER(Options)]
        keep = [_i32(c, "c", True), _i32(r, "r", True), _u64(lC, "lC"), _i32(lR, "lR"), _u64(Li_ptr, "Li_ptr"),
                _i32(blockSet, "blockSet"), _i32(aTree, "aTree", True), _i32(col2Sup, "col2Sup"),
                _i32(levelPtr, "levelPtr", True), _i32(parPtr, "parPtr", True), _i32(partition, "partition", True)]
        p = [k[1] for k in keep]
        rc = f(byref(self._h), int(n), p[0], p[1], p[2], p[3], p[4], p[5], int(supNo), p[6], p[7], int(nLevels), p[8],
               p[9], p[10], byref(opt))
        if rc != OK:
            self._h = c_void_p()
            raise ParsyCudaError(rc, "parsy_cuda_create")
        self.n = int(n)
        st = self.stats()
        self.xsize, self.nnzA = st["xsize"], st["nnzA"]

    def _call(self, name, *args, ok=(OK,)):
        f = getattr(self._L, name)
        f.restype = c_int
        f.argtypes = [c_void_p] + [c_void_p if not isinstance(a, int) else c_int for a in args]
        rc = f(self._h, *args)
        if rc not in ok:
            raise ParsyCudaError(rc, name)
        return rc

    def set_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.size != self.nnzA:
            raise ValueError("values has the wrong length")
        self._call("parsy_cuda_set_values", v.ctypes.data_as(c_void_p))
        self._call("parsy_cuda_sync", ok=(OK, ERR_NOT_SPD))  # host buffer may be released after return

    def factor(self):
        self._call("parsy_cuda_factor")

    def sync(self) -> bool:
        """Waits for the device; returns False if the last factorization met a non-positive pivot."""
        return self._call("parsy_cuda_sync", ok=(OK, ERR_NOT_SPD)) == OK

    def get_factor(self, out=None):
        if out is None:
            out = np.empty(self.xsize, np.float64)
        self._call("parsy_cuda_get_factor", _f64_inplace(out, "out"))
        return out

    def set_factor(self, lValues):
        v = np.ascontiguousarray(lValues, dtype=np.float64)
        if v.size != self.xsize:
            raise ValueError("lValues has the wrong length")
        self._call("parsy_cuda_set_factor", v.ctypes.data_as(c_void_p))
        self._call("parsy_cuda_sync", ok=(OK, ERR_NOT_SPD))

    def set_rhs(self, b):
        v = np.ascontiguousarray(b, dtype=np.float64)
        if v.size != self.n:
            raise ValueError("rhs has the wrong length")
        self._call("parsy_cuda_set_rhs", v.ctypes.data_as(c_void_p))
        self._call("parsy_cuda_sync", ok=(OK, ERR_NOT_SPD))

    def get_rhs(self, out=None):
        if out is None:
            out = np.empty(self.n, np.float64)
        self._call("parsy_cuda_get_rhs", _f64_inplace(out, "out"))
        return out

    def solve(self, which=SOLVE_FWD | SOLVE_BWD):
        self._call("parsy_cuda_solve", int(which))

    def set_permutation(self, perm):
        """perm[k] = caller's index of the unknown at position k of the factored matrix (the inspector's ``Perm``);
        None = identity."""
        pp, ptr = _i32(perm, "perm", allow_none=True)
        if pp is not None and pp.size != self.n:
            raise ValueError("perm must have n entries")
        self._call("parsy_cuda_set_permutation", ptr)

    def solve_system(self, b, refine_steps=0, residuals=False):
        """A x = b in the caller's ordering: permute, forward + backward sweep, ``refine_steps`` rounds of iterative
        refinement, permute back.  ``b``: (n,) or (nrhs, n) C-contiguous rows (one right-hand side per row).
        Returns x (same shape), or (x, rel) with rel[j, k] = relative residual of column j before refinement step k
        (last entry: final) when ``residuals`` is true."""
        B = np.ascontiguousarray(b, np.float64)
        one = B.ndim == 1
        B2 = B.reshape(1, -1) if one else B
        if B2.ndim != 2 or B2.shape[1] != self.n:
            raise ValueError("b must be (n,) or (nrhs, n)")
        nrhs = B2.shape[0]
        X = np.empty_like(B2)
        rel = np.zeros((nrhs, refine_steps + 1)) if residuals else None
        f = self._L.parsy_cuda_solve_system
        f.restype = c_int
        f.argtypes = [c_void_p, c_void_p, c_void_p, c_int, ctypes.c_int64, c_int, c_void_p]
        rc = f(self._h, B2.ctypes.data_as(c_void_p), X.ctypes.data_as(c_void_p), nrhs, self.n, int(refine_steps),
               None if rel is None else rel.ctypes.data_as(c_void_p))
        if rc != OK:
            raise ParsyCudaError(rc, "parsy_cuda_solve_system")
        X = X[0] if one else X
        return (X, rel[0] if one else rel) if residuals else X

    def factor_times(self):
        t = np.zeros(3, np.float64)
        self._call("parsy_cuda_factor_times", t.ctypes.data_as(c_void_p))
        return {"levels": t[0], "last_level": t[1], "assemble": t[2]}

    KERNEL_CLASSES = ("factor_small", "potrf_block", "trsm_tiles_dmma", "update_tiles128_dmma", "update_tiles64_dmma",
                      "update_small")

    def factor_profiled(self):
        """One event-instrumented factorization (no graphs): {class: (ms, launches, algorithmic flops)}."""
        ms = np.zeros(6, np.float64)
        nl = np.zeros(6, np.int64)
        fl = np.zeros(6, np.float64)
        self._call("parsy_cuda_factor_profiled", ms.ctypes.data_as(c_void_p), nl.ctypes.data_as(c_void_p),
                   fl.ctypes.data_as(c_void_p))
        return {k: {"ms": float(ms[i]), "launches": int(nl[i]), "flops": float(fl[i])}
                for i, k in enumerate(self.KERNEL_CLASSES)}

    def factor_trace(self, max_records=1 << 16):
        """One event-instrumented factorization: per-launch arrays (step, kernel class index, ms)."""
        step = np.zeros(max_records, np.int32)
        cls = np.zeros(max_records, np.int32)
        ms = np.zeros(max_records, np.float32)
        f = self._L.parsy_cuda_factor_trace
        f.restype = c_int
        f.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        cnt = f(self._h, int(max_records), step.ctypes.data_as(c_void_p), cls.ctypes.data_as(c_void_p),
                ms.ctypes.data_as(c_void_p))
        if cnt < 0:
            raise ParsyCudaError(ERR_CUDA, "parsy_cuda_factor_trace")
        cnt = min(cnt, max_records)
        return step[:cnt], cls[:cnt], ms[:cnt]

    def stats(self):
        st = Stats()
        f = self._L.parsy_cuda_get_stats
        f.restype = c_int
        f.argtypes = [c_void_p, POINTER(Stats)]
        rc = f(self._h, byref(st))
        if rc != OK:
            raise ParsyCudaError(rc, "parsy_cuda_get_stats")
        return st.as_dict()

    def owned_ranges(self, rank):
        """Contiguous [begin, end) runs of lValues (in doubles) owned by `rank` in the sharded factorization."""
        f = self._L.parsy_cuda_owned_ranges
        f.restype = c_int
        f.argtypes = [c_void_p, c_int, c_void_p, c_int]
        cnt = f(self._h, int(rank), None, 0)
        if cnt < 0:
            raise ParsyCudaError(ERR_BAD_ARG, "parsy_cuda_owned_ranges")
        out = np.zeros(2 * max(cnt, 1), np.int64)
        f(self._h, int(rank), out.ctypes.data_as(c_void_p), cnt)
        return out[:2 * cnt].reshape(-1, 2)

    def device_pointers(self):
        L = self._L
        out = {}
        for k in ("factor", "rhs", "values", ):
            f = getattr(L, f"parsy_cuda_device_{k}")
            f.restype = c_void_p
            f.argtypes = [c_void_p]
            out[k] = f(self._h)
        f = L.parsy_cuda_stream
        f.restype = c_void_p
        f.argtypes = [c_void_p]
        out["stream"] = f(self._h)
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            f = self._L.parsy_cuda_destroy
            f.restype = None
            f.argtypes = [c_void_p]
            f(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Borrowed(Solver):
    """A plan owned by a Sharded handle (introspection only: stats, owned_ranges)."""

    def __init__(self, L, h, owner):   # noqa: super().__init__ not called on purpose
        self._L, self._h, self._owner = L, c_void_p(h), owner
        st = self.stats()
        self.n, self.xsize, self.nnzA = st["n"], st["xsize"], st["nnzA"]

    def close(self):
        self._h = c_void_p()


def nccl_unique_id() -> bytes:
    """128 bytes from ncclGetUniqueId (rank 0 calls it and hands the bytes to every rank)."""
    buf = ctypes.create_string_buffer(128)
    f = lib().parsy_cuda_nccl_unique_id
    f.restype = c_int
    f.argtypes = [c_void_p]
    rc = f(buf)
    if rc != OK:
        raise ParsyCudaError(rc, "parsy_cuda_nccl_unique_id")
    return buf.raw


class Sharded:
    """ONE factorization + solve sharded over the GPUs of a node, one process per GPU (include/parsy_cuda.h section 3,
    DESIGN.md section 8).  ``unique_id``: the bytes of :func:`nccl_unique_id` from rank 0; ``None`` emulates all ``world``
    ranks in this process on ``device`` (tests)."""

    def __init__(self, n, c, r, lC, lR, Li_ptr, blockSet, supNo, aTree, col2Sup, nLevels, levelPtr, parPtr, partition,
                 rank, world, unique_id, device=0, block_cols=0, top_levels=1, top_distributed=True, use_graph=True,
                 lookahead=True, top_chunk=0, lanes=0):
        L = lib()
        self._L = L
        self._h = c_void_p()
        opt = Options()
        opt.device, opt.block_cols, opt.use_graph = int(device), int(block_cols), int(use_graph)
        opt.rank, opt.world = int(rank), int(world)
        opt.reserved[0], opt.reserved[3], opt.reserved[4] = int(not lookahead), int(top_levels), int(not top_distributed)
        opt.reserved[7], opt.reserved[8] = int(top_chunk), int(lanes)      # 0 = library defaults
        f = L.parsy_cuda_sharded_create
        f.restype = c_int
        f.argtypes = [POINTER(c_void_p), c_int] + [c_void_p] * 6 + [c_int, c_void_p, c_void_p, c_int] + \
            [c_void_p] * 3 + [POINTER(Options), c_void_p]
        keep = [_i32(c, "c"), _i32(r, "r"), _u64(lC, "lC"), _i32(lR, "lR"), _u64(Li_ptr, "Li_ptr"),
                _i32(blockSet, "blockSet"), _i32(aTree, "aTree", True), _i32(col2Sup, "col2Sup"),
                _i32(levelPtr, "levelPtr"), _i32(parPtr, "parPtr"), _i32(partition, "partition")]
        p = [k[1] for k in keep]
        uid = None
        if unique_id is not None:
            if len(unique_id) != 128:
                raise ValueError("unique_id must be the 128 bytes of nccl_unique_id()")
            uid = ctypes.create_string_buffer(bytes(unique_id), 128)
        rc = f(byref(self._h), int(n), p[0], p[1], p[2], p[3], p[4], p[5], int(supNo), p[6], p[7], int(nLevels), p[8],
               p[9], p[10], byref(opt), uid)
        if rc != OK:
            self._h = c_void_p()
            raise ParsyCudaError(rc, "parsy_cuda_sharded_create")
        self.n, self.rank, self.world, self.local = int(n), int(rank), int(world), unique_id is None
        self.xsize = int(np.asarray(lC)[n])
        self.nnzA = int(np.asarray(c)[n])

    _call = Solver._call

    def set_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.size != self.nnzA:
            raise ValueError("values has the wrong length")
        self._call("parsy_cuda_sharded_set_values", v.ctypes.data_as(c_void_p))
        self._call("parsy_cuda_sharded_sync", ok=(OK, ERR_NOT_SPD))

    def factor(self):
        self._call("parsy_cuda_sharded_factor")

    def sync(self) -> bool:
        return self._call("parsy_cuda_sharded_sync", ok=(OK, ERR_NOT_SPD)) == OK

    def set_rhs(self, b, sync=True):
        v = np.ascontiguousarray(b, dtype=np.float64)
        if v.size != self.n:
            raise ValueError("b has the wrong length")
        self._call("parsy_cuda_sharded_set_rhs", v.ctypes.data_as(c_void_p))
        if sync:
            self._call("parsy_cuda_sharded_sync", ok=(OK, ERR_NOT_SPD))

    def solve(self, which=SOLVE_FWD | SOLVE_BWD):
        self._call("parsy_cuda_sharded_solve", int(which))

    def get_rhs(self, out=None):
        x = np.empty(self.n) if out is None else out
        self._call("parsy_cuda_sharded_get_rhs", _f64_inplace(x, "x"))
        return x

    def get_factor(self, out=None):
        """The panels this process holds complete (own subtrees + top; emulation: all) written into ``out`` (reference
        layout); other entries keep their previous content (zeros for a fresh array)."""
        lv = np.zeros(self.xsize) if out is None else out
        self._call("parsy_cuda_sharded_get_factor", _f64_inplace(lv, "lValues"))
        return lv

    def phase_times(self):
        t = np.zeros(3)
        self._call("parsy_cuda_sharded_phase_times", t.ctypes.data_as(c_void_p))
        return {"phase1": t[0], "sum_top": t[1], "top": t[2]}

    def stats(self):
        v = np.zeros(12, np.int64)
        self._call("parsy_cuda_sharded_stats", v.ctypes.data_as(c_void_p))
        keys = ("launches_factor", "nccl_broadcasts", "nccl_allreduces", "bytes_broadcast", "bytes_summed", "device_bytes",
                "top_chain_steps", "nccl_version", "launches_fwd", "launches_bwd", "owned_supernodes", "top_supernodes")
        return {k: int(x) for k, x in zip(keys, v)}

    def trace_top(self, max_steps=4096):
        """(times[steps, 7] in ms, owner[steps]) of one un-graphed factorization: see parsy_cuda_sharded_trace_top."""
        out = np.zeros((max_steps, 7), np.float32)
        own = np.zeros(max_steps, np.int32)
        f = self._L.parsy_cuda_sharded_trace_top
        f.restype = c_int
        f.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
        cnt = f(self._h, int(max_steps), out.ctypes.data_as(c_void_p), own.ctypes.data_as(c_void_p))
        if cnt < 0:
            raise ParsyCudaError(ERR_STATE, "parsy_cuda_sharded_trace_top")
        return out[:min(cnt, max_steps)], own[:min(cnt, max_steps)]

    def plan(self, phase, emulated_rank=0):
        f = self._L.parsy_cuda_sharded_plan
        f.restype = c_void_p
        f.argtypes = [c_void_p, c_int, c_int]
        h = f(self._h, int(emulated_rank), int(phase))
        if not h:
            raise ParsyCudaError(ERR_BAD_ARG, "parsy_cuda_sharded_plan")
        return _Borrowed(self._L, h, self)

    def device_pointers(self):
        out = {}
        for k in ("device_factor", "device_rhs", "stream"):
            f = getattr(self._L, f"parsy_cuda_sharded_{k}")
            f.restype = c_void_p
            f.argtypes = [c_void_p]
            out[k.replace("device_", "")] = f(self._h)
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            f = self._L.parsy_cuda_sharded_destroy
            f.restype = None
            f.argtypes = [c_void_p]
            f(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
