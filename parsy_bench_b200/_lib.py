"""ctypes loader of libparsy_cuda.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libparsy_cuda.so")
_lib = None

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_size_t_p = ctypes.POINTER(ctypes.c_size_t)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class Options(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("block_cols", ctypes.c_int), ("use_graph", ctypes.c_int),
                ("ignore_hlevels", ctypes.c_int), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("reserved", ctypes.c_int * 10)]


class Stats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in ("n", "nsuper", "xsize", "ssize", "nnzA", "n_pairs", "n_pairs_small",
                                               "n_pairs_tiled", "n_steps", "n_block_cols", "rel_entries",
                                               "launches_factor", "launches_fwd", "launches_bwd")] + \
               [(k, ctypes.c_double) for k in ("flops_potrf", "flops_trsm", "flops_update", "bytes_solve")] + \
               [("device_bytes", ctypes.c_int64), ("reserved", ctypes.c_int64 * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["reserved"] = list(self.reserved)
        return d


def lib():
    """Loads the shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the executor)")
    L = ctypes.CDLL(LIB_PATH)
    L.parsy_cuda_last_error.restype = ctypes.c_char_p
    L.parsy_cuda_device_factor.restype = ctypes.c_void_p
    L.parsy_cuda_device_rhs.restype = ctypes.c_void_p
    L.parsy_cuda_device_values.restype = ctypes.c_void_p
    L.parsy_cuda_stream.restype = ctypes.c_void_p
    for name in ("parsy_cuda_device_factor", "parsy_cuda_device_rhs", "parsy_cuda_device_values", "parsy_cuda_stream",
                 "parsy_cuda_destroy", "parsy_cuda_set_values", "parsy_cuda_factor", "parsy_cuda_sync",
                 "parsy_cuda_get_factor", "parsy_cuda_set_factor", "parsy_cuda_set_rhs", "parsy_cuda_get_rhs",
                 "parsy_cuda_factor_times", "parsy_cuda_get_stats"):
        getattr(L, name).argtypes = None
    _lib = L
    return L


def last_error() -> str:
    return lib().parsy_cuda_last_error().decode()
