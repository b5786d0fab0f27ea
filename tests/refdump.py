"""Test-side access to the compiled reference (oracle/_ref/parsy_ref): runs it on a synthetic Laplacian and
loads the arrays it dumps.  TEST INFRASTRUCTURE — never imported by the product package."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "parsy_ref")
REF_GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "parsy_ref_gpu")   # same driver, call sites forwarded to libparsy_cuda
_DT = {"i32": np.int32, "u64": np.uint64, "f64": np.float64}
_CACHE = {}


def have_ref():
    return os.path.exists(REF_BIN)


class RefCase(dict):
    __getattr__ = dict.__getitem__


def ref_case(kind, N, cost=8, level=1, div=2, threads=1, factor=True, solve=True, iters=1, keep_values=True, mtx=None,
             blas_threads=None, csc=True, cache=True, binary=None):
    """mtx: path of a lower-half Matrix-Market file read by the reference's readMatrix instead of the synthetic grid.
    blas_threads: BLAS threads of the LAST (sequential) H-level only (parallel_PB_Cholesky_05.h:271); the parallel levels
    stay on `threads` OpenMP threads — 1 for goldens, because of the `top` race (:43,69,115)."""
    key = (kind, N, cost, level, div, threads, factor, solve, mtx, blas_threads, csc, binary, iters)
    if cache and key in _CACHE:
        return _CACHE[key]
    d = tempfile.mkdtemp(prefix="parsy_ref_")
    cmd = [binary or REF_BIN, "--kind", kind, "--N", str(N), "--cost", str(cost), "--level", str(level), "--div", str(div),
           "--threads", str(threads), "--iters", str(iters), "--dump", d]
    if mtx is not None:
        cmd += ["--mtx", str(mtx)]
    if blas_threads is not None:
        cmd += ["--blas-threads", str(blas_threads)]
    if not csc:
        cmd.append("--no-csc")
    if not factor:
        cmd.append("--no-factor")
    if not solve:
        cmd.append("--no-solve")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads))
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=env).stdout
    meta = json.loads(out.strip().splitlines()[-1])
    case = RefCase(meta=meta, dir=d)
    for f in os.listdir(d):
        name, dt = f.rsplit(".", 1)
        case[name] = np.fromfile(os.path.join(d, f), dtype=_DT[dt])
        os.unlink(os.path.join(d, f))
    os.rmdir(d)
    if cache:
        _CACHE[key] = case
    return case


def rel_err(a, b, floor_scale=1e-6):
    """max |a-b| / max(|a|,|b|, floor) with floor = floor_scale * max|b| (SURVEY.md §7 'scatter epilogue')."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = floor_scale * float(np.max(np.abs(b))) if b.size else 1.0
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return float(np.max(np.abs(a - b) / den)) if b.size else 0.0
