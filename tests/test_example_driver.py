"""The C++ integration example (examples/choleskyTest_b200.cpp): the reference driver's flow — readMatrix, analyze,
five timed cholesky_left_par_05 calls, CSV line (examples/choleskyTest01.cpp:118-276) — plus A x = b, over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import refdump
from parsy_bench_b200 import executor as ex, matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "choleskyTest_b200")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="examples/choleskyTest_b200 not built (run build())")


@needs_exe
def test_usage_and_loud_failure_without_a_device():
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if ex.device_count() == 0:
        r = subprocess.run([EXE, "2d5:12"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CUDA device" in r.stderr      # no CPU fallback
    r = subprocess.run([EXE, os.path.join(ROOT, "does_not_exist.mtx")], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid header" in r.stderr          # readMatrix's message for an unreadable file


@needs_exe
@pytest.mark.gpu
@pytest.mark.parametrize("spec", ["2d5:60", "3d27:10", "file"])
def test_driver_flow_on_the_gpu(tmp_path, spec):
    if spec == "file":
        n, Ap, Ai, Ax = matrices.laplacian("3d7", 12)
        spec = str(tmp_path / "lap.mtx")
        matrices.write_mtx(spec, n, Ap, Ai, Ax)
    r = subprocess.run([EXE, spec, "16", "1", "2", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    fields = r.stdout.strip().split(",")
    assert fields[0] == spec and len(fields) == 13
    t_all, t_levels, t_last = (float(v) for v in fields[7:10])
    assert t_all > 0 and t_levels >= 0 and t_last > 0 and t_all >= t_last
    tail = dict(kv.split("=") for kv in fields[12].split())
    assert float(tail["residual"]) < 1e-12 and abs(float(tail["device_residual"]) - float(tail["residual"])) < 1e-12


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(refdump.REF_BIN) and os.path.exists(refdump.REF_GPU_BIN)),
                    reason="compiled reference drivers (oracle/_ref) not present")
@pytest.mark.parametrize("case", [("2d5", 150, 8, 1, 2), ("3d27", 16, 16, 0, 2), ("3d7", 24, 592, 1, 4)])
def test_reference_driver_with_call_sites_forwarded_to_the_gpu(case):
    """oracle/_ref/parsy_ref_gpu is oracle/ref_driver.cpp — the reference's inspector (analyze_p2), its harness
    (rhsInitBlocked / testTriangular, common/Util.h:277-306) and its call sites — compiled with
    include/parsy_cuda_dropin.h, the forwarding header of INTEGRATION.md section 1, so that cholesky_left_par_05 and every
    forward solve run in libparsy_cuda on the arrays the REFERENCE's inspector produced.  Three factorizations per run
    (the second and third hit the structure cache of the drop-in entry points).  Compared with the unforwarded driver."""
    kind, N, c, l, d = case
    R = refdump.ref_case(kind, N, cost=c, level=l, div=d, threads=1)
    G = refdump.ref_case(kind, N, cost=c, level=l, div=d, threads=1, iters=3, binary=refdump.REF_GPU_BIN)
    assert G.meta["factor_ok"] == 1 and G.meta["solve_ok"] == 1          # testTriangular: x == 1 for b = L*1
    assert np.array_equal(G.partition, R.partition) and np.array_equal(G.p, R.p)
    assert refdump.rel_err(G.valL, R.valL) < 1e-9
    assert np.array_equal(G.valL == 0.0, R.valL == 0.0)
    for k in ("x_blocked", "x_h2", "y_ramp", "y_ramp_csc"):
        assert refdump.rel_err(G[k], R[k]) < 1e-9, k
