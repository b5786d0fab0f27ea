"""Timeline of the general sweep kernels (parsy_cuda_sweep_trace): where the time of k_fwd_dataflow / k_bwd_dataflow
goes.  Usage: trace_sweep.py <2d5|3d7|3d27> <N> [bwd]"""
import ctypes
import os
import sys
from ctypes import c_int, c_void_p

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsy_bench_b200 import executor as ex, inspector, matrices  # noqa: E402

kind, N = sys.argv[1], int(sys.argv[2])
which = ex.SOLVE_BWD if len(sys.argv) > 3 and sys.argv[3] == "bwd" else ex.SOLVE_FWD
n, Ap, Ai, Ax = matrices.laplacian(kind, N)
S = inspector.analyze(n, Ap, Ai, Ax, 592, 1, 4)
H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels, S.levelPtr,
              S.parPtr, S.partition)
H.set_values(S.A2_x)
H.factor()
assert H.sync()
b = 1.0 + np.arange(n) / n
for rep in range(2):
    H.set_rhs(b)
    cap = 1 << 20
    k = np.zeros(cap, np.int32); nr = np.zeros(cap, np.int32)
    ts, tr, te = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    f = H._L.parsy_cuda_sweep_trace
    f.restype = c_int
    f.argtypes = [c_void_p, c_int, c_int] + [c_void_p] * 5
    if which == ex.SOLVE_BWD:
        H.solve(ex.SOLVE_FWD)
    cnt = f(H._h, which, cap, *[a.ctypes.data_as(c_void_p) for a in (k, nr, ts, tr, te)])
k, nr, ts, tr, te = k[:cnt], nr[:cnt], ts[:cnt], tr[:cnt], te[:cnt]
print(f"# {kind} {N}: {cnt} CTAs in {'k_bwd_dataflow' if which == ex.SOLVE_BWD else 'k_fwd_dataflow'}, span {te.max():.1f} us")
if which == ex.SOLVE_BWD:      # the backward kernel walks the plan from its end
    k, nr, ts, tr, te = k[::-1], nr[::-1], ts[::-1], tr[::-1], te[::-1]
for kk, name in ((0, "narrow x8"), (2, "narrow tall"), (1, "block slice")):
    m = k == kk
    if m.any():
        print(f"{name:12s} n={m.sum():6d}  wait (ready-start) mean {np.mean(tr[m]-ts[m]):7.2f} us  "
              f"life (end-start) mean {np.mean(te[m]-ts[m]):6.2f} max {np.max(te[m]-ts[m]):6.2f}  "
              f"work (end-ready) mean {np.mean(te[m]-tr[m]):6.2f} p50 {np.median(te[m]-tr[m]):6.2f} p95 {np.percentile(te[m]-tr[m],95):6.2f} max {np.max(te[m]-tr[m]):6.2f} us")
for rows in (0, 64, 256):
    m = (k == 1) & (nr == rows)
    if m.any():
        print(f"  slices with {rows:3d} rows: n={m.sum():5d} work mean {np.mean(te[m]-tr[m]):6.2f} us")
# progress: time at which the i-th decile of CTAs (ticket order) is ready / done
print("ticket decile: start / ready / end (us)")
for q in range(0, 101, 10):
    i = min(cnt - 1, cnt * q // 100)
    print(f"  {q:3d}%  cta {i:6d} kind {k[i]}  {ts[i]:8.1f} {tr[i]:8.1f} {te[i]:8.1f}")
# chain estimate: successive 'ready' times of CTAs that had to wait (ready - start > 1 us) near the end of the kernel
w = np.where((tr - ts) > 1.0)[0]
print(f"CTAs that waited > 1 us: {len(w)} of {cnt}; total wait {np.sum(tr - ts)/1e3:.2f} CTA-ms, total work {np.sum(te - tr)/1e3:.2f} CTA-ms")
order = np.argsort(tr)
gaps = np.diff(tr[order])
print(f"ready-time gaps between consecutive readiness events: mean {gaps.mean():.3f} us, p99 {np.percentile(gaps,99):.2f}, max {gaps.max():.2f}")
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"sweep_trace_{kind}_{N}_{'bwd' if which == ex.SOLVE_BWD else 'fwd'}.npz"),
                    kind=k, nrows=nr, start=ts, ready=tr, end=te)
