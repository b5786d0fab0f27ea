#!/usr/bin/env python
"""bench.py — ParSy hot path on B200: LBC supernodal Cholesky (FP64 GFLOP/s) + supernodal SpTRSV (ms).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg1|cfg4|...]

A "step" is one pass of the numeric hot path over the workload of BASELINE.json configs[1] (2D 5-point Laplacian
1000x1000, n = 1e6): numeric factorization (zero L, scatter A, every LBC level) + forward sweep + backward sweep,
with structure, A and b already resident in HBM.  `value` = sum_j ColCount[j]^2 / (factorization share of the step,
CUDA events) in GFLOP/s (cholesky/ColumnCount.h:486-498 flop count); `e2e` is the same metric through the public handle API with
host buffers: H2D of A's values and b from pinned memory, factor, both sweeps, D2H of x, every step.
One JSON line on stdout (rank 0).  `--impl reference` times the reference's own OpenMP executor (oracle/_ref, built
from /root/reference at build time) on the box's host cores for the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (stencil, N, description)
    "cfg1": ("2d5", 100, "2D 5-point Laplacian 100x100 (n=1e4)"),
    "cfg2": ("2d5", 1000, "2D 5-point Laplacian 1000x1000 (n=1e6) factor + forward/backward solve"),
    "cfg3": ("3d7", 100, "3D 7-point Laplacian 100^3 (n=1e6)"),
    "cfg4": ("3d27", 64, "3D 27-point Laplacian 64^3 (n=262144)"),
    "cfg5": ("3d27", 160, "3D 27-point Laplacian 160^3 (n=4.1M)"),
}
FP64_PEAK_TFLOPS = 37.1      # tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak.jsonl): DMMA.8x8x4 issue rate
HBM_FALLBACK_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback


_JSON_FD = None


def quiet_stdout():
    """Everything but the final JSON line goes to stderr: libraries print banners on stdout (NCCL prints its version
    there at communicator creation), and the contract is ONE JSON line on stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_binary(kind, N, threads, iters, solve=True):
    """oracle/_ref/parsy_ref = the reference's own headers compiled as they are (oracle/build_ref.sh)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "parsy_ref")
    if not os.path.exists(exe):
        return None
    cmd = [exe, "--kind", kind, "--N", str(N), "--cost", str(threads), "--level", "1", "--div", "2", "--threads",
           str(threads), "--iters", str(iters), "--no-dump-values"]
    if not solve:
        cmd.append("--no-solve")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3000)
    if out.returncode != 0:
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_baseline(kind, N, flops_hint=None):
    """Reference OpenMP executor on the host cores; bounded: one warm + two timed factorizations of the workload
    (falls back to the oracle's scalar C port on a 1/4-size grid if the compiled reference is absent)."""
    threads = host_threads()
    r = run_reference_binary(kind, N, threads, 3, solve=True)
    if r is not None:
        t = float(np.median(r["t_factor_all"][1:])) if len(r["t_factor_all"]) > 1 else r["t_factor"]
        return {"value": r["flops"] / t / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "reference",
                "sample": f"full workload, median of 2 warm factorizations of cholesky_left_par_05 (costParam={threads}, "
                          f"levelParam=1, divRate=2); inspector {r['t_inspector']:.1f}s not counted",
                "factor_s": t, "levels_s": r["t_levels"], "last_level_s": r["t_last"],
                "sptrsv_ms": {"blockedLsolve": r["t_blockedLsolve"] * 1e3, "leveledBlockedLsolve": r["t_leveled"] * 1e3,
                              "H2LeveledBlockedLsolve": r["t_h2"] * 1e3,
                              "H2LeveledBlockedLsolve_Peeled": r["t_h2_peeled"] * 1e3}}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import parsy_oracle as orc
    from parsy_bench_b200 import inspector, matrices
    Ns = max(8, N // 4)
    n, Ap, Ai, Ax = matrices.laplacian(kind, Ns)
    S = inspector.analyze(n, Ap, Ai, Ax, 8, 1, 2)
    t0 = time.time()
    orc.cholesky_left_par_05(S)
    t = time.time() - t0
    return {"value": S.flops / t / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "sample": f"oracle C port, one factorization of the {kind} grid N={Ns} (compiled reference absent)",
            "factor_s": t}


def reference_arm(args, kind, N, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    iters = args.warmup + args.steps
    t_wall = time.time()
    r = run_reference_binary(kind, N, threads, iters, solve=True)
    if r is None:
        base = cpu_baseline(kind, N)
        val, ms, cb = base["value"], base["factor_s"] * 1e3, base
        sptrsv = None
    else:
        tf = r["t_factor_all"][args.warmup:]
        th = r["t_h2_all"][args.warmup:] if len(r.get("t_h2_all", [])) > args.warmup else [r["t_h2"]]
        t = float(np.mean(tf))
        val, ms = r["flops"] / t / 1e9, (t + float(np.mean(th))) * 1e3
        sptrsv = {"H2LeveledBlockedLsolve": float(np.mean(th)) * 1e3, "blockedLsolve": r["t_blockedLsolve"] * 1e3}
        cb = {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "reference",
              "sample": f"full workload, {len(tf)} timed factorizations + forward solves after {args.warmup} warm-up "
                        f"(costParam={threads}, levelParam=1, divRate=2)"}
    line = {"impl": "reference", "metric": "cholesky_factor_gflops", "value": val, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "executor": "ParSy OpenMP cholesky_left_par_05 + H2LeveledBlockedLsolve on "
                                                       f"{threads} host threads, OpenBLAS 0.3.15"},
            "cpu_baseline": cb, "sptrsv_ms": sptrsv,
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t_wall}
    emit(line)
    return 0


def sharded_arm(args, kind, N, desc, rank, world, local):
    """N > 1: ONE factorization sharded over the GPUs (strong scaling): owned bottom subtrees -> NCCL broadcast of the
    owners' panels over NVLink -> shared top on every rank.  SpTRSV is not sharded (replicas only)."""
    import torch
    import torch.distributed as dist
    from parsy_bench_b200 import inspector, matrices
    from parsy_bench_b200.sharded import ShardedCholesky

    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    t0 = time.time()
    S = inspector.analyze(n, Ap, Ai, Ax, args.cost, args.level, args.div)
    t_insp = time.time() - t0
    t0 = time.time()
    SC = ShardedCholesky(S, rank, world, local, top_levels=args.top_levels, block_cols=args.block_cols,
                         top_distributed=not args.replicate_top)
    t_create = time.time() - t0
    F = S.flops
    h_vals = torch.from_numpy(S.A2_x.copy()).pin_memory()
    SC.set_values(h_vals.numpy())
    st1, st2 = SC.h1.stats(), SC.h2.stats()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        SC.factor(dist)
    barrier()
    if not SC.sync():
        raise SystemExit("factorization failed: matrix not positive definite")
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(SC.s1)
    for _ in range(args.steps):
        SC.factor(dist)
    SC.s1.wait_stream(SC.s2)
    e1.record(SC.s1)
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    p1_ms, ex_ms, p2_ms = SC.phase_times_ms()
    # end to end: A's values from pinned host memory every step, a device->host read of the result's status + tail
    tail = torch.empty(1024, dtype=torch.float64).pin_memory()
    barrier()
    tw = time.perf_counter()
    for _ in range(args.steps):
        SC.set_values(h_vals.numpy())
        SC.factor(dist)
        ok = SC.sync()
        tail.copy_(SC.lv[-1024:])
    barrier()
    e2e_ms = (time.perf_counter() - tw) * 1e3 / args.steps
    # parity spot check against the unsharded device factorization of rank 0 is done in tests/; here ||L||_F^2 = tr(A)
    fro = float((SC.lv * SC.lv).sum().item())
    trace = float(S.A2_x[S.A2_p[:-1]].sum())
    prof = SC.h2.factor_profiled() if rank == 0 else None
    t_loc = torch.tensor([total_ms, e2e_ms, p1_ms, ex_ms, p2_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, p1_ms, ex_ms, p2_ms = [float(v) for v in t_loc.tolist()]
    per_step = total_ms / args.steps
    if rank == 0:
        dmma = [k for k in prof if k.endswith("dmma")]
        dm_ms = sum(prof[k]["ms"] for k in dmma)
        dm_fl = sum(prof[k]["flops"] for k in dmma)
        dm_n = sum(prof[k]["launches"] for k in dmma)
        ach = dm_fl / (dm_ms * 1e-3) / 1e12 if dm_ms > 0 else 0.0
        line = {
            "metric": "cholesky_factor_gflops", "value": F / (per_step * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + " (factorization only at N>1; SpTRSV: replicas only)", "n": n,
                       "nnzL": int(S.xsize), "flops_sum_cc2": F,
                       "lbc": {"costParam": args.cost, "levelParam": args.level, "divRate": args.div,
                               "hlevels": int(S.nLevels), "wpartitions": int(S.nParts)},
                       "parallelism": f"bottom subtrees sharded over {world} GPUs; {args.top_levels} top H-level(s) "
                                      + ("block-cyclic by target block column, panel broadcast before each step; "
                                         if SC.top_distributed else "replicated; ")
                                      + f"{SC.n_broadcasts} NCCL broadcasts per factorization",
                       "l2": f"working set {8 * S.xsize / 1e6:.0f} MB (factor) > 126 MB L2, no flush needed"},
            "breakdown_ms": {"phase1_owned_subtrees_max_rank": p1_ms, "bottom_panel_exchange_max_rank": ex_ms,
                             "phase2_top_separators_max_rank": p2_ms,
                             "bottom_exchange_bytes_received_rank0": SC.exchange_bytes},
            "e2e": {"value": F / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * S.nnzA), "d2h_bytes_per_step": 8 * 1024 + 4,
                    "api": "ShardedCholesky.set_values (pinned) + factor + sync + read-back of the factor's tail"},
            "gpu_launches": int(args.steps * (st1["launches_factor"] + st2["launches_factor"])),
            "roofline": {"bound": "tensor", "kernel": "k_gemm_tiles (FP64 DMMA) in the shared-top phase of rank 0",
                         "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                         "traffic": None, "launches": dm_n,
                         "whole_factor_frac_of_fp64_peak": F / (per_step * 1e-3) / 1e12 / (FP64_PEAK_TFLOPS * world)},
            "cpu_baseline": None, "clocks": clocks, "fro2_over_trace": fro / trace, "factor_ok": bool(ok),
            "setup_s": {"inspector": t_insp, "create": t_create},
        }
        emit(line)
    SC.close()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 100 (cfg1, cfg2), 20 (cfg4), 5 (cfg3), 2 (cfg5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--cost", type=int, default=592, help="LBC innerParts handed to the inspector (GPU choice)")
    ap.add_argument("--div", type=int, default=4)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--block-cols", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true", help="single stream, no overlap of POTRF/TRSM with the bulk updates")
    ap.add_argument("--ignore-hlevels", action="store_true", help="schedule by dependencies only (no LBC H-level barriers)")
    ap.add_argument("--no-fan-out", action="store_true", help="kernel classes of a step on one stream (A/B)")
    ap.add_argument("--general-sweeps", action="store_true", help="leaf region of the sweeps on the general dataflow kernel (A/B)")
    ap.add_argument("--replicas", action="store_true", help="N>1: one full factorization per GPU (default for cfg1/cfg2/cfg4)")
    ap.add_argument("--sharded", action="store_true", help="N>1: ONE factorization sharded over the GPUs (default for cfg3/cfg5)")
    ap.add_argument("--replicate-top", action="store_true", help="N>1: every rank computes the top separators")
    ap.add_argument("--top-levels", type=int, default=1, help="N>1: LBC H-levels kept shared (computed by every rank)")
    args = ap.parse_args()
    quiet_stdout()
    if args.steps is None:     # long enough a timed region for the clock sampler, short enough to end within a minute
        args.steps = {"cfg1": 100, "cfg2": 100, "cfg4": 20, "cfg3": 5, "cfg5": 2}[args.config]
        if args.impl == "reference":
            args.steps = min(args.steps, 10)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    kind, N, desc = CONFIGS[args.config]
    if args.impl == "reference":
        return reference_arm(args, kind, N, desc)

    import torch
    import torch.distributed as dist
    from parsy_bench_b200 import executor as ex, inspector, matrices

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the executor has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # N > 1: a factorization that takes a few milliseconds on one GPU (cfg1/cfg2/cfg4) has nothing to gain from sharding —
    # measured: cfg2 3.7 ms on one GPU, 4.8 ms sharded over two — so those configs run one independent
    # factorization per GPU (weak scaling, no data-path collective); cfg3/cfg5, the configurations BASELINE.json names for
    # multi-GPU, run ONE factorization sharded over the GPUs (strong scaling).  --sharded / --replicas override.
    sharded = args.sharded or (args.config in ("cfg3", "cfg5") and not args.replicas)
    if world > 1 and sharded:
        return sharded_arm(args, kind, N, desc, rank, world, local)

    # ---- setup (untimed): synthetic matrix, host inspector, device-resident structure -------------------------
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    t0 = time.time()
    S = inspector.analyze(n, Ap, Ai, Ax, args.cost, args.level, args.div)
    t_insp = time.time() - t0
    t0 = time.time()
    H = ex.Solver(n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels, S.levelPtr,
                  S.parPtr, S.partition, device=local, block_cols=args.block_cols, ignore_hlevels=args.ignore_hlevels,
                  lookahead=not args.no_lookahead, narrow_sweeps=not args.general_sweeps, fan_out=not args.no_fan_out)
    t_create = time.time() - t0
    st = H.stats()
    F = S.flops
    F_sn = st["flops_potrf"] + st["flops_trsm"] + st["flops_update"]
    # pinned host buffers for the end-to-end path
    h_vals = torch.from_numpy(S.A2_x.copy()).pin_memory()
    b_host = (1.0 + np.arange(n) / n)[S.Perm]          # b_i = 1 + i/n in the original ordering, permuted
    h_b = torch.from_numpy(b_host.copy()).pin_memory()
    h_x = torch.empty(n, dtype=torch.float64).pin_memory()
    ptr = H.device_pointers()
    stream = torch.cuda.ExternalStream(ptr["stream"], device=torch.device("cuda", local))

    class DevArray:   # zero-copy torch view of the solver's device buffers
        def __init__(self, p, cnt):
            self.__cuda_array_interface__ = {"shape": (cnt,), "typestr": "<f8", "data": (p, False), "version": 3}
    d_rhs = torch.as_tensor(DevArray(ptr["rhs"], n), device=torch.device("cuda", local))
    d_b = torch.from_numpy(b_host).to(torch.device("cuda", local))
    H.set_values(h_vals.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(ev=None):
        with torch.cuda.stream(stream):
            d_rhs.copy_(d_b, non_blocking=True)       # device-to-device: b is resident
            if ev: ev[0].record(stream)
            H.factor()
            if ev: ev[1].record(stream)
            H.solve(ex.SOLVE_FWD)
            if ev: ev[2].record(stream)
            H.solve(ex.SOLVE_BWD)
            if ev: ev[3].record(stream)

    for _ in range(args.warmup):
        resident_step()
    barrier()
    if not H.sync():
        raise SystemExit("factorization failed: matrix not positive definite")
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        resident_step(evs[k])
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    fac_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    fwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    times = H.factor_times()

    # ---- end to end through the public handle API with host buffers ------------------------------------------------
    def e2e_step():
        H.set_values(h_vals.numpy())       # H2D nnz(A) doubles (pinned)
        H.set_rhs(h_b.numpy())             # H2D n doubles
        H.factor()
        H.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
        H.get_rhs(h_x.numpy())             # D2H n doubles (synchronises)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    x = h_x.numpy().copy()
    # residual of the permuted system (cheap sanity check, outside every timed region)
    import scipy.sparse as sp
    A2 = sp.csc_matrix((S.A2_x, S.A2_i, S.A2_p), shape=(n, n))
    res = float(np.linalg.norm(A2 @ x + sp.tril(A2, -1).T @ x - b_host) / np.linalg.norm(b_host))

    # strict drop-in call (reference signature, host arrays, factor downloaded) — once, reported beside e2e
    lv = np.empty(S.xsize)
    tm = np.zeros(4)
    t0 = time.perf_counter()
    ok = ex.cholesky_left_par_05(n, S.A2_p, S.A2_i, S.A2_x, S.p, S.s, S.i_ptr, lv, S.super, S.nsuper, tm, S.sParent,
                                 S.A1_p, S.A1_i, S.col2Sup, S.nLevels, S.levelPtr, None, 0, S.parPtr, S.partition)
    dropin_s = time.perf_counter() - t0
    del lv

    # ---- roofline of the dominant kernel: event-instrumented pass over the same workload --------------------------
    prof = H.factor_profiled()
    H.sync()
    tot_prof = sum(v["ms"] for v in prof.values())
    dmma = [k for k in prof if k.endswith("dmma")]
    dom = max(dmma, key=lambda k: prof[k]["ms"])
    dm_ms = sum(prof[k]["ms"] for k in dmma)
    dm_fl = sum(prof[k]["flops"] for k in dmma)
    dm_n = sum(prof[k]["launches"] for k in dmma)
    peaks, peak_kind = measured_peaks()
    achieved_tf = dm_fl / (dm_ms * 1e-3) / 1e12 if dm_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get(args.config, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "k_gemm_tiles (FP64 DMMA m8n8k4: SYRK/GEMM update + TRSM)",
                "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved_tf / FP64_PEAK_TFLOPS, "traffic": traffic,
                "traffic_note": "dram bytes per launch from the committed ncu --set full capture (profiles/r01_ncu_gemm_cfg2.txt)",
                "peak_source": "measured DMMA issue rate, tools/fp64_peak.cu (MEASURED_PEAKS.json has no FP64 entry)",
                "launches": dm_n, "avg_launch_us": dm_ms * 1e3 / max(dm_n, 1),
                "algorithmic_flops_per_step": dm_fl, "share_of_factor_time": dm_ms / tot_prof if tot_prof else None,
                "dominant_class": dom,
                "whole_factor_frac_of_fp64_peak": F / (fac_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS}
    hbm = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    roofline_solve = {"bound": "hbm", "kernel": "forward sweep (k_fwd_narrow for the leaf region + k_fwd_dataflow)",
                      "achieved": st["bytes_solve"] / (fwd_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                      "frac": st["bytes_solve"] / (fwd_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                      "peak_source": peak_kind + " copy bandwidth", "algorithmic_bytes": st["bytes_solve"],
                      "backward_frac": st["bytes_solve"] / (bwd_ms * 1e-3) / 1e9 / hbm}

    # ---- aggregate over ranks (max time; every rank runs the full workload: replicas) --------------------------------
    t_loc = torch.tensor([total_ms, fac_ms, fwd_ms, bwd_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
    total_ms, fac_ms, fwd_ms, bwd_ms, e2e_ms = [float(v) for v in t_loc.tolist()]
    value = world * F / (fac_ms * 1e-3) / 1e9
    e2e_value = world * F / (e2e_ms * 1e-3) / 1e9

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(kind, N)
    if rank == 0:
        launches = args.steps * (st["launches_factor"] + st["launches_fwd"] + st["launches_bwd"])
        line = {
            "metric": "cholesky_factor_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n": n, "nnzA": int(S.nnzA), "nsuper": int(S.nsuper), "nnzL": int(S.xsize),
                       "flops_sum_cc2": F, "flops_supernodal": F_sn,
                       "lbc": {"costParam": args.cost, "levelParam": args.level, "divRate": args.div,
                               "hlevels": int(S.nLevels), "wpartitions": int(S.nParts)},
                       "step": "zero L + scatter A + factor (all LBC levels) + forward sweep + backward sweep, resident",
                       "l2": f"working set {8 * S.xsize / 1e6:.0f} MB (factor) > 126 MB L2, no flush needed",
                       "parallelism": "single GPU" if world == 1 else f"{world} replicas: one independent factorization + solve per GPU, "
                                                                        "no data-path collective (sharding one factorization of "
                                                                        "this size over GPUs is slower than one GPU: --sharded, "
                                                                        "DESIGN.md §8)"},
            "breakdown_ms": {"factor": fac_ms, "fwd_solve": fwd_ms, "bwd_solve": bwd_ms,
                             "factor_levels": times["levels"] * 1e3, "factor_last_level": times["last_level"] * 1e3,
                             "assemble": times["assemble"] * 1e3},
            "sptrsv_ms": {"forward": fwd_ms, "backward": bwd_ms},
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * S.nnzA + 8 * n), "d2h_bytes_per_step": int(8 * n),
                    "api": "Solver.set_values + set_rhs + factor + solve(FWD|BWD) + get_rhs, pinned host buffers",
                    "dropin_cholesky_left_par_05_s": dropin_s, "dropin_ok": bool(ok),
                    "dropin_note": "reference signature: builds the plan, uploads, factors, downloads L (one cold call)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_solve": roofline_solve,
            "kernel_classes": prof,
            "cpu_baseline": cb,
            "clocks": clocks,
            "residual": res,
            "setup_s": {"inspector": t_insp, "inspector_metis": S.t_ordering, "create": t_create},
            "device_bytes": st["device_bytes"],
        }
        emit(line)
    H.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
