#!/usr/bin/env python
"""bench.py — ParSy hot path on B200: LBC supernodal Cholesky (FP64 GFLOP/s) + supernodal SpTRSV (ms).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg1|cfg4|...]

A "step" is one pass of the numeric hot path over the workload: numeric factorization (zero L, scatter A, every LBC
level) + forward sweep + backward sweep, with structure, A and b already resident in HBM.  Default workload: the largest
single-GPU configuration BASELINE.json names (configs[2], 3D 7-point Laplacian 100^3); at N = 1 the same line carries
configs[1] (2D 5-point 1000x1000, the configuration the metric is quoted on) as `sub_records.cfg2`; at N > 1 ONE
factorization + solve of configs[2] is sharded over the GPUs (`scaling: "strong"`).  `value` = sum_j ColCount[j]^2 / (factorization share of the step,
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
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from before the warm-up on; reported over the samples
    that fall into the timed regions (mark()), else over the samples taken under load."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc, self.windows = index, [], None, []

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

    def mark(self, t0, t1):
        """a timed region, in time.time() seconds"""
        self.windows.append((t0, t1))

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[4]),
                             [nme for nme, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass
        inwin = [r for r in rows if any(a - 0.05 <= r[0] <= b + 0.05 for a, b in self.windows)]
        load = [r for r in rows if r[3] > 0]
        use, which = (inwin, "timed regions") if len(inwin) >= 3 else ((load, "under load") if load else (rows, "all"))
        sm = [r[1] for r in use]
        reasons = sorted({x for r in use for x in r[4]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "samples": len(sm), "samples_total": len(rows), "over": which, "reasons": reasons}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_binary(kind, N, threads, iters, solve=True, level=1, div=2, timeout=3000):
    """oracle/_ref/parsy_ref = the reference's own headers compiled as they are (oracle/build_ref.sh)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "parsy_ref")
    if not os.path.exists(exe):
        return None
    cmd = [exe, "--kind", kind, "--N", str(N), "--cost", str(threads), "--level", str(level), "--div", str(div),
           "--threads", str(threads), "--iters", str(iters), "--no-dump-values", "--no-csc"]
    if not solve:
        cmd.append("--no-solve")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS="1")
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None
    if out.returncode != 0:
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


# LBC parameters of the reference's own evaluation script (scripts/eval.sh:5-17: costParam = threads, chunk = 1,
# levelParam in {2,1,0,-1,-2}, divRate in {2,4}; the best is reported), in the order they are tried under a wall budget
LBC_SWEEP = [(1, 2), (0, 2), (1, 4), (2, 2), (0, 4), (-1, 2), (2, 4), (-1, 4), (-2, 2), (-2, 4)]


def reference_sweep(kind, N, iters, budget_s, first_iters=None):
    """Runs the compiled reference over LBC_SWEEP while the wall budget lasts; every run is inspector + `iters`
    factorizations + the forward solves.  Returns (best run, list of (level, div, GFLOP/s, iters)) or (None, [])."""
    threads = host_threads()
    t_start = time.time()
    tried, best, best_gf, per_run = [], None, -1.0, None
    for k, (level, div) in enumerate(LBC_SWEEP):
        it = (first_iters or iters) if k == 0 else iters
        if per_run is not None and time.time() - t_start + per_run > budget_s:
            break
        t0 = time.time()
        r = run_reference_binary(kind, N, threads, it, solve=True, level=level, div=div,
                                 timeout=max(60.0, 4 * budget_s))
        if r is None or not r.get("factor_ok", 0):
            if k == 0:
                return None, []
            continue
        tf = r["t_factor_all"][1:] if len(r["t_factor_all"]) > 1 else r["t_factor_all"]
        gf = r["flops"] / float(np.mean(tf)) / 1e9
        r["_timed"], r["_level"], r["_div"], r["_gf"] = tf, level, div, gf
        tried.append({"levelParam": level, "divRate": div, "gflops": gf, "timed_factorizations": len(tf)})
        if gf > best_gf:
            best, best_gf = r, gf
        # the later runs time one factorization less often than the first: scale the estimate by the wall of this one
        per_run = (time.time() - t0) * (1.0 if k else (1 + iters) / (1 + it) if it != iters else 1.0)
    return best, tried


def cpu_baseline(kind, N, budget_s, iters=2):
    """Reference OpenMP executor on the host cores, bounded by a wall budget (falls back to the oracle's scalar C port
    on a 1/4-size grid if the compiled reference is absent)."""
    threads = host_threads()
    best, tried = reference_sweep(kind, N, iters, budget_s)
    if best is not None:
        t = float(np.mean(best["_timed"]))
        return {"value": best["_gf"], "unit": "GFLOP/s", "cores": threads, "kind": "reference",
                "sample": f"full workload; cholesky_left_par_05 with costParam={threads}, chunk=1 over the LBC triples of "
                          f"scripts/eval.sh that fit a {budget_s:.0f} s wall budget ({len(tried)} of {len(LBC_SWEEP)}), best "
                          f"reported: levelParam={best['_level']}, divRate={best['_div']}, mean of {len(best['_timed'])} "
                          f"factorization(s); inspector {best['t_inspector']:.1f} s per triple not counted",
                "lbc_sweep": tried, "factor_s": t, "levels_s": best["t_levels"], "last_level_s": best["t_last"],
                "sptrsv_ms": {"blockedLsolve": best["t_blockedLsolve"] * 1e3, "leveledBlockedLsolve": best["t_leveled"] * 1e3,
                              "H2LeveledBlockedLsolve": best["t_h2"] * 1e3,
                              "H2LeveledBlockedLsolve_Peeled": best["t_h2_peeled"] * 1e3}}
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


def workload_config(name):
    kind, N, desc = CONFIGS[name]
    n = N * N if kind == "2d5" else N ** 3
    return {"workload": desc, "name": name, "stencil": kind, "grid": N, "n": n}


def reference_arm(args, name):
    """The reference's own CPU implementation of the path on all host threads, under a wall budget: a 'step' is one
    factorization + forward solve of the full workload; as many of the requested steps as fit are timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, N, desc = CONFIGS[name]
    threads = host_threads()
    t_wall = time.time()
    budget = args.ref_budget
    iters = args.warmup + args.steps
    # size the first run from a quick probe of the factorization time: one warm + what fits a third of the budget
    probe = {"cfg1": 0.1, "cfg2": 1.0, "cfg4": 5.0, "cfg3": 55.0, "cfg5": 1500.0}[name]
    first = int(max(2, min(iters, 1 + (budget / 3.0) // probe)))
    later = int(max(1, min(args.steps, 3, (budget / 8.0) // probe)))     # the other triples: a few factorizations each
    best, tried = reference_sweep(kind, N, later, budget, first_iters=first)
    if best is None:
        base = cpu_baseline(kind, N, budget)
        val, ms, cb, sptrsv, eff = base["value"], base["factor_s"] * 1e3, base, None, 1
    else:
        tf = best["_timed"]
        th = best.get("t_h2_all", [best["t_h2"]])
        t, eff = float(np.mean(tf)), len(tf)
        val, ms = best["_gf"], (t + float(np.mean(th))) * 1e3
        sptrsv = {"H2LeveledBlockedLsolve": float(np.mean(th)) * 1e3, "blockedLsolve": best["t_blockedLsolve"] * 1e3,
                  "leveledBlockedLsolve": best["t_leveled"] * 1e3, "H2LeveledBlockedLsolve_Peeled": best["t_h2_peeled"] * 1e3}
        cb = {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "reference",
              "sample": f"full workload; {eff} timed factorization(s) + forward solves of the best LBC triple "
                        f"(levelParam={best['_level']}, divRate={best['_div']}, costParam={threads}) out of {len(tried)} "
                        f"tried under a {budget:.0f} s wall budget (scripts/eval.sh sweep order)",
              "lbc_sweep": tried}
    line = {"impl": "reference", "metric": "cholesky_factor_gflops", "value": val, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_effective": eff, "ms_per_step": ms,
            "step_note": "one step = cholesky_left_par_05 + H2LeveledBlockedLsolve (the reference has no backward sweep)",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(name),
            "executor": f"ParSy OpenMP cholesky_left_par_05 + H2LeveledBlockedLsolve on {threads} host threads, OpenBLAS 0.3.15",
            "cpu_baseline": cb, "sptrsv_ms": sptrsv,
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t_wall}
    emit(line)
    return 0


class DevArray:   # zero-copy torch view of a device buffer owned by the library
    def __init__(self, p, cnt):
        self.__cuda_array_interface__ = {"shape": (int(cnt),), "typestr": "<f8", "data": (int(p), False), "version": 3}


def traffic_for(name):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS config, or None."""
    for f in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", f)) as fh:
                t = json.load(fh).get(name)
            if t:
                src = t.get("source", f)
                if t.get("launch_flops"):
                    src += (f" [captured launch: {t['launch_flops'] / 1e9:.1f} GFLOP in {t.get('launch_ms')} ms, "
                            f"{t['dram_bytes_per_launch'] / t['launch_flops'] * 1e3:.2f} DRAM bytes per kFLOP]")
                return t.get("dram_bytes_per_launch"), src
        except Exception:
            pass
    return None, None


def sharded_arm(args, name, rank, world, local, sampler):
    """N > 1: ONE factorization + solve sharded over the GPUs (strong scaling), all kernels and NCCL collectives inside
    libparsy_cuda's captured graphs."""
    import torch
    import torch.distributed as dist
    from parsy_bench_b200 import executor as ex, inspector, matrices
    from parsy_bench_b200.sharded import make_sharded

    kind, N, desc = CONFIGS[name]
    dev = torch.device("cuda", local)
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    t0 = time.time()
    S = inspector.analyze(n, Ap, Ai, Ax, args.cost, args.level, args.div)
    t_insp = time.time() - t0
    t0 = time.time()
    SH = make_sharded(S, rank, world, local, dist, top_levels=args.top_levels, block_cols=args.block_cols,
                      top_distributed=not args.replicate_top, lookahead=not args.no_lookahead, top_chunk=args.top_chunk)
    t_create = time.time() - t0
    F = S.flops
    ptr = SH.device_pointers()
    stream = torch.cuda.ExternalStream(ptr["stream"], device=dev)
    d_rhs = torch.as_tensor(DevArray(ptr["rhs"], n), device=dev)
    d_lv = torch.as_tensor(DevArray(ptr["factor"], S.xsize), device=dev)
    h_vals = torch.from_numpy(S.A2_x.copy()).pin_memory()
    b_host = (1.0 + np.arange(n) / n)[S.Perm]
    h_b = torch.from_numpy(b_host.copy()).pin_memory()
    h_x = torch.empty(n, dtype=torch.float64).pin_memory()
    d_b = torch.from_numpy(b_host).to(dev)
    SH.set_values(h_vals.numpy())
    st = SH.stats()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def resident_step(ev=None):
        with torch.cuda.stream(stream):
            d_rhs.copy_(d_b, non_blocking=True)
            if ev: ev[0].record(stream)
            SH.factor()
            if ev: ev[1].record(stream)
            SH.solve(ex.SOLVE_FWD)
            if ev: ev[2].record(stream)
            SH.solve(ex.SOLVE_BWD)
            if ev: ev[3].record(stream)

    for _ in range(args.warmup):
        resident_step()
    barrier()
    if not SH.sync():
        raise SystemExit("factorization failed: matrix not positive definite")
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = np.zeros(3)
    barrier()
    tm0 = time.time()
    e0.record(stream)
    for k in range(args.steps):
        resident_step(evs[k])
    e1.record(stream)
    barrier()
    sampler.mark(tm0, time.time())
    total_ms = e0.elapsed_time(e1)
    fac_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    fwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    pt = SH.phase_times()
    phase = np.array([pt["phase1"], pt["sum_top"], pt["top"]]) * 1e3
    x_res = d_rhs.cpu().numpy().copy()

    # end to end through the public API with host buffers: A's values and b from pinned memory, x back, every step
    def e2e_step():
        SH.set_values(h_vals.numpy())
        SH.set_rhs(h_b.numpy(), sync=False)
        SH.factor()
        SH.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
        SH.get_rhs(h_x.numpy())
    for _ in range(2):
        e2e_step()
    barrier()
    tm0 = time.time()
    tw = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - tw) * 1e3 / args.steps
    sampler.mark(tm0, time.time())
    clocks = sampler.stop() if rank == 0 else None
    ok = SH.sync()
    # checks outside every timed region: ||L||_F^2 = trace(A) over the panels this rank holds complete, summed over the
    # ranks (top counted once), and the residual of the sharded solve
    import scipy.sparse as sp
    own = SH.plan(1).owned_ranges(rank)
    fro = sum(float((d_lv[int(b):int(e)] ** 2).sum().item()) for b, e in own)
    if rank == 0:
        fro += sum(float((d_lv[int(b):int(e)] ** 2).sum().item()) for b, e in SH.plan(1).owned_ranges(-1))
    A2 = sp.csc_matrix((S.A2_x, S.A2_i, S.A2_p), shape=(n, n))
    res = float(np.linalg.norm(A2 @ x_res + sp.tril(A2, -1).T @ x_res - b_host) / np.linalg.norm(b_host))
    t_loc = torch.tensor([total_ms, fac_ms, fwd_ms, bwd_ms, e2e_ms, *phase.tolist()], dtype=torch.float64, device="cuda")
    s_loc = torch.tensor([fro], dtype=torch.float64, device="cuda")
    dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
    dist.all_reduce(s_loc, op=dist.ReduceOp.SUM)
    total_ms, fac_ms, fwd_ms, bwd_ms, e2e_ms, p1_ms, sum_ms, top_ms = [float(v) for v in t_loc.tolist()]
    trace = float(S.A2_x[S.A2_p[:-1]].sum())
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        tf_per_gpu = F / (fac_ms * 1e-3) / 1e12 / world
        line = {
            "metric": "cholesky_factor_gflops", "value": F / (fac_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name),
            "details": {"nnzL": int(S.xsize), "flops_sum_cc2": F,
                        "lbc": {"costParam": args.cost, "levelParam": args.level, "divRate": args.div,
                                "hlevels": int(S.nLevels), "wpartitions": int(S.nParts)},
                        "step": "zero + scatter A + sharded factor + sharded forward + backward sweep, resident",
                        "parallelism": f"ONE factorization over {world} GPUs: bottom subtrees by owner with fan-in of their "
                                       f"updates into the top ({st['nccl_allreduces']} ncclAllReduce, "
                                       f"{st['bytes_summed'] / 1e9:.2f} GB), {args.top_levels} top H-level(s) "
                                       + (f"block-cyclic by block column, owner factors + {st['nccl_broadcasts']} "
                                          f"ncclBroadcast ({st['bytes_broadcast'] / 1e9:.2f} GB) in the captured graph"
                                          if not args.replicate_top else "replicated")
                                       + "; solve sharded the same way (2 all-reduces)",
                        "nccl_version": st["nccl_version"], "top_chain_steps": st["top_chain_steps"],
                        "l2": f"working set {8 * S.xsize / 1e6:.0f} MB (factor) > 126 MB L2, no flush needed"},
            "breakdown_ms": {"factor": fac_ms, "fwd_solve": fwd_ms, "bwd_solve": bwd_ms,
                             "phase1_owned_subtrees_max_rank": p1_ms, "sum_top_panels_max_rank": sum_ms,
                             "distributed_top_max_rank": top_ms,
                             "allreduce_bus_gbs": 2 * (world - 1) / world * st["bytes_summed"] / (sum_ms * 1e-3) / 1e9 if sum_ms > 0 else None},
            "sptrsv_ms": {"forward": fwd_ms, "backward": bwd_ms},
            "e2e": {"value": F / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * S.nnzA + 8 * n), "d2h_bytes_per_step": int(8 * n),
                    "api": "Sharded.set_values + set_rhs (pinned, every rank) + factor + solve(FWD|BWD) + get_rhs (x to the host)"},
            "gpu_launches": int(args.steps * (st["launches_factor"] + st["launches_fwd"] + st["launches_bwd"])),
            "roofline": {"bound": "tensor", "kernel": "whole sharded factorization (dominant kernel k_gemm_tiles, FP64 DMMA)",
                         "achieved": tf_per_gpu, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s per GPU",
                         "frac": tf_per_gpu / FP64_PEAK_TFLOPS, "traffic": None,
                         "note": "algorithmic flops (sum cc^2) / max-over-ranks factor time / N; the single-GPU line carries "
                                 "the per-kernel roofline"},
            "cpu_baseline": None, "clocks": clocks, "fro2_over_trace": float(s_loc.item()) / trace, "residual": res,
            "factor_ok": bool(ok), "setup_s": {"inspector": t_insp, "create": t_create},
            "device_bytes_rank0": st["device_bytes"],
        }
        emit(line)
    SH.close()
    dist.destroy_process_group()
    return 0


def run_single(args, name, local, steps, warmup, sampler, with_dropin=True):
    """One GPU, one config: resident steps (CUDA events), end-to-end steps (host buffers), drop-in call, rooflines."""
    import torch
    from parsy_bench_b200 import executor as ex, inspector, matrices
    kind, N, desc = CONFIGS[name]
    dev = torch.device("cuda", local)
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
    h_vals = torch.from_numpy(S.A2_x.copy()).pin_memory()
    b_host = (1.0 + np.arange(n) / n)[S.Perm]          # b_i = 1 + i/n in the original ordering, permuted
    h_b = torch.from_numpy(b_host.copy()).pin_memory()
    h_x = torch.empty(n, dtype=torch.float64).pin_memory()
    ptr = H.device_pointers()
    stream = torch.cuda.ExternalStream(ptr["stream"], device=dev)
    d_rhs = torch.as_tensor(DevArray(ptr["rhs"], n), device=dev)
    d_b = torch.from_numpy(b_host).to(dev)
    H.set_values(h_vals.numpy())

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

    for _ in range(warmup):
        resident_step()
    torch.cuda.synchronize()
    if not H.sync():
        raise SystemExit("factorization failed: matrix not positive definite")
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    tm0 = time.time()
    e0.record(stream)
    for k in range(steps):
        resident_step(evs[k])
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.mark(tm0, time.time())
    total_ms = e0.elapsed_time(e1)
    fac_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    fwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    bwd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    times = H.factor_times()

    def e2e_step():
        H.set_values(h_vals.numpy())       # H2D nnz(A) doubles (pinned)
        H.set_rhs(h_b.numpy())             # H2D n doubles
        H.factor()
        H.solve(ex.SOLVE_FWD | ex.SOLVE_BWD)
        H.get_rhs(h_x.numpy())             # D2H n doubles (synchronises)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    tm0 = time.time()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    sampler.mark(tm0, time.time())
    x = h_x.numpy().copy()
    import scipy.sparse as sp
    A2 = sp.csc_matrix((S.A2_x, S.A2_i, S.A2_p), shape=(n, n))
    res = float(np.linalg.norm(A2 @ x + sp.tril(A2, -1).T @ x - b_host) / np.linalg.norm(b_host))

    # strict drop-in calls (reference signature, host arrays, factor downloaded): the reference's driver calls the
    # executor five times on one structure (examples/choleskyTest01.cpp:199-222) — first call cold, the rest warm
    dropin = None
    if with_dropin:
        lv = np.empty(S.xsize)
        tm = np.zeros(4)
        ts = []
        ok = True
        for _ in range(3):
            t0 = time.perf_counter()
            ok = ok and ex.cholesky_left_par_05(n, S.A2_p, S.A2_i, S.A2_x, S.p, S.s, S.i_ptr, lv, S.super, S.nsuper, tm,
                                                S.sParent, S.A1_p, S.A1_i, S.col2Sup, S.nLevels, S.levelPtr, None, 0,
                                                S.parPtr, S.partition)
            ts.append(time.perf_counter() - t0)
        dropin = {"cold_s": ts[0], "warm_s": min(ts[1:]), "ok": bool(ok), "d2h_bytes": int(8 * S.xsize),
                  "note": "parsy_cuda_cholesky_left_par_05 with host arrays: plan + upload + factor + download of L; "
                          "warm = structure cached by content hash"}
        del lv

    # column (CSC) forward solve on the same factor (Triangular_CSC.h:76, lsolveParH2): the LBC schedule expanded to
    # columns, one dataflow launch; algorithmic traffic 12 B per stored entry + 16 B per column
    csc = None
    if S.xsize < 3e8:
        Cp, Ci, Cx = S.bcsc2csc(H.get_factor())
        order = np.concatenate([np.arange(S.super[sn], S.super[sn + 1], dtype=np.int32) for sn in S.partition])
        C = ex.CscSolver(n, Cp, Ci, order=order, device=local)
        C.set_values(Cx)
        ms = []
        for _ in range(5):
            xs = b_host.copy()
            ms.append(C.solve(xs))
        C.close()
        cbytes = 12.0 * len(Ci) + 16.0 * n
        peaks0, _ = measured_peaks()
        hbm0 = float(peaks0.get("hbm_gbs", HBM_FALLBACK_GBS))
        csc = {"lsolveParH2_ms": float(np.median(ms)), "nnz": int(len(Ci)), "algorithmic_bytes": cbytes,
               "achieved_gbs": cbytes / (float(np.median(ms)) * 1e-3) / 1e9,
               "hbm_frac": cbytes / (float(np.median(ms)) * 1e-3) / 1e9 / hbm0,
               "max_abs_diff_vs_supernodal_forward": None}
        del Cp, Ci, Cx

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
    traffic, traffic_src = traffic_for(name)
    roofline = {"bound": "tensor", "kernel": "k_gemm_tiles (FP64 DMMA m8n8k4: SYRK/GEMM update + TRSM)",
                "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved_tf / FP64_PEAK_TFLOPS, "traffic": traffic, "traffic_source": traffic_src,
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
    H.close()
    return {
        "config": workload_config(name), "value": F / (fac_ms * 1e-3) / 1e9, "ms_per_step": total_ms / steps,
        "steps": steps, "warmup": warmup,
        "details": {"nnzA": int(S.nnzA), "nsuper": int(S.nsuper), "nnzL": int(S.xsize), "flops_sum_cc2": F,
                    "flops_supernodal": F_sn,
                    "lbc": {"costParam": args.cost, "levelParam": args.level, "divRate": args.div,
                            "hlevels": int(S.nLevels), "wpartitions": int(S.nParts)},
                    "step": "zero L + scatter A + factor (all LBC levels) + forward sweep + backward sweep, resident",
                    "l2": f"working set {8 * S.xsize / 1e6:.0f} MB (factor) > 126 MB L2, no flush needed",
                    "parallelism": "single GPU"},
        "breakdown_ms": {"factor": fac_ms, "fwd_solve": fwd_ms, "bwd_solve": bwd_ms,
                         "factor_levels": times["levels"] * 1e3, "factor_last_level": times["last_level"] * 1e3,
                         "assemble": times["assemble"] * 1e3},
        "sptrsv_ms": {"forward": fwd_ms, "backward": bwd_ms}, "sptrsv_csc": csc,
        "e2e": {"value": F / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(8 * S.nnzA + 8 * n), "d2h_bytes_per_step": int(8 * n),
                "api": "Solver.set_values + set_rhs + factor + solve(FWD|BWD) + get_rhs, pinned host buffers",
                "dropin_cholesky_left_par_05": dropin},
        "gpu_launches": int(steps * (st["launches_factor"] + st["launches_fwd"] + st["launches_bwd"])),
        "roofline": roofline, "roofline_solve": roofline_solve, "kernel_classes": prof, "residual": res,
        "setup_s": {"inspector": t_insp, "inspector_metis": S.t_ordering, "create": t_create},
        "device_bytes": st["device_bytes"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 10 (cfg3), 100 (cfg1, cfg2), 20 (cfg4), 2 (cfg5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="default: cfg3 (the largest single-GPU configuration BASELINE.json names; N > 1: sharded), with "
                         "cfg2 — the configuration the metric is quoted on — as a sub-record of the same line at N = 1")
    ap.add_argument("--cost", type=int, default=592, help="LBC innerParts handed to the inspector (GPU choice)")
    ap.add_argument("--div", type=int, default=4)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--block-cols", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-record", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=75.0, help="wall seconds for the cpu_baseline leg per config")
    ap.add_argument("--ref-budget", type=float, default=420.0, help="wall seconds for --impl reference")
    ap.add_argument("--no-lookahead", action="store_true", help="single stream, no overlap of POTRF/TRSM with the bulk updates")
    ap.add_argument("--ignore-hlevels", action="store_true", help="schedule by dependencies only (no LBC H-level barriers)")
    ap.add_argument("--no-fan-out", action="store_true", help="kernel classes of a step on one stream (A/B)")
    ap.add_argument("--general-sweeps", action="store_true", help="leaf region of the sweeps on the general dataflow kernel (A/B)")
    ap.add_argument("--replicate-top", action="store_true", help="N>1: every rank computes the top separators")
    ap.add_argument("--top-chunk", type=int, default=0, help="N>1: consecutive top block columns per owner (0 = library default)")
    ap.add_argument("--top-levels", type=int, default=1, help="N>1: LBC H-levels whose supernodes are shared (distributed by block column)")
    args = ap.parse_args()
    quiet_stdout()
    default_config = args.config is None
    name = args.config or "cfg3"
    if args.steps is None:     # long enough a timed region for the clock sampler, short enough to end within a minute
        args.steps = {"cfg1": 100, "cfg2": 100, "cfg4": 20, "cfg3": 10, "cfg5": 2}[name]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args, name)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the executor has no CPU fallback")
    torch.cuda.set_device(local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # before warm-up: a short timed region must not end up with zero samples
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return sharded_arm(args, name, rank, world, local, sampler)

    rec = run_single(args, name, local, args.steps, args.warmup, sampler)
    sub = None
    if default_config and not args.no_sub_record:
        # BASELINE.json configs[1] — the configuration the metric is quoted on (a 15-GFLOP, latency-bound problem)
        sub = run_single(args, "cfg2", local, max(args.steps, 50), args.warmup, sampler)
    clocks = sampler.stop()
    cb = None
    if not args.no_cpu_baseline:
        kind, N, _ = CONFIGS[name]
        cb = cpu_baseline(kind, N, args.cpu_budget, iters=1 if name in ("cfg3", "cfg5") else 2)
        if sub is not None:
            sub["cpu_baseline"] = cpu_baseline("2d5", 1000, min(args.cpu_budget, 40.0))
    line = {
        "metric": "cholesky_factor_gflops", "value": rec["value"], "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "strong",    # the --gpus series keeps the total work fixed: one factorization + solve of the same matrix
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": rec["config"], "details": rec["details"],
        "breakdown_ms": rec["breakdown_ms"], "sptrsv_ms": rec["sptrsv_ms"], "sptrsv_csc": rec["sptrsv_csc"], "e2e": rec["e2e"],
        "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"], "roofline_solve": rec["roofline_solve"],
        "kernel_classes": rec["kernel_classes"], "cpu_baseline": cb, "clocks": clocks, "residual": rec["residual"],
        "setup_s": rec["setup_s"], "device_bytes": rec["device_bytes"],
    }
    if sub is not None:
        line["sub_records"] = {"cfg2": sub}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
