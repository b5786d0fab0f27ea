"""Dev tool (torchrun, N GPUs): one inspector run, then the sharded factorization under several (top_chunk, top_levels,
lookahead) settings; prints max-over-ranks phase times.  python -m torch.distributed.run --nproc-per-node N tools/shard_sweep.py cfg3 4:1 1:1 8:1"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CONFIGS  # noqa: E402
from parsy_bench_b200 import executor as ex, inspector, matrices  # noqa: E402
from parsy_bench_b200.sharded import make_sharded  # noqa: E402


def main():
    name = sys.argv[1]
    combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(4, 1)]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, N, _ = CONFIGS[name]
    n, Ap, Ai, Ax = matrices.laplacian(kind, N)
    S = inspector.analyze(n, Ap, Ai, Ax, 592, 1, 4)
    for combo in combos:
        chunk, top = combo[0], combo[1]
        la = combo[2] if len(combo) > 2 else 1
        lanes = combo[3] if len(combo) > 3 else 0
        t0 = time.time()
        SH = make_sharded(S, rank, world, local, dist, top_levels=top, top_chunk=chunk, lookahead=bool(la), lanes=lanes)
        tc = time.time() - t0
        SH.set_values(S.A2_x)
        for _ in range(2):
            SH.factor()
        ok = SH.sync()
        dist.barrier()
        ts = []
        for _ in range(4):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            SH.factor()
            SH.sync()
            t1 = time.perf_counter()
            pt = SH.phase_times()
            ts.append([(t1 - t0) * 1e3, pt["phase1"] * 1e3, pt["sum_top"] * 1e3, pt["top"] * 1e3])
        t = torch.tensor(np.median(np.array(ts), axis=0), device="cuda")
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            st = SH.stats()
            print(f"{name} N={world} chunk={chunk} top_levels={top} lookahead={la} lanes={lanes}: wall {t[0]:.1f} ms  phase1 {tmin[1]:.1f}..{t[1]:.1f}  sum {tmin[2]:.1f}..{t[2]:.1f}  "
                  f"top {tmin[3]:.1f}..{t[3]:.1f}  steps {st['top_chain_steps']} bcasts {st['nccl_broadcasts']} ok={ok} create {tc:.1f}s",
                  flush=True)
        if os.environ.get("PARSY_TRACE_TOP") and combo == combos[0]:
            tms, own = SH.trace_top()
            SH.sync()
            np.savez(os.path.join(ROOT, "gpurun_out", f"top_trace_{name}_N{world}_r{rank}.npz"), t=tms, owner=own)
        SH.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
