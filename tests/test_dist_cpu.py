"""CPU, world_size 2 over gloo: host logic of the sharded factorization (ownership split + panel exchange
bookkeeping).  The numeric kernels need a GPU; here the exchanged "panels" are stand-in values."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parsy_bench_b200 import executor as ex, inspector, matrices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, Ap, Ai, Ax = matrices.laplacian("2d5", 60)
        S = inspector.analyze(n, Ap, Ai, Ax, 64, 1, 2)
        args = (n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
        ranges = [ex.plan_owned_ranges(*args, world, r, top_levels=2) for r in range(world)]
        # stand-in factor: every rank only knows the panels it owns
        lv = torch.full((S.xsize,), float("nan"), dtype=torch.float64)
        for b, e in ranges[rank]:
            lv[b:e] = torch.arange(b, e, dtype=torch.float64) * (rank + 1)
        for owner, runs in enumerate(ranges):
            for b, e in runs:
                dist.broadcast(lv[int(b):int(e)], src=owner)
        # after the exchange every rank holds every owner's values
        for owner, runs in enumerate(ranges):
            for b, e in runs:
                assert torch.equal(lv[int(b):int(e)], torch.arange(int(b), int(e), dtype=torch.float64) * (owner + 1))
        owned = torch.zeros(S.xsize, dtype=torch.int32)
        for runs in ranges:
            for b, e in runs:
                owned[int(b):int(e)] += 1
        assert int(owned.max()) == 1                       # disjoint
        # what nobody owns is exactly the shared top (still NaN), and its supernode count matches the planner
        rc, st = ex.plan_check(*args, rank=rank, world=world, phase=2, top_levels=2)
        assert rc == ex.OK
        top_sup = st["reserved"][3]
        assert st["reserved"][0] == top_sup                # every rank factors every top supernode (POTRF/TRSM replicated)
        rc1, st1 = ex.plan_check(*args, rank=rank, world=world, phase=1, top_levels=2)
        mine = torch.tensor([st1["reserved"][0], st1["reserved"][4] + st["reserved"][4]], dtype=torch.int64)
        tot = mine.clone()
        dist.all_reduce(tot)
        full = ex.plan_check(*args)[1]
        assert int(tot[0]) + top_sup == S.nsuper           # every bottom supernode is factored exactly once
        # every update (descendant pair or trailing block update) runs on exactly one rank: the flops add up
        assert abs(int(tot[1]) - full["reserved"][4]) <= 4 * world
        assert bool(torch.isnan(lv[owned == 0]).all()) and not bool(torch.isnan(lv[owned == 1]).any())
        # checksum agreement across ranks
        chk = torch.nan_to_num(lv).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi)
        q.put((rank, "ok", int(mine[0])))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc), 0))
    finally:
        dist.destroy_process_group()


def test_sharded_ownership_and_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    counts = sorted(r[2] for r in res)
    assert counts[0] > 0 and counts[1] < 2.5 * counts[0]      # both ranks got work, roughly balanced


def test_single_rank_plan_is_unsharded():
    n, Ap, Ai, Ax = matrices.laplacian("3d7", 6)
    S = inspector.analyze(n, Ap, Ai, Ax)
    args = (n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
    r = ex.plan_owned_ranges(*args, 1, 0)
    assert r.shape == (1, 2) and r[0, 0] == 0 and r[0, 1] == S.xsize
