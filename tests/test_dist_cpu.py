"""CPU, world_size 2 over gloo: host logic of the sharded factorization (ownership split, fan-in sum of the top panels,
every supernode factored once, every update applied once).  The numeric kernels need a GPU; here the summed "panels"
are stand-in values."""
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
        top = ex.plan_owned_ranges(*args, world, -1, top_levels=2)
        # ownership: the ranks' subtrees and the shared top partition the panels
        owned = torch.zeros(S.xsize, dtype=torch.int32)
        for runs in ranges + [top]:
            for b, e in runs:
                owned[int(b):int(e)] += 1
        assert int(owned.min()) == 1 and int(owned.max()) == 1
        # fan-in over gloo with stand-in numbers: every rank adds "its subtrees' contribution" into its own copy of the
        # top panels, rank 0 alone carries A's entries; the all-reduce must deliver A_top - sum of all contributions
        lv = torch.zeros(S.xsize, dtype=torch.float64)
        for b, e in top:
            if rank == 0:
                lv[int(b):int(e)] = 1000.0
            lv[int(b):int(e)] -= float(rank + 1)
        for b, e in top:
            dist.all_reduce(lv[int(b):int(e)])
        want = 1000.0 - sum(range(1, world + 1))
        for b, e in top:
            assert bool((lv[int(b):int(e)] == want).all())
        # plans: phase 1 of rank r factors exactly the supernodes r owns; phase 2 lists every top supernode on every
        # rank (the sweeps need them all) and, between the ranks, every update runs exactly once: the flops add up
        rc, st = ex.plan_check(*args, rank=rank, world=world, phase=2, top_levels=2)
        assert rc == ex.OK
        top_sup = st["reserved"][3]
        assert st["reserved"][0] == top_sup
        rc1, st1 = ex.plan_check(*args, rank=rank, world=world, phase=1, top_levels=2)
        assert rc1 == ex.OK and st1["reserved"][0] == st1["reserved"][2]
        mine = torch.tensor([st1["reserved"][0], st1["reserved"][4] + st["reserved"][4]], dtype=torch.int64)
        tot = mine.clone()
        dist.all_reduce(tot)
        full = ex.plan_check(*args)[1]
        assert int(tot[0]) + top_sup == S.nsuper           # every bottom supernode is factored exactly once
        assert abs(int(tot[1]) - full["reserved"][4]) <= 4 * world
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
