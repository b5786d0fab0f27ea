"""Digests of the executor plan (parsy_cuda_plan_digest) for a set of matrices, schedules and sharding options.
Host only.  `python tools/plan_digests.py [--large] [--json out.json]` — used to check that a planner change leaves
the device work untouched (compare the table before and after) and to (re)generate tests/golden/plan_digests.json."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsy_bench_b200 import executor as ex, inspector, matrices  # noqa: E402

SMALL = [("2d5", 60, 64, 1, 2), ("2d5", 150, 8, 1, 2), ("3d7", 12, 16, 0, 2), ("3d7", 24, 592, 1, 4),
         ("3d27", 10, 16, 1, 2), ("3d27", 16, 16, 0, 2), ("2d5", 200, 64, 1, 3), ("3d7", 30, 64, 1, 2),
         ("2d5+dag", 120, 16, 1, 2), ("3d27+dag", 16, 16, 0, 2), ("rnd", 6000, 16, 1, 2)]
MEDIUM = [("2d5", 300, 64, 1, 2), ("3d7", 40, 64, 1, 2)]      # large enough for the planner to split its lists over threads
LARGE = [("2d5", 1000, 64, 1, 2), ("3d7", 60, 64, 1, 2), ("3d27", 40, 64, 1, 2)]
SHARD = [(0, 1, 0, 1, 0), (0, 2, 1, 1, 0), (1, 2, 1, 1, 0), (0, 2, 2, 1, 0), (1, 2, 2, 1, 0), (3, 8, 1, 1, 0),
         (3, 8, 2, 1, 0), (5, 8, 2, 2, 2), (1, 4, 1, 2, 0), (2, 4, 2, 2, 1)]      # rank, world, phase, top_levels, top_chunk


def random_pattern(n, extra_per_col, seed):
    """Lower half of a symmetric, strictly diagonally dominant matrix with a random pattern plus a path."""
    import numpy as np
    rng = np.random.default_rng(seed)
    cols = [sorted({j + 1} | {int(i) for i in rng.integers(j + 1, n, size=extra_per_col)}) for j in range(n - 1)] + [[]]
    deg = np.zeros(n)
    for j, rows in enumerate(cols):
        deg[j] += len(rows)
        for i in rows:
            deg[i] += 1
    Ap, Ai, Ax = [0], [], []
    for j, rows in enumerate(cols):
        Ai += [j] + rows
        Ax += [deg[j] + 1.0] + [-1.0] * len(rows)
        Ap.append(len(Ai))
    return n, np.array(Ap, np.int32), np.array(Ai, np.int32), np.array(Ax)


def table(cases):
    out = {}
    for kind, N, cost, level, div in cases:
        if kind == "rnd":
            n, Ap, Ai, Ax = random_pattern(N, 2, 11)
        else:
            n, Ap, Ai, Ax = matrices.laplacian(kind.split("+")[0], N)
        S = inspector.analyze(n, Ap, Ai, Ax, cost, level, div)
        args = (n, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.col2Sup, S.nLevels, S.levelPtr, S.parPtr, S.partition)
        if kind.endswith("+dag"):      # the DAG-based LBC schedule over the factor's blocks instead of the tree-based one
            nl, lp, pp, part = inspector.dag_lbc_bcsc(S, cost, level, div)
            args = args[:7] + (nl, lp, pp, part)
        for rank, world, phase, tl, tc in SHARD:
            for nb, ign in ((0, False), (64, False), (0, True)):
                if world > 1 and (nb or ign):
                    continue
                key = f"{kind}:{N}:{cost}:{level}:{div}|r{rank}w{world}p{phase}t{tl}c{tc}|nb{nb}i{int(ign)}"
                t = time.time()
                try:
                    d = ex.plan_digest(*args, block_cols=nb, ignore_hlevels=ign, rank=rank, world=world, phase=phase,
                                       top_levels=tl, top_chunk=tc)
                    out[key] = f"{d:016x}"
                except ex.ParsyCudaError as e:
                    out[key] = f"error {e.code}"
                if os.environ.get("PLAN_DIGEST_TIMES"):
                    print(f"{key}  {out[key]}  {1e3 * (time.time() - t):.1f} ms", file=sys.stderr)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--medium", action="store_true", help="only the MEDIUM cases (thread-independence check)")
    ap.add_argument("--json")
    a = ap.parse_args()
    T = table(MEDIUM if a.medium else SMALL + (LARGE if a.large else []))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(T, f, indent=0, sort_keys=True)
    else:
        for k in sorted(T):
            print(k, T[k])
