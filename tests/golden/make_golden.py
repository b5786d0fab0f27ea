"""Generates the committed golden vectors from the compiled reference (oracle/_ref/parsy_ref, built from
/root/reference by oracle/build_ref.sh).  Run in the build container only:  python tests/golden/make_golden.py

Small cases are stored whole (every symbolic / schedule / numeric array the reference produces); the cfg1-size
case is stored as a digest (sizes, SHA-256 of every integer array, ||L||_F^2, sampled factor entries)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refdump import ref_case  # noqa: E402

FULL = [("2d5", 12, 8, 1, 2), ("2d5", 30, 8, 1, 2), ("2d5", 30, 64, 0, 4), ("3d7", 7, 8, 1, 2), ("3d27", 6, 4, 0, 2)]
MTX = {"mtx_rand150_c8_l1_d2": ("rand150.mtx", 8, 1, 2)}   # file written by write_rand_mtx() below
DIGEST = [("2d5", 100, 8, 1, 2), ("2d5", 100, 592, 1, 4), ("3d7", 20, 8, 1, 2), ("3d27", 16, 8, 1, 2)]
# BASELINE.json configs 2, 3 and 4 at full size: inspector only (sizes + SHA-256 of every integer array), with the
# LBC triple bench.py hands to the GPU executor and, for cfg2 / cfg4, the reference's CPU convention as well.  These
# are the sizes where the int-overflow / INT_MAX paths of SURVEY.md App. B.8 and B.10 bite.
DIGEST_LARGE = [("2d5", 1000, 592, 1, 4), ("2d5", 1000, 8, 1, 2), ("3d27", 64, 592, 1, 4), ("3d27", 64, 8, 1, 2),
                ("3d7", 100, 592, 1, 4)]
INT_ARRAYS = ["Perm", "ColCount", "super", "sParent", "col2Sup", "pi", "s", "p", "i_ptr", "levelPtr", "parPtr",
              "partition", "A2_p", "A2_i", "A1_p", "A1_i", "etree_levelPtr", "etree_levelSet"]


def name(c):
    return f"{c[0]}_N{c[1]}_c{c[2]}_l{c[3]}_d{c[4]}".replace("-", "m")


def write_rand_mtx():
    """A random-pattern, strictly diagonally dominant SPD matrix (not a stencil) as a lower-half Matrix-Market file."""
    from test_mmio import random_spd_lower
    from parsy_bench_b200 import matrices
    n, Ap, Ai, Ax = random_spd_lower(150, 2, seed=5)
    matrices.write_mtx(os.path.join(HERE, "rand150.mtx"), n, Ap, Ai, Ax, comment="random_spd_lower(150, 2, seed=5)")


def large_digests(path):
    """digests_large.json: run with `python tests/golden/make_golden.py --large` (cfg3 alone is ~35 s of the reference
    inspector and a few GB of RAM)."""
    dig = {}
    ints = [k for k in INT_ARRAYS if not k.startswith("etree_")]
    for c in DIGEST_LARGE:
        R = ref_case(c[0], c[1], cost=c[2], level=c[3], div=c[4], threads=1, factor=False, solve=False)
        m = R.meta
        d = {k: m[k] for k in ("n", "nnzA", "nsuper", "xsize", "ssize", "nLevels", "nParts", "maxSupWid", "maxCol", "flops")}
        R["s"] = R["s"][:m["ssize"]]
        d["sha256"] = {k: hashlib.sha256(np.ascontiguousarray(R[k]).tobytes()).hexdigest() for k in ints}
        dig[name(c)] = d
        print("digest", name(c), flush=True)
        with open(path, "w") as f:
            json.dump(dig, f, indent=1)


def main():
    if "--large" in sys.argv:
        return large_digests(os.path.join(HERE, "digests_large.json"))
    write_rand_mtx()
    for nm, (f, c, l, d) in MTX.items():
        R = ref_case("2d5", 0, cost=c, level=l, div=d, threads=1, mtx=os.path.join(HERE, f))
        arrays = {k: v for k, v in R.items() if isinstance(v, np.ndarray)}
        arrays["s"] = arrays["s"][:R.meta["ssize"]]
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), meta=json.dumps(R.meta), **arrays)
        print("wrote", nm, {k: R.meta[k] for k in ("n", "nsuper", "xsize")})
    for c in FULL:
        R = ref_case(c[0], c[1], cost=c[2], level=c[3], div=c[4], threads=1)
        arrays = {k: v for k, v in R.items() if isinstance(v, np.ndarray)}
        arrays["s"] = arrays["s"][:R.meta["ssize"]]
        np.savez_compressed(os.path.join(HERE, name(c) + ".npz"), meta=json.dumps(R.meta), **arrays)
        print("wrote", name(c), {k: R.meta[k] for k in ("n", "nsuper", "xsize")})
    dig = {}
    for c in DIGEST:
        R = ref_case(c[0], c[1], cost=c[2], level=c[3], div=c[4], threads=1)
        m = R.meta
        d = {k: m[k] for k in ("n", "nnzA", "nsuper", "xsize", "ssize", "nLevels", "nParts", "maxSupWid", "maxCol",
                                "flops", "fro2", "trace", "etree_levels", "nnzLcsc")}
        R["s"] = R["s"][:m["ssize"]]
        d["sha256"] = {k: hashlib.sha256(np.ascontiguousarray(R[k]).tobytes()).hexdigest() for k in INT_ARRAYS}
        idx = np.linspace(0, m["xsize"] - 1, 257).astype(np.int64)
        d["valL_sample_idx"] = idx.tolist()
        d["valL_sample"] = R.valL[idx].tolist()
        d["y_ramp_sample"] = R.y_ramp[:: max(1, m["n"] // 64)].tolist()
        dig[name(c)] = d
        print("digest", name(c))
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(dig, f, indent=1)


if __name__ == "__main__":
    main()
