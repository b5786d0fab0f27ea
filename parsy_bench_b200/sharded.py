"""Sharded factorization + solve across the GPUs of one node under torch.distributed (one process per GPU).

Everything that moves or computes lives in libparsy_cuda (include/parsy_cuda.h section 3: kernels and NCCL collectives
captured into the same CUDA graphs); torch.distributed is used once, to hand rank 0's NCCL unique id to the other ranks.
"""
import numpy as np

from . import executor as ex


def make_sharded(S, rank, world, device, dist=None, **kw):
    """S: inspector.Symbolic.  ``dist``: an initialised torch.distributed module (any backend) or None to emulate all
    ranks in this process."""
    uid = None
    if dist is not None and world > 1:
        box = [ex.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return ex.Sharded(S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels,
                      S.levelPtr, S.parPtr, S.partition, rank, world, uid, device=device, **kw)


def merge_factor(sh, dist, xsize):
    """Full factor on every rank from the ranks' pieces (tests / verification; moves the whole factor through the host)."""
    lv = sh.get_factor(np.zeros(xsize))
    if dist is None or sh.local:
        return lv
    import torch
    own = sh.plan(1).owned_ranges(sh.rank)
    pieces = [None] * sh.world
    dist.all_gather_object(pieces, [(int(b), int(e), lv[int(b):int(e)].copy()) for b, e in own])
    for lst in pieces:
        for b, e, v in lst:
            lv[b:e] = v
    return lv
