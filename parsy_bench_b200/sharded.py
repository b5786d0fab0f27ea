"""Sharded factorization across the GPUs of one node (one process per GPU): DESIGN.md §8.

Phase 1: every rank factors the bottom subtrees it owns (no communication: LBC's lower levels are disjoint subtrees,
cholesky/InspectionLevel_06.h:208-216).  Exchange: the owners' panels — a handful of contiguous runs of lValues per
rank — are broadcast over NVLink with NCCL.  Phase 2: the top separators, 1-D block-cyclic: every rank factors every
top block column (POTRF/TRSM are latency-bound and cheap) but applies only the trailing / descendant updates into the
block columns it owns; right before a block column is factored its owner broadcasts the panel.  torch.distributed is only the transport; all arithmetic runs in libparsy_cuda."""
import numpy as np

from . import executor as ex


class _DevArray:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


class ShardedCholesky:
    def __init__(self, S, rank, world, device, top_levels=1, block_cols=0, top_distributed=True):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, torch.device("cuda", device)
        args = (S.n, S.A2_p, S.A2_i, S.p, S.s, S.i_ptr, S.super, S.nsuper, S.sParent, S.col2Sup, S.nLevels, S.levelPtr,
                S.parPtr, S.partition)
        self.h1 = ex.Solver(*args, device=device, block_cols=block_cols, rank=rank, world=world, phase=1,
                            top_levels=top_levels)
        self.h2 = ex.Solver(*args, device=device, block_cols=block_cols, rank=rank, world=world, phase=2,
                            top_levels=top_levels, top_distributed=top_distributed)
        self.top_distributed = top_distributed and world > 1
        self.h2.adopt_factor(self.h1)
        self.ranges = [self.h1.owned_ranges(r) for r in range(world)]
        p1, p2 = self.h1.device_pointers(), self.h2.device_pointers()
        self.lv = torch.as_tensor(_DevArray(p1["factor"], S.xsize), device=self.device)
        self.s1 = torch.cuda.ExternalStream(p1["stream"], device=self.device)
        self.s2 = torch.cuda.ExternalStream(p2["stream"], device=self.device)
        self.s2side = torch.cuda.ExternalStream(self.h2.stream2(), device=self.device)
        self.lookahead = True
        self.p2p_exchange = False     # grouped point-to-point was slower than per-owner broadcasts at 8 GPUs (350 vs 286 ms, cfg5)
        self.p2p_chunk = 1 << 28      # doubles per message (2 GiB)
        self.exchange_bytes = int(sum(int((r[:, 1] - r[:, 0]).sum()) for i, r in enumerate(self.ranges) if i != rank) * 8)
        self.n_broadcasts = int(sum(len(r) for r in self.ranges))
        self.nsteps, self.first_top = self.h2.num_steps(), self.h2.first_top_step()
        self.step_bcasts = {}
        if self.top_distributed:
            for st in range(self.first_top, self.nsteps):
                bc = self.h2.step_bcasts(st)
                if len(bc):
                    self.step_bcasts[st] = [(int(o), int(b), int(e)) for o, b, e in bc]
            self.n_broadcasts += sum(len(v) for v in self.step_bcasts.values())

    def set_values(self, values):
        self.h1.set_values(values)

    def factor(self, dist=None):
        """Enqueues phase 1, the NVLink exchange and phase 2; returns without synchronising."""
        torch = self.torch
        self.s1.wait_stream(self.s2)      # phase 1 re-zeroes the buffer the previous phase 2 may still be writing
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.events = evs
        evs[0].record(self.s1)
        self.h1.factor()
        evs[1].record(self.s1)
        with torch.cuda.stream(self.s1):
            if dist is not None and self.world > 1:
                if self.p2p_exchange:
                    # all owners send at once (one NCCL group of point-to-point ops): every rank's ingress link is busy
                    # for the whole exchange instead of waiting for one broadcast root at a time
                    ops = []
                    for owner, runs in enumerate(self.ranges):
                        for b, e in runs:
                            for c0 in range(int(b), int(e), self.p2p_chunk):
                                t = self.lv[c0:min(c0 + self.p2p_chunk, int(e))]
                                if owner == self.rank:
                                    ops += [dist.P2POp(dist.isend, t, peer) for peer in range(self.world) if peer != owner]
                                else:
                                    ops.append(dist.P2POp(dist.irecv, t, owner))
                    if ops:
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()
                else:
                    for owner, runs in enumerate(self.ranges):
                        for b, e in runs:
                            dist.broadcast(self.lv[int(b):int(e)], src=owner)
            ev = evs[2]
            ev.record(self.s1)
        self.s2.wait_event(ev)
        if not self.top_distributed:
            self.h2.factor()
            evs[3].record(self.s2)
            return
        # distributed top: updates from the bottom into the owned top block columns, then step by step:
        # owners broadcast the block columns about to be factored, every rank factors them, owners update theirs
        if not self.lookahead:
            with torch.cuda.stream(self.s2):
                self.h2.factor_steps(0, self.first_top)
                for st in range(self.first_top, self.nsteps):
                    for owner, b, e in self.step_bcasts.get(st, ()):
                        dist.broadcast(self.lv[b:e], src=owner)
                    self.h2.factor_steps(st, st + 1)
                evs[3].record(self.s2)
            return
        # look-ahead: the side stream carries broadcast -> POTRF/TRSM -> updates into the next block column, the main
        # stream the bulk of this rank's trailing updates
        self.h2.factor_steps(0, self.first_top)
        for st in range(self.first_top, self.nsteps):
            self.h2.step_begin(st, st == self.first_top)
            bc = self.step_bcasts.get(st, ())
            if bc:
                with torch.cuda.stream(self.s2side):
                    for owner, b, e in bc:
                        dist.broadcast(self.lv[b:e], src=owner)
            self.h2.step_run(st)
        self.h2.steps_end()
        evs[3].record(self.s2)

    def phase_times_ms(self):
        """(phase 1, bottom exchange, phase 2 incl. its per-step broadcasts) of the last factor(), after sync()."""
        e = self.events
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])

    def sync(self):
        ok1 = self.h1.sync()
        ok2 = self.h2.sync()
        return ok1 and ok2

    def get_factor(self):
        return self.h2.get_factor()

    def close(self):
        self.h2.close()
        self.h1.close()
