"""Data-parallel gradient averaging without DistributedDataParallel.

Every encoder layer's backward (ops.BertLayerFn) leaves ALL its parameter gradients in one flat fp32 buffer
(the ``.grad`` tensors are views of it).  ``GradSync`` all-reduces each buffer over NCCL on a side stream as
soon as that layer's backward has been enqueued, so the exchange of layer i overlaps the backward kernels of
layers i-1, i-2, ... -- no bucket copies, no unused-parameter search, one collective per layer over NVLink.

    sync = GradSync(model)
    loss = model(...)[0]
    with sync:                 # collect + reduce while backward runs
        loss.backward()
    optimizer.step()           # the compute stream has waited for the last all-reduce

Reducing a flat buffer IN PLACE while backward is still running is only sound when autograd adopts its views as the
parameters' ``.grad`` and nothing adds to them afterwards.  That holds exactly when (a) the layer's Function ran once in
the forward this backward belongs to and (b) none of its parameters has a gradient yet.  ``submit`` checks both (the
forward counts come from ``ops.FWD_CALLS``); everything else -- the reference's default q_len != p_len (towers cannot be
fused, every layer runs 2-3 times per backward), gradient accumulation over several backwards, the tied word
embedding of the COCO model (MLM head + embedding table) -- is DEFERRED: those parameters' final ``.grad`` tensors are
reduced in ``__exit__``, after backward has finished with them.

``DistributedDataParallel(model, find_unused_parameters=True)`` (what the reference's driver does,
run_ann.py:178-184) keeps working on the same modules; this is the faster native path.
"""
import torch
import torch.distributed as dist

from . import ops


import os as _os
_SKIP = _os.environ.get("CDR_GRADSYNC_SKIP", "0") != "0"
_AT_END = _os.environ.get("CDR_GRADSYNC_AT_END", "0") != "0"


class GradSync:
    def __init__(self, model, group=None, arena=None):
        """arena: a peeropt.PeerArena -- gradients the backward writes into it are NOT reduced here (the optimizer's
        cdr_adam_multi_peer reads every rank's copy over NVLink); everything else still goes through NCCL."""
        self.model, self.group, self.arena = model, group, arena
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.comm = torch.cuda.Stream() if self.world > 1 and torch.cuda.is_available() else None
        # NCCL averages in the collective itself (ncclAvg); other backends (gloo in the CPU tests) sum, then scale
        nccl = dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._avg = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        self._bufs = []
        self._deferred = {}   # id(param) -> param whose final .grad is reduced in __exit__
        self.stats = {"overlapped": 0, "deferred": 0}

    def _reduce(self, t):
        if _SKIP:  # timing experiment only (tools / bench A-B): gradients stay local
            return
        dist.all_reduce(t, op=self._avg, group=self.group)  # averaged inside NCCL: no extra pass over the buffer
        if self._avg is dist.ReduceOp.SUM:
            t.mul_(1.0 / self.world)

    # called from the autograd Functions (ops.GRAD_SYNC.submit) right after a backward has been enqueued
    def submit(self, flat, params=None, fwd_calls=1):
        """flat: the zero-initialised buffer the Function's parameter gradients are views of; params: those
        parameters; fwd_calls: how many times the Function ran on them in the current forward (ops.FWD_CALLS)."""
        if self.world == 1:
            return
        if self.arena is not None and self.arena.contains(flat):
            self.stats["peer"] = self.stats.get("peer", 0) + 1
            return
        safe = params is not None and fwd_calls == 1 and all(p.grad is None and id(p) not in self._deferred for p in params)
        if _AT_END:  # experiment: no overlap with backward, everything reduced in __exit__
            safe = False
        if not safe:
            # autograd will ADD this buffer's views to existing gradients (or other buffers' views to these): the
            # parameters' final .grad is reduced once backward is over
            for p in (params or ()):
                self._deferred[id(p)] = p
            self.stats["deferred"] += 1
            return
        self.stats["overlapped"] += 1
        if self.comm is None:  # CPU tensors (gloo tests): no stream to overlap on
            self._reduce(flat)
        else:
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm.wait_event(ev)
            self._comm_used = True
            flat.record_stream(self.comm)
            with torch.cuda.stream(self.comm):
                self._reduce(flat)
        self._bufs.append(flat)

    def __enter__(self):
        self._comm_used = False
        self._bufs = []
        self._deferred = {}
        ops.GRAD_SYNC = self
        return self

    def __exit__(self, *exc):
        ops.GRAD_SYNC = None
        ops.FWD_CALLS.clear()
        if self.world == 1:
            return False
        # parameters whose gradient did not come through an overlapped flat buffer (heads outside the encoder,
        # deferred layers, or a gradient autograd had to copy instead of adopting the view): reduce them now
        covered = [(b.data_ptr(), b.data_ptr() + b.numel() * 4) for b in self._bufs]
        rest = []
        for p in self.model.parameters():
            if p.grad is None:
                continue
            if self.arena is not None and self.arena.contains(p):
                # exchanged inside the optimizer kernel, which reads the gradient of every rank from the arena: a
                # gradient autograd accumulated elsewhere (unfused towers, accumulation steps) is moved in first
                if not self.arena.contains(p.grad):
                    slot = self.arena.grad_slot(p)
                    slot.copy_(p.grad)
                    p.grad = slot
                    self.stats["peer_copied"] = self.stats.get("peer_copied", 0) + 1
                continue
            a = p.grad.data_ptr()
            if id(p) in self._deferred or not any(lo <= a < hi for lo, hi in covered):
                rest.append(p.grad)
        if rest:
            if self.comm is None:
                for g in self._coalesce(rest):
                    self._reduce(g)
            else:
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                self.comm.wait_event(ev)
                self._comm_used = True
                with torch.cuda.stream(self.comm):
                    for g in self._coalesce(rest):
                        g.record_stream(self.comm)
                        self._reduce(g)
        if self.comm is not None and self._comm_used:  # (nothing to join when every gradient went through the arena)
            torch.cuda.current_stream().wait_stream(self.comm)
        self._bufs = []
        self._deferred = {}
        return False

    @staticmethod
    def _coalesce(grads):
        """Gradients that are adjacent views of one allocation (a layer's flat buffer adopted by autograd) are reduced
        as one tensor: one collective per layer instead of sixteen.  The ORDER of the returned tensors must be the same
        on every rank (collectives are matched by issue order), so nothing here may depend on addresses: groups follow
        the first appearance of their storage in parameter order, members are ordered by storage offset."""
        groups, order = {}, []
        for g in grads:
            key = g.untyped_storage().data_ptr() if g.is_contiguous() else ("nc", id(g))
            if key not in groups:
                groups[key] = []
                order.append(key)
            groups[key].append(g)
        out = []
        for key in order:
            members = groups[key]
            if isinstance(key, tuple) or len(members) == 1:
                out.extend(members)
                continue
            members = sorted(members, key=lambda t: t.storage_offset())
            run_first, run_n = members[0], members[0].numel()
            for g in members[1:]:
                if g.dtype == run_first.dtype and g.storage_offset() == run_first.storage_offset() + run_n:
                    run_n += g.numel()
                else:
                    out.append(run_first if run_n == run_first.numel() else
                               torch.as_strided(run_first, (run_n,), (1,), run_first.storage_offset()))
                    run_first, run_n = g, g.numel()
            out.append(run_first if run_n == run_first.numel() else
                       torch.as_strided(run_first, (run_n,), (1,), run_first.storage_offset()))
        return out
