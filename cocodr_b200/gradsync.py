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

``DistributedDataParallel(model, find_unused_parameters=True)`` (what the reference's driver does,
run_ann.py:178-184) keeps working on the same modules; this is the faster native path.
"""
import torch
import torch.distributed as dist

from . import ops


class GradSync:
    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.comm = torch.cuda.Stream() if self.world > 1 else None
        # NCCL averages in the collective itself (ncclAvg); other backends (gloo in the CPU tests) sum, then scale
        nccl = dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._avg = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        self._bufs = []

    # called from the autograd Functions (ops.GRAD_SYNC.submit) right after a backward has been enqueued
    def submit(self, flat):
        if self.world == 1:
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm.wait_event(ev)
        flat.record_stream(self.comm)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(flat, op=self._avg, group=self.group)  # averaged inside NCCL: no extra pass over the buffer
            if self._avg is dist.ReduceOp.SUM:
                flat.mul_(1.0 / self.world)
        self._bufs.append(flat)

    def __enter__(self):
        self._bufs = []
        ops.GRAD_SYNC = self
        return self

    def __exit__(self, *exc):
        ops.GRAD_SYNC = None
        if self.world == 1:
            return False
        cur = torch.cuda.current_stream()
        # parameters whose gradient did not come through a flat buffer (heads outside the encoder, or a
        # gradient autograd had to copy instead of adopting the view): reduce them individually
        covered = [(b.data_ptr(), b.data_ptr() + b.numel() * 4) for b in self._bufs]
        rest = []
        for p in self.model.parameters():
            if p.grad is None:
                continue
            a = p.grad.data_ptr()
            if not any(lo <= a < hi for lo, hi in covered):
                rest.append(p.grad)
        if rest:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                for g in rest:
                    g.record_stream(self.comm)
                    dist.all_reduce(g, op=self._avg, group=self.group)
                    if self._avg is dist.ReduceOp.SUM:
                        g.mul_(1.0 / self.world)
        cur.wait_stream(self.comm)
        self._bufs = []
        return False
