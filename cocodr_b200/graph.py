"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer).

A BERT-base contrastive step is ~290 kernel launches of 5-120 us each; issued from Python the host side
(autograd dispatch, allocator, ctypes) is on the critical path once the kernels are fast.  Capturing the step
once and replaying it removes that: every ``cdr_*`` entry point only enqueues work on the caller's stream
(no allocation, no host sync), so the sequence is capturable as is.  Inputs live in static buffers that
``__call__`` overwrites before each replay.
"""
import torch

from . import ops


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, example_inputs); loss = step(*inputs)

    ``model(*inputs)`` must return the scalar loss or a tuple whose first element is it (the reference's
    calling convention, run_ann.py:320-325).  The optimizer must be capturable (e.g.
    ``torch.optim.AdamW(..., fused=True, capturable=True)``)."""

    def __init__(self, model, optimizer, example_inputs, warmup=3, backward_ctx=None):
        """backward_ctx: optional context manager entered around ``loss.backward()`` (e.g. a ``gradsync.GradSync``:
        its side-stream NCCL all-reduces are captured into the same graph)."""
        self.model, self.optimizer, self.backward_ctx = model, optimizer, backward_ctx
        self.static_inputs = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        from . import kernels
        n0 = kernels.launches
        # the captured sequence must contain the weight-shadow cast -- unless the optimizer rewrites the shadows itself
        ops.FORCE_SHADOW_REFRESH = not getattr(optimizer, "manages_shadows", False)
        try:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._forward_backward()
                optimizer.step()
        finally:
            ops.FORCE_SHADOW_REFRESH = False
        self.launches_per_replay = kernels.launches - n0

    def _forward_backward(self):
        out = self.model(*self.static_inputs)
        loss = out[0] if isinstance(out, (tuple, list)) else out
        if self.backward_ctx is not None:
            with self.backward_ctx:
                loss.backward()
        else:
            loss.backward()
        return loss

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self._forward_backward()
        self.optimizer.step()
        return loss

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        sync = getattr(self.optimizer, "sync_hyperparams", None)
        if sync is not None:  # cocodr_b200.optim: scheduler-driven lr changes reach the device scalar the graph reads
            sync()
        self.graph.replay()
        return self.static_loss
