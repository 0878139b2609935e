"""Data-parallel gradient exchange fused into the optimizer over NVLink peer memory (host side of
``cdr_adam_multi_peer``, csrc/optim.cu).

The reference wraps the model in ``DistributedDataParallel`` (ANCE/drivers/run_ann.py:178-184): a bucketed NCCL
all-reduce of every gradient (440 MB fp32 for BERT-base) followed by the same optimizer step on every rank.  On a B200
node the all-reduce kernels and the persistent one-CTA-per-SM GEMMs of the backward cannot share SMs: measured at
N = 2, the overlapped all-reduce still costs 0.75 of the 0.92 ms it takes alone.  Here no collective runs at all:

* ``PeerArena`` puts the parameters, their fp16 operand shadows and persistent gradient buffers of a model into ONE
  symmetric allocation (``torch.distributed._symmetric_memory``) with the same layout on every rank, so rank r's copy
  of any of those addresses is ``address + (base_r - base_local)``;
* the backward writes its gradients into the arena (``ops._grad_flat``; ``gradsync.GradSync(model, arena=...)`` then
  skips NCCL for them);
* ``optim.AdamW.step`` launches ``cdr_adam_multi_peer``: each rank owns every ``world``-th 16 K-element chunk, reads that
  chunk of the gradient from all ranks (peer loads), averages, updates ITS exp_avg / exp_avg_sq and stores the new
  parameter and shadow into every rank's arena (peer stores) -- reduce-scatter + 1/world of the optimizer work +
  all-gather in one kernel, with two flag rounds (gradients final / updates landed) instead of stream-level collectives.

``exp_avg`` / ``exp_avg_sq`` of chunks a rank does not own stay at their initial zeros: ``consolidate_state`` sums
them over the ranks (what ``optimizer.state_dict()`` should save), ``shard_state`` re-zeroes the foreign chunks after a
``load_state_dict``.  Gradient clipping (``optimizer.clip_grad_norm_``: a global norm BEFORE the update) splits the
pass in two kernels -- reduce the owned chunks in place + exchange partial squared norms, then update from the local
reduced chunks; Lamb (per-tensor norms) is not supported in this mode: use the NCCL path (plain ``GradSync``).  No CPU
path.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, kernels, ops
from ._lib import check, stream_ptr
from .peer import PeerArgs

_ALIGN = 256  # bytes: every tensor of the arena starts on a 256-byte boundary (16-byte vector paths, TMA maps)


def _round(n, a=_ALIGN):
    return (n + a - 1) // a * a


class PeerArena:
    def __init__(self, model, optimizer=None, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import bert
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerArena: at most 8 ranks (one NVLink domain)")
        if optimizer is not None:
            from .optim import AdamW
            if not isinstance(optimizer, AdamW):
                raise NotImplementedError("PeerArena: only cocodr_b200.optim.AdamW exchanges gradients in its kernel")
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise RuntimeError("PeerArena needs fp32 CUDA parameters (no CPU path)")
        self.device = dev = params[0].device
        encoders = [m for m in model.modules() if isinstance(m, bert.BertModel)]
        # ---- layout: parameters | shadows | gradients (same on every rank: same model, same iteration order)
        p_bytes = sum(_round(p.numel() * 4) for p in params)
        s_bytes = 0
        for enc in encoders:
            cfg = enc.config
            H, I = cfg.hidden_size, cfg.intermediate_size
            per_layer = _round(3 * H * H * 2) + _round(3 * H * 4) + _round(H * H * 2) + 2 * _round(I * H * 2)
            s_bytes += per_layer * len(enc.encoder.layer)
        # gradients: the per-Function flat buffers (no padding between their views) + a per-parameter slot for gradients
        # that autograd had to accumulate outside the arena (unfused towers, gradient accumulation: GradSync copies
        # them in) -- sized for both
        g_bytes = 2 * p_bytes + _ALIGN * (2 * len(params) + 64)
        total = p_bytes + s_bytes + g_bytes
        self.arena = symm_mem.empty((total,), dtype=torch.uint8, device=dev)
        self.flags = symm_mem.empty((16,), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.norms = symm_mem.empty((16,), dtype=torch.float32, device=dev)  # [2][8] partial sums of squares (clipping)
        self.norms.zero_()
        self._h_arena = symm_mem.rendezvous(self.arena, group=self.group)
        self._h_flags = symm_mem.rendezvous(self.flags, group=self.group)
        self._h_norms = symm_mem.rendezvous(self.norms, group=self.group)
        self.sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.clip = torch.zeros(2, dtype=torch.float32, device=dev)  # coefficient, norm of the reduced gradient
        self.base = self.arena.data_ptr()
        self.total = total
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.done = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._off = 0
        # ---- parameters move into the arena (values kept)
        with torch.no_grad():
            for p in params:
                view = self._take(p.numel() * 4).view(torch.float32)[:p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        # ---- shadows: re-allocated from the arena, then re-cast
        prev, ops.SHADOW_ALLOC = ops.SHADOW_ALLOC, self._alloc_shadow
        try:
            for enc in encoders:
                for sh in enc._shadow_set.layers:
                    sh.key = None
                enc._shadow_set.ptr_key = enc._shadow_set.ver_key = enc._shadow_set.table = None
                enc._shadow_set.refresh([bert.shadow_sources(layer) for layer in enc.encoder.layer])
        finally:
            ops.SHADOW_ALLOC = prev
        self._grad_lo = self._off
        self._flats = {}
        self.model = model
        if optimizer is not None:
            self.attach(optimizer)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)  # every rank's flags are zero and its parameters are in place

    def attach(self, optimizer):
        """Make ``optimizer`` (a cocodr_b200.optim.AdamW over this model's parameters) exchange gradients in its kernel."""
        from .optim import AdamW
        if not isinstance(optimizer, AdamW):
            raise NotImplementedError("PeerArena: only cocodr_b200.optim.AdamW exchanges gradients in its kernel")
        optimizer.attach_shadows(self.model)
        optimizer.peer = self
        return optimizer

    # ------------------------------------------------------------------------------------------ allocation
    def _take(self, nbytes):
        off = self._off
        self._off = off + _round(nbytes)
        if self._off > self.total:
            raise RuntimeError("PeerArena: out of arena space")
        return self.arena[off:off + nbytes]

    def _alloc_shadow(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        return self._take(n * torch.empty((), dtype=dtype).element_size()).view(dtype)[:n].view(*shape)

    def flat(self, key_param, n, dev):
        """Persistent zeroed fp32 [n] gradient buffer of the Function whose first parameter is ``key_param``."""
        buf = self._flats.get(id(key_param))
        if buf is None or buf.numel() != n:
            buf = self._flats[id(key_param)] = self._take(n * 4).view(torch.float32)[:n]
        buf.zero_()
        return buf

    def grad_slot(self, p):
        """Persistent arena gradient of parameter ``p`` (for gradients that were accumulated outside the arena)."""
        buf = self._flats.get(("slot", id(p)))
        if buf is None:
            buf = self._flats[("slot", id(p))] = self._take(p.numel() * 4).view(torch.float32)[:p.numel()].view(p.shape)
        return buf

    def contains(self, t):
        a = t.data_ptr()
        return self.base <= a < self.base + self.total

    # ------------------------------------------------------------------------------------------ kernel arguments
    def peer_args(self):
        a = PeerArgs()
        a.world, a.rank, a.epoch = self.world, self.rank, self.epoch_dev.data_ptr()
        for r in range(self.world):
            a.peer_buf[r] = self._h_arena.buffer_ptrs[r]
            a.peer_flag[r] = self._h_flags.buffer_ptrs[r]
        a.done_counter = self.done.data_ptr()
        return a

    def adam_step(self, opt_args):
        pa = self.peer_args()
        check(_lib.load().cdr_adam_multi_peer(C.byref(opt_args), C.byref(pa), C.c_void_p(self.epoch_dev.data_ptr()),
                                              C.c_void_p(self.err.data_ptr()), stream_ptr()), "cdr_adam_multi_peer")
        kernels._count(2)

    def reduce_clip(self, opt_args, max_norm):
        """Phase 1 of the clipped step: reduce the owned chunks in place, exchange the partial squared norms; afterwards
        ``self.clip`` = (coefficient, norm) on the device."""
        pa = self.peer_args()
        slots = (C.c_void_p * 8)(*[self._h_norms.buffer_ptrs[r] if r < self.world else 0 for r in range(8)])
        check(_lib.load().cdr_grad_reduce_clip_peer(C.byref(opt_args), C.byref(pa), C.c_void_p(self.epoch_dev.data_ptr()),
                                                    C.c_float(max_norm), C.c_void_p(self.sq.data_ptr()), slots,
                                                    C.c_void_p(self.clip.data_ptr()), C.c_void_p(self.err.data_ptr()),
                                                    stream_ptr()), "cdr_grad_reduce_clip_peer")
        kernels._count(2)

    def adam_step_reduced(self, opt_args):
        """Phase 2: the update from the locally reduced gradients (opt_args.grad_scale = the clip coefficient)."""
        pa = self.peer_args()
        check(_lib.load().cdr_adam_multi_peer_reduced(C.byref(opt_args), C.byref(pa), C.c_void_p(self.epoch_dev.data_ptr()),
                                                      C.c_void_p(self.err.data_ptr()), stream_ptr()),
              "cdr_adam_multi_peer_reduced")
        kernels._count(2)

    def check(self):
        """Raises if a rank ever gave up waiting for a peer (host synchronisation: call outside the hot loop)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("cdr_adam_multi_peer: a peer did not arrive (timeout); parameters are not consistent")

    # ------------------------------------------------------------------------------------------ optimizer state
    def _owned_mask(self, n, first_chunk):
        from .optim import OPT_CHUNK
        idx = torch.arange(n, device=self.device) // OPT_CHUNK + first_chunk
        return (idx % self.world) == self.rank

    def consolidate_state(self, optimizer):
        """exp_avg / exp_avg_sq of every peer-updated parameter summed over the ranks (each rank holds its own chunks,
        zeros elsewhere): afterwards ``optimizer.state_dict()`` is the full state on every rank."""
        for tab in optimizer._tables.values():
            if not tab.get("peer"):
                continue
            for p in tab["params"]:
                st = optimizer.state[p]
                dist.all_reduce(st["exp_avg"], group=self.group)
                dist.all_reduce(st["exp_avg_sq"], group=self.group)

    def shard_state(self, optimizer):
        """Inverse of consolidate_state (after load_state_dict of a full state): zero the chunks other ranks own."""
        for tab in optimizer._tables.values():
            if not tab.get("peer"):
                continue
            first = 0
            from .optim import OPT_CHUNK
            for p in tab["params"]:
                st = optimizer.state[p]
                own = self._owned_mask(p.numel(), first).view(p.shape)
                st["exp_avg"].mul_(own)
                st["exp_avg_sq"].mul_(own)
                first += (p.numel() + OPT_CHUNK - 1) // OPT_CHUNK
