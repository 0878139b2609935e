"""Peer-memory exchange of the contrastive head: host side of csrc/peer.cu.

``PeerExchange`` owns the symmetric buffers (``torch.distributed._symmetric_memory``: the same allocation mapped on
every rank of the node over NVLink) the kernels write into:

    gather [2, world * n_rows, dim] fp32 forward: rank r's rows land in slot r on EVERY rank, written by the last
                                                  LayerNorm kernel itself (cdr_ln_fwd_push); double-buffered by epoch
                                                  parity, consumed through a private copy (cdr_peer_wait_fetch)
    recv   [world, n_rows, dim]  fp32   backward: rank r's gradient block for this rank lands in slot r
    flags  [16] uint32                  epochs published by the writers, polled by cdr_peer_wait / _reduce_slots

``gather_passages`` is the autograd-visible all-gather (forward: wait for the pushes, return the gather buffer;
backward: scatter the gradient blocks to their owners, wait, sum) -- a drop-in for the NCCL
``all_gather_into_tensor`` / ``reduce_scatter_tensor`` pair of ``models.gather_with_grad``.  No CPU path.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, kernels
from ._lib import check, stream_ptr


class PeerArgs(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_void_p),
                ("peer_buf", C.c_void_p * 8), ("peer_flag", C.c_void_p * 8), ("done_counter", C.c_void_p)]


class PeerExchange:
    def __init__(self, n_rows, dim, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerExchange: at most 8 ranks (one NVLink domain)")
        self.n_rows, self.dim, self.device = n_rows, dim, device
        self.gather = symm_mem.empty((2, self.world * n_rows, dim), dtype=torch.float32, device=device)
        self.recv = symm_mem.empty((self.world, n_rows, dim), dtype=torch.float32, device=device)
        self.flags = symm_mem.empty((16,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self._h = [symm_mem.rendezvous(t, group=self.group) for t in (self.gather, self.recv, self.flags)]
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=device)  # advanced on the device: graph capturable
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's flags are zeroed before anyone publishes an epoch

    def _args(self, which):
        a = PeerArgs()
        a.world, a.rank, a.epoch = self.world, self.rank, self.epoch_dev.data_ptr()
        for r in range(self.world):
            a.peer_buf[r] = self._h[which].buffer_ptrs[r]
            a.peer_flag[r] = self._h[2].buffer_ptrs[r]
        a.done_counter = self.done.data_ptr()
        return a

    # ---- forward -------------------------------------------------------------------------------
    def begin_step(self):
        """A new exchange round: call once per training step before the encoder runs."""
        check(_lib.load().cdr_peer_next_epoch(C.c_void_p(self.epoch_dev.data_ptr()), stream_ptr()), "cdr_peer_next_epoch")
        kernels._count(1)

    @property
    def epoch(self):
        return int(self.epoch_dev.item())

    def push_args(self):
        """cdr_peer_args for cdr_ln_fwd_push (target: the gather buffers)."""
        return self._args(0)

    def wait_gather(self):
        """Wait for every rank's push of this epoch and return a PRIVATE copy [world * n_rows, dim] of the gathered
        rows (never aliases memory the peers write into)."""
        out = torch.empty(self.world * self.n_rows, self.dim, dtype=torch.float32, device=self.device)
        check(_lib.load().cdr_peer_wait_fetch(C.c_void_p(self.flags.data_ptr()), C.c_int32(self.world),
                                              C.c_void_p(self.epoch_dev.data_ptr()), C.c_void_p(self.gather.data_ptr()),
                                              C.c_int64(out.numel()), C.c_void_p(out.data_ptr()), stream_ptr()),
              "cdr_peer_wait_fetch")
        kernels._count(1)
        return out

    # ---- backward ------------------------------------------------------------------------------
    def reduce_scatter(self, g):
        """g [world * n_rows, dim] fp32 -> sum over ranks of their block for this rank, [n_rows, dim]."""
        lib = _lib.load()
        g = g.contiguous().float()
        a = self._args(1)
        check(lib.cdr_peer_scatter_rows(C.c_void_p(g.data_ptr()), C.c_int32(self.n_rows), C.c_int32(self.dim),
                                        C.byref(a), stream_ptr()), "cdr_peer_scatter_rows")
        out = torch.empty(self.n_rows, self.dim, dtype=torch.float32, device=g.device)
        check(lib.cdr_peer_reduce_slots(C.c_void_p(self.recv.data_ptr()), C.c_void_p(self.flags[8:].data_ptr()),
                                        C.c_int32(self.world), C.c_int64(self.n_rows * self.dim),
                                        C.c_void_p(self.epoch_dev.data_ptr()), C.c_void_p(out.data_ptr()), stream_ptr()),
              "cdr_peer_reduce_slots")
        kernels._count(2)
        return out


class _GatherPassages(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_local, xchg):
        ctx.xchg = xchg
        # p_local already sits in slot `rank` of every rank's gather area (pushed by the LayerNorm kernel)
        return xchg.wait_gather()

    @staticmethod
    def backward(ctx, g):
        return ctx.xchg.reduce_scatter(g), None


def gather_passages(p_local, xchg):
    return _GatherPassages.apply(p_local, xchg)
