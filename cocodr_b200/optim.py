"""Fused multi-tensor optimizers (SURVEY f-4): host-side mirror of the optimizers the reference's train loop builds.

The reference does (ANCE/drivers/run_ann.py:134-147, 345-353)

    optimizer = Lamb(grouped_parameters, lr=..., eps=...)          # utils/lamb.py, the default
    optimizer = AdamW(grouped_parameters, lr=..., eps=...)         # transformers.AdamW
    torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm); optimizer.step()

``Lamb`` / ``AdamW`` below keep those constructors, ``param_groups`` / ``state_dict`` layout (``exp_avg``,
``exp_avg_sq``, ``step`` per parameter) and ``step()`` / ``zero_grad()``, but one ``cdr_*_multi`` launch per group
updates every tensor and -- when the model registers its operand shadows (``attach_shadows``) -- rewrites the
fp16 copies the tcgen05 GEMMs read, so no separate weight cast runs per step.  lr / step counters live on the
device, which makes ``step()`` CUDA-graph capturable.  There is no CPU path.
"""
import ctypes as C

import torch

from . import _lib, kernels
from ._lib import check, stream_ptr

MODE_TORCH, MODE_HF = 0, 1


class OptItem(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("shadow", C.c_void_p),
                ("n", C.c_int64), ("shadow_f32", C.c_int32), ("reserved", C.c_int32)]


class OptChunk(C.Structure):
    _fields_ = [("start", C.c_int64), ("item", C.c_int32), ("reserved", C.c_int32)]


OPT_CHUNK = 16384  # CDR_OPT_CHUNK


class OptArgs(C.Structure):
    _fields_ = [("items", C.c_void_p), ("chunks", C.c_void_p), ("count", C.c_int32), ("n_chunks", C.c_int32),
                ("mode", C.c_int32), ("reserved", C.c_int32),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("lr", C.c_void_p), ("step", C.c_void_p), ("grad_scale", C.c_void_p), ("norms", C.c_void_p),
                ("trust", C.c_void_p)]


class _FusedOptimizer(torch.optim.Optimizer):
    manages_shadows = False  # True once a model's shadows are attached: graph.py then drops the per-step cast

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._shadow_of = {}      # id(param) -> (tensor view, is_f32)
        self._shadow_sets = []    # ops.ShadowSet objects to re-snapshot after a step
        self._tables = {}         # group index -> dict(key, dev table, pinned host table, max_n, params)
        self._clip = None         # device [3]: sum of squares scratch, coefficient, norm
        self._scale = None        # device scalar the next step() multiplies into every gradient (clip coefficient)
        self._use_clip = False
        self._peer_reduced = False  # peer path: clip_grad_norm_ already reduced the owned chunks in place
        self.peer = None          # a peeropt.PeerArena: AdamW.step exchanges the gradients inside cdr_adam_multi_peer

    # ------------------------------------------------------------------------------------------ shadows
    def attach_shadows(self, *models):
        """Register the fp16 / packed operand shadows of every cocodr_b200 encoder found in ``models``: step() then
        writes them together with the parameters (and the encoder skips its own cast)."""
        from . import bert
        for model in models:
            for mod in model.modules():
                if isinstance(mod, bert.BertModel):
                    for param, (dst, is_f32) in mod.shadow_map().items():
                        self._shadow_of[id(param)] = (dst, is_f32)
                    self._shadow_sets.append(mod._shadow_set)
        self.manages_shadows = bool(self._shadow_of)
        self._tables.clear()
        return self

    # ------------------------------------------------------------------------------------------ tables
    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _table(self, gi, group, subset=None):
        """subset: None (all parameters of the group with a gradient), or 'peer' / 'local' = those whose parameter,
        gradient and shadow all live in the peer arena / the rest (tables are cached per (group, subset))."""
        params = [p for p in group["params"] if p.grad is not None]
        if subset is not None:
            in_arena = lambda p: (self.peer.contains(p) and self.peer.contains(p.grad) and  # noqa: E731
                                  (id(p) not in self._shadow_of or self.peer.contains(self._shadow_of[id(p)][0])))
            params = [p for p in params if in_arena(p) == (subset == "peer")]
            gi = (gi, subset)
        if not params:
            return None
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("cocodr_b200.optim needs contiguous fp32 CUDA parameters (no CPU fallback)")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise RuntimeError("cocodr_b200.optim needs contiguous fp32 gradients")
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        tab = self._tables.get(gi)
        if tab is not None and tab["key"] == key:
            return tab
        n = len(params)
        if tab is None or tab["host_ring"][0].numel() != n * C.sizeof(OptItem):
            dev = params[0].device
            # the pinned staging copy of the table is a small RING: in eager mode the gradient pointers change almost
            # every step, and the non-blocking H2D copy of step n may still be queued when step n + 1 rewrites the
            # host side -- each slot is rewritten only after the event recorded behind its last copy has completed
            tab = {"host_ring": [torch.empty(n * C.sizeof(OptItem), dtype=torch.uint8).pin_memory() for _ in range(4)],
                   "host_ev": [None] * 4, "host_i": 0,
                   "dev": torch.empty(n * C.sizeof(OptItem), dtype=torch.uint8, device=dev),
                   "step": torch.zeros((), dtype=torch.float32, device=dev),
                   "lr": torch.zeros((), dtype=torch.float32, device=dev), "lr_host": None,
                   "norms": torch.zeros(2 * n, dtype=torch.float32, device=dev),
                   "trust": torch.zeros(n, dtype=torch.float32, device=dev)}
            self._tables[gi] = tab
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            # a captured H2D copy is replayed from this host buffer at every graph launch: it gets a buffer of its own
            # that is never rewritten (kept alive by the table)
            host = torch.empty(n * C.sizeof(OptItem), dtype=torch.uint8).pin_memory()
            tab.setdefault("host_captured", []).append(host)
        else:
            slot = tab["host_i"] = (tab["host_i"] + 1) % len(tab["host_ring"])
            if tab["host_ev"][slot] is not None:
                tab["host_ev"][slot].synchronize()  # (long done in practice: the ring is 4 steps deep)
            host = tab["host_ring"][slot]
        arr = (OptItem * n).from_address(host.data_ptr())
        for i, p in enumerate(params):
            st = self._state_for(p)
            sh = self._shadow_of.get(id(p))
            arr[i].p, arr[i].g = p.data_ptr(), p.grad.data_ptr()
            arr[i].m, arr[i].v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            arr[i].shadow = sh[0].data_ptr() if sh is not None else 0
            arr[i].shadow_f32 = int(sh[1]) if sh is not None else 0
            arr[i].n = p.numel()
            arr[i].reserved = int(any(a & 15 for a in (arr[i].p, arr[i].g, arr[i].m, arr[i].v)) or
                                  bool((arr[i].shadow or 0) & 15))  # unaligned entries take the scalar path
        tab["dev"].copy_(host, non_blocking=True)
        if not capturing:
            tab["host_ev"][slot] = torch.cuda.Event()
            tab["host_ev"][slot].record()
        sizes = tuple(p.numel() for p in params)
        if tab.get("sizes") != sizes:  # chunk work list: depends on the tensor sizes only
            import numpy as np
            per = [(n_ + OPT_CHUNK - 1) // OPT_CHUNK for n_ in sizes]
            ck = np.zeros(sum(per), dtype=np.dtype([("start", "<i8"), ("item", "<i4"), ("reserved", "<i4")]))
            ck["item"] = np.repeat(np.arange(len(sizes), dtype=np.int32), per)
            ck["start"] = np.concatenate([np.arange(c, dtype=np.int64) * OPT_CHUNK for c in per])
            tab["chunks"] = torch.from_numpy(ck.view(np.uint8)).to(tab["dev"].device)
            tab["n_chunks"], tab["sizes"] = len(ck), sizes
        tab["key"], tab["params"], tab["peer"] = key, params, subset == "peer"
        # one fp32 step counter per group on the device; seeded from the per-parameter state (state_dict round trips)
        # (a reference-format optimizer.pt stores ``step`` as a python int: ANCE/utils/lamb.py:93, transformers AdamW)
        tab["step"].copy_(torch.as_tensor(self.state[params[0]]["step"], dtype=torch.float32, device=tab["step"].device))
        return tab

    def _args(self, tab, group, mode=0):
        lr = float(group["lr"])
        if tab["lr_host"] != lr:  # schedulers change group['lr'] on the host: mirror it on the device when it moves
            tab["lr"].fill_(lr)
            tab["lr_host"] = lr
        a = OptArgs()
        a.items, a.count, a.mode = tab["dev"].data_ptr(), len(tab["params"]), mode
        a.chunks, a.n_chunks = tab["chunks"].data_ptr(), tab["n_chunks"]
        a.beta1, a.beta2 = group["betas"]
        a.eps, a.weight_decay = group["eps"], group["weight_decay"]
        a.lr, a.step = tab["lr"].data_ptr(), tab["step"].data_ptr()
        a.grad_scale = self._scale.data_ptr() if self._use_clip else 0
        a.norms, a.trust = tab["norms"].data_ptr(), tab["trust"].data_ptr()
        return a

    def sync_hyperparams(self):
        """Mirror host-side ``group['lr']`` changes (LR schedulers) into the device scalars the kernels read.  step()
        does this itself; a CUDA-graph replay does not run Python, so ``graph.GraphedTrainStep`` calls it before
        every replay."""
        for key, tab in self._tables.items():
            group = self.param_groups[key[0] if isinstance(key, tuple) else key]
            if tab["lr_host"] != float(group["lr"]):
                tab["lr"].fill_(float(group["lr"]))
                tab["lr_host"] = float(group["lr"])

    def _finish(self, tab):
        for p in tab["params"]:
            self.state[p]["step"] = tab["step"]  # shared device counter (what state_dict() stores)
            torch.autograd.graph.increment_version(p)  # the kernel wrote p behind autograd's back

    def _after_step(self):
        self._use_clip = False
        self._peer_reduced = False
        for ss in self._shadow_sets:
            ss.mark_fresh()

    # ------------------------------------------------------------------------------------------ clipping
    def clip_grad_norm_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_(all parameters of this optimizer, max_norm) without touching the
        gradients: the coefficient stays on the device and is folded into the next step().  Returns the norm
        (device scalar tensor)."""
        if self.peer is not None:
            return self._clip_peer(max_norm)
        tabs = [t for t in (self._table(gi, g) for gi, g in enumerate(self.param_groups)) if t is not None]
        if not tabs:
            return None
        dev = tabs[0]["dev"].device
        if self._clip is None:
            self._clip = torch.zeros(3, dtype=torch.float32, device=dev)
        lib = _lib.load()
        self._clip[0].zero_()
        for tab in tabs:  # all groups accumulate into one sum of squares ...
            check(lib.cdr_grad_sqnorm_multi(C.c_void_p(tab["dev"].data_ptr()), C.c_void_p(tab["chunks"].data_ptr()),
                                            C.c_int32(tab["n_chunks"]), C.c_void_p(self._clip.data_ptr()), stream_ptr()),
                  "cdr_grad_sqnorm_multi")
        check(lib.cdr_grad_clip_coef(C.c_void_p(self._clip.data_ptr()), C.c_float(max_norm),  # ... -> coefficient
                                     C.c_void_p(self._clip[1:].data_ptr()), C.c_void_p(self._clip[2:].data_ptr()),
                                     stream_ptr()), "cdr_grad_clip_coef")
        self._scale = self._clip[1:]
        self._use_clip = True
        return self._clip[2]


    def _clip_peer(self, max_norm):
        """clip_grad_norm_ with the peer-memory optimizer: phase 1 of the split pass (peeropt.PeerArena.reduce_clip) --
        every rank reduces the chunks it owns in place and the partial squared norms are exchanged; step() then runs
        phase 2 with the coefficient.  One parameter group whose gradients all live in the arena."""
        if len(self.param_groups) != 1:
            raise NotImplementedError("peer-memory clipping supports a single parameter group")
        group = self.param_groups[0]
        if self._table(0, group, "local") is not None:
            raise NotImplementedError("peer-memory clipping needs every gradient in the arena")
        tab = self._table(0, group, "peer")
        if tab is None:
            return None
        self.peer.reduce_clip(self._args(tab, group, getattr(self, "mode", 0)), float(max_norm))
        self._scale = self.peer.clip  # [coefficient, norm]: _args hands the coefficient to the update as grad_scale
        self._use_clip = True
        self._peer_reduced = True
        return self.peer.clip[1]


class AdamW(_FusedOptimizer):
    """AdamW with the constructor of ``transformers.AdamW`` (``correct_bias`` must stay True) / ``torch.optim.AdamW``.

    ``semantics='hf'`` (default, what the reference imports) or ``'torch'`` picks the update rule (they differ in
    where eps and the weight decay enter; see include/cocodr_b200.h)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 semantics="hf"):
        if not correct_bias:
            raise NotImplementedError("correct_bias=False is not implemented")
        if semantics not in ("hf", "torch"):
            raise ValueError("semantics must be 'hf' or 'torch'")
        self.mode = MODE_HF if semantics == "hf" else MODE_TORCH
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            if self.peer is not None:
                tab = self._table(gi, group, "peer")
                if tab is not None:
                    if self._peer_reduced:  # clip_grad_norm_ ran phase 1: update from the local reduced chunks
                        self.peer.adam_step_reduced(self._args(tab, group, self.mode))
                    else:
                        self.peer.adam_step(self._args(tab, group, self.mode))
                    self._finish(tab)
                tab = self._table(gi, group, "local")  # gradients outside the arena were all-reduced by GradSync
            else:
                tab = self._table(gi, group)
            if tab is None:
                continue
            a = self._args(tab, group, self.mode)
            check(lib.cdr_adam_multi(C.byref(a), stream_ptr()), "cdr_adam_multi")
            kernels._count(2)  # step counter + update
            self._finish(tab)
        self._after_step()
        return loss


class Lamb(_FusedOptimizer):
    """utils/lamb.py ``Lamb`` (ANCE/utils/lamb.py:24-121): same constructor, same update (no bias correction,
    weight norm clamped to [0, 10], trust ratio 1 when either norm is 0); ``state[p]['trust_ratio']`` is a device
    scalar view."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False):
        if adam:
            raise NotImplementedError("adam=True (trust ratio forced to 1) is not implemented: use AdamW")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            tab = self._table(gi, group)
            if tab is None:
                continue
            a = self._args(tab, group)
            check(lib.cdr_lamb_multi(C.byref(a), stream_ptr()), "cdr_lamb_multi")
            kernels._count(3)  # step counter + two phases
            self._finish(tab)
            for i, p in enumerate(tab["params"]):
                self.state[p]["trust_ratio"] = tab["trust"][i]
        self._after_step()
        return loss
