"""Group-DRO losses of the ANCE stage, same surface as the reference's ``ANCE/model/dro_loss.py``
(``DROGreedyLoss`` :11-135, ``AverageMeter`` :138-158, ``iDROLoss`` :160-254): constructor arguments,
registered buffers (``h_fun``, ``sum_losses``, ``count_cat`` -- they ride along in the state dict),
``forward`` signatures and return tuples are identical.

Group statistics (K10) and the Gram matrix of per-group gradients (K12) run through the C ABI.  In a
multi-GPU run the reference all-reduces the full [G, P_last] gradient matrix (dro_loss.py:232, 4.25 GB at
G=50 / BERT-base); here every rank reduce-scatters it along P_last, takes the Gram of its column shard
and all-reduces the [G, G] result -- the same numbers with half the bytes on NVLink and no full-matrix
re-read.
"""
import os
from collections import defaultdict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import kernels as K
from . import ops


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class AverageMeter(object):
    """Running average (ANCE/model/dro_loss.py:138-158)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count > 0 else 0.


class MeterBank:
    """The (2 + 2G) ``.item()`` host reads per step of the reference's meters (ANCE/model/models.py:269-271) kept on
    the device: ``add`` accumulates (val, val * n, n) per meter with three tiny device ops -- no synchronisation, so a
    whole DRO step can be captured into a CUDA graph -- and the host values are fetched (one transfer for all meters)
    only when somebody READS a meter."""

    def __init__(self, n):
        self.n = n
        self.dev = None
        self.host = [[0.0, 0.0, 0.0] for _ in range(n)]  # val, sum, count contributed through update() on the host
        self.snap = None                                  # last device snapshot
        self.dirty = False

    def add(self, vals, ns):
        """vals, ns: device tensors [n]."""
        if self.dev is None or self.dev.device != vals.device:
            self.dev = torch.zeros(self.n, 3, dtype=torch.float64, device=vals.device)
        v, c = vals.detach().to(torch.float64), ns.detach().to(torch.float64)
        self.dev[:, 0] = v
        self.dev[:, 1] += v * c
        self.dev[:, 2] += c
        self.dirty = True

    def read(self, i):
        if self.dirty:
            self.snap = self.dev.tolist()
            self.dirty = False
        d = self.snap[i] if self.snap is not None else (0.0, 0.0, 0.0)
        h = self.host[i]
        return (d[0] if self.snap is not None else h[0]), h[1] + d[1], h[2] + d[2]

    def meters(self):
        return [BankMeter(self, i) for i in range(self.n)]


class BankMeter(object):
    """AverageMeter look-alike (val / sum / count / avg / update / reset) backed by a MeterBank slot."""

    def __init__(self, bank, i):
        self._bank, self._i = bank, i

    val = property(lambda self: self._bank.read(self._i)[0])
    sum = property(lambda self: self._bank.read(self._i)[1])
    count = property(lambda self: self._bank.read(self._i)[2])

    @property
    def avg(self):
        _, s, c = self._bank.read(self._i)
        return s / c if c > 0 else 0.

    def update(self, val, n=1):
        h = self._bank.host[self._i]
        h[0], h[1], h[2] = val, h[1] + val * n, h[2] + n

    def reset(self):
        self._bank.host[self._i] = [0.0, 0.0, 0.0]
        if self._bank.dev is not None:
            self._bank.dev[self._i].zero_()
            self._bank.dirty = True


class DROGreedyLoss(nn.Module):
    def __init__(self, args, n_groups, alpha, eps, ema=0.1, weight_ema=False, weight_cutoff=True, fraction=None):
        super().__init__()
        self.args = args
        self.alpha = alpha
        self.ema = ema
        self.eps = eps
        self.weight_cutoff = weight_cutoff
        self.weight_ema = weight_ema
        self.n_groups = n_groups
        self.id2group = {str(x): f"group{x}" for x in range(n_groups)}
        self.n_splits = n_groups
        self.register_buffer('h_fun', torch.ones(self.n_splits))
        self.register_buffer('sum_losses', torch.zeros(self.n_splits))
        if fraction is not None:
            self.register_buffer('fraction', torch.as_tensor(fraction).float())
            self.register_buffer('count_cat', None)
        else:
            self.register_buffer('count_cat', torch.ones(self.n_splits))
        self.idx_dict = defaultdict(lambda: len(self.idx_dict))
        for i in range(self.n_groups):
            _ = self.idx_dict['[' + str(i) + ']']

    def reset(self):
        self.h_fun.fill_(1.)
        self.sum_losses.fill_(0.)
        if self.count_cat is not None:
            self.count_cat.fill_(1.)

    def reset_loss(self):
        self.h_fun.fill_(1.)
        self.sum_losses.fill_(0.)

    def forward(self, losses, g, w=None):
        """dro_loss.py:49-86 -> (robust_loss, group_losses[G], group_counts[G])."""
        if w is not None:
            losses = losses * w
        batch_size = losses.size(0)
        gdro_losses, gdro_counts = ops.group_stats(losses, g, self.n_groups)
        # (the state buffers are updated IN PLACE below -- a captured CUDA graph must find them at the same address at
        # every replay -- so the loss multiplies a copy: autograd saves it for the backward)
        robust_loss = (gdro_losses * self.h_fun.clone()).sum() / batch_size
        with torch.no_grad():
            if self.training:
                if _world() > 1:
                    # one fused exchange instead of the reference's two all_gathers (:64-65)
                    packed = torch.stack([gdro_losses.detach(), gdro_counts])
                    dist.all_reduce(packed)
                    losses_agg, counts_agg = packed[0], packed[1]
                else:
                    losses_agg, counts_agg = gdro_losses.detach(), gdro_counts
                group_losses_agg = losses_agg / (counts_agg + (counts_agg == 0).float())
                valid = counts_agg > 0
                self.sum_losses.copy_(torch.where(valid, self.sum_losses * (1 - self.ema) + group_losses_agg * self.ema,
                                                  self.sum_losses))
                if self.count_cat is not None:
                    self.count_cat.copy_(self.count_cat * (1 - self.ema) + counts_agg * self.ema)
                self.update_mw()
            group_losses = gdro_losses.detach() / (gdro_counts + (gdro_counts == 0).float())
        return robust_loss, group_losses, gdro_counts

    def update_mw(self):
        """dro_loss.py:88-120: greedy re-weighting of the worst groups (alpha-fraction cut-off)."""
        past_losses = self.sum_losses
        if self.count_cat is not None:
            past_frac = self.count_cat / self.count_cat.sum()
        else:
            past_frac = self.fraction
        sorted_losses, sort_id = torch.sort(past_losses, descending=True)
        sorted_frac = past_frac[sort_id]
        cutoff_count = torch.sum(torch.cumsum(sorted_frac, 0) < self.alpha)
        cutoff_count = torch.clamp(cutoff_count, max=sorted_frac.numel() - 1)
        idx = torch.arange(sorted_frac.numel(), device=sorted_frac.device)
        top = idx < cutoff_count
        tmp_sorted = torch.where(top, torch.full_like(sorted_frac, 1.0 / self.alpha),
                                 torch.full_like(sorted_frac, self.eps))
        leftover_mass = 1.0 - (sorted_frac * top.float()).sum() / self.alpha
        tiebreak = torch.clamp(leftover_mass / sorted_frac[cutoff_count], min=self.eps)
        tmp_sorted = torch.where(idx == cutoff_count, tiebreak.expand_as(tmp_sorted), tmp_sorted)
        tmp = torch.empty_like(tmp_sorted)
        tmp[sort_id] = tmp_sorted
        if self.weight_ema:
            tmp = torch.clamp(tmp, min=self.eps)
            self.h_fun.copy_(self.h_fun * (1 - self.ema) + tmp * self.ema)
        else:
            self.h_fun.copy_(tmp)


class iDROLoss(DROGreedyLoss):
    def __init__(self, args, n_groups, alpha, eps, ema, rho, reg=0, weight_ema=False, weight_cutoff=True,
                 fraction=None):
        super().__init__(args, n_groups, alpha, eps, ema, weight_ema, weight_cutoff, fraction)
        self.rho = rho
        self.reg = reg
        self.tol = 1e-5
        self.max_iter = 1200
        self.para_name = {}

    def _params(self, model):
        """dro_loss.py:174-190: parameters of the last 3 (base) / 2 (large) encoder layers."""
        if self.args.model_size == 'large':
            select = ['layer.23', 'layer.22']
        else:
            select = ['layer.10', 'layer.11', 'layer.9']
        if len(self.para_name) == 0:
            params = []
            for name, param in model.named_parameters():
                if any(name.find(s) >= 0 for s in select):
                    params.append(param)
                    self.para_name[name] = 1
        else:
            params = [param for (name, param) in model.named_parameters() if name in self.para_name]
        return params

    def _get_grad(self, params, gdro_losses_agg, gdro_counts_agg):
        """[G, P_last] fp32 matrix of per-group gradients of the local group means (rows of absent groups = 0).

        Same contract as dro_loss.py:192-204 (one ``autograd.grad(..., retain_graph=True)`` per present
        group through the per-layer autograd Functions), written row-by-row into one preallocated matrix
        instead of cat-ing G vectors."""
        dim = sum(p.numel() for p in params)
        mat = torch.zeros(self.n_groups, dim, dtype=torch.float32, device=gdro_losses_agg.device)
        present = torch.nonzero(gdro_counts_agg > 0).flatten().tolist()
        for li in present:
            grads = torch.autograd.grad(gdro_losses_agg[li], params, retain_graph=True, allow_unused=True)
            off = 0
            for p, gr in zip(params, grads):
                n = p.numel()
                if gr is not None:
                    mat[li, off:off + n].copy_(gr.reshape(-1))
                off += n
        return mat

    def _get_grad_grouped(self, params, gdro_losses_agg, gdro_counts_agg, g, n_towers):
        """The same [G, P_last] matrix as ``_get_grad`` from ONE partial backward (K11).

        Valid when sample i's loss depends on the encoder outputs of sample i only (the triplet NLL,
        ANCE/model/models.py:101-108) and the towers ran as one pass over ``n_towers * B`` sequences (sequence s belongs
        to sample s % B).  Then in the backward of sum_g mean_g every row of an activation gradient carries
        d mean_{g(row)} only, and the gradient of mean_g w.r.t. a weight is dY[rows of g]^T X[rows of g]: the layer
        backwards hand over their wgrad operands (ops.GroupCapture), rows are regrouped so that each group is one
        contiguous K range, and the wgrad of every weight adds each group's product into the group's row of the matrix
        -- as one launch over (tile, group) work items with the k-ranges in device memory (cdr_gemm_grouped,
        ``grouped_kernel``) or as one launch per present group (cdr_gemm_segments).  Bias / LayerNorm gradients are
        per-group column sums: one GEMM with a one-hot [rows, G] operand.
        """
        cap = ops.GroupCapture()
        ops.GROUP_CAPTURE = cap
        try:
            torch.autograd.grad(gdro_losses_agg.sum(), params, retain_graph=True, allow_unused=True)
        finally:
            ops.GROUP_CAPTURE = None
        return self._grouped_from_records(cap.records, params, gdro_counts_agg, g, n_towers)

    def _grouped_from_records(self, records, params, gdro_counts_agg, g, n_towers):
        G = self.n_groups
        dev = gdro_counts_agg.device
        offs, dim = {}, 0
        for p in params:
            offs[id(p)] = (dim, p)
            dim += p.numel()
        mat = torch.zeros(G, dim, dtype=torch.float32, device=dev)
        covered = {k for rec in records for k in rec["keys"]}
        if not set(offs) <= covered:
            raise RuntimeError("iDROLoss: grouped gradients need every selected parameter to belong to an encoder layer "
                               "that ran through cocodr_b200.ops (set iDROLoss.grouped_wgrad = False otherwise)")
        seq_group = g.to(torch.int64).repeat(n_towers)  # towers are concatenated: sequence s -> sample s % B
        order = torch.argsort(seq_group, stable=True)
        if self.grouped_kernel:  # group sizes stay on the device (cdr_gemm_grouped reads its k-ranges there)
            per_group = torch.zeros(G, dtype=torch.int64, device=dev).scatter_add_(0, seq_group, torch.ones_like(seq_group))
        else:
            per_group = (gdro_counts_agg.to(torch.int64) * n_towers).tolist()  # sequences per group (one host transfer)
        Gp = (G + 127) // 128 * 128
        onehot = {}
        for rec in records:
            if any(k in offs for k in rec["keys"]):
                self._grouped_layer(rec, mat, offs, seq_group, order, per_group, Gp, onehot)
        return mat

    @staticmethod
    def _padded_layout(seq_group, order, cnt_seq, r, n_groups):
        """Row layout for cdr_gemm_grouped when a sequence owns r rows: every group's rows contiguous and padded to
        whole 64-row k-blocks.  -> (seg_kb int32 [G + 1], src rows, destination rows, static row bound); all on the
        device, nothing read back."""
        n_seq = seq_group.numel()
        dev = seq_group.device
        rows = cnt_seq * r
        padded = (rows + 63) // 64 * 64
        ends = torch.cumsum(padded, 0)
        seg_kb = (torch.cat([ends.new_zeros(1), ends]) // 64).to(torch.int32).contiguous()
        first = torch.cumsum(cnt_seq, 0) - cnt_seq       # first sorted sequence of every group
        gs = seq_group[order]                            # group of the j-th sorted sequence
        dest0 = (ends - padded)[gs] + (torch.arange(n_seq, device=dev) - first[gs]) * r
        ar = torch.arange(r, device=dev)
        dest = (dest0.unsqueeze(1) + ar).reshape(-1)
        src = (order.unsqueeze(1) * r + ar).reshape(-1)
        return seg_kb, src, dest, n_seq * r + 64 * n_groups

    def _grouped_layer(self, rec, mat, offs, seq_group, order, per_group, Gp, onehot):
        n_seq = rec["n_seq"]
        if n_seq != seq_group.numel():
            raise RuntimeError("iDROLoss: grouped gradients need the towers in one encoder pass "
                               f"({n_seq} sequences in the layer, {seq_group.numel()} expected)")
        if self.grouped_kernel:
            self._grouped_weights_one_launch(rec, mat, offs, seq_group, order, per_group, onehot)
        else:
            self._grouped_weights_segments(rec, mat, offs, order, per_group)
        self._grouped_vectors(rec, mat, offs, seq_group, Gp, onehot)

    def _grouped_weights_one_launch(self, rec, mat, offs, seq_group, order, cnt_seq, cache):
        """Weights through cdr_gemm_grouped: one launch per weight, (tile, group) work items, k-ranges on the device."""
        keys, rps, L, n_seq, S = rec["keys"], rec["rps"], rec["L"], rec["n_seq"], rec["S"]
        inv = 1.0 / S
        G = self.n_groups
        H = rec["x"].shape[1]
        dim = mat.shape[1]

        def layout(r):
            if ("layout", r) not in cache:
                cache[("layout", r)] = self._padded_layout(seq_group, order, cnt_seq, r, G)
            return cache[("layout", r)]

        def place(t, r):  # rows regrouped (and zero-padded per group when r is not a whole number of k-blocks)
            seg_kb, src, dest, bound = layout(r)
            if r % 64 == 0:
                C = t.shape[1]
                return t.reshape(n_seq, r * C)[order].reshape(n_seq * r, C)
            buf = torch.zeros(bound, t.shape[1], dtype=t.dtype, device=t.device)
            buf.index_copy_(0, dest, t.index_select(0, src))
            return buf

        def wgrads(k, a, b, r):
            ent = offs.get(keys[k])
            if ent is not None:
                K.gemm_grouped(a, b, mat, M=a.shape[1], N=b.shape[1], seg_kb=layout(r)[0], n_groups=G,
                               out_group_stride=dim, out_offset=ent[0], alpha=inv)

        dy2, gl, dz, x1, dy1, att = (place(rec[n], rps) for n in ("dy2", "gl", "dz", "x1", "dy1", "att"))
        dqkv, x = place(rec["dqkv"], L), place(rec["x"], L)
        wgrads(12, dy2, gl, rps)
        wgrads(10, dz, x1, rps)
        wgrads(6, dy1, att, rps)
        for j, k in enumerate((0, 2, 4)):
            wgrads(k, dqkv[:, j * H:(j + 1) * H], x, L)

    def _grouped_weights_segments(self, rec, mat, offs, order, per_group):
        """Weights through cdr_gemm_segments: one wgrad per (weight, present group), enqueued by one C call per weight."""
        keys, rps, L, n_seq, S = rec["keys"], rec["rps"], rec["L"], rec["n_seq"], rec["S"]
        inv = 1.0 / S
        H = rec["x"].shape[1]

        def regroup(t, r):  # [n_seq * r, C] rows -> the sequences of every group contiguous
            C = t.shape[1]
            return t.reshape(n_seq, r * C)[order].reshape(n_seq * r, C)

        # ---- weights: per present group one wgrad over that group's rows (K range), stored at the group's row of mat
        present = [gi for gi, ns in enumerate(per_group) if ns > 0]
        first, o = {}, 0
        for gi in present:
            first[gi] = o
            o += per_group[gi]
        dim = mat.shape[1]

        def wgrads(k, a, b, r):
            ent = offs.get(keys[k])
            if ent is not None:
                K.gemm_segments(a, b, mat, M=a.shape[1], N=b.shape[1], row_begin=[first[gi] * r for gi in present],
                                row_count=[per_group[gi] * r for gi in present],
                                out_offset=[gi * dim + ent[0] for gi in present], alpha=inv)

        dy2, gl, dz, x1, dy1, att = (regroup(rec[n], rps) for n in ("dy2", "gl", "dz", "x1", "dy1", "att"))
        dqkv, x = regroup(rec["dqkv"], L), regroup(rec["x"], L)
        wgrads(12, dy2, gl, rps)   # output.dense.weight [H, I]
        wgrads(10, dz, x1, rps)    # intermediate.dense.weight [I, H]
        wgrads(6, dy1, att, rps)   # attention.output.dense.weight [H, H]
        for j, k in enumerate((0, 2, 4)):  # query / key / value weights: column blocks of dQKV
            wgrads(k, dqkv[:, j * H:(j + 1) * H], x, L)

    def _grouped_vectors(self, rec, mat, offs, seq_group, Gp, onehot):
        """Bias / LayerNorm gradients: per-group column sums = one-hot^T R through the same GEMM (exact products, fp32
        accumulation)."""
        keys, rps, L, n_seq, S = rec["keys"], rec["rps"], rec["L"], rec["n_seq"], rec["S"]
        inv = 1.0 / S
        G = self.n_groups
        dev = mat.device
        H = rec["x"].shape[1]

        def hot(r):
            if r not in onehot:
                e = torch.zeros(n_seq * r, Gp, dtype=torch.float16, device=dev)
                e.scatter_(1, seq_group.repeat_interleave(r).unsqueeze(1), 1.0)
                onehot[r] = e
            return onehot[r]

        def colsums(r, R, alpha, ks):
            R = R.contiguous()
            n = R.shape[1]
            out = torch.zeros(Gp, n, dtype=torch.float32, device=dev)
            K.gemm(hot(r), R, out, M=Gp, N=n, K=R.shape[0], a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0,
                   alpha=alpha)
            w = n // len(ks)
            for j, k in enumerate(ks):
                ent = offs.get(keys[k])
                if ent is not None:
                    mat[:, ent[0]:ent[0] + w] = out[:G, j * w:(j + 1) * w]

        colsums(rps, rec["dy2"], inv, (13,))          # output.dense.bias
        colsums(rps, rec["dz"], inv, (11,))           # intermediate.dense.bias
        colsums(rps, rec["dy1"], inv, (7,))           # attention.output.dense.bias
        colsums(L, rec["dqkv"], inv, (1, 3, 5))       # query / key / value bias
        xhat1 = (rec["y1"].float() - rec["mean1"].unsqueeze(1)) * rec["rstd1"].unsqueeze(1)
        d1 = rec["dx1"].float()
        colsums(rps, rec["dx1"], inv, (9,))                        # attention.output.LayerNorm.bias
        colsums(rps, (d1 * xhat1).half(), inv, (8,))               # attention.output.LayerNorm.weight
        dy, dcls = rec["din2"]
        if dy is not None:  # fp16, already in the scaled-gradient domain; the [CLS] gradient joins it through S
            d2, a2 = dy.float(), inv
            if dcls is not None:
                d2.view(n_seq, rps, H)[:, 0] += S * dcls
        else:
            d2, a2 = torch.zeros(n_seq * rps, H, dtype=torch.float32, device=dev), 1.0
            d2.view(n_seq, rps, H)[:, 0] = dcls
        xhat2 = (rec["y2"].float() - rec["mean2"].unsqueeze(1)) * rec["rstd2"].unsqueeze(1)
        colsums(rps, d2.half(), a2, (15,))                         # output.LayerNorm.bias
        colsums(rps, (d2 * xhat2).half(), a2, (14,))               # output.LayerNorm.weight

    def _gram(self, all_grads):
        """Gram matrix of the rank-summed gradient rows (what :232-237 compute through a full all-reduce)."""
        G, P = all_grads.shape
        gram = torch.zeros(G, G, dtype=torch.float32, device=all_grads.device)
        W = _world()
        if W == 1:
            K.gram_f32(all_grads, gram)
            return gram
        shard = (P + W - 1) // W
        shard = (shard + 3) // 4 * 4
        padded = torch.zeros(W, G, shard, dtype=torch.float32, device=all_grads.device)
        for r in range(W):  # column shard r of every row, laid out [W, G, shard] for reduce_scatter
            lo, hi = r * shard, min(P, (r + 1) * shard)
            if hi > lo:
                padded[r, :, :hi - lo].copy_(all_grads[:, lo:hi])
        mine = torch.empty(G, shard, dtype=torch.float32, device=all_grads.device)
        dist.reduce_scatter_tensor(mine, padded.view(W * G, shard))
        K.gram_f32(mine, gram)
        dist.all_reduce(gram)
        return gram

    # K11: take the group gradients from one shared partial backward + per-group wgrads when the caller vouches that
    # the per-sample losses do not mix samples (``sample_towers``); False = one partial backward per present group
    grouped_wgrad = os.environ.get("CDR_IDRO_GROUPED", "1") == "1"
    # weights of the grouped path through ONE cdr_gemm_grouped launch each (device-side group table, no host transfer);
    # False = one cdr_gemm_segments wgrad per present group
    grouped_kernel = os.environ.get("CDR_IDRO_GROUPED_KERNEL", "1") == "1"

    def forward(self, model, losses, g, sample_towers=0, grad_losses=None):
        """dro_loss.py:216-254 -> (robust_loss, group mean losses[G] detached, group counts[G]).

        ``sample_towers``: n > 0 promises that loss i depends on encoder sequences {i, i + B, ..} of ONE pass over
        n * B sequences only (the triplet NLL), which lets ``_get_grad_grouped`` replace the per-group backwards.
        ``grad_losses``: the same per-sample loss VALUES on a different autograd graph, used for the group gradients
        only (the robust loss and the training gradient always come from ``losses``).  The in-batch head passes a view
        whose graph contains no collective (per-group partial backwards must not issue rank-dependent collectives) --
        see models.BertDot_InBatch_NLL_LN."""
        if not self.training:
            raise RuntimeError("iDROLoss.forward is only defined in training mode (as in the reference, where "
                               "gdro_counts_agg is undefined otherwise: dro_loss.py:222-226)")
        sums, counts = ops.group_stats(losses, g, self.n_groups)
        means = sums / (counts + (counts == 0).float())
        robust_loss = (means * self.h_fun.clone()).sum()  # pre-update weights; h_fun itself is updated in place below

        mask = (counts > 0).float()
        params = self._params(model)
        gmeans = means
        if grad_losses is not None:
            gsums, _ = ops.group_stats(grad_losses, g, self.n_groups)
            gmeans = gsums / (counts + (counts == 0).float())
        if sample_towers and self.grouped_wgrad and len(params) > 0:
            all_grads = self._get_grad_grouped(params, gmeans, counts, g, sample_towers)
        else:
            all_grads = self._get_grad(params, gmeans, counts)
        gram = self._gram(all_grads)
        with torch.no_grad():
            norm = torch.sqrt(torch.diagonal(gram).clamp_min(0)).unsqueeze(-1)  # ||G_g||
            denom = 1e-12 + norm
            RTG = gram / (denom * denom.t())
            _gl = torch.pow(means.detach().unsqueeze(-1), self.alpha)
            RTG = torch.mm(_gl, _gl.t()) * RTG
            _exp = self.rho * torch.mean(RTG, dim=0)
            _exp = _exp * mask
            _exp = _exp - _exp.max()
            weight = torch.exp(_exp)
            h = torch.pow(self.h_fun, self.ema) * weight * (counts != 0).float()
            h = h / h.sum()
            self.h_fun.copy_(torch.clamp(h, min=self.eps))
        return robust_loss, means.detach(), counts.detach()
