"""ANCE-stage bi-encoder, same surface as the reference's ``ANCE/model/models.py``:
``EmbeddingMixin`` (:41-77), ``NLL.forward_model`` (:80-115), ``BertDot_NLL_LN`` (:194-290) and the
``MSMarcoConfigDict`` registry (:418-445, key ``rdot_nll_condenser``) -- constructor, attributes
(``bert``, ``classifier``, ``embeddingHead``, ``norm``, ``total``, ``correct``, ``dro_type``, ``n_groups``,
``accum_loss``, ``accum_group_loss``), methods (``query_emb``, ``body_emb``, ``forward``, ``add_group_loss``,
``output_state``, ``gather_tensors``) and state-dict keys are the reference's, so
``MSMarcoConfigDict[name].model_class.from_pretrained(path, config=...)`` (run_ann.py:896-901) yields a
drop-in whose encoder and losses run on the sm_100a kernels.

Differences that are deliberate and documented:
  * q / p+ / p- towers of equal length run as ONE encoder launch (the reference runs three sequential
    passes over the same weights, models.py:97-99);
  * the per-forward ``all_reduce(train_size)`` + ``.item()`` bookkeeping (:256-258) becomes arithmetic
    (batch * world size) and the (2 + 2G) ``.item()`` syncs of the meters (:269-271) become one transfer;
  * ``BertDot_InBatch_NLL_LN`` adds the q x p in-batch InfoNCE over all-gathered passages that
    BASELINE.json's configs describe (K9'); the reference defines the gather helper (:282-290) but
    never calls it.
"""
import logging

import torch
import torch.distributed as dist
from torch import nn
from transformers import BertConfig, BertForSequenceClassification, BertTokenizer

from . import ops, peer
from .bert import BertModel
from .dro_loss import AverageMeter, DROGreedyLoss, MeterBank, iDROLoss

logger = logging.getLogger(__name__)


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class AllGatherWithGrad(torch.autograd.Function):
    """all_gather whose backward returns every rank's gradient for the local slice (reduce-scatter), so the
    multi-GPU loss equals the single-process loss on the concatenated batch once DDP averages over ranks."""

    @staticmethod
    def forward(ctx, t):
        W = _world()
        out = torch.empty(W * t.shape[0], *t.shape[1:], dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    @staticmethod
    def backward(ctx, g):
        W = _world()
        out = torch.empty(g.shape[0] // W, *g.shape[1:], dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g.contiguous())
        return out


def gather_with_grad(t):
    return AllGatherWithGrad.apply(t) if _world() > 1 else t


class EmbeddingMixin:
    """models.py:41-77."""

    def __init__(self, model_argobj):
        self.use_mean = False if model_argobj is None else model_argobj.use_mean

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding, nn.Conv1d)):
            module.weight.data.normal_(mean=0.0, std=0.02)

    def masked_mean(self, t, mask):
        s = torch.sum(t * mask.unsqueeze(-1).float(), axis=1)
        d = mask.sum(axis=1, keepdim=True).float()
        return s / d

    def masked_mean_or_first(self, emb_all, mask):
        assert isinstance(emb_all, tuple)
        if self.use_mean:
            return self.masked_mean(emb_all[0], mask)
        return emb_all[0][:, 0]

    def query_emb(self, input_ids, attention_mask):
        raise NotImplementedError("Please Implement this method")

    def body_emb(self, input_ids, attention_mask):
        raise NotImplementedError("Please Implement this method")


class NLL(EmbeddingMixin):
    def _towers(self, q_ids, q_mask, a_ids, a_mask, b_ids=None, b_mask=None):
        """CLS embeddings of 2 or 3 towers; one fused encoder launch when the towers share a length."""
        ids, masks = [q_ids, a_ids], [q_mask, a_mask]
        if b_ids is not None:
            ids.append(b_ids)
            masks.append(b_mask)
        if all(t.shape == ids[0].shape for t in ids):
            emb = self.query_emb(torch.cat(ids, 0), torch.cat(masks, 0))
            return emb.chunk(len(ids), 0)
        embs = [self.query_emb(q_ids, q_mask)] + [self.body_emb(i, m) for i, m in zip(ids[1:], masks[1:])]
        return tuple(embs)

    def forward_model(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None, input_ids_b=None,
                      attention_mask_b=None, is_query=True, group_ids=None):
        """models.py:80-115: embeddings only, or (loss[B], accs[B], logits[B,2]) of the triplet NLL."""
        if input_ids_b is None and is_query:
            return self.query_emb(query_ids, attention_mask_q)
        elif input_ids_b is None:
            return self.body_emb(query_ids, attention_mask_q)
        q_embs, a_embs, b_embs = self._towers(query_ids, attention_mask_q, input_ids_a, attention_mask_a, input_ids_b,
                                              attention_mask_b)
        loss, accs, logit_matrix = ops.pair_nll(q_embs, a_embs, b_embs)
        # loss i reads the embeddings of triplet i only; when the three towers ran as one pass (sequence s = sample
        # s % B) iDRO may take its group gradients from one shared backward (dro_loss.iDROLoss._get_grad_grouped)
        self._sample_towers = 3 if query_ids.shape == input_ids_a.shape == input_ids_b.shape else 0
        return loss, accs, logit_matrix


class BertDot_NLL_LN(NLL, BertForSequenceClassification):
    """models.py:194-290 with ``self.bert`` running on the B200 kernels."""

    def __init__(self, config, model_argobj=None):
        NLL.__init__(self, model_argobj)
        BertForSequenceClassification.__init__(self, config)
        BertModel.adopt(self.bert)
        self.embeddingHead = nn.Linear(config.hidden_size, 768)
        self.norm = nn.LayerNorm(768)
        self.use_moco = False
        self.apply(self._init_weights)
        self.total = 0
        self.correct = 0
        self.prob_diff = []
        self.dro_type = 'erm'

    def add_group_loss(self, args, n_groups, dro_type, alpha, eps, ema=0.1, rho=0.1, weight_ema=True):
        if dro_type == 'dro-greedy':
            self.dro_type = dro_type
            self.loss = DROGreedyLoss(args, n_groups, alpha, eps, ema, weight_ema)
        elif dro_type == 'idro':
            self.dro_type = dro_type
            self.loss = iDROLoss(args, n_groups, alpha, eps, ema, rho)
        else:
            logger.info("Warning! No training strategy selected")
        if hasattr(self, "loss"):
            self.loss.to(self.bert.embeddings.word_embeddings.weight.device)
        self.n_groups = n_groups
        # same attributes as the reference (AverageMeter API), backed by device accumulators: no host sync per step
        self._meters = MeterBank(1 + n_groups)
        meters = self._meters.meters()
        self.accum_loss = meters[0]
        self.accum_group_loss = meters[1:]

    def query_emb(self, input_ids, attention_mask):
        """``self.bert(input_ids, attention_mask)[0][:, 0]`` (models.py:225-229), fp32 [B, H]."""
        if not isinstance(self.bert, BertModel):
            BertModel.adopt(self.bert)
        return self.bert.encode_cls(input_ids, attention_mask)

    def body_emb(self, input_ids, attention_mask):
        return self.query_emb(input_ids, attention_mask)

    def _route_loss(self, loss, train_acc, logits, group_ids, weights):
        """models.py:256-273: bookkeeping, ERM mean or DRO loss, meters."""
        self.total += int(train_acc.shape[0]) * _world()
        if group_ids is None:
            if weights is not None:
                loss = loss * weights
            return loss.mean(), train_acc, logits
        if self.dro_type == 'idro':
            robust_loss, group_losses, group_counts = self.loss(self.bert, loss, group_ids,
                                                                sample_towers=getattr(self, "_sample_towers", 0),
                                                                grad_losses=getattr(self, "_idro_grad_losses", None))
            self._idro_grad_losses = None
        else:
            robust_loss, group_losses, group_counts = self.loss(loss, group_ids, weights)
        # the reference's accum_loss.update(robust.item(), B) / accum_group_loss[i].update(gl[i].item(), gc[i].item())
        # (:269-271), accumulated on the device and fetched when a meter is read (logging time)
        self._meters.add(torch.cat([robust_loss.detach().reshape(1), group_losses]),
                         torch.cat([group_counts.new_full((1,), float(loss.size(0))), group_counts]))
        return robust_loss, train_acc, group_losses, group_counts

    def forward(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None, input_ids_b=None,
                attention_mask_b=None, is_query=True, group_ids=None, weights=None):
        out = self.forward_model(query_ids, attention_mask_q, input_ids_a, attention_mask_a, input_ids_b,
                                 attention_mask_b, is_query, group_ids)
        if not isinstance(out, tuple):
            return out
        loss, train_acc, logits = out
        return self._route_loss(loss, train_acc, logits, group_ids, weights)

    def output_state(self):
        h = self.loss.h_fun.detach().cpu().numpy()
        h_fun = {self.loss.id2group[str(i)]: h[i] for i in range(self.n_groups)}
        sum_loss = {self.loss.id2group[str(i)]: self.accum_group_loss[i].avg for i in range(self.n_groups)}
        return h_fun, sum_loss

    def _gather_tensor(self, t):
        all_tensors = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(all_tensors, t)
        all_tensors[_rank()] = t
        return all_tensors

    def gather_tensors(self, *tt):
        return [torch.cat(self._gather_tensor(t)) for t in tt]


class BertDot_InBatch_NLL_LN(BertDot_NLL_LN):
    """q x p in-batch InfoNCE with cross-GPU all-gathered passages (K9').

    ``loss_i = CE(q_i . P_all^T, rank*B + i)`` over the passages of every rank (hard negatives
    ``input_ids_b``, when given, are appended to the key set).  ERM / DRO-greedy routing of the per-sample
    losses is inherited unchanged.

    iDRO differentiates every group's mean loss w.r.t. the last layers, one partial backward per group
    (ANCE/model/dro_loss.py:192-204).  Through the gathered keys those backwards would issue collectives
    (reduce-scatter of the key gradients) a rank-dependent number of times -- a hang or a mismatch as soon as two ranks
    hold different groups.  The group gradients are therefore taken from a VIEW of the same loss values whose graph
    stays on this rank (the robust loss and the training gradient still use the fully differentiable gather):

      ``idro_group_grads = "local-batch"`` (default): keys of other ranks are constants, this rank's passages keep
          their graph -- exactly what the reference's own gather convention does (``all_tensors[rank] = t``,
          COCO/modeling.py:182-186).  Every local passage is a negative of every local query, so an activation-gradient
          row of the passage tower mixes all groups: K11 (one shared backward + grouped wgrad) cannot apply, and the
          gradients are G_present partial backwards through the last layers.
      ``idro_group_grads = "own-pair"``: in addition the in-batch negatives are constants -- sample i's loss view
          depends on (q_i, p_i) only, the same locality the reference's triplet loss has by construction
          (models.py:101-108) -- so K11 applies and the [G, P_last] matrix costs one partial backward.  This changes
          the gradient-similarity statistics that drive ``h_fun`` (the negatives' share of each group gradient is
          dropped), not the loss or its training gradient."""

    peer_gather = False  # enable_peer_gather(): exchange the passage embeddings through peer memory, not NCCL
    idro_group_grads = "local-batch"

    def enable_peer_gather(self, on=True):
        """Multi-GPU, one node: all-gather the passage CLS embeddings by letting the last LayerNorm kernel store them
        into every rank's symmetric buffer over NVLink (peer.py) and reduce-scatter their gradients the same way."""
        self.peer_gather = bool(on)
        self._xchg = None
        return self

    def _peer_exchange(self, q_ids, a_ids, b_ids):
        if not (self.peer_gather and _world() > 1) or b_ids is not None or q_ids.shape != a_ids.shape:
            return None
        if not torch.is_grad_enabled():
            return None
        B, H = a_ids.shape[0], self.config.hidden_size
        x = getattr(self, "_xchg", None)
        if x is None or x.n_rows != B or x.dim != H:
            x = self._xchg = peer.PeerExchange(B, H, a_ids.device)
        return x

    def forward_model(self, query_ids, attention_mask_q, input_ids_a=None, attention_mask_a=None, input_ids_b=None,
                      attention_mask_b=None, is_query=True, group_ids=None):
        if input_ids_a is None:
            return super().forward_model(query_ids, attention_mask_q, is_query=is_query)
        self._sample_towers = 0  # the in-batch loss of sample i reads every passage: group gradients do not separate
        xchg = self._peer_exchange(query_ids, input_ids_a, input_ids_b)
        if xchg is not None:
            # fused path: the last LayerNorm kernel writes the passage CLS rows into every rank's gather buffer
            xchg.begin_step()
            ops.CLS_PUSH = (xchg, query_ids.shape[0])
            try:
                embs = self._towers(query_ids, attention_mask_q, input_ids_a, attention_mask_a)
            finally:
                ops.CLS_PUSH = None
            q_embs = embs[0]
            B = q_embs.shape[0]
            keys = peer.gather_passages(embs[1], xchg)
        else:
            embs = self._towers(query_ids, attention_mask_q, input_ids_a, attention_mask_a, input_ids_b,
                                attention_mask_b)
            q_embs = embs[0]
            B = q_embs.shape[0]
            keys = gather_with_grad(embs[1])
        offset = _rank() * B
        if len(embs) == 3:
            keys = torch.cat([keys, gather_with_grad(embs[2])], 0)
        loss = ops.qp_infonce(q_embs, keys, row_offset=offset)
        self._idro_grad_losses = None
        if group_ids is not None and self.dro_type == 'idro' and torch.is_grad_enabled():
            if self.idro_group_grads == "own-pair" and len(embs) == 2:
                self._idro_grad_losses = ops.OwnPairCE.apply(q_embs, embs[1], keys, offset)
                if query_ids.shape == input_ids_a.shape:
                    self._sample_towers = 2  # loss view i reads sequences {i, B + i} of the fused pass only
            elif self.idro_group_grads in ("own-pair", "local-batch"):
                kd = keys.detach()
                n_all = kd.shape[0] // (len(embs) - 1)  # gathered passages (| gathered hard negatives)
                parts = [kd[:offset], embs[1], kd[offset + B:n_all]]
                if len(embs) == 3:
                    parts += [kd[n_all:n_all + offset], embs[2], kd[n_all + offset + B:]]
                self._idro_grad_losses = ops.qp_infonce(q_embs, torch.cat(parts, 0), row_offset=offset)
            else:
                raise ValueError(f"idro_group_grads must be 'local-batch' or 'own-pair', not {self.idro_group_grads!r}")
        with torch.no_grad():
            pos = (q_embs * embs[1]).sum(-1)
            neg = (q_embs * embs[2]).sum(-1) if len(embs) == 3 else pos
            logits = torch.stack([pos, neg], 1)
            accs = torch.argmax(logits, dim=1)
        return loss, accs, logits


default_process_fn = None  # tokenisation lives outside the hot path


class MSMarcoConfig:
    def __init__(self, name, model, process_fn=default_process_fn, use_mean=True, tokenizer_class=BertTokenizer,
                 config_class=BertConfig):
        self.name = name
        self.process_fn = process_fn
        self.model_class = model
        self.use_mean = use_mean
        self.tokenizer_class = tokenizer_class
        self.config_class = config_class


configs = [
    MSMarcoConfig(name="rdot_nll_condenser", model=BertDot_NLL_LN, tokenizer_class=BertTokenizer,
                  config_class=BertConfig, use_mean=False),
    MSMarcoConfig(name="rdot_nll_condenser_inbatch", model=BertDot_InBatch_NLL_LN, tokenizer_class=BertTokenizer,
                  config_class=BertConfig, use_mean=False),
]

MSMarcoConfigDict = {cfg.name: cfg for cfg in configs}
