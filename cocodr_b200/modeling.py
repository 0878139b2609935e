"""COCO pre-training model, same surface as the reference's ``COCO/modeling.py``:
``CondenserForPretraining`` (:34-131) and ``CoCondenserForPretraining`` (:162-248) -- constructor
``(bert, model_args, data_args, train_args)``, attributes ``lm`` / ``c_head`` / ``co_target``,
``forward(model_input, labels, ...)`` returning the scalar loss, ``mlm_loss``, ``compute_contrastive_loss``,
``gather_tensors``, ``from_pretrained`` / ``from_config`` / ``save_pretrained`` (backbone via HF + head-only
``model.pt``), registered in ``CONDENSER_TYPE_MAP['bert']`` like run_coco_pre_training.py:46-48.

``self.lm`` stays a HuggingFace ``BertForMaskedLM`` as the parameter container (state-dict keys ``lm.*``,
``c_head.{i}.*`` unchanged); every FLOP of ``forward`` runs on the sm_100a kernels: the backbone through
``cocodr_b200.bert.BertModel`` (with hidden-state taps), the Condenser head layers through the same per-layer
operator, the two MLM losses through the masked-row MLM head (K14: rows with label != -100 are gathered first,
so the vocabulary GEMMs and the 30 522-way softmax touch ~15 % of the positions), the sequence-contrastive
loss through K9 (``ops.coco_contrastive``).
"""
import os
import warnings

import torch
import torch.distributed as dist
from torch import nn
from transformers import AutoModelForMaskedLM
from transformers.models.bert.modeling_bert import BertLayer

from . import ops
from .bert import BertModel, key_bias_from_mask, run_layer


class CondenserForPretraining(nn.Module):
    def __init__(self, bert, model_args, data_args, train_args):
        super().__init__()
        self.lm = bert
        self.c_head = nn.ModuleList([BertLayer(bert.config) for _ in range(model_args.n_head_layers)])
        self.c_head.apply(self.lm._init_weights)
        self.cross_entropy = nn.CrossEntropyLoss()
        self.cross_entropy2 = nn.CrossEntropyLoss()
        self.model_args = model_args
        self.train_args = train_args
        self.data_args = data_args
        object.__setattr__(self, "_c_shadows", [ops.LayerShadow() for _ in range(model_args.n_head_layers)])
        object.__setattr__(self, "_mlm_shadow", ops.MLMShadow())

    # ---- pieces -------------------------------------------------------------------------------
    def _backbone(self):
        if not isinstance(self.lm.bert, BertModel):
            BertModel.adopt(self.lm.bert)
        return self.lm.bert

    def _encode(self, model_input):
        """-> (cls fp32 [B,H], last hidden, hidden taps) in the internal fp16 [T,H] layout."""
        return self._backbone().encode(model_input['input_ids'], model_input.get('attention_mask'),
                                       token_type_ids=model_input.get('token_type_ids'), want_hidden=True)

    def _head(self, last, hidden, model_input):
        """Condenser head (modeling.py:76-85): [CLS of the last layer ; tokens of layer skip_from] -> c_head."""
        ids = model_input['input_ids']
        n_seq, L = ids.shape
        skip = hidden[self.model_args.skip_from]
        is_cls = (torch.arange(n_seq * L, device=ids.device) % L == 0).unsqueeze(1)
        h = torch.where(is_cls, last, skip)
        kb = key_bias_from_mask(model_input.get('attention_mask'))
        cfg = self.lm.config
        # the head layers are ordinary BertLayers of THIS module: they drop in train() mode even though the reference
        # puts the backbone in eval() (COCO/modeling.py:198, 216-220); their sites continue the backbone's numbering
        drop = self._backbone().dropout_spec(self.training, ids.device)
        for i, (layer, shadow) in enumerate(zip(self.c_head, self._c_shadows)):
            h = run_layer(layer, shadow, h, kb, n_seq, L, cfg, drop=drop, layer_index=cfg.num_hidden_layers + i)
        return h

    # Fraction of the positions the MLM head is sized for when set (e.g. 0.25 for 15 % masking): the masked rows are
    # then gathered into a FIXED-size buffer (unused slots carry label -100 and contribute nothing), no host
    # synchronisation sizes the GEMMs and the whole step can be captured into a CUDA graph.  None = exact dynamic size
    # (one host sync per step).  ``mlm_overflow`` counts the masked positions that did not fit (must stay 0).
    mlm_capacity = None

    def _mlm_rows(self, labels):
        """-> (row indices, their labels, number of valid rows or None)."""
        lab = labels.reshape(-1)
        if self.mlm_capacity is None:
            idx = torch.nonzero(lab != -100).flatten()  # (one host sync: the row count sizes the GEMMs)
            return idx, lab.index_select(0, idx), None
        cap = min(lab.numel(), (int(self.mlm_capacity * lab.numel()) + 63) // 64 * 64)
        valid = lab != -100
        count = valid.sum()
        idx = torch.nonzero_static(valid, size=cap, fill_value=0).flatten()
        row_labels = torch.where(torch.arange(cap, device=lab.device) < count, lab.index_select(0, idx),
                                 torch.full((cap,), -100, dtype=lab.dtype, device=lab.device))
        if not hasattr(self, "mlm_overflow") or self.mlm_overflow.device != lab.device:
            object.__setattr__(self, "mlm_overflow", torch.zeros((), dtype=torch.int64, device=lab.device))
        self.mlm_overflow += (count - cap).clamp(min=0)
        return idx, row_labels, count

    def _mlm(self, hidden_internal, idx, row_labels, count=None):
        """mean CE over the masked rows == CrossEntropyLoss()(scores.view(-1, V), labels.view(-1)) (:87-93)."""
        p = self.lm.cls.predictions
        if idx.numel() == 0:
            return hidden_internal.new_zeros((), dtype=torch.float32) * float("nan")
        rows = ops.GatherRows.apply(hidden_internal, idx)
        per_row = ops.MLMHead.apply(rows, row_labels, p.transform.dense.weight, p.transform.dense.bias,
                                    p.transform.LayerNorm.weight, p.transform.LayerNorm.bias, p.decoder.weight, p.bias,
                                    self._mlm_shadow, float(self.lm.config.layer_norm_eps))
        if count is None:
            return per_row.mean()
        return per_row.sum() / count.to(per_row.dtype)  # slots with label -100 are exact zeros (0 / 0 = nan like the reference)

    def mlm_loss(self, hiddens, labels):
        """Reference signature (:87-93): ``hiddens`` fp32 [B, L, H] as returned to callers."""
        idx, row_labels, count = self._mlm_rows(labels)
        return self._mlm(ops.FloatToHidden.apply(hiddens), idx, row_labels, count)

    def forward(self, model_input, labels, groups=None, **kwargs):
        cls, last, hidden = self._encode(model_input)
        idx, row_labels, count = self._mlm_rows(labels)
        loss = self._mlm(self._head(last, hidden, model_input), idx, row_labels, count)
        if self.model_args.late_mlm:
            loss = loss + self._mlm(last, idx, row_labels, count)
        return loss

    # ---- persistence (modeling.py:96-131) -----------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_args, data_args, train_args, *args, **kwargs):
        hf_model = AutoModelForMaskedLM.from_pretrained(*args, **kwargs)
        model = cls(hf_model, model_args, data_args, train_args)
        path = args[0]
        if os.path.exists(os.path.join(path, 'model.pt')):
            model_dict = torch.load(os.path.join(path, 'model.pt'), map_location="cpu")
            model.load_state_dict(model_dict, strict=False)
        return model

    @classmethod
    def from_config(cls, config, model_args, data_args, train_args):
        hf_model = AutoModelForMaskedLM.from_config(config)
        return cls(hf_model, model_args, data_args, train_args)

    def save_pretrained(self, output_dir):
        self.lm.save_pretrained(output_dir)
        model_dict = self.state_dict()
        hf_weight_keys = [k for k in model_dict.keys() if k.startswith('lm')]
        warnings.warn(f'omiting {len(hf_weight_keys)} transformer weights')
        for k in hf_weight_keys:
            model_dict.pop(k)
        torch.save(model_dict, os.path.join(output_dir, 'model.pt'))
        torch.save([self.data_args, self.model_args, self.train_args], os.path.join(output_dir, 'args.pt'))


class CoCondenserForPretraining(CondenserForPretraining):
    def __init__(self, bert, model_args, data_args, train_args):
        super().__init__(bert, model_args, data_args, train_args)
        effective_bsz = train_args.per_device_train_batch_size * self._world_size() * 2
        target = torch.arange(effective_bsz, dtype=torch.long).view(-1, 2).flip([1]).flatten().contiguous()
        self.register_buffer('co_target', target)
        self.train_method = getattr(data_args, "train_method", None)  # (undefined in the reference's own arguments.py)

    def _gather_tensor(self, t):
        all_tensors = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(all_tensors, t)
        all_tensors[self.train_args.local_rank] = t
        return all_tensors

    def gather_tensors(self, *tt):
        return [torch.cat(self._gather_tensor(t)) for t in tt]

    def forward(self, model_input, labels, groups=None, grad_cache=None, chunk_offset=None):
        """modeling.py:192-235: Condenser MLM loss (+ backbone MLM loss when late_mlm) + mean contrastive loss."""
        if grad_cache is not None:
            raise NotImplementedError("the GradCache path (COCO/trainer.py:110-192) is out of scope (SURVEY §2.1 #5)")
        self.lm.eval()
        cls, last, hidden = self._encode(model_input)
        if self.train_args.local_rank > -1 and dist.is_available() and dist.is_initialized():
            co_cls_hiddens = self.gather_tensors(cls.contiguous())[0]
        else:
            co_cls_hiddens = cls
        idx, row_labels, count = self._mlm_rows(labels)
        loss = self._mlm(self._head(last, hidden, model_input), idx, row_labels, count)
        if self.model_args.late_mlm:
            loss = loss + self._mlm(last, idx, row_labels, count)
        co_loss = self.compute_contrastive_loss(co_cls_hiddens).mean()
        return loss + co_loss

    @staticmethod
    def _world_size():
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def compute_contrastive_loss(self, co_cls_hiddens):
        """modeling.py:244-248: S = E E^T, diag = -inf, CE(S, i^1) * world -> [N]."""
        return ops.coco_contrastive(co_cls_hiddens, loss_scale=float(self._world_size()))


CONDENSER_TYPE_MAP = {'bert': CoCondenserForPretraining}
