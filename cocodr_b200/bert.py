"""Drop-in ``BertModel`` whose forward/backward run on the sm_100a kernels.

The reference never implements the encoder: it calls HuggingFace ``BertModel`` through ``self.bert(...)``
(ANCE/model/models.py:225-229) / ``AutoModelForMaskedLM`` (COCO/modeling.py:199-204).  This class *is* a
``transformers.BertModel`` as far as construction, ``from_pretrained`` / ``save_pretrained``, parameter names
and ``state_dict`` go (the HF-named fp32 parameters stay the source of truth) -- only ``forward`` is replaced:
embeddings+LN (K1), packed-QKV GEMM (K2), fused attention (K3), output / FFN GEMMs with fused
bias / GELU / residual epilogues + LayerNorm (K4-K6) and fp32 CLS pooling (K7), all through the C ABI.

Dropout: in ``train()`` mode the four nn.Dropout sites of HF BERT (embeddings, attention probabilities, both dense
outputs of every layer; ``hidden_dropout_prob`` / ``attention_probs_dropout_prob``) are applied inside the kernels
with counter-based masks (``ops.DropSpec``; include/cocodr_b200.h ``cdr_dropout``): every encoder pass advances a
device-resident (seed, offset) state, so a CUDA-graph replay draws fresh masks and the backward regenerates the
forward's.  ``set_dropout_seed`` fixes the stream; ``eval()`` or p = 0 turns it off.
"""
import torch
import torch.distributed as dist
from transformers import BertModel as _HFBertModel
from transformers.modeling_outputs import BaseModelOutputWithPoolingAndCrossAttentions

from . import ops


def key_bias_from_mask(attention_mask):
    """HF get_extended_attention_mask semantics on the key axis: (1 - mask) * finfo(float32).min, [B, L] fp32."""
    if attention_mask is None:
        return None
    return (1.0 - attention_mask.to(torch.float32)) * torch.finfo(torch.float32).min


def layer_params(layer):
    """The 16 HF parameters of one BertLayer in the order ops.BertLayerFn expects."""
    a, o = layer.attention, layer.output
    return (a.self.query.weight, a.self.query.bias, a.self.key.weight, a.self.key.bias, a.self.value.weight,
            a.self.value.bias, a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight,
            a.output.LayerNorm.bias, layer.intermediate.dense.weight, layer.intermediate.dense.bias, o.dense.weight,
            o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias)


def shadow_sources(layer):
    """(wq, bq, wk, bk, wv, bv, wo, wi, wo2): the parameters that have fp16 / packed shadows."""
    p = layer_params(layer)
    return (p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[10], p[12])


def check_config(config):
    if config.hidden_size % config.num_attention_heads or config.hidden_size // config.num_attention_heads != 64:
        raise RuntimeError("cocodr_b200 BERT kernels need head_dim == 64")
    if getattr(config, "hidden_act", "gelu") != "gelu":
        raise RuntimeError("cocodr_b200 BERT kernels implement exact-erf GELU only (hidden_act='gelu')")
    if getattr(config, "position_embedding_type", "absolute") not in (None, "absolute"):
        raise RuntimeError("cocodr_b200 BERT kernels implement absolute position embeddings only")


def run_last_layer_cls(layer, shadow, x, key_bias, n_seq, L, config, drop=None, layer_index=0):
    """Last BertLayer when only the [CLS] embedding is consumed: FFN / LayerNorms / output projection on n_seq rows."""
    return ops.BertLastLayerCLSFn.apply(x, key_bias, *layer_params(layer), shadow, n_seq, L,
                                        config.num_attention_heads, float(config.layer_norm_eps), drop, layer_index)


def run_layer(layer, shadow, x, key_bias, n_seq, L, config, emit_cls=False, drop=None, layer_index=0):
    """One BertLayer (HF module used as the parameter container) on internal fp16 [T, H] activations.
    ``drop`` (ops.DropSpec) / ``layer_index`` select the layer's dropout sites; None = no dropout."""
    return ops.BertLayerFn.apply(x, key_bias, *layer_params(layer), shadow, n_seq, L, config.num_attention_heads,
                                 float(config.layer_norm_eps), emit_cls, drop, layer_index)


class DropoutState:
    """Device-resident (seed, offset) of the counter-based dropout.  ``next()`` advances the offset ON THE DEVICE (so a
    captured CUDA graph draws new masks at every replay) and returns a snapshot that the pass's forward and backward
    kernels both read -- later passes do not disturb it."""

    def __init__(self):
        self.seed = None
        self.state = None

    def set_seed(self, seed):
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        self.state = None

    def next(self, device):
        if self.state is None or self.state.device != device:
            seed = self.seed
            if seed is None:  # default: torch's seed, decorrelated across data-parallel ranks
                rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
                seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF
            self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self.state[1] += 1
        return self.state.clone()


class BertModel(_HFBertModel):
    """``transformers.BertModel`` with the compute path replaced (see module docstring)."""

    def __init__(self, config, add_pooling_layer=True):
        super().__init__(config, add_pooling_layer=add_pooling_layer)
        check_config(config)
        self._cdr_init()

    def _cdr_init(self):
        object.__setattr__(self, "_shadow_set", ops.ShadowSet(len(self.encoder.layer)))
        object.__setattr__(self, "_shadows", self._shadow_set.layers)
        if not hasattr(self, "_dropout_state"):
            object.__setattr__(self, "_dropout_state", DropoutState())

    def set_dropout_seed(self, seed):
        """Fix the dropout stream (default: torch.initial_seed(), decorrelated per rank); offset restarts at 0."""
        self._dropout_state.set_seed(seed)

    def dropout_spec(self, training, device):
        """ops.DropSpec of the next pass (advances the device-side offset), or None in eval mode / with p = 0."""
        ph, pa = float(self.config.hidden_dropout_prob), float(self.config.attention_probs_dropout_prob)
        if not training or (ph <= 0.0 and pa <= 0.0):
            return None
        if not (0.0 <= ph < 1.0 and 0.0 <= pa < 1.0):
            raise RuntimeError("dropout probabilities must be in [0, 1)")
        return ops.DropSpec(self._dropout_state.next(device), ph, pa)

    @classmethod
    def adopt(cls, hf_bert):
        """Turn an already-constructed HF BertModel into this class in place (keeps its parameters)."""
        check_config(hf_bert.config)
        hf_bert.__class__ = cls
        hf_bert._cdr_init()
        return hf_bert

    def shadow_map(self):
        """{parameter: (shadow view, is_f32)} of every parameter that has an fp16 / packed operand copy; a fused
        optimizer (optim.py) writes these together with the parameters."""
        if len(self._shadows) != len(self.encoder.layer):
            self._cdr_init()
        srcs = [shadow_sources(layer) for layer in self.encoder.layer]
        return {param: (dst, is_f32) for param, dst, is_f32 in self._shadow_set.pairs(srcs)}

    # ------------------------------------------------------------------------------------------
    def _check_inputs(self, input_ids, token_type_ids, position_ids, inputs_embeds):
        if inputs_embeds is not None or input_ids is None:
            raise NotImplementedError("cocodr_b200.BertModel takes input_ids (inputs_embeds is not supported)")
        if not input_ids.is_cuda:
            raise RuntimeError("cocodr_b200.BertModel needs CUDA inputs: there is no CPU / eager fallback")
        if position_ids is not None:
            raise NotImplementedError("custom position_ids are not supported (positions are 0..L-1)")
        if token_type_ids is not None and bool(token_type_ids.any()):
            raise NotImplementedError("non-zero token_type_ids are not supported (the reference never passes them)")
        if input_ids.shape[1] > self.config.max_position_embeddings:
            raise RuntimeError("sequence longer than max_position_embeddings")

    cls_only_last_layer = True  # encode_cls(): run the last layer's FFN / LayerNorms on the [CLS] rows only

    def encode(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
               want_hidden=False, last_layer=None, cls_only=False):
        """Internal entry: returns (cls fp32 [B,H], last hidden fp16 [T,H], [hidden fp16] * (layers+1) or None).
        ``cls_only``: the caller only reads ``cls`` -- the last hidden state is not produced (None)."""
        self._check_inputs(input_ids, token_type_ids, position_ids, inputs_embeds)
        if len(self._shadows) != len(self.encoder.layer):
            self._cdr_init()
        n_seq, L = input_ids.shape
        e = self.embeddings
        self._shadow_set.refresh([shadow_sources(layer) for layer in self.encoder.layer])
        drop = self.dropout_spec(self.training, input_ids.device)
        drop = ops.prefill_attn_bits(drop, len(self.encoder.layer), n_seq, self.config.num_attention_heads, L,
                                     input_ids.device)
        x = ops.EmbedLN.apply(input_ids.long(), e.word_embeddings.weight, e.position_embeddings.weight,
                              e.token_type_embeddings.weight, e.LayerNorm.weight, e.LayerNorm.bias,
                              float(self.config.layer_norm_eps), drop)
        kb = key_bias_from_mask(attention_mask)
        hidden = [x] if want_hidden else None
        cls = None
        n_layers = len(self.encoder.layer)
        for i, layer in enumerate(self.encoder.layer):
            if i == n_layers - 1 and cls_only and not want_hidden and self.cls_only_last_layer:
                cls = run_last_layer_cls(layer, self._shadows[i], x, kb, n_seq, L, self.config, drop, i)
                x = None
            elif i == n_layers - 1:
                x, cls = run_layer(layer, self._shadows[i], x, kb, n_seq, L, self.config, emit_cls=True, drop=drop,
                                   layer_index=i)
            else:
                x = run_layer(layer, self._shadows[i], x, kb, n_seq, L, self.config, drop=drop, layer_index=i)
            if want_hidden:
                hidden.append(x)
        return cls, x, hidden

    def encode_cls(self, input_ids, attention_mask=None):
        """``self(input_ids, attention_mask)[0][:, 0]`` without materialising the fp32 hidden states."""
        return self.encode(input_ids, attention_mask, cls_only=True)[0]

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                output_hidden_states=None, return_dict=None, **kwargs):
        n_seq, L = input_ids.shape
        want_hidden = bool(output_hidden_states) or bool(getattr(self.config, "output_hidden_states", False))
        cls, last, hidden = self.encode(input_ids, attention_mask, token_type_ids, position_ids, inputs_embeds,
                                        want_hidden=want_hidden)
        seq_out = ops.HiddenToFloat.apply(last, n_seq, L)
        # row 0 of the caller-visible tensor is the fp32 CLS straight from the last LayerNorm
        seq_out = torch.cat([cls.unsqueeze(1), seq_out[:, 1:]], dim=1)
        hs = tuple(ops.HiddenToFloat.apply(h, n_seq, L) for h in hidden) if want_hidden else None
        out = BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=seq_out, pooler_output=None,
                                                           hidden_states=hs)
        if return_dict is False:
            return out.to_tuple()
        return out
