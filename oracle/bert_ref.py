"""fp32 torch restatement of the BERT encoder the reference drives through HF.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference never implements the encoder itself; it calls HuggingFace
``BertModel`` (ANCE/model/models.py:225-229, COCO/modeling.py:199-204).  The
arithmetic restated here follows the installed transformers 5.5.0
``models/bert/modeling_bert.py``:
  BertEmbeddings.forward            :76-112   word + type0 + pos -> LayerNorm
  eager_attention_forward           :115-140  softmax(QK^T/sqrt(d) + mask) V
  BertSelfAttention.forward         :168-207
  BertSelfOutput.forward            :287-298  LN(x + dense(ctx))
  BertIntermediate.forward          :330-342  gelu_erf(dense(x))
  BertOutput.forward                :345-356  LN(x + dense(h))
Dropout: every function takes an optional ``drop`` (``oracle.dropout_ref.DropSpec``): the four nn.Dropout sites
(embeddings :110, attention probabilities :133, BertSelfOutput :296, BertOutput :354) then multiply by the
counter-based masks of ``oracle/dropout_ref.py`` -- the same masks the CUDA kernels regenerate -- so training-mode
parity is exact in the mask and fp16-vs-fp32 in the arithmetic.  ``drop=None`` is eval mode / p = 0.

All functions take a plain ``dict`` state (HF parameter names *without* the
``bert.`` prefix) so the same weights can be loaded into the reference, the
oracle and the CUDA drop-in.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # BertConfig.layer_norm_eps


def make_config(hidden=768, layers=12, heads=12, inter=3072, vocab=30522,
                max_pos=512, type_vocab=2):
    return dict(hidden=hidden, layers=layers, heads=heads, inter=inter,
                vocab=vocab, max_pos=max_pos, type_vocab=type_vocab)


def layer_param_names(i: int) -> List[str]:
    p = f"encoder.layer.{i}."
    return [p + s for s in (
        "attention.self.query.weight", "attention.self.query.bias",
        "attention.self.key.weight", "attention.self.key.bias",
        "attention.self.value.weight", "attention.self.value.bias",
        "attention.output.dense.weight", "attention.output.dense.bias",
        "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
        "intermediate.dense.weight", "intermediate.dense.bias",
        "output.dense.weight", "output.dense.bias",
        "output.LayerNorm.weight", "output.LayerNorm.bias")]


def param_shapes(cfg) -> Dict[str, tuple]:
    H, I = cfg["hidden"], cfg["inter"]
    shapes = {
        "embeddings.word_embeddings.weight": (cfg["vocab"], H),
        "embeddings.position_embeddings.weight": (cfg["max_pos"], H),
        "embeddings.token_type_embeddings.weight": (cfg["type_vocab"], H),
        "embeddings.LayerNorm.weight": (H,),
        "embeddings.LayerNorm.bias": (H,),
    }
    for i in range(cfg["layers"]):
        p = f"encoder.layer.{i}."
        for n in ("query", "key", "value"):
            shapes[p + f"attention.self.{n}.weight"] = (H, H)
            shapes[p + f"attention.self.{n}.bias"] = (H,)
        shapes[p + "attention.output.dense.weight"] = (H, H)
        shapes[p + "attention.output.dense.bias"] = (H,)
        shapes[p + "attention.output.LayerNorm.weight"] = (H,)
        shapes[p + "attention.output.LayerNorm.bias"] = (H,)
        shapes[p + "intermediate.dense.weight"] = (I, H)
        shapes[p + "intermediate.dense.bias"] = (I,)
        shapes[p + "output.dense.weight"] = (H, I)
        shapes[p + "output.dense.bias"] = (H,)
        shapes[p + "output.LayerNorm.weight"] = (H,)
        shapes[p + "output.LayerNorm.bias"] = (H,)
    return shapes


def synth_state(cfg, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoints offline).

    Matrices ~ N(0, std) like HF's init; LayerNorm weights/biases and linear
    biases get small non-trivial values so parity tests exercise them.
    Generated name-by-name in sorted order from one CPU generator so every
    consumer (reference, oracle, CUDA) reproduces the same tensors.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shp in sorted(param_shapes(cfg).items()):
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            t = std * torch.randn(shp, generator=g)
        out[name] = t
    return out


def synth_batch(n_seq: int, seq_len: int, vocab: int, seed: int, full: bool = False,
                mean_len: float = 0.6):
    """Synthetic token ids / right-padded masks (SURVEY.md §8d "Synthetic inputs")."""
    g = torch.Generator().manual_seed(seed)
    lo = min(1000, vocab // 2)
    ids = torch.randint(lo, vocab, (n_seq, seq_len), generator=g)
    if full:
        lens = torch.full((n_seq,), seq_len, dtype=torch.long)
    else:
        lens = (torch.rand(n_seq, generator=g) * seq_len * mean_len * 2).long().clamp(4, seq_len)
        lens[0] = seq_len  # always one full-length row
    pos = torch.arange(seq_len)[None, :]
    mask = (pos < lens[:, None]).long()
    ids[:, 0] = min(101, vocab - 2)
    ids[torch.arange(n_seq), lens - 1] = min(102, vocab - 1)
    ids = ids * mask  # pad id 0
    return ids, mask


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def embeddings_fwd(st, ids, cfg, drop=None):
    L = ids.shape[1]
    H = cfg["hidden"]
    # nn.Embedding(padding_idx=pad_token_id=0): the pad row receives no gradient (:58)
    x = (F.embedding(ids, st["embeddings.word_embeddings.weight"], padding_idx=0)
         + st["embeddings.token_type_embeddings.weight"][0][None, None, :]
         + st["embeddings.position_embeddings.weight"][:L][None, :, :])
    x = F.layer_norm(x, (H,), st["embeddings.LayerNorm.weight"],
                     st["embeddings.LayerNorm.bias"], LN_EPS)
    if drop is not None:
        m = drop.hidden(ids.shape[0] * L, H, 0)
        if m is not None:
            x = x * m.view(ids.shape[0], L, H)
    return x


def layer_fwd(st, prefix, x, add_mask, cfg, drop=None, layer_index=0):
    """One BertLayer.  ``prefix`` e.g. 'encoder.layer.3.' or 'c_head.0.'; ``layer_index`` numbers the dropout sites
    (4 * layer_index + 1 attention probabilities, + 2 attention output dense, + 3 FFN output dense)."""
    B, L, H = x.shape
    nh = cfg["heads"]
    d = H // nh
    lin = lambda t, n: F.linear(t, st[prefix + n + ".weight"], st[prefix + n + ".bias"])
    q = lin(x, "attention.self.query").view(B, L, nh, d).transpose(1, 2)
    k = lin(x, "attention.self.key").view(B, L, nh, d).transpose(1, 2)
    v = lin(x, "attention.self.value").view(B, L, nh, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5) + add_mask
    p = torch.softmax(s, dim=-1)
    ma = mb = mc = None
    if drop is not None:
        ma = drop.attn(B, nh, L, 4 * layer_index + 1)
        mb = drop.hidden(B * L, H, 4 * layer_index + 2)
        mc = drop.hidden(B * L, H, 4 * layer_index + 3)
    if ma is not None:
        p = p * ma
    ctx = torch.matmul(p, v).transpose(1, 2).reshape(B, L, H)
    so = lin(ctx, "attention.output.dense")
    if mb is not None:
        so = so * mb.view(B, L, H)
    x1 = F.layer_norm(x + so, (H,),
                      st[prefix + "attention.output.LayerNorm.weight"],
                      st[prefix + "attention.output.LayerNorm.bias"], LN_EPS)
    h = gelu_erf(lin(x1, "intermediate.dense"))
    fo = lin(h, "output.dense")
    if mc is not None:
        fo = fo * mc.view(B, L, H)
    x2 = F.layer_norm(x1 + fo, (H,),
                      st[prefix + "output.LayerNorm.weight"],
                      st[prefix + "output.LayerNorm.bias"], LN_EPS)
    return x2


def additive_mask(mask, dtype=torch.float32):
    """HF get_extended_attention_mask: (1-mask) * finfo.min on key positions."""
    return (1.0 - mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min


def encoder_fwd(st, ids, mask, cfg, output_hidden_states: bool = False, drop=None):
    """Returns last hidden state [B,L,H] (and list of L+1 hidden states)."""
    x = embeddings_fwd(st, ids, cfg, drop)
    am = additive_mask(mask, x.dtype)
    hs = [x]
    for i in range(cfg["layers"]):
        x = layer_fwd(st, f"encoder.layer.{i}.", x, am, cfg, drop, i)
        hs.append(x)
    return (x, hs) if output_hidden_states else x


def cls_embedding(st, ids, mask, cfg, drop=None):
    """BertDot_NLL_LN.query_emb / body_emb: last_hidden[:, 0] (models.py:225-232)."""
    return encoder_fwd(st, ids, mask, cfg, drop=drop)[:, 0]


def mlm_head_fwd(st, hidden, cfg, prefix="cls.predictions."):
    """BertOnlyMLMHead (modeling_bert.py:471-511): LN(gelu(dense(x))) @ E_word^T + b."""
    H = cfg["hidden"]
    t = F.linear(hidden, st[prefix + "transform.dense.weight"], st[prefix + "transform.dense.bias"])
    t = gelu_erf(t)
    t = F.layer_norm(t, (H,), st[prefix + "transform.LayerNorm.weight"],
                     st[prefix + "transform.LayerNorm.bias"], LN_EPS)
    return F.linear(t, st["embeddings.word_embeddings.weight"], st[prefix + "bias"])


def fwd_flops_per_seq(cfg, L):
    """Algorithmic GEMM+attention FLOPs of one sequence forward."""
    H, I = cfg["hidden"], cfg["inter"]
    return cfg["layers"] * (2 * L * H * 3 * H + 2 * 2 * L * L * H + 2 * L * H * H + 2 * 2 * L * H * I)
