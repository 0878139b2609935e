"""Counter-based dropout masks: numpy restatement of the mask generator the CUDA kernels use.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference trains with HF BERT's ``nn.Dropout`` modules (p = 0.1: ``BertEmbeddings`` :110, ``BertSelfAttention``
attention probabilities (eager_attention_forward :133), ``BertSelfOutput`` :296, ``BertOutput`` :354 of the installed
``modeling_bert.py``, reached through ``self.bert(...)`` in ANCE/model/models.py:226 under ``model.train()``,
ANCE/drivers/run_ann.py:320; COCO/modeling.py:216-220 for the ``c_head`` layers).  torch's generator stream cannot be
reproduced on another device or kernel decomposition, so -- like every fused-dropout implementation -- the CUDA path
draws its masks from a counter-based generator; parity is therefore defined as "same arithmetic given the same mask",
and this module regenerates the masks of ``include/cocodr_b200.h`` (``cdr_dropout``) on the CPU:

    Philox4x32-10,  counter = (group index, site, offset lo, offset hi),  key = (seed lo, seed hi)
    one call -> 4 x u32 = 8 x u16 (low half first) for 8 consecutive elements of a row;
    element kept iff u16 >= threshold = round(p * 65536); kept values are multiplied by 1 / (1 - p)
    group index of element (m, n) of a [rows, cols] activation: (m * row_mul) * (cols / 8) + n / 8
    group index of attention probability (seq, head, r, c):     ((seq * heads + head) * L + r) * 64 + c / 8

Known-answer check for the generator itself: ``philox4x32_10`` reproduces the Random123 test vectors (tests/).
Site numbering (cocodr_b200/bert.py): 0 = embeddings; layer l: 4l + 1 attention probabilities, 4l + 2 attention
output dense, 4l + 3 FFN output dense; Condenser head layer i continues as layer n_layers + i.
"""
import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counter words: uint32 arrays (broadcastable); key words: python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & MASK32, p0 & MASK32, n0, n2
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def threshold(p):
    return int(round(float(p) * 65536.0))


def keep_groups(gidx, site, seed, offset, p):
    """gidx: integer array of group indices -> bool array [..., 8] (True = kept)."""
    gidx = np.asarray(gidx, dtype=np.uint64)
    seed, offset = int(seed), int(offset)
    w = philox4x32_10(gidx, np.uint64(site), np.uint64(offset & 0xFFFFFFFF), np.uint64((offset >> 32) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    thr = threshold(p)
    halves = []
    for x in w:
        halves.append((x & np.uint32(0xFFFF)) >= thr)
        halves.append((x >> np.uint32(16)) >= thr)
    return np.stack(halves, axis=-1)


def hidden_mask(rows, cols, site, seed, offset, p, row_mul=1):
    """float32 torch [rows, cols] multiplier (0 or 1 / (1 - p)) of a dense-output / embedding dropout site."""
    assert cols % 8 == 0
    m = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(row_mul) * np.uint64(cols // 8)
    g = m + np.arange(cols // 8, dtype=np.uint64)[None, :]
    keep = keep_groups(g, site, seed, offset, p).reshape(rows, cols)
    return torch.from_numpy(keep.astype(np.float32) / (1.0 - float(p)))


def attention_mask(n_seq, heads, L, site, seed, offset, p):
    """float32 torch [n_seq, heads, L, L] multiplier of the attention probabilities."""
    item = np.arange(n_seq * heads, dtype=np.uint64)[:, None, None]
    r = np.arange(L, dtype=np.uint64)[None, :, None]
    kg = np.arange((L + 7) // 8, dtype=np.uint64)[None, None, :]
    g = (item * np.uint64(L) + r) * np.uint64(64) + kg
    keep = keep_groups(g, site, seed, offset, p).reshape(n_seq * heads, L, -1)[:, :, :L]
    return torch.from_numpy(keep.astype(np.float32) / (1.0 - float(p))).reshape(n_seq, heads, L, L)


class DropSpec:
    """(seed, offset, p_hidden, p_attn) of one encoder pass; hands out the site masks in the order of the module
    docstring.  ``layer_base`` shifts the layer numbering (Condenser head layers follow the backbone's)."""

    def __init__(self, seed, offset, p_hidden, p_attn):
        self.seed, self.offset, self.p_hidden, self.p_attn = int(seed), int(offset), float(p_hidden), float(p_attn)

    def hidden(self, x2d_rows, cols, site, row_mul=1):
        if self.p_hidden <= 0:
            return None
        return hidden_mask(x2d_rows, cols, site, self.seed, self.offset, self.p_hidden, row_mul)

    def attn(self, n_seq, heads, L, site):
        if self.p_attn <= 0:
            return None
        return attention_mask(n_seq, heads, L, site, self.seed, self.offset, self.p_attn)


class TorchDropSpec(DropSpec):
    """Same interface, masks drawn from torch's CPU generator (Bernoulli(1 - p) / (1 - p)): the work the reference's own
    nn.Dropout modules do.  Used where the oracle is TIMED as the CPU baseline (bench.py) -- regenerating Philox masks
    in numpy would charge the CPU arm for work the reference does not do.  Not reproducible against the CUDA masks."""

    def __init__(self, p_hidden, p_attn, generator=None):
        super().__init__(0, 0, p_hidden, p_attn)
        self.generator = generator

    def hidden(self, x2d_rows, cols, site, row_mul=1):
        if self.p_hidden <= 0:
            return None
        keep = torch.empty(x2d_rows, cols).bernoulli_(1.0 - self.p_hidden, generator=self.generator)
        return keep / (1.0 - self.p_hidden)

    def attn(self, n_seq, heads, L, site):
        if self.p_attn <= 0:
            return None
        keep = torch.empty(n_seq, heads, L, L).bernoulli_(1.0 - self.p_attn, generator=self.generator)
        return keep / (1.0 - self.p_attn)
