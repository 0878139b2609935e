"""CPU: the oracle's counter-based dropout masks (oracle/dropout_ref.py) -- the generator is pinned to the published
Random123 known-answer vectors for Philox4x32-10, the mask layout to the rule in include/cocodr_b200.h."""
import numpy as np
import torch

from oracle import bert_ref, dropout_ref


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors (philox4x32 10)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = dropout_ref.philox4x32_10(*[np.uint64(c) for c in ctr], *key)
        assert tuple(int(x) for x in got) == want


def test_mask_layout_and_rate():
    p, seed, off = 0.1, 1234, 7
    m = dropout_ref.hidden_mask(64, 768, 5, seed, off, p)
    assert m.shape == (64, 768)
    vals = set(np.unique(m.numpy()).tolist())
    assert vals == {0.0, np.float32(1.0 / 0.9).item()}
    keep = (m > 0).float().mean().item()
    assert abs(keep - (1 - dropout_ref.threshold(p) / 65536.0)) < 0.01
    # group g of row m is Philox counter (m * row_mul) * (cols / 8) + g: the [CLS]-only view (row_mul = L) picks the
    # full tensor's rows 0, L, 2L, ...
    full = dropout_ref.hidden_mask(4 * 16, 128, 11, seed, off, p)
    cls = dropout_ref.hidden_mask(4, 128, 11, seed, off, p, row_mul=16)
    assert torch.equal(cls, full[::16])
    # different site / offset / seed -> different masks; same arguments -> same mask
    assert torch.equal(m, dropout_ref.hidden_mask(64, 768, 5, seed, off, p))
    for other in (dropout_ref.hidden_mask(64, 768, 6, seed, off, p), dropout_ref.hidden_mask(64, 768, 5, seed + 1, off, p),
                  dropout_ref.hidden_mask(64, 768, 5, seed, off + 1, p)):
        assert not torch.equal(m, other)
    a = dropout_ref.attention_mask(3, 2, 40, 9, seed, off, p)
    assert a.shape == (3, 2, 40, 40) and abs((a > 0).float().mean().item() - 0.9) < 0.02
    # element (item, r, c) is half-word c % 8 of the call with counter (item * L + r) * 64 + c / 8
    k = dropout_ref.keep_groups(np.array([(4 * 40 + 17) * 64 + 3], dtype=np.uint64), 9, seed, off, p)[0]
    assert np.array_equal(k, (a.view(6, 40, 40)[4, 17, 24:32] > 0).numpy())


def test_oracle_encoder_with_dropout_is_unbiased_and_eval_is_identity():
    cfg = dict(hidden=64, layers=2, heads=1, inter=128, vocab=500, max_pos=32, type_vocab=2)
    st = bert_ref.synth_state(cfg, 0)
    ids, mask = bert_ref.synth_batch(3, 16, cfg["vocab"], 3)
    base = bert_ref.cls_embedding(st, ids, mask, cfg)
    assert torch.equal(base, bert_ref.cls_embedding(st, ids, mask, cfg, drop=None))
    zero = dropout_ref.DropSpec(1, 1, 0.0, 0.0)
    assert torch.equal(base, bert_ref.cls_embedding(st, ids, mask, cfg, drop=zero))
    d1 = bert_ref.cls_embedding(st, ids, mask, cfg, drop=dropout_ref.DropSpec(1, 1, 0.1, 0.1))
    d2 = bert_ref.cls_embedding(st, ids, mask, cfg, drop=dropout_ref.DropSpec(1, 2, 0.1, 0.1))
    assert not torch.equal(d1, base) and not torch.equal(d1, d2)
    assert torch.equal(d1, bert_ref.cls_embedding(st, ids, mask, cfg, drop=dropout_ref.DropSpec(1, 1, 0.1, 0.1)))
