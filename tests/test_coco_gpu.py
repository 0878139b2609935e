"""GPU: drop-in CoCondenserForPretraining (backbone + Condenser head + two MLM losses + sequence-contrastive
loss, all on the CUDA kernels) vs the fixture written by the UNMODIFIED reference class
(tests/golden/coco_tiny.npz, oracle/make_golden.py::gen_coco), plus a BERT-large-shaped L = 256 smoke."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(hidden=128, layers=12, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)


def hf_config(cfg, **kw):
    from transformers import BertConfig
    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)


def build_from_golden(g):
    from transformers import BertForMaskedLM

    from cocodr_b200 import modeling
    from oracle import bert_ref
    lm = BertForMaskedLM(hf_config(TINY))
    margs = types.SimpleNamespace(n_head_layers=int(g["n_head_layers"]), skip_from=int(g["skip_from"]), late_mlm=True)
    dargs = types.SimpleNamespace(train_method="coco")
    targs = types.SimpleNamespace(per_device_train_batch_size=int(g["n_docs"]), local_rank=-1)
    m = modeling.CONDENSER_TYPE_MAP['bert'](lm, margs, dargs, targs)
    state = {k[len("state."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state.")}
    res = m.load_state_dict(state, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    m.lm.bert.load_state_dict(bert_ref.synth_state(TINY, 0), strict=False)
    p = m.lm.cls.predictions
    assert p.decoder.weight is m.lm.bert.embeddings.word_embeddings.weight  # tied
    return m.cuda()


def rel(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max() / (np.abs(ref).max() + 1e-12)


def test_cocondenser_forward_backward_matches_reference(golden_dir):
    from oracle import bert_ref
    g = np.load(os.path.join(golden_dir, "coco_tiny.npz"))
    m = build_from_golden(g)
    m.train()
    n_docs, L = int(g["n_docs"]), int(g["L"])
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(2 * n_docs, L, TINY["vocab"], int(g["seed"])))
    labels = torch.from_numpy(g["labels"]).cuda()
    inp = {"input_ids": ids, "attention_mask": mask}
    total = m(inp, labels)
    assert abs(total.item() - float(g["total"])) < 1e-2 * abs(float(g["total"])), (total.item(), float(g["total"]))
    with torch.no_grad():
        cls, last, hidden = m._encode(inp)
        co = m.compute_contrastive_loss(cls)
        idx, row_labels, count = m._mlm_rows(labels)
        lm_mlm = m._mlm(last, idx, row_labels, count)
    assert rel(cls.cpu().numpy(), g["cls"]) < 1e-2
    np.testing.assert_allclose(co.cpu().numpy(), g["co_loss"], rtol=1e-2, atol=5e-3)
    assert abs(lm_mlm.item() - float(g["lm_mlm_loss"])) < 1e-2 * float(g["lm_mlm_loss"])
    m.zero_grad()
    total.backward()
    named = dict(m.named_parameters())
    got = named["lm.bert.encoder.layer.0.attention.self.query.weight"].grad.cpu().numpy()
    ref = g["grad_l0_query"]
    cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref)))
    assert rel(got, ref) < 0.1 and cos > 0.995, (rel(got, ref), cos)
    rn = named["lm.bert.embeddings.word_embeddings.weight"].grad.norm(dim=1).cpu().numpy()  # lookup + tied decoder
    assert rel(rn, g["grad_word_rownorm"]) < 5e-2
    for k in ("c_head.1.output.dense.weight", "lm.cls.predictions.transform.dense.weight", "lm.cls.predictions.bias"):
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all() and named[k].grad.abs().sum() > 0, k


def test_cocondenser_persistence_and_reference_signatures(golden_dir, tmp_path):
    from oracle import bert_ref
    g = np.load(os.path.join(golden_dir, "coco_tiny.npz"))
    m = build_from_golden(g).eval()
    m.save_pretrained(str(tmp_path))
    assert os.path.exists(tmp_path / "model.pt") and os.path.exists(tmp_path / "args.pt")
    from cocodr_b200 import modeling
    m2 = modeling.CoCondenserForPretraining.from_pretrained(m.model_args, m.data_args, m.train_args, str(tmp_path)).cuda()
    assert set(m2.state_dict()) == set(m.state_dict())
    n_docs, L = int(g["n_docs"]), int(g["L"])
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(2 * n_docs, L, TINY["vocab"], int(g["seed"])))
    labels = torch.from_numpy(g["labels"]).cuda()
    with torch.no_grad():
        a = m({"input_ids": ids, "attention_mask": mask}, labels)
        b = m2({"input_ids": ids, "attention_mask": mask}, labels)
        # reference-style call of mlm_loss on caller-visible fp32 hidden states
        out = m.lm.bert(input_ids=ids, attention_mask=mask)
        ml = m.mlm_loss(out[0], labels)
    assert abs(a.item() - b.item()) < 1e-5
    assert abs(ml.item() - float(g["lm_mlm_loss"])) < 2e-2 * float(g["lm_mlm_loss"])


def test_cocondenser_large_shape_l256_runs():
    """cfg4 shape class (hidden 1024, 16 heads, L = 256, tiled attention) at reduced depth: finite loss and
    gradients through every component; contrastive part checked against the oracle on the produced CLS."""
    from transformers import BertConfig, BertForMaskedLM

    from cocodr_b200 import modeling
    from oracle import heads_ref
    torch.manual_seed(0)
    cfg = BertConfig(hidden_size=1024, num_hidden_layers=3, num_attention_heads=16, intermediate_size=4096,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    lm = BertForMaskedLM(cfg)
    m = modeling.CoCondenserForPretraining(lm, types.SimpleNamespace(n_head_layers=2, skip_from=1, late_mlm=True),
                                           types.SimpleNamespace(train_method="coco"),
                                           types.SimpleNamespace(per_device_train_batch_size=4, local_rank=-1)).cuda()
    m.train()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1000, cfg.vocab_size, (8, 256), generator=g)
    mask = torch.ones(8, 256, dtype=torch.long)
    mask[1, 200:] = 0
    ids = ids * mask
    labels = torch.where((torch.rand(ids.shape, generator=g) < 0.15) & (mask > 0), ids, torch.full_like(ids, -100))
    ids, mask, labels = ids.cuda(), mask.cuda(), labels.cuda()
    total = m({"input_ids": ids, "attention_mask": mask}, labels)
    total.backward()
    assert torch.isfinite(total)
    for n, p in m.named_parameters():
        if "pooler" in n or "position_embeddings" in n or "token_type" in n:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    with torch.no_grad():
        cls = m._encode({"input_ids": ids, "attention_mask": mask})[0]
        co = m.compute_contrastive_loss(cls)
    np.testing.assert_allclose(co.cpu().numpy(), heads_ref.coco_contrastive(cls.cpu()).numpy(), rtol=1e-4, atol=1e-4)


def test_fixed_capacity_mlm_gather_equals_dynamic_and_is_graph_capturable(golden_dir):
    """mlm_capacity: masked rows gathered into a fixed-size buffer (no host sync) -> same loss and gradients as the
    exact dynamic gather; the whole step replays from a CUDA graph with fresh dropout masks per replay."""
    from oracle import bert_ref
    from cocodr_b200.graph import GraphedTrainStep
    g = np.load(os.path.join(golden_dir, "coco_tiny.npz"))
    m = build_from_golden(g)
    m.eval()  # no dropout: the two gathers must agree exactly up to summation order
    n_docs, L = int(g["n_docs"]), int(g["L"])
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(2 * n_docs, L, TINY["vocab"], int(g["seed"])))
    labels = torch.from_numpy(g["labels"]).cuda()
    inp = {"input_ids": ids, "attention_mask": mask}
    out = {}
    for cap in (None, 0.5):
        m.mlm_capacity = cap
        m.zero_grad(set_to_none=True)
        loss = m(inp, labels)
        loss.backward()
        out[cap] = (loss.item(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    assert abs(out[None][0] - out[0.5][0]) < 1e-5 * abs(out[None][0])
    for n, gr in out[None][1].items():
        err = (out[0.5][1][n] - gr).abs().max().item()
        assert err <= 1e-6 + 2e-3 * gr.abs().max().item(), (n, err)
    assert int(m.mlm_overflow) == 0
    m.mlm_capacity = 0.01  # too small on purpose (every position labelled): the overflow counter must say so
    m(inp, torch.where(mask.bool(), ids, torch.full_like(ids, -100)))
    assert int(m.mlm_overflow) > 0
    del loss, out, m  # (autograd nodes of the eager default-stream passes must not outlive into the capture)
    m = build_from_golden(g)
    m.mlm_capacity = 0.5
    m.lm.config.hidden_dropout_prob = m.lm.config.attention_probs_dropout_prob = 0.1  # (the fixture model has p = 0)
    m.train()
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.0)
    step = GraphedTrainStep(lambda i, a, l: m({"input_ids": i, "attention_mask": a}, l), opt, (ids, mask, labels))
    losses = [step(ids, mask, labels).item() for _ in range(3)]
    assert all(np.isfinite(losses)) and len(set(losses)) == 3  # c_head dropout draws new masks at every replay
