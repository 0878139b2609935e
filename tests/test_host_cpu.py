"""CPU: host-side mirror of the reference surface -- class / attribute / state-dict contract, registry, loud
failure without CUDA, and the pure-host parts of the DRO losses against the oracle."""
import types

import numpy as np
import pytest
import torch

TINY = dict(hidden=128, layers=12, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)


def hf_config(cfg, **kw):
    from transformers import BertConfig
    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)


def test_ance_surface_and_state_dict_contract():
    from transformers import BertForSequenceClassification

    from cocodr_b200 import models
    from cocodr_b200.bert import BertModel
    from oracle import bert_ref
    cfgd = models.MSMarcoConfigDict["rdot_nll_condenser"]
    m = cfgd.model_class(hf_config(TINY, num_labels=2))
    assert isinstance(m.bert, BertModel) and isinstance(m, BertForSequenceClassification)
    ref_keys = set(BertForSequenceClassification(hf_config(TINY, num_labels=2)).state_dict())
    ref_keys |= {"embeddingHead.weight", "embeddingHead.bias", "norm.weight", "norm.bias"}  # models.py:202-203
    assert set(m.state_dict()) == ref_keys
    for k, shp in bert_ref.param_shapes(TINY).items():
        assert tuple(m.state_dict()["bert." + k].shape) == shp
    for attr in ("total", "correct", "dro_type", "use_mean", "query_emb", "body_emb", "forward_model", "add_group_loss",
                 "output_state", "gather_tensors"):
        assert hasattr(m, attr), attr
    assert m.dro_type == "erm" and m.total == 0
    m.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=0), 5, "idro", 0.25, 0.01, 0.1, 0.05, True)
    assert m.dro_type == "idro" and m.n_groups == 5 and len(m.accum_group_loss) == 5
    assert {"loss.h_fun", "loss.sum_losses", "loss.count_cat"} <= set(m.state_dict())  # DRO state rides along
    names = [n for n, _ in m.bert.named_parameters()]
    from oracle import heads_ref
    picked = m.loss._params(m.bert)
    assert len(picked) == len(heads_ref.idro_param_names(names, "base")) == 48


def test_no_cpu_fallback_anywhere():
    from cocodr_b200 import models, ops, scan
    m = models.BertDot_NLL_LN(hf_config(TINY, num_labels=2))
    ids, mask = torch.ones(2, 8, dtype=torch.long), torch.ones(2, 8, dtype=torch.long)
    with pytest.raises(RuntimeError):
        m(ids, mask, ids, mask, ids, mask, weights=torch.ones(2))
    with pytest.raises(RuntimeError):
        ops.pair_nll(torch.randn(2, 8), torch.randn(2, 8), torch.randn(2, 8))
    with pytest.raises(RuntimeError):
        ops.coco_contrastive(torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        scan.search(torch.randn(2, 8).half(), torch.randn(9, 8).half(), 3)
    with pytest.raises(RuntimeError):
        scan.IndexFlatIP(8).search(np.zeros((1, 8), np.float32), 1)
    with pytest.raises(RuntimeError):
        scan.search_async(torch.randn(2, 8).half(), torch.randn(9, 8).half(), 3)
    # the widened rows (optimizers, ANN episode) are CUDA-only as well
    from cocodr_b200 import mining, optim
    with pytest.raises(RuntimeError):
        mining.mine_negatives(torch.zeros(2, 4, dtype=torch.long), torch.arange(9), torch.zeros(2, dtype=torch.long), 2)
    with pytest.raises(RuntimeError):
        mining.kmeans(torch.randn(16, 8).half(), 2)
    w = torch.nn.Parameter(torch.randn(8, 8))
    w.grad = torch.randn(8, 8)
    for opt in (optim.AdamW([w], lr=1e-3), optim.Lamb([w], lr=1e-3)):
        with pytest.raises(RuntimeError):
            opt.step()


def test_optimizer_surface_matches_the_reference_constructors():
    """Lamb(params, lr, betas, eps, weight_decay, adam) (ANCE/utils/lamb.py:44-58) and AdamW(params, lr, eps)
    (ANCE/drivers/run_ann.py:139-144): same keywords, same param_groups defaults, torch state_dict layout."""
    from cocodr_b200 import optim
    w = torch.nn.Parameter(torch.zeros(4))
    lamb = optim.Lamb([{"params": [w]}], lr=5e-6, eps=1e-8, weight_decay=0.01)
    g = lamb.param_groups[0]
    assert (g["lr"], g["eps"], g["weight_decay"], g["betas"]) == (5e-6, 1e-8, 0.01, (0.9, 0.999))
    adamw = optim.AdamW([w], lr=1e-5, eps=1e-8)
    assert adamw.param_groups[0]["betas"] == (0.9, 0.999) and adamw.mode == optim.MODE_HF
    assert set(adamw.state_dict()) == {"state", "param_groups"}
    with pytest.raises(NotImplementedError):
        optim.Lamb([w], adam=True)
    with pytest.raises(ValueError):
        optim.AdamW([w], semantics="nope")


def test_unsupported_configs_are_rejected():
    from cocodr_b200.bert import BertModel
    with pytest.raises(RuntimeError):
        BertModel(hf_config(dict(TINY, heads=4)))  # head_dim 32
    from transformers import BertConfig
    with pytest.raises(RuntimeError):
        BertModel(BertConfig(hidden_size=128, num_attention_heads=2, num_hidden_layers=1, intermediate_size=256,
                             hidden_act="relu"))


def test_dro_greedy_update_mw_matches_oracle():
    """update_mw (dro_loss.py:88-120) is host/device-agnostic torch code: drive it on CPU against the oracle."""
    from cocodr_b200.dro_loss import DROGreedyLoss
    from oracle import heads_ref
    G, alpha, eps, ema = 7, 0.25, 0.01, 0.1
    torch.manual_seed(3)
    for weight_ema in (True, False):
        mod = DROGreedyLoss(types.SimpleNamespace(local_rank=0), G, alpha, eps, ema, weight_ema)
        h, sl, cc = torch.ones(G), torch.zeros(G), torch.ones(G)
        for step in range(4):
            losses, g = torch.rand(16) * 3, torch.randint(0, G, (16,))
            _, _, _, h, sl, cc = heads_ref.dro_greedy_forward(losses, g, h, sl, cc, G, alpha, eps, ema, weight_ema)
            # same EMA inputs, then the module's own greedy re-weighting
            mod.sum_losses, mod.count_cat = sl.clone(), cc.clone()
            mod.update_mw()
            np.testing.assert_allclose(mod.h_fun.numpy(), h.numpy(), rtol=1e-6, atol=1e-7)


def test_bench_reference_arm_contract():
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""  # ranks > 0 exit without work
    sys.path.insert(0, root)
    import bench
    assert abs(bench.fwd_flops_per_seq() / 1e9 - 22.347) < 0.01  # SURVEY.md §8d / BASELINE.md §3
    p = bench.measured_peaks()
    assert p["tensor"] > 1000 and p["hbm"] > 5000


def test_plan_buckets_covers_all_sequences_within_the_token_budget():
    """mining.plan_buckets (length-bucketed inference, SURVEY 8d mode B): consecutive groups over the sorted lengths,
    padded length = longest of the group rounded up to 16, tokens per group within the budget."""
    import random
    from cocodr_b200.mining import plan_buckets
    rng = random.Random(0)
    for n, lo, hi, budget in ((0, 1, 1, 1024), (1, 5, 5, 1024), (2048, 16, 128, 16384), (2048, 4, 16, 16384), (300, 1, 512, 4096)):
        lens = sorted(rng.randint(lo, hi) for _ in range(n))
        plan = plan_buckets(lens, token_budget=budget, max_len=hi)
        assert [a for a, _, _ in plan] == [0] * (n > 0) + [b for _, b, _ in plan][:-1]  # consecutive, starts at 0
        assert (plan[-1][1] if plan else 0) == n
        for a, b, L in plan:
            assert b > a and L % 16 == 0 or L == hi
            assert L >= lens[b - 1] and (L - lens[b - 1] < 16 or L == hi)
            assert (b - a) * L <= budget or b - a == 1


def test_forward_counts_do_not_outlive_their_parameters():
    """GradSync decides per layer whether its flat gradient buffer may be all-reduced in place from the number of
    forwards since the last exit (ops.FWD_CALLS).  The counts are keyed WEAKLY by the parameter object: keyed by id(), a
    deleted model's stale entry was inherited by a new parameter that happened to reuse the id -- on one rank of eight
    only, which changed the order of that rank's collectives (round-2 N = 8 hang)."""
    import gc
    import torch
    from cocodr_b200 import ops

    class Ctx:
        needs_input_grad = (True,)

    ops.FWD_CALLS.clear()
    p = torch.nn.Parameter(torch.zeros(4))
    ops._note_forward(Ctx(), (p,))
    ops._note_forward(Ctx(), (p,))
    assert ops.FWD_CALLS.get(p, 1) == 2 and len(ops.FWD_CALLS) == 1
    del p
    gc.collect()
    assert len(ops.FWD_CALLS) == 0  # the entry died with the parameter: nothing for a recycled id to inherit
    q = torch.nn.Parameter(torch.zeros(4))
    assert ops.FWD_CALLS.get(q, 1) == 1
    Ctx.needs_input_grad = (False,)
    ops._note_forward(Ctx(), (q,))  # forwards autograd does not record (no_grad) are not counted
    assert len(ops.FWD_CALLS) == 0
