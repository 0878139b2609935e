"""CPU: pin the oracle restatements against outputs of the UNMODIFIED reference classes
(fixtures written by oracle/make_golden.py in the dev container)."""
import os

import numpy as np
import pytest
import torch

from oracle import bert_ref, heads_ref, scan_ref

TINY = dict(hidden=128, layers=12, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def triplet(cfg, B, L, seed, full=False):
    out = []
    for k in range(3):
        out += list(bert_ref.synth_batch(B, L, cfg["vocab"], seed + k, full))
    return out


def test_encoder_and_pair_nll_tiny(golden_dir):
    g = load(golden_dir, "ance_tiny.npz")
    st = bert_ref.synth_state(TINY, 0)
    q, mq, a, ma, b, mb = triplet(TINY, int(g["B"]), int(g["L"]), int(g["seed"]))
    qe = bert_ref.cls_embedding(st, q, mq, TINY)
    ae = bert_ref.cls_embedding(st, a, ma, TINY)
    be = bert_ref.cls_embedding(st, b, mb, TINY)
    np.testing.assert_allclose(qe.numpy(), g["q_emb"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ae.numpy(), g["a_emb"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(be.numpy(), g["b_emb"], rtol=1e-4, atol=1e-5)
    loss, accs, logits = heads_ref.pair_nll(qe, ae, be)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-4, atol=1e-5)
    assert (accs.numpy() == g["accs"]).all()
    erm = heads_ref.erm_reduce(loss, torch.from_numpy(g["weights"]))
    assert abs(erm.item() - float(g["erm_loss"])) < 1e-5
    np.testing.assert_allclose(heads_ref.qp_infonce(qe, ae).numpy(), g["qp_infonce"], rtol=1e-4, atol=1e-5)


def test_encoder_grads_tiny(golden_dir):
    g = load(golden_dir, "ance_tiny.npz")
    st = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(TINY, 0).items()}
    q, mq, a, ma, b, mb = triplet(TINY, int(g["B"]), int(g["L"]), int(g["seed"]))
    loss, _, _ = heads_ref.pair_nll(bert_ref.cls_embedding(st, q, mq, TINY), bert_ref.cls_embedding(st, a, ma, TINY),
                                    bert_ref.cls_embedding(st, b, mb, TINY))
    heads_ref.erm_reduce(loss, torch.from_numpy(g["weights"])).backward()
    for key in g.files:
        if not key.startswith("grad."):
            continue
        name = key[5:]
        if name.endswith(".rownorm"):
            got = st[name[:-8]].grad.norm(dim=1).numpy()
        elif name.endswith(".norm"):
            got = st[name[:-5]].grad.norm().item()
        else:
            got = st[name].grad.numpy()
        atol = 1e-3 * float(np.abs(g[key]).max()) + 1e-7  # sums of cancelling terms: scale-relative
        np.testing.assert_allclose(got, g[key], rtol=2e-3, atol=atol, err_msg=key)


def test_cfg1_base_embeddings(golden_dir):
    """cfg1: BERT-base, B=8, L=128 (BASELINE.json configs[0]) -- reference outputs vs oracle."""
    g = load(golden_dir, "ance_cfg1_base.npz")
    cfg = bert_ref.make_config()
    st = bert_ref.synth_state(cfg, 0)
    q, mq, a, ma, b, mb = triplet(cfg, int(g["B"]), int(g["L"]), int(g["seed"]), full=True)
    with torch.no_grad():
        qe = bert_ref.cls_embedding(st, q, mq, cfg)
        ae = bert_ref.cls_embedding(st, a, ma, cfg)
    np.testing.assert_allclose(qe.numpy(), g["q_emb"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(ae.numpy(), g["a_emb"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(heads_ref.qp_infonce(qe, ae).numpy(), g["qp_infonce"], rtol=1e-3, atol=1e-4)


def _run_dro(g, kind):
    st = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(TINY, 0).items()}
    B, L, seed, G = int(g["B"]), int(g["L"]), int(g["seed"]), int(g["n_groups"])
    alpha, eps, ema, rho = float(g["alpha"]), float(g["eps"]), float(g["ema"]), float(g["rho"])
    names = heads_ref.idro_param_names(list(st.keys()), "base")
    params = [st[n] for n in st if n in set(names)]
    h = torch.ones(G)
    sum_losses, count_cat = torch.zeros(G), torch.ones(G)
    for s in range(int(g["steps"])):
        q, mq, a, ma, b, mb = triplet(TINY, B, L, seed + 10 * s)
        gid = torch.from_numpy(g[f"group_ids_{s}"])
        for v in st.values():
            v.grad = None
        loss, _, _ = heads_ref.pair_nll(bert_ref.cls_embedding(st, q, mq, TINY), bert_ref.cls_embedding(st, a, ma, TINY),
                                        bert_ref.cls_embedding(st, b, mb, TINY))
        if kind == "idro":
            robust, gl, gc, h = heads_ref.idro_forward(loss, gid, params, h, G, alpha, ema, rho, eps)
        else:
            robust, gl, gc, h, sum_losses, count_cat = heads_ref.dro_greedy_forward(
                loss, gid, h, sum_losses, count_cat, G, alpha, eps, ema, True, torch.ones(B))
            np.testing.assert_allclose(sum_losses.numpy(), g[f"sum_losses_{s}"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(count_cat.numpy(), g[f"count_cat_{s}"], rtol=1e-5)
        robust.backward()
        assert abs(robust.item() - float(g[f"robust_{s}"])) < 1e-4 * max(1, abs(robust.item()))
        np.testing.assert_allclose(gl.numpy(), g[f"group_losses_{s}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(gc.numpy(), g[f"group_counts_{s}"])
        np.testing.assert_allclose(h.numpy(), g[f"h_fun_{s}"], rtol=2e-3, atol=1e-5)
        np.testing.assert_allclose(st["encoder.layer.11.attention.self.query.weight"].grad.numpy(), g[f"grad_q11_{s}"],
                                   rtol=5e-3, atol=1e-6)


def test_idro_trajectory(golden_dir):
    _run_dro(load(golden_dir, "idro_tiny.npz"), "idro")


def test_dro_greedy_trajectory(golden_dir):
    _run_dro(load(golden_dir, "dro_greedy_tiny.npz"), "greedy")


def contrastive_inputs(n, h, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(h, generator=g) * 14.7 / h ** 0.5
    trained = base[None, :] + 0.015 * torch.randn(n, h, generator=g)
    gauss = torch.randn(n, h, generator=g) * 0.5
    return {"trained": trained, "gauss": gauss}


@pytest.mark.parametrize("n,h,nm", [(16, 128, "small"), (512, 1024, "cfg4")])
def test_coco_contrastive(golden_dir, n, h, nm):
    g = load(golden_dir, "contrastive.npz")
    for tag, e in contrastive_inputs(n, h, 11).items():
        e = e.clone().requires_grad_(True)
        loss = heads_ref.coco_contrastive(e)
        loss.mean().backward()
        np.testing.assert_allclose(loss.detach().numpy(), g[f"{nm}_{tag}_loss"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(e.grad[:8].numpy(), g[f"{nm}_{tag}_grad_head"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(e.grad.norm(dim=1).numpy(), g[f"{nm}_{tag}_grad_rownorm"], rtol=1e-4, atol=1e-6)


def test_coco_cls_and_contrastive_tiny(golden_dir):
    g = load(golden_dir, "coco_tiny.npz")
    st = bert_ref.synth_state(TINY, 0)
    ids, mask = bert_ref.synth_batch(2 * int(g["n_docs"]), int(g["L"]), TINY["vocab"], int(g["seed"]))
    with torch.no_grad():
        cls = bert_ref.cls_embedding(st, ids, mask, TINY)
    np.testing.assert_allclose(cls.numpy(), g["cls"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(heads_ref.coco_contrastive(cls).numpy(), g["co_loss"], rtol=1e-4, atol=1e-5)


def test_scan_small(golden_dir):
    g = load(golden_dir, "scan_small.npz")
    for kind in ("exact", "gauss"):
        Q, P = scan_ref.synth_corpus(6000, 37, 768, seed=7, kind=kind)
        D, I = scan_ref.search(Q, P, 100, chunk=1500)
        assert (I == g[f"{kind}_I"]).all()
        np.testing.assert_array_equal(D, g[f"{kind}_D"])
        # brute-force definition on a few rows
        s = (Q[:3].float() @ P.float().t()).numpy()
        for r in range(3):
            order = sorted(range(s.shape[1]), key=lambda j: (-s[r, j], j))[:100]
            assert order == list(I[r])


def test_lamb_oracle_matches_reference_fixture(golden_dir):
    """oracle/optim_ref.lamb_step vs the parameters and trust ratios the UNMODIFIED reference Lamb produced
    (tests/golden/lamb_tiny.npz, written by oracle/make_golden.py gen_lamb)."""
    import os
    import numpy as np
    import torch
    from oracle import optim_ref
    from oracle.make_golden import lamb_inputs
    g = np.load(os.path.join(golden_dir, "lamb_tiny.npz"))
    for tag, wd in (("wd0", 0.0), ("wd01", 0.01)):
        params, grads = lamb_inputs(int(g["seed"]))
        ms, vs = [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params]
        trust = [1.0] * len(params)
        for step in range(3):
            for i, (p, m, v) in enumerate(zip(params, ms, vs)):
                trust[i] = optim_ref.lamb_step(p, grads[step][i].clone(), m, v, lr=1e-3, eps=1e-6, weight_decay=wd)
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.numpy(), g[f"{tag}.p{i}"], rtol=1e-6, atol=1e-8)
            assert abs(trust[i] - float(g[f"{tag}.trust{i}"])) <= 1e-5 * max(1.0, abs(trust[i]))


def test_mining_oracle_matches_reference_fixture(golden_dir):
    """oracle/mining_ref.generate_negatives vs the UNMODIFIED GenerateNegativePassaageID (its AST node executed by
    oracle/make_golden.py gen_mining): both the SelectTopK and the shuffled branch, per-query negatives and
    reciprocal ranks."""
    import os
    import numpy as np
    from oracle import mining_ref
    g = np.load(os.path.join(golden_dir, "mining_tiny.npz"))
    I, doc_pid, pos, n_neg = g["I"], g["doc_pid"], g["pos"], int(g["n_neg"])
    for tag in ("topk", "shuf"):
        for q in range(I.shape[0]):
            order = None if tag == "topk" else list(g["shuf.order"][q])
            negs, rr = mining_ref.generate_negatives(list(I[q]), doc_pid, int(pos[q]), n_neg, order=order)
            want = [int(v) for v in g[f"{tag}.neg"][q] if v >= 0]
            assert negs == want, (tag, q)
            assert abs(rr - float(g[f"{tag}.rr"][q])) < 1e-12
