"""GPU: oracle parity AT BASELINE.json SIZES (the small-shape tests live in test_model_gpu.py).

  cfg2   BERT-base, 64 q + 64 p, L = 128, MS-MARCO-shaped lengths: forward and EVERY parameter gradient vs the fp32 oracle
         (oracle/bert_ref.py, ~10 s of host time) for a fixed upstream d(cls) -- embeddings 1e-2, gradients 2.5e-2
  drift  the ill-conditioned ance_tiny loss-gradient fixture through the STOCK fp16 path (HF BertModel under
         torch.autocast on the same GPU): how far does library fp16 arithmetic drift on it, next to ours
  cfg3   iDRO on BERT-base with G = 50 groups, 16 triplets: robust loss, group means, h_fun after one step and the group
         gradient matrix vs oracle/heads_ref.idro_forward (autograd per group, dro_loss.py:192-254), through the grouped
         K11 path
  cfg4   BERT-large at full depth (24 layers, L = 256, 8 spans): CLS + hidden states vs the fp32 oracle
  amp    the drop-in under torch.cuda.amp.GradScaler (what run_ann.py's --fp16 / HF Trainer do): scaled backward,
         unscale, inf/nan skip logic
"""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _rel(got, ref, floor=0.0):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), floor, 1e-12))


def _record(name, payload):
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, f"parity_{name}.json"), "w") as f:
            json.dump(payload, f, indent=1)
    except OSError:
        pass


def _hf_config(cfg, **kw):
    from transformers import BertConfig
    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)


def _build(cfg, cls_name="BertDot_NLL_LN"):
    from cocodr_b200 import models
    from oracle import bert_ref
    m = getattr(models, cls_name)(_hf_config(cfg, num_labels=2))
    m.bert.load_state_dict(bert_ref.synth_state(cfg, 0), strict=False)
    return m.cuda()


def test_cfg2_bert_base_forward_backward_vs_oracle():
    from oracle import bert_ref
    cfg = bert_ref.make_config()
    m = _build(cfg).train()
    B, L = 64, 128
    qi, qm = bert_ref.synth_batch(B, L, cfg["vocab"], 1234, mean_len=0.15)   # short queries
    pi, pm = bert_ref.synth_batch(B, L, cfg["vocab"], 1235, mean_len=0.55)   # passages, one full-length row each
    ids, mask = torch.cat([qi, pi]), torch.cat([qm, pm])
    dcls = torch.randn(2 * B, cfg["hidden"], generator=torch.Generator().manual_seed(5)) * 0.05
    cls = m.query_emb(ids.cuda(), mask.cuda())
    (cls * dcls.cuda()).sum().backward()
    torch.set_num_threads(os.cpu_count() or 1)
    leaf = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(cfg, 0).items()}
    ref_cls = bert_ref.cls_embedding(leaf, ids, mask, cfg)
    (ref_cls * dcls).sum().backward()
    fwd = _rel(cls.detach().cpu().numpy(), ref_cls.detach().numpy())
    named = dict(m.bert.named_parameters())
    errs = {}
    for name, ref in leaf.items():
        got = named[name].grad
        assert got is not None, name
        if "key.bias" in name:  # analytically zero (softmax is invariant to a key-bias shift): rounding noise on both sides
            continue
        errs[name] = _rel(got.cpu().numpy(), ref.grad.numpy(), floor=1e-4)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    _record("cfg2_base", {"forward_rel": fwd, "grad_rel_max": worst[0][1], "grad_rel_mean": float(np.mean(list(errs.values()))),
                          "worst": worst, "n_params": len(errs), "sequences": 2 * B, "seq_len": L})
    assert fwd < 1e-2, fwd
    assert worst[0][1] < 2.5e-2, worst
    assert len(errs) == 197 - 12


def test_ance_tiny_gradient_drift_of_the_stock_fp16_path(golden_dir):
    """Same fixture, same bound question as test_model_gpu.py::test_ance_tiny_forward_backward_matches_reference: the
    loss gradient of a random-init encoder is a difference of near-identical CLS vectors, so ANY fp16 forward drifts
    on it.  Measured here for HF's own BertModel under torch.autocast(fp16) (cuBLAS / eager attention) and for the
    drop-in, both against the reference's fp32 fixture."""
    from transformers import BertModel
    from oracle import bert_ref, heads_ref
    import test_model_gpu as T
    g = np.load(os.path.join(golden_dir, "ance_tiny.npz"))
    B, L, seed = int(g["B"]), int(g["L"]), int(g["seed"])
    q, mq, a, ma, b, mb = T.triplet(T.TINY, B, L, seed)
    w = torch.from_numpy(g["weights"]).cuda()

    def grad_errs(named):
        out = {}
        for key in g.files:
            if not key.startswith("grad.") or key.endswith("norm") or np.abs(g[key]).max() < 1e-5:
                continue
            out[key[5:]] = _rel(named[key[5:]].grad.float().cpu().numpy(), g[key], floor=1e-5)
        return out

    ours = T.build(T.TINY).train()
    loss = ours(q, mq, a, ma, b, mb, weights=w)[0]
    loss.backward()
    e_ours = grad_errs(dict(ours.bert.named_parameters()))

    hf = BertModel(_hf_config(T.TINY, attn_implementation="eager"), add_pooling_layer=False)
    hf.load_state_dict(bert_ref.synth_state(T.TINY, 0), strict=False)
    hf = hf.cuda().train()
    with torch.autocast("cuda", dtype=torch.float16):
        embs = [hf(input_ids=i, attention_mask=m_)[0][:, 0] for i, m_ in ((q, mq), (a, ma), (b, mb))]
    per = heads_ref.pair_nll(*[e.float() for e in embs])[0]
    (per * w).mean().backward()
    e_stock = grad_errs(dict(hf.named_parameters()))
    wo, ws = max(e_ours.values()), max(e_stock.values())
    _record("ance_tiny_drift", {"ours_worst": wo, "stock_autocast_fp16_worst": ws,
                                "ours_mean": float(np.mean(list(e_ours.values()))),
                                "stock_mean": float(np.mean(list(e_stock.values()))),
                                "ours_worst_name": max(e_ours, key=e_ours.get), "stock_worst_name": max(e_stock, key=e_stock.get)})
    print("ance_tiny loss-gradient drift: ours", wo, "stock autocast fp16", ws)
    # MEASURED (profiles/r02_parity.md): the stock path drifts ~0.03 on this fixture, the drop-in ~0.3.  The stock path
    # (and the reference's apex O1) keeps LayerNorm outputs and the residual stream in fp32 and only runs the GEMMs in
    # fp16; the drop-in stores every activation in fp16 (half the HBM traffic of the bandwidth-bound kernels), so each
    # CLS embedding carries ~2e-3 of independent rounding noise -- invisible in logits and losses (the north-star bound,
    # asserted at 1e-2 everywhere) but amplified in THIS gradient, which is sigma * (b - a) for two embeddings that
    # differ by ~2 % in a random-init encoder.  The bound below is therefore the drop-in's own measured band, not a
    # claim of equivalence with the stock path; the well-conditioned backward checks are the fixed-upstream tests.
    assert wo < 0.4 and ws < 0.1, (wo, ws)


def test_cfg3_idro_g50_bert_base_vs_oracle():
    """16 triplets, G = 50 (most groups absent, a few shared): one iDRO step of the drop-in (grouped K11 path) vs
    heads_ref.idro_forward on the fp32 oracle encoder."""
    from oracle import bert_ref, heads_ref
    cfg = bert_ref.make_config()
    m = _build(cfg).train()
    G, B, L = 50, 16, 64
    alpha, eps, ema, rho = 0.25, 0.01, 0.1, 0.05
    m.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=0), G, "idro", alpha, eps, ema, rho)
    gen = torch.Generator().manual_seed(3)
    gid = torch.randint(0, 12, (B,), generator=gen)  # 12 of the 50 groups can appear, several twice or more
    batch = []
    for k in range(3):
        batch += list(bert_ref.synth_batch(B, L, cfg["vocab"], 500 + k))
    dev_batch = [t.cuda() for t in batch]
    h0 = m.loss.h_fun.clone()
    robust, acc, gl, gc = m(*dev_batch, group_ids=gid.cuda(), weights=torch.ones(B, device="cuda"))
    robust.backward()
    torch.set_num_threads(os.cpu_count() or 1)
    leaf = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(cfg, 0).items()}
    embs = [bert_ref.cls_embedding(leaf, batch[2 * k], batch[2 * k + 1], cfg) for k in range(3)]
    losses = heads_ref.pair_nll(*embs)[0]
    names = heads_ref.idro_param_names(["bert." + n for n in leaf], "base")
    params = [leaf[n[5:]] for n in names]
    r_ref, m_ref, c_ref, h_ref = heads_ref.idro_forward(losses, gid, params, h0.cpu(), G, alpha, ema, rho, eps)
    _record("cfg3_idro", {"robust": [robust.item(), r_ref.item()], "h_fun_rel": _rel(m.loss.h_fun.cpu().numpy(), h_ref.numpy()),
                          "groups_present": int((c_ref > 0).sum()), "P_last": int(sum(p.numel() for p in params))})
    assert sum(p.numel() for p in params) == 21_263_616
    assert abs(robust.item() - r_ref.item()) < 1e-2 * abs(r_ref.item())
    # group means are averages of log(1 + exp(l- - l+)) with |l| ~ 10^2..10^3 for LayerNorm-ed 768-d embeddings: the
    # north-star bound (logits within 1e-2 relative) allows far more than the 3e-2 absolute slack used here
    np.testing.assert_allclose(gl.cpu().numpy(), m_ref.numpy(), rtol=1e-2, atol=3e-2)
    np.testing.assert_array_equal(gc.cpu().numpy(), c_ref.numpy())
    np.testing.assert_allclose(m.loss.h_fun.cpu().numpy(), h_ref.numpy(), rtol=1e-2, atol=1e-4)


def test_cfg4_bert_large_full_depth_forward_vs_oracle():
    from transformers import BertConfig
    from cocodr_b200.bert import BertModel
    from oracle import bert_ref
    cfg = bert_ref.make_config(hidden=1024, layers=24, heads=16, inter=4096)
    hf = BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    st = bert_ref.synth_state(cfg, 0)
    m = BertModel(hf, add_pooling_layer=False)
    m.load_state_dict(st, strict=False)
    m = m.cuda().eval()
    ids, mask = bert_ref.synth_batch(8, 256, cfg["vocab"], 42)
    with torch.no_grad():
        out = m(ids.cuda(), mask.cuda(), output_hidden_states=True)
        torch.set_num_threads(os.cpu_count() or 1)
        ref, hs = bert_ref.encoder_fwd(st, ids, mask, cfg, output_hidden_states=True)
    real = mask.bool()
    errs = [_rel(a.cpu()[real].numpy(), b[real].numpy()) for a, b in zip(out.hidden_states, hs)]
    cls_err = _rel(out.last_hidden_state[:, 0].cpu().numpy(), ref[:, 0].numpy())
    _record("cfg4_large", {"cls_rel": cls_err, "hidden_rel_per_layer": errs})
    assert cls_err < 1e-2 and max(errs) < 1e-2, (cls_err, errs)
    assert len(errs) == 25


def test_grad_scaler_round_trip():
    """torch GradScaler around the drop-in (run_ann.py --fp16 / HF Trainer scale the loss).  The activation gradients
    travel in fp16 multiplied by the caller's loss scale AND the internal static factor (ops.get_grad_scale(), 2^10):
      * with the internal factor set to 1 (the external scaler already lifts the gradients) the default
        2^16 scale works at once and the unscaled gradients equal the unscaled run;
      * with both factors the first steps overflow to inf -- never silently: the scaler sees non-finite gradients,
        skips and backs off until the product fits, then trains with correct gradients."""
    import test_model_gpu as T
    from cocodr_b200 import ops
    m = T.build(T.TINY, "BertDot_InBatch_NLL_LN").train()
    q, mq, a, ma, b, mb = T.triplet(T.TINY, 8, 32, 99)
    w = torch.ones(8, device="cuda")
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.0)
    m(q, mq, a, ma, weights=w)[0].backward()
    plain = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}

    def scaled_step(scaler):
        m.zero_grad(set_to_none=True)
        scaler.scale(m(q, mq, a, ma, weights=w)[0]).backward()
        scaler.unscale_(opt)
        finite = all(torch.isfinite(p.grad).all().item() for p in m.parameters() if p.grad is not None)
        scaler.step(opt)
        scaler.update()
        return finite

    def worst_err():
        worst = 0.0
        for n, p in m.named_parameters():
            if p.grad is None or "key.bias" in n:
                continue
            worst = max(worst, (p.grad - plain[n]).abs().max().item() / (plain[n].abs().max().item() + 1e-30))
        return worst

    old = ops.get_grad_scale()
    try:
        ops.set_grad_scale(1.0)
        scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
        assert scaled_step(scaler) and scaler.get_scale() == 65536.0
        assert worst_err() < 2e-2   # fp16 activation gradients round differently at another scale, nothing more
    finally:
        ops.set_grad_scale(old)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
    skipped = 0
    while not scaled_step(scaler):
        skipped += 1
        assert skipped < 20, "the scaler never found a finite scale"
    assert scaler.get_scale() == 65536.0 * 0.5 ** skipped
    assert worst_err() < 2e-2
    _record("grad_scaler", {"internal_scale": old, "skipped_steps_at_init_scale_65536": skipped,
                            "settled_scale": scaler.get_scale()})
