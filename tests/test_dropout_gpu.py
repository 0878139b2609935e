"""GPU: fused counter-based dropout (through the C ABI) vs the oracle's regenerated masks (oracle/dropout_ref.py).

The masks are bit-exact by construction (same Philox counters), so every check is "same arithmetic given the same
mask": kernels vs torch-fp32 restatements with the oracle mask applied at HF's four nn.Dropout sites, and the whole
encoder (forward, fixed-upstream-gradient backward) vs oracle/bert_ref.py with the same DropSpec -- at the same
tolerances as the p = 0 tests (1e-2 on embeddings, 2.5e-2 on parameter gradients)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SEED, OFF, P = 20240611, 5, 0.1


def _state(seed=SEED, off=OFF):
    return torch.tensor([seed, off], dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("rows,cols,row_mul", [(37, 768, 1), (128, 1024, 1), (5, 64, 128), (1000, 8, 3)])
def test_dropout_kernel_matches_oracle_mask_bit_exactly(rows, cols, row_mul):
    from cocodr_b200 import kernels as k
    from oracle import dropout_ref
    x = torch.randn(rows, cols, generator=torch.Generator().manual_seed(rows)).half().cuda()
    out = torch.empty_like(x)
    k.dropout_f16(x, out, drop=k.drop_args(_state(), 7, P, row_mul))
    m = dropout_ref.hidden_mask(rows, cols, 7, SEED, OFF, P, row_mul).cuda()
    assert torch.equal(out, (x.float() * m).half())
    assert 0.85 < (out != 0).float().mean().item() < 0.95 or rows * cols < 2000
    # in place
    y = x.clone()
    k.dropout_f16(y, y, drop=k.drop_args(_state(), 7, P, row_mul))
    assert torch.equal(y, out)


@pytest.mark.parametrize("M,N,K,row_mul", [(256, 768, 768, 1), (1024, 768, 3072, 1), (100, 128, 64, 1), (64, 768, 768, 128),
                                           (300, 1024, 256, 1)])
def test_gemm_dropout_residual_epilogue(M, N, K, row_mul):
    """out = dropout(a b^T + bias) + residual (HF BertSelfOutput / BertOutput before the LayerNorm)."""
    from cocodr_b200 import kernels as k
    from oracle import dropout_ref
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    b = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).half().cuda()
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    k.gemm(a, b, out, M=M, N=N, K=K, bias=bias, aux=res, epilogue=k.EPI_BIAS_DROP_RESIDUAL,
           drop=k.drop_args(_state(), 14, P, row_mul))
    m = dropout_ref.hidden_mask(M, N, 14, SEED, OFF, P, row_mul).cuda()
    ref = (a.float() @ b.float().t() + bias) * m + res.float()
    assert (out.float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()
    # dropped positions are EXACTLY the residual (nothing of the dense output leaks through)
    dropped = m == 0
    assert torch.equal(out[dropped], res[dropped])


def _attn_ref(qkv, bias, n_seq, L, heads, mask):
    H = heads * 64
    x = qkv.view(n_seq, L, 3, heads, 64)
    q, kk, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = q @ kk.transpose(2, 3) * 0.125
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = torch.softmax(s, dim=-1) * mask
    return (p @ v).transpose(1, 2).reshape(n_seq * L, H), torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("n_seq,L,heads,masked", [(3, 128, 2, True), (64, 128, 12, False), (2, 100, 3, True),
                                                  (5, 32, 2, True), (3, 256, 2, True), (2, 200, 3, True),
                                                  (2, 512, 1, False)])
def test_attention_dropout_fwd_bwd(n_seq, L, heads, masked):
    """dropout(softmax(S)) V and its backward, one-tile (L <= 128) and tiled (L > 128) kernels."""
    from cocodr_b200 import kernels as k
    from oracle import dropout_ref
    g = torch.Generator().manual_seed(n_seq * 1000 + L)
    H, T = heads * 64, n_seq * L
    qkv = (torch.randn(T, 3 * H, generator=g) * 1.5).half().cuda()
    bias = None
    if masked:
        lens = torch.randint(1, L + 1, (n_seq,), generator=g)
        lens[0] = L
        bias = ((torch.arange(L)[None, :] >= lens[:, None]).float() * torch.finfo(torch.float32).min).cuda()
    site = 4 * 3 + 1
    drop = k.drop_args(_state(), site, P)
    out = torch.zeros(T, H, dtype=torch.float16, device="cuda")
    lse = torch.zeros(n_seq, heads, L, dtype=torch.float32, device="cuda")
    bits = k.attn_dropout_bits(n_seq, heads, L, "cuda")  # filled by the forward, read by the backward
    k.attn_fwd(qkv, bias, out, lse, n_seq=n_seq, seq_len=L, heads=heads, drop=drop, drop_bits=bits)
    mask = dropout_ref.attention_mask(n_seq, heads, L, site, SEED, OFF, P).cuda()
    qf = qkv.float().requires_grad_(True)
    o_ref, lse_ref = _attn_ref(qf, bias, n_seq, L, heads, mask)
    assert (out.float() - o_ref).abs().max().item() <= 4e-3 * max(1.0, o_ref.abs().max().item())
    assert (lse - lse_ref).abs().max().item() <= 2e-3  # the log-sum-exp is that of the UNDROPPED probabilities
    d_out = torch.randn(T, H, generator=g).half().cuda()
    dqkv = torch.zeros(T, 3 * H, dtype=torch.float16, device="cuda")
    dbias = torch.zeros(3 * H, dtype=torch.float32, device="cuda") if L <= 128 else None
    k.attn_bwd(qkv, bias, out, lse, d_out, dqkv, n_seq=n_seq, seq_len=L, heads=heads, dbias=dbias, dbias_scale=1.0,
               drop=drop, drop_bits=bits)
    (o_ref * d_out.float()).sum().backward()
    ref = qf.grad
    for nm, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = (dqkv[:, sl].float() - ref[:, sl]).abs().max().item()
        assert e <= 1.5e-2 * ref[:, sl].abs().max().item(), f"{nm} err {e}"
    if dbias is not None:
        assert (dbias - ref.sum(0)).abs().max().item() <= 5e-3 * max(1.0, ref.abs().sum(0).max().item())
    # without the descriptor the same call is the p = 0 kernel: different output
    out0 = torch.zeros_like(out)
    k.attn_fwd(qkv, bias, out0, lse, n_seq=n_seq, seq_len=L, heads=heads)
    assert not torch.equal(out0, out)


@pytest.mark.parametrize("rows,H,row_mul", [(1024, 768, 1), (128, 768, 128), (77, 1024, 1), (9, 128, 1), (512, 256, 4)])
def test_ln_bwd_drop(rows, H, row_mul):
    """LayerNorm backward after a dropped dense output: dx as without dropout, dx_drop = mask * dx / (1 - p), and the
    dense bias gradient = column sums of dx_drop."""
    from cocodr_b200 import kernels as k
    from oracle import dropout_ref
    g = torch.Generator().manual_seed(rows + H)
    x = torch.randn(rows, H, generator=g).half().cuda()
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    beta = (0.1 * torch.randn(H, generator=g)).cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    k.ln_fwd(x, gamma, beta, y, mean, rstd, None, n_seq=rows, seq_len=1, hidden=H, eps=1e-12)
    dy = torch.randn(rows, H, generator=g).half().cuda()
    dx, dxm = torch.empty_like(x), torch.empty_like(x)
    dgamma, dbeta, dbias = (torch.ones(H, device="cuda") for _ in range(3))
    k.ln_bwd_drop(dy, x, gamma, mean, rstd, dx, dxm, dgamma, dbeta, dbias, rows=rows, hidden=H, out_scale=0.5,
                  drop=k.drop_args(_state(), 10, P, row_mul))
    dx0 = torch.empty_like(x)
    dg0, db0, dc0 = (torch.ones(H, device="cuda") for _ in range(3))
    k.ln_bwd(dy, None, x, gamma, mean, rstd, dx0, dg0, db0, dc0, n_seq=rows, seq_len=1, hidden=H, out_scale=0.5,
             row_ws=torch.empty(2 * rows, device="cuda"))
    assert torch.equal(dx, dx0)
    np.testing.assert_allclose(dgamma.cpu().numpy(), dg0.cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dbeta.cpu().numpy(), db0.cpu().numpy(), rtol=1e-4, atol=1e-4)
    m = dropout_ref.hidden_mask(rows, H, 10, SEED, OFF, P, row_mul).cuda()
    xf = x.float().requires_grad_(True)
    (F.layer_norm(xf, (H,), gamma, beta, 1e-12) * dy.float()).sum().backward()
    ref = xf.grad * m
    assert (dxm.float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 1e-3
    assert (dxm[m == 0] == 0).all()
    cs = 1.0 + 0.5 * ref.sum(0)
    assert (dbias - cs).abs().max().item() <= 5e-3 * (0.5 * ref).abs().sum(0).max().item() + 2e-2


TINY = dict(hidden=128, layers=4, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)


def _model(cfg, p_hidden=P, p_attn=P, cls_name="BertDot_NLL_LN"):
    from transformers import BertConfig
    from cocodr_b200 import models
    from oracle import bert_ref
    hf = BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                    num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                    max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                    hidden_dropout_prob=p_hidden, attention_probs_dropout_prob=p_attn, num_labels=2)
    m = getattr(models, cls_name)(hf)
    m.bert.load_state_dict(bert_ref.synth_state(cfg, 0), strict=False)
    return m.cuda()


def _rel(got, ref, floor=0.0):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), floor, 1e-12)


@pytest.mark.parametrize("p_hidden,p_attn", [(0.1, 0.1), (0.1, 0.0), (0.0, 0.2)])
def test_encoder_train_mode_matches_oracle_with_same_masks(p_hidden, p_attn):
    """Whole encoder in train() mode (all four dropout sites, the [CLS]-only last layer included) vs the fp32 oracle
    driven with the same (seed, offset): CLS embeddings 1e-2, every parameter gradient 2.5e-2 for a fixed upstream
    gradient."""
    from oracle import bert_ref, dropout_ref
    m = _model(TINY, p_hidden, p_attn).train()
    m.bert.set_dropout_seed(SEED)
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(6, 32, TINY["vocab"], 77))
    dcls = torch.randn(6, TINY["hidden"], generator=torch.Generator().manual_seed(5)) * 0.05
    cls = m.query_emb(ids, mask)  # first pass of this model: offset 1
    (cls * dcls.cuda()).sum().backward()
    spec = dropout_ref.DropSpec(SEED, 1, p_hidden, p_attn)
    leaf = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(TINY, 0).items()}
    ref_cls = bert_ref.cls_embedding(leaf, ids.cpu(), mask.cpu(), TINY, drop=spec)
    (ref_cls * dcls).sum().backward()
    err = _rel(cls.detach().cpu().numpy(), ref_cls.detach().numpy())
    assert err < 1e-2
    # ... and it is NOT the eval-mode embedding: the distance to the undropped oracle is several times the error
    # against the oracle that applies the same masks
    off = _rel(cls.detach().cpu().numpy(), bert_ref.cls_embedding(leaf, ids.cpu(), mask.cpu(), TINY).detach().numpy())
    assert off > 4 * max(err, 1e-3), (off, err)
    named = dict(m.bert.named_parameters())
    for name, ref in leaf.items():
        r = _rel(named[name].grad.cpu().numpy(), ref.grad.numpy(), floor=1e-4)
        assert r < 2.5e-2, f"{name}: rel err {r}"
    # second pass: offset 2 -> different masks, again equal to the oracle's
    with torch.no_grad():
        cls2 = m.query_emb(ids, mask)
    ref2 = bert_ref.cls_embedding(leaf, ids.cpu(), mask.cpu(), TINY, drop=dropout_ref.DropSpec(SEED, 2, p_hidden, p_attn))
    assert _rel(cls2.cpu().numpy(), ref2.detach().numpy()) < 1e-2
    # eval(): dropout off, deterministic
    m.eval()
    with torch.no_grad():
        e1, e2 = m.query_emb(ids, mask), m.query_emb(ids, mask)
    assert torch.equal(e1, e2)


def test_full_last_layer_and_hidden_states_use_the_same_masks():
    """BertModel.forward (full last layer, fp32 hidden states for callers) under dropout == oracle with the same spec;
    its row 0 equals the [CLS]-only path run at the same offset."""
    from oracle import bert_ref, dropout_ref
    m = _model(TINY).train()
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(4, 32, TINY["vocab"], 9))
    m.bert.set_dropout_seed(77)
    out = m.bert(ids, mask, output_hidden_states=True)
    st = bert_ref.synth_state(TINY, 0)
    ref, hs = bert_ref.encoder_fwd(st, ids.cpu(), mask.cpu(), TINY, output_hidden_states=True,
                                   drop=dropout_ref.DropSpec(77, 1, P, P))
    real = mask.bool().cpu()
    assert _rel(out.last_hidden_state.detach().cpu()[real].numpy(), ref[real].numpy()) < 1e-2
    for a, b in zip(out.hidden_states, hs):
        assert _rel(a.detach().cpu()[real].numpy(), b[real].numpy()) < 1e-2
    m.bert.set_dropout_seed(77)
    with torch.no_grad():
        cls = m.query_emb(ids, mask)
    assert _rel(cls.cpu().numpy(), out.last_hidden_state[:, 0].detach().cpu().numpy()) < 2e-3


def test_graph_replays_draw_fresh_masks():
    """The (seed, offset) state advances on the device: every replay of a captured training step sees new masks."""
    from cocodr_b200 import optim
    from cocodr_b200.graph import GraphedTrainStep
    from oracle import bert_ref
    m = _model(TINY, cls_name="BertDot_InBatch_NLL_LN").train()
    opt = optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=0.0, eps=1e-8, semantics="torch").attach_shadows(m)
    q, mq = (t.cuda() for t in bert_ref.synth_batch(4, 32, TINY["vocab"], 1))
    p, mp = (t.cuda() for t in bert_ref.synth_batch(4, 32, TINY["vocab"], 2))
    w = torch.ones(4, device="cuda")
    step = GraphedTrainStep(m, opt, (q, mq, p, mp, None, None, True, None, w))
    losses = [step(q, mq, p, mp).item() for _ in range(4)]
    assert len(set(losses)) == 4, losses  # lr = 0: only the masks change between replays
    m.eval()
    with torch.no_grad():
        a = m(q, mq, p, mp, weights=w)[0].item()
        b = m(q, mq, p, mp, weights=w)[0].item()
    assert a == b
