"""GPU: embedding+LayerNorm and LayerNorm fwd/bwd kernels vs torch fp32 (oracle/bert_ref.py embeddings_fwd)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_seq,L,H", [(3, 32, 128), (4, 128, 768), (2, 64, 1024), (3, 7, 768), (2, 16, 1536)])
def test_ln_fwd_bwd(n_seq, L, H):
    from cocodr_b200 import kernels as k
    g = torch.Generator().manual_seed(H + L)
    T = n_seq * L
    x = torch.randn(T, H, generator=g).half().cuda()
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    beta = (0.1 * torch.randn(H, generator=g)).cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    cls = torch.empty(n_seq, H, device="cuda")
    k.ln_fwd(x, gamma, beta, y, mean, rstd, cls, n_seq=n_seq, seq_len=L, hidden=H, eps=1e-12)
    xf = x.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xf, (H,), gf, bf, 1e-12)
    assert (y.float() - yr).abs().max().item() < 5e-3
    assert (cls - yr.view(n_seq, L, H)[:, 0]).abs().max().item() < 1e-5

    dy = torch.randn(T, H, generator=g).half().cuda()
    dcls = torch.randn(n_seq, H, generator=g).cuda()
    dx = torch.empty_like(x)
    dgamma, dbeta, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    k.ln_bwd(dy, dcls, x, gamma, mean, rstd, dx, dgamma, dbeta, dbias, n_seq=n_seq, seq_len=L, hidden=H,
             in_scale=2.0, out_scale=0.5)
    dyt = dy.float().view(n_seq, L, H).clone()
    dyt[:, 0] += 2.0 * dcls  # in_scale lifts the fp32 CLS gradient into the (scaled) fp16 gradient domain
    (yr * dyt.view(T, H)).sum().backward()
    s = xf.grad.abs().max().item()
    assert (dx.float() - xf.grad).abs().max().item() < 4e-3 * s + 1e-3
    for got, ref in ((dgamma, 0.5 * gf.grad), (dbeta, 0.5 * bf.grad)):  # out_scale on parameter gradients
        assert (got - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 1e-3
    ref = 0.5 * xf.grad.sum(0)
    assert (dbias - ref).abs().max().item() < 5e-3 * ref.abs().max().item() + 2e-2


@pytest.mark.parametrize("n_seq,L,H", [(3, 32, 128), (16, 128, 768), (2, 64, 1024), (5, 8, 64), (3, 7, 768), (1, 1, 768),
                                       (2, 16, 1536), (64, 128, 768)])
def test_ln_bwd_split_path(n_seq, L, H):
    """fp16 dy only (no fp32 CLS gradient): the staged single-pass backward for hidden <= 1024 (ragged last tile
    included), the two-pass backward (dx pass + column-sum pass) above that."""
    from cocodr_b200 import kernels as k
    g = torch.Generator().manual_seed(H + L + 1)
    T = n_seq * L
    x = torch.randn(T, H, generator=g).half().cuda()
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).cuda()
    beta = (0.1 * torch.randn(H, generator=g)).cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    k.ln_fwd(x, gamma, beta, y, mean, rstd, None, n_seq=n_seq, seq_len=L, hidden=H, eps=1e-12)
    dy = torch.randn(T, H, generator=g).half().cuda()
    dx = torch.empty_like(x)
    dgamma, dbeta, dbias = (torch.ones(H, device="cuda") for _ in range(3))  # accumulate semantics
    ws = torch.empty(2 * T, device="cuda")
    k.ln_bwd(dy, None, x, gamma, mean, rstd, dx, dgamma, dbeta, dbias, n_seq=n_seq, seq_len=L, hidden=H,
             out_scale=0.25, row_ws=ws)
    xf = x.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    (F.layer_norm(xf, (H,), gf, bf, 1e-12) * dy.float()).sum().backward()
    assert (dx.float() - xf.grad).abs().max().item() < 4e-3 * xf.grad.abs().max().item() + 1e-3
    for got, ref in ((dgamma, gf.grad), (dbeta, bf.grad), (dbias, xf.grad.sum(0))):
        ref = 1.0 + 0.25 * ref
        assert (got - ref).abs().max().item() < 3e-3 * ref.abs().max().item() + 5e-3


def test_embed_ln_fwd_bwd():
    from cocodr_b200 import kernels as k
    from oracle import bert_ref
    cfg = bert_ref.make_config(hidden=128, layers=1, heads=2, inter=512, vocab=2000, max_pos=64)
    st = {n: t.cuda() for n, t in bert_ref.synth_state(cfg, 3).items() if n.startswith("embeddings")}
    ids, mask = bert_ref.synth_batch(6, 32, cfg["vocab"], 5)
    ids = ids.cuda()
    n_seq, L, H = 6, 32, 128
    T = n_seq * L
    out = torch.empty(T, H, dtype=torch.float16, device="cuda")
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    w, p, t = (st[f"embeddings.{n}_embeddings.weight"] for n in ("word", "position", "token_type"))
    gam, bet = st["embeddings.LayerNorm.weight"], st["embeddings.LayerNorm.bias"]
    k.embed_ln_fwd(ids, w, p, t[0].contiguous(), gam, bet, out, mean, rstd, n_seq=n_seq, seq_len=L, hidden=H,
                   vocab=cfg["vocab"], eps=1e-12)
    leaf = {n: v.clone().requires_grad_(True) for n, v in st.items()}
    ref = bert_ref.embeddings_fwd(leaf, ids, cfg).view(T, H)
    assert (out.float() - ref).abs().max().item() < 5e-3
    dy = torch.randn(T, H).half().cuda()
    (ref * dy.float()).sum().backward()
    dword, dpos = torch.zeros_like(w), torch.zeros_like(p)
    dtype0, dgam, dbet = (torch.zeros(H, device="cuda") for _ in range(3))
    k.embed_ln_bwd(dy, ids, w, p, t[0].contiguous(), gam, mean, rstd, dword, dpos, dtype0, dgam, dbet, n_seq=n_seq,
                   seq_len=L, hidden=H, vocab=cfg["vocab"], pad_id=0, in_scale=1.0, out_scale=1.0)
    for got, name in ((dword, "embeddings.word_embeddings.weight"), (dpos, "embeddings.position_embeddings.weight"),
                      (dgam, "embeddings.LayerNorm.weight"), (dbet, "embeddings.LayerNorm.bias")):
        r = leaf[name].grad
        assert (got - r).abs().max().item() < 3e-3 * r.abs().max().item() + 1e-3, name
    r = leaf["embeddings.token_type_embeddings.weight"].grad[0]
    assert (dtype0 - r).abs().max().item() < 3e-3 * r.abs().max().item() + 1e-2
