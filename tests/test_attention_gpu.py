"""GPU: tcgen05 fused attention fwd/bwd (through the C ABI) vs a torch fp32 restatement of
HF eager_attention_forward (oracle/bert_ref.py layer_fwd attention block)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, bias, n_seq, L, heads):
    H = heads * 64
    x = qkv.float().view(n_seq, L, 3, heads, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))  # [B, h, L, 64]
    s = q @ k.transpose(2, 3) * 0.125
    if bias is not None:
        s = s + bias[:, None, None, :]
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(n_seq * L, H), lse


@pytest.mark.parametrize("n_seq,L,heads,masked", [(2, 128, 2, False), (3, 64, 2, True), (5, 32, 2, True),
                                                  (4, 128, 12, True), (2, 100, 3, True), (64, 128, 12, False),
                                                  (3, 256, 2, True), (2, 200, 3, True), (2, 384, 2, False),
                                                  (2, 512, 1, True), (8, 256, 16, True)])
def test_attention_fwd_bwd(n_seq, L, heads, masked):
    from cocodr_b200 import kernels as k
    g = torch.Generator().manual_seed(n_seq * 1000 + L)
    H = heads * 64
    T = n_seq * L
    qkv = (torch.randn(T, 3 * H, generator=g) * 1.5).half().cuda()
    bias = None
    if masked:
        lens = torch.randint(1, L + 1, (n_seq,), generator=g)
        lens[0] = L
        bias = ((torch.arange(L)[None, :] >= lens[:, None]).float() * torch.finfo(torch.float32).min).cuda()
    out = torch.zeros(T, H, dtype=torch.float16, device="cuda")
    lse = torch.zeros(n_seq, heads, L, dtype=torch.float32, device="cuda")
    k.attn_fwd(qkv, bias, out, lse, n_seq=n_seq, seq_len=L, heads=heads)
    torch.cuda.synchronize()

    qf = qkv.float().requires_grad_(True)
    o_ref, lse_ref = _ref(qf, bias, n_seq, L, heads)
    err = (out.float() - o_ref).abs().max().item()
    assert err <= 4e-3 * max(1.0, o_ref.abs().max().item()), f"fwd max err {err}"
    assert (lse - lse_ref).abs().max().item() <= 2e-3

    d_out = (torch.randn(T, H, generator=g)).half().cuda()
    dqkv = torch.zeros(T, 3 * H, dtype=torch.float16, device="cuda")
    fused_db = L <= 128  # the one-tile backward can emit the QKV bias gradient (column sums of dQKV) as well
    dbias = torch.ones(3 * H, dtype=torch.float32, device="cuda") if fused_db else None
    k.attn_bwd(qkv, bias, out, lse, d_out, dqkv, n_seq=n_seq, seq_len=L, heads=heads, dbias=dbias, dbias_scale=0.5)
    torch.cuda.synchronize()
    (o_ref * d_out.float()).sum().backward()
    ref = qf.grad
    if fused_db:
        cs = 1.0 + 0.5 * ref.sum(0)
        e = (dbias - cs).abs().max().item()
        assert e <= 5e-3 * max(1.0, (0.5 * ref).abs().sum(0).max().item()), f"fused bias gradient err {e}"
    err = (dqkv.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item(), f"bwd max err {err} vs scale {ref.abs().max().item()}"
    for nm, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = (dqkv[:, sl].float() - ref[:, sl]).abs().max().item()
        assert e <= 1.5e-2 * ref[:, sl].abs().max().item(), f"{nm} err {e}"


def test_attention_rejects_long_sequences():
    from cocodr_b200 import kernels as k
    qkv = torch.zeros(640, 192, dtype=torch.float16, device="cuda")
    out = torch.zeros(640, 64, dtype=torch.float16, device="cuda")
    lse = torch.zeros(1, 1, 640, device="cuda")
    with pytest.raises(RuntimeError):
        k.attn_fwd(qkv, None, out, lse, n_seq=1, seq_len=640, heads=1)
