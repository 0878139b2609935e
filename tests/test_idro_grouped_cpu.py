"""CPU: host logic of the iDRO grouped-gradient path (dro_loss.iDROLoss._grouped_from_records, K11).

The layer backwards hand over their wgrad operands; this test feeds the reducer synthetic operands of a real
(torch, fp32) post-LN transformer layer whose per-group parameter gradients autograd can compute directly -- the way
the reference takes them (one ``autograd.grad`` per group, ANCE/model/dro_loss.py:192-204) -- with ``K.gemm`` replaced
by a torch restatement of the one configuration the reducer uses (both operands MN-major, fp32 accumulate-add).
What is under test is the row regrouping, the K ranges, the parameter-to-column map, the one-hot column sums and the
LayerNorm gradient formulas; the CUDA GEMM itself is covered by tests/test_gemm_gpu.py.
"""
import types

import pytest
import torch

from cocodr_b200 import dro_loss
from cocodr_b200 import kernels as K


def _fake_gemm(a, b, out, *, M, N, K, a_major=0, b_major=0, epilogue=0, split_k=1, alpha=1.0, **kw):  # noqa: N803
    from cocodr_b200 import kernels
    assert (a_major, b_major, epilogue) == (1, 1, kernels.EPI_F32_ATOMIC)
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and out.dtype == torch.float32
    assert a.shape == (K, M) and b.shape == (K, N) and out.shape == (M, N)
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.is_contiguous()
    out += alpha * (a.float().t() @ b.float())
    return out


def _fake_gemm_segments(a, b, out, *, M, N, row_begin, row_count, out_offset, alpha=1.0, **kw):  # noqa: N803
    """cdr_gemm_segments (include/cocodr_b200.h) restated: segment i reduces rows [begin, begin + count) of a and b
    into the [M, N] fp32 block that starts out_offset[i] elements into out."""
    assert a.shape[1] == M and b.shape[1] == N and a.stride(1) == 1 and b.stride(1) == 1 and out.is_contiguous()
    flat = out.view(-1)
    for rb, rc, off in zip(row_begin, row_count, out_offset):
        assert rc > 0 and rb + rc <= a.shape[0]
        _fake_gemm(a[rb:rb + rc], b[rb:rb + rc], flat[off:off + M * N].view(M, N), M=M, N=N, K=rc, a_major=1, b_major=1,
                   epilogue=K.EPI_F32_ATOMIC, alpha=alpha)
    return out


def _fake_gemm_grouped(a, b, out, *, M, N, seg_kb, n_groups, out_group_stride, out_offset=0, alpha=1.0, **kw):  # noqa: N803
    """cdr_gemm_grouped restated: group g adds a[rows_g]^T b[rows_g], rows_g = 64-row k-blocks [seg_kb[g], seg_kb[g+1])."""
    assert seg_kb.dtype == torch.int32 and seg_kb.numel() == n_groups + 1
    assert a.shape[1] == M and b.shape[1] == N and a.shape[0] == b.shape[0] and a.stride(1) == 1 and b.stride(1) == 1
    kb = seg_kb.tolist()
    assert kb[0] == 0 and all(x <= y for x, y in zip(kb, kb[1:])) and kb[-1] * 64 <= a.shape[0]
    flat = out.view(-1)
    for g in range(n_groups):
        r0, r1 = kb[g] * 64, kb[g + 1] * 64
        if r1 > r0:
            o = out_offset + g * out_group_stride
            flat[o:o + M * N].view(M, N).add_(alpha * (a[r0:r1].float().t() @ b[r0:r1].float()))
    return out


def _layer(x, p, n_seq, L, cls_only):
    """A post-LN layer restricted to what the reducer sees: y = LN2(x1 + W2 gelu(W1 x1 + b1) + b2),
    x1 = LN1(x + Wo att + bo), att = a fixed per-sequence mixing of V-like projections, qkv = x Wqkv^T + b."""
    wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, wi, bi, wo2, bo2, g2, be2 = p
    qkv = torch.cat([x @ wq.t() + bq, x @ wk.t() + bk, x @ wv.t() + bv], 1)
    H = x.shape[1]
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    att = torch.tanh(q) * torch.sigmoid(k) + v  # stand-in for attention (keeps samples separate)
    att = att + att.view(n_seq, L, H).mean(1, keepdim=True).expand(n_seq, L, H).reshape(-1, H)  # mixes a sequence's rows
    if cls_only:  # ops.BertLastLayerCLSFn: everything after the attention core runs on row 0 of every sequence
        x, att = x.view(n_seq, L, H)[:, 0], att.view(n_seq, L, H)[:, 0]
    y1 = x + att @ wo.t() + bo
    x1 = torch.nn.functional.layer_norm(y1, (H,), g1, be1, 1e-12)
    z = x1 @ wi.t() + bi
    gl = torch.nn.functional.gelu(z)
    y2 = x1 + gl @ wo2.t() + bo2
    y = torch.nn.functional.layer_norm(y2, (H,), g2, be2, 1e-12)
    return dict(qkv=qkv, att=att, y1=y1, x1=x1, z=z, gl=gl, y2=y2, y=y)


@pytest.mark.parametrize("cls_only,L,one_launch", [(False, 4, False), (True, 4, False), (False, 4, True),
                                                  (True, 4, True), (False, 64, True), (True, 64, True)])
def test_grouped_reducer_matches_per_group_autograd(monkeypatch, cls_only, L, one_launch):  # noqa: N803
    monkeypatch.setattr(K, "gemm", _fake_gemm)
    monkeypatch.setattr(K, "gemm_segments", _fake_gemm_segments)
    monkeypatch.setattr(K, "gemm_grouped", _fake_gemm_grouped)
    monkeypatch.setattr(dro_loss.iDROLoss, "grouped_kernel", one_launch)
    rps = 1 if cls_only else L
    torch.manual_seed(3 + rps)
    B, towers, G, H, I = 5, 3, 4, 8, 16  # noqa: E741
    n_seq = B * towers
    rows = n_seq * L
    S = 64.0
    g = torch.tensor([2, 0, 2, 3, 0])  # group 1 absent
    counts = torch.bincount(g, minlength=G).float()
    shapes = [(H, H), (H,), (H, H), (H,), (H, H), (H,), (H, H), (H,), (H,), (H,), (I, H), (I,), (H, I), (H,), (H,), (H,)]
    params = [torch.nn.Parameter(torch.randn(s) * (0.3 if len(s) == 2 else 0.1) + (1.0 if i in (8, 14) else 0.0))
              for i, s in enumerate(shapes)]
    x = torch.randn(rows, H)
    t = _layer(x, params, n_seq, L, rps == 1)
    # per-sample loss: a fixed random functional of the sample's own rows; group means as in iDROLoss.forward
    wsel = torch.randn(n_seq * rps, H)
    per_seq = (t["y"] * wsel).view(n_seq, -1).sum(1)
    loss = per_seq.view(towers, B).sum(0)
    means = torch.zeros(G).index_add(0, g, loss) / (counts + (counts == 0).float())
    want = torch.zeros(G, sum(p.numel() for p in params))
    for gi in range(G):
        if counts[gi] > 0:
            gr = torch.autograd.grad(means[gi], params, retain_graph=True)
            want[gi] = torch.cat([v.reshape(-1) for v in gr])

    # the operands a layer backward would hand over, for the backward of means.sum(), in the S-scaled fp16 domain
    inter = [t[n] for n in ("y", "y2", "z", "x1", "y1", "att", "qkv")]
    d_y, d_y2, d_z, d_x1tot, d_y1, d_att, d_qkv = torch.autograd.grad(means.sum(), inter, retain_graph=True)
    # d(x1) as the LayerNorm-1 backward sees it: its incoming gradient is the total d(x1); dx1 in ops.py is exactly that
    mean1, rstd1 = t["y1"].mean(1), (t["y1"].var(1, unbiased=False) + 1e-12).rsqrt()
    mean2, rstd2 = t["y2"].mean(1), (t["y2"].var(1, unbiased=False) + 1e-12).rsqrt()
    h = lambda v: (v * S).half()  # noqa: E731
    rec = dict(keys=tuple(id(p) for p in params), rps=rps, L=L, n_seq=n_seq, S=S,
               dy2=h(d_y2), gl=t["gl"].detach().half(), dz=h(d_z), x1=t["x1"].detach().half(), dy1=h(d_y1),
               att=t["att"].detach().half(), dx1=h(d_x1tot), y1=t["y1"].detach().half(), mean1=mean1.detach(),
               rstd1=rstd1.detach(), y2=t["y2"].detach().half(), mean2=mean2.detach(), rstd2=rstd2.detach(),
               din2=(h(d_y), None), dqkv=h(d_qkv), x=x.half())
    if rps == 1:  # the [CLS]-only layer hands over an unscaled fp32 [CLS] gradient instead of a scaled fp16 one
        rec["din2"] = (None, d_y.detach().clone())

    crit = dro_loss.iDROLoss(types.SimpleNamespace(model_size="base", local_rank=0), G, 0.25, 0.01, 0.1, 0.05)
    got = crit._grouped_from_records([rec], params, counts, g, towers)
    assert got.shape == want.shape
    assert torch.count_nonzero(got[1]) == 0  # absent group: zero row, as in the reference (dro_loss.py:199-201)
    off = 0
    for i, p in enumerate(params):
        n = p.numel()
        w_, g_ = want[:, off:off + n], got[:, off:off + n]
        err = (w_ - g_).abs().max().item()
        assert err <= 2e-2 * w_.abs().max().item() + 1e-4, (i, tuple(p.shape), err, w_.abs().max().item())
        off += n


def test_grouped_reducer_rejects_uncovered_parameters(monkeypatch):
    monkeypatch.setattr(K, "gemm", _fake_gemm)
    monkeypatch.setattr(K, "gemm_segments", _fake_gemm_segments)
    crit = dro_loss.iDROLoss(types.SimpleNamespace(model_size="base", local_rank=0), 2, 0.25, 0.01, 0.1, 0.05)
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(RuntimeError):
        crit._grouped_from_records([], [p], torch.tensor([1.0, 1.0]), torch.tensor([0, 1]), 3)
