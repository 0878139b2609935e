"""GPU: contrastive heads / DRO statistics through the C ABI vs the CPU oracle (oracle/heads_ref.py) and
the reference-generated fixture tests/golden/contrastive.npz."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def contrastive_inputs(n, h, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(h, generator=g) * 14.7 / h ** 0.5
    trained = base[None, :] + 0.015 * torch.randn(n, h, generator=g)
    gauss = torch.randn(n, h, generator=g) * 0.5
    return {"trained": trained, "gauss": gauss}


def test_pair_nll_matches_oracle():
    from cocodr_b200 import kernels as k
    from oracle import heads_ref
    torch.manual_seed(0)
    n, d = 37, 768
    q, a, b = (torch.randn(n, d) * 0.5 for _ in range(3))
    a[3] = b[3]  # tie -> argmax 0
    leaf = [t.clone().requires_grad_(True) for t in (q, a, b)]
    loss_r, acc_r, logit_r = heads_ref.pair_nll(*leaf)
    w = torch.rand(n)
    (loss_r * w).sum().backward()
    qc, ac, bc = (t.cuda() for t in (q, a, b))
    loss, accs, logits = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, 2, device="cuda")
    k.pair_nll_fwd(qc, ac, bc, loss, accs, logits)
    np.testing.assert_allclose(loss.cpu().numpy(), loss_r.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), logit_r.detach().numpy(), rtol=1e-4, atol=1e-4)
    assert (accs.cpu() == acc_r).all()
    dq, da, db = (torch.empty(n, d, device="cuda") for _ in range(3))
    k.pair_nll_bwd(qc, ac, bc, logits, w.cuda(), dq, da, db)
    for got, ref in zip((dq, da, db), leaf):
        np.testing.assert_allclose(got.cpu().numpy(), ref.grad.numpy(), rtol=1e-4, atol=1e-5)


def _coco(e, row_offset=0, n_rows=None, loss_scale=1.0, dloss=None):
    from cocodr_b200 import kernels as k
    n, h = e.shape
    n_rows = n if n_rows is None else n_rows
    ec = e.cuda().contiguous()
    q = ec[row_offset:row_offset + n_rows]
    loss, lse = torch.empty(n_rows, device="cuda"), torch.empty(n_rows, device="cuda")
    k.simmat_ce_fwd(q, ec, loss, lse, mode=k.SIM_COCO, row_offset=row_offset, loss_scale=loss_scale)
    dq, dk = torch.empty(n_rows, h, device="cuda"), torch.empty(n, h, device="cuda")
    dl = torch.full((n_rows,), 1.0 / n_rows, device="cuda") if dloss is None else dloss.cuda()
    k.simmat_ce_bwd(q, ec, lse, dl, dq, dk, mode=k.SIM_COCO, row_offset=row_offset, loss_scale=loss_scale)
    return loss.cpu(), dq.cpu(), dk.cpu()


@pytest.mark.parametrize("n,h,nm", [(16, 128, "small"), (512, 1024, "cfg4")])
def test_coco_contrastive_matches_reference_golden(golden_dir, n, h, nm):
    g = np.load(os.path.join(golden_dir, "contrastive.npz"))
    for tag, e in contrastive_inputs(n, h, 11).items():
        loss, dq, dk = _coco(e)
        grad = dq + dk  # every row is both a query row and a key
        np.testing.assert_allclose(loss.numpy(), g[f"{nm}_{tag}_loss"], rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(grad[:8].numpy(), g[f"{nm}_{tag}_grad_head"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(grad.norm(dim=1).numpy(), g[f"{nm}_{tag}_grad_rownorm"], rtol=2e-3, atol=2e-5)


def test_coco_local_rows_and_world_scale():
    """rank-local rows of the gathered matrix, loss * world (COCO/modeling.py:247)."""
    from oracle import heads_ref
    e = contrastive_inputs(32, 64, 5)["gauss"]
    ref = heads_ref.coco_contrastive(e, world_size=4)
    loss, _, _ = _coco(e, row_offset=8, n_rows=8, loss_scale=4.0)
    np.testing.assert_allclose(loss.numpy(), ref[8:16].numpy(), rtol=1e-4, atol=1e-4)


def test_qp_infonce_matches_oracle():
    from cocodr_b200 import kernels as k
    from oracle import heads_ref
    torch.manual_seed(1)
    B, W, h, r = 16, 4, 96, 2
    Q = (torch.randn(B, h) * 0.4).requires_grad_(True)
    P = (torch.randn(B * W, h) * 0.4).requires_grad_(True)
    ref = heads_ref.qp_infonce(Q, P, r * B + torch.arange(B))
    w = torch.rand(B)
    (ref * w).sum().backward()
    Qc, Pc = Q.detach().cuda(), P.detach().cuda()
    loss, lse = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    k.simmat_ce_fwd(Qc, Pc, loss, lse, mode=k.SIM_QP, row_offset=r * B)
    np.testing.assert_allclose(loss.cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-5)
    dq, dk = torch.empty(B, h, device="cuda"), torch.empty(B * W, h, device="cuda")
    k.simmat_ce_bwd(Qc, Pc, lse, w.cuda(), dq, dk, mode=k.SIM_QP, row_offset=r * B)
    np.testing.assert_allclose(dq.cpu().numpy(), Q.grad.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dk.cpu().numpy(), P.grad.numpy(), rtol=1e-3, atol=1e-5)


def test_group_reduce_and_gram():
    from cocodr_b200 import kernels as k
    from oracle import heads_ref
    torch.manual_seed(2)
    n, G = 64, 50
    loss, g = torch.rand(n), torch.randint(0, G, (n,))
    sums_r, cnts_r, _ = heads_ref.group_stats(loss, g, G)
    sums, cnts = torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    k.group_reduce_fwd(loss.cuda(), g.cuda(), sums, cnts, n_groups=G)
    np.testing.assert_allclose(sums.cpu().numpy(), sums_r.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(cnts.cpu().numpy(), cnts_r.numpy())
    ds = torch.rand(G)
    dl = torch.empty(n, device="cuda")
    k.group_reduce_bwd(ds.cuda(), g.cuda(), dl, n_groups=G)
    np.testing.assert_array_equal(dl.cpu().numpy(), ds[g].numpy())

    X = torch.randn(G, 100003)  # row stride not a multiple of 4 floats: the fp32 CUDA-core kernel
    gram = torch.zeros(G, G, device="cuda")
    k.gram_f32(X.cuda(), gram)
    ref = (X.double() @ X.double().t()).float()
    np.testing.assert_allclose(gram.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-2)


@pytest.mark.parametrize("G,P", [(50, 1_000_000), (64, 4096), (7, 100_004), (1, 130), (33, 21_263_616 // 8)])
def test_gram_tensor_core_path(G, P):
    """cdr_gram_f32 through tcgen05 kind::tf32 (16-byte aligned rows): operands keep 10 mantissa bits, accumulation is
    fp32 -- entries within 2e-3 of the fp64 Gram relative to the row norms, ragged column tail and G < 64 included;
    accumulates into the given matrix."""
    from cocodr_b200 import kernels as k
    g = torch.Generator().manual_seed(G + P)
    X = torch.randn(G, P, generator=g)
    X[:, ::3] *= 0.01
    if G > 2:
        X[2] = 0.5 * X[1] + 0.1 * X[2]  # a strongly correlated pair
    gram = torch.ones(G, G, device="cuda")
    k.gram_f32(X.cuda(), gram)
    ref = X.double() @ X.double().t()
    nrm = torch.sqrt(torch.diagonal(ref))
    err = ((gram.cpu().double() - 1.0 - ref).abs() / (nrm[:, None] * nrm[None, :])).max().item()
    assert err < 2e-3, err


@pytest.mark.parametrize("n_rows,n_keys,dim,mode,off", [(64, 512, 768, "qp", 128), (7, 23, 40, "qp", 5), (37, 130, 96, "qp", 0),
                                                        (512, 512, 1024, "coco", 0), (20, 64, 128, "coco", 16),
                                                        (3, 6, 8, "coco", 2), (200, 1000, 2048, "qp", 300),
                                                        # > 2^20 scores: the tiled kernels (smaller problems take the
                                                        # one-block-per-vector kernels)
                                                        (1100, 1200, 96, "qp", 50), (1024, 2048, 64, "coco", 512)])
def test_fused_simmat_ragged_shapes(n_rows, n_keys, dim, mode, off):
    """The fused K9 / K9' kernels (scores never written; small problems: one block per vector, large ones: score tiles in
    registers, online softmax, recomputed in the backward) on shapes that do not divide the tilings, with few and many
    key splits, vs the torch restatement."""
    from cocodr_b200 import kernels as k
    g = torch.Generator().manual_seed(n_rows * 7 + n_keys)
    K_ = (torch.randn(n_keys, dim, generator=g) * (3.0 / dim ** 0.5)).requires_grad_(True)
    if mode == "coco":
        Q = K_[off:off + n_rows]
        S = Q @ K_.t()
        idx = torch.arange(n_rows)
        S = S.masked_fill(torch.nn.functional.one_hot(off + idx, n_keys).bool(), float("-inf"))
        tgt = (off + idx) ^ 1
        Qd = Q.detach()
    else:
        Qd = (torch.randn(n_rows, dim, generator=g) * (3.0 / dim ** 0.5)).requires_grad_(True)
        S = Qd @ K_.t()
        tgt = off + torch.arange(n_rows)
    ref = torch.nn.functional.cross_entropy(S, tgt, reduction="none") * 2.0
    w = torch.rand(n_rows, generator=g)
    (ref * w).sum().backward()
    qc, kc = Qd.detach().cuda().contiguous(), K_.detach().cuda()
    loss, lse = torch.empty(n_rows, device="cuda"), torch.empty(n_rows, device="cuda")
    k.simmat_ce_fwd(qc, kc, loss, lse, mode=k.SIM_COCO if mode == "coco" else k.SIM_QP, row_offset=off, loss_scale=2.0)
    np.testing.assert_allclose(loss.cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    dq, dk = torch.full((n_rows, dim), 7.0, device="cuda"), torch.full((n_keys, dim), 7.0, device="cuda")
    k.simmat_ce_bwd(qc, kc, lse, w.cuda(), dq, dk, mode=k.SIM_COCO if mode == "coco" else k.SIM_QP, row_offset=off, loss_scale=2.0)
    if mode == "coco":  # rows are a slice of the keys: autograd sums both roles into K_.grad
        tot = dk.cpu().clone()
        tot[off:off + n_rows] += dq.cpu()
        np.testing.assert_allclose(tot.numpy(), K_.grad.numpy(), rtol=2e-3, atol=2e-5)
    else:
        np.testing.assert_allclose(dq.cpu().numpy(), Qd.grad.numpy(), rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(dk.cpu().numpy(), K_.grad.numpy(), rtol=2e-3, atol=2e-5)
