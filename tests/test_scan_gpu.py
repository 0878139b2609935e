"""GPU: corpus scan (cocodr_b200.scan.search -> cdr_scan_topk) vs the CPU oracle (oracle/scan_ref.py) and the
committed fixture tests/golden/scan_small.npz.  Bit-exact ranks on the exact-arithmetic corpus."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_scan_small_golden(golden_dir):
    from cocodr_b200 import scan
    from oracle import scan_ref
    g = np.load(os.path.join(golden_dir, "scan_small.npz"))
    for kind in ("exact", "gauss"):
        Q, P = scan_ref.synth_corpus(6000, 37, 768, seed=7, kind=kind)
        D, I = scan.search(Q.cuda(), P.cuda(), 100)
        D, I = D.cpu().numpy(), I.cpu().numpy()
        if kind == "exact":
            assert (I == g["exact_I"]).all()
            np.testing.assert_array_equal(D, g["exact_D"])
        else:
            np.testing.assert_allclose(D, g["gauss_D"], rtol=1e-5, atol=1e-4)
            assert (I == g["gauss_I"]).mean() > 0.999


@pytest.mark.parametrize("n_docs,n_q,k", [(200_000, 64, 100), (120_000, 200, 1000), (9000, 8, 10)])
def test_scan_sampled_path_exact_ranks(n_docs, n_q, k):
    """Large enough that the thresholds come from the corpus sample; exact corpus => bit-exact ranks."""
    from cocodr_b200 import scan
    from oracle import scan_ref
    Q, P = scan_ref.synth_corpus(n_docs, n_q, 768, seed=11, kind="exact")
    D, I = scan.search(Q.cuda(), P.cuda(), k)
    Dr, Ir = scan_ref.search(Q, P, k)
    assert (I.cpu().numpy() == Ir).all()
    np.testing.assert_array_equal(D.cpu().numpy(), Dr)


def test_scan_forced_chunked_fallback_and_sharded_merge():
    from cocodr_b200 import scan
    from oracle import scan_ref
    Q, P = scan_ref.synth_corpus(30_000, 16, 256, seed=3, kind="exact")
    Dr, Ir = scan_ref.search(Q, P, 50)
    D, I = scan.search(Q.cuda(), P.cuda(), 50, force_exhaustive=True)
    assert (I.cpu().numpy() == Ir).all()
    # two document shards with global ids, merged like the multi-GPU path
    Pc = P.cuda()
    parts = [scan.search(Q.cuda(), Pc[lo:hi], 50, doc_base=lo) for lo, hi in ((0, 17_000), (17_000, 30_000))]
    Dm, Im = scan.merge_topk(torch.cat([p[0] for p in parts], 1), torch.cat([p[1] for p in parts], 1), 50)
    assert (Im.cpu().numpy() == Ir).all()
    np.testing.assert_array_equal(Dm.cpu().numpy(), Dr)


def test_scan_edge_cases():
    from cocodr_b200 import scan
    from oracle import scan_ref
    Q, P = scan_ref.synth_corpus(40, 3, 64, seed=1, kind="exact")
    D, I = scan.search(Q.cuda(), P.cuda(), 100)  # k > n_docs -> clamped like the oracle
    Dr, Ir = scan_ref.search(Q, P, 100)
    assert I.shape == (3, 40) and (I.cpu().numpy() == Ir).all()
    with pytest.raises(RuntimeError):
        scan.search(Q, P, 10)  # CPU tensors: no fallback


@pytest.mark.parametrize("n_in,k", [(7, 5), (1000, 100), (1500, 1000), (3000, 1000), (8000, 1000), (12000, 2000)])
def test_topk_merge_register_sort_sizes(n_in, k):
    """cdr_topk_merge across every keys-per-thread variant of the block sort, with ties and empty slots."""
    from cocodr_b200 import scan
    g = torch.Generator().manual_seed(n_in)
    D = torch.randint(-50, 50, (5, n_in), generator=g).float() / 8.0  # many ties -> id tie-break matters
    I = torch.stack([torch.randperm(4 * n_in, generator=g)[:n_in] for _ in range(5)]).long()
    I[:, ::7] = -1  # empty slots
    Dm, Im = scan.merge_topk(D.cuda(), I.cuda(), k)
    for r in range(5):
        valid = I[r] >= 0
        d, i = D[r][valid], I[r][valid]
        order = sorted(range(len(d)), key=lambda t: (-d[t].item(), i[t].item()))[:k]
        exp_i = [i[t].item() for t in order] + [-1] * (k - len(order))
        assert Im[r].cpu().tolist() == exp_i


def test_fp16_corpus_keeps_the_fp32_ranking_on_trained_like_embeddings():
    """IndexFlatIP.add stores fp32 embeddings as fp16 (half the bytes of the HBM-bound scan).  Trained COCO-DR
    embeddings are a large shared component plus small per-item differences (norm ~14.7, dot products ~217, gaps ~0.3:
    SURVEY H4 / 8d), the case where that rounding could reorder results.  Against the reference semantics (fp32 inner
    products, evaluate_beir.py:220-224): the top-k SETS agree except for boundary ties, and two documents may only
    swap places where their fp32 scores differ by less than 1e-3 relative (SURVEY 8d's criterion)."""
    from cocodr_b200 import scan
    g = torch.Generator().manual_seed(3)
    H, n_docs, n_q, k = 768, 50_000, 64, 100
    base = torch.randn(H, generator=g) * 14.7 / H ** 0.5
    P = base[None, :] + 0.015 * torch.randn(n_docs, H, generator=g)
    Q = base[None, :] + 0.015 * torch.randn(n_q, H, generator=g)
    index = scan.IndexFlatIP(H)
    index.add(P.cuda())
    D, I = index.search(Q.cuda(), k)
    I = I.cpu().numpy() if torch.is_tensor(I) else np.asarray(I)
    S = (Q.double() @ P.double().T).numpy()  # reference scores (fp32 inputs, exact accumulation)
    assert 200 < np.median(S) < 235
    ref_order = np.argsort(-S, axis=1, kind="stable")[:, :k]
    recall = np.mean([len(set(I[q]) & set(ref_order[q])) / k for q in range(n_q)])
    assert recall > 0.97, recall
    worst = 0.0
    for q in range(n_q):
        s = S[q, I[q]]                       # reference scores in OUR order: must be non-increasing up to the tolerance
        inv = np.maximum.accumulate(s[::-1])[::-1]  # max of everything ranked at or below position i
        worst = max(worst, float(((inv - s) / np.abs(s)).max()))
        missing = set(ref_order[q]) - set(I[q])
        if missing:                          # a missed document must sit within the tolerance of the k-th score
            kth = S[q, I[q]].min()
            worst = max(worst, float(max((S[q, d] - kth) / abs(kth) for d in missing)))
    assert worst < 1e-3, worst
