"""GPU: ANN negative mining / k-means / episode glue (cocodr_b200.mining) against the oracle restatement of the
reference's GenerateNegativePassaageID and a float64 Lloyd step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_q,k,n_neg,shuffled", [(37, 200, 20, False), (64, 100, 5, True), (5, 33, 40, True), (3, 8, 7, False)])
def test_mine_negatives_matches_reference_walk(n_q, k, n_neg, shuffled):
    from cocodr_b200 import mining
    from oracle import mining_ref
    rng = np.random.RandomState(n_q + k)
    n_docs = 500
    doc_pid = rng.permutation(5000)[:n_docs].astype(np.int64)
    doc_pid[::7] = doc_pid[1]  # several document rows share a passage id (multi-chunk passages): repeats must be skipped
    I = rng.randint(0, n_docs, size=(n_q, k)).astype(np.int64)
    I[0, 5:] = -1  # short list
    pos = doc_pid[rng.randint(0, n_docs, size=n_q)].copy()
    pos[1] = 999_999  # positive not retrieved
    I[2, 0] = int(np.where(doc_pid == pos[2])[0][0])  # positive at rank 1
    order = None
    if shuffled:
        order = np.stack([rng.permutation(k) for _ in range(n_q)]).astype(np.int32)
    neg, cnt, rr = mining.mine_negatives(torch.from_numpy(I).cuda(), torch.from_numpy(doc_pid).cuda(),
                                         torch.from_numpy(pos).cuda(), n_neg,
                                         order=None if order is None else torch.from_numpy(order).cuda())
    neg, cnt, rr = neg.cpu().numpy(), cnt.cpu().numpy(), rr.cpu().numpy()
    for q in range(n_q):
        ref_neg, ref_rr = mining_ref.generate_negatives(list(I[q]), doc_pid, pos[q], n_neg,
                                                        order=None if order is None else list(order[q]))
        assert cnt[q] == len(ref_neg)
        assert list(neg[q, :cnt[q]]) == ref_neg and (neg[q, cnt[q]:] == -1).all()
        assert abs(rr[q] - ref_rr) < 1e-7


def test_kmeans_matches_lloyd_steps():
    from cocodr_b200 import mining
    from oracle import mining_ref
    rng = np.random.RandomState(0)
    k, dim, n = 5, 64, 3000
    centers = rng.randn(k, dim) * 3.0
    X = (centers[rng.randint(0, k, size=n)] + 0.3 * rng.randn(n, dim)).astype(np.float16)
    init = (centers + 0.5 * rng.randn(k, dim)).astype(np.float32)  # well separated: assignments are unambiguous
    cent, assign = mining.kmeans(torch.from_numpy(X).cuda(), k, niter=4, init=torch.from_numpy(init))
    c = init.astype(np.float64)
    for _ in range(4):
        c, _ = mining_ref.kmeans_step(X, c)
    _, ref_assign = mining_ref.kmeans_step(X, c)
    assert (assign.cpu().numpy() == ref_assign).mean() > 0.999  # fp16 centroids in the score GEMM: ties may flip
    assert np.abs(cent.cpu().numpy() - c).max() < 2e-2


def test_episode_on_resident_embeddings():
    """scan -> negatives -> groups, all on the device: positives planted as each query's nearest passage."""
    from cocodr_b200 import mining
    g = torch.Generator().manual_seed(3)
    n_docs, n_q, dim = 20000, 96, 128
    P = torch.randn(n_docs, dim, generator=g).half().cuda()
    pid = (torch.randperm(10 * n_docs, generator=g)[:n_docs]).cuda()
    rows = torch.randperm(n_docs, generator=g)[:n_q].cuda()
    Q = (P[rows].float() * 1.5).half()  # the planted passage has the largest inner product
    out = mining.ann_episode(Q, torch.arange(n_q).cuda(), P, pid, pid[rows], top_k=50, n_neg=8, n_groups=4,
                             kmeans_iters=3)
    assert (out["I"][:, 0] == rows).all() and torch.allclose(out["rr"], torch.ones(n_q, device="cuda"))
    assert (out["neg_count"] == 8).all()
    assert not (out["neg"] == pid[rows][:, None]).any()
    assert out["group"].shape == (n_q,) and int(out["group"].max()) < 4
    shuf = mining.ann_episode(Q, torch.arange(n_q).cuda(), P, pid, pid[rows], top_k=50, n_neg=8, shuffle_seed=1)
    cand = pid[out["I"]]
    assert all(set(shuf["neg"][q].tolist()) <= set(cand[q].tolist()) for q in range(n_q))


def test_mine_negatives_matches_reference_fixture(golden_dir):
    """cdr_mine_negatives vs the outputs of the UNMODIFIED reference function (tests/golden/mining_tiny.npz)."""
    import os
    from cocodr_b200 import mining
    g = np.load(os.path.join(golden_dir, "mining_tiny.npz"))
    I, doc_pid, pos = (torch.from_numpy(g[k]).cuda() for k in ("I", "doc_pid", "pos"))
    n_neg = int(g["n_neg"])
    for tag in ("topk", "shuf"):
        order = None if tag == "topk" else torch.from_numpy(g["shuf.order"]).cuda()
        neg, cnt, rr = mining.mine_negatives(I, doc_pid, pos, n_neg, order=order)
        assert (neg.cpu().numpy() == g[f"{tag}.neg"]).all()
        assert (cnt.cpu().numpy() == (g[f"{tag}.neg"] >= 0).sum(1)).all()
        np.testing.assert_allclose(rr.cpu().numpy(), g[f"{tag}.rr"], rtol=1e-6)


def test_encode_inference_loop_matches_oracle_embeddings():
    """mining.encode == the reference's InferenceEmbeddingFromStreamDataLoader loop
    (evaluate/drivers/run_ann_data_gen.py:152-206): batches in GetProcessingFn's format (int32 ids, bool mask, uint8
    type ids, int64 idx; host tensors are copied inside the loop), eval() mode even when the model is in train(),
    ids and embeddings returned in order; embeddings vs the fp32 oracle at 1e-2, fp16 and fp32 outputs."""
    from cocodr_b200 import mining
    from oracle import bert_ref
    import test_model_gpu as T
    m = T.build(T.TINY).train()  # train(): encode must switch dropout-free eval() on and restore the mode
    cfg = T.TINY
    batches, ref_ids, ref_embs = [], [], []
    st = bert_ref.synth_state(cfg, 0)
    for b, (n, L) in enumerate([(5, 32), (7, 32), (3, 48)]):  # ragged batch sizes and two sequence lengths
        ids, mask = bert_ref.synth_batch(n, L, cfg["vocab"], 40 + b)
        idx = torch.arange(100 * b, 100 * b + n)
        batches.append((ids.int(), mask.bool(), torch.zeros(n, L, dtype=torch.uint8), idx))
        ref_ids.append(idx)
        with torch.no_grad():
            ref_embs.append(bert_ref.cls_embedding(st, ids, mask, cfg))
    ref = torch.cat(ref_embs).numpy()
    for dtype in (torch.float16, torch.float32):
        emb, ids = mining.encode(m, [tuple(t.pin_memory() for t in bt) for bt in batches], is_query=False, out_dtype=dtype)
        assert emb.is_cuda and emb.dtype == dtype and emb.shape == ref.shape
        assert torch.equal(ids.cpu(), torch.cat(ref_ids))
        assert np.abs(emb.float().cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-2
    assert m.training
    q_emb, _ = mining.encode(m, batches[:1], is_query=True)  # 3-tuples / device tensors are accepted as well
    q2, _ = mining.encode(m, [(batches[0][0].cuda(), batches[0][1].cuda(), batches[0][3])], is_query=True)
    assert torch.equal(q_emb, q2)


def test_encode_trimmed_equals_padded_run():
    """SURVEY 8d mode B: re-batching by real length (mining.encode trim=True) must not change an embedding -- padded
    positions reach a [CLS] row only as masked keys.  Several groups per pool (small token budget), fp32 outputs
    compared with the padded run of the same sequences; a mask that is not right-padded falls back to the padded run."""
    from cocodr_b200 import mining
    from oracle import bert_ref
    import test_model_gpu as T
    m = T.build(T.TINY).eval()
    cfg = T.TINY
    g = torch.Generator().manual_seed(5)
    batches = []
    for b in range(4):
        n, L = 24, 64
        lens = torch.randint(1, L + 1, (n,), generator=g)
        lens[0] = L
        ids = torch.randint(1000, cfg["vocab"], (n, L), generator=g, dtype=torch.int32)
        mask = torch.arange(L)[None, :] < lens[:, None]
        ids = ids * mask
        ids[:, 0] = 101
        batches.append((ids.cuda(), mask.cuda(), torch.arange(b * n, (b + 1) * n)))
    pad, ids_pad = mining.encode(m, batches, is_query=False, out_dtype=torch.float32, trim=False)
    for budget in (16384, 512, 64):  # one group, several groups, one sequence per group
        tr, ids_tr = mining.encode(m, batches, is_query=False, out_dtype=torch.float32, trim=True, token_budget=budget)
        assert torch.equal(ids_tr, ids_pad)
        err = (tr - pad).abs().max().item() / pad.abs().max().item()
        assert err < 2e-3, (budget, err)  # (different GEMM tile shapes: summation order, not values)
    holes = [(i, mk.clone(), idx) for i, mk, idx in batches]
    holes[1][1][3, 1] = False  # a hole in the mask: not a right-padded batch
    a, _ = mining.encode(m, holes, is_query=False, out_dtype=torch.float32, trim=True)
    b_, _ = mining.encode(m, holes, is_query=False, out_dtype=torch.float32, trim=False)
    assert torch.equal(a, b_)


def test_kmeans_restarts_keep_the_best_objective():
    """faiss.Kmeans(nredo=5, niter=500) semantics (run_ann_data_gen.py:340-351): several random initialisations, the
    smallest sum of squared distances wins; iterations stop once the assignment is stable."""
    from cocodr_b200 import mining
    g = torch.Generator().manual_seed(11)
    k, dim = 8, 64
    centers = torch.randn(k, dim, generator=g) * 3.0
    X = (centers[torch.randint(0, k, (4000,), generator=g)] + 0.3 * torch.randn(4000, dim, generator=g)).half().cuda()

    def objective(cent, assign):
        return ((X.float() - cent[assign.long()]) ** 2).sum().item()

    singles = [objective(*mining.kmeans(X, k, niter=100, seed=s)) for s in range(3)]
    cent, assign = mining.kmeans(X, k, niter=100, seed=0, nredo=3)
    best = objective(cent, assign)
    assert best <= min(singles) * (1 + 1e-4), (best, singles)
    # the winner is a fixed point: one more assignment against its centroids changes nothing
    d2 = ((X.float()[:, None, :] - cent[None, :, :]) ** 2).sum(-1)
    assert (d2.argmin(1).int() == assign).float().mean().item() > 0.999
    with pytest.raises(ValueError):
        mining.kmeans(X, k, init=cent, nredo=2)
