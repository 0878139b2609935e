"""CPU restatement of the brute-force corpus scan (faiss IndexFlatIP add/search).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows evaluate/evaluation/evaluate_beir.py:220-224 and
ANCE/drivers/run_ann_data_gen.py:310-317 (``IndexFlatIP(dim).add(P); D, I =
search(Q, k)``): exact inner products, top-k by descending score.  faiss-cpu
1.6.4 (warmup/commands/install.sh:4) is a third-party dependency absent from
/root/reference and from this image; its published algorithm for IndexFlatIP
is a blocked SGEMM followed by a per-query heap.  faiss leaves the order of
equal scores unspecified; the contract here (and of the CUDA path) strengthens
it to (score desc, doc index asc) so rank lists are comparable bit-for-bit.
"""
import numpy as np
import torch


def synth_corpus(n_docs, n_q, dim, seed=7, kind="exact"):
    """SURVEY.md §8d corpora.  'exact': entries in {-4..4}/8 (every fp32 partial sum is
    exact => scores independent of summation order, many ties); 'gauss': N(0,1) in fp16."""
    g = torch.Generator().manual_seed(seed)
    if kind == "exact":
        P = (torch.randint(-4, 5, (n_docs, dim), generator=g).float() / 8).half()
        Q = (torch.randint(-4, 5, (n_q, dim), generator=g).float() / 8).half()
    else:
        P = torch.randn(n_docs, dim, generator=g).half()
        Q = torch.randn(n_q, dim, generator=g).half()
    return Q, P


def topk_desc_stable(scores: np.ndarray, k: int):
    """Top-k per row ordered by (score desc, index asc)."""
    k = min(k, scores.shape[1])
    # stable argsort of the negated scores keeps ascending index among ties
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, axis=1), order.astype(np.int64)


def search(Q, P, k, chunk=131072):
    """D float32 [nq,k], I int64 [nq,k] -- exact IP in fp32, chunked over docs and merged."""
    Qf = torch.as_tensor(Q).float()
    Pt = torch.as_tensor(P)
    nq = Qf.shape[0]
    best_s = np.empty((nq, 0), np.float32)
    best_i = np.empty((nq, 0), np.int64)
    for lo in range(0, Pt.shape[0], chunk):
        s = (Qf @ Pt[lo:lo + chunk].float().t()).numpy()
        cs, ci = topk_desc_stable(s, k)
        ci = ci + lo
        ms = np.concatenate([best_s, cs], axis=1)
        mi = np.concatenate([best_i, ci], axis=1)
        # merge: candidates are already (score desc, idx asc) within each part and part
        # indices are increasing, so a stable sort on -score preserves idx-asc ties
        o = np.argsort(-ms, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(ms, o, axis=1)
        best_i = np.take_along_axis(mi, o, axis=1)
    return best_s, best_i


def search_fast(Q, P, k, chunk=200000):
    """Same result set as ``search`` but via torch.topk per chunk (used as the timed CPU
    baseline: SGEMM + heap like faiss).  Final order fixed up to (score desc, idx asc)."""
    Qf = torch.as_tensor(Q).float()
    Pt = torch.as_tensor(P)
    parts_s, parts_i = [], []
    for lo in range(0, Pt.shape[0], chunk):
        s = Qf @ Pt[lo:lo + chunk].float().t()
        kk = min(k, s.shape[1])
        v, i = torch.topk(s, kk, dim=1)
        parts_s.append(v)
        parts_i.append(i + lo)
    s = torch.cat(parts_s, 1)
    i = torch.cat(parts_i, 1)
    kk = min(k, s.shape[1])
    v, j = torch.topk(s, kk, dim=1)
    return v.numpy(), torch.gather(i, 1, j).numpy()
