"""ANN episode on the device (SURVEY f-2): embedding inference -> corpus scan -> negative mining (+ query clustering
for the iDRO group ids) without leaving HBM.

Mirrors the steps of the reference's ``generate_new_ann`` (ANCE/drivers/run_ann_data_gen.py:265-429):

    InferenceEmbeddingFromStreamDataLoader (:157-206)   -> ``encode``       embeddings stay on the GPU as fp16
    faiss.IndexFlatIP.add / search (:310-317, :390)      -> ``scan.search`` / ``scan.search_sharded``
    GenerateNegativePassaageID (:497-570)                -> ``mine_negatives``  (cdr_mine_negatives)
    faiss.Kmeans + IndexFlatL2.search(q, 1) (:340-351)   -> ``kmeans``          (tcgen05 scores + cdr_kmeans_*)

The reference moves every embedding through ``.cpu().numpy()``, pickles them per rank (utils/util.py:87-155) and
walks the top-k lists in Python; here only the final (query id, positive, negatives, group) table leaves the GPU.
No CPU path: CUDA tensors in, CUDA tensors out.
"""
import ctypes as C

import torch

from . import kernels as K
from . import scan
from ._lib import check, load, stream_ptr


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def plan_buckets(sorted_lens, token_budget=16384, max_seqs=1024, multiple=16, max_len=None):
    """Split sequences SORTED BY LENGTH (ascending) into consecutive groups that are each run at their own padded
    length: group = [a, b) with L = round_up(longest, multiple) and (b - a) * L <= token_budget.  Returns
    [(a, b, L)].  Pure host logic (tests/test_host_cpu.py)."""
    n = len(sorted_lens)
    out, a = [], 0
    while a < n:
        b = a
        while b < n:
            L = -(-max(int(sorted_lens[b]), 1) // multiple) * multiple
            if max_len is not None:
                L = min(L, max_len)
            if b > a and ((b - a + 1) * L > token_budget or b - a + 1 > max_seqs):
                break
            b += 1
        L = -(-max(int(sorted_lens[b - 1]), 1) // multiple) * multiple
        if max_len is not None:
            L = min(L, max_len)
        out.append((a, b, L))
        a = b
    return out


@torch.no_grad()
def encode(model, batches, is_query=True, out_dtype=torch.float16, trim=True, pool=16, token_budget=16384):
    """``batches`` yields the reference's inference tuples ``(input_ids, attention_mask, token_type_ids, idx)`` (or
    ``(input_ids, attention_mask, idx)``); returns ``(emb [N, H] out_dtype, ids int64 [N])`` on the model's device, rows
    in the order the batches delivered them.  Host tensors are copied asynchronously; nothing is copied back.

    ``trim`` (SURVEY 8d mode B, MS-MARCO-shaped lengths): the reference pads every sequence to max_seq_length
    (evaluate/data/msmarco_data.py GetProcessingFn) and runs the encoder on the padding (queries average 8 of 64
    positions, passages 75 of 128).  Padded positions never reach a [CLS] embedding -- keys are masked, every other
    operator is row-wise -- so ``pool`` batches at a time are sorted by real length on the device and re-batched into
    groups of similar length, each run at its own padded length (a multiple of 16) with ~``token_budget`` tokens per
    encoder pass (the GEMMs keep their full-wave shapes).  One host synchronisation per pool (the group boundaries);
    the embeddings are the same as the padded run's.  Masks that are not right-padded fall back to the padded run."""
    mod = model.module if hasattr(model, "module") else model
    dev = next(mod.parameters()).device
    was_training = mod.training
    mod.eval()
    fwd = mod.query_emb if is_query else mod.body_emb
    embs, ids = [], []

    def to_out(e):
        if out_dtype == torch.float16:
            h = torch.empty(e.shape, dtype=torch.float16, device=dev)
            K.cast_f32_f16(e.contiguous(), h)
            return h
        return e.to(out_dtype)

    def flush(pending):
        if not pending:
            return
        Lmax = max(t.shape[1] for t, _ in pending)
        if any(t.shape[1] != Lmax for t, _ in pending):  # ragged pools: pad to the widest batch
            pending = [(torch.nn.functional.pad(t, (0, Lmax - t.shape[1])), torch.nn.functional.pad(m, (0, Lmax - m.shape[1])))
                       for t, m in pending]
        inp = torch.cat([t for t, _ in pending], 0)
        mask = torch.cat([m for _, m in pending], 0)
        lens = mask.sum(1)
        prefix_ok = (mask[:, :-1] >= mask[:, 1:]).all() if Lmax > 1 else torch.ones((), dtype=torch.bool, device=dev)
        slen, order = torch.sort(lens)
        host = torch.cat([slen, prefix_ok.reshape(1).long()]).cpu()  # the pool's ONE host synchronisation
        if not bool(host[-1]):
            for t, m in pending:
                embs.append(to_out(fwd(input_ids=t, attention_mask=m)))
            return
        out = torch.empty(inp.shape[0], mod.config.hidden_size, dtype=out_dtype, device=dev)
        for a, b, L in plan_buckets(host[:-1].tolist(), token_budget=token_budget, max_len=Lmax):
            sel = order[a:b]
            e = fwd(input_ids=inp[sel, :L].contiguous(), attention_mask=mask[sel, :L].contiguous())
            out[sel] = to_out(e)
        embs.append(out)

    pending = []
    for batch in batches:
        inp, mask, idx = batch[0], batch[1], batch[-1]
        inp = inp.to(dev, non_blocking=True).long()
        mask = mask.to(dev, non_blocking=True).long()
        ids.append(idx.to(dev, non_blocking=True).long().reshape(-1))
        if not trim:
            embs.append(to_out(fwd(input_ids=inp, attention_mask=mask)))
            continue
        pending.append((inp, mask))
        if len(pending) >= pool:
            flush(pending)
            pending = []
    flush(pending)
    mod.train(was_training)
    return torch.cat(embs, 0), torch.cat(ids, 0)


def mine_negatives(I, doc_pid, pos_pid, n_neg, n_sel=None, order=None):
    """I [n_q, k] int64 document ROWS from the scan (-1 = empty), doc_pid [n_docs] int64 (``passage_embedding2id``),
    pos_pid [n_q] int64 (``training_query_positive_id[query_id]``).

    Returns ``(neg [n_q, n_neg] int64 padded with -1, counts int32 [n_q], rr float32 [n_q])``: the negatives the
    reference keeps (first ``n_neg`` distinct non-positive passage ids among the selected candidates) and the
    reciprocal rank of the positive over all k results.  ``n_sel`` defaults to ``n_neg + 1`` (the SelectTopK branch,
    ``top_ann_pid[:negative_sample + 1]``); ``order`` [n_q, n_sel] int32 walks the candidates in a caller-chosen
    order (the shuffled branch: pass a per-query permutation of ``range(k)``)."""
    if not (I.is_cuda and doc_pid.is_cuda and pos_pid.is_cuda):
        raise RuntimeError("cocodr_b200.mining needs CUDA tensors (no CPU fallback)")
    I, doc_pid, pos_pid = I.contiguous().long(), doc_pid.contiguous().long(), pos_pid.contiguous().long()
    n_q, k = I.shape
    if order is not None:
        order = order.contiguous().int()
        n_sel = order.shape[1]
    elif n_sel is None:
        n_sel = min(k, n_neg + 1)
    dev = I.device
    neg = torch.empty(n_q, n_neg, dtype=torch.int64, device=dev)
    cnt = torch.empty(n_q, dtype=torch.int32, device=dev)
    rr = torch.empty(n_q, dtype=torch.float32, device=dev)
    check(load().cdr_mine_negatives(_p(I), C.c_int32(n_q), C.c_int32(k), _p(doc_pid), C.c_int64(doc_pid.numel()),
                                    _p(pos_pid), _p(order), C.c_int32(n_sel), C.c_int32(n_neg), _p(rr), _p(neg), _p(cnt),
                                    stream_ptr()), "cdr_mine_negatives")
    K._count(1)
    return neg, cnt, rr


def kmeans(X, k, niter=20, init=None, seed=0, nredo=1, check_every=10):
    """Lloyd k-means of fp16 rows X [n, dim] (the train-query embeddings) -> (centroids fp32 [k, dim], assign int32 [n]).

    Mirrors ``faiss.Kmeans(dim, k, nredo=5, niter=500).train(q)`` + ``IndexFlatL2(centroids).search(q, 1)`` of
    ANCE/drivers/run_ann_data_gen.py:340-351: ``nredo`` runs from different random initialisations (k distinct rows drawn
    with ``seed + redo``), the run with the smallest sum of squared distances wins; ``niter`` Lloyd iterations each,
    stopping early once the assignment no longer changes (checked every ``check_every`` iterations: one host
    synchronisation per check).  Assignment = nearest centroid in L2 (``argmax_c x.c - |c|^2/2``; the x.c matrix comes
    from the tcgen05 GEMM against fp16 centroids), update = mean of the members.  Empty clusters keep their centroid
    (faiss re-seeds them by splitting a large cluster); faiss is not importable here, so this step is NOT pinned against
    the reference's outputs.  ``init`` [k, dim] fixes the starting centroids (then nredo must be 1)."""
    if not X.is_cuda or X.dtype != torch.float16:
        raise RuntimeError("cocodr_b200.mining.kmeans needs a CUDA fp16 matrix (no CPU fallback)")
    if init is not None and nredo != 1:
        raise ValueError("kmeans: an explicit init and nredo > 1 exclude each other")
    X = X.contiguous()
    n, dim = X.shape
    dev = X.device
    kp = (k + 63) // 64 * 64
    c16 = torch.zeros(kp, dim, dtype=torch.float16, device=dev)
    scores = torch.empty(n, kp, dtype=torch.float32, device=dev)
    sums = torch.empty(k, dim, dtype=torch.float32, device=dev)
    counts = torch.empty(k, dtype=torch.float32, device=dev)
    lib = load()
    best = None
    for redo in range(max(1, nredo)):
        if init is None:
            g = torch.Generator(device="cpu").manual_seed(seed + redo)
            start = X[torch.randperm(n, generator=g)[:k].to(dev)].float()
        else:
            start = init
        cent = start.to(dev).float().contiguous().clone()
        assign = torch.empty(n, dtype=torch.int32, device=dev)
        prev = None
        for it in range(niter + 1):
            K.cast_f32_f16(cent, c16[:k])
            K.gemm(X, c16, scores, M=n, N=kp, K=dim, epilogue=K.EPI_F32_STORE)
            half_sq = 0.5 * (c16[:k].float() ** 2).sum(1)
            check(lib.cdr_kmeans_assign(_p(scores), C.c_int64(kp), _p(half_sq), C.c_int64(n), C.c_int32(k), _p(assign),
                                        stream_ptr()), "cdr_kmeans_assign")
            K._count(1)
            if it == niter:
                break
            if check_every > 0 and it % check_every == check_every - 1:
                if prev is not None and torch.equal(prev, assign):  # converged: later iterations would change nothing
                    break
                prev = assign.clone()
            sums.zero_()
            counts.zero_()
            check(lib.cdr_kmeans_accumulate(_p(X), _p(assign), C.c_int64(n), C.c_int32(dim), C.c_int32(k), _p(sums),
                                            _p(counts), stream_ptr()), "cdr_kmeans_accumulate")
            K._count(1)
            nz = counts > 0
            cent = torch.where(nz[:, None], sums / counts.clamp(min=1.0)[:, None], cent)
        if nredo <= 1:
            return cent, assign
        # sum of squared distances up to the constant sum |x|^2:  -2 * sum_i (x_i . c_a - |c_a|^2 / 2)
        a64 = assign.long()
        obj = -2.0 * (scores.gather(1, a64[:, None]).squeeze(1) - half_sq[a64]).double().sum()
        if best is None or float(obj) < best[0]:
            best = (float(obj), cent, assign)
    return best[1], best[2]


def ann_episode(query_emb, query_ids, passage_emb, passage_ids, positive_pid, top_k, n_neg, shuffle_seed=None,
                n_groups=0, kmeans_iters=500, kmeans_redo=5):
    """One ANN data-generation episode on resident embeddings: scan -> (MRR, negatives) [-> group ids].

    positive_pid [n_q] int64: positive passage id of every query row.  Returns a dict of CUDA tensors:
    ``neg`` [n_q, n_neg], ``neg_count``, ``rr`` (reciprocal ranks; ``rr.mean()`` is the reference's ANN MRR when the
    list is cut at 10), ``I`` (document rows), ``D`` (scores) and, if ``n_groups`` > 0, ``group``/``centroids``."""
    D, I = scan.search(query_emb, passage_emb, top_k)
    order = None
    if shuffle_seed is not None:  # the reference's non-SelectTopK branch: a fresh permutation of the k candidates per query
        g = torch.Generator(device=I.device).manual_seed(shuffle_seed)
        order = torch.rand(I.shape, generator=g, device=I.device).argsort(1).int()
    neg, cnt, rr = mine_negatives(I, passage_ids, positive_pid, n_neg, order=order)
    out = {"neg": neg, "neg_count": cnt, "rr": rr, "I": I, "D": D}
    if n_groups > 0:
        out["centroids"], out["group"] = kmeans(query_emb, n_groups, niter=kmeans_iters, nredo=kmeans_redo)
    return out
