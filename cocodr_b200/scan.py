"""Corpus scan: host-side mirror of the reference's faiss usage for brute-force retrieval.

The reference does (evaluate/evaluation/evaluate_beir.py:220-224, ANCE/drivers/run_ann_data_gen.py:310-317,390)

    cpu_index = faiss.IndexFlatIP(dim); cpu_index.add(passage_embedding)
    D, I = cpu_index.search(query_embedding, topN)

``IndexFlatIP`` below has that surface (add / search / ntotal / d / reset) over HBM-resident fp16 document
embeddings; ``search`` / ``merge_topk`` / ``search_sharded`` are the tensor-level entry points.  All compute
goes through the C ABI (cdr_scan_topk, cdr_topk_merge); there is no CPU path.  Result order is
(score desc, doc index asc) -- faiss leaves ties unspecified.
"""
import numpy as np
import torch

from . import kernels as K


def _as_f16_cuda(x, device=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        x = x.pin_memory().to(device, non_blocking=True)
    if not x.is_cuda:
        raise RuntimeError("cocodr_b200.scan needs CUDA tensors (no CPU fallback)")
    if x.dtype == torch.float32:
        out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        K.cast_f32_f16(x.contiguous(), out)
        return out
    if x.dtype != torch.float16:
        raise RuntimeError(f"embeddings must be float16 or float32, got {x.dtype}")
    return x


def _scan_once(Q, P, k, doc_base):
    n_q = Q.shape[0]
    D = torch.empty(n_q, k, dtype=torch.float32, device=Q.device)
    I = torch.empty(n_q, k, dtype=torch.int64, device=Q.device)
    ws = torch.empty(K.scan_workspace_bytes(P.shape[0], n_q, k, P.shape[1]), dtype=torch.uint8, device=Q.device)
    status = torch.zeros(1, dtype=torch.int32, device=Q.device)
    K.scan_topk(P, Q, D, I, ws, status, k=k, doc_base=doc_base)
    return D, I, status


def merge_topk(D, I, k):
    """Top-k by (score desc, id asc) of ``n_in`` candidates per query (ids < 0 are empty slots)."""
    D, I = D.contiguous(), I.contiguous()
    outD = torch.empty(D.shape[0], k, dtype=torch.float32, device=D.device)
    outI = torch.empty(D.shape[0], k, dtype=torch.int64, device=D.device)
    K.topk_merge(D, I, outD, outI, k=k)
    return outD, outI


def search_async(Q, P, k, doc_base=0):
    """One sampled-threshold scan without the host-side status check: returns (D, I, status) where ``status`` is a
    device int32[1]; the results are valid iff it is 0 (``check_status`` / ``search`` handle the other case).
    Lets a caller keep several searches in flight and pay one synchronisation for all of them."""
    if not (torch.is_tensor(Q) and Q.is_cuda and torch.is_tensor(P) and P.is_cuda):
        raise RuntimeError("cocodr_b200.scan.search needs CUDA tensors (no CPU fallback)")
    Q, P = _as_f16_cuda(Q).contiguous(), _as_f16_cuda(P)
    if P.stride(1) != 1:
        P = P.contiguous()
    return _scan_once(Q, P, min(k, P.shape[0]), doc_base)


def check_status(results, Q, P, k, doc_base=0):
    """Resolve a list of ``search_async`` results: one host sync; any search whose status is non-zero is redone on
    the guaranteed (exhaustive, chunked) path."""
    flags = torch.cat([r[2] for r in results]).cpu().tolist()
    out = []
    for (D, I, _), bad in zip(results, flags):
        out.append(search(Q, P, k, doc_base=doc_base, force_exhaustive=True) if bad else (D, I))
    return out


def search(Q, P, k, doc_base=0, force_exhaustive=False):
    """D [nq,k'] fp32, I [nq,k'] int64 (k' = min(k, n_docs)) for fp16 (or fp32 -> cast) CUDA embeddings."""
    if not (torch.is_tensor(Q) and Q.is_cuda and torch.is_tensor(P) and P.is_cuda):
        raise RuntimeError("cocodr_b200.scan.search needs CUDA tensors (no CPU fallback)")
    Q, P = _as_f16_cuda(Q).contiguous(), _as_f16_cuda(P)
    if P.stride(1) != 1:
        P = P.contiguous()
    n_docs = P.shape[0]
    k = min(k, n_docs)
    if not force_exhaustive:
        D, I, status = _scan_once(Q, P, k, doc_base)
        if int(status.item()) == 0:
            return D, I
    # guaranteed path: every document of each chunk is admitted; running merge of the chunk winners
    chunk = K.scan_exhaustive_docs(k)
    runD = runI = None
    for lo in range(0, n_docs, chunk):
        part = P[lo:lo + chunk]
        D, I, status = _scan_once(Q, part, min(k, part.shape[0]), doc_base + lo)
        if int(status.item()) != 0:
            raise RuntimeError("cocodr_b200.scan: exhaustive chunk reported a candidate-buffer failure")
        if runD is None:
            runD, runI = D, I
        else:
            runD, runI = merge_topk(torch.cat([runD, D], 1), torch.cat([runI, I], 1), min(k, runD.shape[1] + D.shape[1]))
    return runD, runI


def shard_list_len(k, world):
    """Entries each shard contributes to a sharded top-k: a shard of a randomly split corpus holds ~k / W of the global
    top k (binomial), so 1.5 k / W + 64 covers it with a wide margin; the merge VERIFIES that it sufficed."""
    if world * k <= 2048:
        return k
    return min(k, (int(1.5 * k / world) + 64 + 63) // 64 * 64)


def _shard_keys(Q, P_local, ks, doc_base):
    """This shard's top-ks as packed keys (+ status words); everything stays enqueued on the stream."""
    k_eff = min(ks, P_local.shape[0])
    D, I, status = search_async(Q, P_local, k_eff, doc_base=doc_base)
    keys = torch.empty(Q.shape[0] * ks + 2, dtype=torch.int64, device=Q.device)
    K.topk_pack(D, I, status, keys, ks=ks, all_returned=(k_eff == P_local.shape[0]))
    return keys


def _merge_gathered(allk, world, n_q, ks, k):
    outD = torch.empty(n_q, k, dtype=torch.float32, device=allk.device)
    outI = torch.empty(n_q, k, dtype=torch.int64, device=allk.device)
    flag = torch.zeros(1, dtype=torch.int32, device=allk.device)
    K.topk_merge_keys(allk, outD, outI, flag, world=world, n_q=n_q, ks=ks, k=k)
    return outD, outI, flag


def search_sharded(Q, P_local, k, doc_base, group=None):
    """Documents sharded over ranks: every rank scans its shard for a SHORT list (``shard_list_len``), the
    lists travel as packed 8-byte keys in ONE all-gather, and each rank merges them where they land.  The merge kernel
    proves exactness (no truncated shard can hold a better document than the k-th merged one) and the host reads one
    flag per search; if the proof fails -- a corpus clustered by shard, or a shard scan that overflowed -- the search is
    repeated with full-length lists through the guaranteed path.  No document embedding crosses NVLink.  Results are
    bit-identical to a single-device search: same total order (score desc, id asc)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return search(Q, P_local, k, doc_base=doc_base)
    W = dist.get_world_size(group)
    n_q = Q.shape[0]
    for ks in dict.fromkeys((shard_list_len(k, W), k)):
        if ks == k and W * ks > 16384:
            break
        mine = _shard_keys(Q, P_local, ks, doc_base)
        allk = torch.empty(W * mine.numel(), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(allk, mine, group=group)
        D, I, flag = _merge_gathered(allk, W, n_q, ks, min(k, W * ks))
        if int(flag.item()) == 0:  # (identical on every rank: all ranks merge the same gathered keys)
            return D, I
    # guaranteed path: exact per-shard top-k (chunked exhaustive scan if need be), gathered and merged
    D, I = search(Q, P_local, k, doc_base=doc_base)
    if D.shape[1] < k:  # ragged shards: pad with empty slots
        pad = k - D.shape[1]
        D = torch.cat([D, D.new_full((D.shape[0], pad), float("-inf"))], 1)
        I = torch.cat([I, I.new_full((I.shape[0], pad), -1)], 1)
    allD = torch.empty(W * n_q, k, dtype=D.dtype, device=D.device)  # rank-major blocks of [n_q, k]
    allI = torch.empty(W * n_q, k, dtype=I.dtype, device=I.device)
    dist.all_gather_into_tensor(allD, D.contiguous(), group=group)
    dist.all_gather_into_tensor(allI, I.contiguous(), group=group)
    return merge_topk(allD.view(W, n_q, k).permute(1, 0, 2).reshape(n_q, -1),
                      allI.view(W, n_q, k).permute(1, 0, 2).reshape(n_q, -1), k)


class IndexFlatIP:
    """faiss.IndexFlatIP look-alike (exact inner product) over fp16 embeddings resident in HBM."""

    def __init__(self, d, device=None):
        self.d = int(d)
        self.device = device
        self._parts = []
        self._docs = None

    @property
    def ntotal(self):
        return sum(p.shape[0] for p in self._parts)

    def reset(self):
        self._parts, self._docs = [], None

    def add(self, x):
        x = _as_f16_cuda(x, self.device)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"IndexFlatIP.add: expected [n, {self.d}], got {tuple(x.shape)}")
        self._parts.append(x.contiguous())
        self._docs = None

    def search(self, x, k):
        if self.ntotal == 0:
            raise RuntimeError("IndexFlatIP.search on an empty index")
        if self._docs is None:
            self._docs = self._parts[0] if len(self._parts) == 1 else torch.cat(self._parts, 0)
            self._parts = [self._docs]
        as_numpy = isinstance(x, np.ndarray)
        q = _as_f16_cuda(x, self._docs.device)
        D, I = search(q, self._docs, k)
        if D.shape[1] < k:  # faiss pads missing neighbours with -inf / -1
            pad = k - D.shape[1]
            D = torch.cat([D, D.new_full((D.shape[0], pad), float("-inf"))], 1)
            I = torch.cat([I, I.new_full((I.shape[0], pad), -1)], 1)
        if as_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I
