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


def search_sharded(Q, P_local, k, doc_base, group=None):
    """Documents sharded over ranks: local top-k with global ids, all-gather of the [nq,k]
    candidate lists (12 B per entry), k-way merge on every rank.  No document embedding crosses NVLink."""
    import torch.distributed as dist
    D, I = search(Q, P_local, k, doc_base=doc_base)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return D, I
    W = dist.get_world_size(group)
    if D.shape[1] < k:  # ragged shards: pad with empty slots
        pad = k - D.shape[1]
        D = torch.cat([D, D.new_full((D.shape[0], pad), float("-inf"))], 1)
        I = torch.cat([I, I.new_full((I.shape[0], pad), -1)], 1)
    n_q = D.shape[0]
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
