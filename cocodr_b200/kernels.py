"""Thin tensor-level wrappers over the C ABI (one python function per cdr_* entry point).

These only translate torch tensors into raw pointers / sizes and raise on error; all arithmetic is in
the CUDA library.  Device memory and streams are torch's (plumbing, not product).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BIAS_DROP_RESIDUAL, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_F32_ATOMIC,  # noqa: F401
                   EPI_F32_STORE, EPI_STORE_F16, check, ptr, stream_ptr)


# ---- instrumentation used by bench.py: kernel-launch counter and optional per-GEMM CUDA-event timing
launches = 0          # number of CUDA kernels launched by this library so far (our own kernels only)
gemm_events = None    # when a list: gemm() appends (flops, start_event, end_event)


op_events = None      # when a list: every cdr_* call appends (name, start_event, end_event)


def _count(n):
    global launches
    launches += n


def _run(name, call):
    """Issue one C-ABI call (optionally bracketed by CUDA events on the launching stream) and raise on error."""
    if op_events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = call()
        e1.record()
        op_events.append((name, e0, e1))
    else:
        rc = call()
    check(rc, name)


def drop_args(state, site, p, row_mul=1, keep_bits=None):
    """cdr_dropout descriptor: ``state`` = int64 device tensor {seed, offset}; keep iff u16 >= round(p * 65536).
    ``keep_bits`` (uint8 [rows * cols / 8]): the dropout GEMM epilogue stores its masks there and cdr_ln_bwd_drop reads
    them back instead of regenerating them."""
    assert state.is_cuda and state.dtype == torch.int64 and state.numel() == 2 and state.is_contiguous()
    assert 0.0 < p < 1.0
    d = _lib.Dropout()
    d.state, d.site, d.threshold, d.scale, d.row_mul = state.data_ptr(), site, int(round(p * 65536.0)), 1.0 / (1.0 - p), row_mul
    assert 0 < d.threshold < 65536
    if keep_bits is not None:
        assert keep_bits.is_cuda and keep_bits.dtype == torch.uint8 and keep_bits.is_contiguous()
        d.keep_bits = keep_bits.data_ptr()
    d._keepalive = (state, keep_bits)  # the descriptor only carries raw pointers: the tensors must outlive every launch
    return d


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cocodr_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, b, out, *, M, N, K, a_major=0, b_major=0, epilogue=EPI_STORE_F16, bias=None, aux=None, out2=None,
         alpha=1.0, split_k=1, lda=None, ldb=None, ldo=None, ldaux=None, dbg_lbo=0, dbg_sbo=0, colsum=None,
         colsum_scale=1.0, drop=None):
    """out[M,N] = alpha * A[M,K] @ B[N,K]^T with a fused epilogue (see include/cocodr_b200.h).
    ``drop`` (a drop_args descriptor) goes with EPI_BIAS_DROP_RESIDUAL."""
    _need_cuda(a, b, out)
    assert a.dtype == torch.float16 and b.dtype == torch.float16
    g = _lib.GemmArgs()
    g.a, g.b, g.out, g.out2 = a.data_ptr(), b.data_ptr(), out.data_ptr(), (out2.data_ptr() if out2 is not None else 0)
    g.bias = bias.data_ptr() if bias is not None else 0
    g.aux = aux.data_ptr() if aux is not None else 0
    g.M, g.N, g.K = M, N, K
    g.lda = lda if lda is not None else a.stride(0)
    g.ldb = ldb if ldb is not None else b.stride(0)
    g.ldo = ldo if ldo is not None else out.stride(0)
    g.ldaux = ldaux if ldaux is not None else (aux.stride(0) if aux is not None else 0)
    g.a_major, g.b_major, g.epilogue, g.split_k = a_major, b_major, epilogue, split_k
    g.alpha = alpha
    g.dbg_lbo, g.dbg_sbo = dbg_lbo, dbg_sbo
    g.colsum = colsum.data_ptr() if colsum is not None else 0
    g.colsum_scale = colsum_scale
    if drop is not None:
        g.drop = drop
    if bias is not None:
        assert bias.dtype == torch.float32
    if gemm_events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.load().cdr_gemm(C.byref(g), stream_ptr()), "cdr_gemm")
        e1.record()
        gemm_events.append((2.0 * M * N * K, e0, e1))
    else:
        _run("cdr_gemm[epi%d]" % epilogue, lambda: _lib.load().cdr_gemm(C.byref(g), stream_ptr()))
    _count(1)
    return out


def gemm_segments(a, b, out, *, M, N, row_begin, row_count, out_offset, alpha=1.0, epilogue=EPI_F32_ATOMIC,
                  split_k=0, ldo=None):
    """cdr_gemm_segments: for every i, out.view(-1)[out_offset[i]:][M, N] (+)= alpha * a[rows_i]^T @ b[rows_i] with
    rows_i = [row_begin[i], row_begin[i] + row_count[i]) -- both operands MN-major ([rows, M] / [rows, N] fp16),
    fp32 output; one C call enqueues all segments (iDRO per-group wgrad)."""
    _need_cuda(a, b, out)
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and out.dtype == torch.float32
    n = len(row_begin)
    assert len(row_count) == n and len(out_offset) == n
    g = _lib.GemmArgs()
    g.a, g.b, g.out, g.out2 = a.data_ptr(), b.data_ptr(), out.data_ptr(), 0
    g.bias, g.aux = 0, 0
    g.M, g.N, g.K = M, N, 0
    g.lda, g.ldb, g.ldo, g.ldaux = a.stride(0), b.stride(0), (ldo if ldo is not None else N), 0
    g.a_major, g.b_major, g.epilogue, g.split_k = 1, 1, epilogue, split_k
    g.alpha = alpha
    g.dbg_lbo, g.dbg_sbo = 0, 0
    g.colsum, g.colsum_scale = 0, 1.0
    arr = C.c_int64 * n
    _run("cdr_gemm_segments", lambda: _lib.load().cdr_gemm_segments(C.byref(g), C.c_int32(n), arr(*row_begin),
                                                                   arr(*row_count), arr(*out_offset), stream_ptr()))
    _count(sum(1 for c in row_count if c > 0))
    return out


def gemm_grouped(a, b, out, *, M, N, seg_kb, n_groups, out_group_stride, out_offset=0, alpha=1.0, ldo=None):
    """cdr_gemm_grouped: ONE launch; group g adds alpha * a[rows_g]^T @ b[rows_g] to the [M, N] block at
    out.view(-1)[out_offset + g * out_group_stride:], rows_g = [64 * seg_kb[g], 64 * seg_kb[g + 1]) -- seg_kb is an int32 DEVICE tensor of n_groups + 1 ascending
    entries, so nothing about the group sizes crosses to the host; rows padding a group to 64 must be zero."""
    _need_cuda(a, b, out, seg_kb)
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and out.dtype == torch.float32
    assert seg_kb.dtype == torch.int32 and seg_kb.numel() == n_groups + 1 and seg_kb.is_contiguous()
    assert a.shape[0] == b.shape[0]
    g = _lib.GemmArgs()
    g.a, g.b, g.out, g.out2 = a.data_ptr(), b.data_ptr(), out.data_ptr() + 4 * out_offset, 0
    g.bias, g.aux = 0, 0
    g.M, g.N, g.K = M, N, a.shape[0]
    g.lda, g.ldb, g.ldo, g.ldaux = a.stride(0), b.stride(0), (ldo if ldo is not None else N), 0
    g.a_major, g.b_major, g.epilogue, g.split_k = 1, 1, EPI_F32_ATOMIC, 0
    g.alpha = alpha
    g.dbg_lbo, g.dbg_sbo = 0, 0
    g.colsum, g.colsum_scale = 0, 1.0
    _run("cdr_gemm_grouped", lambda: _lib.load().cdr_gemm_grouped(C.byref(g), C.c_int32(n_groups), _p(seg_kb),
                                                                  C.c_int64(out_group_stride), stream_ptr()))
    _count(1)
    return out


# ------------------------------------------------------------------------------------------------
# helpers: explicit ctypes scalars (no argtypes are registered, so every scalar is wrapped here)
# ------------------------------------------------------------------------------------------------
def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


_i32, _i64, _f32 = C.c_int32, C.c_int64, C.c_float


def _lib_():
    return _lib.load()


def embed_ln_fwd(ids, word, pos, type0, gamma, beta, out, mean, rstd, *, n_seq, seq_len, hidden, vocab, eps):
    _need_cuda(ids, word, out)
    assert ids.dtype == torch.int64 and word.dtype == torch.float32 and out.dtype == torch.float16
    _run("cdr_embed_ln_fwd", lambda: _lib_().cdr_embed_ln_fwd(_p(ids), _p(word), _p(pos), _p(type0), _p(gamma), _p(beta), _p(out), _p(mean),
                                   _p(rstd), _i32(n_seq), _i32(seq_len), _i32(hidden), _i32(vocab), _f32(eps),
                                   stream_ptr()))
    _count(1)


def embed_ln_bwd(dy, ids, word, pos, type0, gamma, mean, rstd, dword, dpos, dtype0, dgamma, dbeta, *, n_seq, seq_len,
                 hidden, vocab, pad_id, in_scale, out_scale):
    _need_cuda(dy, ids, dword)
    _run("cdr_embed_ln_bwd", lambda: _lib_().cdr_embed_ln_bwd(_p(dy), _p(ids), _p(word), _p(pos), _p(type0), _p(gamma), _p(mean), _p(rstd),
                                   _p(dword), _p(dpos), _p(dtype0), _p(dgamma), _p(dbeta), _i32(n_seq), _i32(seq_len),
                                   _i32(hidden), _i32(vocab), _i32(pad_id), _f32(in_scale), _f32(out_scale),
                                   stream_ptr()))
    _count(1)


def ln_fwd(x, gamma, beta, y, mean, rstd, cls_out, *, n_seq, seq_len, hidden, eps, push=None):
    """push = (peer.PeerExchange, first_seq): the CLS rows of sequences >= first_seq are also written into every
    rank's gather buffer by the same kernel (fused all-gather over NVLink, cdr_ln_fwd_push)."""
    _need_cuda(x, y)
    assert x.dtype == torch.float16 and y.dtype == torch.float16
    if push is not None:
        xchg, first_seq = push
        assert n_seq - first_seq == xchg.n_rows and hidden == xchg.dim
        a = xchg.push_args()
        _run("cdr_ln_fwd_push", lambda: _lib_().cdr_ln_fwd_push(
            _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), _p(cls_out), _i32(n_seq), _i32(seq_len), _i32(hidden),
            _f32(eps), _i32(first_seq), C.byref(a), stream_ptr()))
        _count(1)
        return
    _run("cdr_ln_fwd", lambda: _lib_().cdr_ln_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), _p(cls_out), _i32(n_seq),
                             _i32(seq_len), _i32(hidden), _f32(eps), stream_ptr()))
    _count(1)


def ln_bwd(dy, dy_cls, x, gamma, mean, rstd, dx, dgamma, dbeta, dbias, *, n_seq, seq_len, hidden, in_scale=1.0,
           out_scale=1.0, row_ws=None):
    _need_cuda(x, dx)
    if row_ws is not None:
        assert row_ws.dtype == torch.float32 and row_ws.numel() >= 2 * n_seq * seq_len
    _run("cdr_ln_bwd", lambda: _lib_().cdr_ln_bwd(_p(dy), _p(dy_cls), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dgamma), _p(dbeta),
                             _p(dbias), _p(row_ws), _i32(n_seq), _i32(seq_len), _i32(hidden), _f32(in_scale),
                             _f32(out_scale), stream_ptr()))
    staged = dy is not None and dy_cls is None and hidden <= 1024  # one staged pass (elementwise.cu)
    _count(2 if (not staged and row_ws is not None and dy is not None and dy_cls is None) else 1)


def ln_bwd_drop(dy, x, gamma, mean, rstd, dx, dx_drop, dgamma, dbeta, dbias, *, rows, hidden, out_scale, drop):
    """LayerNorm backward after a dropped dense output: dx (residual branch), dx_drop = dropout'(dx) and
    dbias += out_scale * colsum(dx_drop) in one staged pass (cdr_ln_bwd_drop)."""
    _need_cuda(dy, x, dx, dx_drop)
    assert dy.dtype == torch.float16 and dy.is_contiguous() and x.is_contiguous()
    _run("cdr_ln_bwd_drop", lambda: _lib_().cdr_ln_bwd_drop(
        _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dx_drop), _p(dgamma), _p(dbeta), _p(dbias), _i32(rows),
        _i32(hidden), _f32(out_scale), C.byref(drop), stream_ptr()))
    _count(1)


def dropout_f16(x, out, *, drop):
    """out = dropout(x) for a contiguous fp16 [rows, cols] tensor (may alias)."""
    _need_cuda(x, out)
    assert x.dtype == torch.float16 and out.dtype == torch.float16 and x.is_contiguous() and out.is_contiguous()
    assert x.dim() == 2 and x.shape == out.shape
    _run("cdr_dropout_f16", lambda: _lib_().cdr_dropout_f16(_p(x), _p(out), _i64(x.shape[0]), _i32(x.shape[1]),
                                                           C.byref(drop), stream_ptr()))
    _count(1)
    return out


def colsum(x, out, *, rows, cols, ld=None, scale=1.0):
    _need_cuda(x, out)
    assert x.dtype == torch.float16 and out.dtype == torch.float32
    _run("cdr_colsum_f16", lambda: _lib_().cdr_colsum_f16(_p(x), _p(out), _i64(rows), _i64(cols), _i64(ld if ld is not None else x.stride(0)),
                                 _f32(scale), stream_ptr()))
    _count(1)


def cast_f32_f16(src, dst):
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and dst.dtype == torch.float16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    _run("cdr_cast_f32_f16", lambda: _lib_().cdr_cast_f32_f16(_p(src), _p(dst), _i64(src.numel()), stream_ptr()))
    _count(1)


def cast_table(entries, device):
    """Device-resident cdr_cast_item table for cast_multi: entries = [(src fp32, dst fp16|fp32), ...]."""
    rows = []
    for src, dst in entries:
        assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
        assert src.numel() == dst.numel() and src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0
        rows.append([src.data_ptr(), dst.data_ptr(), src.numel(), 1 if dst.dtype == torch.float32 else 0])
    # struct cdr_cast_item {ptr, ptr, int64, int32, int32}: 4 little-endian int64 words per entry
    table = torch.tensor(rows, dtype=torch.int64).to(device)
    return table, max(r[2] for r in rows)


def cast_multi(table, max_n):
    _run("cdr_cast_multi", lambda: _lib_().cdr_cast_multi(_p(table), _i32(table.shape[0]), _i64(max_n), stream_ptr()))
    _count(1)


def _attn_args(qkv, key_bias, out, lse, n_seq, seq_len, heads, scale, d_out=None, dqkv=None, dq_ws=None):
    a = _lib.AttnArgs()
    a.dq_workspace = dq_ws.data_ptr() if dq_ws is not None else 0
    a.qkv, a.out, a.lse = qkv.data_ptr(), out.data_ptr(), lse.data_ptr()
    a.key_bias = key_bias.data_ptr() if key_bias is not None else 0
    a.d_out = d_out.data_ptr() if d_out is not None else 0
    a.dqkv = dqkv.data_ptr() if dqkv is not None else 0
    a.n_seq, a.seq_len, a.heads, a.head_dim, a.scale = n_seq, seq_len, heads, 64, scale
    return a


def attn_dropout_bits(n_seq, heads, seq_len, device):
    """Buffer for the keep bits of the attention-probability dropout (filled by attn_fwd, read by attn_bwd)."""
    n = int(_lib_().cdr_attn_dropout_bits_bytes(_i32(n_seq), _i32(heads), _i32(seq_len)))
    return torch.empty(n, dtype=torch.uint8, device=device)


def attn_fill_bits(drop_bits, *, n_seq, seq_len, heads, drop):
    """The generator pass of attn_fwd on its own (current stream): afterwards attn_fwd(..., bits_ready=True)."""
    assert drop_bits.is_cuda and drop_bits.dtype == torch.uint8 and drop is not None
    a = _lib.AttnArgs()
    a.n_seq, a.seq_len, a.heads, a.head_dim = n_seq, seq_len, heads, 64
    a.drop, a.drop_bits = drop, drop_bits.data_ptr()
    _run("cdr_attn_dropout_bits_fill", lambda: _lib_().cdr_attn_dropout_bits_fill(C.byref(a), stream_ptr()))
    _count(1)


def attn_fwd(qkv, key_bias, out, lse, *, n_seq, seq_len, heads, scale=0.125, drop=None, drop_bits=None, bits_ready=False):
    """drop (a drop_args descriptor) needs drop_bits = attn_dropout_bits(...): fwd fills it (unless bits_ready: filled
    earlier by attn_fill_bits), bwd reads it."""
    _need_cuda(qkv, out, lse)
    assert qkv.dtype == torch.float16 and out.dtype == torch.float16 and lse.dtype == torch.float32
    assert qkv.is_contiguous() and out.is_contiguous()
    a = _attn_args(qkv, key_bias, out, lse, n_seq, seq_len, heads, scale)
    if drop is not None:
        assert drop_bits is not None and drop_bits.is_cuda and drop_bits.dtype == torch.uint8
        a.drop, a.drop_bits = drop, drop_bits.data_ptr()
        a.drop_bits_ready = 1 if bits_ready else 0
    _run("cdr_attn_fwd", lambda: _lib_().cdr_attn_fwd(C.byref(a), stream_ptr()))
    _count(1 if drop is None or bits_ready else 2)


def attn_bwd(qkv, key_bias, out, lse, d_out, dqkv, *, n_seq, seq_len, heads, scale=0.125, dbias=None, dbias_scale=1.0,
             drop=None, drop_bits=None):
    """dbias (optional fp32 [3*heads*64], seq_len <= 128): += dbias_scale * column sums of dqkv, fused."""
    _need_cuda(qkv, out, lse, d_out, dqkv)
    assert d_out.dtype == torch.float16 and dqkv.dtype == torch.float16 and d_out.is_contiguous()
    dq_ws = None
    if seq_len > 128:  # tiled backward: key tiles add their dQ shares in an fp32 scratch
        dq_ws = torch.empty(n_seq * seq_len, heads * 64, dtype=torch.float32, device=qkv.device)
    a = _attn_args(qkv, key_bias, out, lse, n_seq, seq_len, heads, scale, d_out, dqkv, dq_ws)
    if drop is not None:
        assert drop_bits is not None
        a.drop, a.drop_bits = drop, drop_bits.data_ptr()
    if dbias is not None:
        assert dbias.dtype == torch.float32 and dbias.numel() == 3 * heads * 64
        a.dbias_qkv, a.dbias_scale = dbias.data_ptr(), dbias_scale
    _run("cdr_attn_bwd", lambda: _lib_().cdr_attn_bwd(C.byref(a), stream_ptr()))
    _count(1 if seq_len <= 128 else 2)


def pair_nll_fwd(q, a, b, loss, accs, logits):
    _need_cuda(q, a, b)
    n, dim = q.shape
    _run("cdr_pair_nll_fwd", lambda: _lib_().cdr_pair_nll_fwd(_p(q), _p(a), _p(b), _i32(n), _i32(dim), _p(loss), _p(accs), _p(logits),
                                   stream_ptr()))
    _count(1)


def pair_nll_bwd(q, a, b, logits, dloss, dq, da, db):
    n, dim = q.shape
    _run("cdr_pair_nll_bwd", lambda: _lib_().cdr_pair_nll_bwd(_p(q), _p(a), _p(b), _p(logits), _p(dloss), _i32(n), _i32(dim), _p(dq), _p(da),
                                   _p(db), stream_ptr()))
    _count(1)


SIM_QP, SIM_COCO = 0, 1


def _simmat_args(q, k, lse, mode, row_offset, loss_scale):
    a = _lib.SimmatArgs()
    a.q, a.k, a.lse = q.data_ptr(), k.data_ptr(), lse.data_ptr()
    a.n_rows, a.n_keys, a.dim = q.shape[0], k.shape[0], q.shape[1]
    a.mode, a.row_offset, a.loss_scale = mode, row_offset, loss_scale
    return a


def simmat_workspace(n_rows, n_keys, device):
    """fp32 workspace of the fused similarity / cross-entropy forward (per-split softmax partials, NOT the scores)."""
    n = int(_lib_().cdr_simmat_workspace_bytes(_i32(n_rows), _i32(n_keys)))
    return torch.empty(max(1, n // 4), dtype=torch.float32, device=device)


def simmat_ce_fwd(q, k, loss, lse, *, mode, row_offset=0, loss_scale=1.0, workspace=None):
    _need_cuda(q, k, loss, lse)
    assert q.dtype == torch.float32 and k.dtype == torch.float32 and q.is_contiguous() and k.is_contiguous()
    ws = workspace if workspace is not None else simmat_workspace(q.shape[0], k.shape[0], q.device)
    a = _simmat_args(q, k, lse, mode, row_offset, loss_scale)
    a.loss, a.scores = loss.data_ptr(), ws.data_ptr()
    _run("cdr_simmat_ce_fwd", lambda: _lib_().cdr_simmat_ce_fwd(C.byref(a), stream_ptr()))
    _count(2)


def simmat_ce_bwd(q, k, lse, dloss, dq, dk, *, mode, row_offset=0, loss_scale=1.0):
    _need_cuda(q, k, lse, dloss)
    assert dloss.dtype == torch.float32 and dloss.is_contiguous()
    a = _simmat_args(q, k, lse, mode, row_offset, loss_scale)
    a.dloss = dloss.data_ptr()
    a.dq = dq.data_ptr() if dq is not None else 0
    a.dk = dk.data_ptr() if dk is not None else 0
    _run("cdr_simmat_ce_bwd", lambda: _lib_().cdr_simmat_ce_bwd(C.byref(a), stream_ptr()))
    _count((dq is not None) + (dk is not None))


def own_key_grad(q, loss, dloss, dk_own):
    """dk_own[i, :] = dloss[i] * (exp(-loss[i]) - 1) * q[i, :]: the gradient an in-batch InfoNCE row sends to its own
    (positive) key (cdr_simmat_own_key_grad)."""
    _need_cuda(q, loss, dloss, dk_own)
    assert q.dtype == torch.float32 and q.is_contiguous() and dk_own.is_contiguous() and dk_own.shape == q.shape
    _run("cdr_simmat_own_key_grad", lambda: _lib_().cdr_simmat_own_key_grad(
        _p(q), _p(loss), _p(dloss), _i32(q.shape[0]), _i32(q.shape[1]), _p(dk_own), stream_ptr()))
    _count(1)


def vocab_ce_fwd(logits, bias, labels, loss, lse, *, n_cols):
    _need_cuda(logits, bias, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64 and logits.stride(1) == 1
    _run("cdr_vocab_ce_fwd", lambda: _lib_().cdr_vocab_ce_fwd(_p(logits), _p(bias), _p(labels), _p(loss), _p(lse),
                                                             _i32(logits.shape[0]), _i32(n_cols), _i64(logits.stride(0)),
                                                             stream_ptr()))
    _count(1)


def vocab_ce_bwd(logits, bias, labels, lse, dloss, dlogits, *, n_cols, scale):
    assert dlogits.dtype == torch.float16 and dlogits.stride(0) == logits.stride(0)
    _run("cdr_vocab_ce_bwd", lambda: _lib_().cdr_vocab_ce_bwd(_p(logits), _p(bias), _p(labels), _p(lse), _p(dloss),
                                                             _p(dlogits), _i32(logits.shape[0]), _i32(n_cols),
                                                             _i64(logits.stride(0)), _f32(scale), stream_ptr()))
    _count(1)


def dgelu(dt, z, dz):
    assert dt.dtype == torch.float16 and z.dtype == torch.float16 and dt.is_contiguous() and z.is_contiguous()
    _run("cdr_dgelu_f16", lambda: _lib_().cdr_dgelu_f16(_p(dt), _p(z), _p(dz), _i64(dt.numel()), stream_ptr()))
    _count(1)


def group_reduce_fwd(loss, g, sums, counts, *, n_groups):
    _need_cuda(loss, g, sums, counts)
    assert g.dtype == torch.int64 and loss.dtype == torch.float32
    _run("cdr_group_reduce_fwd", lambda: _lib_().cdr_group_reduce_fwd(_p(loss), _p(g), _i32(loss.numel()), _i32(n_groups), _p(sums), _p(counts),
                                       stream_ptr()))
    _count(1)


def group_reduce_bwd(dsums, g, dloss, *, n_groups):
    _run("cdr_group_reduce_bwd", lambda: _lib_().cdr_group_reduce_bwd(_p(dsums), _p(g), _i32(g.numel()), _i32(n_groups), _p(dloss), stream_ptr()))
    _count(1)


def gram_f32(x, gram):
    """gram[G,G] += x x^T for a row-major fp32 [G, P] matrix (row stride x.stride(0))."""
    _need_cuda(x, gram)
    assert x.dtype == torch.float32 and gram.dtype == torch.float32 and x.stride(1) == 1
    _run("cdr_gram_f32", lambda: _lib_().cdr_gram_f32(_p(x), _i32(x.shape[0]), _i64(x.shape[1]), _i64(x.stride(0)), _p(gram), stream_ptr()))
    _count(1)


def scan_workspace_bytes(n_docs, n_q, k, dim=768):
    return int(_lib_().cdr_scan_workspace_bytes(_i64(n_docs), _i32(n_q), _i32(k), _i32(dim)))


def scan_exhaustive_docs(k):
    return int(_lib_().cdr_scan_exhaustive_docs(_i32(k)))


def scan_topk(docs, queries, out_scores, out_ids, workspace, status, *, k, doc_base=0):
    _need_cuda(docs, queries, out_scores, out_ids, workspace, status)
    assert docs.dtype == torch.float16 and queries.dtype == torch.float16 and queries.is_contiguous()
    assert out_scores.dtype == torch.float32 and out_ids.dtype == torch.int64 and status.dtype == torch.int32
    a = _lib.ScanArgs()
    a.docs, a.queries, a.out_scores, a.out_ids = docs.data_ptr(), queries.data_ptr(), out_scores.data_ptr(), out_ids.data_ptr()
    a.workspace, a.workspace_bytes, a.status = workspace.data_ptr(), workspace.numel() * workspace.element_size(), status.data_ptr()
    a.n_docs, a.ld_docs, a.doc_base = docs.shape[0], docs.stride(0), doc_base
    a.n_q, a.dim, a.k = queries.shape[0], queries.shape[1], k
    _run("cdr_scan_topk", lambda: _lib_().cdr_scan_topk(C.byref(a), stream_ptr()))
    _count(3 if docs.shape[0] <= scan_exhaustive_docs(k) else 5)


def topk_merge(scores, ids, out_scores, out_ids, *, k):
    _need_cuda(scores, ids, out_scores, out_ids)
    assert scores.is_contiguous() and ids.is_contiguous() and ids.dtype == torch.int64
    n_q, n_in = scores.shape
    _run("cdr_topk_merge", lambda: _lib_().cdr_topk_merge(_p(scores), _p(ids), _i32(n_q), _i32(n_in), _i32(k), _p(out_scores), _p(out_ids),
                                 stream_ptr()))
    _count(1)


def topk_pack(scores, ids, status, keys, *, ks, all_returned):
    """[n_q, k_in] shard result -> u64 keys [n_q * ks + 2] (cdr_topk_pack)."""
    _need_cuda(scores, ids, status, keys)
    n_q, k_in = scores.shape
    assert keys.dtype == torch.int64 and keys.numel() == n_q * ks + 2 and scores.is_contiguous() and ids.is_contiguous()
    _run("cdr_topk_pack", lambda: _lib_().cdr_topk_pack(_p(scores), _p(ids), _i32(n_q), _i32(k_in), _i32(ks), _p(status),
                                                       _i32(int(all_returned)), _p(keys), stream_ptr()))
    _count(1)


def topk_merge_keys(gathered, out_scores, out_ids, flag, *, world, n_q, ks, k):
    """Merge the all-gathered key blocks of every shard into the global top k (cdr_topk_merge_keys)."""
    _need_cuda(gathered, out_scores, out_ids, flag)
    assert gathered.dtype == torch.int64 and gathered.numel() == world * (n_q * ks + 2) and flag.dtype == torch.int32
    _run("cdr_topk_merge_keys", lambda: _lib_().cdr_topk_merge_keys(_p(gathered), _i32(world), _i32(n_q), _i32(ks), _i32(k),
                                                                   _p(out_scores), _p(out_ids), _p(flag), stream_ptr()))
    _count(1)
