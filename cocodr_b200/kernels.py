"""Thin tensor-level wrappers over the C ABI (one python function per cdr_* entry point).

These only translate torch tensors into raw pointers / sizes and raise on error; all arithmetic is in
the CUDA library.  Device memory and streams are torch's (plumbing, not product).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_F32_ATOMIC, EPI_F32_STORE,  # noqa: F401
                   EPI_STORE_F16, check, ptr, stream_ptr)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cocodr_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, b, out, *, M, N, K, a_major=0, b_major=0, epilogue=EPI_STORE_F16, bias=None, aux=None, out2=None,
         alpha=1.0, split_k=1, lda=None, ldb=None, ldo=None, ldaux=None, dbg_lbo=0, dbg_sbo=0):
    """out[M,N] = alpha * A[M,K] @ B[N,K]^T with a fused epilogue (see include/cocodr_b200.h)."""
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
    if bias is not None:
        assert bias.dtype == torch.float32
    check(_lib.load().cdr_gemm(C.byref(g), stream_ptr()), "cdr_gemm")
    return out
