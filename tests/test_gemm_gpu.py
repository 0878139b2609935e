"""GPU: tcgen05 GEMM (all operand layouts / epilogues) vs a torch fp32 reference of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def _check(got, ref, tol=2e-2, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (384, 768, 768), (200, 384, 136), (1024, 2304, 768),
                                   (8192, 768, 3072)])
def test_nt_store(M, N, K):
    from cocodr_b200 import kernels as k
    a, b = _rand((M, K), 1), _rand((N, K), 2, 0.1)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    k.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
    _check(out, a.float() @ b.float().t() + bias, what=f"NT {M}x{N}x{K}")


def test_nt_gelu_residual():
    from cocodr_b200 import kernels as k
    M, N, K = 512, 1024, 256
    a, b = _rand((M, K), 3), _rand((N, K), 4, 0.1)
    bias = torch.randn(N, device="cuda") * 0.1
    res = _rand((M, N), 5)
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    gp = torch.empty_like(out)
    k.gemm(a, b, out, M=M, N=N, K=K, bias=bias, epilogue=k.EPI_BIAS_GELU, out2=gp)
    zr = (a.float() @ b.float().t() + bias).requires_grad_(True)
    ref = torch.nn.functional.gelu(zr)
    ref.sum().backward()
    _check(out, ref.detach(), tol=2e-3, what="gelu(z)")
    _check(gp, zr.grad, tol=2e-3, what="gelu'(z) saved for the backward")
    zr = zr.detach()
    k.gemm(a, b, out, M=M, N=N, K=K, bias=bias, epilogue=k.EPI_BIAS_RESIDUAL, aux=res)
    _check(out, zr + res.float(), what="bias+residual")
    o32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    k.gemm(a, b, o32, M=M, N=N, K=K, epilogue=k.EPI_F32_STORE, alpha=0.5)
    _check(o32, 0.5 * (a.float() @ b.float().t()), tol=1e-3, what="f32 store")


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (512, 768, 3072), (1000, 3072, 768)])
def test_dgrad_b_mn_major(M, N, K):
    """dx[M,N] = dy[M,K] @ W[K,N]   (B stored [K, N] row-major = MN-major)."""
    from cocodr_b200 import kernels as k
    dy, w = _rand((M, K), 6), _rand((K, N), 7, 0.1)
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    k.gemm(dy, w, out, M=M, N=N, K=K, b_major=1)
    _check(out, dy.float() @ w.float(), what="dgrad")
    gp = _rand((M, N), 8, 0.5)  # the derivative tensor the forward GELU epilogue saves
    csum = torch.ones(N, dtype=torch.float32, device="cuda")
    k.gemm(dy, w, out, M=M, N=N, K=K, b_major=1, epilogue=k.EPI_DGELU, aux=gp, colsum=csum, colsum_scale=0.5)
    ref = (dy.float() @ w.float()) * gp.float()
    _check(out, ref, what="dgrad*gelu'")
    _check(csum, 1.0 + 0.5 * ref.sum(0), tol=2e-3, what="fused bias gradient (column sums)")


@pytest.mark.parametrize("M,N,K,split", [(128, 128, 256, 1), (768, 768, 4096, 0), (3072, 768, 2048, 0), (128, 512, 1000, 3)])
def test_wgrad_both_mn_major(M, N, K, split):
    """dW[M,N] += dy[K,M]^T @ x[K,N]   (both operands MN-major, split-K with fp32 atomics)."""
    from cocodr_b200 import kernels as k
    dy, x = _rand((K, M), 9), _rand((K, N), 10)
    out = torch.ones(M, N, dtype=torch.float32, device="cuda")
    k.gemm(dy, x, out, M=M, N=N, K=K, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=split, alpha=0.25)
    _check(out, 1.0 + 0.25 * (dy.float().t() @ x.float()), tol=2e-3, what="wgrad")


def test_row_segments():
    """cdr_gemm_segments: the wgrad GEMM on row ranges of its operands, each into its own fp32 block (iDRO K11)."""
    from cocodr_b200 import kernels as k
    rows, M, N = 1000, 128, 192
    a, b = _rand((rows, 2 * M), 11), _rand((rows, N), 12)
    a_cols = a[:, M:]  # a column block of a wider tensor (dQKV -> query / key / value), row stride 2M
    begin, count = [0, 128, 131, 640], [128, 3, 509, 360]
    out = torch.ones(6, M * N, dtype=torch.float32, device="cuda")
    offs = [4 * M * N, 0, 2 * M * N, 5 * M * N]
    k.gemm_segments(a_cols, b, out, M=M, N=N, row_begin=begin, row_count=count, out_offset=offs, alpha=0.5)
    for r0, n, o in zip(begin, count, offs):
        ref = 1.0 + 0.5 * (a_cols[r0:r0 + n].float().t() @ b[r0:r0 + n].float())
        _check(out.view(-1)[o:o + M * N].view(M, N), ref, tol=2e-3, what=f"segment rows {r0}+{n}")
    assert torch.equal(out[1], torch.ones_like(out[1])) and torch.equal(out[3], torch.ones_like(out[3]))
    with pytest.raises(RuntimeError):
        k.gemm_segments(a_cols, b, out, M=M, N=N, row_begin=[0], row_count=[8], out_offset=[0], epilogue=k.EPI_STORE_F16)


@pytest.mark.parametrize("M,N", [(128, 192), (256, 384), (768, 128)])
def test_grouped_k_ranges(M, N):
    """cdr_gemm_grouped: one launch, (tile, group) work items, per-group k-block ranges read from device memory."""
    from cocodr_b200 import kernels as k
    kblocks = [2, 0, 1, 5, 0, 3]  # 64-row blocks per group (two empty groups)
    G, rows = len(kblocks), 64 * sum(kblocks) + 64  # one spare block behind the last group: must not be read
    a, b = _rand((rows, M), 21), _rand((rows, N), 22)
    seg = torch.tensor([sum(kblocks[:i]) for i in range(G + 1)], dtype=torch.int32, device="cuda")
    stride, off = M * N + 64, 32
    out = torch.ones(G * stride + off, dtype=torch.float32, device="cuda")
    k.gemm_grouped(a, b, out, M=M, N=N, seg_kb=seg, n_groups=G, out_group_stride=stride, out_offset=off, alpha=0.25)
    r0 = 0
    for g, nb in enumerate(kblocks):
        blk = out[off + g * stride: off + g * stride + M * N].view(M, N)
        ref = 1.0 + 0.25 * (a[r0:r0 + 64 * nb].float().t() @ b[r0:r0 + 64 * nb].float())
        _check(blk, ref, tol=2e-3, what=f"group {g} ({nb} k-blocks)")
        if nb == 0:
            assert torch.equal(blk, torch.ones_like(blk))
        tail = out[off + g * stride + M * N: off + (g + 1) * stride]
        assert torch.equal(tail, torch.ones_like(tail))  # nothing written between the groups' blocks
        r0 += 64 * nb


def test_errors_are_loud():
    from cocodr_b200 import kernels as k
    a, b = _rand((128, 60), 1), _rand((128, 60), 2)
    out = torch.empty(128, 128, dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError):
        k.gemm(a, b, out, M=128, N=128, K=60)
