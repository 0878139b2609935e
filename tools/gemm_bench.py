#!/usr/bin/env python
"""Per-shape timing of the 12 GEMMs of one BERT-base layer step (T = 16384 tokens) through the C ABI, with the
epilogues the layer uses, rotating operand sets so consecutive launches do not hit L2.  GPU only (gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocodr_b200 import kernels as k  # noqa: E402

T, H, I = 16384, 768, 3072
ROT = 3


def rnd(*s):
    return [(torch.randn(*s, device="cuda") * 0.1).half() for _ in range(ROT)]


def bench(name, fn, flops, iters=30):
    for i in range(3):
        fn(i % ROT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ROT)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:34s} {us:8.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)
    return us


def main():
    xTH, xTI, xT3H = rnd(T, H), rnd(T, I), rnd(T, 3 * H)
    wqkv, wo, wi, wo2 = rnd(3 * H, H), rnd(H, H), rnd(I, H), rnd(H, I)
    oTH, oTI, oT3H, oTI2 = rnd(T, H), rnd(T, I), rnd(T, 3 * H), rnd(T, I)
    b3h, bh, bi = torch.randn(3 * H, device="cuda"), torch.randn(H, device="cuda"), torch.randn(I, device="cuda")
    dW = {n: torch.zeros(s, device="cuda") for n, s in dict(qkv=(3 * H, H), o=(H, H), i=(I, H), o2=(H, I)).items()}
    cs = torch.zeros(I, device="cuda")
    tot = 0.0
    tot += bench("fwd QKV   [T,2304,768] bias", lambda r: k.gemm(xTH[r], wqkv[r], oT3H[r], M=T, N=3 * H, K=H, bias=b3h), 2 * T * 3 * H * H)
    tot += bench("fwd O     [T,768,768] bias+res", lambda r: k.gemm(xTH[r], wo[r], oTH[r], M=T, N=H, K=H, bias=bh, epilogue=k.EPI_BIAS_RESIDUAL, aux=xTH[(r + 1) % ROT]), 2 * T * H * H)
    tot += bench("fwd FFN1  [T,3072,768] gelu", lambda r: k.gemm(xTH[r], wi[r], oTI[r], M=T, N=I, K=H, bias=bi, epilogue=k.EPI_BIAS_GELU, out2=oTI2[r]), 2 * T * I * H)
    tot += bench("fwd FFN2  [T,768,3072] bias+res", lambda r: k.gemm(xTI[r], wo2[r], oTH[r], M=T, N=H, K=I, bias=bh, epilogue=k.EPI_BIAS_RESIDUAL, aux=xTH[(r + 1) % ROT]), 2 * T * H * I)
    tot += bench("dgrad FFN2 [T,3072,768] dgelu+cs", lambda r: k.gemm(xTH[r], wo2[r], oTI[r], M=T, N=I, K=H, b_major=1, epilogue=k.EPI_DGELU, aux=xTI[r], colsum=cs, colsum_scale=1e-3), 2 * T * I * H)
    tot += bench("wgrad FFN2 [768,3072,T]", lambda r: k.gemm(xTH[r], xTI[r], dW["o2"], M=H, N=I, K=T, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=0), 2 * T * I * H)
    tot += bench("dgrad FFN1 [T,768,3072] res", lambda r: k.gemm(xTI[r], wi[r], oTH[r], M=T, N=H, K=I, b_major=1, epilogue=k.EPI_BIAS_RESIDUAL, aux=xTH[(r + 1) % ROT]), 2 * T * I * H)
    tot += bench("wgrad FFN1 [3072,768,T]", lambda r: k.gemm(xTI[r], xTH[r], dW["i"], M=I, N=H, K=T, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=0), 2 * T * I * H)
    tot += bench("dgrad O   [T,768,768]", lambda r: k.gemm(xTH[r], wo[r], oTH[r], M=T, N=H, K=H, b_major=1), 2 * T * H * H)
    tot += bench("wgrad O   [768,768,T]", lambda r: k.gemm(xTH[r], xTH[(r + 1) % ROT], dW["o"], M=H, N=H, K=T, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=0), 2 * T * H * H)
    tot += bench("dgrad QKV [T,768,2304] res", lambda r: k.gemm(xT3H[r], wqkv[r], oTH[r], M=T, N=H, K=3 * H, b_major=1, epilogue=k.EPI_BIAS_RESIDUAL, aux=xTH[(r + 1) % ROT]), 2 * T * 3 * H * H)
    tot += bench("wgrad QKV [2304,768,T]", lambda r: k.gemm(xT3H[r], xTH[r], dW["qkv"], M=3 * H, N=H, K=T, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=0), 2 * T * 3 * H * H)
    fl = 3 * 2 * T * (3 * H * H + H * H + 2 * H * I)
    print(f"layer total: {tot:.1f} us, {fl / tot / 1e6:.1f} TFLOP/s; x12 layers = {12 * tot / 1e3:.2f} ms", flush=True)
    for (M, N, K) in [(T, 3 * H, H), (T, H, I)]:
        a, b = xTH if K == H else xTI, wqkv if N == 3 * H else wo2
        bench(f"cuBLAS [{M},{N},{K}]", lambda r: torch.matmul(a[r], b[r].t()), 2 * M * N * K)


if __name__ == "__main__":
    main()
