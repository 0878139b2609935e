#!/usr/bin/env python
"""The tcgen05 GEMM against cuBLAS on large square-ish problems (steady-state main loop, no wave or epilogue
effects): where does the kernel sit relative to the library peak MEASURED_PEAKS.json quotes?  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocodr_b200 import kernels as k  # noqa: E402

ROT = 2


def bench(name, fn, flops, iters=10):
    for i in range(3):
        fn(i % ROT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ROT)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:40s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)


for (M, N, K) in [(8192, 8192, 8192), (16384, 3072, 3072), (16384, 768, 3072), (16384, 2304, 768)]:
    a = [(torch.randn(M, K, device="cuda") * 0.1).half() for _ in range(ROT)]
    b = [(torch.randn(N, K, device="cuda") * 0.1).half() for _ in range(ROT)]
    o = [torch.empty(M, N, dtype=torch.float16, device="cuda") for _ in range(ROT)]
    bench(f"ours   nt [{M},{N},{K}]", lambda r: k.gemm(a[r], b[r], o[r], M=M, N=N, K=K), 2.0 * M * N * K)
    bench(f"cuBLAS nt [{M},{N},{K}]", lambda r: torch.matmul(a[r], b[r].t(), out=o[r]), 2.0 * M * N * K)
    del a, b, o
# wgrad shape: both operands MN-major, fp32 accumulate-add, split-K
T, H, I = 16384, 768, 3072
x = [(torch.randn(T, H, device="cuda") * 0.1).half() for _ in range(ROT)]
g = [(torch.randn(T, I, device="cuda") * 0.1).half() for _ in range(ROT)]
dw = torch.zeros(H, I, device="cuda")
bench("ours   wgrad [768,3072,16384]", lambda r: k.gemm(x[r], g[r], dw, M=H, N=I, K=T, a_major=1, b_major=1, epilogue=k.EPI_F32_ATOMIC, split_k=0), 2.0 * T * H * I)
bench("cuBLAS wgrad [768,3072,16384]", lambda r: torch.matmul(x[r].t(), g[r]), 2.0 * T * H * I)
