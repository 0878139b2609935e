#!/usr/bin/env python
"""Corpus-scan timing: 1M x 768 fp16 docs, query tiles of 1/16/128/1000, whole search and (under ncu) per kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocodr_b200 import scan  # noqa: E402

N, D = 1_000_000, 768
g = torch.Generator(device="cuda").manual_seed(7)
P = torch.randn(N, D, generator=g, device="cuda", dtype=torch.float16)
Q = torch.randn(1000, D, generator=g, device="cuda", dtype=torch.float16)
iters = int(os.environ.get("SCAN_ITERS", "10"))
for nq, k in ((1, 100), (16, 100), (128, 100), (128, 1000), (1000, 1000)):
    q = Q[:nq].contiguous()
    for _ in range(2):
        scan.search(q, P, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        scan.search(q, P, k)
    e1.record()
    torch.cuda.synchronize()
    ms_sync = e0.elapsed_time(e1) / iters
    scan.check_status([scan.search_async(q, P, k) for _ in range(iters)], q, P, k)  # allocator warm-up
    e0.record()
    res = [scan.search_async(q, P, k) for _ in range(iters)]
    scan.check_status(res, q, P, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"   (status checked per search: {ms_sync * 1e3:.1f} us; below: {iters} searches in flight, one check)")
    print(f"nq={nq:5d} k={k:5d}: {ms * 1e3:8.1f} us/search  {N * D * 2 / ms / 1e6:7.1f} GB/s  "
          f"{2.0 * nq * N * D / ms / 1e9:7.1f} TFLOP/s  {nq / ms * 1e3:9.0f} q/s", flush=True)
