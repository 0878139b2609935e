#!/usr/bin/env python
"""LayerNorm fwd / bwd timing at the bench shape (16384 x 768), rotating buffers (> L2). GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocodr_b200 import kernels as k  # noqa: E402

n_seq, L, H = 128, 128, 768
T = n_seq * L
ROT = 6
x = [torch.randn(T, H, device="cuda").half() for _ in range(ROT)]
dy = [torch.randn(T, H, device="cuda").half() for _ in range(ROT)]
y = [torch.empty(T, H, device="cuda", dtype=torch.float16) for _ in range(ROT)]
gamma, beta = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
dg, db, dc = (torch.zeros(H, device="cuda") for _ in range(3))
ws = torch.empty(2 * T, device="cuda")


def bench(name, fn, bytes_, iters=60):
    for i in range(6):
        fn(i % ROT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ROT)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:10s} {us:7.1f} us  {bytes_ / us / 1e3:7.1f} GB/s", flush=True)


bench("ln fwd", lambda r: k.ln_fwd(x[r], gamma, beta, y[r], mean, rstd, None, n_seq=n_seq, seq_len=L, hidden=H, eps=1e-12), T * H * 4)
bench("ln bwd", lambda r: k.ln_bwd(dy[r], None, x[r], gamma, mean, rstd, y[r], dg, db, dc, n_seq=n_seq, seq_len=L, hidden=H, row_ws=ws), T * H * 6)
state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
dxm = [torch.empty(T, H, device="cuda", dtype=torch.float16) for _ in range(ROT)]
drop = k.drop_args(state, 3, 0.1)
bench("ln bwd drop", lambda r: k.ln_bwd_drop(dy[r], x[r], gamma, mean, rstd, y[r], dxm[r], dg, db, dc, rows=T, hidden=H, out_scale=1.0, drop=drop), T * H * 8)
