#!/usr/bin/env python
"""Timing of the fused attention kernels at the bench shape (128 sequences x 128 tokens x 12 heads). GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocodr_b200 import kernels as k  # noqa: E402

n_seq, L, heads = 128, 128, 12
H, T = heads * 64, n_seq * L
ROT = 4
qkv = [(torch.randn(T, 3 * H, device="cuda") * 1.5).half() for _ in range(ROT)]
out = [torch.zeros(T, H, dtype=torch.float16, device="cuda") for _ in range(ROT)]
lse = [torch.zeros(n_seq, heads, L, device="cuda") for _ in range(ROT)]
do = [torch.randn(T, H, device="cuda").half() for _ in range(ROT)]
dqkv = [torch.zeros(T, 3 * H, dtype=torch.float16, device="cuda") for _ in range(ROT)]
db = torch.zeros(3 * H, device="cuda")


NCU = "--ncu" in sys.argv  # under ncu: two launches per variant, nothing timed


def bench(name, fn, flops, bytes_, iters=40):
    if NCU:
        fn(0), fn(1)
        return
    for i in range(4):
        fn(i % ROT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ROT)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:12s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {bytes_ / us / 1e3:7.1f} GB/s", flush=True)


fw = 4 * L * L * 64 * heads * n_seq
bench("attn fwd", lambda r: k.attn_fwd(qkv[r], None, out[r], lse[r], n_seq=n_seq, seq_len=L, heads=heads), fw, T * 4 * H * 2)
bench("attn bwd", lambda r: k.attn_bwd(qkv[r], None, out[r], lse[r], do[r], dqkv[r], n_seq=n_seq, seq_len=L, heads=heads,
                                      dbias=db, dbias_scale=1e-3), 2.5 * fw, T * (3 + 1 + 3) * H * 2)
bench("bwd no-dbias", lambda r: k.attn_bwd(qkv[r], None, out[r], lse[r], do[r], dqkv[r], n_seq=n_seq, seq_len=L, heads=heads),
      2.5 * fw, T * (3 + 1 + 3) * H * 2)
state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
drop = k.drop_args(state, 5, 0.1)
bits = k.attn_dropout_bits(n_seq, heads, L, "cuda")
bench("fwd drop", lambda r: k.attn_fwd(qkv[r], None, out[r], lse[r], n_seq=n_seq, seq_len=L, heads=heads, drop=drop, drop_bits=bits),
      fw, T * 4 * H * 2)
bench("bwd drop", lambda r: k.attn_bwd(qkv[r], None, out[r], lse[r], do[r], dqkv[r], n_seq=n_seq, seq_len=L, heads=heads,
                                      dbias=db, dbias_scale=1e-3, drop=drop, drop_bits=bits), 2.5 * fw, T * (3 + 1 + 3) * H * 2)
