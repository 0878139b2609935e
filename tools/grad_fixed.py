#!/usr/bin/env python
"""Backward accuracy for a FIXED upstream gradient (well conditioned): per-parameter relative error of the CUDA
encoder backward vs fp32 autograd through the oracle.  Run from a tree root (old or new) for A/B comparisons."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_model_gpu as T
from oracle import bert_ref
for cfgname, cfg, n, L in (("tiny", T.TINY, 6, 32),):
    m = T.build(cfg); m.train()
    ids, mask = (t.cuda() for t in bert_ref.synth_batch(n, L, cfg["vocab"], 77))
    g = torch.Generator().manual_seed(5)
    dcls = torch.randn(n, cfg["hidden"], generator=g) * 0.05
    cls = m.query_emb(ids, mask)
    (cls * dcls.cuda()).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(cfg, 0).items()}
    ref_cls = bert_ref.cls_embedding(leaf, ids.cpu(), mask.cpu(), cfg)
    (ref_cls * dcls).sum().backward()
    named = dict(m.bert.named_parameters())
    res = sorted(((T.rel(named[k].grad.cpu().numpy(), v.grad.numpy(), floor=1e-4), k) for k, v in leaf.items()), reverse=True)
    print(cfgname, "fwd rel", T.rel(cls.detach().cpu().numpy(), ref_cls.detach().numpy()), "mean grad rel", np.mean([r for r, _ in res]))
    for r, k in res[:6]: print(f"   {r:.5f} {k}")
    by = {}
    for r, k in res:
        kind = k.split(".")[-2] + "." + k.split(".")[-1]
        by.setdefault(kind, []).append(r)
    print("   by kind:", {k: round(float(np.mean(v)), 5) for k, v in sorted(by.items())})
