#!/usr/bin/env python
"""Print the per-step deviations the DRO trajectory tests assert on (tests/test_model_gpu.py::_run_dro)."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_model_gpu as T
for fname, kind in (("dro_greedy_tiny.npz", "dro-greedy"), ("idro_tiny.npz", "idro")):
    g = np.load(os.path.join("tests", "golden", fname))
    m = T.build(T.TINY)
    B, L, seed, G = int(g["B"]), int(g["L"]), int(g["seed"]), int(g["n_groups"])
    m.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=0), G, kind, float(g["alpha"]), float(g["eps"]),
                     float(g["ema"]), float(g["rho"]), True)
    m.train()
    for s in range(int(g["steps"])):
        q, mq, a, ma, b, mb = T.triplet(T.TINY, B, L, seed + 10 * s)
        gid = torch.from_numpy(g[f"group_ids_{s}"]).cuda()
        m.zero_grad()
        robust, acc, gl, gc = m(q, mq, a, ma, b, mb, group_ids=gid, weights=torch.ones(B, device="cuda"))
        robust.backward()
        got = dict(m.bert.named_parameters())["encoder.layer.11.attention.self.query.weight"].grad.cpu().numpy()
        print(kind, s, "robust rel", abs(robust.item() - float(g[f"robust_{s}"])) / abs(float(g[f"robust_{s}"])),
              "grad_q11 rel", T.rel(got, g[f"grad_q11_{s}"]), flush=True)
