#!/usr/bin/env python
"""Print the forward / gradient deviations of the tiny ANCE model against the reference fixture (the numbers
tests/test_model_gpu.py::test_ance_tiny_forward_backward_matches_reference asserts on).  Run from a tree root."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_model_gpu as T
g = np.load(os.path.join("tests", "golden", "ance_tiny.npz"))
m = T.build(T.TINY); m.train()
q, mq, a, ma, b, mb = T.triplet(T.TINY, int(g["B"]), int(g["L"]), int(g["seed"]))
w = torch.from_numpy(g["weights"]).cuda()
loss, acc, logits = m(q, mq, a, ma, b, mb, weights=w)
print("logits rel", T.rel(logits.detach().cpu().numpy(), g["logits"]), "loss", loss.item(), float(g["erm_loss"]))
with torch.no_grad():
    print("q_emb rel", T.rel(m.query_emb(q, mq).cpu().numpy(), g["q_emb"]), "b_emb rel", T.rel(m.body_emb(b, mb).cpu().numpy(), g["b_emb"]))
m.zero_grad(); loss.backward()
named = dict(m.bert.named_parameters())
res = []
for key in g.files:
    if not key.startswith("grad."): continue
    name = key[5:]
    if name.endswith(".rownorm"): got = named[name[:-8]].grad.norm(dim=1).cpu().numpy()
    elif name.endswith(".norm"): got = named[name[:-5]].grad.norm().item()
    else: got = named[name].grad.cpu().numpy()
    if np.abs(g[key]).max() < 1e-5: continue
    res.append((T.rel(got, g[key], floor=1e-5), key))
res.sort(reverse=True)
for r, k in res[:8]: print(f"{r:.4f} {k}")
print("mean rel", np.mean([r for r, _ in res]))
