"""CPU, world_size 2 over gloo: the multi-GPU host logic of the hot path.

  * AllGatherWithGrad (models.py): per-rank loss / gradients equal the single-process values on the
    concatenated batch once DDP's mean over ranks is applied (SURVEY.md A.2/A.3 contract)
  * iDROLoss._gram: reduce-scatter of column shards + local Gram + all-reduce == Gram of the all-reduced
    [G, P] matrix (what dro_loss.py:232-237 computes); the CUDA Gram kernel is replaced by a torch stand-in
    here because only the sharding / collective logic is under test
  * CoCondenserForPretraining._gather_tensor / gather_tensors (COCO/modeling.py:182-190): the own slot keeps its
    autograd edge, remote slots carry none; with the reference's loss x world scaling the local-row gradients, once
    DDP's mean over ranks is applied, equal the single-process gradient of the full 2BW-span batch
  * DROGreedyLoss.forward (dro_loss.py:49-120): the group statistics are exchanged as one all-reduce of per-group
    sums / counts instead of the reference's two all_gathers of per-sample values (:64-65); h_fun, the EMA losses and
    the EMA counts after two steps equal the oracle fed with the gathered batch
  * iDROLoss.forward end to end (dro_loss.py:216-254) on a small torch model: local group means, rank-summed group
    gradients through the sharded Gram, h_fun after two steps == the oracle that all-reduces the [G, P] matrix
  * gradsync.GradSync: mean-over-ranks gradients whether a layer's flat buffer can be reduced in place during
    backward (one use, no prior gradient) or must be deferred (layer used twice, accumulation, tied parameters)
  * scan.search_sharded: per-rank lists as packed keys, one all-gather, merge with the exactness proof; full-length
    lists with a shard smaller than k, short lists on an even split, and short lists on a corpus clustered by shard
    (proof fails -> guaranteed path) all == the single-process oracle scan, ids bit-exact in (score desc, id asc) order;
    the per-shard scan, the pack and the merge kernels are replaced by numpy stand-ins with the same key layout
"""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _gather_worker(rank, world, port, ret):
    _init(rank, world, port)
    from cocodr_b200 import models
    torch.manual_seed(0)
    B, H = 4, 16
    Q, P = torch.randn(world * B, H), torch.randn(world * B, H)
    q = Q[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    p = P[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    keys = models.gather_with_grad(p)
    tgt = rank * B + torch.arange(B)
    loss = torch.nn.functional.cross_entropy(q @ keys.t(), tgt, reduction="none").mean()
    loss.backward()
    # DDP would average parameter gradients over ranks: emulate on the embedding gradients
    gq, gp = q.grad / world, p.grad / world
    Qf, Pf = Q.clone().requires_grad_(True), P.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(Qf @ Pf.t(), torch.arange(world * B), reduction="none").mean()
    ref.backward()
    ok = (torch.allclose(gq, Qf.grad[rank * B:(rank + 1) * B], atol=1e-6)
          and torch.allclose(gp, Pf.grad[rank * B:(rank + 1) * B], atol=1e-6))
    losses = [torch.zeros(()) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    ok = ok and abs(torch.stack(losses).mean().item() - ref.item()) < 1e-6
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _gram_worker(rank, world, port, ret):
    _init(rank, world, port)
    from cocodr_b200 import dro_loss, kernels

    def gram_cpu(x, gram):
        gram += x @ x.t()
    kernels.gram_f32 = gram_cpu  # stand-in for the CUDA kernel (collective logic under test)
    G, P = 5, 1003  # P not divisible by the world size
    torch.manual_seed(rank)
    local = torch.randn(G, P)
    loss = dro_loss.iDROLoss(types.SimpleNamespace(model_size="base", local_rank=rank), G, 0.25, 0.01, 0.1, 0.05)
    got = loss._gram(local)
    summed = local.clone()
    dist.all_reduce(summed)
    ref = summed @ summed.t()
    ret[rank] = bool(torch.allclose(got, ref, rtol=1e-4, atol=1e-3))
    dist.destroy_process_group()


def _coco_worker(rank, world, port, ret):
    _init(rank, world, port)
    from cocodr_b200 import modeling
    from oracle import heads_ref
    torch.manual_seed(0)
    n, H = 6, 16  # 3 documents = 6 spans per rank
    E = torch.randn(world * n, H)
    e = E[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    me = types.SimpleNamespace(train_args=types.SimpleNamespace(local_rank=rank))
    me._gather_tensor = lambda t: modeling.CoCondenserForPretraining._gather_tensor(me, t)
    all_e = modeling.CoCondenserForPretraining.gather_tensors(me, e)[0]
    ok = torch.equal(all_e.detach(), E)
    loss = heads_ref.coco_contrastive(all_e, world_size=world).mean()  # every rank scores all rows, x world
    loss.backward()
    Ef = E.clone().requires_grad_(True)
    ref = heads_ref.coco_contrastive(Ef, world_size=1).mean()
    ref.backward()
    ok = ok and torch.allclose(e.grad / world, Ef.grad[rank * n:(rank + 1) * n], atol=1e-6)  # DDP's mean over ranks
    ok = ok and abs(loss.item() / world - ref.item()) < 1e-6
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _greedy_worker(rank, world, port, ret):
    _init(rank, world, port)
    from cocodr_b200 import dro_loss, ops
    from oracle import heads_ref
    ops.group_stats = lambda losses, g, n: heads_ref.group_stats(losses, g, n)[:2]  # stand-in for the CUDA kernel
    G, B = 6, 8
    crit = dro_loss.DROGreedyLoss(types.SimpleNamespace(model_size="base", local_rank=rank), G, 0.25, 0.01, 0.1, True)
    crit.train()
    h, sl, cc = torch.ones(G), torch.zeros(G), torch.ones(G)
    ok = True
    for step in range(2):
        gen = torch.Generator().manual_seed(100 + step)
        losses_all = torch.rand(world * B, generator=gen) * 3
        g_all = torch.randint(0, G - 1, (world * B,), generator=gen)  # the last group never appears
        mine = slice(rank * B, (rank + 1) * B)
        robust, gl, gc = crit(losses_all[mine].clone().requires_grad_(True), g_all[mine])
        r_ref, gl_ref, gc_ref, h, sl, cc = heads_ref.dro_greedy_forward(losses_all[mine], g_all[mine], h, sl, cc, G, 0.25,
                                                                        0.01, 0.1, True, gathered=(g_all, losses_all))
        ok = ok and torch.allclose(robust, r_ref, atol=1e-6) and torch.allclose(gl, gl_ref, atol=1e-6)
        ok = ok and torch.equal(gc, gc_ref) and torch.allclose(crit.h_fun, h, atol=1e-6)
        ok = ok and torch.allclose(crit.sum_losses, sl, atol=1e-6) and torch.allclose(crit.count_cat, cc, atol=1e-6)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _idro_worker(rank, world, port, ret):
    _init(rank, world, port)
    from cocodr_b200 import dro_loss, kernels, ops
    from oracle import heads_ref

    def gram_cpu(x, gram):
        gram += x @ x.t()
    kernels.gram_f32 = gram_cpu
    ops.group_stats = lambda losses, g, n: heads_ref.group_stats(losses, g, n)[:2]

    class Net(torch.nn.Module):  # parameter names layer.0 .. layer.11: iDRO selects layer.9 / .10 / .11
        def __init__(self):
            super().__init__()
            self.layer = torch.nn.ModuleList(torch.nn.Linear(8, 8, bias=(i % 2 == 0)) for i in range(12))

        def forward(self, x):
            for lin in self.layer:
                x = torch.tanh(lin(x))
            return x.pow(2).sum(1)

    torch.manual_seed(0)
    net = Net()
    G, B = 5, 6
    crit = dro_loss.iDROLoss(types.SimpleNamespace(model_size="base", local_rank=rank), G, 0.25, 0.01, 0.1, 0.05)
    crit.train()
    h = torch.ones(G)
    names = heads_ref.idro_param_names([n for n, _ in net.named_parameters()])
    params = [dict(net.named_parameters())[n] for n in names]
    ok = len(params) == 4  # layer.9 weight, layer.10 weight + bias, layer.11 weight

    def summed(m):
        dist.all_reduce(m)
        return m

    for step in range(2):
        gen = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(B, 8, generator=gen)
        g = torch.randint(0, G - 1, (B,), generator=gen)
        robust, means, counts = crit(net, net(x), g)
        r_ref, m_ref, c_ref, h = heads_ref.idro_forward(net(x), g, params, h, G, 0.25, 0.1, 0.05, 0.01, all_reduce=summed)
        ok = ok and torch.allclose(robust, r_ref, atol=1e-6) and torch.allclose(means, m_ref, atol=1e-6)
        ok = ok and torch.equal(counts, c_ref) and torch.allclose(crit.h_fun, h, rtol=1e-5, atol=1e-7)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _scan_worker(rank, world, port, ret):
    _init(rank, world, port)
    import numpy as np
    from cocodr_b200 import scan
    from oracle import scan_ref

    def flip(sc):
        u = np.asarray(sc, dtype=np.float32).view(np.uint32).astype(np.uint64)
        return np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))

    def shard_keys_cpu(Q, P, ks, doc_base):  # stand-in for search_async + cdr_topk_pack (same key layout)
        k_eff = min(ks, P.shape[0])
        D, I = scan_ref.search(Q, P, k_eff)
        n_q = Q.shape[0]
        keys = np.zeros((n_q, ks), dtype=np.uint64)
        ids = (np.asarray(I, dtype=np.int64) + doc_base).astype(np.uint64)
        keys[:, :k_eff] = (flip(D) << np.uint64(32)) | (~ids & np.uint64(0xFFFFFFFF))
        out = np.concatenate([keys.reshape(-1), np.array([0, int(k_eff == P.shape[0])], dtype=np.uint64)])
        return torch.from_numpy(out.view(np.int64))

    def merge_cpu(allk, W, n_q, ks, k):  # stand-in for cdr_topk_merge_keys, incl. the exactness proof
        a = allk.numpy().view(np.uint64).reshape(W, n_q * ks + 2)
        flag = int(a[:, n_q * ks].sum())
        D, I = np.full((n_q, k), -np.inf, dtype=np.float32), np.full((n_q, k), -1, dtype=np.int64)
        for q in range(n_q):
            cand = np.sort(np.concatenate([a[w, q * ks:(q + 1) * ks] for w in range(W)]))[::-1][:k]
            kth = cand[k - 1] if len(cand) >= k else np.uint64(0)
            for w in range(W):
                last = a[w, q * ks + ks - 1]
                if a[w, n_q * ks + 1] == 0 and last != 0 and last > kth:
                    flag += 1
            live = cand[cand != 0]
            u = (live >> np.uint64(32)).astype(np.uint32)
            u = np.where(u & np.uint32(0x80000000), u & np.uint32(0x7FFFFFFF), ~u)
            D[q, :len(live)] = u.view(np.float32)
            I[q, :len(live)] = (~live & np.uint64(0xFFFFFFFF)).astype(np.int64)
        return torch.from_numpy(D), torch.from_numpy(I), torch.tensor([flag], dtype=torch.int32)

    def search_cpu(Q, P, k, doc_base=0, force_exhaustive=False):  # the per-shard scan of the guaranteed path
        k_eff = min(k, P.shape[0])
        D, I = scan_ref.search(Q, P, k_eff)
        return torch.from_numpy(np.asarray(D, dtype=np.float32)), torch.from_numpy(np.asarray(I, dtype=np.int64)) + doc_base

    def merge_topk_cpu(D, I, k):  # (score desc, id asc), ids < 0 = empty slots
        outD, outI = torch.empty(D.shape[0], k), torch.empty(D.shape[0], k, dtype=torch.int64)
        for r in range(D.shape[0]):
            live = [(-float(d), int(i)) for d, i in zip(D[r].tolist(), I[r].tolist()) if i >= 0]
            live.sort()
            outD[r] = torch.tensor([-d for d, _ in live[:k]])
            outI[r] = torch.tensor([i for _, i in live[:k]])
        return outD, outI

    scan._shard_keys, scan._merge_gathered, scan.search, scan.merge_topk = shard_keys_cpu, merge_cpu, search_cpu, merge_topk_cpu
    k, n_docs = 40, 1000
    Q, P = scan_ref.synth_corpus(n_docs, 7, 64, seed=11, kind="exact")  # exact arithmetic, many ties
    Dr, Ir = scan_ref.search(Q, P, k)
    ok = True
    # (a) full-length lists; rank 1 holds fewer documents than k: its list is padded with empty slots
    cut = n_docs - 25
    lo, hi = (0, cut) if rank == 0 else (cut, n_docs)
    D, I = scan.search_sharded(Q, P[lo:hi], k, doc_base=lo)
    ok = ok and bool((I.numpy() == np.asarray(Ir)).all() and (D.numpy() == np.asarray(Dr, dtype=np.float32)).all())
    # (b) short lists (24 < k per shard), even split: the proof holds or the search silently repeats with full lists
    scan.shard_list_len = lambda k_, w_: 24
    lo, hi = (0, n_docs // 2) if rank == 0 else (n_docs // 2, n_docs)
    D, I = scan.search_sharded(Q, P[lo:hi], k, doc_base=lo)
    ok = ok and bool((I.numpy() == np.asarray(Ir)).all() and (D.numpy() == np.asarray(Dr, dtype=np.float32)).all())
    # (c) short lists on a corpus CLUSTERED by shard (every query's best documents sit on rank 0): the proof must fail
    #     and the fallback must still return the exact answer
    order = np.argsort(-(np.asarray(Q[:1], dtype=np.float32) @ np.asarray(P, dtype=np.float32).T)[0], kind="stable")
    Ps = P[torch.from_numpy(order.copy())]
    Drs, Irs = scan_ref.search(Q[:1], Ps, k)
    D, I = scan.search_sharded(Q[:1], Ps[lo:hi], k, doc_base=lo)
    ok = ok and bool((I.numpy() == np.asarray(Irs)).all() and (D.numpy() == np.asarray(Drs, dtype=np.float32)).all())
    ret[rank] = ok
    dist.destroy_process_group()


def _gradsync_worker(rank, world, port, ret):
    """GradSync must give mean-over-ranks gradients in every autograd pattern, not only "each layer once, no prior
    gradient": a toy Function with the same flat-buffer protocol as ops.BertLayerFn (views of one zero-filled buffer,
    ops._note_forward / ops._submit) is driven through (1) one use per layer, (2) the same layer twice in one forward
    (unfused q / p towers), (3) two backward passes accumulating into .grad, (4) a parameter that already received a
    gradient from an op outside the protocol (the tied word embedding of the COCO model)."""
    _init(rank, world, port)
    from cocodr_b200 import ops
    from cocodr_b200.gradsync import GradSync

    class ToyLayer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b):
            ctx.save_for_backward(x, w)
            ctx.params = ops._note_forward(ctx, (w, b))
            return torch.tanh(x @ w.t() + b)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            y = torch.tanh(x @ w.t() + ctx.params[1])
            dz = dy * (1 - y * y)
            flat = torch.zeros(w.numel() + w.shape[0])
            dw, db = flat[:w.numel()].view_as(w), flat[w.numel():]
            dw += dz.t() @ x
            db += dz.sum(0)
            ops._submit(ctx, flat)
            return dz @ w, dw, db

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.ParameterList(torch.nn.Parameter(torch.randn(6, 6) * 0.5) for _ in range(3))
            self.b = torch.nn.ParameterList(torch.nn.Parameter(torch.randn(6) * 0.1) for _ in range(3))
            self.head = torch.nn.Linear(6, 1)

        def tower(self, x):
            for w, b in zip(self.w, self.b):
                x = ToyLayer.apply(x, w, b)
            return x

    torch.manual_seed(0)
    net = Net()
    ref_net = Net()
    ref_net.load_state_dict(net.state_dict())
    gen = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(5, 6, generator=gen) for _ in range(4)]

    def losses(n):
        return {
            "once": lambda: n.head(n.tower(xs[0])).sum(),
            "twice": lambda: (n.head(n.tower(xs[0])) * n.tower(xs[1]).sum(1, keepdim=True)).sum(),
            "tied": lambda: (n.head(n.tower(xs[0])).sum() + (n.w[0] ** 2).sum() * 0.1),
        }

    def reference(names, n_backward):
        ref_net.zero_grad(set_to_none=True)
        for _ in range(n_backward):
            for nm in names:
                losses(ref_net)[nm]().backward()
        out = {}
        for k, p in ref_net.named_parameters():
            g = p.grad.clone()
            dist.all_reduce(g)
            out[k] = g / world
        return out

    sync = GradSync(net)
    ok = True
    checks = {"once": (["once"], 1, True), "twice": (["twice"], 1, False), "accumulate": (["once"], 2, None),
              "tied": (["tied"], 1, None)}
    for label, (names, n_backward, expect_overlap) in checks.items():
        net.zero_grad(set_to_none=True)
        sync.stats = {"overlapped": 0, "deferred": 0}
        for _ in range(n_backward):
            for nm in names:
                loss = losses(net)[nm]()
                with sync:
                    loss.backward()
        want = reference(names, n_backward)
        for k, p in net.named_parameters():
            ok = ok and torch.allclose(p.grad, want[k], rtol=1e-5, atol=1e-6)
        if expect_overlap is True:
            ok = ok and sync.stats["overlapped"] == 3 and sync.stats["deferred"] == 0
        if expect_overlap is False:
            ok = ok and sync.stats["overlapped"] == 0 and sync.stats["deferred"] == 6
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("worker,port", [(_gather_worker, 29641), (_gram_worker, 29643), (_scan_worker, 29645),
                                         (_coco_worker, 29647), (_greedy_worker, 29649),
                                         (_idro_worker, 29651), (_gradsync_worker, 29653)])
def test_world2_gloo(worker, port):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)
