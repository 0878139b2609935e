"""torchrun worker for the multi-GPU parity checks (launched by tests/test_multigpu_gpu.py).

Every rank checks, against a single-process computation on the concatenated inputs:
  1. sharded corpus scan (documents split over ranks, all-gather of [nq,k] lists, merge) -> bit-exact ranks
  2. in-batch InfoNCE over all-gathered passage CLS embeddings: loss and (DDP-averaged) embedding gradients
  3. iDRO Gram through reduce-scatter + local cdr_gram_f32 + all-reduce
  4. COCO contrastive loss with the reference's gather convention (own slot keeps the graph, loss * world)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))  # a mismatch must fail fast
    from cocodr_b200 import dro_loss, models, ops, scan
    from oracle import heads_ref, scan_ref

    # 1. sharded scan
    Q, P = scan_ref.synth_corpus(40_000, 32, 256, seed=21, kind="exact")
    per = P.shape[0] // world
    lo, hi = rank * per, (P.shape[0] if rank == world - 1 else (rank + 1) * per)
    D, I = scan.search_sharded(Q.to(dev), P[lo:hi].to(dev), 100, doc_base=lo)
    Dr, Ir = scan_ref.search(Q, P, 100)
    assert (I.cpu().numpy() == Ir).all() and (D.cpu().numpy() == Dr).all(), "sharded scan mismatch"

    # 2. in-batch InfoNCE with gathered passages
    torch.manual_seed(0)
    B, H = 8, 64
    Qa, Pa = torch.randn(world * B, H) * 0.3, torch.randn(world * B, H) * 0.3
    q = Qa[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
    p = Pa[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
    keys = models.gather_with_grad(p)
    loss = ops.qp_infonce(q, keys, row_offset=rank * B).mean()
    loss.backward()
    Qf, Pf = Qa.clone().requires_grad_(True), Pa.clone().requires_grad_(True)
    ref = heads_ref.qp_infonce(Qf, Pf).mean()
    ref.backward()
    np.testing.assert_allclose((q.grad / world).cpu().numpy(), Qf.grad[rank * B:(rank + 1) * B].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose((p.grad / world).cpu().numpy(), Pf.grad[rank * B:(rank + 1) * B].numpy(), rtol=1e-4, atol=1e-6)
    lall = torch.zeros(world, device=dev)
    lall[rank] = loss.detach()
    dist.all_reduce(lall)
    assert abs(lall.mean().item() - ref.item()) < 1e-5

    # 3. iDRO Gram: reduce-scatter shards + local Gram + all-reduce == Gram of the all-reduced matrix
    G, Pn = 7, 100_003
    g = torch.Generator().manual_seed(rank)
    local_m = torch.randn(G, Pn, generator=g).to(dev)
    mod = dro_loss.iDROLoss(types.SimpleNamespace(model_size="base", local_rank=local), G, 0.25, 0.01, 0.1, 0.05)
    got = mod._gram(local_m)
    summed = local_m.clone()
    dist.all_reduce(summed)
    refg = (summed.double() @ summed.double().t()).float()
    # (the local Gram runs on tcgen05 kind::tf32: operands keep 10 mantissa bits -> ~1e-3 of the row norms)
    nrm = torch.sqrt(torch.diagonal(refg)).cpu().numpy()
    assert (np.abs(got.cpu().numpy() - refg.cpu().numpy()) / (nrm[:, None] * nrm[None, :])).max() < 2e-3

    # 4. COCO contrastive, reference gather convention
    n_loc = 6
    E = torch.randn(world * n_loc, H) * 0.4
    e = E[rank * n_loc:(rank + 1) * n_loc].to(dev).requires_grad_(True)
    parts = [torch.empty_like(e) for _ in range(world)]
    dist.all_gather(parts, e.detach())
    parts[rank] = e
    co = ops.coco_contrastive(torch.cat(parts), loss_scale=float(world))
    co.mean().backward()
    Ef = E.clone().requires_grad_(True)
    refc = heads_ref.coco_contrastive(Ef)
    refc.mean().backward()
    np.testing.assert_allclose(co.detach().cpu().numpy(), world * refc.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose((e.grad / world).cpu().numpy(), Ef.grad[rank * n_loc:(rank + 1) * n_loc].numpy(),
                               rtol=1e-3, atol=1e-6)
    # 5. GradSync: per-layer flat-buffer all-reduce during backward == mean over ranks of the local gradients
    from transformers import BertConfig

    from cocodr_b200.gradsync import GradSync
    from oracle import bert_ref
    tiny = dict(hidden=128, layers=3, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)
    hf = BertConfig(vocab_size=2000, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=512,
                    max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_labels=2)
    model = models.BertDot_InBatch_NLL_LN(hf)
    model.bert.load_state_dict(bert_ref.synth_state(tiny, 0), strict=False)
    model = model.to(dev).train()
    qi, qm = (t.to(dev) for t in bert_ref.synth_batch(4, 32, 2000, 100 + rank))
    pi, pm = (t.to(dev) for t in bert_ref.synth_batch(4, 32, 2000, 200 + rank))
    w = torch.ones(4, device=dev)
    model(qi, qm, pi, pm, weights=w)[0].backward()
    local_g = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for t in local_g.values():
        dist.all_reduce(t)
        t /= world
    model.zero_grad(set_to_none=True)
    sync = GradSync(model)
    loss = model(qi, qm, pi, pm, weights=w)[0]
    with sync:
        loss.backward()
    torch.cuda.synchronize()
    assert len(local_g) > 50
    for n, p in model.named_parameters():
        if n in local_g:
            ref_g = local_g[n]
            err = (p.grad - ref_g).abs().max().item()
            assert err <= 1e-5 + 1e-3 * ref_g.abs().max().item(), (n, err)

    # 6. peer-memory exchange (cdr_ln_fwd_push / cdr_peer_*): the last LayerNorm pushes the passage CLS rows into
    #    every rank's symmetric buffer and the gradients come back the same way == the NCCL all-gather path
    model.zero_grad(set_to_none=True)
    l_nccl = model(qi, qm, pi, pm, weights=w)[0]
    l_nccl.backward()
    g_nccl = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.enable_peer_gather(True)
    for rep in range(3):  # several rounds: epochs advance, buffers are reused
        model.zero_grad(set_to_none=True)
        l_peer = model(qi, qm, pi, pm, weights=w)[0]
        l_peer.backward()
        torch.cuda.synchronize()
        assert model._xchg is not None and model._xchg.epoch == rep + 1
        assert abs(l_peer.item() - l_nccl.item()) <= 1e-5 * max(1.0, abs(l_nccl.item())), (l_peer.item(), l_nccl.item())
        for n, p in model.named_parameters():
            if n in g_nccl:
                err = (p.grad - g_nccl[n]).abs().max().item()
                assert err <= 1e-6 + 2e-3 * g_nccl[n].abs().max().item(), (n, err)
        dist.barrier()
    model.enable_peer_gather(False)

    # 7. GradSync when a layer's flat gradient buffer must NOT be reduced in place during backward: (a) unfused towers
    #    (q_len != p_len, the reference's default 64 / 128: every layer runs twice per backward), (b) gradient
    #    accumulation over two backward passes.  Reference: all-reduce of the locally accumulated gradients.
    qs, qsm = (t.to(dev) for t in bert_ref.synth_batch(4, 16, 2000, 300 + rank))

    def two_steps(use_sync):
        model.zero_grad(set_to_none=True)
        s_ = GradSync(model) if use_sync else None
        for rep in range(2):
            loss = model(qs, qsm, pi, pm, weights=w)[0]  # 16-token queries, 32-token passages: towers run separately
            if s_ is not None:
                with s_:
                    loss.backward()
            else:
                loss.backward()
        torch.cuda.synchronize()
        return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, s_

    ref_g, _ = two_steps(False)
    for t in ref_g.values():
        dist.all_reduce(t)
        t /= world
    got_g, s_ = two_steps(True)
    assert s_.stats["overlapped"] == 0 and s_.stats["deferred"] > 0, s_.stats  # second backward of two: all deferred
    for n, t in ref_g.items():
        err = (got_g[n] - t).abs().max().item()
        assert err <= 1e-5 + 2e-3 * t.abs().max().item(), (n, err)

    # 8. iDRO on the in-batch head at world > 1: ranks hold DIFFERENT groups, so per-group partial backwards through the
    #    gathered keys would issue mismatched collectives; the group gradients come from collective-free loss views
    #    (models.BertDot_InBatch_NLL_LN.idro_group_grads).  Must not hang or produce non-finite weights / gradients.
    tiny12 = dict(tiny, layers=12)  # iDRO differentiates layers 9-11 (dro_loss.py:176-190)
    hf12 = BertConfig(vocab_size=2000, hidden_size=128, num_hidden_layers=12, num_attention_heads=2, intermediate_size=512,
                      max_position_embeddings=64, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, num_labels=2)
    m12 = models.BertDot_InBatch_NLL_LN(hf12)
    m12.bert.load_state_dict(bert_ref.synth_state(tiny12, 0), strict=False)
    m12 = m12.to(dev).train()
    for mode in ("own-pair", "local-batch"):
        m12.idro_group_grads = mode
        m12.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=local), 6, "idro", 0.25, 0.01, 0.1, 0.05)
        gid = torch.tensor([0, 1, 1, 2] if rank == 0 else [3, 3, 4, 0], device=dev)
        for rep in range(2):
            m12.zero_grad(set_to_none=True)
            robust = m12(qi, qm, pi, pm, group_ids=gid, weights=w)[0]
            s2 = GradSync(m12)
            with s2:
                robust.backward()
        torch.cuda.synchronize()
        # (h_fun is per rank: the reference mixes LOCAL group means / masks with the rank-summed gradients)
        h = m12.loss.h_fun.clone()
        assert torch.isfinite(h).all() and abs(h.sum().item() - 1.0) < 0.2, (mode, h)
        assert all(torch.isfinite(p.grad).all() for p in m12.parameters() if p.grad is not None)

    # 9. gradient exchange fused into the optimizer (peeropt.PeerArena + cdr_adam_multi_peer): parameters, shadows and
    #    gradient buffers in one symmetric arena, every rank updates its chunks from all ranks' gradients and stores
    #    the result everywhere == GradSync's NCCL all-reduce + the same AdamW on every rank.  Three steps, fused and
    #    unfused towers (the unfused step leaves its gradients outside the arena -> NCCL + local update), then the
    #    sharded optimizer state is consolidated and compared as well.
    import copy

    from cocodr_b200 import optim, peeropt

    def make():
        m = models.BertDot_InBatch_NLL_LN(hf)
        m.bert.load_state_dict(bert_ref.synth_state(tiny, 0), strict=False)
        return m.to(dev).train()

    def run(m, opt, sync, steps):
        for q_ids, q_mask in steps:
            opt.zero_grad(set_to_none=True)
            loss = m(q_ids, q_mask, pi, pm, weights=w)[0]
            with sync:
                loss.backward()
            opt.step()
        torch.cuda.synchronize()

    def worst(a_, b_):
        out = (0.0, "")
        for (n, x), (_, y) in zip(a_.named_parameters(), b_.named_parameters()):
            if x.grad is None and y.grad is None:
                continue
            e = (x - y).abs().max().item() / max(y.abs().max().item(), 1e-3)
            if e > out[0]:
                out = (e, n)
        return out

    steps = [(qi, qm), (qi, qm), (qs, qsm), (qi, qm)]
    m_ref, m_ref2, m_peer = make(), make(), make()
    o_ref = optim.AdamW(m_ref.parameters(), lr=1e-3, weight_decay=0.01).attach_shadows(m_ref)
    o_ref2 = optim.AdamW(m_ref2.parameters(), lr=1e-3, weight_decay=0.01).attach_shadows(m_ref2)
    o_peer = optim.AdamW(m_peer.parameters(), lr=1e-3, weight_decay=0.01)
    arena = peeropt.PeerArena(m_peer, o_peer)
    s_ref, s_ref2, sync_p = GradSync(m_ref), GradSync(m_ref2), GradSync(m_peer, arena=arena)
    # one step: the update is a smooth function of the gradients -> tight agreement
    for m_, o_, s_x in ((m_ref, o_ref, s_ref), (m_ref2, o_ref2, s_ref2), (m_peer, o_peer, sync_p)):
        run(m_, o_, s_x, steps[:1])
    e1 = worst(m_peer, m_ref)
    assert e1[0] < 2e-5, ("peer adam, first step", e1)
    # more steps: Adam's normalised update amplifies the run-to-run noise of the atomically accumulated gradients, so
    # two runs of the SAME (NCCL) path drift apart; the peer path must stay as close to the reference as that
    for m_, o_, s_x in ((m_ref, o_ref, s_ref), (m_ref2, o_ref2, s_ref2), (m_peer, o_peer, sync_p)):
        run(m_, o_, s_x, steps[1:])
    arena.check()
    assert sync_p.stats.get("peer", 0) > 0 and sync_p.stats.get("peer_copied", 0) > 0, sync_p.stats
    e_rr, e_pr = worst(m_ref2, m_ref), worst(m_peer, m_ref)
    assert e_pr[0] <= 3.0 * e_rr[0] + 1e-5, ("peer adam", e_pr, "reference vs itself", e_rr)
    ref_p = dict(m_ref.named_parameters())
    # every rank holds the same parameters and the same fp16 shadows
    flat_p = torch.cat([p_.detach().reshape(-1) for p_ in m_peer.parameters()])
    lo_, hi_ = flat_p.clone(), flat_p.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_, hi_), "parameters differ across ranks"
    sh_ref, sh_peer = m_ref.bert._shadows[1], m_peer.bert._shadows[1]
    assert arena.contains(sh_peer.wqkv) and (sh_peer.wqkv.float() - sh_ref.wqkv.float()).abs().max().item() < 1e-3
    # clipped step (torch.nn.utils.clip_grad_norm_ semantics: norm of the rank-averaged gradient): two-phase peer path
    # vs NCCL all-reduce + cdr_grad_sqnorm_multi / clip coefficient, from identical weights
    with torch.no_grad():
        for p_, pr in zip(m_peer.parameters(), m_ref.parameters()):
            pr.copy_(p_)
    for st_ in (o_ref.state, o_peer.state):
        for v_ in st_.values():
            v_["exp_avg"].zero_()
            v_["exp_avg_sq"].zero_()
    norms = []
    for m_, o_, s_x in ((m_ref, o_ref, s_ref), (m_peer, o_peer, sync_p)):
        o_.zero_grad(set_to_none=True)
        loss = m_(qi, qm, pi, pm, weights=w)[0]
        with s_x:
            loss.backward()
        norms.append(o_.clip_grad_norm_(0.01))
        o_.step()
    torch.cuda.synchronize()
    arena.check()
    n_ref, n_peer = norms[0].item(), norms[1].item()
    assert n_ref > 0.01 and abs(n_peer - n_ref) <= 1e-4 * n_ref, ("clip norm", n_peer, n_ref)
    e_clip = worst(m_peer, m_ref)
    assert e_clip[0] < 5e-5, ("peer adam, clipped step", e_clip)
    arena.consolidate_state(o_peer)
    for (n, p_), (_, pr) in zip(m_peer.named_parameters(), m_ref.named_parameters()):
        if pr.grad is None or not arena.contains(p_.grad):
            continue
        a_, b_ = o_peer.state[p_]["exp_avg"], o_ref.state[pr]["exp_avg"]
        assert (a_ - b_).abs().max().item() <= 1e-7 + 5e-2 * b_.abs().max().item(), ("exp_avg", n)  # (same drift)
    del copy

    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
