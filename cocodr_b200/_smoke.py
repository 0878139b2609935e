"""__graft_entry__.smoke(): one tiny invocation of the hot path on cuda:0, checked against the CPU oracle
(oracle/ is imported here only as the checker)."""
import torch


def run():
    from transformers import BertConfig

    from oracle import bert_ref, heads_ref, scan_ref

    from . import _lib, kernels, models, ops, optim, scan

    _lib.check(_lib.load().cdr_device_check(), "cdr_device_check")
    torch.cuda.set_device(0)
    cfg = dict(hidden=128, layers=4, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)
    hf = BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                    num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                    max_position_embeddings=cfg["max_pos"], type_vocab_size=2, hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0, num_labels=2)
    m = models.BertDot_InBatch_NLL_LN(hf)
    st = bert_ref.synth_state(cfg, 0)
    m.bert.load_state_dict(st, strict=False)
    m = m.cuda().train()
    B, L = 4, 32
    q, mq = bert_ref.synth_batch(B, L, cfg["vocab"], 1)
    p, mp = bert_ref.synth_batch(B, L, cfg["vocab"], 2)
    n0 = kernels.launches
    loss, acc, logits = m(q.cuda(), mq.cuda(), p.cuda(), mp.cuda(), weights=torch.ones(B, device="cuda"))
    loss.backward()
    torch.cuda.synchronize()
    # oracle: fp32 CPU restatement of the same step
    leaf = {k: v.clone().requires_grad_(True) for k, v in st.items()}
    qe, pe = bert_ref.cls_embedding(leaf, q, mq, cfg), bert_ref.cls_embedding(leaf, p, mp, cfg)
    ref = heads_ref.qp_infonce(qe, pe).mean()
    ref.backward()
    err = abs(loss.item() - ref.item()) / abs(ref.item())
    gname = "encoder.layer.3.output.dense.weight"
    got = dict(m.bert.named_parameters())[gname].grad.cpu()
    # (the InfoNCE gradient is a difference of near-identical CLS vectors for a random-init encoder, so the
    # bound on the max error is loose; direction is what must agree)
    gerr = ((got - leaf[gname].grad).abs().max() / leaf[gname].grad.abs().max()).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), leaf[gname].grad.flatten(), dim=0).item()
    assert err < 1e-2, f"loss {loss.item()} vs oracle {ref.item()}"
    assert gerr < 0.25 and cos > 0.99, f"grad rel err {gerr}, cosine {cos}"
    # one fused optimizer step (cdr_adam_multi): parameters move, fp16 operand shadows follow in the same launch
    opt = optim.AdamW([t for t in m.parameters() if t.requires_grad], lr=1e-3, eps=1e-8, semantics="torch").attach_shadows(m)
    w_before = dict(m.bert.named_parameters())[gname].detach().clone()
    opt.step()
    torch.cuda.synchronize()
    assert (dict(m.bert.named_parameters())[gname] - w_before).abs().max().item() > 1e-5
    for param, (dst, is_f32) in m.bert.shadow_map().items():
        want = param.detach() if is_f32 else param.detach().half()
        assert torch.equal(dst.view_as(want), want)
    # corpus scan, exact-arithmetic corpus => bit-exact ranks
    Q, P = scan_ref.synth_corpus(20000, 16, 128, seed=5, kind="exact")
    D, I = scan.search(Q.cuda(), P.cuda(), 10)
    Dr, Ir = scan_ref.search(Q, P, 10)
    assert (I.cpu().numpy() == Ir).all() and (D.cpu().numpy() == Dr).all()
    print(f"smoke ok: loss {loss.item():.5f} (oracle {ref.item():.5f}, rel {err:.2e}), grad rel {gerr:.2e} cos {cos:.5f}, "
          f"scan ranks bit-exact, {kernels.launches - n0} kernel launches, grad_scale {ops.get_grad_scale()}")
