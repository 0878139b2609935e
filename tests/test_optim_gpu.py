"""GPU: fused multi-tensor AdamW / Lamb / gradient clipping (cdr_adam_multi, cdr_lamb_multi, cdr_grad_*) against
torch.optim.AdamW and the oracle restatements of the reference's optimizers (oracle/optim_ref.py)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(768, 768), (3072, 768), (768,), (7,), (1000, 3), (30522, 64)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.05).cuda()) for s in SHAPES]


def _grads(params, seed):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        p.grad = (torch.randn(*p.shape, generator=g) * 0.01).cuda()


def test_adamw_torch_semantics_matches_torch():
    from cocodr_b200 import optim
    ours, ref = _params(1), _params(1)
    o = optim.AdamW(ours, lr=3e-4, eps=1e-8, weight_decay=0.01, semantics="torch")
    r = torch.optim.AdamW(ref, lr=3e-4, eps=1e-8, weight_decay=0.01)
    for step in range(4):
        _grads(ours, 10 + step)
        _grads(ref, 10 + step)
        o.step()
        r.step()
    for a, b in zip(ours, ref):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
    sd = o.state_dict()  # torch layout: per-parameter step / exp_avg / exp_avg_sq
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 4.0


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adamw_hf_semantics_and_lamb_match_oracle(wd):
    from cocodr_b200 import optim
    from oracle import optim_ref
    for kind in ("hf", "lamb"):
        ours = _params(2)
        ref = [p.detach().clone() for p in ours]
        ms, vs = [torch.zeros_like(p) for p in ref], [torch.zeros_like(p) for p in ref]
        o = (optim.AdamW(ours, lr=1e-3, eps=1e-6, weight_decay=wd) if kind == "hf"
             else optim.Lamb(ours, lr=1e-3, eps=1e-6, weight_decay=wd))
        for step in range(1, 4):
            _grads(ours, 20 + step)
            trusts = []
            for p, q, m, v in zip(ours, ref, ms, vs):
                if kind == "hf":
                    optim_ref.hf_adamw_step(q, p.grad.clone(), m, v, step, lr=1e-3, eps=1e-6, weight_decay=wd)
                else:
                    trusts.append(optim_ref.lamb_step(q, p.grad.clone(), m, v, lr=1e-3, eps=1e-6, weight_decay=wd))
            o.step()
            if kind == "lamb":
                got = [float(o.state[p]["trust_ratio"]) for p in ours]
                for a, b in zip(got, trusts):
                    assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
        for a, b in zip(ours, ref):
            assert (a - b).abs().max().item() <= 5e-6 * max(1.0, b.abs().max().item()), kind


def test_clip_coefficient_is_folded_into_the_step():
    from cocodr_b200 import optim
    from oracle import optim_ref
    ours, ref = _params(3), _params(3)
    o = optim.AdamW(ours, lr=1e-3, eps=1e-8, semantics="torch")
    r = torch.optim.AdamW(ref, lr=1e-3, eps=1e-8, weight_decay=0.0)
    _grads(ours, 30)
    _grads(ref, 30)
    coef, total = optim_ref.clip_coef([p.grad for p in ref], 0.05)
    assert coef < 1.0
    n = o.clip_grad_norm_(0.05)
    torch.nn.utils.clip_grad_norm_(ref, 0.05)
    o.step()
    r.step()
    assert abs(float(n) - total) <= 1e-4 * total
    for a, b in zip(ours, ref):
        assert (a - b).abs().max().item() <= 5e-6 * max(1.0, b.abs().max().item())


def test_optimizer_refreshes_encoder_shadows_and_graph_step():
    """attach_shadows: the fp16 operand copies follow the parameters without a separate cast; two eager steps with
    the fused optimizer == two steps with torch AdamW + the encoder's own cast."""
    from transformers import BertConfig
    from cocodr_b200 import models, optim
    cfg = BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     vocab_size=500, max_position_embeddings=64, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, num_labels=2)
    torch.manual_seed(0)
    ma = models.BertDot_InBatch_NLL_LN(cfg).cuda().train()
    mb = copy.deepcopy(ma)
    oa = optim.AdamW([p for p in ma.parameters() if p.requires_grad], lr=1e-3, eps=1e-8, weight_decay=0.01,
                     semantics="torch").attach_shadows(ma)
    ob = torch.optim.AdamW([p for p in mb.parameters() if p.requires_grad], lr=1e-3, eps=1e-8, weight_decay=0.01)
    assert oa.manages_shadows
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(5, 500, (8, 32), generator=g).cuda()
    mask = torch.ones_like(ids)
    w = torch.ones(4, device="cuda")
    for _ in range(2):
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad(set_to_none=True)
            loss = m(ids[:4], mask[:4], ids[4:], mask[4:], weights=w)[0]
            loss.backward()
            o.step()
    la = ma(ids[:4], mask[:4], ids[4:], mask[4:], weights=w)[0]
    lb = mb(ids[:4], mask[:4], ids[4:], mask[4:], weights=w)[0]
    assert abs(la.item() - lb.item()) <= 1e-2 * abs(lb.item())  # Adam amplifies the noise of near-zero gradients
    # the operand shadows ARE the updated parameters (fp16 / packed fp32), with no cast launched by the encoder;
    # (parameters themselves are not compared across the two runs: Adam turns the run-to-run noise of
    # near-zero gradients -- split-K atomics -- into O(lr) differences)
    n_checked = 0
    for param, (dst, is_f32) in ma.bert.shadow_map().items():
        want = param.detach() if is_f32 else param.detach().half()
        assert torch.equal(dst.view_as(want), want)
        n_checked += 1
    assert n_checked == 2 * 9


def test_graphed_step_follows_host_lr_changes():
    """A captured step reads lr from a device scalar; GraphedTrainStep mirrors scheduler-driven host changes before
    every replay (lr = 0 freezes the parameters, a non-zero lr moves them again)."""
    from transformers import BertConfig
    from cocodr_b200 import models, optim
    from cocodr_b200.graph import GraphedTrainStep
    cfg = BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     vocab_size=500, max_position_embeddings=64, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, num_labels=2)
    torch.manual_seed(0)
    m = models.BertDot_InBatch_NLL_LN(cfg).cuda().train()
    opt = optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, eps=1e-8, semantics="torch").attach_shadows(m)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(5, 500, (8, 32), generator=g).cuda()
    mask = torch.ones_like(ids)
    w = torch.ones(4, device="cuda")
    step = GraphedTrainStep(m, opt, (ids[:4], mask[:4], ids[4:], mask[4:], None, None, True, None, w))
    probe = m.bert.encoder.layer[0].output.dense.weight
    step(ids[:4], mask[:4], ids[4:], mask[4:])
    torch.cuda.synchronize()
    before = probe.detach().clone()
    opt.param_groups[0]["lr"] = 0.0
    step(ids[:4], mask[:4], ids[4:], mask[4:])
    torch.cuda.synchronize()
    assert torch.equal(probe.detach(), before)
    opt.param_groups[0]["lr"] = 1e-3
    step(ids[:4], mask[:4], ids[4:], mask[4:])
    torch.cuda.synchronize()
    assert (probe.detach() - before).abs().max().item() > 1e-5


def test_lamb_matches_reference_fixture(golden_dir):
    """cdr_lamb_multi vs the UNMODIFIED reference Lamb (fixture tests/golden/lamb_tiny.npz): parameters after 3
    steps and the last trust ratios, with and without weight decay, including an all-zero tensor (trust ratio 1)."""
    import os
    import numpy as np
    from cocodr_b200 import optim
    from oracle.make_golden import lamb_inputs
    g = np.load(os.path.join(golden_dir, "lamb_tiny.npz"))
    for tag, wd in (("wd0", 0.0), ("wd01", 0.01)):
        params, grads = lamb_inputs(int(g["seed"]))
        ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
        opt = optim.Lamb(ps, lr=1e-3, eps=1e-6, weight_decay=wd)
        for step in range(3):
            for p, gr in zip(ps, grads[step]):
                p.grad = gr.clone().cuda()
            opt.step()
        for i, p in enumerate(ps):
            np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"{tag}.p{i}"], rtol=2e-5, atol=1e-7)
            t = float(opt.state[p]["trust_ratio"])
            assert abs(t - float(g[f"{tag}.trust{i}"])) <= 1e-4 * max(1.0, abs(t))


@pytest.mark.parametrize("kind", ["adamw", "lamb"])
def test_resume_from_reference_format_state_dict(kind):
    """The reference's optimizers store ``state['step']`` as a python int (ANCE/utils/lamb.py:93, transformers AdamW)
    and run_ann.py reloads optimizer.pt on resume: loading such a state dict and stepping must continue from it."""
    from cocodr_b200 import optim
    mk = (lambda ps: optim.AdamW(ps, lr=1e-3, eps=1e-8, semantics="torch")) if kind == "adamw" else \
         (lambda ps: optim.Lamb(ps, lr=1e-3, eps=1e-6))
    a, b = _params(7), _params(7)
    oa, ob = mk(a), mk(b)
    for step in range(2):
        _grads(a, 40 + step)
        _grads(b, 40 + step)
        oa.step()
        ob.step()
    sd = copy.deepcopy(oa.state_dict())
    for st in sd["state"].values():  # reference layout: python int step, tensors for the moments
        st["step"] = int(round(float(st["step"])))
        st.pop("trust_ratio", None)
    c = _params(7)
    for p, q in zip(c, a):
        p.data.copy_(q.data)
    oc = mk(c)
    oc.load_state_dict(sd)
    _grads(b, 50)
    _grads(c, 50)
    ob.step()
    oc.step()
    for p, q in zip(c, b):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item())
    assert float(oc.state_dict()["state"][0]["step"]) == 3.0
