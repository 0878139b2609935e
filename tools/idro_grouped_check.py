"""iDRO group gradients at BASELINE size (BERT-base, 64 triplets, L = 128, G = 50): the grouped-wgrad path
(iDROLoss._get_grad_grouped: one shared partial backward + per-group wgrads) against the per-group partial backwards
(iDROLoss._get_grad, the reference's way), parameter by parameter, with timings of both and of the whole iDRO step.
Writes gpurun_out/idro_grouped.json.  Usage: python tools/idro_grouped_check.py [--steps 5]
"""
import argparse
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/idro_grouped.json")
    ap.add_argument("--skip-pytest", action="store_true")
    ap.add_argument("--kernel-check", action="store_true")
    args = ap.parse_args()
    os.environ["CDR_IDRO_GROUPED"] = "1"  # read when cocodr_b200.dro_loss is imported
    if args.kernel_check:
        os.environ["CDR_IDRO_GROUPED_KERNEL"] = "1"
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    from transformers import BertConfig

    from cocodr_b200 import dro_loss, models, ops, optim
    assert dro_loss.iDROLoss.grouped_wgrad
    if not args.skip_pytest:  # the pinned 3-step reference trajectory (tests/golden/idro_tiny.npz) through the grouped path
        import pytest
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        rc = pytest.main([os.path.join(root, "tests", "test_model_gpu.py"), os.path.join(root, "tests", "test_gemm_gpu.py"),
                          "-q", "-k", "idro or segments or grouped_k", "-p", "no:cacheprovider"])
        print(json.dumps({"idro_trajectory_and_segments_pytest_rc": int(rc)}), flush=True)
    B, L, G = 64, 128, 50
    dev = torch.device("cuda:0")
    cfg = BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_labels=2)
    torch.manual_seed(0)
    model = models.BertDot_NLL_LN(cfg).to(dev).train()
    model.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=0), G, "idro", 0.25, 0.01, 0.1, 0.05)
    gen = torch.Generator().manual_seed(4321)
    ids = torch.randint(1000, cfg.vocab_size, (3 * B, L), generator=gen)
    ids[:, 0], ids[:, -1] = 101, 102
    ids, mask = ids.to(dev), torch.ones(3 * B, L, dtype=torch.long, device=dev)
    gid = torch.randint(0, G, (B,), generator=gen).to(dev)
    inp = (ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], ids[2 * B:], mask[2 * B:])
    crit = model.loss
    res = {"config": f"BERT-base, {B} triplets, L={L}, G={G}", "params": []}

    def ev_time(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    if args.kernel_check:  # cdr_gemm_grouped (one launch per weight) against cdr_gemm_segments (one per present group)
        loss, _, _ = model.forward_model(*inp)
        sums, counts = ops.group_stats(loss, gid, G)
        means = sums / (counts + (counts == 0).float())
        params = crit._params(model.bert)
        out = {}
        for flag in (False, True):
            type(crit).grouped_kernel = flag
            ms, m = ev_time(lambda: crit._get_grad_grouped(params, means, counts, gid, 3), 3)
            out[flag] = (ms, m)
        d = (out[True][1] - out[False][1]).abs().max().item()
        res = {"ms_segments": round(out[False][0], 2), "ms_one_launch": round(out[True][0], 2), "max_abs_diff": d,
               "max_abs": out[False][1].abs().max().item()}
        print(json.dumps(res), flush=True)
        del out
        opt = optim.AdamW(list(model.bert.parameters()), lr=5e-6, eps=1e-8, weight_decay=0.01,
                          semantics="torch").attach_shadows(model)

        def kstep():
            o = model(*inp, group_ids=gid)[0]
            opt.zero_grad(set_to_none=True)
            o.backward()
            opt.step()

        res["ms_step_one_launch"] = round(ev_time(kstep, args.steps)[0], 2)
        print(json.dumps(res), flush=True)
        with open(args.out.replace(".json", "_kernel.json"), "w") as f:
            json.dump(res, f, indent=1)
        return
    loss, _, _ = model.forward_model(*inp)
    sums, counts = ops.group_stats(loss, gid, G)
    means = sums / (counts + (counts == 0).float())
    params = crit._params(model.bert)
    names = [n for n, p in model.bert.named_parameters() if any(p is q for q in params)]
    t_slow, slow = ev_time(lambda: crit._get_grad(params, means, counts), 2)
    try:
        t_fast, fast = ev_time(lambda: crit._get_grad_grouped(params, means, counts, gid, 3), 3)
    except Exception:
        import traceback
        traceback.print_exc()
        print(json.dumps({"grouped_failed": True, "ms_per_group_backwards": round(t_slow, 2)}), flush=True)
        return
    res["ms_per_group_backwards"], res["ms_grouped_wgrad"] = round(t_slow, 2), round(t_fast, 2)
    off, worst = 0, 0.0
    for n, p in zip(names, params):
        k = p.numel()
        a, b = slow[:, off:off + k], fast[:, off:off + k]
        scale = a.abs().max().item()
        err = (a - b).abs().max().item()
        rel = err / (scale + 1e-30)
        if "key.bias" not in n:  # analytically zero: rounding noise on both sides
            worst = max(worst, rel)
        res["params"].append({"name": n, "max_abs": scale, "max_err": err, "rel": round(rel, 5)})
        off += k
    res["worst_rel_err"] = worst
    gs, gf = slow @ slow.t(), fast @ fast.t()
    res["gram_rel_err"] = ((gs - gf).abs().max() / gs.abs().max()).item()
    res["present_groups"] = int((counts > 0).sum().item())
    print(json.dumps({k: v for k, v in res.items() if k != "params"}), flush=True)
    for r in sorted(res["params"], key=lambda r: -r["rel"])[:6]:
        print(r, flush=True)
    del slow, fast, gs, gf

    # whole iDRO step, both ways
    opt = optim.AdamW(list(model.bert.parameters()), lr=5e-6, eps=1e-8, weight_decay=0.01,
                      semantics="torch").attach_shadows(model)

    def step():
        out = model(*inp, group_ids=gid)[0]
        opt.zero_grad(set_to_none=True)
        out.backward()
        opt.step()

    for flag in (False, True):
        type(crit).grouped_wgrad = flag
        ms, _ = ev_time(step, args.steps)
        res["ms_step_grouped" if flag else "ms_step_per_group"] = round(ms, 2)
    res["h_fun_after"] = crit.h_fun.tolist()[:8]
    print(json.dumps({k: res[k] for k in ("ms_step_per_group", "ms_step_grouped")}), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
