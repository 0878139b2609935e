"""SURVEY.md section 8(d), "the real bar": the same contrastive step on the SAME B200 through the stock library path
(HuggingFace BertModel under torch.autocast(fp16) + GradScaler, cuBLAS GEMMs, SDPA or eager attention, torch's fused
AdamW), timed with CUDA events next to this repo's step -- plus the two secondary figures section 8(d) asks for:
triplets/s of the reference-exact ANCE mode (3 encoder passes, per-triplet NLL) and the iDRO step (G = 50).

Not part of bench.py's contract line (that compares against the reference's CPU arm); the output goes to
gpurun_out/stock_bar.json and is summarised in profiles/.  Usage: python tools/stock_gpu_bar.py [--steps 10]
(--tiny --cpu runs the stock arm on a 2-layer model on the CPU: a syntax/API check for this container).
"""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

B, L = 64, 128


def make_cfg(tiny, **kw):
    from transformers import BertConfig
    extra = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256) if tiny else {}
    return BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_labels=2, **extra, **kw)


def synth(cfg, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, cfg.vocab_size, (n, L), generator=g)
    ids[:, 0], ids[:, -1] = 101, 102
    return ids.to(dev), torch.ones(n, L, dtype=torch.long, device=dev)


def timed(fn, steps, warmup, cuda):
    for _ in range(warmup):
        fn()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    t = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t) * 1e3 / steps


def stock_arm(args, dev, attn, one_pass):
    """HF BertModel, two towers (query batch, then passage batch -- the reference runs them sequentially,
    ANCE/model/models.py:97-99) or one concatenated pass; CE(q p^T, arange) in fp32; GradScaler + fused AdamW."""
    from transformers import BertModel
    cuda = dev.type == "cuda"
    cfg = make_cfg(args.tiny)
    cfg._attn_implementation = attn
    torch.manual_seed(0)
    model = BertModel(cfg, add_pooling_layer=False).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-6, fused=cuda)
    scaler = torch.amp.GradScaler("cuda", enabled=cuda)
    ids, mask = synth(cfg, 2 * B, 1234, dev)
    tgt = torch.arange(B, device=dev)
    dt = torch.float16 if cuda else torch.bfloat16

    def step():
        with torch.autocast(dev.type, dtype=dt):
            if one_pass:
                e = model(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0]
                q, p = e[:B], e[B:]
            else:
                q = model(input_ids=ids[:B], attention_mask=mask[:B]).last_hidden_state[:, 0]
                p = model(input_ids=ids[B:], attention_mask=mask[B:]).last_hidden_state[:, 0]
        loss = torch.nn.functional.cross_entropy(q.float() @ p.float().t(), tgt)
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    ms = timed(step, args.steps, 3, cuda)
    return {"arm": f"stock HF BertModel autocast fp16, attn={attn}, {'one 128-seq pass' if one_pass else 'two 64-seq towers'}, "
                   "GradScaler + torch fused AdamW", "ms_per_step": round(ms, 3), "pairs_per_s": round(B / ms * 1e3, 1)}


def ours_pairs(args, dev):
    from cocodr_b200 import models, optim
    from cocodr_b200.graph import GraphedTrainStep
    cfg = make_cfg(False)
    torch.manual_seed(0)
    model = models.BertDot_InBatch_NLL_LN(cfg).to(dev).train()
    opt = optim.AdamW(list(model.parameters()), lr=5e-6, eps=1e-8, weight_decay=0.01, semantics="torch").attach_shadows(model)
    ids, mask = synth(cfg, 2 * B, 1234, dev)
    ones = torch.ones(B, device=dev)
    g = GraphedTrainStep(model, opt, (ids[:B], mask[:B], ids[B:], mask[B:], None, None, True, None, ones))
    ms = timed(lambda: g(ids[:B], mask[:B], ids[B:], mask[B:]), args.steps, 3, True)
    return {"arm": "cocodr_b200 in-batch pair step (bench.py's step, CUDA graph, own AdamW)", "ms_per_step": round(ms, 3),
            "pairs_per_s": round(B / ms * 1e3, 1)}


def ours_triplet(args, dev):
    """Reference-exact ANCE mode: query + positive + negative towers, per-triplet NLL (models.py:80-115), ERM mean."""
    from cocodr_b200 import models, optim
    from cocodr_b200.graph import GraphedTrainStep
    cfg = make_cfg(False)
    torch.manual_seed(0)
    model = models.BertDot_NLL_LN(cfg).to(dev).train()
    opt = optim.AdamW([p for p in model.bert.parameters()], lr=5e-6, eps=1e-8, weight_decay=0.01,
                      semantics="torch").attach_shadows(model)
    ids, mask = synth(cfg, 3 * B, 4321, dev)
    inp = (ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], ids[2 * B:], mask[2 * B:])
    g = GraphedTrainStep(model, opt, inp)
    ms = timed(lambda: g(*inp), args.steps, 3, True)
    return {"arm": "cocodr_b200 BertDot_NLL_LN triplet step (3 towers, CUDA graph, own AdamW)", "ms_per_step": round(ms, 3),
            "triplets_per_s": round(B / ms * 1e3, 1), "tflops": round(B * 201.1e9 / (ms * 1e-3) / 1e12, 1)}


def ours_idro(args, dev):
    """Triplet step with the iDRO loss, G = 50 (dro_loss.py:216-254): per-group gradients of the last three layers,
    Gram matrix, h update.  Eager (the meters read one host transfer per step, as the reference's do)."""
    from cocodr_b200 import models, optim
    cfg = make_cfg(False)
    torch.manual_seed(0)
    model = models.BertDot_NLL_LN(cfg).to(dev).train()
    model.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=0), 50, "idro", 0.25, 0.01, 0.1, 0.05)
    opt = optim.AdamW([p for p in model.bert.parameters()], lr=5e-6, eps=1e-8, weight_decay=0.01,
                      semantics="torch").attach_shadows(model)
    ids, mask = synth(cfg, 3 * B, 4321, dev)
    gid = torch.randint(0, 50, (B,), generator=torch.Generator().manual_seed(5)).to(dev)

    def step():
        loss = model(ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], ids[2 * B:], mask[2 * B:], group_ids=gid)[0]
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    ms = timed(step, args.steps, 3, True)
    return {"arm": "cocodr_b200 BertDot_NLL_LN triplet step + iDRO (G=50), eager launches", "ms_per_step": round(ms, 3),
            "triplets_per_s": round(B / ms * 1e3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", default="gpurun_out/stock_bar.json")
    args = ap.parse_args()
    dev = torch.device("cpu" if args.cpu else "cuda:0")
    arms = [lambda: stock_arm(args, dev, "sdpa", False), lambda: stock_arm(args, dev, "sdpa", True),
            lambda: stock_arm(args, dev, "eager", False)]
    if not args.cpu:
        arms += [lambda: ours_pairs(args, dev), lambda: ours_triplet(args, dev), lambda: ours_idro(args, dev)]
    res = []
    for a in arms:
        try:
            r = a()
        except Exception as e:  # keep the other arms: one GPU call pays for all of them
            r = {"error": f"{type(e).__name__}: {e}"[:400]}
        res.append(r)
        print(json.dumps(r), flush=True)
        if not args.cpu:
            torch.cuda.empty_cache()
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"batch": f"{B} queries + {B} passages (+{B} negatives in triplet mode), L={L}, full-length masks",
                       "device": torch.cuda.get_device_name(0) if not args.cpu else "cpu", "torch": torch.__version__,
                       "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
