"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes.

Run in the dev container only (needs /root/reference; it does not exist on the
GPU box):   python oracle/make_golden.py

Import recipe and the two transformers-5.5 shims for COCO follow SURVEY.md §8c.
Weights come from oracle.bert_ref.synth_state (seeded, deterministic), so the
fixtures hold only inputs' seeds and the reference's outputs.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bert_ref  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")

TINY = dict(hidden=128, layers=12, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)
BASE = bert_ref.make_config()


def hf_config(cfg, **kw):
    from transformers import BertConfig
    return BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                      num_attention_heads=cfg["heads"], intermediate_size=cfg["inter"],
                      max_position_embeddings=cfg["max_pos"], type_vocab_size=cfg["type_vocab"],
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                      attn_implementation="eager", **kw)


def import_ance():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k == "data" or k.startswith("data.")]:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(REF, "ANCE"))
    from model.models import BertDot_NLL_LN  # noqa
    sys.path.pop(0)
    return BertDot_NLL_LN


def build_ance(cfg, seed=0):
    cls = import_ance()
    m = cls(hf_config(cfg, num_labels=2))
    st = bert_ref.synth_state(cfg, seed)
    missing = m.bert.load_state_dict(st, strict=False)
    assert all("pooler" in k or "position_ids" in k for k in missing.missing_keys), missing
    assert not missing.unexpected_keys
    return m, st


def triplet_batch(cfg, B, L, seed, full=False):
    q, mq = bert_ref.synth_batch(B, L, cfg["vocab"], seed, full)
    a, ma = bert_ref.synth_batch(B, L, cfg["vocab"], seed + 1, full)
    b, mb = bert_ref.synth_batch(B, L, cfg["vocab"], seed + 2, full)
    return q, mq, a, ma, b, mb


def gen_ance(cfg, tag, B, L, seed, full, with_grads):
    m, st = build_ance(cfg)
    m.train()  # dropout probs are 0 in the config
    q, mq, a, ma, b, mb = triplet_batch(cfg, B, L, seed, full)
    w = torch.linspace(0.5, 1.5, B)
    loss, acc, logits = m(q, mq, a, ma, b, mb, weights=w)
    out = dict(seed=seed, B=B, L=L, full=int(full), erm_loss=loss.item(), accs=acc.numpy(), logits=logits.detach().numpy(),
               weights=w.numpy())
    with torch.no_grad():
        out["q_emb"] = m.query_emb(q, mq).numpy()
        out["a_emb"] = m.body_emb(a, ma).numpy()
        out["b_emb"] = m.body_emb(b, mb).numpy()
        per, _, _ = m.forward_model(q, mq, a, ma, b, mb)
        out["loss"] = per.numpy()
        out["qp_infonce"] = torch.nn.functional.cross_entropy(
            torch.from_numpy(out["q_emb"]) @ torch.from_numpy(out["a_emb"]).t(), torch.arange(B), reduction="none").numpy()
    if with_grads:
        m.zero_grad()
        loss.backward()
        named = dict(m.bert.named_parameters())
        for n in ("embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                  "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias",
                  *bert_ref.layer_param_names(0), *bert_ref.layer_param_names(cfg["layers"] - 1)):
            g = named[n].grad
            if n.endswith("word_embeddings.weight"):
                out["grad." + n + ".rownorm"] = g.norm(dim=1).numpy()  # (row sums vanish: LN bwd is mean-free)
                out["grad." + n + ".norm"] = g.norm().item()
            else:
                out["grad." + n] = g.numpy()
    np.savez_compressed(os.path.join(OUT, f"ance_{tag}.npz"), **out)
    print("ance", tag, "erm_loss", out["erm_loss"])


def gen_idro(cfg, tag, B, L, seed, n_groups, dro_type, steps=3):
    m, st = build_ance(cfg)
    args = types.SimpleNamespace(model_size="base", local_rank=0)
    hp = dict(alpha=0.25, eps=0.01, ema=0.1, rho=0.05)
    m.add_group_loss(args, n_groups, dro_type, hp["alpha"], hp["eps"], hp["ema"], hp["rho"], True)
    m.train()
    out = dict(seed=seed, B=B, L=L, n_groups=n_groups, steps=steps, **hp)
    for s in range(steps):
        q, mq, a, ma, b, mb = triplet_batch(cfg, B, L, seed + 10 * s)
        gid = torch.randint(0, n_groups, (B,), generator=torch.Generator().manual_seed(seed + s))
        w = torch.ones(B)
        m.zero_grad()
        robust, acc, gl, gc = m(q, mq, a, ma, b, mb, group_ids=gid, weights=w)
        robust.backward()
        out[f"group_ids_{s}"] = gid.numpy()
        out[f"robust_{s}"] = robust.item()
        out[f"group_losses_{s}"] = gl.numpy()
        out[f"group_counts_{s}"] = gc.numpy()
        out[f"h_fun_{s}"] = m.loss.h_fun.detach().numpy().copy()
        out[f"grad_q11_{s}"] = dict(m.bert.named_parameters())[
            f"encoder.layer.{cfg['layers'] - 1}.attention.self.query.weight"].grad.numpy().copy()
        if dro_type == "dro-greedy":
            out[f"sum_losses_{s}"] = m.loss.sum_losses.numpy().copy()
            out[f"count_cat_{s}"] = m.loss.count_cat.numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print(tag, [out[f"robust_{s}"] for s in range(steps)], out[f"h_fun_{steps - 1}"])


def import_coco():
    for k in [k for k in sys.modules if k in ("modeling", "arguments")]:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(REF, "COCO"))
    import modeling  # noqa
    sys.path.pop(0)
    return modeling


class _TupleLayer(torch.nn.Module):
    """Shim (2): installed BertLayer returns a Tensor; reference indexes [0] (modeling.py:216-220)."""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer

    def forward(self, h, mask):
        return (self.layer(h, mask),)


def gen_coco(cfg, tag, n_docs, L, seed):
    modeling = import_coco()
    from transformers import BertForMaskedLM
    torch.manual_seed(seed)
    lm = BertForMaskedLM(hf_config(cfg))
    st = bert_ref.synth_state(cfg, 0)
    lm.bert.load_state_dict(st, strict=False)
    margs = types.SimpleNamespace(n_head_layers=2, skip_from=2, late_mlm=True)
    dargs = types.SimpleNamespace(train_method="coco")
    targs = types.SimpleNamespace(per_device_train_batch_size=n_docs, local_rank=-1)
    m = modeling.CoCondenserForPretraining(lm, margs, dargs, targs)
    # shim (1): reference passes `device` where installed HF expects `dtype` (modeling.py:193-197)
    orig = m.lm.get_extended_attention_mask
    m.lm.get_extended_attention_mask = lambda mask, shape, device=None: orig(mask, shape)
    head_state = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.startswith("lm.bert.")}
    raw_heads = list(m.c_head)
    m.c_head = torch.nn.ModuleList([_TupleLayer(l) for l in raw_heads])
    m.train()
    ids, mask = bert_ref.synth_batch(2 * n_docs, L, cfg["vocab"], seed)
    g = torch.Generator().manual_seed(seed + 5)
    labels = torch.where((torch.rand(ids.shape, generator=g) < 0.15) & (mask > 0), ids, torch.full_like(ids, -100))
    total = m({"input_ids": ids, "attention_mask": mask}, labels)
    with torch.no_grad():
        lm_out = m.lm(input_ids=ids, attention_mask=mask, labels=labels, output_hidden_states=True, return_dict=True)
        cls = lm_out.hidden_states[-1][:, 0]
        co = m.compute_contrastive_loss(cls.clone())
    m.zero_grad()
    total.backward()
    named = dict(m.named_parameters())
    out = dict(seed=seed, n_docs=n_docs, L=L, skip_from=2, n_head_layers=2, total=total.item(), cls=cls.numpy(),
               co_loss=co.numpy(), lm_mlm_loss=lm_out.loss.item(), labels=labels.numpy(),
               grad_word_rownorm=named["lm.bert.embeddings.word_embeddings.weight"].grad.norm(dim=1).numpy(),
               grad_l0_query=named["lm.bert.encoder.layer.0.attention.self.query.weight"].grad.numpy())
    for k, v in head_state.items():
        out["state." + k.replace("c_head.", "c_head.")] = v.numpy()
    np.savez_compressed(os.path.join(OUT, f"coco_{tag}.npz"), **out)
    print("coco", tag, "total", out["total"], "co_mean", co.mean().item())


def contrastive_inputs(n, h, seed):
    """Deterministic inputs shared by the generator and the tests (SURVEY §8d 'trained-like')."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(h, generator=g) * 14.7 / h ** 0.5
    trained = base[None, :] + 0.015 * torch.randn(n, h, generator=g)
    gauss = torch.randn(n, h, generator=g) * 0.5
    return {"trained": trained, "gauss": gauss}


def gen_contrastive_only():
    modeling = import_coco()
    out = {}
    for n, h, nm in ((16, 128, "small"), (512, 1024, "cfg4")):
        dummy = types.SimpleNamespace(co_target=bert_ref_target(n), _world_size=lambda: 1)
        for tag, e in contrastive_inputs(n, h, 11).items():
            e = e.clone().requires_grad_(True)
            loss = modeling.CoCondenserForPretraining.compute_contrastive_loss(dummy, e)
            loss.mean().backward()
            out[f"{nm}_{tag}_loss"] = loss.detach().numpy()
            out[f"{nm}_{tag}_grad_head"] = e.grad[:8].numpy()
            out[f"{nm}_{tag}_grad_rownorm"] = e.grad.norm(dim=1).numpy()
    np.savez_compressed(os.path.join(OUT, "contrastive.npz"), **out)
    print("contrastive", {k: v.shape for k, v in out.items() if k.endswith("loss")})


def bert_ref_target(n):
    return torch.arange(n, dtype=torch.long).view(-1, 2).flip([1]).flatten().contiguous()


def gen_scan():
    from oracle import scan_ref
    out = {}
    for kind in ("exact", "gauss"):
        Q, P = scan_ref.synth_corpus(6000, 37, 768, seed=7, kind=kind)
        D, I = scan_ref.search(Q, P, 100, chunk=2048)
        D2, I2 = scan_ref.search(Q, P, 100, chunk=100000)
        assert (I == I2).all() and (D == D2).all()
        out[f"{kind}_D"] = D
        out[f"{kind}_I"] = I
    np.savez_compressed(os.path.join(OUT, "scan_small.npz"), **out)
    print("scan ok")


LAMB_SHAPES = [(96, 64), (64,), (7,), (300, 3)]


def lamb_inputs(seed):
    """Seeded parameters and per-step gradients shared by the fixture generator and the tests."""
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(*s, generator=g) * 0.05 for s in LAMB_SHAPES]
    params[2].zero_()  # an all-zero tensor: weight_norm == 0 -> trust ratio 1 (lamb.py:112-113)
    grads = [[torch.randn(*s, generator=g) * 0.01 for s in LAMB_SHAPES] for _ in range(3)]
    return params, grads


def gen_lamb():
    """The reference's own Lamb (ANCE/utils/lamb.py, unmodified; only the absent tensorboardX import is stubbed)
    driven for 3 steps -> parameters and trust ratios.  Pins oracle/optim_ref.lamb_step and the CUDA Lamb."""
    import importlib.util
    sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))
    spec = importlib.util.spec_from_file_location("ref_lamb", os.path.join(REF, "ANCE", "utils", "lamb.py"))
    ref_lamb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_lamb)
    out = {"seed": 77}
    for tag, wd in (("wd0", 0.0), ("wd01", 0.01)):
        params, grads = lamb_inputs(77)
        ps = [torch.nn.Parameter(p.clone()) for p in params]
        opt = ref_lamb.Lamb(ps, lr=1e-3, eps=1e-6, weight_decay=wd)
        for step in range(3):
            for p, g in zip(ps, grads[step]):
                p.grad = g.clone()
            opt.step()
        for i, p in enumerate(ps):
            out[f"{tag}.p{i}"] = p.detach().numpy()
            out[f"{tag}.trust{i}"] = np.float32(float(opt.state[p]["trust_ratio"]))
    np.savez_compressed(os.path.join(OUT, "lamb_tiny.npz"), **out)
    print("lamb ok")


def gen_mining():
    """The reference's own GenerateNegativePassaageID (ANCE/drivers/run_ann_data_gen.py:497-570), executed from its
    source: the driver module does not import here (faiss, pytrec_eval, transformers.AdamW), so the function's AST node
    is compiled on its own with the names it uses (random, np, trange).  Both branches: SelectTopK, and the shuffled one
    with Python's ``random`` seeded -- the per-query permutations it draws are replayed and stored so the oracle and the
    CUDA kernel can walk the candidates in the same order."""
    import ast
    import random
    path = os.path.join(REF, "ANCE", "drivers", "run_ann_data_gen.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "GenerateNegativePassaageID"][0]
    ns = {"random": random, "np": np, "trange": range, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    ref_fn = ns["GenerateNegativePassaageID"]
    rng = np.random.RandomState(5)
    n_q, k, n_docs, n_neg = 40, 60, 400, 7
    doc_pid = rng.permutation(4000)[:n_docs].astype(np.int64)
    doc_pid[::9] = doc_pid[2]  # repeated passage ids
    I = rng.randint(0, n_docs, size=(n_q, k)).astype(np.int64)
    qids = np.arange(1000, 1000 + n_q)
    pos = {int(q): int(doc_pid[rng.randint(0, n_docs)]) for q in qids}
    pos[int(qids[1])] = 999_999  # positive never retrieved
    I[2, 0] = int(np.where(doc_pid == pos[int(qids[2])])[0][0])
    out = {"I": I, "doc_pid": doc_pid, "pos": np.array([pos[int(q)] for q in qids], dtype=np.int64), "n_neg": n_neg}
    for tag, topk in (("topk", True), ("shuf", False)):
        args = types.SimpleNamespace(ann_measure_topk_mrr=topk, negative_sample=n_neg, rank=0)
        random.seed(99)
        negs, mrr = ref_fn(args, qids, doc_pid, pos, I, set(int(q) for q in qids))
        table = -np.ones((n_q, n_neg), dtype=np.int64)
        for i, q in enumerate(qids):
            table[i, :len(negs[q])] = negs[q]
        out[f"{tag}.neg"], out[f"{tag}.rr"] = table, np.asarray(mrr, dtype=np.float64)
        if not topk:  # replay the permutations the function drew (one random.shuffle(list(range(k))) per query)
            random.seed(99)
            orders = []
            for _ in range(n_q):
                o = list(range(k))
                random.shuffle(o)
                orders.append(o)
            out["shuf.order"] = np.asarray(orders, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "mining_tiny.npz"), **out)
    print("mining ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=0, world_size=1)
    which = sys.argv[1:] or ["ance", "idro", "greedy", "coco", "contrastive", "scan", "base", "lamb", "mining"]
    if "ance" in which:
        gen_ance(TINY, "tiny", B=4, L=32, seed=100, full=False, with_grads=True)
    if "idro" in which:
        gen_idro(TINY, "idro_tiny", B=8, L=32, seed=200, n_groups=5, dro_type="idro")
    if "greedy" in which:
        gen_idro(TINY, "dro_greedy_tiny", B=8, L=32, seed=300, n_groups=5, dro_type="dro-greedy")
    if "coco" in which:
        gen_coco(TINY, "tiny", n_docs=4, L=32, seed=400)
    if "contrastive" in which:
        gen_contrastive_only()
    if "scan" in which:
        gen_scan()
    if "lamb" in which:
        gen_lamb()
    if "mining" in which:
        gen_mining()
    if "base" in which:
        gen_ance(BASE, "cfg1_base", B=8, L=128, seed=500, full=True, with_grads=False)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
