import os, sys, datetime, types
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
from transformers import BertConfig
from cocodr_b200 import models, optim, peeropt
from cocodr_b200.gradsync import GradSync
from oracle import bert_ref
tiny = dict(hidden=128, layers=3, heads=2, inter=512, vocab=2000, max_pos=64, type_vocab=2)
hf = BertConfig(vocab_size=2000, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=512,
                max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_labels=2)
qi, qm = (t.to(dev) for t in bert_ref.synth_batch(4, 32, 2000, 100 + rank))
pi, pm = (t.to(dev) for t in bert_ref.synth_batch(4, 32, 2000, 200 + rank))
w = torch.ones(4, device=dev)
def make():
    m = models.BertDot_InBatch_NLL_LN(hf); m.bert.load_state_dict(bert_ref.synth_state(tiny, 0), strict=False)
    return m.to(dev).train()
def one(m, opt, sync):
    opt.zero_grad(set_to_none=True)
    loss = m(qi, qm, pi, pm, weights=w)[0]
    with sync: loss.backward()
    g = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    opt.step(); torch.cuda.synchronize()
    return g
ms = [make() for _ in range(3)]
os_ = [optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01) for m in ms]
os_[0].attach_shadows(ms[0]); os_[1].attach_shadows(ms[1])
arena = peeropt.PeerArena(ms[2], os_[2])
syncs = [GradSync(ms[0]), GradSync(ms[1]), GradSync(ms[2], arena=arena)]
def cmp(a, b):
    worst = (0, "")
    for (n, x), (_, y) in zip(a.named_parameters(), b.named_parameters()):
        if x.grad is None and y.grad is None: continue
        e = (x - y).abs().max().item()
        if e > worst[0]: worst = (e, n, int(((x - y).abs() > 1e-6).sum()), x.numel())
    return worst
for step in range(3):
    gs = [one(m, o, s) for m, o, s in zip(ms, os_, syncs)]
    # gradient comparison: ref grads are averaged; peer grads local -> average them for comparison
    gp = {n: t.clone() for n, t in gs[2].items()}
    for t in gp.values():
        dist.all_reduce(t); t /= world
    ge = max(((gp[n] - gs[0][n]).abs().max().item() / (gs[0][n].abs().max().item() + 1e-20), n) for n in gp if "key.bias" not in n)
    g01 = max(((gs[1][n] - gs[0][n]).abs().max().item() / (gs[0][n].abs().max().item() + 1e-20), n) for n in gp if "key.bias" not in n)
    if rank == 0:
        print("step", step, "ref-vs-ref", cmp(ms[0], ms[1]), "peer-vs-ref", cmp(ms[2], ms[0]), "grad rel peer", ge, "grad rel ref2", g01, flush=True)
arena.check()
dist.barrier(); dist.destroy_process_group()
