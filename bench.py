#!/usr/bin/env python
"""bench.py -- headline benchmark of the COCO-DR contrastive hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (configs[1] of BASELINE.json; the same per-GPU work at N > 1 = weak scaling): BERT-base bi-encoder,
seq_len 128, per-GPU batch 64 queries + 64 passages, synthetic full-length token ids, seeded random-init weights,
dropout p = 0.1 (HF defaults, fused).  One *step* = encoder forward of both towers (one fused launch sequence) -> fp32
CLS embeddings -> (all-gather of the passage embeddings when N > 1) -> in-batch InfoNCE -> backward -> (gradient
all-reduce) -> fused AdamW update.  Metric: query+passage pairs / s, whole job.

  value        inputs already resident in HBM when the timed region starts.  The K timed steps are repeated
               ``inner_repeats`` times inside ONE timed region so that it lasts >= 2 s (sustained clocks / power cap);
               ms_per_step is the mean over all K * inner_repeats steps
  e2e          the same step through the public model API with the batch in pinned host memory: H2D copy of ids/masks
               and a D2H read of the loss inside the timed region
  roofline     the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of every GEMM launch of one
               instrumented step; ``frac`` against the measured SUSTAINED bf16 peak (the regime of a >= 2 s region),
               ``frac_burst`` against the burst peak (MEASURED_PEAKS.json)
  cpu_baseline the CPU oracle port of the same step (oracle/, torch fp32, all host threads), bounded sample, N = 1
  gpu_baseline the same step through the STOCK library path on the same GPU (HF BertModel, torch.autocast fp16, SDPA,
               torch fused AdamW), CUDA-graphed like ours, with its own clock record (N = 1)
  parity       N > 1: one step through the peer-memory exchange + GradSync vs the same step through NCCL all-gather and
               an explicit all-reduce of the local gradients (loss and gradient agreement, measured on the job)
  scan         corpus-scan sub-metric (configs[4]): 1M x 768 fp16 docs sharded over the N GPUs, 1000 queries, k = 1000
  inference    embedding-inference loop (evaluate/drivers/run_ann_data_gen.py:152-206): sequences / s
  idro         configs[2]: the step with iDRO group weights (G = 50), reference-exact triplet model and in-batch head
  coco         configs[3]: BERT-large, L = 256, 64 spans / GPU, Condenser head + MLM + sequence-contrastive loss

``--impl reference`` times the CPU port alone (the reference is pure Python on top of HF/PyTorch and cannot travel to
the GPU box).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query+passage pairs/sec (contrastive step)"
UNIT = "pairs/s"
SEQ_LEN, PER_GPU_BATCH = 128, 64
WORKLOAD = "BERT-base seq_len=128, per-GPU batch=64 q + 64 p, in-batch InfoNCE (fwd+loss+bwd+AdamW)"
MIN_REGION_MS = 2000.0


def fwd_flops_per_seq(H=768, I=3072, layers=12, L=SEQ_LEN):
    return layers * (2 * L * H * 3 * H + 2 * 2 * L * L * H + 2 * L * H * H + 2 * 2 * L * H * I)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tensor": float(p.get("bf16_tflops_sustained") or p["bf16_tflops"]), "tensor_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:
        return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe):
    one streaming ``nvidia-smi -lms`` process; only samples that fall between mark_start() and mark_end()
    are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index, period_ms=20):
        self.index, self.period_ms = index, period_ms
        self.samples, self.proc, self._t = [], None, None
        self.t0 = self.t1 = None

    def _loop(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 6:
                self.samples.append((time.perf_counter(), parts))

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
            time.sleep(0.3)  # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None
        return self

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self._t.join(timeout=5)

    def summary(self):
        inside = [p for t, p in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        use = inside if inside else [p for _, p in self.samples[-3:]]
        sm = [int(s[0]) for s in use if s[0].isdigit()]
        mx = [int(s[1]) for s in use if s[1].isdigit()]
        pw = []
        for s in use:
            try:
                pw.append(float(s[6]))
            except (IndexError, ValueError):
                pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in use for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_in_timed_region": len(inside),
                "power_w_median": statistics.median(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port_step_fn(n_pairs, threads=None, dropout=0.1):
    """One contrastive step of the CPU oracle port on ``n_pairs`` pairs; returns a callable step()."""
    import torch

    from oracle import bert_ref, dropout_ref, heads_ref
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    cfg = bert_ref.make_config()
    st = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(cfg, 0).items()}
    opt = torch.optim.AdamW(list(st.values()), lr=5e-6)
    q, mq = bert_ref.synth_batch(n_pairs, SEQ_LEN, cfg["vocab"], 1234, full=True)
    p, mp = bert_ref.synth_batch(n_pairs, SEQ_LEN, cfg["vocab"], 1235, full=True)
    drop = dropout_ref.TorchDropSpec(dropout, dropout) if dropout > 0 else None  # the reference's nn.Dropout work

    def step():
        e = bert_ref.cls_embedding(st, torch.cat([q, p]), torch.cat([mq, mp]), cfg, drop=drop)
        loss = heads_ref.qp_infonce(e[:n_pairs], e[n_pairs:]).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, torch.get_num_threads()


def time_cpu_port(n_pairs, steps, warmup, dropout=0.1):
    step, threads = cpu_port_step_fn(n_pairs, dropout=dropout)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_pairs / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_pairs = 8
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    v, dt, threads = time_cpu_port(n_pairs, steps, warmup, args.dropout)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n_pairs} pairs per step on the host CPU",
                       "dropout": f"p={args.dropout} (torch CPU Bernoulli masks at HF's four nn.Dropout sites)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{steps} steps of {n_pairs} q+p pairs (BERT-base, L=128, fp32, AdamW)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers (ours)
class Ctx:
    """World / device / timing plumbing shared by the headline metric and the sub-metrics."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import datetime
            if args.nccl_ctas > 0:
                # NCCL's kernels and the persistent one-CTA-per-SM GEMMs cannot share an SM (shared memory): cap the
                # collective's CTAs and size the persistent grids for the remaining SMs, so no GEMM runs a second wave
                os.environ["NCCL_MAX_CTAS"] = str(args.nccl_ctas)
                os.environ["NCCL_MIN_CTAS"] = "1"
            # a mismatched collective must abort within minutes, not hold N GPUs for NCCL's default 10
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=90))
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, sampler=None):
        """ms for ``steps`` calls of fn(i), device-timed, max over ranks, barrier + synchronize on both sides."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.mark_start()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        if sampler is not None:
            sampler.mark_end()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def gemm_roofline(self, run_one_step, note):
        """Every cdr_gemm launch of one eagerly launched step bracketed by CUDA events on the launching stream."""
        from cocodr_b200 import kernels
        torch = self.torch
        self.barrier()
        kernels.gemm_events = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_one_step()
        e1.record()
        torch.cuda.synchronize()
        ev, kernels.gemm_events = kernels.gemm_events, None
        gemm_ms = sum(a.elapsed_time(b) for _, a, b in ev)
        flops = sum(f for f, _, _ in ev)
        achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        pk = self.peaks
        return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": pk["tensor"],
                "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "peak_burst": pk["tensor_burst"],
                "frac_burst": achieved / pk["tensor_burst"], "peak_source": pk["source"] + " (bf16: sustained / burst)",
                "launches_per_step": len(ev), "avg_launch_us": gemm_ms * 1e3 / max(1, len(ev)),
                "gemm_share_of_step": gemm_ms / e0.elapsed_time(e1), "algorithmic_tflop_per_step": flops / 1e12,
                "note": note}

    def op_breakdown(self, run_one_step):
        from cocodr_b200 import kernels
        kernels.op_events = []
        run_one_step()
        self.torch.cuda.synchronize()
        opev, kernels.op_events = kernels.op_events, None
        op_ms = {}
        for name, a, b in opev:
            t = op_ms.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(b)
        return {k: {"calls": v[0], "ms": round(v[1], 3)} for k, v in sorted(op_ms.items(), key=lambda kv: -kv[1][1])}


def _safe(name, fn):
    """Sub-metrics must not take the headline line down with them: an exception becomes {'error': ...}."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        import traceback
        sys.stderr.write(f"[bench] sub-metric {name} failed:\n{traceback.format_exc()}\n")
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def synth_ids(torch, gen, n, L, vocab):
    ids = torch.randint(1000, vocab, (n, L), generator=gen)
    ids[:, 0], ids[:, -1] = 101, 102
    return ids, torch.ones(n, L, dtype=torch.long)


# ------------------------------------------------------------------------------------------------ headline
def run_ours(args):
    import torch
    from transformers import BertConfig

    from cocodr_b200 import _lib, kernels, models

    cx = Ctx(args)
    dist, world, rank, dev = cx.dist, cx.world, cx.rank, cx.dev
    _lib.check(_lib.load().cdr_device_check(), "cdr_device_check")
    if world > 1 and args.nccl_ctas > 0:
        _lib.check(_lib.load().cdr_set_sm_budget(148 - args.nccl_ctas), "cdr_set_sm_budget")

    torch.manual_seed(0)
    cfg = BertConfig(hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout, num_labels=2)
    model = models.BertDot_InBatch_NLL_LN(cfg).to(dev).train()
    net = model
    sync = None
    arena = None
    if world > 1:
        if args.ddp:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[cx.local], find_unused_parameters=True,
                                                            gradient_as_bucket_view=True)
        else:  # native path: per-layer flat gradient buffers all-reduced on a side stream during backward
            from cocodr_b200.gradsync import GradSync
            for p_ in model.parameters():  # identical initial weights on every rank
                dist.broadcast(p_.data, 0)
            sync = GradSync(model)
            if not args.torch_adamw and not args.nccl_grads:
                # gradient exchange inside the optimizer kernel over NVLink peer memory (cocodr_b200.peeropt): the
                # parameters, shadows and gradient buffers move into one symmetric arena; no NCCL all-reduce runs
                from cocodr_b200 import peeropt
                try:
                    arena = peeropt.PeerArena(model)
                except Exception as e:  # noqa: BLE001  (no symmetric memory on this box: every rank falls back together)
                    sys.stderr.write(f"[bench] peer-memory optimizer unavailable ({type(e).__name__}: {e}); NCCL gradient path\n")
                    arena = None
                ok = torch.tensor([1 if arena is not None else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0:
                    arena = None
    if world > 1 and not args.nccl_gather:
        model.enable_peer_gather(True)  # CLS all-gather / gradient reduce-scatter through peer memory (NVLink stores)
    # N > 1: the NCCL gradient all-reduces and the peer-memory exchange are captured into the same CUDA graph
    use_graph = not args.no_graph and (world == 1 or not args.ddp)
    if args.torch_adamw:  # library optimizer (A/B only): torch's fused AdamW + the encoder's own weight-shadow cast
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=5e-6, fused=True,
                                capturable=use_graph)
    else:  # our multi-tensor AdamW: one launch updates all parameters AND rewrites the fp16 operand shadows
        from cocodr_b200 import optim as cdr_optim
        opt = cdr_optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=5e-6, eps=1e-8,
                              weight_decay=0.01, semantics="torch").attach_shadows(model)

    B, L = PER_GPU_BATCH, SEQ_LEN
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4  # distinct pinned host batches cycled by the e2e loop
    host = [tuple(t.pin_memory() for t in synth_ids(torch, g, 2 * B, L, cfg.vocab_size)) for _ in range(n_host)]
    dev_batches = [tuple(t.to(dev) for t in hb) for hb in host]
    ones = torch.ones(B, device=dev)

    # ---- N > 1: parity of the multi-GPU step, measured on the job itself before anything is captured
    parity = None
    if world > 1 and sync is not None and not args.no_parity:
        parity = _safe("parity", lambda: multi_gpu_parity(cx, model, sync, dev_batches[0], ones))
        if arena is not None and isinstance(parity, dict) and "error" not in parity:
            parity["peer_adam"] = _safe("peer_adam", lambda: peer_adam_parity(cx, model, arena, dev_batches[0], ones))
    if arena is not None:
        from cocodr_b200.gradsync import GradSync
        arena.attach(opt)
        sync = GradSync(model, arena=arena)

    def step(ids, mask):
        loss = net(ids[:B], mask[:B], ids[B:], mask[B:], weights=ones)[0]
        opt.zero_grad(set_to_none=True)
        if sync is not None:
            with sync:
                loss.backward()
        else:
            loss.backward()
        opt.step()
        return loss

    graphed = None
    if use_graph:
        # the repo's public step helper: the whole step captured once, replayed per batch
        from cocodr_b200.graph import GraphedTrainStep
        ids0, mask0 = dev_batches[0]
        graphed = GraphedTrainStep(net, opt, (ids0[:B], mask0[:B], ids0[B:], mask0[B:], None, None, True, None, ones),
                                   backward_ctx=sync)

        def step(ids, mask):  # noqa: F811
            return graphed(ids[:B], mask[:B], ids[B:], mask[B:])

    def resident_step(i):
        ids, mask = dev_batches[i % n_host]
        step(ids, mask)

    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_step(i):
        hi, hm = host[i % n_host]
        if graphed is not None:  # pinned host -> the graph's static input buffers (H2D), replay, read the loss
            loss = graphed(hi[:B], hm[:B], hi[B:], hm[B:])
        else:
            loss = step(hi.to(dev, non_blocking=True), hm.to(dev, non_blocking=True))
        return loss.item()  # D2H read of the step's result

    warm = max(args.warmup, 3)
    for i in range(warm):
        resident_step(i)
    est = cx.timed(resident_step, 5) / 5  # sizes the timed region; these 5 steps are extra warm-up
    inner = max(1, math.ceil(MIN_REGION_MS / max(est * args.steps, 1e-3)))
    n_timed = args.steps * inner
    l0 = kernels.launches
    with ClockSampler(cx.local) as clocks:
        ms = cx.timed(resident_step, n_timed, clocks)
    launches = kernels.launches - l0
    if graphed is not None:
        launches = graphed.launches_per_replay * n_timed
    for i in range(2):
        e2e_step(i)
    n_e2e = args.steps * max(1, inner // 2)
    ms_e2e = cx.timed(e2e_step, n_e2e)

    value = B * world * n_timed / (ms * 1e-3)
    e2e_value = B * world * n_e2e / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: every GEMM launch of one instrumented EAGER step (same kernels and shapes
    # as the replayed graph; the events add no GPU work)
    eager = graphed._eager_step if graphed is not None else (lambda: resident_step(0))
    eager()  # warm the eager path (allocator) before the instrumented steps
    roofline = cx.gemm_roofline(eager, "per-launch CUDA events over one eagerly launched step; same kernels/shapes as the timed graph")
    traffic = None  # DRAM bytes per GEMM launch from the committed ncu pass over one step (same command, --no-graph)
    for f in ("r02_gemm_traffic.json", "r01d_gemm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", f)) as fh:
                traffic = float(json.load(fh)["traffic_bytes_per_launch"])
            roofline["traffic_source"] = "profiles/" + f
            break
        except Exception:
            pass
    roofline["traffic"] = traffic
    roofline["traffic_unit"] = "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches of one step)"
    op_breakdown = cx.op_breakdown(eager)
    # executed tensor-core FLOPs of the step (GEMM launches as counted above + the attention matmuls, fwd + 2.5x bwd)
    attn_flops = 12 * 2 * B * 12 * (4 * L * L * 64) * 3.5
    exec_tflop = roofline["algorithmic_tflop_per_step"] + attn_flops / 1e12
    step_tflops = exec_tflop / (ms / n_timed * 1e-3)

    line = None
    if rank == 0:
        par = f"dp{world}"
        if world > 1:
            par += (" + NCCL all-gather of passage CLS + " if args.nccl_gather else
                    " + passage CLS pushed into every rank's HBM by the last LayerNorm kernel (NVLink peer stores) + ")
            par += ("DDP all-reduce" if args.ddp else
                    "gradient exchange inside the AdamW kernel over NVLink peer memory (reduce-scatter + sharded update + "
                    "all-gather in one pass, no NCCL kernel)" if arena is not None else
                    "per-layer NCCL all-reduce of flat gradient buffers overlapped with backward")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / n_timed, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "inner_repeats": inner, "timed_steps": n_timed, "timed_region_ms": ms,
                "config": {"workload": WORKLOAD, "global_pairs_per_step": B * world, "parallelism": par,
                           "l2": "per-step working set (~5 GB of activations) exceeds the 126 MB L2; 4 input batches cycled",
                           "optimizer": ("torch fused AdamW" if args.torch_adamw else "cdr_adam_multi (own fused multi-tensor AdamW + fp16 shadow refresh)") + " inside the timed step",
                           "dropout": (f"p={args.dropout} fused (HF defaults: embeddings, attention probabilities, both dense outputs "
                                       "of every layer; Philox4x32-10 masks regenerated in the backward, offset advanced on the "
                                       "device inside the captured graph)") if args.dropout > 0 else "p=0 (--dropout 0)",
                           "cuda_graph": graphed is not None,
                           "nccl_ctas": args.nccl_ctas,
                           "timed_region": f">= {MIN_REGION_MS / 1e3:.0f} s: the K steps are repeated inner_repeats times inside one event pair"},
                "clocks": clocks.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / n_e2e, "timed_steps": n_e2e},
                "gpu_launches": launches,
                "roofline": roofline,
                "step_tensor_tflops": {"executed_tflop_per_step": exec_tflop, "achieved": step_tflops,
                                       "frac_sustained": step_tflops / cx.peaks["tensor"],
                                       "frac_burst": step_tflops / cx.peaks["tensor_burst"],
                                       "note": "FLOPs the kernels execute (GEMM launches + attention matmuls) / whole step time"},
                "op_breakdown_ms_per_step": op_breakdown,
                "parity": parity}

    # ---- release the headline model before the sub-metrics
    if graphed is not None:
        graphed.graph.reset()
    del graphed, opt, net, model, sync, dev_batches, step, resident_step, e2e_step, eager
    torch.cuda.empty_cache()

    subs = {}
    if not args.no_scan:
        subs["scan"] = _safe("scan", lambda: bench_scan(cx))
    if not args.quick:
        subs["inference"] = _safe("inference", lambda: bench_inference(cx))
        subs["idro"] = _safe("idro", lambda: bench_idro(cx))
        subs["coco"] = _safe("coco", lambda: bench_coco(cx))
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, threads = time_cpu_port(8, 3, 1, args.dropout)
        subs["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "3 steps of 8 q+p pairs (BERT-base, L=128, fp32, AdamW, torch dropout p="
                                          f"{args.dropout}), oracle port on host CPU"}
    else:
        subs["cpu_baseline"] = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and not args.quick:
        subs["gpu_baseline"] = _safe("gpu_baseline", lambda: bench_gpu_baseline(cx))

    if rank == 0:
        line.update(subs)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown after the result is out; a watchdog ends the process if the (already useless) teardown ever wedges.
        def _bail():
            time.sleep(20)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ N > 1 parity
def peer_adam_parity(cx, model, arena, batch, ones):
    """One optimizer step through the peer-memory path (cdr_adam_multi_peer) vs the NCCL path (GradSync all-reduce +
    the same AdamW on every rank) from the same weights, batch and dropout masks: max relative parameter difference.
    (One step only: Adam's normalised update amplifies run-to-run summation noise over several steps.)"""
    torch, dist = cx.torch, cx.dist
    from cocodr_b200 import optim as cdr_optim
    from cocodr_b200.gradsync import GradSync
    B = ones.shape[0]
    ids, mask = batch
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    params = [p for _, p in named]
    saved = [p.detach().clone() for p in params]
    # (the key bias has an analytically zero gradient -- softmax is invariant to it -- so Adam normalises pure rounding
    # noise there and the sign of its update is arbitrary: excluded, as in the gradient comparison above)
    skip = {i for i, (n, _) in enumerate(named) if "key.bias" in n}

    def one_step(use_peer):
        with torch.no_grad():
            for p, s0 in zip(params, saved):
                p.copy_(s0)
        opt = cdr_optim.AdamW(params, lr=1e-4, eps=1e-8, weight_decay=0.01, semantics="torch")
        if use_peer:
            arena.attach(opt)
            sync = GradSync(model, arena=arena)
        else:
            opt.attach_shadows(model)
            sync = GradSync(model)
        model.bert.set_dropout_seed(4321 + cx.rank)
        opt.zero_grad(set_to_none=True)
        loss = model(ids[:B], mask[:B], ids[B:], mask[B:], weights=ones)[0]
        with sync:
            loss.backward()
        opt.step()
        torch.cuda.synchronize()
        out = {i: p.detach().clone() for i, p in enumerate(params) if p.grad is not None}
        opt.zero_grad(set_to_none=True)
        return out

    a = one_step(True)
    arena.check()
    b = one_step(False)
    with torch.no_grad():
        for p, s0 in zip(params, saved):
            p.copy_(s0)
    model.bert.set_dropout_seed(torch.initial_seed() + 977 * cx.rank)
    upd = max((b[i] - saved[i]).abs().max().item() for i in b)
    diffs = torch.cat([(a[i] - b[i]).abs().reshape(-1) for i in b if i not in skip]) / max(upd, 1e-30)
    # (first Adam step: every element moves by ~lr * sign(g); where the rank-summed gradient is pure cancellation noise
    # the two summation orders may disagree on the sign, i.e. differ by up to 2 updates -- a handful of elements)
    st = torch.stack([diffs.max(), diffs.mean(), (diffs > 0.01).float().mean()])
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    return {"largest_update": upd, "max_diff_in_updates": st[0].item(), "mean_diff_in_updates": st[1].item(),
            "fraction_of_elements_off_by_more_than_1pct_of_an_update": st[2].item(),
            "compared": "cdr_adam_multi_peer (gradients read from every rank's arena, sharded update, stores to all "
                        "ranks) vs NCCL all-reduce + replicated cdr_adam_multi; lr 1e-4, one step, max over ranks"}


def multi_gpu_parity(cx, model, sync, batch, ones):
    """One training step's loss and gradients through (a) the peer-memory CLS exchange + GradSync and (b) NCCL
    all-gather / reduce-scatter + an explicit all-reduce of the local gradients; same weights, inputs and dropout masks.
    Returns the relative disagreement (the checks of tests/dist_worker.py #2 / #5 / #6, here on the benchmarked job)."""
    torch, dist, world = cx.torch, cx.dist, cx.world
    from cocodr_b200 import ops
    B = PER_GPU_BATCH
    ids, mask = batch

    def run(peer, use_sync):
        model.bert.set_dropout_seed(4242 + cx.rank)  # both arms draw identical masks
        model.peer_gather = peer
        model.zero_grad(set_to_none=True)
        loss = model(ids[:B], mask[:B], ids[B:], mask[B:], weights=ones)[0]
        if use_sync:
            with sync:
                loss.backward()
        else:
            loss.backward()
            ops.FWD_CALLS.clear()  # (no GradSync exit on this path: do not leave forward counts behind)
            for p in model.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad /= world
        torch.cuda.synchronize()
        return loss.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    was_peer = model.peer_gather
    la, ga = run(was_peer, True)
    stats_sync = dict(sync.stats)
    lb, gb = run(False, False)
    model.peer_gather = was_peer
    model.zero_grad(set_to_none=True)
    model.bert.set_dropout_seed(torch.initial_seed() + 977 * cx.rank)
    worst, worst_name = 0.0, ""
    for n in gb:
        if "key.bias" in n:
            continue
        r = (ga[n] - gb[n]).abs().max().item() / (gb[n].abs().max().item() + 1e-30)
        if r > worst:
            worst, worst_name = r, n
    stats = torch.tensor([abs(la.item() - lb.item()) / (abs(lb.item()) + 1e-30), worst], device=cx.dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    return {"loss_peer_vs_nccl": stats[0].item(), "grad_max_rel": stats[1].item(), "grad_worst_param_rank0": worst_name,
            "compared": "peer-memory CLS exchange + GradSync vs NCCL all-gather/reduce-scatter + explicit all-reduce of the "
                        "local gradients; same weights, batch and dropout masks; max over ranks",
            "gradsync_layers": stats_sync}


# ------------------------------------------------------------------------------------------------ scan (configs[4])
def bench_scan(cx):
    torch = cx.torch
    from cocodr_b200 import scan
    world, rank, dev, peaks = cx.world, cx.rank, cx.dev, cx.peaks
    n_docs, n_q, k, dim = 1_000_000 // world, 1000, 1000, 768
    gs = torch.Generator(device=dev).manual_seed(7 + rank)
    P = torch.randn(n_docs, dim, generator=gs, device=dev, dtype=torch.float16)
    gq = torch.Generator().manual_seed(7)
    Qh = torch.randn(n_q, dim, generator=gq).half().pin_memory()
    Qd = Qh.to(dev)

    def scan_dev(i):
        scan.search_sharded(Qd, P, k, doc_base=rank * n_docs)

    def scan_e2e(i):
        D, I = scan.search_sharded(Qh.to(dev, non_blocking=True), P, k, doc_base=rank * n_docs)
        return D.cpu(), I.cpu()

    for i in range(3):
        scan_dev(i)
    sms = cx.timed(scan_dev, 20) / 20
    sms_e2e = cx.timed(scan_e2e, 5) / 5
    # HBM-bound regime: one pass over the corpus with a 128-query tile
    Q128 = Qd[:128].contiguous()

    def scan128(n):  # n searches in flight, ONE status check (host sync) for all of them, inside the timed region
        scan.check_status([scan.search_async(Q128, P, 100) for _ in range(n)], Q128, P, 100)

    for i in range(2):
        scan128(10)
    hms = cx.timed(lambda i: scan128(10), 5) / 50
    bytes_pass = n_docs * dim * 2
    tf = 2.0 * n_q * n_docs * dim / (sms * 1e-3) / 1e12
    res = {"metric": "corpus-scan queries/s", "value": n_q / (sms * 1e-3), "unit": "q/s",
           "e2e": {"value": n_q / (sms_e2e * 1e-3), "unit": "q/s", "h2d_bytes_per_step": n_q * dim * 2,
                   "d2h_bytes_per_step": n_q * k * 12},
           "config": {"workload": f"{n_docs * world} x {dim} fp16 docs ({n_docs}/GPU), {n_q} queries, k={k}",
                      "ms_per_search": sms},
           "roofline_q1000": {"bound": "tensor", "achieved": tf, "peak": peaks["tensor"], "unit": "TFLOP/s",
                              "frac": tf / peaks["tensor"], "peak_burst": peaks["tensor_burst"],
                              "frac_burst": tf / peaks["tensor_burst"], "note": "per GPU; whole search incl. thresholds + select (+ merge)"},
           "roofline_q128": {"bound": "hbm", "achieved": bytes_pass / (hms * 1e-3) / 1e9, "peak": peaks["hbm"],
                             "unit": "GB/s", "frac": bytes_pass / (hms * 1e-3) / 1e9 / peaks["hbm"],
                             "note": "128 queries, k=100: one pass over the shard; whole search (thresholds + filter pass + select), 10 searches in flight per status check"}}
    del P
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not cx.args.no_cpu:
        # torch fp32 Q @ P^T + topk on the host cores (faiss is absent); bounded: 100 queries x 1M docs
        torch.set_num_threads(os.cpu_count() or 1)
        nq_cpu = 100
        Pc = torch.randn(1_000_000, dim, generator=torch.Generator().manual_seed(7))
        Qc = Qh[:nq_cpu].float()
        t0 = time.perf_counter()
        best = None
        for lo in range(0, Pc.shape[0], 250_000):
            S = Qc @ Pc[lo:lo + 250_000].t()
            d, i = torch.topk(S, k, dim=1)
            i = i + lo
            if best is not None:
                d, j = torch.topk(torch.cat([best[0], d], 1), k, dim=1)
                i = torch.cat([best[1], i], 1).gather(1, j)
            best = (d, i)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": nq_cpu / dt, "unit": "q/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{nq_cpu} queries x 1,000,000 x {dim} fp32 docs, torch matmul + topk (k={k}) in 4 chunks"}
        del Pc
    return res


# ------------------------------------------------------------------------------------------------ inference (a11)
def bench_inference(cx):
    """The reference's InferenceEmbeddingFromStreamDataLoader loop (evaluate/drivers/run_ann_data_gen.py:152-206): per
    batch an H2D copy of (ids int32, mask, type ids, idx), query_emb / body_emb under no_grad, embeddings kept.
    Passage-shaped batches (L = 128) of 128 sequences in the format of evaluate/data/msmarco_data.py GetProcessingFn."""
    torch = cx.torch
    from transformers import BertConfig

    from cocodr_b200 import mining, models
    dev, world = cx.dev, cx.world
    cfg = BertConfig(num_labels=2)
    torch.manual_seed(0)
    model = models.BertDot_NLL_LN(cfg).to(dev).eval()
    Bi, L, n_batches = 128, 128, 16
    g = torch.Generator().manual_seed(99 + cx.rank)
    lens = (torch.randn(n_batches, Bi, generator=g) * 25 + 75).round().clamp(16, L).long()  # MS-MARCO-shaped passages
    batches = []
    for b in range(n_batches):
        ids = torch.randint(1000, cfg.vocab_size, (Bi, L), generator=g, dtype=torch.int32)
        mask = torch.arange(L)[None, :] < lens[b][:, None]
        ids = ids * mask
        ids[:, 0] = 101
        batches.append((ids.pin_memory(), mask.pin_memory(), torch.zeros(Bi, L, dtype=torch.uint8).pin_memory(),
                        torch.arange(b * Bi, (b + 1) * Bi)))
    dev_batches = [tuple(t.to(dev) for t in bt) for bt in batches]
    mining.encode(model, dev_batches[:2], is_query=False)
    mining.encode(model, dev_batches[:2], is_query=False, trim=False)

    def resident(i):  # mode B of SURVEY 8d: sequences re-batched by real length (mining.encode, trim=True)
        mining.encode(model, dev_batches, is_query=False)

    def resident_padded(i):  # mode A: every sequence run at the padded length L, as the reference does
        mining.encode(model, dev_batches, is_query=False, trim=False)

    def e2e(i):
        emb, ids = mining.encode(model, batches, is_query=False)
        return emb.float().cpu(), ids.cpu()  # what the reference hands to numpy / pickle

    ms_pad = cx.timed(resident_padded, 3) / 3
    ms = cx.timed(resident, 3) / 3
    ms_e2e = cx.timed(e2e, 2) / 2
    n_seq = Bi * n_batches
    fl = fwd_flops_per_seq() * n_seq
    real_frac = float(lens.sum()) / (n_seq * L)
    # queries: clamp(round(N(8, 3)), 4, 64) real tokens in 64 positions
    Lq = 64
    qlens = (torch.randn(n_batches, Bi, generator=g) * 3 + 8).round().clamp(4, Lq).long()
    qb = []
    for b in range(n_batches):
        ids = torch.randint(1000, cfg.vocab_size, (Bi, Lq), generator=g, dtype=torch.int32)
        mask = torch.arange(Lq)[None, :] < qlens[b][:, None]
        ids = ids * mask
        ids[:, 0] = 101
        qb.append((ids.to(dev), mask.to(dev), torch.arange(b * Bi, (b + 1) * Bi).to(dev)))
    mining.encode(model, qb[:2], is_query=True)
    mining.encode(model, qb[:2], is_query=True, trim=False)
    ms_q_pad = cx.timed(lambda i: mining.encode(model, qb, is_query=True, trim=False), 3) / 3
    ms_q = cx.timed(lambda i: mining.encode(model, qb, is_query=True), 3) / 3
    res = {"metric": "embedding-inference sequences/s", "value": n_seq * world / (ms * 1e-3), "unit": "sequences/s",
           "config": {"workload": f"BERT-base body_emb under no_grad, {n_batches} batches x {Bi} passages, L={L}, "
                                  "MS-MARCO-shaped lengths (SURVEY 8d mode B: re-batched by real length, padded to a "
                                  "multiple of 16 per group), fp16 embeddings kept in HBM",
                      "real_token_fraction": real_frac},
           "padded": {"value": n_seq * world / (ms_pad * 1e-3), "unit": "sequences/s",
                      "note": "mode A: every sequence run at L = 128 as the reference does (same embeddings)"},
           "queries": {"value": n_seq * world / (ms_q * 1e-3), "padded_value": n_seq * world / (ms_q_pad * 1e-3),
                       "unit": "sequences/s", "note": "query_emb, 4-16 real tokens in 64 positions; trimmed vs padded"},
           "e2e": {"value": n_seq * world / (ms_e2e * 1e-3), "unit": "sequences/s",
                   "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in batches[0]) * n_batches,
                   "d2h_bytes_per_step": n_seq * 768 * 4 + n_seq * 8},
           "roofline": {"bound": "tensor", "achieved": fl / (ms_pad * 1e-3) / 1e12, "peak": cx.peaks["tensor"], "unit": "TFLOP/s",
                        "frac": fl / (ms_pad * 1e-3) / 1e12 / cx.peaks["tensor"],
                        "useful_tflops_trimmed": fl * real_frac / (ms * 1e-3) / 1e12,
                        "note": "padded run: padded-length model FLOPs (22.35 GFLOP / sequence) over the whole loop; "
                                "useful_tflops_trimmed counts real tokens only over the trimmed run"}}
    if cx.rank == 0 and world == 1 and not cx.args.no_cpu:
        from oracle import bert_ref
        torch.set_num_threads(os.cpu_count() or 1)
        c = bert_ref.make_config()
        st = bert_ref.synth_state(c, 0)
        ids, mask = bert_ref.synth_batch(16, L, c["vocab"], 3)
        with torch.no_grad():
            bert_ref.cls_embedding(st, ids[:4], mask[:4], c)
            t0 = time.perf_counter()
            bert_ref.cls_embedding(st, ids, mask, c)
            dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 16 / dt, "unit": "sequences/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "one no_grad forward of 16 sequences (BERT-base, L=128, fp32), oracle port"}
    del model, dev_batches
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ iDRO (configs[2])
def bench_idro(cx):
    """configs[2]: the contrastive step with iDRO group weights, G = 50 (ANCE/model/dro_loss.py:216-254; README
    hyper-parameters alpha .25, ema .1, rho .05, eps .01).  Three arms: the reference-exact triplet model
    (BertDot_NLL_LN: group gradients through the grouped wgrad K11), and the in-batch head with both group-gradient
    views (models.BertDot_InBatch_NLL_LN.idro_group_grads).  At N > 1 the [G, P_last] matrix is reduce-scattered and the
    Gram is taken per shard (dro_loss.iDROLoss._gram) -- SURVEY 8(e) e3 on hardware."""
    torch, dist = cx.torch, cx.dist
    from transformers import BertConfig

    from cocodr_b200 import models
    from cocodr_b200 import optim as cdr_optim
    from cocodr_b200.gradsync import GradSync
    dev, world, rank, args = cx.dev, cx.world, cx.rank, cx.args
    B, L, G = PER_GPU_BATCH, SEQ_LEN, 50
    cfg = BertConfig(hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout, num_labels=2)
    g = torch.Generator().manual_seed(4321 + rank)
    ids, mask = (t.to(dev) for t in synth_ids(torch, g, 3 * B, L, cfg.vocab_size))
    gid = torch.randint(0, G, (B,), generator=g).to(dev)
    ones = torch.ones(B, device=dev)
    P_last = 21_263_616
    out = {"metric": "contrastive step with iDRO group weights", "unit": "samples/s",
           "config": {"workload": f"BERT-base L={L}, {B} samples/GPU, G={G} groups (uniform ids), N={world}",
                      "group_gradient_bytes": G * P_last * 4}}

    def arm(kind, mode=None):
        torch.manual_seed(0)
        cls = models.BertDot_NLL_LN if kind == "triplet" else models.BertDot_InBatch_NLL_LN
        model = cls(cfg).to(dev).train()
        if mode is not None:
            model.idro_group_grads = mode
        model.add_group_loss(types.SimpleNamespace(model_size="base", local_rank=cx.local), G, "idro", 0.25, 0.01, 0.1, 0.05)
        sync = None
        if world > 1:
            for p_ in model.parameters():
                dist.broadcast(p_.data, 0)
            sync = GradSync(model)
        opt = cdr_optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=5e-6, eps=1e-8, weight_decay=0.01,
                              semantics="torch").attach_shadows(model)

        def step(i):
            if kind == "triplet":
                loss = model(ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], ids[2 * B:], mask[2 * B:], group_ids=gid)[0]
            else:
                loss = model(ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], group_ids=gid, weights=ones)[0]
            opt.zero_grad(set_to_none=True)
            if sync is not None:
                with sync:
                    loss.backward()
            else:
                loss.backward()
            opt.step()
            return loss

        eager_step = step
        graphed = None
        capturable = (kind == "triplet" or mode == "own-pair") and not args.no_graph  # (per-group backwards read the present groups on the host)
        if capturable:
            from cocodr_b200.graph import GraphedTrainStep
            if kind == "triplet":
                inputs = (ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], ids[2 * B:], mask[2 * B:], True, gid)
            else:
                inputs = (ids[:B], mask[:B], ids[B:2 * B], mask[B:2 * B], None, None, True, gid, ones)
            graphed = GraphedTrainStep(model, opt, inputs, backward_ctx=sync)

            def step(i):  # noqa: F811
                return graphed(*[t for t in inputs if torch.is_tensor(t)])
        for i in range(3):
            step(i)
        n = 20 if capturable else 4
        ms = cx.timed(step, n) / n
        ms_e2e = cx.timed(lambda i: step(i).item(), n) / n
        r = {"ms_per_step": ms, "value": B * world / (ms * 1e-3), "unit": "triplets/s" if kind == "triplet" else "pairs/s",
             "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                     "note": "loss read back every step; ids resident (the headline e2e covers the H2D of ids)"},
             "launch_mode": "one CUDA graph per step (meters accumulate on the device)" if graphed is not None else
                            "eager (one partial backward per present group: the present groups are read on the host)",
             "h_fun_minmax": [model.loss.h_fun.min().item(), model.loss.h_fun.max().item()]}
        if kind == "triplet" or mode == "own-pair":
            ops_ms = cx.op_breakdown(lambda: eager_step(0))
            gg = ops_ms.get("cdr_gemm_grouped")
            if gg and gg["ms"] > 0:
                bw = G * P_last * 4 / (gg["ms"] * 1e-3) / 1e9
                r["roofline"] = {"bound": "hbm", "kernel": "gemm_tcgen05_kernel<F32_GROUPED> (K11 grouped wgrad)",
                                 "achieved": bw, "peak": cx.peaks["hbm"], "unit": "GB/s", "frac": bw / cx.peaks["hbm"],
                                 "note": f"{gg['calls']} launches write the [G, P_last] fp32 matrix once (4.25 GB algorithmic), event-timed in an eager step"}
            r["op_breakdown_ms_per_step"] = dict(list(ops_ms.items())[:10])
        if graphed is not None:
            graphed.graph.reset()
        del model, opt, sync, graphed, step, eager_step
        torch.cuda.empty_cache()
        return r

    out["triplet_reference_exact"] = _safe("idro.triplet", lambda: arm("triplet"))
    out["inbatch_own_pair"] = _safe("idro.own_pair", lambda: arm("inbatch", "own-pair"))
    out["inbatch_local_batch"] = _safe("idro.local_batch", lambda: arm("inbatch", "local-batch"))
    t = out["triplet_reference_exact"]
    if isinstance(t, dict) and "value" in t:
        out["value"] = t["value"]
        out["unit"] = t["unit"]
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = _safe("idro.cpu", cpu_idro_sample)
    return out


def cpu_idro_sample():
    """Oracle iDRO step (per-group autograd, dro_loss.py:192-254) on 8 triplets / 4 groups, L = 64, host cores."""
    import torch

    from oracle import bert_ref, heads_ref
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = bert_ref.make_config()
    leaf = {k: v.clone().requires_grad_(True) for k, v in bert_ref.synth_state(cfg, 0).items()}
    B, L, G = 8, 64, 50
    bt = [bert_ref.synth_batch(B, L, cfg["vocab"], 70 + k, full=True) for k in range(3)]
    gid = torch.arange(B) % 4
    names = heads_ref.idro_param_names(["bert." + n for n in leaf], "base")
    params = [leaf[n[5:]] for n in names]
    t0 = time.perf_counter()
    embs = [bert_ref.cls_embedding(leaf, i, m, cfg) for i, m in bt]
    losses = heads_ref.pair_nll(*embs)[0]
    robust, _, _, _ = heads_ref.idro_forward(losses, gid, params, torch.ones(G), G, 0.25, 0.1, 0.05, 0.01)
    robust.backward()
    dt = time.perf_counter() - t0
    return {"value": B / dt, "unit": "triplets/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one step of {B} triplets (L={L}), 4 groups present of G={G}: forward, 4 partial backwards, Gram, h update, backward"}


# ------------------------------------------------------------------------------------------------ COCO (configs[3])
def bench_coco(cx):
    """configs[3]: BERT-large, L = 256, per-GPU batch 32 documents = 64 spans, Condenser head (2 layers, skip_from 6,
    late_mlm) + MLM losses + sequence-contrastive loss over the gathered spans (COCO/modeling.py:192-235)."""
    torch, dist = cx.torch, cx.dist
    from transformers import BertConfig, BertForMaskedLM

    from cocodr_b200 import modeling
    from cocodr_b200 import optim as cdr_optim
    from cocodr_b200.gradsync import GradSync
    dev, world, rank, args = cx.dev, cx.world, cx.rank, cx.args
    docs, L = 32, 256
    n = 2 * docs
    cfg = BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                     hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout)
    torch.manual_seed(0)
    lm = BertForMaskedLM(cfg)
    model = modeling.CoCondenserForPretraining(
        lm, types.SimpleNamespace(n_head_layers=2, skip_from=6, late_mlm=True), types.SimpleNamespace(train_method="coco"),
        types.SimpleNamespace(per_device_train_batch_size=docs, local_rank=(cx.local if world > 1 else -1))).to(dev).train()
    model._backbone()
    sync = None
    if world > 1:
        for p_ in model.parameters():
            dist.broadcast(p_.data, 0)
        sync = GradSync(model)
    opt = cdr_optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, eps=1e-8, weight_decay=0.01,
                          semantics="torch").attach_shadows(model)
    g = torch.Generator().manual_seed(77 + rank)
    ids_h, mask_h = synth_ids(torch, g, n, L, cfg.vocab_size)
    lab_h = torch.where(torch.rand(n, L, generator=g) < 0.15, ids_h, torch.full_like(ids_h, -100))
    host = tuple(t.pin_memory() for t in (ids_h, mask_h, lab_h))
    ids, mask, lab = (t.to(dev) for t in host)

    def call(i_, m_, l_):
        return model({"input_ids": i_, "attention_mask": m_}, l_)

    def eager_step(i_, m_, l_):
        loss = call(i_, m_, l_)
        opt.zero_grad(set_to_none=True)
        if sync is not None:
            with sync:
                loss.backward()
        else:
            loss.backward()
        opt.step()
        return loss

    graphed = None
    if not args.no_graph:
        # masked rows gathered into a fixed-size buffer (25 % of the positions for 15 % masking): no host sync sizes the
        # MLM GEMMs, so the whole step -- NCCL gather and gradient all-reduces included -- is one CUDA graph
        from cocodr_b200.graph import GraphedTrainStep
        model.mlm_capacity = 0.25
        graphed = GraphedTrainStep(call, opt, (ids, mask, lab), backward_ctx=sync)
    step_on = graphed if graphed is not None else eager_step
    for i in range(3):
        step_on(ids, mask, lab)
    nst = 16
    with ClockSampler(cx.local) as clocks:
        ms = cx.timed(lambda i: step_on(ids, mask, lab), nst, clocks) / nst
    ms_e2e = cx.timed(lambda i: step_on(*((t if graphed is not None else t.to(dev, non_blocking=True)) for t in host)).item(), nst) / nst
    overflow = int(model.mlm_overflow) if model.mlm_capacity is not None else 0
    eager_step(ids, mask, lab)
    roof = cx.gemm_roofline(lambda: eager_step(ids, mask, lab), "per-launch CUDA events over one eagerly launched step; same kernels/shapes as the timed graph")
    res = {"metric": "COCO pre-training spans/s", "value": n * world / (ms * 1e-3), "unit": "spans/s", "ms_per_step": ms,
           "config": {"workload": f"BERT-large L={L}, {docs} docs = {n} spans/GPU, n_head_layers=2 skip_from=6 late_mlm, "
                                  f"15% MLM labels, N={world}" + (", all_gather of CLS spans (reference convention)" if world > 1 else ""),
                      "dropout": f"c_head p={args.dropout} (backbone in eval(), as COCO/modeling.py:198)",
                      "launch_mode": "one CUDA graph per step (mlm_capacity = 0.25: fixed-size masked-row gather)" if graphed is not None else "eager",
                      "mlm_overflow_rows": overflow},
           "e2e": {"value": n * world / (ms_e2e * 1e-3), "unit": "spans/s",
                   "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host), "d2h_bytes_per_step": 4},
           "roofline": roof, "clocks": clocks.summary()}
    if graphed is not None:
        graphed.graph.reset()
    del model, opt, sync, lm, graphed, step_on
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ stock GPU path
def bench_gpu_baseline(cx):
    """The headline step through the stock library path on the same GPU: HF BertModel (SDPA attention), torch.autocast
    fp16, one 128-sequence pass (its fastest arrangement; the reference runs the towers sequentially), in-batch CE in
    fp32, static loss scale, torch fused AdamW -- captured into a CUDA graph like our arm."""
    torch = cx.torch
    from transformers import BertConfig, BertModel
    dev, args = cx.dev, cx.args
    B, L = PER_GPU_BATCH, SEQ_LEN
    cfg = BertConfig(hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    model = BertModel(cfg, add_pooling_layer=False).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-6, fused=True, capturable=True)
    g = torch.Generator().manual_seed(1234)
    ids, mask = (t.to(dev) for t in synth_ids(torch, g, 2 * B, L, cfg.vocab_size))
    tgt = torch.arange(B, device=dev)

    def fwd_bwd():
        with torch.autocast("cuda", dtype=torch.float16):
            e = model(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0]
        loss = torch.nn.functional.cross_entropy(e[:B].float() @ e[B:].float().t(), tgt)
        (loss * 1024.0).backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            fwd_bwd()
            opt.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_loss = fwd_bwd()
        opt.step()
    for _ in range(3):
        graph.replay()
    est = cx.timed(lambda i: graph.replay(), 5) / 5
    n = max(20, math.ceil(MIN_REGION_MS / est))
    with ClockSampler(cx.local) as clocks:
        ms = cx.timed(lambda i: graph.replay(), n, clocks) / n
    res = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "timed_steps": n, "loss": static_loss.item(),
           "arm": "HF BertModel (sdpa) under torch.autocast(fp16), one 128-sequence pass, fp32 in-batch CE, static loss scale, "
                  f"torch fused AdamW, dropout p={args.dropout}; whole step in one CUDA graph",
           "clocks": clocks.summary()}
    graph.reset()
    del model, opt, graph
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-scan", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="hidden / attention dropout probability (HF default 0.1; 0 = off)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline metric (+ scan) only: skip the inference / iDRO / COCO / stock-GPU sub-metrics")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--ddp", action="store_true", help="N > 1: wrap in DistributedDataParallel instead of GradSync")
    ap.add_argument("--torch-adamw", action="store_true", help="use torch.optim.AdamW(fused=True) instead of cdr_adam_multi")
    ap.add_argument("--nccl-grads", action="store_true", help="N > 1: NCCL all-reduce of the gradients (GradSync) + the same AdamW on every rank instead of the peer-memory optimizer")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: NCCL all-gather of the CLS embeddings instead of the fused peer-memory push")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--nccl-ctas", type=int, default=0, help="N > 1: cap NCCL at this many CTAs and leave that many SMs free of persistent kernels (0 = off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
