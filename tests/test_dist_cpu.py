"""CPU, world_size 2 over gloo: the multi-GPU host logic of the hot path.

  * AllGatherWithGrad (models.py): per-rank loss / gradients equal the single-process values on the
    concatenated batch once DDP's mean over ranks is applied (SURVEY.md A.2/A.3 contract)
  * iDROLoss._gram: reduce-scatter of column shards + local Gram + all-reduce == Gram of the all-reduced
    [G, P] matrix (what dro_loss.py:232-237 computes); the CUDA Gram kernel is replaced by a torch stand-in
    here because only the sharding / collective logic is under test
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


@pytest.mark.parametrize("worker,port", [(_gather_worker, 29641), (_gram_worker, 29643)])
def test_world2_gloo(worker, port):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)
