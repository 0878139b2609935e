import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm_mem.empty((world * 4, 8), dtype=torch.float32, device=f"cuda:{rank}")
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad", [hex(p) for p in hdl.signal_pad_ptrs], hdl.signal_pad_size, flush=True)
t.fill_(-1)
dist.barrier(); torch.cuda.synchronize()
# write my block into every peer's buffer through the peer view
for r in range(world):
    peer = hdl.get_buffer(r, (world * 4, 8), torch.float32)
    peer[rank * 4:(rank + 1) * 4] = rank + 1
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
print(rank, "rows", t[:, 0].tolist(), flush=True)
print(rank, "multicast_ptr", getattr(hdl, "multicast_ptr", None), flush=True)
dist.destroy_process_group()
