"""Probe: does torch symmetric memory (peer pointers over NVLink) work on this box?  torchrun, >= 2 GPUs."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
t = symm.empty((world, 16), dtype=torch.int32, device="cuda")
t.zero_()
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
hdl.barrier(channel=0)
for p in range(world):   # write my rank into row `rank` of every peer's buffer
    peer = hdl.get_buffer(p, (world, 16), torch.int32)
    peer[rank].fill_(rank + 1)
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "rows", t[:, 0].tolist(), flush=True)
# timing of the barrier alone
for _ in range(10):
    hdl.barrier(channel=0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier(channel=0)
e1.record()
torch.cuda.synchronize()
print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
dist.destroy_process_group()
