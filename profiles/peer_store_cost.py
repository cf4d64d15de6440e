"""Weak-scaling step (B=1024 per GPU) under torchrun: where do the extra microseconds against one GPU come from?
   (a) kernel alone, no peer description            (b) kernel with peer stores to all ranks, no barrier
   (c) kernel + peer stores + symmetric-memory barrier (the shipping step)      (d) kernel + NCCL all-gather in line
   torchrun --nproc-per-node N profiles/peer_store_cost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from art_tts_b200 import monotonic_align, _lib
from art_tts_b200.distributed import PeerDurationGather
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
B, T_X, T_Y, F = 1024, bench.T_X, bench.T_Y, bench.N_FEATS
tx_np, ty_np = bench.make_lengths(B, 1000 + rank)
tx, ty = torch.from_numpy(tx_np).to(dev), torch.from_numpy(ty_np).to(dev)
g = torch.Generator(device=dev).manual_seed(rank)
mu_x = torch.randn(B, F, T_X, device=dev, generator=g)
y = torch.randn(B, F, T_Y, device=dev, generator=g)
peer = PeerDurationGather(B, T_X, dev)
dur_all = torch.empty(world * B, T_X, dtype=torch.int32, device=dev)


def step(mode):
    if mode == "alone":
        monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty)
    elif mode == "stores":
        monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, peer=peer.desc())
        peer.step += 1                  # next slot, no barrier
    elif mode == "stores+barrier":
        monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, peer=peer.desc())
        peer.finish()
    else:
        _, d = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty)
        dist.all_gather_into_tensor(dur_all, d)


def timed(mode, n=40):
    for _ in range(5):
        step(mode)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(mode)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return float(mx.item()), float(mn.item())


for mode in ("alone", "stores", "stores+barrier", "nccl", "alone"):
    mx, mn = timed(mode)
    if rank == 0:
        print(f"N={world} {mode:15s} max over ranks {mx:.4f} ms   min over ranks {mn:.4f} ms", flush=True)
dist.destroy_process_group()
