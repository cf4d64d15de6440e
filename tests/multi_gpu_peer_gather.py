"""Multi-GPU check of the fused all-gather over peer memory (not collected by pytest: needs >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_peer_gather.py

Every rank runs the fused op on its own shard with PeerDurationGather active, through the device entry and
through the chunked host-buffer entry, and compares the gathered buffer with the NCCL all-gather of the
local durations (bit-exact), on ragged LJSpeech-shape utterances."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from art_tts_b200 import monotonic_align
from art_tts_b200.distributed import PeerDurationGather

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl")
B, F, T_x, T_y = 300, 80, 190, 872
rng = np.random.default_rng(100 + rank)
t_x = rng.integers(60, T_x + 1, B).astype(np.int32)
t_y = np.minimum(870, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
order = np.argsort(-(t_x.astype(np.int64) * t_y), kind="stable")
t_x, t_y = t_x[order], t_y[order]
torch.manual_seed(rank)
mu_x = torch.randn(B, F, T_x, device=dev)
y = torch.randn(B, F, T_y, device=dev)
tx_d, ty_d = torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev)
assert PeerDurationGather.supported(B, F, T_x, T_y)
peer = PeerDurationGather(B, T_x, dev)
want = torch.empty(world * B, T_x, dtype=torch.int32, device=dev)
ok = True
for it in range(3):
    peer.all.fill_(-7)
    torch.cuda.synchronize()
    dist.barrier()
    path, dur = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx_d, ty_d)
    got = peer.finish()
    dist.all_gather_into_tensor(want, dur)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(got, want))
# chunked host-buffer entry: rows of every chunk land at the right offset
h = [mu_x.cpu().pin_memory(), y.cpu().pin_memory(), torch.from_numpy(t_x).pin_memory(), torch.from_numpy(t_y).pin_memory()]
peer.all.fill_(-7)
torch.cuda.synchronize()
dist.barrier()
out = monotonic_align.maximum_path_from_prior_host(h[0], h[1], h[2], h[3], dev, chunk=64)
got = peer.finish()
dist.all_gather_into_tensor(want, out[1])
torch.cuda.synchronize()
ok_host = bool(torch.equal(got, want))
peer.close()
# switched off again: nothing may be written
peer.all.fill_(-7)
torch.cuda.synchronize()
monotonic_align.maximum_path_from_prior(mu_x, None, y, tx_d, ty_d)
torch.cuda.synchronize()
ok_off = bool((peer.all == -7).all())
print(f"rank {rank}: device entry {'ok' if ok else 'MISMATCH'}, host entry {'ok' if ok_host else 'MISMATCH'}, "
      f"off {'ok' if ok_off else 'WRITES'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and ok_host and ok_off) else 1)
