"""Multi-GPU worker of tests/test_gpu_multi.py (>= 2 GPUs; launched under torch.distributed.run):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_peer_gather.py

Every rank runs the fused op on its own shard with the fused peer-memory all-gather
(distributed.PeerDurationGather, double-buffered, description passed with every call), through the
device entry and through the chunked host-buffer entry, and compares what landed in its buffer with
the NCCL all-gather of the local durations (bit-exact).  The steps run BACK TO BACK -- no host
synchronisation and no extra barrier between them, different inputs every step, ranks deliberately
skewed -- so a write-after-read race across ranks (a single-buffered gather) would show up as stale
or mixed rows."""
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
dist.init_process_group("nccl", device_id=dev)
B, F, T_x, T_y = 300, 80, 190, 872
NSETS, STEPS = 3, 24
sets = []
for k in range(NSETS):
    rng = np.random.default_rng(100 + 17 * k + rank)
    t_x = rng.integers(60, T_x + 1, B).astype(np.int32)
    t_y = np.minimum(870, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
    order = np.argsort(-(t_x.astype(np.int64) * t_y), kind="stable")
    t_x, t_y = t_x[order], t_y[order]
    torch.manual_seed(1000 * k + rank)
    sets.append((torch.randn(B, F, T_x, device=dev), torch.randn(B, F, T_y, device=dev),
                 torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev), t_x, t_y))
assert PeerDurationGather.supported(B, F, T_x, T_y)
peer = PeerDurationGather(B, T_x, dev, frame_idx_len=T_y)
want = torch.empty(world * B, T_x, dtype=torch.int32, device=dev)
want_fi = torch.empty(world * B, T_y, dtype=torch.int32, device=dev)
bad = torch.zeros(3, dtype=torch.int64, device=dev)
spin = torch.empty(1 << 22, device=dev)
for it in range(STEPS):
    mu_x, y, tx_d, ty_d, _, _ = sets[it % NSETS]
    if (it + rank) % 3 == 0:          # skew the ranks: this one is late into the step
        for _ in range(4):
            spin.normal_()
    path, dur, fi = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx_d, ty_d, return_frame_idx=True,
                                                            peer=peer.desc())
    got, got_fi = peer.finish()
    # consumers on the same stream (the lifetime contract of PeerDurationGather)
    dist.all_gather_into_tensor(want, dur)
    dist.all_gather_into_tensor(want_fi, fi)
    bad[0] += (got != want).sum()
    bad[1] += (got_fi != want_fi).sum()
# chunked host-buffer entry: rows of every chunk land at the right offset
mu_x, y, tx_d, ty_d, t_x, t_y = sets[0]
h = [mu_x.cpu().pin_memory(), y.cpu().pin_memory(), torch.from_numpy(t_x).pin_memory(), torch.from_numpy(t_y).pin_memory()]
out = monotonic_align.maximum_path_from_prior_host(h[0], h[1], h[2], h[3], dev, chunk=64, peer=peer.desc())
got, _ = peer.finish()
dist.all_gather_into_tensor(want, out[1])
bad[2] += (got != want).sum()
# no description: nothing may be written to the peer buffers
torch.cuda.synchronize()
dist.barrier()
peer.buf.fill_(-7)
torch.cuda.synchronize()
dist.barrier()
monotonic_align.maximum_path_from_prior(mu_x, None, y, tx_d, ty_d)
torch.cuda.synchronize()
dist.barrier()
ok_off = bool((peer.buf == -7).all())
b = bad.cpu().tolist()
print(f"rank {rank}: device entry {'ok' if b[0] == 0 else f'MISMATCH({b[0]})'}, frame idx "
      f"{'ok' if b[1] == 0 else f'MISMATCH({b[1]})'}, host entry {'ok' if b[2] == 0 else f'MISMATCH({b[2]})'}, "
      f"off {'ok' if ok_off else 'WRITES'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (sum(b) == 0 and ok_off) else 1)
