"""Batch-sharded MAS over the GPUs of one box.

MAS is embarrassingly parallel over utterances (core.pyx:44-45; the reference's own DDP
script runs one independent call per rank, train_v1_1_dist.py:249), so the data path has NO
collective.  The only exchange is the re-assembly of results: one all-gather of the int32
durations [B/g, T_x] (and optionally scores) over NCCL/NVLink; dense paths are rebuilt
locally from durations with the generate_path kernel (exact, utils.py:26-43) instead of
moving 4 B/cell across the links.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n utterances for `rank` (cf. balance_batch.py:148-150);
    the first n % world ranks get one extra utterance."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def length_bucket_order(t_x: torch.Tensor, t_y: torch.Tensor, world: int = 1) -> torch.Tensor:
    """Permutation that sorts utterances by work (t_x*t_y, longest first) and deals them
    round-robin to ranks, so that every contiguous shard holds the same length mix and each
    GPU starts its longest utterances first (LPT).  Returns int64 indices [B]."""
    work = t_x.to(torch.int64) * t_y.to(torch.int64)
    order = torch.argsort(work, descending=True, stable=True)
    if world <= 1:
        return order
    n = order.numel()
    parts = [order[r::world] for r in range(world)]
    # shard_bounds gives the first n%world ranks one more element: same as the strided deal
    assert sum(p.numel() for p in parts) == n
    return torch.cat(parts)


def all_gather_rows(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None
                    ) -> torch.Tensor:
    """All-gather row-shards of unequal size (differing by at most one row) into [n_total, ...].
    Uses one all_gather_into_tensor on padded shards (NCCL all-gather over NVLink on GPUs)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    lo, hi = shard_bounds(n_total, rank, world)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    pad = local
    if local.shape[0] != per:
        pad = local.new_zeros((per,) + tuple(local.shape[1:]))
        pad[: local.shape[0]] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if n_total == world * per:
        return out
    pieces = []
    for r in range(world):
        rlo, rhi = shard_bounds(n_total, r, world)
        pieces.append(out[r * per: r * per + (rhi - rlo)])
    return torch.cat(pieces)


class DurationGatherer:
    """Asynchronous, double-buffered all-gather of per-rank durations [B_local, T_x] int32.

    The gather of step i runs on a side stream while the MAS kernel of step i+1 runs on the
    caller's stream (the only exchange of the sharded path is this re-assembly).
    `gather(dur)` returns the buffer that will hold all ranks' durations once `wait()` (or the
    next-but-one `gather`) has been ordered after it."""

    def __init__(self, b_local: int, t_x: int, device, group: Optional[dist.ProcessGroup] = None,
                 depth: int = 2):
        self.group = group
        self.world = dist.get_world_size(group)
        self.bufs = [torch.empty(self.world * b_local, t_x, dtype=torch.int32, device=device)
                     for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=device)
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.i = 0

    def gather(self, dur: torch.Tensor) -> torch.Tensor:
        slot = self.i % len(self.bufs)
        self.i += 1
        main = torch.cuda.current_stream(dur.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)           # durations produced
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.bufs[slot], dur, group=self.group)
            self.done[slot].record(self.stream)
        dur.record_stream(self.stream)          # keep the allocator from recycling it early
        return self.bufs[slot]

    def wait(self) -> None:
        """Order the caller's stream after every gather issued so far."""
        torch.cuda.current_stream().wait_stream(self.stream)


class PeerDurationGather:
    """All-gather of the durations done BY the fused kernel over NVLink peer memory.

    Every rank owns `slots` (default 2) [world * b_local, t_x_max] int32 buffers in symmetric memory
    (torch.distributed._symmetric_memory), each mapped into every process.  `desc()` describes the
    buffer set of the next step (a `mas_peer_gather`, passed WITH the call -- the library keeps no
    peer state); the backtrack warp of the tensor-core kernel stores each utterance's durations into
    row `rank * b_local + b` of every rank's buffer.  No collective kernel runs: after the MAS kernel
    the ranks only meet in `finish()` (a symmetric-memory barrier on the caller's stream, ~7 us at 2
    GPUs against 25-44 us for the NCCL all-gather), which returns the gathered [world*b_local, t_x]
    view for that step.

    Buffer lifetime (the write-after-read hazard of a single buffer): step i uses slot i % slots.  A
    rank's kernel of step i+slots can only start after that rank passed the barrier of step
    i+slots-1, i.e. after EVERY rank enqueued-and-ran everything that precedes its own barrier
    i+slots-1 on the stream it calls finish() on.  So the tensor returned by finish() for step i is
    valid for consumers enqueued on that same stream before the rank's finish() of step i+slots-1
    (with the default two slots: before its next finish()).  Consumers on other streams must make
    the calling stream wait for them before the next finish() (or copy the rows out).
    Only the tensor-core engine of maximum_path_from_prior writes peer memory (`supported()`);
    everything else must use the NCCL gathers above."""

    def __init__(self, b_local: int, t_x: int, device, group: Optional[dist.ProcessGroup] = None,
                 slots: int = 2, channel: int = 0, frame_idx_len: int = 0):
        import ctypes

        import torch.distributed._symmetric_memory as symm

        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("PeerDurationGather: at most 16 ranks (one NVLink domain)")
        if slots < 2:
            raise ValueError("PeerDurationGather needs >= 2 slots (see the class docstring)")
        self.b_local, self.t_x, self.slots, self.channel = b_local, t_x, slots, channel
        self.t_y = int(frame_idx_len)
        self.step = 0
        rows = self.world * b_local
        # ONE symmetric allocation (one rendezvous) carved into the slots
        per = rows * t_x + rows * self.t_y
        self.buf = symm.empty((slots * per,), dtype=torch.int32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self.all, self.all_fi, self._dptr, self._fptr = [], [], [], []
        for s in range(slots):
            o = s * per
            self.all.append(self.buf[o:o + rows * t_x].view(rows, t_x))
            self._dptr.append((ctypes.c_uint64 * self.world)(*[p + 4 * o for p in base]))
            if self.t_y:
                o2 = o + rows * t_x
                self.all_fi.append(self.buf[o2:o2 + rows * self.t_y].view(rows, self.t_y))
                self._fptr.append((ctypes.c_uint64 * self.world)(*[p + 4 * o2 for p in base]))
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=self.channel)

    @staticmethod
    def supported(B: int, F: int, T_x: int, T_y: int, flags: int = 0) -> bool:
        from . import _lib
        return bool(_lib.load().mas_peer_durations_supported(B, F, T_x, T_y, flags))

    def desc(self, b_call: Optional[int] = None):
        """`mas_peer_gather` for the step about to be launched (slot = step % slots).  `b_call`:
        utterances of the call when fewer than b_local (a ragged last shard)."""
        ct = self._ct
        s = self.step % self.slots
        d = self._lib.PeerGatherDesc()
        d.n_peers = self.world
        d.durations_ptrs = ct.cast(self._dptr[s], ct.POINTER(ct.c_uint64))
        d.row0 = self.rank * self.b_local
        d.rows = self.b_local if b_call is None else int(b_call)
        d.row_stride = self.t_x
        if self.t_y:
            d.frame_idx_ptrs = ct.cast(self._fptr[s], ct.POINTER(ct.c_uint64))
            d.frame_idx_stride = self.t_y
        else:
            d.frame_idx_ptrs = None
            d.frame_idx_stride = 0
        return d

    def finish(self):
        """Call after maximum_path_from_prior(..., peer=self.desc()) on the same stream: once every
        rank's kernel of this step has completed (its peer stores are then visible), returns the
        gathered durations of the step (and the frame index when frame_idx_len was given); advances to
        the next slot.  See the class docstring for how long the returned view stays valid."""
        s = self.step % self.slots
        self.hdl.barrier(channel=self.channel)
        self.step += 1
        return (self.all[s], self.all_fi[s]) if self.t_y else self.all[s]

    def close(self) -> None:
        pass   # nothing process-wide to undo: the description travels with every call


def maximum_path_from_prior_sharded(mu_x, y, t_x, t_y, *, group=None, rebuild_path=False):
    """Every rank holds the FULL batch description (or just its shard, see below), computes MAS
    for its contiguous shard and all-gathers durations.  Returns (durations [B,T_x] int32 for
    the whole batch, local path for the shard, optional rebuilt global path)."""
    from . import monotonic_align, utils
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = mu_x.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    path, dur = monotonic_align.maximum_path_from_prior(mu_x[lo:hi], None, y[lo:hi], t_x[lo:hi],
                                                        t_y[lo:hi])
    dur_all = all_gather_rows(dur, B, group)
    full = None
    if rebuild_path:
        full = utils.generate_path_lengths(dur_all, t_x, t_y, y.shape[2], out_dtype=mu_x.dtype)
    return dur_all, path, full
