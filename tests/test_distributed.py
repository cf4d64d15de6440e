"""CPU (gloo, world_size 2): the host-side logic of the batch-sharded path -- shard bounds,
length bucketing, and the all-gather of durations with unequal shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from art_tts_b200 import distributed as D


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_partition_every_batch():
    for n in (0, 1, 7, 8, 9, 1024, 1027):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_length_bucket_order_is_lpt_and_balanced():
    g = torch.Generator().manual_seed(0)
    t_x = torch.randint(60, 191, (1024,), generator=g)
    t_y = torch.minimum(torch.tensor(870), 4 * t_x + torch.randint(0, 100, (1024,), generator=g))
    work = (t_x * t_y).double()
    order = D.length_bucket_order(t_x, t_y, world=1)
    assert sorted(order.tolist()) == list(range(1024))
    assert bool((work[order][:-1] >= work[order][1:]).all())          # longest first
    for world in (2, 4, 8):
        order = D.length_bucket_order(t_x, t_y, world=world)
        assert sorted(order.tolist()) == list(range(1024))
        loads = []
        for r in range(world):
            lo, hi = D.shard_bounds(1024, r, world)
            shard = work[order[lo:hi]]
            assert bool((shard[:-1] >= shard[1:]).all())              # LPT inside every shard
            loads.append(shard.sum().item())
        assert max(loads) / min(loads) < 1.02                         # same length mix per rank


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T_x = 11
        full = torch.arange(n_total * T_x, dtype=torch.int32).reshape(n_total, T_x)
        lo, hi = D.shard_bounds(n_total, rank, world)
        got = D.all_gather_rows(full[lo:hi].clone(), n_total)
        q.put((rank, bool(torch.equal(got, full)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 9, 1])
def test_all_gather_rows_world2_gloo(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, (rank, shape)
        assert shape == (n_total, 11)


def test_durations_to_logw_matches_reference_formula():
    """tts.py:503-505: logw_ = log(1e-8 + sum_y attn) * x_mask."""
    from art_tts_b200.utils import mas_durations_to_logw
    rng = np.random.default_rng(0)
    dur = torch.from_numpy(rng.integers(0, 5, (3, 17)).astype(np.int32))
    x_mask = (torch.arange(17)[None, :] < torch.tensor([17, 9, 1])[:, None]).float().unsqueeze(1)
    attn = torch.zeros(3, 17, 80)
    for b in range(3):
        c = 0
        for x in range(17):
            attn[b, x, c:c + int(dur[b, x])] = 1
            c += int(dur[b, x])
    want = torch.log(1e-8 + torch.sum(attn.unsqueeze(1), -1)) * x_mask
    assert torch.equal(mas_durations_to_logw(dur, x_mask), want)
