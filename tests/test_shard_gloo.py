"""world_size-2 (and 3) gloo tests of the time-sharding exchanges on CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, halo, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from poor_man_gplvm_b200.shard import TimeShard
        sh = TimeShard()
        assert sh.active and sh.rank == rank and sh.world == world
        full = torch.arange(T * 3, dtype=torch.float32).reshape(T, 3)
        per = T // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else T
        block = full[lo:hi].clone()
        ext, hl, hr = sh.halo_exchange(block, halo)
        want = full[max(0, lo - halo):min(T, hi + halo)]
        assert hl == (halo if rank > 0 else 0) and hr == (halo if rank < world - 1 else 0)
        assert torch.equal(ext, want)
        # boundary messages in both directions, then leftwards only (the backward-pass pattern)
        first, last = block[0].clone(), block[-1].clone()
        from_left, from_right = sh.boundary(first, last)
        assert (from_left is None) == (rank == 0) and (from_right is None) == (rank == world - 1)
        if from_left is not None:
            assert torch.equal(from_left, full[lo - 1])
        if from_right is not None:
            assert torch.equal(from_right, full[hi])
        from_left, from_right = sh.boundary(first, None)
        assert from_left is None
        if rank < world - 1:
            assert torch.equal(from_right, full[hi])
        # packed all-reduce of statistics-like tensors and the scalar "any seam failing" reduction
        a = torch.full((4, 5), float(rank + 1))
        b = torch.full((7,), float(10 * (rank + 1)))
        sh.allreduce_sum_(a, b)
        tot = sum(range(1, world + 1))
        assert torch.all(a == tot) and torch.all(b == 10 * tot)
        assert sh.max_int(rank, torch.device("cpu")) == world - 1
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_time_shard_exchanges(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 101, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_single_process_shard_is_inert():
    from poor_man_gplvm_b200.shard import TimeShard
    sh = TimeShard(None, single=True)
    x = torch.arange(12.0).reshape(6, 2)
    ext, hl, hr = sh.halo_exchange(x, 3)
    assert ext is x and hl == 0 and hr == 0
    assert sh.boundary(x[0], x[-1]) == (None, None)
    assert sh.max_int(5, torch.device("cpu")) == 5 and sh.is_first and sh.is_last
