"""CPU, world_size 2, gloo: the host-side sharding / gather / max-over-ranks logic that bench.py and
multi-GPU drivers use (the N > 1 path has no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG  # noqa: F401  (puts the package on sys.path)
from foto_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, row_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.shard_indices(n_items, rank, world)
    rows = np.stack([np.full(row_len, 10.0 * i) + np.arange(row_len) for i in idx]) if idx else np.zeros((0, row_len))
    shard.barrier()
    mx = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
    full = shard.gather_rows(rows, idx, n_items, row_len)
    q.put((rank, idx, mx, None if full is None else full.tolist()))
    dist.destroy_process_group()


def test_sharding_gather_and_max_with_gloo():
    world, n_items, row_len = 2, 5, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, row_len, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, idx0, mx0, full0), (r1, idx1, mx1, full1) = res
    assert idx0 == [0, 2, 4] and idx1 == [1, 3]
    assert mx0 == mx1 == [2.0, 5.0]
    assert full1 is None
    expect = np.stack([np.full(row_len, 10.0 * i) + np.arange(row_len) for i in range(n_items)])
    np.testing.assert_array_equal(np.array(full0), expect)


def _queue_worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wq = shard.WorkQueue("t1", n_items, chunk=2 if rank else 1)
    shard.barrier()
    got = []
    while True:
        i = wq.next()
        if i is None:
            break
        got.append(i)
    stats = shard.gather_objects({"rank": rank, "n": len(got)})
    q.put((rank, got, stats))
    shard.barrier()
    dist.destroy_process_group()


def test_work_queue_hands_every_item_out_once_with_gloo():
    """The shared queue bench.py and multi-process drivers draw pairs from: every index exactly once over the ranks."""
    world, n_items = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_queue_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    items = sorted(res[0][1] + res[1][1])
    assert items == list(range(n_items))
    assert res[0][2] == res[1][2] and sum(s["n"] for s in res[0][2]) == n_items


def test_single_process_identity():
    wq = shard.WorkQueue("solo", 3)
    assert [wq.next(), wq.next(), wq.next(), wq.next()] == [0, 1, 2, None]
    assert shard.gather_objects(5) == [5]
    assert shard.shard_indices(7, 0, 1) == list(range(7))
    assert shard.max_over_ranks([3.0]) == [3.0]
    out = shard.gather_rows(np.ones((2, 4)), [0, 1], 2, 4)
    assert out.shape == (2, 4)


def test_slab_plan_partitions():
    from foto_b200 import slab
    g = slab.plan(64, 2160, 8)
    assert g["t"][0] == (0, 8) and g["t"][-1] == (56, 64) and sum(b - a for a, b in g["t"]) == 64
    assert sum(b - a for a, b in g["y"]) == 2160 and all(b > a for a, b in g["y"])
    g = slab.plan(5, 7, 3)                                  # uneven
    assert [b - a for a, b in g["t"]] == [1, 2, 2] and [b - a for a, b in g["y"]] == [2, 2, 3]
    import pytest
    with pytest.raises(ValueError):
        slab.plan(2, 100, 4)
