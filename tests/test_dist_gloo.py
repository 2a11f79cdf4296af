"""CPU tests (-m "not gpu") of the N > 1 host logic with the gloo backend, world_size 2 (127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from olavm_b200 import dist as odist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # column shard: disjoint cover of 200 columns
        lo, hi = odist.shard_range(200, rank, world)
        cover = torch.zeros(200, dtype=torch.int64)
        cover[lo:hi] = 1
        dist.all_reduce(cover)
        assert bool((cover == 1).all())
        # coset shard + cap all-gather: every rank ends with the same full cap, entries in cap order
        full = np.arange(16 * 4, dtype=np.int64).reshape(16, 4) * 1000003
        slots = odist.cap_slots(3, 4, rank, world)
        got = odist.allgather_cap(full[slots], 3, 4)
        assert got.shape == (16, 4) and bool((got.numpy() == full).all())
        # timing rule: max over ranks
        assert odist.max_over_ranks(1.0 + rank) == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_partitions():
    for total in (1, 7, 94, 200):
        for world in (1, 2, 4, 8):
            ranges = [odist.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert odist.coset_range(3, 3, 4) == (6, 8)
    assert odist.cap_slots(3, 4, 1, 8) == [2, 3]  # one coset = 2 of the 16 cap subtrees
    # leaf ownership under the coset shard: contiguous leaf ranges
    assert odist.leaf_owner(0, 10, 3, 4) == (0, 0)
    assert odist.leaf_owner((1 << 13) - 1, 10, 3, 4) == (3, (1 << 11) - 1)
    assert odist.leaf_owner(5000, 10, 3, 8) == (4, 5000 - 4096)
