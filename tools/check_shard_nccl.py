#!/usr/bin/env python3
"""Multi-GPU check of the coset-shard commitment (run under torchrun, one rank per GPU, NCCL):
every rank commits its cosets of the same trace, one all-gather assembles the cap; rank 0 also commits the whole
trace alone and compares the cap, and every rank checks its leaves / Merkle paths against rank 0's full batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import olavm_b200
from olavm_b200 import PolynomialBatch
from olavm_b200 import dist as odist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = olavm_b200.Context(local)
    lg, ncols = 14, 37
    rng = np.random.Generator(np.random.PCG64(5))
    vals = rng.integers(0, 0xFFFFFFFF00000001, size=(ncols, 1 << lg), dtype=np.uint64)
    shard = odist.commit_sharded(ctx, vals, 3, 4, device=dev)
    full = PolynomialBatch.from_values(ctx, vals, 3, False, 4)  # every rank: reference for its own checks
    assert (shard.merkle_cap.hashes == full.merkle_cap.hashes).all(), "cap mismatch"
    L = 1 << (lg + 3)
    per = L // world
    assert (shard.leaves() == full.leaves(rank * per, per)).all(), "leaf range mismatch"
    for g in (rank * per, rank * per + 12345 % per, (rank + 1) * per - 1):
        owner, loc = odist.leaf_owner(g, lg, 3, world)
        assert owner == rank
        assert (shard.prove(loc) == full.prove(g)).all(), "path mismatch"
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f"coset-shard commit over {world} GPUs: cap, leaves and paths equal the single-GPU commitment ({int(ok.item())} ranks ok)")
    shard.free()
    full.free()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
