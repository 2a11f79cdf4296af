"""Multi-GPU plumbing (torch.distributed; NCCL on the GPU box, gloo in CPU tests).

The proving path shards without a data-path collective at the granularity used so far:
  * column shard  -- NTT / LDE / leaf-column hashing inputs are independent per column: rank r owns a contiguous
    column range (`shard_range`);
  * coset shard   -- with blowup 2^rate_bits, rank r owns LDE cosets (= contiguous leaf ranges = cap subtrees):
    `coset_range` gives the cosets and `cap_slots` the Merkle-cap entries a rank produces; the only exchange is an
    all-gather of those cap entries (`allgather_cap`).
Timing of any multi-rank measurement is the max over ranks of device time (`max_over_ranks`).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block partition of `total` items: sizes differ by at most one, earlier ranks get the extras."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def coset_range(rate_bits, rank, world):
    """LDE cosets (leaf blocks of n rows) owned by `rank`; world must divide 2^rate_bits."""
    ncosets = 1 << rate_bits
    if ncosets % world:
        raise ValueError("world size must divide the blowup factor")
    per = ncosets // world
    return rank * per, (rank + 1) * per


def cap_slots(rate_bits, cap_height, rank, world):
    """Indices of the Merkle-cap entries whose subtrees lie entirely inside this rank's coset range."""
    lo, hi = coset_range(rate_bits, rank, world)
    ncap, ncosets = 1 << cap_height, 1 << rate_bits
    if ncap % ncosets:
        raise ValueError("cap must be at least as fine as the coset partition")
    per = ncap // ncosets
    return list(range(lo * per, hi * per))


def allgather_cap(local_cap, rate_bits, cap_height, device=None):
    """All-gather the cap entries each rank produced into the full cap [2^cap_height, 4] (int64 view of u64)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    slots = cap_slots(rate_bits, cap_height, rank, world)
    local = torch.as_tensor(local_cap, dtype=torch.int64, device=device).reshape(len(slots), 4).contiguous()
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local)
    return torch.cat(out, dim=0)


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def commit_sharded(ctx, values, rate_bits, cap_height, is_coeffs=False, device=None, **kw):
    """PolynomialBatch::from_values across the ranks of the default process group (coset shard): every
    rank passes the SAME trace columns, evaluates / hashes / reduces only its cosets, and one all-gather of
    2^cap_height / world digests per rank assembles the cap.  Returns the local shard with the full `merkle_cap`."""
    from .pcs import MerkleCap, PolynomialBatch

    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = coset_range(rate_bits, rank, world)
    b = PolynomialBatch._commit(ctx, values, is_coeffs, rate_bits, cap_height, coset_first=lo, coset_count=hi - lo, **kw)
    full = allgather_cap(b.merkle_cap.hashes.view("int64"), rate_bits, cap_height, device=device)
    b.merkle_cap = MerkleCap(full.cpu().numpy().view("uint64"))
    return b


def leaf_owner(leaf_index, degree_log, rate_bits, world):
    """(rank, local leaf index) of a global leaf under the coset shard."""
    per = (1 << (degree_log + rate_bits)) // world
    return leaf_index // per, leaf_index % per
