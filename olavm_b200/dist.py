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


# ---------------------------------------------------------------------------------------------------------------------
# Communicators for the coset-sharded prover (ola_set_comm): the library takes two collectives as C callbacks.
# ---------------------------------------------------------------------------------------------------------------------
import ctypes  # noqa: E402

_ALLGATHER = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
_ALLREDUCE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


class _DevView:
    """Zero-copy view of `count` int64 at a raw device pointer (CUDA array interface) for torch.as_tensor."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i8", "data": (int(ptr), False), "version": 3}


def set_comm_torch(ctx, group=None):
    """Bind the library's collectives to torch.distributed (NCCL over NVLink on a GPU box): all-gather and a wrapping
    u64 sum, both enqueued on the context's stream.  Call once per rank after init_process_group; keeps the callbacks
    alive on the context."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

    def allgather(_user, send, recv, nbytes, _stream):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(stream):
                s = torch.as_tensor(_DevView(send, nbytes // 8), device=dev)
                r = torch.as_tensor(_DevView(recv, world * (nbytes // 8)), device=dev)
                dist.all_gather_into_tensor(r, s, group=group)
            return 0
        except Exception as e:  # pragma: no cover
            print("olavm_b200.dist allgather failed:", repr(e))
            return 1

    def allreduce(_user, buf, count, _stream):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(stream):
                t = torch.as_tensor(_DevView(buf, count), device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)  # int64 wrap-around == u64 wrap-around
            return 0
        except Exception as e:  # pragma: no cover
            print("olavm_b200.dist allreduce failed:", repr(e))
            return 1

    ctx._comm_keep = (_ALLGATHER(allgather), _ALLREDUCE(allreduce))
    ctx.check(ctx._lib.ola_set_comm(ctx.handle, rank, world, ctypes.cast(ctx._comm_keep[0], ctypes.c_void_p),
                                    ctypes.cast(ctx._comm_keep[1], ctypes.c_void_p), None))
    return rank, world


def set_comm(ctx, group=None, libnccl_path=None):
    """Bind the library's collectives to ITS OWN NCCL communicator (ola_set_comm_nccl: libnccl is dlopened by the library,
    the collectives are issued from C++ on the context's stream).  torch.distributed only carries the 128-byte
    ncclUniqueId from rank 0 to the others.  OLA_COMM=torch selects the callback binding (set_comm_torch) instead."""
    import os

    import numpy as np

    if os.environ.get("OLA_COMM", "nccl") == "torch":
        return set_comm_torch(ctx, group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = np.zeros(128, dtype=np.uint8)
    path = libnccl_path.encode() if libnccl_path else None
    if rank == 0:
        ctx.check(ctx._lib.ola_nccl_unique_id(path, uid.ctypes.data_as(ctypes.c_void_p)))
    backend = dist.get_backend(group)
    dev = torch.device("cuda", ctx.device) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(uid.copy()).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    uid = t.cpu().numpy()
    ctx.check(ctx._lib.ola_set_comm_nccl(ctx.handle, path, rank, world, uid.ctypes.data_as(ctypes.c_void_p)))
    return rank, world


def comm_bytes(ctx):
    """Bytes this rank received through the library's communicator so far (ola_comm_bytes)."""
    return int(ctx._lib.ola_comm_bytes(ctx.handle))


class LocalComm:
    """In-process communicator for `world` host threads that each own a Context on the SAME GPU (test harness for the
    sharded prover on a one-GPU box): collectives go through host staging buffers and a threading.Barrier.
    Use: comm = LocalComm(world); in thread r: comm.attach(ctx_r, r)."""

    def __init__(self, world):
        import threading

        self.world = world
        self.barrier = threading.Barrier(world)
        self.stage = [None] * world

    def attach(self, ctx, rank):
        import numpy as np

        def exchange(ptr, count):
            ctx.sync()
            self.stage[rank] = ctx.download(ctypes.c_void_p(ptr), (count,))
            self.barrier.wait()
            parts = [self.stage[r] for r in range(self.world)]
            self.barrier.wait()
            return parts

        def allgather(_user, send, recv, nbytes, _stream):
            try:
                parts = exchange(send, nbytes // 8)
                ctx.upload(np.concatenate(parts), ctypes.c_void_p(recv))
                return 0
            except Exception as e:  # pragma: no cover
                print("LocalComm allgather failed:", repr(e))
                return 1

        def allreduce(_user, buf, count, _stream):
            try:
                parts = exchange(buf, count)
                total = parts[0].copy()
                for p in parts[1:]:
                    total += p  # uint64 wrap-around
                ctx.upload(total, ctypes.c_void_p(buf))
                return 0
            except Exception as e:  # pragma: no cover
                print("LocalComm allreduce failed:", repr(e))
                return 1

        ctx._comm_keep = (_ALLGATHER(allgather), _ALLREDUCE(allreduce))
        ctx.check(ctx._lib.ola_set_comm(ctx.handle, rank, self.world, ctypes.cast(ctx._comm_keep[0], ctypes.c_void_p),
                                        ctypes.cast(ctx._comm_keep[1], ctypes.c_void_p), None))


def prove_sharded_local(device, world, table_ids, traces, **kw):
    """Run the coset-sharded prover with `world` ranks as host threads on one GPU; returns every rank's proof bytes."""
    import threading

    from .context import Context
    from .prover import prove_with_traces

    comm = LocalComm(world)
    out, err = [None] * world, [None] * world
    hasher = kw.pop("hasher", 0)

    def run(rank):
        try:
            ctx = Context(device)
            ctx.hasher = hasher
            comm.attach(ctx, rank)
            out[rank] = prove_with_traces(ctx, table_ids, traces, **kw)
            ctx.close()
        except BaseException as e:  # noqa: B902
            err[rank] = e
            comm.barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in err if e is not None and "callback failed" not in str(e)]
    for e in real + [e for e in err if e is not None]:
        raise e  # the root cause first; ranks that merely lost their peers (broken barrier) last
    return out
