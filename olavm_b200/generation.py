"""Mirror of the trace-generation tail (circuits/src/generation), the step directly in front of prove_with_traces:
`generate_poseidon_trace` (generation/poseidon.rs:5-130, the round states of core/src/util/poseidon_utils.rs included)
on the GPU, `permuted_cols` (stark/lookup.rs:68-131) and `generate_rc_trace` (generation/builtin.rs:249-316) on the GPU,
and the Bitwise / Program compress challenge (generation/builtin.rs:118-131, generation/prog.rs:23-29)."""
import ctypes

import numpy as np

from . import _lib


def generate_poseidon_trace(ctx, inputs, filters=None, log_n=None):
    """inputs [k, 12] permutation inputs, filters [k, 4] (looked_normal, looked_treekey, looked_storage_leaf,
    looked_storage_branch; default all 0) -> the Poseidon table [134, 2^log_n]; rows past k are zero-input padding rows."""
    a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1, 12)
    k = a.shape[0]
    if log_n is None:
        log_n = max(1, (max(k, 1) - 1).bit_length())
    if k > (1 << log_n):
        raise ValueError("more rows than 2^log_n")
    f = None
    if filters is not None:
        f = np.ascontiguousarray(filters, dtype=np.uint64).reshape(-1, 4)
        if f.shape[0] != k:
            raise ValueError("one filter quadruple per input")
    out = np.empty((134, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_poseidon_trace(ctx.handle, _lib.hptr(a) if k else None, _lib.hptr(f), k, log_n, _lib.hptr(out), 0))
    return out


def permuted_cols(ctx, inputs, table):
    """permuted_cols (stark/lookup.rs:68-131): (permuted inputs, permuted table) of one lookup, on the GPU."""
    a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1)
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1)
    if a.shape != t.shape:
        raise ValueError("input and table columns of one length")
    pi, pt = np.empty_like(a), np.empty_like(a)
    ctx.check(ctx._lib.ola_permuted_cols(ctx.handle, _lib.hptr(a), _lib.hptr(t), a.shape[0], _lib.hptr(pi), _lib.hptr(pt), 0))
    return pi, pt


def generate_rc_trace(ctx, vals, kinds, log_n=None):
    """generate_rc_trace (generation/builtin.rs:249-316): values and the table that looks each one up (0 cpu, 1 memory
    sort, 2 memory region, 3 comparison) -> the RangeCheck table [12, 2^log_n] (log_n >= 16)."""
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1)
    k = np.ascontiguousarray(kinds, dtype=np.uint64).reshape(-1)
    if v.shape != k.shape:
        raise ValueError("one kind per value")
    if log_n is None:
        log_n = max(16, (max(v.shape[0], 1) - 1).bit_length())
    out = np.empty((12, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_rangecheck_trace(ctx.handle, _lib.hptr(v) if v.shape[0] else None, _lib.hptr(k) if v.shape[0] else None,
                                                      v.shape[0], log_n, _lib.hptr(out), 0))
    return out


def generate_bitwise_trace(ctx, tags, op0, op1, res, log_n=None):
    """generate_bitwise_trace (generation/builtin.rs:35-206): (the Bitwise table [59, 2^log_n], its compress challenge)."""
    t, a, b, r = (np.ascontiguousarray(x, dtype=np.uint64).reshape(-1) for x in (tags, op0, op1, res))
    k = t.shape[0]
    if not (a.shape[0] == b.shape[0] == r.shape[0] == k):
        raise ValueError("one tag, op0, op1 and res per operation")
    if log_n is None:
        log_n = max(18, (max(k, 1) - 1).bit_length())
    out = np.empty((59, 1 << log_n), dtype=np.uint64)
    beta = ctypes.c_uint64(0)
    ptr = (lambda x: _lib.hptr(x) if k else None)
    ctx.check(ctx._lib.ola_generate_bitwise_trace(ctx.handle, ptr(t), ptr(a), ptr(b), ptr(r), k, log_n, _lib.hptr(out), ctypes.byref(beta), 0))
    return out, int(beta.value)


def generate_cmp_trace(ctx, cells, log_n=None):
    """generate_cmp_trace (generation/builtin.rs:208-247): cells [k, 6] -> the Cmp table [6, 2^log_n]."""
    c = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 6)
    k = c.shape[0]
    if log_n is None:
        log_n = max(1, (max(k, 1) - 1).bit_length())
    out = np.empty((6, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_cmp_trace(ctx.handle, _lib.hptr(c) if k else None, k, log_n, _lib.hptr(out), 0))
    return out


def generate_cpu_trace(ctx, steps, log_n=None):
    """generate_cpu_trace (generation/cpu.rs:11-218): Step records [k, 66] (layout: include/ola_gpu.h) -> the CPU table
    [94, 2^log_n]."""
    r = np.ascontiguousarray(steps, dtype=np.uint64).reshape(-1, 66)
    k = r.shape[0]
    if log_n is None:
        log_n = max(0, (max(k, 1) - 1).bit_length())
    out = np.empty((94, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_cpu_trace(ctx.handle, _lib.hptr(r) if k else None, k, log_n, _lib.hptr(out), 0))
    return out


def generate_memory_trace(ctx, cells, log_n=None):
    """generate_memory_trace (generation/memory.rs:8-155): MemoryTraceCell records [k, 15] (layout: include/ola_gpu.h) -> the
    Memory table [29, 2^log_n]."""
    r = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 15)
    k = r.shape[0]
    if log_n is None:
        log_n = max(1, (max(k, 2) - 1).bit_length())
    out = np.empty((29, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_memory_trace(ctx.handle, _lib.hptr(r) if k else None, k, log_n, _lib.hptr(out), 0))
    return out


def generate_prog_trace(ctx, steps, prog_rows, roots, log_n):
    """generate_prog_trace (generation/prog.rs:18-157): Step records [k, 66], program lines [m, 6], roots[8] = start_root,
    end_root -> (the Program table [18, 2^log_n], its compress challenge)."""
    r = np.ascontiguousarray(steps, dtype=np.uint64).reshape(-1, 66)
    pr = np.ascontiguousarray(prog_rows, dtype=np.uint64).reshape(-1, 6)
    ro = np.ascontiguousarray(roots, dtype=np.uint64).reshape(8)
    out = np.empty((18, 1 << log_n), dtype=np.uint64)
    beta = ctypes.c_uint64(0)
    ctx.check(ctx._lib.ola_generate_program_trace(ctx.handle, _lib.hptr(r) if r.shape[0] else None, r.shape[0], _lib.hptr(pr) if pr.shape[0] else None,
                                                   pr.shape[0], _lib.hptr(ro), log_n, _lib.hptr(out), ctypes.byref(beta), 0))
    return out, int(beta.value)


def compress_challenge(columns):
    """Challenger::new(); observe_elements(column) for every column; get_challenge()."""
    cols = [np.ascontiguousarray(c, dtype=np.uint64).reshape(-1) for c in columns]
    n = cols[0].shape[0] if cols else 0
    if any(c.shape[0] != n for c in cols):
        raise ValueError("columns of one length")
    ptrs = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    beta = ctypes.c_uint64(0)
    rc = _lib.load().ola_compress_challenge(ptrs, len(cols), n, ctypes.byref(beta))
    if rc != 0:
        raise _lib.OlaError(rc, "ola_compress_challenge")
    return int(beta.value)


class Hasher:
    """The two Poseidon services workload generators need (`poseidon(state)`, `poseidon_table_row(input)`), served by the
    library's device entry points."""

    def __init__(self, ctx):
        self.ctx = ctx

    def poseidon(self, state):
        from . import hashing

        return hashing.poseidon(self.ctx, state)

    def poseidon_table_row(self, inp):
        return generate_poseidon_trace(self.ctx, np.asarray(inp, dtype=np.uint64).reshape(1, 12), log_n=1)[:, 0].copy()


def _padded_log(k, log_n):
    return max(1, (max(k, 1) - 1).bit_length()) if log_n is None else log_n


def _small(ctx, fn, rows, rec, ncols, log_n, *extra):
    r = np.ascontiguousarray(rows, dtype=np.uint64).reshape(-1, rec)
    k = r.shape[0]
    log_n = _padded_log(k, log_n)
    out = np.empty((ncols, 1 << log_n), dtype=np.uint64)
    args = extra if extra else (k,)
    ctx.check(fn(ctx.handle, _lib.hptr(r) if k else None, *args, log_n, _lib.hptr(out), 0))
    return out


def generate_poseidon_chunk_trace(ctx, rows, log_n=None):
    """generate_poseidon_chunk_trace (generation/poseidon_chunk.rs:7-88): PoseidonChunkRow records [k, 32] (layout:
    include/ola_gpu.h) -> the PoseidonChunk table [53, 2^log_n]."""
    return _small(ctx, ctx._lib.ola_generate_poseidon_chunk_trace, rows, 32, 53, log_n)


def generate_storage_access_trace(ctx, accesses, prog_hash_reads=(), log_n=None):
    """generate_storage_access_trace (generation/storage.rs:7-123): StorageHashRow records [k, 38] of the storage accesses and
    of the program-hash reads -> the StorageAccess table [48, 2^log_n]."""
    a = np.ascontiguousarray(accesses, dtype=np.uint64).reshape(-1, 38)
    b = np.ascontiguousarray(prog_hash_reads, dtype=np.uint64).reshape(-1, 38)
    return _small(ctx, ctx._lib.ola_generate_storage_access_trace, np.concatenate([a, b]), 38, 48, log_n, a.shape[0], b.shape[0])


def generate_tape_trace(ctx, rows, log_n=None):
    """generate_tape_trace (generation/tape.rs:10-73): TapeRow records [k, 5] -> the Tape table [6, 2^log_n]."""
    return _small(ctx, ctx._lib.ola_generate_tape_trace, rows, 5, 6, log_n)


def generate_sccall_trace(ctx, rows, log_n=None):
    """generate_sccall_trace (generation/sccall.rs:11-64): SCCallRow records [k, 24] -> the SCCall table [26, 2^log_n]."""
    return _small(ctx, ctx._lib.ola_generate_sccall_trace, rows, 24, 26, log_n)


def generate_prog_chunk_trace(ctx, prog_rows, log_n=None):
    """generate_prog_chunk_trace (generation/prog.rs:158-249): prog_rows [m, 6] = (code address 0..3, pc, word) -> the
    ProgChunk table [40, 2^log_n]."""
    r = np.ascontiguousarray(prog_rows, dtype=np.uint64).reshape(-1, 6)
    m = r.shape[0]
    if log_n is None:
        lines = int(((r[:, 4] % np.uint64(8)) == 0).sum()) if m else 0
        log_n = _padded_log(lines, None)
    out = np.empty((40, 1 << log_n), dtype=np.uint64)
    ctx.check(ctx._lib.ola_generate_prog_chunk_trace(ctx.handle, _lib.hptr(r) if m else None, m, log_n, _lib.hptr(out), 0))
    return out
