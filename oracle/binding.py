"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/libola_oracle.so.

Every function cites, in oracle/*.c, the reference file:line it restates.  numpy uint64 arrays in,
numpy uint64 arrays out; all values canonical Goldilocks representatives.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libola_oracle.so")
P = 0xFFFFFFFF00000001


def build(force=False):
    """Compile the C restatement (gcc + OpenMP).  Called by __graft_entry__.build()."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".cpp", ".hpp"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _set_sigs(_lib)
    return _lib


_u64p = ctypes.POINTER(ctypes.c_uint64)
_sz = ctypes.c_size_t
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_int = ctypes.c_int


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _set_sigs(L):
    L.orc_get_twiddles.argtypes = [_u64p, _sz, _int]
    L.orc_evaluate_poly.argtypes = [_u64p, _sz]
    L.orc_interpolate_poly.argtypes = [_u64p, _sz]
    L.orc_evaluate_poly_with_offset.argtypes = [_u64p, _sz, _u64, _sz, _u64p]
    L.orc_interpolate_poly_with_offset.argtypes = [_u64p, _sz, _u64]
    L.orc_fft_classic.argtypes = [_u64p, _sz]
    L.orc_poly_eval.argtypes = [_u64p, _sz, _u64]
    L.orc_poly_eval.restype = _u64
    L.orc_ifft_batch.argtypes = [_u64p, _sz, _sz]
    L.orc_lde_batch.argtypes = [_u64p, _sz, _sz, _u64, _sz, _u64p]
    L.orc_poseidon.argtypes = [_u64p]
    L.orc_poseidon_naive.argtypes = [_u64p]
    L.orc_poseidon_table_row.argtypes = [_u64p, _u64p]
    L.orc_poseidon_table_row.restype = None
    L.orc_hash_no_pad.argtypes = [_u64p, _sz, _u64p]
    L.orc_two_to_one.argtypes = [_u64p, _u64p, _u64p]
    L.orc_hash_rows.argtypes = [_u64p, _sz, _sz, _u64p]
    L.orc_set_hasher.argtypes = [_int]
    L.orc_get_hasher.restype = _int
    L.orc_challenger_permute.argtypes = [_u64p]
    L.orc_blake3.argtypes = [ctypes.c_char_p, _sz, ctypes.POINTER(ctypes.c_uint8)]
    L.orc_blake3_hash_no_pad.argtypes = [_u64p, _sz, _u64p]
    L.orc_blake3_two_to_one.argtypes = [_u64p, _u64p, _u64p]
    L.orc_blake3_permute.argtypes = [_u64p]
    L.orc_bytes_hash_to_fields.argtypes = [_u64p, _u64p]
    L.orc_build_merkle_nodes.argtypes = [_u64p, _sz, _u64p]
    L.orc_merkle_new_v2.argtypes = [_u64p, _sz, _sz, _u32, _u64p, _u64p]
    L.orc_merkle_new_v2.restype = _int
    L.orc_merkle_prove.argtypes = [_u64p, _sz, _u32, _sz, _u64p]
    L.orc_merkle_prove.restype = _int
    L.orc_merkle_verify.argtypes = [_u64p, _sz, _sz, _u64p, _u64p, _sz]
    L.orc_merkle_verify.restype = _int
    L.orc_commit.argtypes = [_u64p, _sz, _sz, _int, _u32, _u32, _u64p, _u64p, _u64p, _u64p]
    L.orc_commit.restype = _int
    L.orc_stark_prove.argtypes = [ctypes.POINTER(_int), _u32, ctypes.POINTER(_u64p), ctypes.POINTER(_u32), _u64p, _int,
                                  ctypes.POINTER(ctypes.c_uint8), _sz, ctypes.POINTER(_sz), ctypes.c_char_p, _sz]
    L.orc_stark_prove.restype = _int
    L.orc_stark_verify.argtypes = [ctypes.POINTER(_int), _u32, ctypes.POINTER(ctypes.c_uint8), _sz, ctypes.c_char_p, _sz]
    L.orc_stark_verify.restype = _int
    L.orc_table_columns.argtypes = [_int]
    L.orc_air_first_failure.argtypes = [_int, _u64p, _u32, _u64, ctypes.POINTER(_u64), ctypes.POINTER(_int)]
    L.orc_table_columns.restype = _int
    L.orc_air_constraints.argtypes = [_int, _u64p, _u64p, _u64, _u64p, ctypes.POINTER(_int), _int]
    L.orc_air_constraints.restype = _int
    L.orc_permuted_cols.argtypes = [_u64p, _u64p, _sz, _u64p, _u64p]
    L.orc_permuted_cols.restype = None
    L.orc_generate_rc_trace.argtypes = [_u64p, ctypes.POINTER(ctypes.c_uint8), _sz, _u64p, _sz]
    L.orc_generate_rc_trace.restype = _sz
    L.orc_generate_bitwise_trace.argtypes = [_u64p, _u64p, _u64p, _u64p, _sz, _u64p, _sz, _u64p]
    L.orc_generate_bitwise_trace.restype = _sz
    L.orc_generate_cmp_trace.argtypes = [_u64p, _sz, _u64p, _sz]
    L.orc_generate_cmp_trace.restype = _sz
    L.orc_generate_cpu_trace.argtypes = [_u64p, _sz, _sz, _u64p]
    L.orc_generate_cpu_trace.restype = None
    L.orc_generate_memory_trace.argtypes = [_u64p, _sz, _sz, _u64p]
    L.orc_generate_memory_trace.restype = None
    L.orc_generate_prog_trace.argtypes = [_u64p, _sz, _u64p, _sz, _u64p, _u64p, _sz, _u64p]
    L.orc_generate_prog_trace.restype = _sz
    for name, nsz in (("poseidon_chunk", 1), ("storage_access", 2), ("tape", 1), ("sccall", 1), ("prog_chunk", 1)):
        f = getattr(L, "orc_generate_%s_trace" % name)
        f.argtypes = [_u64p] + [_sz] * nsz + [_u64p, _sz]
        f.restype = _sz
    L.orc_compress_challenge.argtypes = [ctypes.POINTER(ctypes.c_void_p), _u32, _sz]
    L.orc_compress_challenge.restype = _u64


# ---------------------------------------------------------------- helpers
def splitmix64(seed, n):
    """Deterministic canonical Goldilocks elements (splitmix64 stream mod p)."""
    out = np.empty(n, dtype=np.uint64)
    x = np.uint64(seed)
    M = (1 << 64) - 1
    s = int(seed) & M
    for i in range(n):
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z = z ^ (z >> 31)
        out[i] = z % P
    return out


def rand_elems(seed, shape):
    """Fast vectorised uniform canonical elements (numpy PCG64), for larger test inputs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, P, size=shape, dtype=np.uint64, endpoint=False)
    return np.ascontiguousarray(a)


# ---------------------------------------------------------------- NTT
def evaluate_poly(coeffs):
    v = np.ascontiguousarray(coeffs, dtype=np.uint64).copy()
    lib().orc_evaluate_poly(_p(v), v.size)
    return v


def interpolate_poly(values):
    v = np.ascontiguousarray(values, dtype=np.uint64).copy()
    lib().orc_interpolate_poly(_p(v), v.size)
    return v


def evaluate_poly_with_offset(coeffs, shift, blowup):
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.empty(c.size * blowup, dtype=np.uint64)
    lib().orc_evaluate_poly_with_offset(_p(c), c.size, int(shift), blowup, _p(out))
    return out


def interpolate_poly_with_offset(values, shift):
    v = np.ascontiguousarray(values, dtype=np.uint64).copy()
    lib().orc_interpolate_poly_with_offset(_p(v), v.size, int(shift))
    return v


def fft_classic(coeffs):
    v = np.ascontiguousarray(coeffs, dtype=np.uint64).copy()
    lib().orc_fft_classic(_p(v), v.size)
    return v


def poly_eval(coeffs, x):
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    return int(lib().orc_poly_eval(_p(c), c.size, int(x)))


def ifft_batch(cols):
    v = np.ascontiguousarray(cols, dtype=np.uint64).copy()
    lib().orc_ifft_batch(_p(v), v.shape[0], v.shape[1])
    return v


def lde_batch(coeffs, shift=7, blowup=8):
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.empty((c.shape[0], c.shape[1] * blowup), dtype=np.uint64)
    lib().orc_lde_batch(_p(c), c.shape[0], c.shape[1], int(shift), blowup, _p(out))
    return out


# ---------------------------------------------------------------- Hasher selection (GenericConfig::Hasher)
POSEIDON, BLAKE3 = 0, 1


class hasher:
    """with oracle.hasher(oracle.BLAKE3): ...  -- C::Hasher = Blake3_256<32> (Blake3GoldilocksConfig, plonk/config.rs:153-161)
    for leaf / node hashing, the challenger's permutation and the hash wire format; Poseidon (the default) otherwise."""

    def __init__(self, hid):
        self.hid = int(hid)

    def __enter__(self):
        self.prev = lib().orc_get_hasher()
        lib().orc_set_hasher(self.hid)
        return self

    def __exit__(self, *exc):
        lib().orc_set_hasher(self.prev)
        return False


def blake3(data):
    out = (ctypes.c_uint8 * 32)()
    data = bytes(data)
    lib().orc_blake3(data, len(data), out)
    return bytes(out)


def blake3_permute(state):
    s = np.ascontiguousarray(state, dtype=np.uint64).copy()
    assert s.size == 12
    lib().orc_blake3_permute(_p(s))
    return s


def bytes_hash_to_fields(h):
    a = np.ascontiguousarray(h, dtype=np.uint64)
    out = np.empty(5, dtype=np.uint64)
    lib().orc_bytes_hash_to_fields(_p(a), _p(out))
    return out


# ---------------------------------------------------------------- Poseidon / Merkle
def poseidon(state, naive=False):
    s = np.ascontiguousarray(state, dtype=np.uint64).copy()
    assert s.size == 12
    (lib().orc_poseidon_naive if naive else lib().orc_poseidon)(_p(s))
    return s


def hash_no_pad(inp):
    a = np.ascontiguousarray(inp, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_p(a), a.size, _p(out))
    return out


def two_to_one(l, r):
    l = np.ascontiguousarray(l, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_two_to_one(_p(l), _p(r), _p(out))
    return out


def hash_rows(rows):
    a = np.ascontiguousarray(rows, dtype=np.uint64)
    out = np.empty((a.shape[0], 4), dtype=np.uint64)
    lib().orc_hash_rows(_p(a), a.shape[0], a.shape[1], _p(out))
    return out


def merkle_new_v2(rows, cap_height):
    a = np.ascontiguousarray(rows, dtype=np.uint64)
    n = a.shape[0]
    ncap = 1 << cap_height
    dig = np.empty((max(2 * (n - ncap), 0), 4), dtype=np.uint64)
    cap = np.empty((ncap, 4), dtype=np.uint64)
    rc = lib().orc_merkle_new_v2(_p(a), n, a.shape[1], cap_height, _p(dig) if dig.size else _p(np.zeros(4, np.uint64)), _p(cap))
    assert rc == 0
    return dig, cap


def merkle_prove(digests, nrows, cap_height, index):
    nl = int(np.log2(nrows)) - cap_height
    sib = np.empty((max(nl, 1), 4), dtype=np.uint64)
    k = lib().orc_merkle_prove(_p(digests), nrows, cap_height, index, _p(sib))
    return sib[:k]


def merkle_verify(leaf, index, cap, siblings):
    leaf = np.ascontiguousarray(leaf, dtype=np.uint64)
    sib = np.ascontiguousarray(siblings, dtype=np.uint64).reshape(-1, 4)
    sp = _p(sib) if sib.size else _p(np.zeros(4, np.uint64))
    return bool(lib().orc_merkle_verify(_p(leaf), leaf.size, index, _p(np.ascontiguousarray(cap)), sp, sib.shape[0]))


def commit(cols, is_coeffs=False, rate_bits=3, cap_height=4, want_leaves=True, want_digests=True):
    """PolynomialBatch::from_values / from_coeffs.  Returns dict(coeffs, leaves, digests, cap)."""
    c = np.ascontiguousarray(cols, dtype=np.uint64)
    ncols, n = c.shape
    L = n << rate_bits
    ncap = 1 << cap_height
    coeffs = np.empty((ncols, n), dtype=np.uint64)
    leaves = np.empty((L, ncols), dtype=np.uint64) if want_leaves else None
    dig = np.empty((max(2 * (L - ncap), 1), 4), dtype=np.uint64) if want_digests else None
    cap = np.empty((ncap, 4), dtype=np.uint64)
    rc = lib().orc_commit(_p(c), ncols, n, int(is_coeffs), rate_bits, cap_height, _p(coeffs), _p(leaves), _p(dig), _p(cap))
    assert rc == 0
    return dict(coeffs=coeffs, leaves=leaves, digests=dig, cap=cap)


# ---------------------------------------------------------------- STARK prover / verifier
TABLE_IDS = dict(cpu=0, memory=1, bitwise=2, cmp=3, rangecheck=4, poseidon=5, poseidon_chunk=6, storage=7, tape=8, sccall=9,
                 program=10, prog_chunk=11)


class StarkError(RuntimeError):
    pass


def table_columns(table_id):
    return lib().orc_table_columns(int(table_id))


def air_first_failure(table_id, trace, compress_challenge=0):
    """The reference's per-table acceptance test (cpu_stark.rs:974-1105 and siblings): evaluate the table's AIR on every
    row pair of `trace` ([columns, 2^k]).  None when all constraints vanish, else (row, position of the first non-zero
    constraint in evaluation order)."""
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    row, idx = _u64(0), _int(0)
    rc = lib().orc_air_first_failure(int(table_id), _p(t), int(t.shape[1]).bit_length() - 1, int(compress_challenge), ctypes.byref(row),
                                     ctypes.byref(idx))
    if rc < 0:
        raise StarkError("air_first_failure: unknown table or bad trace")
    return None if rc == 0 else (int(row.value), int(idx.value))


def air_constraints(table_id, lv, nv, compress_challenge=0):
    """The individual constraint values the table's eval_packed_generic emits for one (local, next) row pair, in order,
    with their kinds (0 constraint, 1 transition, 2 first row, 3 last row): (values uint64[K], kinds int32[K])."""
    lv = np.ascontiguousarray(lv, dtype=np.uint64)
    nv = np.ascontiguousarray(nv, dtype=np.uint64)
    cap = 4096
    vals = np.zeros(cap, dtype=np.uint64)
    kinds = np.zeros(cap, dtype=np.int32)
    n = lib().orc_air_constraints(int(table_id), _p(lv), _p(nv), int(compress_challenge), _p(vals), kinds.ctypes.data_as(ctypes.POINTER(_int)), cap)
    if n < 0 or n > cap:
        raise StarkError("air_constraints: unknown table")
    return vals[:n].copy(), kinds[:n].copy()


def permuted_cols(inputs, table):
    """lookup.rs:68-131: (permuted inputs, permuted table) of the Halo2-style lookup argument."""
    a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1)
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1)
    assert a.shape == t.shape
    pi, pt = np.empty_like(a), np.empty_like(a)
    lib().orc_permuted_cols(_p(a), _p(t), a.shape[0], _p(pi), _p(pt))
    return pi, pt


def generate_rc_trace(vals, kinds):
    """generate_rc_trace (builtin.rs:249-316): the 12-column RangeCheck table of the given (value, looking table) rows."""
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1)
    k = np.ascontiguousarray(kinds, dtype=np.uint8).reshape(-1)
    assert v.shape == k.shape
    kp = k.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    n = int(lib().orc_generate_rc_trace(_p(v), kp, v.shape[0], None, 0))
    out = np.empty((12, n), dtype=np.uint64)
    lib().orc_generate_rc_trace(_p(v), kp, v.shape[0], _p(out), n)
    return out


def generate_bitwise_trace(tags, op0, op1, res):
    """generate_bitwise_trace (builtin.rs:35-206): (the 59-column Bitwise table, its compress challenge beta)."""
    t, a, b, r = (np.ascontiguousarray(x, dtype=np.uint64).reshape(-1) for x in (tags, op0, op1, res))
    k = t.shape[0]
    n = int(lib().orc_generate_bitwise_trace(_p(t), _p(a), _p(b), _p(r), k, None, 0, None))
    out = np.empty((59, n), dtype=np.uint64)
    beta = np.zeros(1, dtype=np.uint64)
    lib().orc_generate_bitwise_trace(_p(t), _p(a), _p(b), _p(r), k, _p(out), n, _p(beta))
    return out, int(beta[0])


def generate_cmp_trace(cells):
    """generate_cmp_trace (builtin.rs:208-247): cells [k, 6] -> the Cmp table [6, n]."""
    c = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 6)
    n = int(lib().orc_generate_cmp_trace(_p(c), c.shape[0], None, 0))
    out = np.empty((6, n), dtype=np.uint64)
    lib().orc_generate_cmp_trace(_p(c), c.shape[0], _p(out), n)
    return out


def generate_cpu_trace(steps, log_n=None):
    """generate_cpu_trace (generation/cpu.rs:11-218): step records [k, 66] (layout: oracle/generation_cpu.c) -> the CPU table [94, n]."""
    r = np.ascontiguousarray(steps, dtype=np.uint64).reshape(-1, 66)
    k = r.shape[0]
    n = 1 << (log_n if log_n is not None else max(0, (max(k, 1) - 1).bit_length()))
    assert n >= k
    out = np.empty((94, n), dtype=np.uint64)
    lib().orc_generate_cpu_trace(_p(r), k, n, _p(out))
    return out


def generate_memory_trace(cells, log_n=None):
    """generate_memory_trace (generation/memory.rs:8-155): MemoryTraceCell records [k, 15] (layout: oracle/generation_cpu.c)
    -> the Memory table [29, n]."""
    r = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 15)
    k = r.shape[0]
    n = 1 << (log_n if log_n is not None else max(1, (max(k, 2) - 1).bit_length()))
    assert n >= max(k, 2)
    out = np.empty((29, n), dtype=np.uint64)
    lib().orc_generate_memory_trace(_p(r), k, n, _p(out))
    return out


def generate_prog_trace(steps, prog_rows, roots, log_n=None):
    """generate_prog_trace (generation/prog.rs:18-157): Step records [k, 66], program lines [m, 6] = (addr0..3, pc, inst),
    roots[8] = start_root, end_root -> (the Program table [18, n], its compress challenge beta)."""
    r = np.ascontiguousarray(steps, dtype=np.uint64).reshape(-1, 66)
    pr = np.ascontiguousarray(prog_rows, dtype=np.uint64).reshape(-1, 6)
    ro = np.ascontiguousarray(roots, dtype=np.uint64).reshape(8)
    n = int(lib().orc_generate_prog_trace(_p(r), r.shape[0], _p(pr), pr.shape[0], _p(ro), None, 0, None))
    if log_n is not None:
        assert (1 << log_n) >= n
        n = 1 << log_n
    out = np.empty((18, n), dtype=np.uint64)
    beta = np.zeros(1, dtype=np.uint64)
    lib().orc_generate_prog_trace(_p(r), r.shape[0], _p(pr), pr.shape[0], _p(ro), _p(out), n, _p(beta))
    return out, int(beta[0])


def _small_table(fn, ncols, args, log_n):
    n = int(fn(*args, None, 0))
    if log_n is not None:
        assert (1 << log_n) >= n
        n = 1 << log_n
    out = np.empty((ncols, n), dtype=np.uint64)
    fn(*args, _p(out), n)
    return out


def generate_poseidon_chunk_trace(cells, log_n=None):
    """generate_poseidon_chunk_trace (generation/poseidon_chunk.rs:7-88): PoseidonChunkRow records [k, 32] -> [53, n]."""
    c = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 32)
    return _small_table(lib().orc_generate_poseidon_chunk_trace, 53, (_p(c), c.shape[0]), log_n)


def generate_storage_access_trace(accesses, prog_hash_reads=(), log_n=None):
    """generate_storage_access_trace (generation/storage.rs:7-123): StorageHashRow records [k, 38] -> [48, n]."""
    a = np.ascontiguousarray(accesses, dtype=np.uint64).reshape(-1, 38)
    b = np.ascontiguousarray(prog_hash_reads, dtype=np.uint64).reshape(-1, 38)
    r = np.ascontiguousarray(np.concatenate([a, b]))
    return _small_table(lib().orc_generate_storage_access_trace, 48, (_p(r), a.shape[0], b.shape[0]), log_n)


def generate_tape_trace(cells, log_n=None):
    """generate_tape_trace (generation/tape.rs:10-73): TapeRow records [k, 5] -> [6, n]."""
    c = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 5)
    return _small_table(lib().orc_generate_tape_trace, 6, (_p(c), c.shape[0]), log_n)


def generate_sccall_trace(cells, log_n=None):
    """generate_sccall_trace (generation/sccall.rs:11-64): SCCallRow records [k, 24] -> [26, n]."""
    c = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1, 24)
    return _small_table(lib().orc_generate_sccall_trace, 26, (_p(c), c.shape[0]), log_n)


def generate_prog_chunk_trace(prog_rows, log_n=None):
    """generate_prog_chunk_trace (generation/prog.rs:158-249): program words [m, 6] = (addr0..3, pc, word) -> [40, n]."""
    r = np.ascontiguousarray(prog_rows, dtype=np.uint64).reshape(-1, 6)
    return _small_table(lib().orc_generate_prog_chunk_trace, 40, (_p(r), r.shape[0]), log_n)


def compress_challenge(columns):
    """The Bitwise / Program compress challenge (generation/builtin.rs:118-131, generation/prog.rs:23-29): a fresh
    Challenger observes every column in turn and squeezes one element."""
    cols = [np.ascontiguousarray(c, dtype=np.uint64).reshape(-1) for c in columns]
    n = cols[0].shape[0] if cols else 0
    ptrs = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    return int(lib().orc_compress_challenge(ptrs, len(cols), n))


def poseidon_table_row(inp):
    """Witness row of the Poseidon table (134 columns; the 4 filters left 0) for a 12-element permutation input."""
    a = np.ascontiguousarray(inp, dtype=np.uint64)
    assert a.shape == (12,)
    row = np.zeros(134, dtype=np.uint64)
    lib().orc_poseidon_table_row(_p(a), _p(row))
    return row


def stark_prove(table_ids, traces, check_degree=True, max_bytes=1 << 26, compress_challenges=None, hasher_id=None):
    """prove_with_traces + Buffer::write_all_proof -> bytes.  traces[i]: [columns_i, 2^k_i] uint64 column-major."""
    if hasher_id is not None:
        with hasher(hasher_id):
            return stark_prove(table_ids, traces, check_degree, max_bytes, compress_challenges)
    k = len(table_ids)
    trs = [np.ascontiguousarray(t, dtype=np.uint64) for t in traces]
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    ptrs = (_u64p * k)(*[_p(t) for t in trs])
    logs = (ctypes.c_uint32 * k)(*[int(t.shape[1]).bit_length() - 1 for t in trs])
    out = (ctypes.c_uint8 * max_bytes)()
    n = ctypes.c_size_t(0)
    err = ctypes.create_string_buffer(512)
    cc = None
    if compress_challenges is not None:
        cc_arr = np.ascontiguousarray(compress_challenges, dtype=np.uint64)
        assert cc_arr.shape == (k,)
        cc = _p(cc_arr)
    rc = lib().orc_stark_prove(ids, k, ptrs, logs, cc, 1 if check_degree else 0, out, max_bytes, ctypes.byref(n), err, 512)
    if rc != 0:
        raise StarkError(err.value.decode())
    return bytes(bytearray(out)[: n.value])


def stark_verify(table_ids, proof, hasher_id=None):
    """Buffer::read_all_proof + verify_proof -> (ok, message)."""
    if hasher_id is not None:
        with hasher(hasher_id):
            return stark_verify(table_ids, proof)
    k = len(table_ids)
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    buf = (ctypes.c_uint8 * len(proof)).from_buffer_copy(proof)
    err = ctypes.create_string_buffer(512)
    rc = lib().orc_stark_verify(ids, k, buf, len(proof), err, 512)
    return rc == 0, err.value.decode()
