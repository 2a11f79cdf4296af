"""Mirror of `circuits::stark::prover::prove_with_traces` (circuits/src/stark/prover.rs:79-85) followed by
`Buffer::write_all_proof` (circuits/src/stark/serialization.rs:377-393)."""
import ctypes

import numpy as np

from . import _lib

POSEIDON, BLAKE3 = 0, 1  # OLA_HASH_* of include/ola_gpu.h: C::Hasher (Context.hasher, verify_proof(hasher=))
TABLES = dict(cpu=0, memory=1, bitwise=2, cmp=3, rangecheck=4, poseidon=5, poseidon_chunk=6, storage_access=7, tape=8, sccall=9,
              program=10, prog_chunk=11)


def table_columns(ctx, table_id):
    """S::COLUMNS of a table, or -1 if its constraint kernel is not in this build."""
    return int(ctx._lib.ola_table_columns(int(table_id)))


def prove_with_traces(ctx, table_ids, trace_poly_values, check_quotient_degree=True, max_bytes=1 << 26, compress_challenges=None):
    """-> proof bytes (AllProof in the reference's wire format).

    table_ids: ids of the reference `Table` enum, ascending; trace_poly_values[i]: [columns_i, 2^k_i] uint64
    (Vec<PolynomialValues<F>> column-major); compress_challenges: None or one field element per table (only the
    Bitwise and Program entries are used: the beta of generation/mod.rs:183-188).  Raises OlaError(OLA_ERR_QUOTIENT_DEGREE) where the reference panics with
    "Quotient has failed, ..." and OlaError(OLA_ERR_INVALID_ARG, "Non-binary filter?") like partial_products' assert."""
    k = len(table_ids)
    trs = [np.ascontiguousarray(t, dtype=np.uint64) for t in trace_poly_values]
    if len(trs) != k:
        raise ValueError("one trace per table id")
    for tid, t in zip(table_ids, trs):
        if t.ndim != 2:
            raise ValueError("a trace is a [columns, rows] matrix")
        cols = table_columns(ctx, tid)
        if cols < 0:
            raise ValueError(f"unknown table id {tid}")
        if t.shape[0] != cols:   # the C ABI reads columns * rows elements from each pointer
            raise ValueError(f"table {tid} has {cols} columns, the trace has {t.shape[0]}")
        n = t.shape[1]
        if n == 0 or n & (n - 1):
            raise ValueError("trace length must be a power of 2")
    if compress_challenges is None and any(int(t) in (TABLES["bitwise"], TABLES["program"]) for t in table_ids):
        # generation/mod.rs:183-188 always sets both; proving with beta = 0 would make the compress lookups degenerate
        raise ValueError("the Bitwise and Program tables need their compress challenges (compress_challenges=...)")
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    ptrs = (ctypes.c_void_p * k)(*[t.ctypes.data for t in trs])
    logs = (ctypes.c_uint32 * k)(*[int(t.shape[1]).bit_length() - 1 for t in trs])
    out = np.empty(max_bytes, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    cc = None
    if compress_challenges is not None:
        cc_arr = np.ascontiguousarray(compress_challenges, dtype=np.uint64)
        if cc_arr.shape != (k,):
            raise ValueError("one compress challenge per table")
        cc = cc_arr.ctypes.data_as(ctypes.c_void_p)
    ctx.check(ctx._lib.ola_prove(ctx.handle, ids, k, ptrs, 0, logs, cc, 1 if check_quotient_degree else 0, out.ctypes.data_as(ctypes.c_void_p),
                                 max_bytes, ctypes.byref(n)))
    return out[: n.value].tobytes()


def air_constraints(table_id, lv, nv, compress_challenge=0):
    """The individual constraint values the table's eval_packed_generic emits for one (local, next) row pair over the base
    field, in order, with their kinds (0 constraint, 1 transition, 2 first row, 3 last row): the constraint code of the
    quotient kernels and of verify_proof, evaluated on the host.  -> (values uint64[K], kinds int32[K])."""
    lib = _lib.load()
    lv = np.ascontiguousarray(lv, dtype=np.uint64)
    nv = np.ascontiguousarray(nv, dtype=np.uint64)
    cols = int(lib.ola_table_columns(int(table_id)))
    if cols < 0 or lv.shape != (cols,) or nv.shape != (cols,):
        raise ValueError("unknown table or wrong row width")
    cap = 4096
    vals = np.zeros(cap, dtype=np.uint64)
    kinds = np.zeros(cap, dtype=np.int32)
    k = lib.ola_air_constraints(int(table_id), _lib.hptr(lv), _lib.hptr(nv), int(compress_challenge), _lib.hptr(vals), kinds.ctypes.data_as(ctypes.c_void_p), cap)
    if k < 0 or k > cap:
        raise _lib.OlaError(k, "ola_air_constraints")
    return vals[:k].copy(), kinds[:k].copy()


def prove_with_device_traces(ctx, table_ids, device_ptrs, log_ns, check_quotient_degree=True, max_bytes=1 << 26, compress_challenges=None):
    """prove_with_traces for traces already resident in HBM (ola_prove with on_device = 1): device_ptrs[i] points at table
    i's column-major [columns_i][2^log_ns[i]] u64 block (Context.alloc / Context.upload / generation on the device)."""
    k = len(table_ids)
    if len(device_ptrs) != k or len(log_ns) != k:
        raise ValueError("one device pointer and one log size per table id")
    if compress_challenges is None and any(int(t) in (TABLES["bitwise"], TABLES["program"]) for t in table_ids):
        raise ValueError("the Bitwise and Program tables need their compress challenges (compress_challenges=...)")
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    ptrs = (ctypes.c_void_p * k)(*[p.value if isinstance(p, ctypes.c_void_p) else int(p) for p in device_ptrs])
    logs = (ctypes.c_uint32 * k)(*[int(x) for x in log_ns])
    out = np.empty(max_bytes, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    cc = None
    if compress_challenges is not None:
        cc_arr = np.ascontiguousarray(compress_challenges, dtype=np.uint64)
        if cc_arr.shape != (k,):
            raise ValueError("one compress challenge per table")
        cc = cc_arr.ctypes.data_as(ctypes.c_void_p)
    ctx.check(ctx._lib.ola_prove(ctx.handle, ids, k, ptrs, 1, logs, cc, 1 if check_quotient_degree else 0, out.ctypes.data_as(ctypes.c_void_p),
                                 max_bytes, ctypes.byref(n)))
    return out[: n.value].tobytes()


class TranscriptEvent(ctypes.Structure):
    """ola_transcript_event of include/ola_gpu.h."""
    _fields_ = [("kind", ctypes.c_int), ("stage", ctypes.c_int), ("table", ctypes.c_int), ("elems", ctypes.POINTER(ctypes.c_uint64)), ("count", ctypes.c_size_t)]


EV_OBSERVE, EV_CHALLENGE, EV_COMPACT, EV_DONE, EV_FAILED = 1, 2, 3, 4, 5


def prove_with_challenger(ctx, table_ids, trace_poly_values, challenger, check_quotient_degree=True, max_bytes=1 << 26, compress_challenges=None,
                          log=None):
    """prove_with_traces with the Fiat-Shamir transcript kept by the CALLER (ola_prove_session_*): `challenger` is the host's
    own plonky2 `Challenger` -- any object with observe_elements(list of int), get_n_challenges(n) -> list of int and
    compact().  The library reports every transcript operation as an event; `log`, if a list, receives (kind, stage, table,
    count) per event.  With a faithful Challenger the bytes equal prove_with_traces'."""
    k = len(table_ids)
    trs = [np.ascontiguousarray(t, dtype=np.uint64) for t in trace_poly_values]
    for tid, t in zip(table_ids, trs):
        if t.ndim != 2 or t.shape[0] != table_columns(ctx, tid) or t.shape[1] & (t.shape[1] - 1):
            raise ValueError(f"bad trace for table {tid}")
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    ptrs = (ctypes.c_void_p * k)(*[t.ctypes.data for t in trs])
    logs = (ctypes.c_uint32 * k)(*[int(t.shape[1]).bit_length() - 1 for t in trs])
    cc = None
    if compress_challenges is not None:
        cc_arr = np.ascontiguousarray(compress_challenges, dtype=np.uint64)
        cc = cc_arr.ctypes.data_as(ctypes.c_void_p)
    sess = ctypes.c_void_p()
    lib = ctx._lib
    ctx.check(lib.ola_prove_session_begin(ctx.handle, ids, k, ptrs, 0, logs, cc, 1 if check_quotient_degree else 0, ctypes.byref(sess)))
    ev = TranscriptEvent()
    rc = 0
    try:
        while True:
            rc = lib.ola_prove_session_next(sess, ctypes.byref(ev))
            if log is not None:
                log.append((ev.kind, ev.stage, ev.table, int(ev.count)))
            if rc != 0 or ev.kind in (EV_DONE, EV_FAILED):
                break
            if ev.kind == EV_OBSERVE:
                challenger.observe_elements([int(ev.elems[i]) for i in range(ev.count)])
            elif ev.kind == EV_CHALLENGE:
                vals = np.array([int(x) for x in challenger.get_n_challenges(int(ev.count))], dtype=np.uint64)
                r = lib.ola_prove_session_supply(sess, vals.ctypes.data_as(ctypes.c_void_p), vals.size)
                if r != 0:
                    raise _lib.OlaError(r, "ola_prove_session_supply")
            elif ev.kind == EV_COMPACT:
                challenger.compact()
    finally:
        out = np.empty(max_bytes, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        frc = lib.ola_prove_session_finish(sess, out.ctypes.data_as(ctypes.c_void_p), max_bytes, ctypes.byref(n))
    ctx.check(frc)
    return out[: n.value].tobytes()


def _verify(entry, table_ids, proof, hasher):
    lib = _lib.load()
    k = len(table_ids)
    ids = (ctypes.c_int * k)(*[int(x) for x in table_ids])
    buf = np.frombuffer(bytes(proof), dtype=np.uint8)
    err = ctypes.create_string_buffer(512)
    rc = getattr(lib, entry)(int(hasher), ids, k, buf.ctypes.data_as(ctypes.c_void_p), buf.size, err, 512)
    return rc == 0, err.value.decode()


def verify_proof(table_ids, proof, hasher=0):
    """`circuits::stark::verifier::verify_proof` over `Buffer::read_all_proof`'s bytes (verifier.rs:32-212,
    serialization.rs:395-411) -> (accepted, reason).  Host code: needs the library but no GPU.  Like the reference's it is
    fixed at the full 12-table system (table_ids = 0..11) and checks every cross-table lookup.
    hasher: POSEIDON (PoseidonGoldilocksConfig) or BLAKE3 (Blake3GoldilocksConfig), the config the proof was made under."""
    return _verify("ola_verify_cfg", table_ids, proof, hasher)


def verify_subsystem_proof(table_ids, proof, hasher=0):
    """WEAKER than verify_proof: verifies a proof of an ordered SUBSET of the tables (what prove_with_traces makes when given
    fewer tables).  Lookups with a side outside the subset are not checked.  For tests of subsystems."""
    return _verify("ola_verify_subsystem_cfg", table_ids, proof, hasher)
