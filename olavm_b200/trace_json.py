"""Mirror of the `ola prove` flow (client/src/main.rs:172-207): Trace JSON -> generate_traces -> prove_with_traces -> proof bytes.
The parser is the library's (ola_trace_from_json, host code); generation and the proof run on the GPU and the tables never
leave it."""
import ctypes

import numpy as np

from . import _lib

REC = dict(step=0, memory=1, rc_val=2, rc_kind=3, bitwise_tag=4, bitwise_op0=5, bitwise_op1=6, bitwise_res=7, cmp=8, poseidon_input=9,
           poseidon_filter=10, poseidon_chunk=11, storage_hash=12, tape=13, sccall=14, prog_row=15, roots=16, storage_access_count=17)


class Trace:
    """A parsed core::trace::trace::Trace (trace.rs:320-342): the flat executor records of every table."""

    def __init__(self, text):
        data = text.encode() if isinstance(text, str) else bytes(text)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        err = ctypes.create_string_buffer(256)
        rc = self._lib.ola_trace_from_json(data, len(data), ctypes.byref(h), err, len(err))
        if rc != 0:
            raise ValueError(err.value.decode() or "ola_trace_from_json: error %d" % rc)
        self.handle = h

    @classmethod
    def from_records(cls, **arrays):
        """The trace object over records the caller already holds (what a Rust host does with its own Trace): keyword = a kind of
        REC (`step`, `memory`, `rc_val`, ... `prog_row`, `roots`), value = the record array; `storage_access_count` = int.  The
        arrays are borrowed, not copied: they are kept alive by this object."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.ola_trace_new(ctypes.byref(h))
        if rc != 0:
            raise _lib.OlaError(rc, "ola_trace_new")
        self.handle = h
        self._keep = []
        for kind, a in arrays.items():
            if kind == "storage_access_count":
                rc = self._lib.ola_trace_set_records(h, REC[kind], None, int(a))
            else:
                a = np.ascontiguousarray(a, dtype=np.uint64)
                self._keep.append(a)
                width = 8 if kind == "roots" else None
                n = 1 if kind == "roots" else (a.shape[0] if a.ndim else 0)
                if width and a.size != 8:
                    raise ValueError("roots = start_root[4], end_root[4]")
                rc = self._lib.ola_trace_set_records(h, REC[kind], a.ctypes.data_as(ctypes.c_void_p) if a.size else None, n)
            if rc != 0:
                raise _lib.OlaError(rc, "ola_trace_set_records(%s)" % kind)
        return self

    def records(self, kind):
        rows, n, rec = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32()
        rc = self._lib.ola_trace_records(self.handle, REC[kind], ctypes.byref(rows), ctypes.byref(n), ctypes.byref(rec))
        if rc != 0:
            raise _lib.OlaError(rc, "ola_trace_records")
        if kind == "storage_access_count":
            return int(n.value)
        count = n.value * rec.value
        a = np.ctypeslib.as_array(ctypes.cast(rows, ctypes.POINTER(ctypes.c_uint64)), shape=(count,)).copy() if count else np.zeros(0, dtype=np.uint64)
        return a.reshape(n.value, rec.value) if rec.value > 1 else a

    def table_log_rows(self, table_id):
        v = self._lib.ola_trace_table_log_rows(self.handle, table_id)
        if v < 0:
            raise _lib.OlaError(v, "ola_trace_table_log_rows")
        return v

    def close(self):
        if self.handle:
            self._lib.ola_trace_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_traces(ctx, trace):
    """generate_traces (generation/mod.rs:79-213) -> (12 device pointers, log_ns, compress_challenges); free the pointers with
    ctx.free."""
    tabs = (ctypes.c_void_p * 12)()
    logs = (ctypes.c_uint32 * 12)()
    cc = (ctypes.c_uint64 * 12)()
    ctx.check(ctx._lib.ola_generate_traces(ctx.handle, trace.handle, tabs, logs, cc))
    return [tabs[i] for i in range(12)], [int(x) for x in logs], [int(x) for x in cc]


def prove_trace(ctx, trace, cap=1 << 24):
    """`ola prove`: the proof bytes of the twelve-table system generated from the trace."""
    out = np.empty(cap, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    ctx.check(ctx._lib.ola_prove_trace(ctx.handle, trace.handle, out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n)))
    return out[: n.value].tobytes()
