"""Trace-generation tail: ola_compress_challenge (host) and ola_generate_poseidon_trace (GPU) against
the oracle's restatements of generation/builtin.rs:118-131, generation/prog.rs:23-29, generation/poseidon.rs:5-130."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001


@pytest.mark.parametrize("ncols,n", [(0, 0), (1, 1), (1, 8), (3, 7), (12, 256), (12, 1000)])
def test_compress_challenge_equals_oracle(orc, ncols, n):
    from olavm_b200 import generation

    rng = np.random.default_rng(ncols * 1000 + n)
    cols = [rng.integers(0, P, size=n, dtype=np.uint64) for _ in range(ncols)]
    if ncols and n:
        cols[0][0] = np.uint64(P + 5)  # a non-canonical representative is observed as its canonical value
    assert generation.compress_challenge(cols) == orc.compress_challenge(cols)


def test_program_compress_challenge_is_the_roots_transcript(orc):
    from olavm_b200 import generation

    # generate_prog_trace observes start_root[i], end_root[i] for i in 0..4 as single elements: one column of 8
    roots = np.arange(1, 9, dtype=np.uint64)
    assert generation.compress_challenge([roots]) == orc.compress_challenge([roots])
    assert generation.compress_challenge([roots]) != generation.compress_challenge([roots[::-1].copy()])


@pytest.mark.gpu
@pytest.mark.parametrize("k,log_n", [(0, 1), (1, 1), (5, 3), (300, 9), (1 << 12, 12)])
def test_generate_poseidon_trace_equals_oracle_rows(ctx, orc, k, log_n):
    from olavm_b200 import generation

    rng = np.random.default_rng(k + 1)
    inputs = rng.integers(0, P, size=(k, 12), dtype=np.uint64)
    filters = np.zeros((k, 4), dtype=np.uint64)
    if k:
        filters[np.arange(k), rng.integers(0, 4, size=k)] = 1
        inputs[0, 0] = np.uint64(P + 1)
    t = generation.generate_poseidon_trace(ctx, inputs, filters, log_n)
    assert t.shape == (134, 1 << log_n)
    pad = orc.poseidon_table_row(np.zeros(12, dtype=np.uint64))
    check = range(k) if k <= 300 else list(range(0, k, 97)) + [k - 1]
    for i in check:
        ref = orc.poseidon_table_row(inputs[i])
        ref[0:4] = filters[i]
        assert (t[:, i] == ref).all(), i
    assert (t[:, k:] == pad[:, None]).all()
    # the generated table satisfies the Poseidon AIR wherever the filters allow the input (filters 1-3 pin the capacity)
    free = generation.generate_poseidon_trace(ctx, inputs, None, log_n)
    assert orc.air_first_failure(5, free) is None
    # and the permutation outputs agree with the hash kernels
    if k:
        from olavm_b200 import hashing

        out = hashing.poseidon(ctx, inputs % np.uint64(P))
        assert (free[16:28, :k].T == out).all()


# ---- permuted_cols / generate_rc_trace (stark/lookup.rs:68-131, generation/builtin.rs:249-316) ----------------------------------
def _lookup_case(rng, kind, n):
    """Input / table pairs that reach every branch of the reference's merge walk: values absent from the table, surplus on
    either side, duplicated table entries, inputs above the table's maximum, random 64-bit columns."""
    if kind == 0:  # a valid lookup into 0 .. n-1
        return rng.integers(0, n, size=n, dtype=np.uint64), np.arange(n, dtype=np.uint64)
    if kind == 1:  # the padded fixed table of the RangeCheck AIR (last value repeated)
        return rng.integers(0, n // 2 + 1, size=n, dtype=np.uint64), np.minimum(np.arange(n), n // 2).astype(np.uint64)
    if kind == 2:  # few distinct values, inputs partly above max(table)
        return rng.integers(0, 12, size=n, dtype=np.uint64), rng.integers(0, 8, size=n, dtype=np.uint64)
    if kind == 3:  # random field elements: (almost) nothing matches
        return rng.integers(0, P, size=n, dtype=np.uint64), rng.integers(0, P, size=n, dtype=np.uint64)
    if kind == 4:  # inputs below min(table): pops on an empty list
        return rng.integers(0, 16, size=n, dtype=np.uint64), rng.integers(0, 4, size=n, dtype=np.uint64) + np.uint64(5)
    if kind == 5:  # a permutation of the table
        t = rng.integers(0, 6, size=n, dtype=np.uint64)
        a = t.copy()
        rng.shuffle(a)
        return a, t
    t = rng.integers(0, 5, size=n, dtype=np.uint64) * np.uint64(3)  # gaps in the table
    return rng.integers(0, 20, size=n, dtype=np.uint64), t


def test_oracle_permuted_cols_equals_the_python_restatement(orc):
    """Two independent restatements of lookup.rs:68-131 (oracle/lookup.c, workload/tracegen.py) agree; the permuted pair
    satisfies eval_lookups' row relation (lookup.rs:13-35) whenever every input value occurs in the table."""
    from workload import tracegen

    rng = np.random.default_rng(11)
    for trial in range(140):
        n = 1 << int(rng.integers(1, 8))
        a, t = _lookup_case(rng, trial % 7, n)
        pi, pt = orc.permuted_cols(a, t)
        ri, rt = tracegen.permuted_cols(a, t)
        assert (pi == ri).all() and (pt == rt).all(), (trial, n)
        assert sorted(pt.tolist()) == sorted((t % np.uint64(P)).tolist())  # a permutation of the table
        if trial % 7 in (0, 1, 5):  # valid lookups: each row repeats the previous input or equals its table entry
            same_as_prev = np.concatenate([[False], pi[1:] == pi[:-1]])
            assert (same_as_prev | (pi == pt)).all() and pi[0] == pt[0]


@pytest.mark.gpu
def test_permuted_cols_equals_oracle(ctx, orc):
    from olavm_b200 import generation

    rng = np.random.default_rng(12)
    for trial in range(70):
        n = 1 << int(rng.integers(1, 13))
        a, t = _lookup_case(rng, trial % 7, n)
        if trial % 5 == 0 and int(a[0]) < (1 << 32):
            a[0] += np.uint64(P)  # a non-canonical representative sorts as its canonical value
        pi, pt = generation.permuted_cols(ctx, a, t)
        ri, rt = orc.permuted_cols(a, t)
        assert (pi == ri).all(), (trial, n, "inputs")
        assert (pt == rt).all(), (trial, n, "table")


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [16, 18, 20])
def test_permuted_cols_large(ctx, orc, log_n):
    from olavm_b200 import generation

    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    table = np.minimum(np.arange(n), 65535).astype(np.uint64)  # FIX_RANGE_CHECK_U16 padded with its last value
    inputs = rng.integers(0, 1 << 16, size=n, dtype=np.uint64)
    inputs[: n // 3] = 0  # padding rows of the RangeCheck table: limb 0
    pi, pt = generation.permuted_cols(ctx, inputs, table)
    ri, rt = orc.permuted_cols(inputs, table)
    assert (pi == ri).all() and (pt == rt).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nrows,log_n", [(0, 16), (1, 16), (1000, 16), (70000, 17), (1 << 18, 18)])
def test_generate_rc_trace_equals_oracle_and_satisfies_the_air(ctx, orc, nrows, log_n):
    from olavm_b200 import generation

    rng = np.random.default_rng(nrows + 3)
    vals = rng.integers(0, 1 << 32, size=nrows, dtype=np.uint64)
    kinds = rng.integers(0, 4, size=nrows, dtype=np.uint64)
    t = generation.generate_rc_trace(ctx, vals, kinds, log_n)
    ref = orc.generate_rc_trace(vals, kinds.astype(np.uint8))
    if ref.shape[1] < (1 << log_n):  # the oracle sizes the table itself (next power of two, at least 2^16)
        assert nrows <= ref.shape[1]
    else:
        assert t.shape == ref.shape and (t == ref).all()
    # "all constraints vanish on a real trace" (rangecheck_stark.rs test) on the generated table: table id 4 = RangeCheck
    assert orc.air_first_failure(4, t) is None


# ---- generate_bitwise_trace / generate_cmp_trace (generation/builtin.rs:35-247) ---------------------------------------------------
def _bitwise_ops(rng, k, bits=24):
    tags = np.array([1 << 18, 1 << 17, 1 << 16], dtype=np.uint64)[rng.integers(0, 3, size=k)]  # 1 << Opcode::{AND, OR, XOR}
    a = rng.integers(0, 1 << bits, size=k, dtype=np.uint64)
    b = rng.integers(0, 1 << bits, size=k, dtype=np.uint64)
    r = np.where(tags == (1 << 18), a & b, np.where(tags == (1 << 17), a | b, a ^ b)).astype(np.uint64)
    return tags, a, b, r


def test_oracle_bitwise_trace_follows_the_reference_including_its_fourth_limb(orc):
    """The reference writes limb 3 of op0 / op1 / res to the first column of the NEXT range (builtin.rs:66, :71, :76), where it
    is overwritten: columns 8, 12, 16 stay zero.  The restated generator reproduces that table; with operands below 2^24 it
    satisfies the Bitwise AIR (table id 2), with a 32-bit operand the limb sum check fails as it does in the reference."""
    rng = np.random.default_rng(21)
    tags, a, b, r = _bitwise_ops(rng, 300)
    t, beta = orc.generate_bitwise_trace(tags, a, b, r)
    assert t.shape == (59, 1 << 18)
    assert not t[8].any() and not t[12].any() and not t[16].any()
    assert (t[5, :300] == (a & 255)).all() and (t[11, :300] == ((b >> 16) & 255)).all() and (t[13, :300] == (r & 255)).all()
    assert beta == orc.compress_challenge([t[c] for c in range(5, 17)])
    assert orc.air_first_failure(2, t, compress_challenge=beta) is None
    tags2, a2, b2, r2 = _bitwise_ops(rng, 4, bits=32)
    a2[0] |= np.uint64(1 << 31)
    r2 = np.where(tags2 == (1 << 18), a2 & b2, np.where(tags2 == (1 << 17), a2 | b2, a2 ^ b2)).astype(np.uint64)
    t2, beta2 = orc.generate_bitwise_trace(tags2, a2, b2, r2)
    assert orc.air_first_failure(2, t2, compress_challenge=beta2) is not None


def test_oracle_cmp_trace(orc):
    rng = np.random.default_rng(22)
    cells = []
    for _ in range(5):
        x, y = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
        d = abs(x - y)
        cells.append([x, y, int(x >= y), d, pow(d, P - 2, P) if d else 0, 1])
    t = orc.generate_cmp_trace(np.array(cells, dtype=np.uint64))
    assert t.shape == (6, 8) and (t[:, 5:] == np.array([1, 0, 1, 1, 1, 0], dtype=np.uint64)[:, None]).all()
    assert orc.air_first_failure(3, t) is None


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1, 300, 5000])
def test_generate_bitwise_trace_equals_oracle(ctx, orc, k):
    from olavm_b200 import generation

    rng = np.random.default_rng(30 + k)
    tags, a, b, r = _bitwise_ops(rng, k, bits=32 if k == 300 else 24)
    t, beta = generation.generate_bitwise_trace(ctx, tags, a, b, r)
    ref, rbeta = orc.generate_bitwise_trace(tags, a, b, r)
    assert beta == rbeta
    assert t.shape == ref.shape
    bad = [c for c in range(59) if not (t[c] == ref[c]).all()]
    assert not bad, bad
    if k != 300:
        assert orc.air_first_failure(2, t, compress_challenge=beta) is None


@pytest.mark.gpu
@pytest.mark.parametrize("k,log_n", [(0, 1), (3, 2), (1000, 10), (1 << 16, 16)])
def test_generate_cmp_trace_equals_oracle(ctx, orc, k, log_n):
    from olavm_b200 import generation

    rng = np.random.default_rng(40 + k)
    x = rng.integers(0, 1 << 32, size=k, dtype=np.uint64)
    y = rng.integers(0, 1 << 32, size=k, dtype=np.uint64)
    d = np.where(x >= y, x - y, y - x).astype(np.uint64)
    inv = np.array([pow(int(v), P - 2, P) if v else 0 for v in d[: min(k, 2000)]] + [1] * max(0, k - 2000), dtype=np.uint64)
    cells = np.stack([x, y, (x >= y).astype(np.uint64), d, inv, np.ones(k, dtype=np.uint64)], axis=1) if k else np.zeros((0, 6), dtype=np.uint64)
    t = generation.generate_cmp_trace(ctx, cells, log_n)
    ref = orc.generate_cmp_trace(cells)
    assert ref.shape[1] <= (1 << log_n)
    assert (t[:, : ref.shape[1]] == ref).all()
    assert (t[:, ref.shape[1]:] == np.array([1, 0, 1, 1, 1, 0], dtype=np.uint64)[:, None]).all()


@pytest.mark.gpu
def test_tables_generated_on_the_device_are_proven_where_they_lie(ctx, orc):
    """generation -> prove without a host round trip: Cmp and RangeCheck tables are generated INTO device memory
    (on_device = 1), ola_prove consumes them there (on_device = 1), the subsystem verifier accepts the proof, and the bytes
    equal the oracle prover's on the host copies of the same tables.  The cmp -> rangecheck lookup carries real data: every
    comparison's |a - b| is a RangeCheck row looked up by the Cmp table."""
    import ctypes

    import olavm_b200
    from olavm_b200 import _lib

    rng = np.random.default_rng(77)
    k = 3000
    x = rng.integers(0, 1 << 32, size=k, dtype=np.uint64)
    y = rng.integers(0, 1 << 32, size=k, dtype=np.uint64)
    d = np.where(x >= y, x - y, y - x).astype(np.uint64)
    inv = np.array([pow(int(v), P - 2, P) if v else 0 for v in d], dtype=np.uint64)
    cells = np.stack([x, y, (x >= y).astype(np.uint64), d, inv, np.ones(k, dtype=np.uint64)], axis=1)
    kinds = np.full(k, 3, dtype=np.uint64)  # looked up by the comparison table
    log_cmp, log_rc = 12, 16
    lib = ctx._lib
    d_cells, d_vals, d_kinds = ctx.upload(cells), ctx.upload(d), ctx.upload(kinds)
    d_cmp, d_rc = ctx.alloc(6 << log_cmp), ctx.alloc(12 << log_rc)
    try:
        ctx.check(lib.ola_generate_cmp_trace(ctx.handle, d_cells, k, log_cmp, d_cmp, 1))
        ctx.check(lib.ola_generate_rangecheck_trace(ctx.handle, d_vals, d_kinds, k, log_rc, d_rc, 1))
        proof = olavm_b200.prove_with_device_traces(ctx, [3, 4], [d_cmp, d_rc], [log_cmp, log_rc])
        cmp_host = ctx.download(d_cmp, (6, 1 << log_cmp))
        rc_host = ctx.download(d_rc, (12, 1 << log_rc))
    finally:
        for p in (d_cells, d_vals, d_kinds, d_cmp, d_rc):
            ctx.free(p)
    ok, why = olavm_b200.verify_subsystem_proof([3, 4], proof)
    assert ok, why
    assert orc.air_first_failure(3, cmp_host) is None and orc.air_first_failure(4, rc_host) is None
    assert proof == orc.stark_prove([3, 4], [cmp_host, rc_host], check_degree=True)


# ---- generate_cpu_trace (generation/cpu.rs:11-218) ---------------------------------------------------------------------------------
_CPU_PROGRAMS = ("fibo_recursive", "memory", "mem_gep", "call", "tape", "bitwise", "comparison", "range_check", "context_fetch", "storage",
                 "fibo_loop", "malloc", "poseidon_hash")


def _vm_run(orc, name):
    """One of the reference's assembly test programs through the restated VM -> (its CPU table, the Step records)."""
    import json
    import os

    from workload import tracegen

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_programs.json")))
    prog, prophets = tracegen.parse_ola_prophets({"program": g["programs"][name], "prophets": g["prophets"].get(name, [])})
    if name in tracegen.REFERENCE_CALLDATA:
        tape = tracegen.reference_test_tape(tracegen.REFERENCE_CALLDATA[name])
    else:
        tape = tracegen.CONTEXT_TAPE if name == "context_fetch" else ()
    ids, traces, cc, steps = tracegen.run_system(orc, np.random.default_rng(3), prog, prophets=prophets, init_tape=tape)
    return traces[0], tracegen.steps_to_records(steps)


def _random_step_records(rng, k):
    """Records with every opcode, ext lines, non-zero env / call counters: the branches of cpu.rs:111-177."""
    r = rng.integers(0, P, size=(k, 66), dtype=np.uint64)
    shifts = rng.integers(7, 32, size=k)
    r[:, 27] = np.left_shift(np.uint64(1), shifts.astype(np.uint64))
    r[:, 27][rng.random(k) < 0.05] = 12345          # an opcode that is no instruction: no selector
    r[:, 0] = np.where(rng.random(k) < 0.5, 0, rng.integers(1, 5, size=k)).astype(np.uint64)   # env_idx
    r[:, 13] = rng.integers(0, 2, size=k)           # is_ext_line
    r[:, 14] = rng.integers(0, 3, size=k)           # ext_cnt
    r[:, 26] = rng.integers(0, 2, size=k)           # op1_imm
    r[:, 29] = rng.integers(0, 2, size=k)           # op0 (a flag on tload rows)
    r[:, 30] = rng.integers(0, 4, size=k)           # op1 (a length on tload / tstore rows)
    r[:, 11] = rng.integers(0, 1 << 32, size=k)     # clk is a u32
    return r


@pytest.mark.parametrize("name", _CPU_PROGRAMS)
def test_oracle_cpu_trace_equals_the_vm_table(orc, name):
    """Two restatements of generate_cpu_trace -- oracle/generation_cpu.c from the Rust, and the test VM's own table fill
    (workload/tracegen.py, which carries its own ext_length bookkeeping) -- agree on the reference's assembly programs."""
    table, records = _vm_run(orc, name)
    got = orc.generate_cpu_trace(records, int(table.shape[1]).bit_length() - 1)
    bad = [c for c in range(94) if not (got[c] == table[c]).all()]
    assert not bad, bad
    assert orc.air_first_failure(0, got) is None or name in ()  # the generated table satisfies the CPU AIR


def test_oracle_cpu_trace_of_no_steps_is_the_padding_table(orc):
    from workload import tracegen

    assert (orc.generate_cpu_trace(np.zeros((0, 66), dtype=np.uint64), 4) == tracegen.cpu_padding_trace(4)).all()


@pytest.mark.gpu
def test_generate_cpu_trace_equals_oracle(ctx, orc):
    from olavm_b200 import generation

    for name in ("fibo_recursive", "tape", "storage", "fibo_loop"):
        table, records = _vm_run(orc, name)
        log_n = int(table.shape[1]).bit_length() - 1
        got = generation.generate_cpu_trace(ctx, records, log_n)
        assert (got == table).all(), name
    rng = np.random.default_rng(8)
    for k, log_n in ((0, 3), (1, 0), (777, 10), (1 << 14, 14), (50000, 16)):
        rec = _random_step_records(rng, k)
        got = generation.generate_cpu_trace(ctx, rec, log_n)
        ref = orc.generate_cpu_trace(rec, log_n)
        bad = [c for c in range(94) if not (got[c] == ref[c]).all()]
        assert not bad, (k, bad)


# ---- generate_memory_trace (generation/memory.rs:8-155) ----------------------------------------------------------------------------
_MEM_PROGRAMS = ("fibo_recursive", "memory", "mem_gep", "call", "storage", "malloc", "poseidon_hash", "fibo_loop", "global")


def _vm_memory(orc, name):
    """A reference program through the restated VM -> (its Memory table, the MemoryTraceCell records, log_n)."""
    import json
    import os

    from workload import tracegen

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_programs.json")))
    prog, prophets = tracegen.parse_ola_prophets({"program": g["programs"][name], "prophets": g["prophets"].get(name, [])})
    tape = tracegen.reference_test_tape(tracegen.REFERENCE_CALLDATA[name]) if name in tracegen.REFERENCE_CALLDATA else ()
    mem_log = tracegen.cpu_vm_trace(prog, 13, want_side_tables="all+storage", orc=orc, init_tape=tape, prophets=prophets)[5]
    log_n = max(1, (len(mem_log) - 1).bit_length())
    table, cells = tracegen.memory_trace_from_log(mem_log, log_n, want_cells=True)
    return table, tracegen.memory_cells_to_records(cells), log_n


@pytest.mark.parametrize("name", _MEM_PROGRAMS)
def test_oracle_memory_trace_equals_the_vm_table(orc, name):
    """oracle/generation_cpu.c's generate_memory_trace (from the Rust) against the test VM's own Memory table (stack, heap and
    write-once regions, padding that continues the write-once region), and against the Memory AIR."""
    table, records, log_n = _vm_memory(orc, name)
    got = orc.generate_memory_trace(records, log_n)
    bad = [c for c in range(29) if not (got[c] == table[c]).all()]
    assert not bad, bad
    assert orc.air_first_failure(1, got) is None


def test_oracle_memory_trace_of_no_cells(orc):
    t = orc.generate_memory_trace(np.zeros((0, 15), dtype=np.uint64), 2)
    span = (1 << 32) - 1
    assert t[3].tolist() == [P - span, P - span + 1, P - span + 2, P - span + 3] and t[24].tolist() == [1, 1, 1, 1]
    assert t[16].tolist() == [0, 1, 1, 1] and t[19].tolist() == [0, 1, 1, 1] and t[26].tolist() == [span, span - 1, span - 2, span - 3]


@pytest.mark.gpu
def test_generate_memory_trace_equals_oracle(ctx, orc):
    from olavm_b200 import generation

    for name in ("fibo_recursive", "storage", "malloc", "global"):
        table, records, log_n = _vm_memory(orc, name)
        for extra in (0, 1):  # also with a larger table: more padding rows
            got = generation.generate_memory_trace(ctx, records, log_n + extra)
            assert (got == orc.generate_memory_trace(records, log_n + extra)).all(), (name, extra)
            if not extra:
                assert (got == table).all(), name
    assert (generation.generate_memory_trace(ctx, np.zeros((0, 15), dtype=np.uint64), 3) == orc.generate_memory_trace(np.zeros((0, 15), dtype=np.uint64), 3)).all()
    rng = np.random.default_rng(9)
    for k, log_n in ((1, 1), (5, 3), (4000, 12), (1 << 15, 15)):
        rec = rng.integers(0, P, size=(k, 15), dtype=np.uint64)
        ops = np.array([0, 1 << 22, 1 << 21, 1 << 24, 1 << 23, 1 << 9, 1 << 8, 1 << 7, 1 << 12, 1 << 10, 1 << 11, 12345], dtype=np.uint64)
        rec[:, 4] = ops[rng.integers(0, len(ops), size=k)]
        rec[:, 1] = rng.integers(0, 2, size=k)
        rec[:, 12] = rng.integers(0, 2, size=k)
        rec[:, 13] = rng.integers(0, 2, size=k)
        got = generation.generate_memory_trace(ctx, rec, log_n)
        ref = orc.generate_memory_trace(rec, log_n)
        bad = [c for c in range(29) if not (got[c] == ref[c]).all()]
        assert not bad, (k, bad)


# ---- generate_prog_trace (generation/prog.rs:18-157) --------------------------------------------------------------------------------
def _vm_program(orc, name):
    """A reference program through the restated VM -> (Step records, program lines, executed lines, program)."""
    import json
    import os

    from workload import tracegen

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_programs.json")))
    prog, prophets = tracegen.parse_ola_prophets({"program": g["programs"][name], "prophets": g["prophets"].get(name, [])})
    tape = tracegen.reference_test_tape(tracegen.REFERENCE_CALLDATA[name]) if name in tracegen.REFERENCE_CALLDATA else ()
    steps = tracegen.cpu_vm_trace(prog, 13, want_side_tables="all+storage", orc=orc, init_tape=tape, prophets=prophets)[1]
    prog_rows, exec_rows = tracegen.program_rows_of_run(prog, steps)
    return tracegen.steps_to_records(steps), np.array(prog_rows, dtype=np.uint64), exec_rows


_ROOTS = np.arange(11, 19, dtype=np.uint64)


@pytest.mark.parametrize("name", ("fibo_recursive", "memory", "call", "tape", "storage", "fibo_loop", "poseidon_hash"))
def test_oracle_prog_trace_equals_the_python_generator(orc, name):
    from workload import tracegen

    records, prog_rows, exec_rows = _vm_program(orc, name)
    got, beta = orc.generate_prog_trace(records, prog_rows, _ROOTS)
    inter = np.empty(8, dtype=np.uint64)
    inter[0::2], inter[1::2] = _ROOTS[:4], _ROOTS[4:]
    assert beta == orc.compress_challenge([inter])   # start[i], end[i] interleaved (prog.rs:25-28)
    ref = tracegen.program_valid_trace(np.random.default_rng(0), int(got.shape[1]).bit_length() - 1, beta,
                                       prog_rows=[tuple(int(x) for x in r) for r in prog_rows], exec_rows=exec_rows)
    bad = [c for c in range(18) if not (got[c] == ref[c]).all()]
    assert not bad, bad
    assert orc.air_first_failure(10, got, compress_challenge=beta) is None


@pytest.mark.gpu
def test_generate_prog_trace_equals_oracle(ctx, orc):
    from olavm_b200 import generation

    for name in ("fibo_recursive", "storage", "fibo_loop"):
        records, prog_rows, _ = _vm_program(orc, name)
        ref, rbeta = orc.generate_prog_trace(records, prog_rows, _ROOTS)
        log_n = int(ref.shape[1]).bit_length() - 1
        got, beta = generation.generate_prog_trace(ctx, records, prog_rows, _ROOTS, log_n)
        assert beta == rbeta
        bad = [c for c in range(18) if not (got[c] == ref[c]).all()]
        assert not bad, (name, bad)
    # a long run: many executed words over a short program, the lookup at 2^17 rows
    rng = np.random.default_rng(10)
    prog_rows = np.stack([np.zeros(500, dtype=np.uint64)] * 4 + [np.arange(500, dtype=np.uint64), rng.integers(0, P, size=500, dtype=np.uint64)], axis=1)
    k = 90000
    rec = np.zeros((k, 66), dtype=np.uint64)
    pcs = rng.integers(0, 499, size=k)
    rec[:, 12] = pcs
    rec[:, 25] = prog_rows[pcs, 5]
    rec[:, 26] = rng.integers(0, 2, size=k)            # op1_imm: a second fetched word
    rec[:, 28] = prog_rows[pcs + 1, 5]
    rec[:, 27] = np.uint64(1 << 31)
    rec[:, 13] = (rng.random(k) < 0.1).astype(np.uint64)  # ext lines fetch nothing
    ref, rbeta = orc.generate_prog_trace(rec, prog_rows, _ROOTS)
    got, beta = generation.generate_prog_trace(ctx, rec, prog_rows, _ROOTS, int(ref.shape[1]).bit_length() - 1)
    assert beta == rbeta and (got == ref).all()


# ---- the five small tables: PoseidonChunk, StorageAccess, Tape, SCCall, ProgChunk -----------------------------------------------------
_SMALL_PROGRAMS = ("poseidon_hash", "storage", "tape", "fibo_loop", "context_fetch")


def _vm_tables(orc, name):
    """A reference program through the restated VM -> {table id: table} of the whole system and the program's words."""
    import json
    import os

    from workload import tracegen

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_programs.json")))
    prog, prophets = tracegen.parse_ola_prophets({"program": g["programs"][name], "prophets": g["prophets"].get(name, [])})
    if name in tracegen.REFERENCE_CALLDATA:
        tape = tracegen.reference_test_tape(tracegen.REFERENCE_CALLDATA[name])
    else:
        tape = tracegen.CONTEXT_TAPE if name == "context_fetch" else ()
    ids, traces, cc, steps = tracegen.run_system(orc, np.random.default_rng(3), prog, prophets=prophets, init_tape=tape)
    return dict(zip(ids, traces))


def _tape_records_of_table(t):
    """TapeRow records of a Tape table: the rows up to the last one that is not the padding form of its predecessor."""
    n = t.shape[1]
    k = n
    while k > 1 and t[1, k - 1] == t[1, k - 2] and t[2, k - 1] == (1 << 9) and t[3, k - 1] == t[3, k - 2] and t[4, k - 1] == t[4, k - 2] \
            and t[5, k - 1] == 0:
        k -= 1
    return np.ascontiguousarray(t[1:6, :k].T)


@pytest.mark.parametrize("name", _SMALL_PROGRAMS)
def test_oracle_small_tables_equal_the_vm_tables(orc, name):
    """The VM's tables satisfy the AIRs and prove (tests/test_oracle_stark.py); the oracle's generators, fed the executor records
    read back from those tables, must rebuild them: every derived flag column and every padding row."""
    from workload import tracegen

    tabs = _vm_tables(orc, name)
    checked = 0
    if 6 in tabs:
        t = tabs[6]
        got = orc.generate_poseidon_chunk_trace(tracegen.poseidon_chunk_records_of_table(t), int(t.shape[1]).bit_length() - 1)
        assert not [c for c in range(53) if not (got[c] == t[c]).all()]
        checked += 1
    if 7 in tabs:
        t = tabs[7]
        rec, n_access = tracegen.storage_records_of_table(t)
        got = orc.generate_storage_access_trace(rec[:n_access], rec[n_access:], int(t.shape[1]).bit_length() - 1)
        assert not [c for c in range(48) if not (got[c] == t[c]).all()]
        checked += 1
    if 8 in tabs:
        t = tabs[8]
        got = orc.generate_tape_trace(_tape_records_of_table(t), int(t.shape[1]).bit_length() - 1)
        assert (got == t).all()
        checked += 1
    assert checked >= 1


@pytest.mark.parametrize("name", ("fibo_loop", "storage", "poseidon_hash"))
def test_oracle_prog_chunk_trace_equals_the_vm_table(orc, name):
    """One program: the VM's ProgChunk table (lines of eight words, sponge in overwrite mode, capacity reset on the first line)
    is the reference's; it satisfies the AIR."""
    from workload import tracegen

    _, prog_rows, _ = _vm_program(orc, name)
    words = [int(w) for w in prog_rows[:, 5]]
    addr = [int(x) for x in prog_rows[0, :4]]
    lines = (len(words) + 7) // 8
    log_n = max(1, (lines - 1).bit_length())
    t = tracegen.prog_chunk_valid_trace(orc, np.random.default_rng(0), log_n, programs=[(addr, words)])[0]
    got = orc.generate_prog_chunk_trace(prog_rows)
    assert got.shape == t.shape
    assert not [c for c in range(40) if not (got[c] == t[c]).all()]
    assert orc.air_first_failure(11, got) is None


def _random_small_records(rng, kind, k):
    if kind == "poseidon_chunk":
        r = rng.integers(0, P, size=(k, 32), dtype=np.uint64)
        r[:, 1] = rng.integers(0, 1 << 32, size=k)
        r[:, 31] = rng.integers(0, 2, size=k)
        r[:, 5] = rng.integers(1, 40, size=k)
        r[:, 6] = np.where(rng.random(k) < 0.5, r[:, 5], rng.integers(0, 40, size=k).astype(np.uint64))
        return r
    if kind == "storage":
        r = rng.integers(0, P, size=(k, 38), dtype=np.uint64)
        r[:, 10] = rng.choice([1, 2, 63, 64, 65, 127, 128, 191, 192, 255, 256, 300], size=k)
        r[:, 11] = rng.integers(0, 3, size=k)
        r[:, 9] = rng.integers(0, 2, size=k)
        return r
    if kind == "tape":
        r = rng.integers(0, P, size=(k, 5), dtype=np.uint64)
        r[:, 0] = rng.integers(0, 2, size=k)
        return r
    return rng.integers(0, P, size=(k, 24), dtype=np.uint64)


def test_oracle_small_tables_row_counts_and_padding(orc):
    """The reference's row count (next power of two, at least 2) and the padding rows of each table."""
    rng = np.random.default_rng(5)
    for k, n in ((0, 2), (1, 2), (2, 2), (3, 4), (8, 8), (9, 16)):
        assert orc.generate_poseidon_chunk_trace(_random_small_records(rng, "poseidon_chunk", k)).shape == (53, n)
        assert orc.generate_sccall_trace(_random_small_records(rng, "sccall", k)).shape == (26, n)
        rec = _random_small_records(rng, "tape", k)
        t = orc.generate_tape_trace(rec)
        assert t.shape == (6, n)
        if k < n:
            assert (t[2, k:] == 1 << 9).all() and (t[5, k:] == 0).all() and (t[0] == 0).all()
            assert (t[3, k:] == (rec[k - 1, 2] % P if k else 0)).all() and (t[4, k:] == (rec[k - 1, 3] % P if k else 0)).all()
        rec = _random_small_records(rng, "storage", k)
        t = orc.generate_storage_access_trace(rec[: k // 2], rec[k // 2:])
        assert t.shape == (48, n) and (t[47, k:] == 1).all() and (t[47, :k] == 0).all()
        if 0 < k < n:
            assert (t[5:9, k:] == (rec[k - 1, 5:9] % P)[:, None]).all()
        assert (t[46, : k // 2] == 0).all() and (t[46, k // 2:k] == (rec[k // 2:, 10] == 256)).all()
    # two programs: the sponge state runs on from the first program's last line into the second's first line (prog.rs:199)
    words = rng.integers(0, P, size=13 + 5, dtype=np.uint64)
    rows = [(1, 2, 3, 4, pc, int(words[pc])) for pc in range(13)] + [(5, 6, 7, 8, pc, int(words[13 + pc])) for pc in range(5)]
    t = orc.generate_prog_chunk_trace(np.array(rows, dtype=np.uint64))
    assert t.shape == (40, 4) and list(t[39]) == [0, 0, 0, 1] and list(t[29]) == [1, 0, 1, 0] and list(t[30]) == [0, 1, 1, 0]
    assert list(t[4]) == [0, 8, 0, 0] and list(t[31:39, 1]) == [1] * 5 + [0] * 3
    assert (t[10:13, 1] == t[22:25, 0]).all()      # unused slots 5..7 of line 1 = hash[5..8] of line 0
    assert (t[13:17, 2] == t[25:29, 1]).all() and t[13:17, 2].any()   # the second program's first line carries the capacity
    h = orc.poseidon(np.concatenate([t[5:13, 1], t[13:17, 1]]))
    assert (t[17:29, 1] == h).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k,log_n", [(0, 1), (1, 1), (5, 3), (1000, 10), (40000, 16)])
def test_generate_small_tables_equal_oracle(ctx, orc, k, log_n):
    from olavm_b200 import generation

    rng = np.random.default_rng(k + 77)
    rec = _random_small_records(rng, "poseidon_chunk", k)
    if k:
        rec[0, 0] = np.uint64(P + 3)   # a non-canonical representative lands in the table as its canonical value
    assert (generation.generate_poseidon_chunk_trace(ctx, rec, log_n) == orc.generate_poseidon_chunk_trace(rec, log_n)).all()
    rec = _random_small_records(rng, "storage", k)
    for n_access in (0, k // 3, k):
        got = generation.generate_storage_access_trace(ctx, rec[:n_access], rec[n_access:], log_n)
        assert (got == orc.generate_storage_access_trace(rec[:n_access], rec[n_access:], log_n)).all()
    rec = _random_small_records(rng, "tape", k)
    assert (generation.generate_tape_trace(ctx, rec, log_n) == orc.generate_tape_trace(rec, log_n)).all()
    rec = _random_small_records(rng, "sccall", k)
    assert (generation.generate_sccall_trace(ctx, rec, log_n) == orc.generate_sccall_trace(rec, log_n)).all()


@pytest.mark.gpu
def test_generate_prog_chunk_trace_equals_oracle(ctx, orc):
    from olavm_b200 import generation

    rng = np.random.default_rng(21)
    for lens in ((), (1,), (8,), (13,), (16, 1), (3, 700, 8, 9), (4001,)):
        rows = []
        for p, L in enumerate(lens):
            addr = [int(x) for x in rng.integers(0, P, size=4, dtype=np.uint64)]
            rows += [tuple(addr) + (pc, int(rng.integers(0, P, dtype=np.uint64))) for pc in range(L)]
        rows = np.array(rows, dtype=np.uint64).reshape(-1, 6)
        ref = orc.generate_prog_chunk_trace(rows)
        got = generation.generate_prog_chunk_trace(ctx, rows)
        assert got.shape == ref.shape, lens
        bad = [c for c in range(40) if not (got[c] == ref[c]).all()]
        assert not bad, (lens, bad)
        got = generation.generate_prog_chunk_trace(ctx, rows, int(ref.shape[1]).bit_length() + 1)   # a larger table: more padding lines
        assert (got == orc.generate_prog_chunk_trace(rows, int(ref.shape[1]).bit_length() + 1)).all()
    for name in ("fibo_loop", "storage"):
        _, prog_rows, _ = _vm_program(orc, name)
        ref = orc.generate_prog_chunk_trace(prog_rows)
        assert (generation.generate_prog_chunk_trace(ctx, prog_rows) == ref).all()
        assert orc.air_first_failure(11, ref) is None
