"""The AIR constraints are transcribed TWICE from the Rust, independently: the product's (olavm_b200/csrc/air/*.h, compiled
into the quotient kernels and ola_verify) and the oracle's (oracle/air_*.hpp, oracle/ctl_registry.hpp).  The consumer is
Horner in alpha (constraint_consumer.rs:60-65), so proofs are bit-identical only if both emit the same constraints in the same
order with the same values.  These tests compare the two element by element: per (table, row pair) the vector of raw
constraint values and their kinds, on random rows (a random point separates different polynomials) and on real rows."""
import os
import re

import numpy as np
import pytest

P = 0xFFFFFFFF00000001
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# constraints per table: counted from the Rust sources (yield_constr call sites with their loops expanded)
TABLES = {0: "cpu", 1: "memory", 2: "bitwise", 3: "cmp", 4: "rangecheck", 5: "poseidon", 6: "poseidon_chunk", 7: "storage_access", 8: "tape",
          9: "sccall", 10: "program", 11: "prog_chunk"}


def test_the_oracle_includes_nothing_from_the_product():
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".c", ".h", ".cpp", ".hpp")):
            text = open(os.path.join(ROOT, "oracle", f)).read()
            assert not re.search(r'#include\s+"[^"]*olavm_b200', text), f
            assert "ola::air" not in text, f


@pytest.mark.parametrize("tid", sorted(TABLES))
def test_constraint_vectors_agree_on_random_rows(orc, tid):
    import olavm_b200

    cols = orc.table_columns(tid)
    assert cols == olavm_b200.load().ola_table_columns(tid) > 0
    rng = np.random.default_rng(1000 + tid)
    for trial in range(6):
        lv = rng.integers(0, P, size=cols, dtype=np.uint64)
        nv = rng.integers(0, P, size=cols, dtype=np.uint64)
        if trial == 1:   # small values: selectors in {0, 1}, the shapes real rows have
            lv, nv = lv % np.uint64(2), nv % np.uint64(2)
        if trial == 2 and tid == 1:   # the Memory table's value-level branch: next address == ADDR_HEAP_PTR
            nv[3] = np.uint64(18446744060824649731)
        beta = int(rng.integers(0, P, dtype=np.uint64))
        a_vals, a_kinds = orc.air_constraints(tid, lv, nv, beta)
        b_vals, b_kinds = olavm_b200.prover.air_constraints(tid, lv, nv, beta)
        assert len(a_vals) == len(b_vals) > 0, (TABLES[tid], len(a_vals), len(b_vals))
        assert (a_kinds == b_kinds).all(), (TABLES[tid], int(np.argmax(a_kinds != b_kinds)))
        assert (a_vals == b_vals).all(), (TABLES[tid], int(np.argmax(a_vals != b_vals)))
        if trial == 0:
            assert np.count_nonzero(a_vals) > len(a_vals) // 2   # a random point: most constraints are non-zero


def test_constraint_vectors_agree_on_real_rows(orc):
    import olavm_b200
    from workload import fibloop

    ids, traces, cc, _ = fibloop.fib_loop_system(9, orc)
    for tid, t in zip(ids, traces):
        n = t.shape[1]
        rows = sorted(set([0, 1, 2, n // 2, n - 2, n - 1]) | set(int(x) for x in np.random.default_rng(tid).integers(0, n, size=8)))
        for r in rows:
            lv, nv = np.ascontiguousarray(t[:, r]), np.ascontiguousarray(t[:, (r + 1) % n])
            a_vals, a_kinds = orc.air_constraints(tid, lv, nv, cc[tid])
            b_vals, b_kinds = olavm_b200.prover.air_constraints(tid, lv, nv, cc[tid])
            assert (a_kinds == b_kinds).all() and (a_vals == b_vals).all(), (TABLES[tid], r)
            # a satisfying trace: every constraint vanishes except transition ones on the last row / first-, last-row ones elsewhere
            live = (a_kinds == 0) | ((a_kinds == 1) & (r != n - 1)) | ((a_kinds == 2) & (r == 0)) | ((a_kinds == 3) & (r == n - 1))
            assert not a_vals[live].any(), (TABLES[tid], r)


def test_constraint_counts(orc):
    """Per-table constraint counts of the small tables, pinned by hand from the Rust (a dropped or duplicated yield changes
    them): Cmp 4 (cmp_stark.rs:36-44), RangeCheck 1 + 2 lookups x 2 (rangecheck_stark.rs:41-67), SCCall 1, Program 2 + one
    lookup x 2, Tape 15 (tape_stark.rs:59-137)."""
    expect = {3: 4, 4: 5, 9: 1, 10: 4, 8: 15}
    for tid, k in expect.items():
        cols = orc.table_columns(tid)
        z = np.zeros(cols, dtype=np.uint64)
        assert len(orc.air_constraints(tid, z, z, 1)[0]) == k, TABLES[tid]


def test_ctl_registry_agrees(orc):
    """The 19 cross-table lookups (ola_stark.rs:121-560), both transcriptions: proofs of a system that exercises every lookup
    are byte-identical between the product's verifier-side registry and the oracle's only if columns, filters and order agree;
    here the registries are compared through the proofs of the five-table hash system and the fib-loop system."""
    import olavm_b200
    from workload import fibloop

    ids, traces, cc, _ = fibloop.fib_loop_system(4, orc)
    proof = orc.stark_prove(ids, traces, check_degree=True, compress_challenges=cc)
    ok, msg = olavm_b200.verify_proof(ids, proof)   # the product's registry + AIRs (host code) on the oracle's proof
    assert ok, msg
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    assert not olavm_b200.verify_proof(ids, bytes(bad))[0]
