"""workload/fibloop.py (the lane-parallel generator behind BASELINE configs[2]) against its spec, workload/tracegen.py
(the step-by-step VM restating the reference executor + circuits/src/generation): same tables, every AIR satisfied,
restated prover -> restated verifier accepts with the quotient-degree check on."""
import numpy as np
import pytest

from workload import fibloop as fl
from workload import tracegen as tg

P = tg.P


def _lookup_ok(perm_in, perm_tab):
    """eval_lookups (circuits/src/stark/lookup.rs:13-35) over a whole column pair, cyclic."""
    nxt_in, nxt_tab = np.roll(perm_in, -1), np.roll(perm_tab, -1)
    return bool((((nxt_in == perm_in) | (nxt_in == nxt_tab))).all()) and perm_in[0] == perm_tab[0]


@pytest.mark.parametrize("n", [2, 3, 9, 40])
def test_lane_parallel_run_equals_the_step_by_step_vm(orc, n):
    prog, _ = fl.fib_loop_program(n)
    lg = max(4, (fl.steps_of(n) - 1).bit_length())
    cpu_ref, steps, cmp_pairs, rc_cmp, rc_cpu, mlog, bit_ops = tg.cpu_vm_trace(prog, lg, want_side_tables="all")
    assert len(steps) == fl.steps_of(n)
    _, cpu_t, nsteps, mem, cmp_, bit, exe = fl.cpu_and_logs(n, lg)
    assert nsteps == len(steps) and (cpu_t == cpu_ref).all()
    # Memory / Cmp: identical tables and identical range-check values
    mem_log = max(2, len(mlog).bit_length())
    mem_ref, rc_sort_ref = tg.memory_trace_from_log(mlog, mem_log)
    mem_t, rc_sort = fl.memory_trace_vec(*mem, mem_log)
    assert (mem_t == mem_ref).all() and list(rc_sort) == list(rc_sort_ref)
    cmp_log = max(4, len(cmp_pairs).bit_length())
    cmp_t, rc_c = fl.cmp_trace_vec(*cmp_, cmp_log)
    assert (cmp_t == tg.cmp_trace(cmp_pairs, cmp_log)).all() and list(rc_c) == list(rc_cmp) and not rc_cpu
    # RangeCheck / Bitwise / Program: identical except for the permuted TABLE columns, whose unused entries may be handed
    # out in any order (same multiset, lookup property holds)
    rng = np.random.default_rng(0)
    beta, beta_b = 0x1234567890ABCDEF % P, 0x0FEDCBA987654321 % P
    rc_ref = tg.rangecheck_trace(rc_cmp, mem_sort_vals=rc_sort_ref)
    rc_t = fl.rangecheck_trace_vec(rc_c, 16, mem_sort_vals=rc_sort)
    bw_ref = tg.bitwise_valid_trace(rng, 9, beta_b, ops=bit_ops)
    bw_t = fl.bitwise_trace_vec(*bit, beta_b, 9)
    prog_rows, exec_rows = tg.program_rows_of_run(prog, steps)
    pr_log = max(2, max(len(prog_rows), len(exec_rows)).bit_length())
    pr_ref = tg.program_valid_trace(rng, pr_log, beta, prog_rows=prog_rows, exec_rows=exec_rows)
    pr_t = fl.program_trace_vec([r[5] for r in prog_rows], exe[0], exe[1], beta, pr_log)
    for got, ref, pairs in ((rc_t, rc_ref, [(7, 10), (8, 11)]),
                            (bw_t, bw_ref, [(17 + j, 38 + j) for j in range(12)] + [(33 + j, 55 + j) for j in range(4)]),
                            (pr_t, pr_ref, [(15, 7)])):
        free = [tab for _, tab in pairs]
        keep = [c for c in range(ref.shape[0]) if c not in free]
        assert (got[keep] == ref[keep]).all()
        for inp, tab in pairs:
            assert (np.sort(got[tab]) == np.sort(ref[tab])).all()
            assert _lookup_ok(got[inp], got[tab]) and _lookup_ok(ref[inp], ref[tab])


def test_permuted_cols_vec_property():
    rng = np.random.default_rng(3)
    for _ in range(20):
        n = 1 << int(rng.integers(2, 9))
        table = rng.integers(0, 12, size=n).astype(np.uint64)
        inputs = rng.choice(table, size=n).astype(np.uint64)
        si, perm = fl.permuted_cols_vec(inputs, table)
        assert (si == np.sort(inputs)).all() and (np.sort(perm) == np.sort(table)).all() and _lookup_ok(si, perm)
        ref_si, ref_perm = tg.permuted_cols(inputs, table)
        assert (ref_si == si).all() and (np.sort(ref_perm) == np.sort(perm)).all()


def test_field_helpers():
    rng = np.random.default_rng(5)
    a = rng.integers(0, P, size=500, dtype=np.uint64)
    b = rng.integers(0, P, size=500, dtype=np.uint64)
    a[:3], b[:3] = [P - 1, 0, P - 1], [P - 1, 0, 1]
    assert [int(x) for x in fl.addp(a, b)] == [(int(x) + int(y)) % P for x, y in zip(a, b)]
    assert [int(x) for x in fl.subp(a, b)] == [(int(x) - int(y)) % P for x, y in zip(a, b)]
    assert [int(x) for x in fl.inv_many(a)] == [pow(int(x), P - 2, P) if x else 0 for x in a]


def test_every_table_of_the_system_satisfies_its_air(orc):
    n = fl.bound_for_rows(14)
    ids, traces, cc, info = fl.fib_loop_system(n, orc, log_n_cpu=14)
    assert ids == list(range(12)) and info["table_log_n"][0] == 14
    for i, t in zip(ids, traces):
        assert t.shape[0] == orc.table_columns(i)
        assert orc.air_first_failure(i, t, cc[i]) is None, f"table {i}"


def test_fib_loop_system_proves_and_verifies(orc):
    ids, traces, cc, _ = fl.fib_loop_system(5, orc)
    proof = orc.stark_prove(ids, traces, check_degree=True, compress_challenges=cc)
    ok, msg = orc.stark_verify(ids, proof)
    assert ok, msg
    # a broken fib value is caught
    bad = [t.copy() for t in traces]
    row = fl.PROLOGUE_STEPS + 9   # `add r0 r1 r2` of the first iteration: dst = a + b
    assert bad[0][32, row] == 1
    bad[0][32, row] = 2
    with pytest.raises(orc.StarkError):
        orc.stark_prove(ids, bad, check_degree=True, compress_challenges=cc)
