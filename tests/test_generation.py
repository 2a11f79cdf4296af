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
