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
