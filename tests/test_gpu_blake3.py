"""GPU parity tests (-m gpu) of the Blake3GoldilocksConfig path (ola_set_hasher(OLA_HASH_BLAKE3)): leaf / node / FRI-leaf
BLAKE3 kernels, the commitment and the whole prover must agree with the oracle (oracle/blake3.c, pinned to the official
implementation's vectors) bit for bit -- digests, caps, Merkle paths and proof bytes."""
import numpy as np
import pytest

import olavm_b200
import tracegen
from olavm_b200 import hashing

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001
CMP, RC = 3, 4
B3 = olavm_b200.BLAKE3


@pytest.fixture()
def b3ctx(ctx):
    ctx.hasher = B3
    assert ctx.hasher == B3
    yield ctx
    ctx.hasher = olavm_b200.POSEIDON


@pytest.mark.parametrize("ncols", [1, 3, 8, 9, 16, 94, 127, 128, 129, 134, 256, 257, 300, 1000, 2048])
def test_leaf_digests_equal_oracle(b3ctx, orc, ncols):
    # one block, partial blocks, one full chunk (128), chunk + 1 element, 2..16 chunks (parent nodes, unbalanced trees)
    rows = orc.rand_elems(1000 + ncols, (70, ncols))
    rows[0, 0] = np.uint64(P + 3)  # a non-canonical input word is hashed as its canonical representative
    got = hashing.hash_no_pad_rows(b3ctx, rows)
    with orc.hasher(orc.BLAKE3):
        ref = orc.hash_rows(rows)
    assert (got == ref).all()


def test_leaf_wider_than_sixteen_chunks_is_an_error(b3ctx, orc):
    with pytest.raises(olavm_b200.OlaError, match="2048"):
        hashing.hash_no_pad_rows(b3ctx, orc.rand_elems(1, (4, 2049)))


def test_merkle_tree_nodes_and_cap_equal_oracle(b3ctx, orc):
    rows = orc.rand_elems(77, (256, 20))
    cap, nodes = hashing.merkle_tree(b3ctx, rows, 3, want_nodes=True)
    with orc.hasher(orc.BLAKE3):
        digests, cap_ref = orc.merkle_new_v2(rows, 3)
        assert (cap == cap_ref).all()
        # heap-ordered nodes: the sibling walk of leaf i is the reference's Merkle path
        for i in (0, 1, 100, 255):
            sib = np.array([nodes[((256 + i) >> j) ^ 1] for j in range(8 - 3)], dtype=np.uint64)
            assert orc.merkle_verify(rows[i], i, cap_ref, sib)
            assert (sib == orc.merkle_prove(digests, 256, 3, i)).all()


@pytest.mark.parametrize("ncols,log_n", [(12, 10), (94, 8), (134, 6), (94, 14)])
def test_commitment_equals_oracle(b3ctx, orc, ncols, log_n):
    vals = orc.rand_elems(5 + ncols, (ncols, 1 << log_n))
    batch = olavm_b200.PolynomialBatch.from_values(b3ctx, vals, 3, False, 4)
    with orc.hasher(orc.BLAKE3):
        ref = orc.commit(vals, rate_bits=3, cap_height=4)
        assert (batch.merkle_cap.hashes == ref["cap"]).all()
        leaf = (1 << log_n) + 5
        assert orc.merkle_verify(batch.leaves(leaf, 1)[0], leaf, ref["cap"], batch.prove(leaf))
    assert (batch.polynomials == ref["coeffs"]).all()
    batch.free()


def _valid_cmp_rc(seed, log_cmp):
    rng = np.random.default_rng(seed)
    k = (1 << log_cmp) - 5
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(k, 2))] + [(5, 5), (0, 9)]
    return tracegen.cmp_trace(pairs, log_cmp), tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])


@pytest.mark.parametrize("seed,log_cmp", [(5, 6), (7, 9)])
def test_proof_bytes_equal_oracle_and_verify(b3ctx, orc, seed, log_cmp):
    cmp_t, rc_t = _valid_cmp_rc(seed, log_cmp)
    ref = orc.stark_prove([CMP, RC], [cmp_t, rc_t], hasher_id=orc.BLAKE3)
    got = olavm_b200.prove_with_traces(b3ctx, [CMP, RC], [cmp_t, rc_t])
    assert got == ref
    ok, msg = orc.stark_verify([CMP, RC], got, hasher_id=orc.BLAKE3)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], got, hasher=B3)
    assert ok, msg
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], got)[0]  # not a PoseidonGoldilocksConfig proof


def test_hasher_switch_leaves_the_poseidon_path_unchanged(ctx, orc):
    cmp_t, rc_t = _valid_cmp_rc(3, 5)
    ref = orc.stark_prove([CMP, RC], [cmp_t, rc_t])
    ctx.hasher = B3
    b3 = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])
    ctx.hasher = olavm_b200.POSEIDON
    assert olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t]) == ref
    assert b3 != ref and len(b3) == len(ref)
    with pytest.raises(olavm_b200.OlaError):
        ctx.hasher = 9


def test_wide_tables_and_pipeline_parity(b3ctx, orc):
    # Cpu (94 columns: 12 blocks per leaf, 78 Z columns) and Poseidon (134 columns: two chunks + a parent per leaf) on
    # random columns with binary filters, quotient-degree check off: every kernel of the prover runs under the BLAKE3
    # commitment; bytes must equal the oracle's
    rng = np.random.default_rng(11)
    ids = [0, 5]
    traces = [tracegen.cpu_random_trace(rng, 6), tracegen.poseidon_random_trace(rng, 5)]
    ref = orc.stark_prove(ids, traces, check_degree=False, hasher_id=orc.BLAKE3)
    got = olavm_b200.prove_with_traces(b3ctx, ids, traces, check_quotient_degree=False)
    assert got == ref


def test_sharded_prover_under_blake3_equals_single_gpu(b3ctx, orc):
    from olavm_b200 import dist

    cmp_t, rc_t = _valid_cmp_rc(9, 7)
    single = olavm_b200.prove_with_traces(b3ctx, [CMP, RC], [cmp_t, rc_t])
    for world in (2, 8):
        outs = dist.prove_sharded_local(0, world, [CMP, RC], [cmp_t, rc_t], hasher=B3)
        assert all(o == single for o in outs), world
