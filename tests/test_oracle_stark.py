"""CPU tests (-m "not gpu"): the restated STARK prover is pinned by the restated verifier -- the only
acceptance test the reference itself has (circuits/src/stark/ola_stark.rs:690-812 run prove -> verify_proof)."""
import numpy as np
import pytest

import tracegen

P = 0xFFFFFFFF00000001
CMP, RC = 3, 4


@pytest.fixture(scope="module")
def cmp_rc(orc):
    rng = np.random.default_rng(5)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(37, 2))] + [(5, 5), (0, 9)]
    cmp_t = tracegen.cmp_trace(pairs, 6)
    rc_t = tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])
    proof = orc.stark_prove([CMP, RC], [cmp_t, rc_t])
    return cmp_t, rc_t, proof


def test_cmp_rangecheck_proof_verifies(orc, cmp_rc):
    _, _, proof = cmp_rc
    ok, msg = orc.stark_verify([CMP, RC], proof)
    assert ok, msg
    assert len(proof) > 100_000


def test_tampered_proofs_are_rejected(orc, cmp_rc):
    _, _, proof = cmp_rc
    for off in (200, len(proof) // 3, len(proof) // 2, len(proof) - 60):
        bad = bytearray(proof)
        bad[off] ^= 1
        ok, _ = orc.stark_verify([CMP, RC], bytes(bad))
        assert not ok, off


def test_pow_witness_is_smallest_and_only_free_field(orc, cmp_rc):
    # the last 8 bytes of each StarkProof are pow_witness (serialization.rs write_fri_proof); the verifier accepts
    # any valid nonce; our rule: the prover returns the smallest one
    cmp_t, rc_t, proof = cmp_rc
    assert orc.stark_prove([CMP, RC], [cmp_t, rc_t]) == proof  # deterministic


def test_unsatisfied_constraints_are_rejected_by_the_verifier(orc, cmp_rc):
    # Cmp has constraint degree 3 => quotient_degree_factor 2 == 2^qdb: trim_to_len(n*2) of a 2n-coefficient
    # quotient can never fail (prover.rs:463-473), so -- exactly as in the reference -- the prover emits a proof
    # and the verifier's quotient identity rejects it.
    cmp_t, rc_t, _ = cmp_rc
    bad = cmp_t.copy()
    bad[3, 0] = (int(bad[3, 0]) + 1) % P  # abs_diff wrong
    proof = orc.stark_prove([CMP, RC], [bad, rc_t])
    ok, msg = orc.stark_verify([CMP, RC], proof)
    assert not ok and "Mismatch between evaluation and opening of quotient polynomial" in msg


def test_ctl_multiset_mismatch_is_caught(orc, cmp_rc):
    cmp_t, rc_t, _ = cmp_rc
    bad = rc_t.copy()
    bad[3, 0] = 0  # drop one looked-up value from the RangeCheck side: CTL products differ
    proof = orc.stark_prove([CMP, RC], [cmp_t, bad])
    ok, msg = orc.stark_verify([CMP, RC], proof)
    assert not ok and "Cross-table" in msg


def test_non_binary_filter_is_an_error(orc, cmp_rc):
    cmp_t, rc_t, _ = cmp_rc
    bad = cmp_t.copy()
    bad[5, 1] = 2
    with pytest.raises(orc.StarkError, match="Non-binary filter"):
        orc.stark_prove([CMP, RC], [bad, rc_t])


CPU = 0


def test_cpu_padding_trace_satisfies_the_cpu_air(orc):
    """All-padding CPU trace (generate_cpu_trace with no steps) + padding-only Cmp + lookup-free RangeCheck: the
    degree-7 CPU quotient has qdf = 6 < 8, so trim_to_len really checks divisibility (prover.rs:463-473), and the
    restated verify_proof accepts."""
    cpu_t = tracegen.cpu_padding_trace(5)
    cmp_t = tracegen.cmp_trace([], 4)
    rc_t = tracegen.rangecheck_trace([])
    proof = orc.stark_prove([CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    ok, msg = orc.stark_verify([CPU, CMP, RC], proof)
    assert ok, msg
    # breaking one padding invariant (s_end must be 1 on padding rows) makes the quotient non-divisible
    bad = cpu_t.copy()
    bad[74, 3] = 0
    with pytest.raises(orc.StarkError, match="Quotient has failed"):
        orc.stark_prove([CPU, CMP, RC], [bad, cmp_t, rc_t])
