"""CPU tests (-m "not gpu") of the product-side verifier `ola_verify` (olavm_b200/csrc/verify.h, host code inside
libola_gpu.so; mirrors circuits/src/stark/verifier.rs + plonky2 fri/verifier.rs).  It is an implementation independent
of the oracle's verifier: both must accept the oracle prover's proofs and reject the same broken ones.  (The GPU suite
feeds it the GPU prover's proofs.)"""
import numpy as np
import pytest

import olavm_b200
import tracegen
from test_oracle_stark import BREAK, SINGLE, _valid_single

P = 0xFFFFFFFF00000001
CMP, RC = 3, 4


@pytest.fixture(scope="module")
def cmp_rc_proof(orc):
    rng = np.random.default_rng(5)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(37, 2))] + [(5, 5), (0, 9)]
    cmp_t = tracegen.cmp_trace(pairs, 6)
    rc_t = tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])
    return cmp_t, rc_t, orc.stark_prove([CMP, RC], [cmp_t, rc_t])


def test_accepts_valid_proof_and_agrees_with_oracle_verifier(orc, cmp_rc_proof):
    _, _, proof = cmp_rc_proof
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], proof)
    assert ok, msg
    assert orc.stark_verify([CMP, RC], proof)[0]


def test_rejects_every_kind_of_tampering(orc, cmp_rc_proof):
    _, _, proof = cmp_rc_proof
    rng = np.random.default_rng(1)
    offsets = [4, 200, len(proof) // 3, len(proof) // 2, len(proof) - 60] + [int(x) for x in rng.integers(0, len(proof), size=40)]
    for off in offsets:
        bad = bytearray(proof)
        bad[off] ^= 1
        ok, _ = olavm_b200.verify_subsystem_proof([CMP, RC], bytes(bad))
        ok_ref, _ = orc.stark_verify([CMP, RC], bytes(bad))
        assert ok == ok_ref, off  # both verifiers decide identically (a flipped pow_witness bit may stay valid)
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], proof[:-1])[0]
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], proof + b"\0")[0]
    assert not olavm_b200.verify_subsystem_proof([CMP], proof)[0]
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], b"")[0]


def test_rejects_unsatisfied_constraints_and_ctl_mismatch(orc, cmp_rc_proof):
    cmp_t, rc_t, _ = cmp_rc_proof
    bad = cmp_t.copy()
    bad[3, 0] = (int(bad[3, 0]) + 1) % P
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], orc.stark_prove([CMP, RC], [bad, rc_t]))
    assert not ok and "Mismatch between evaluation and opening of quotient polynomial" in msg
    bad = rc_t.copy()
    bad[3, 0] = 0
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], orc.stark_prove([CMP, RC], [cmp_t, bad]))
    assert not ok and "Cross-table lookup" in msg


@pytest.mark.parametrize("name", SINGLE)
def test_accepts_valid_trace_of_each_table(orc, name):
    ids, traces, cc = _valid_single(orc, name)
    proof = orc.stark_prove(ids, traces, True, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, proof)
    assert ok, msg


@pytest.mark.parametrize("name", ["tape", "bitwise", "poseidon", "program"])
def test_rejects_broken_trace(orc, name):
    ids, traces, cc = _valid_single(orc, name)
    col, row, _ = BREAK[name]
    bad = traces[0].copy()
    bad[col, row] = (int(bad[col, row]) + 1) % P
    proof = orc.stark_prove(ids, [bad], False, compress_challenges=cc)  # degree check off: the prover emits a proof anyway
    ok, msg = olavm_b200.verify_subsystem_proof(ids, proof)
    assert not ok and not orc.stark_verify(ids, proof)[0]


def test_cpu_table_and_five_table_system(orc):
    cpu_t = tracegen.cpu_padding_trace(5)
    cmp_t = tracegen.cmp_trace([], 4)
    rc_t = tracegen.rangecheck_trace([])
    proof = orc.stark_prove([0, CMP, RC], [cpu_t, cmp_t, rc_t])
    ok, msg = olavm_b200.verify_subsystem_proof([0, CMP, RC], proof)
    assert ok, msg
    rng = np.random.default_rng(3)
    ids, traces, cc = tracegen.hash_system_valid(orc, rng)
    proof = orc.stark_prove(ids, traces, True, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, proof)
    assert ok, msg
    tampered = bytearray(proof)
    tampered[-16] ^= 1  # the Program table's compress challenge travels in the proof (verifier.rs:83-86)
    assert not olavm_b200.verify_subsystem_proof(ids, bytes(tampered))[0]
    bad = [t.copy() for t in traces]
    bad[3][17, 2] = 0
    ok, msg = olavm_b200.verify_subsystem_proof(ids, orc.stark_prove(ids, bad, True, compress_challenges=cc))
    assert not ok and "Cross-table lookup" in msg


def test_verify_proof_is_fixed_at_the_full_system(orc, cmp_rc_proof):
    """verify_proof (verifier.rs:32-212) always verifies the 12 tables and every lookup between them: the drop-in entry point
    refuses a subset (which would silently skip the lookups that lost a side); the subsystem entry point is explicit."""
    _, _, proof = cmp_rc_proof
    ok, msg = olavm_b200.verify_proof([CMP, RC], proof)
    assert not ok and "12-table" in msg
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], proof)
    assert ok, msg
    assert not olavm_b200.verify_subsystem_proof([RC, CMP], proof)[0]   # ids in enum order only
    assert not olavm_b200.verify_subsystem_proof([CMP, CMP], proof)[0]
