"""CPU tests (-m "not gpu") of the product-side verifier under Blake3GoldilocksConfig (`ola_verify_cfg`, host code of
libola_gpu.so: olavm_b200/csrc/verify.h + blake3.cuh).  The product's BLAKE3 (the same functions the leaf / node / FRI
kernels compile for the device) and its transcript are an implementation independent of oracle/blake3.c: both verifiers
must accept the oracle prover's BLAKE3 proofs and decide identically on tampered ones."""
import numpy as np
import pytest

import olavm_b200
import tracegen
from test_oracle_stark import _valid_single

CMP, RC = 3, 4
B3 = olavm_b200.BLAKE3


@pytest.fixture(scope="module")
def cmp_rc_b3(orc):
    rng = np.random.default_rng(5)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(37, 2))] + [(5, 5), (0, 9)]
    cmp_t = tracegen.cmp_trace(pairs, 6)
    rc_t = tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])
    return cmp_t, rc_t, orc.stark_prove([CMP, RC], [cmp_t, rc_t], hasher_id=orc.BLAKE3)


def test_accepts_blake3_proof_and_only_under_blake3(orc, cmp_rc_b3):
    cmp_t, rc_t, proof = cmp_rc_b3
    ok, msg = olavm_b200.verify_subsystem_proof([CMP, RC], proof, hasher=B3)
    assert ok, msg
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], proof)[0]  # Poseidon verifier
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], orc.stark_prove([CMP, RC], [cmp_t, rc_t]), hasher=B3)[0]
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], proof, hasher=7)[0]  # unknown hasher id


def test_decides_like_the_oracle_verifier_on_tampered_blake3_proofs(orc, cmp_rc_b3):
    _, _, proof = cmp_rc_b3
    rng = np.random.default_rng(2)
    offsets = [4, 8 + 3, 200, len(proof) // 3, len(proof) // 2, len(proof) - 60] + [int(x) for x in rng.integers(0, len(proof), size=40)]
    rejected = 0
    for off in offsets:
        bad = bytearray(proof)
        bad[off] ^= 1
        ok, _ = olavm_b200.verify_subsystem_proof([CMP, RC], bytes(bad), hasher=B3)
        ok_ref, _ = orc.stark_verify([CMP, RC], bytes(bad), hasher_id=orc.BLAKE3)
        assert ok == ok_ref, off
        rejected += not ok
    assert rejected >= len(offsets) - 2
    assert not olavm_b200.verify_subsystem_proof([CMP, RC], proof[:-1], hasher=B3)[0]


@pytest.mark.parametrize("name", ["poseidon", "tape"])
def test_accepts_valid_trace_of_a_wide_and_a_narrow_table(orc, name):
    # the Poseidon table's rows are 134 columns = 1072 bytes: two BLAKE3 chunks and a parent node per leaf
    ids, traces, cc = _valid_single(orc, name)
    proof = orc.stark_prove(ids, traces, True, compress_challenges=cc, hasher_id=orc.BLAKE3)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, proof, hasher=B3)
    assert ok, msg
    assert orc.stark_verify(ids, proof, hasher_id=orc.BLAKE3)[0]


def test_accepts_the_reference_storage_program_under_blake3(orc):
    """The reference's own sstore / sload program (seven tables of one run, the 134-column Poseidon table among them: two
    BLAKE3 chunks and a parent per leaf) proven by the oracle under Blake3GoldilocksConfig: the product verifier accepts."""
    from test_oracle_stark import _reference_run

    ids, traces, cc, _ = _reference_run(orc, "storage")
    proof = orc.stark_prove(ids, traces, compress_challenges=cc, hasher_id=orc.BLAKE3)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, proof, hasher=B3)
    assert ok, msg
    assert not olavm_b200.verify_subsystem_proof(ids, proof)[0]
