"""BLAKE3 restatement (oracle/blake3.c) pinned to the official implementation's vectors, and the modes the reference's
Blake3GoldilocksConfig builds on it (plonky2/plonky2/src/hash/blake3.rs:165-233, hash_types.rs:142-152)."""
import json
import os
import struct

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blake3_kat.json")
P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def kat():
    with open(GOLD) as f:
        return json.load(f)


def pattern(n):
    return bytes(i % 251 for i in range(n))


def test_blake3_known_answers(orc, kat):
    assert orc.blake3(b"").hex() == kat["known"]["empty"]
    assert orc.blake3(b"abc").hex() == kat["known"]["abc"]


def test_blake3_matches_every_official_vector(orc, kat):
    assert len(kat["vectors"]) >= 40
    for v in kat["vectors"]:
        assert orc.blake3(pattern(v["len"])).hex() == v["hash"], v["len"]


def test_hash_no_pad_is_blake3_of_the_le_u64_image(orc, kat):
    by_len = {v["len"]: v["hash"] for v in kat["vectors"]}
    # rows whose little-endian image is the official test pattern: only lengths whose words are canonical qualify
    for ncols in (1, 4, 12, 16, 94, 128, 134, 256):
        words = np.frombuffer(pattern(8 * ncols), dtype="<u8").astype(np.uint64)
        assert (words < P).all()
        with orc.hasher(orc.BLAKE3):
            h = orc.hash_no_pad(words)
        assert h.astype("<u8").tobytes().hex() == by_len[8 * ncols], ncols
    # non-canonical input words are hashed as their canonical representatives (documented deviation, oracle/blake3.c)
    w = np.array([P + 5, 7], dtype=np.uint64)
    with orc.hasher(orc.BLAKE3):
        assert (orc.hash_no_pad(w) == orc.hash_no_pad(np.array([5, 7], dtype=np.uint64))).all()


def test_two_to_one_is_blake3_of_the_concatenation(orc):
    l = orc.rand_elems(1, (4,))
    r = np.array([P, 2**64 - 1, 0, 1], dtype=np.uint64)  # digest words are bytes, not field elements: never reduced
    with orc.hasher(orc.BLAKE3):
        h = orc.two_to_one(l, r)
    assert h.astype("<u8").tobytes() == orc.blake3(l.astype("<u8").tobytes() + r.astype("<u8").tobytes())


def test_permutation_is_the_hash_onion_with_rejection_sampling(orc):
    for seed in range(20):
        st = orc.rand_elems(100 + seed, (12,))
        got = orc.blake3_permute(st)
        buf = st.astype("<u8").tobytes()
        want = []
        while len(want) < 12:
            buf = orc.blake3(buf)
            want += [w for w in struct.unpack("<4Q", buf) if w < P]
        assert [int(x) for x in got] == want[:12]
        assert all(int(x) < P for x in got)


def test_bytes_hash_to_fields_uses_seven_byte_chunks(orc):
    h = np.frombuffer(bytes(range(1, 33)), dtype="<u8").astype(np.uint64)
    f = orc.bytes_hash_to_fields(h)
    raw = bytes(range(1, 33))
    want = [int.from_bytes(raw[7 * c:7 * c + 7], "little") for c in range(5)]
    assert [int(x) for x in f] == want
    assert int(f[4]) < 2**32  # the last chunk holds 4 bytes


def test_merkle_tree_under_blake3_opens_and_verifies(orc):
    rows = orc.rand_elems(9, (64, 20))
    with orc.hasher(orc.BLAKE3):
        digests, cap = orc.merkle_new_v2(rows, 2)
        for i in (0, 17, 63):
            sib = orc.merkle_prove(digests, 64, 2, i)
            assert orc.merkle_verify(rows[i], i, cap, sib)
            bad = rows[i].copy()
            bad[3] ^= 1
            assert not orc.merkle_verify(bad, i, cap, sib)
    # the Poseidon tree over the same rows has a different cap, and the default hasher is restored
    _, cap_p = orc.merkle_new_v2(rows, 2)
    assert not (cap_p == cap).all()


# ---- the STARK prover / verifier under Blake3GoldilocksConfig (the config of the reference's criterion benches,
# circuits/benches/fibo_loop.rs:26, and of its integration tests, circuits/src/stark/ola_stark.rs:684) ----
CMP, RC = 3, 4


@pytest.fixture(scope="module")
def cmp_rc_blake3(orc):
    import tracegen

    rng = np.random.default_rng(5)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(37, 2))] + [(5, 5), (0, 9)]
    cmp_t = tracegen.cmp_trace(pairs, 6)
    rc_t = tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])
    proof = orc.stark_prove([CMP, RC], [cmp_t, rc_t], hasher_id=orc.BLAKE3)
    return cmp_t, rc_t, proof


def test_blake3_proof_verifies_only_under_blake3(orc, cmp_rc_blake3):
    cmp_t, rc_t, proof = cmp_rc_blake3
    ok, msg = orc.stark_verify([CMP, RC], proof, hasher_id=orc.BLAKE3)
    assert ok, msg
    ok, _ = orc.stark_verify([CMP, RC], proof)  # PoseidonGoldilocksConfig verifier: other transcript, other trees
    assert not ok
    poseidon_proof = orc.stark_prove([CMP, RC], [cmp_t, rc_t])
    assert len(poseidon_proof) == len(proof) and poseidon_proof != proof  # same shape: 32-byte hashes either way
    ok, _ = orc.stark_verify([CMP, RC], poseidon_proof, hasher_id=orc.BLAKE3)
    assert not ok


def test_blake3_tampered_proofs_are_rejected(orc, cmp_rc_blake3):
    _, _, proof = cmp_rc_blake3
    for off in (4 + 5, 200, len(proof) // 3, len(proof) // 2, len(proof) - 60):
        bad = bytearray(proof)
        bad[off] ^= 1
        ok, _ = orc.stark_verify([CMP, RC], bytes(bad), hasher_id=orc.BLAKE3)
        assert not ok, off


def test_blake3_cap_words_are_raw_bytes_on_the_wire(orc, cmp_rc_blake3):
    # write_hash = to_bytes (serialization.rs:115-117): a BytesHash is its 32 bytes, words >= p included; with 3 caps of
    # 16 hashes per table the odds that no word of the first cap has its top 32 bits set are nil -- but it may happen
    # that none is >= p, so only the round trip is asserted: the first cap follows the u32 proof count and u32 cap length
    _, _, proof = cmp_rc_blake3
    assert struct.unpack_from("<II", proof, 0) == (2, 16)
    assert orc.stark_prove([CMP, RC], [cmp_rc_blake3[0], cmp_rc_blake3[1]], hasher_id=orc.BLAKE3) == proof  # deterministic
