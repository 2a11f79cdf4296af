"""CPU tests (-m "not gpu"): pin the oracle against every golden vector / property the reference's own
tests hold for the hot path."""
import json
import os

import numpy as np
import pytest

P = 0xFFFFFFFF00000001
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "poseidon_kat.json")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN))


def u64(x):
    return np.array(x, dtype=np.uint64)


# ---- Poseidon: reference KATs (poseidon_goldilocks.rs:293-314) and round tables (core/.../poseidon_utils.rs)
def test_poseidon_kats(orc, golden):
    for v in golden["kat"]:
        assert orc.poseidon(u64(v["input"])).tolist() == v["output"]


def test_poseidon_fast_equals_naive(orc, golden):
    # check_consistency (poseidon.rs:714-728): fast partial rounds == naive partial rounds
    for v in golden["kat"]:
        assert orc.poseidon(u64(v["input"]), naive=True).tolist() == v["output"]
    for seed in range(8):
        s = orc.rand_elems(100 + seed, 12)
        assert (orc.poseidon(s) == orc.poseidon(s, naive=True)).all()


def test_poseidon_round_tables(orc, golden):
    r = golden["rounds"]
    for tag in ("ZERO", "1000"):
        out = orc.poseidon(u64(r[f"POSEIDON_{tag}_HASH_INPUT"]))
        assert out.tolist() == r[f"POSEIDON_{tag}_HASH_OUTPUT"]
    # zero-input output equals KAT #1
    assert r["POSEIDON_ZERO_HASH_OUTPUT"] == golden["kat"][0]["output"]


def test_poseidon_table_row_golden(orc, golden):
    # the Poseidon TABLE's witness row (S-box inputs per round) against the reference's per-round tables
    # (core/src/util/poseidon_utils.rs:11-287, used by generation/poseidon.rs:83-126 as the padding row)
    r = golden["rounds"]
    for tag in ("ZERO", "1000"):
        row = orc.poseidon_table_row(u64(r[f"POSEIDON_{tag}_HASH_INPUT"])).tolist()
        assert row[0:4] == [0, 0, 0, 0]
        assert row[4:16] == r[f"POSEIDON_{tag}_HASH_INPUT"]
        assert row[16:28] == r[f"POSEIDON_{tag}_HASH_OUTPUT"]
        assert row[28:40] == r[f"POSEIDON_{tag}_HASH_FULL_0_1"]
        assert row[40:52] == r[f"POSEIDON_{tag}_HASH_FULL_0_2"]
        assert row[52:64] == r[f"POSEIDON_{tag}_HASH_FULL_0_3"]
        assert row[64:86] == r[f"POSEIDON_{tag}_HASH_PARTIAL"]
        assert row[86:98] == r[f"POSEIDON_{tag}_HASH_FULL_1_0"]
        assert row[98:110] == r[f"POSEIDON_{tag}_HASH_FULL_1_1"]
        assert row[110:122] == r[f"POSEIDON_{tag}_HASH_FULL_1_2"]
        assert row[122:134] == r[f"POSEIDON_{tag}_HASH_FULL_1_3"]


def test_noncanonical_inputs_are_the_same_element(orc):
    s = orc.rand_elems(7, 12)
    s[:3] = u64([0, 1, 2])
    t = s.copy()
    t[:3] += np.uint64(P)  # 0+p, 1+p, 2+p are valid non-canonical u64 representatives
    assert (orc.poseidon(s) == orc.poseidon(t)).all()


# ---- field constants (goldilocks_field.rs:70-77, goldilocks_extensions.rs:27)
def test_generators(golden):
    c = golden["constants"]
    g = c["power_of_two_generator"]
    assert pow(7, (P - 1) >> 32, P) == g
    assert pow(g, 1 << 32, P) == 1 and pow(g, 1 << 31, P) == P - 1
    # ext generator [0, e]: (e X)^2 = 7 e^2 must equal the base 2^32-order generator
    e = c["ext_power_of_two_generator"][1]
    assert 7 * e * e % P == g


# ---- FFT: the reference checks coset FFT against naive evaluation (polynomial/mod.rs:495-540, fft.rs:219-251)
@pytest.mark.parametrize("lg", [0, 1, 2, 3, 8, 10])
def test_fft_matches_naive(orc, lg):
    n = 1 << lg
    c = orc.rand_elems(lg, n)
    v = orc.evaluate_poly(c)
    g = pow(1753635133440165772, 1 << (32 - lg), P)
    for k in {0, 1 % n, n // 2, n - 1}:
        assert int(v[k]) == orc.poly_eval(c, pow(g, k, P))
    assert (orc.fft_classic(c) == v).all()  # legacy plonky2 fft.rs path agrees with cfft
    assert (orc.interpolate_poly(v) == c).all()


@pytest.mark.parametrize("lg,rb", [(0, 3), (1, 3), (4, 3), (8, 3), (8, 0), (6, 1)])
def test_coset_fft_matches_naive(orc, lg, rb):
    n = 1 << lg
    c = orc.rand_elems(50 + lg, n)
    l = orc.evaluate_poly_with_offset(c, 7, 1 << rb)
    g = pow(1753635133440165772, 1 << (32 - lg - rb), P)
    L = n << rb
    for k in {0, 1 % L, L // 2, L - 1, (3 * L) // 4}:
        assert int(l[k]) == orc.poly_eval(c, 7 * pow(g, k, P) % P)
    # coset_ifft(coset_fft(p)) == p   (polynomial/mod.rs test_coset_fft/test_coset_ifft)
    if rb == 0:
        assert (orc.interpolate_poly_with_offset(l, 7) == c).all()


def test_ntt_config1_size(orc):
    # BASELINE config #1: 2^16 single column, splitmix64(seed=1)
    c = orc.splitmix64(1, 1 << 16)
    v = orc.evaluate_poly(c)
    assert (orc.fft_classic(c) == v).all()
    assert (orc.interpolate_poly(v) == c).all()


# ---- Merkle: build over random leaves and verify every opening (merkle_tree/mod.rs:352-409)
@pytest.mark.parametrize("log_n,cap_h,width", [(8, 0, 7), (8, 1, 7), (8, 4, 7), (4, 4, 12), (5, 4, 3), (3, 0, 1)])
def test_merkle_all_openings(orc, log_n, cap_h, width):
    n = 1 << log_n
    leaves = orc.rand_elems(log_n * 31 + cap_h, (n, width))
    dig, cap = orc.merkle_new_v2(leaves, cap_h)
    for i in range(n):
        sib = orc.merkle_prove(dig, n, cap_h, i)
        assert sib.shape[0] == log_n - cap_h
        assert orc.merkle_verify(leaves[i], i, cap, sib)
    bad = leaves[0].copy()
    bad[0] = (int(bad[0]) + 1) % P
    assert not orc.merkle_verify(bad, 0, cap, orc.merkle_prove(dig, n, cap_h, 0))


def test_commit_structure(orc):
    # PolynomialBatch::from_values: leaves[r] = LDE row bitrev(r); coefficients interpolate the values
    vals = orc.rand_elems(5, (5, 16))
    b = orc.commit(vals, is_coeffs=False, rate_bits=3, cap_height=2)
    for c in range(5):
        assert (orc.evaluate_poly(b["coeffs"][c]) == vals[c]).all()
        lde = orc.evaluate_poly_with_offset(b["coeffs"][c], 7, 8)
        for r in (0, 1, 77, 127):
            assert int(b["leaves"][r, c]) == int(lde[int(format(r, "07b")[::-1], 2)])
    for r in (0, 5, 127):
        sib = orc.merkle_prove(b["digests"], 128, 2, r)
        assert orc.merkle_verify(b["leaves"][r], r, b["cap"], sib)
