"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs -- bit-exact (integer arithmetic: no tolerance anywhere)."""
import json
import os

import numpy as np
import pytest

import olavm_b200
from olavm_b200 import cfft, hashing, PolynomialBatch

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "poseidon_kat.json")


def bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def bitrev_perm(bits):
    idx = np.arange(1 << bits, dtype=np.uint64)
    out = np.zeros_like(idx)
    for b in range(bits):
        out |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(bits - 1 - b)
    return out.astype(np.int64)


# ------------------------------------------------------------------ Poseidon
def test_poseidon_kats_gpu(ctx):
    g = json.load(open(GOLDEN))
    ins = np.array([v["input"] for v in g["kat"]], dtype=np.uint64)
    outs = np.array([v["output"] for v in g["kat"]], dtype=np.uint64)
    assert (hashing.poseidon(ctx, ins) == outs).all()
    r = g["rounds"]
    for tag in ("ZERO", "1000"):
        got = hashing.poseidon(ctx, np.array(r[f"POSEIDON_{tag}_HASH_INPUT"], dtype=np.uint64))
        assert got.tolist() == r[f"POSEIDON_{tag}_HASH_OUTPUT"]


def test_poseidon_random_vs_oracle(ctx, orc):
    s = orc.rand_elems(11, (4096, 12))
    s[0, :] = P - 1
    s[1, :4] = np.array([P, P + 1, 2**64 - 1, 0], dtype=np.uint64)  # non-canonical representatives
    got = hashing.poseidon(ctx, s)
    for i in list(range(64)) + [4095]:
        assert (got[i] == orc.poseidon(s[i])).all(), i


@pytest.mark.parametrize("ncols", [1, 4, 7, 8, 9, 12, 16, 17, 94, 135])
def test_hash_rows_vs_oracle(ctx, orc, ncols):
    rows = orc.rand_elems(200 + ncols, (777, ncols))
    assert (hashing.hash_no_pad_rows(ctx, rows) == orc.hash_rows(rows)).all()


@pytest.mark.parametrize("log_n,cap_h,width", [(10, 4, 7), (4, 4, 12), (5, 4, 3), (12, 0, 9), (3, 0, 1), (1, 0, 2)])
def test_merkle_rows_vs_oracle(ctx, orc, log_n, cap_h, width):
    n = 1 << log_n
    leaves = orc.rand_elems(log_n * 17 + width, (n, width))
    cap, nodes = hashing.merkle_tree(ctx, leaves, cap_h, want_nodes=True)
    dig, ocap = orc.merkle_new_v2(leaves, cap_h)
    assert (cap == ocap).all()
    # reference prove() == sibling walk of the heap
    for i in {0, 1, n // 2, n - 1}:
        sib = orc.merkle_prove(dig, n, cap_h, i)
        mine = np.array([nodes[((n + i) >> j) ^ 1] for j in range(log_n - cap_h)], dtype=np.uint64).reshape(-1, 4)
        assert (sib == mine).all()
        assert orc.merkle_verify(leaves[i], i, cap, mine)


# ------------------------------------------------------------------ NTT family
@pytest.mark.parametrize("lg", list(range(0, 15)) + [16])
def test_ntt_forward_inverse_vs_oracle(ctx, orc, lg):
    n = 1 << lg
    ncols = 3 if lg < 14 else 1
    c = orc.splitmix64(1, n)[None, :] if lg == 16 else orc.rand_elems(lg + 1, (ncols, n))  # config #1 at lg 16
    v = cfft.evaluate_poly(ctx, c)
    for k in range(c.shape[0]):
        assert (v[k] == orc.evaluate_poly(c[k])).all(), (lg, k)
    back = cfft.interpolate_poly(ctx, v)
    assert (back == c).all()


@pytest.mark.parametrize("lg", [20, 22, 23])
def test_ntt_large_vs_oracle(ctx, orc, lg):
    n = 1 << lg
    c = orc.rand_elems(lg, (1, n))
    v = cfft.evaluate_poly(ctx, c)
    assert (v[0] == orc.evaluate_poly(c[0])).all()
    assert (cfft.interpolate_poly(ctx, v) == c).all()


@pytest.mark.parametrize("lg,rb", [(0, 3), (1, 3), (2, 3), (5, 3), (10, 3), (11, 3), (12, 3), (13, 1), (14, 3), (12, 0), (8, 2)])
def test_coset_lde_vs_oracle(ctx, orc, lg, rb):
    n = 1 << lg
    c = orc.rand_elems(lg * 5 + rb, (2, n))
    nat = cfft.evaluate_poly_with_offset(ctx, c, 7, 1 << rb, natural_order=True)
    leaf = cfft.evaluate_poly_with_offset(ctx, c, 7, 1 << rb, natural_order=False)
    perm = bitrev_perm(lg + rb)
    for k in range(2):
        ref = orc.evaluate_poly_with_offset(c[k], 7, 1 << rb)
        assert (nat[k] == ref).all(), (lg, rb)
        assert (leaf[k] == ref[perm]).all(), (lg, rb)  # leaf r = natural row bitrev(r)  (oracle.rs:84-85)


def test_coset_lde_other_shift(ctx, orc):
    c = orc.rand_elems(3, (1, 1 << 12))
    shift = 7**16 % P  # FRI layer shift after one arity-16 reduction (fri/prover.rs:109)
    got = cfft.evaluate_poly_with_offset(ctx, c, shift, 1)
    assert (got[0] == orc.evaluate_poly_with_offset(c[0], shift, 1)).all()


@pytest.mark.parametrize("lg", [0, 3, 9, 12, 15])
def test_coset_intt_vs_oracle(ctx, orc, lg):
    v = orc.rand_elems(900 + lg, (2, 1 << lg))
    got = cfft.interpolate_poly_with_offset(ctx, v, 7)
    for k in range(2):
        assert (got[k] == orc.interpolate_poly_with_offset(v[k], 7)).all()


def test_cfft_error_behaviour(ctx):
    with pytest.raises(ValueError):
        cfft.evaluate_poly(ctx, np.zeros(12, dtype=np.uint64))  # not a power of two (cfft/mod.rs:26-29)
    with pytest.raises(ValueError):
        cfft.evaluate_poly_with_offset(ctx, np.zeros(8, dtype=np.uint64), 0, 8)  # zero offset (mod.rs:97)
    with pytest.raises(olavm_b200.OlaError):  # beyond two-adicity (mod.rs:37-41): 2^30 * 2^3 > 2^32
        ctx.check(ctx._lib.ola_coset_lde(ctx.handle, 1, 1, 1, 1, 30, 3, 7, 0))


# ------------------------------------------------------------------ PolynomialBatch commit
@pytest.mark.parametrize(
    "lg,ncols,is_coeffs,cap_h", [(4, 1, False, 4), (4, 3, True, 4), (6, 6, False, 4), (9, 29, False, 4), (10, 94, False, 4),
                                 (12, 12, True, 4), (13, 5, False, 0), (1, 2, False, 4), (2, 78, False, 4)]
)
def test_commit_vs_oracle(ctx, orc, lg, ncols, is_coeffs, cap_h):
    vals = orc.rand_elems(lg * 100 + ncols, (ncols, 1 << lg))
    vals[0, 0] = np.uint64(2**64 - 1)  # non-canonical input representative
    ref = orc.commit(vals, is_coeffs=is_coeffs, rate_bits=3, cap_height=cap_h)
    make = PolynomialBatch.from_coeffs if is_coeffs else PolynomialBatch.from_values
    b = make(ctx, vals, 3, False, cap_h)
    assert (b.merkle_cap.hashes == ref["cap"]).all()
    assert (b.polynomials == ref["coeffs"]).all()
    L = 1 << (lg + 3)
    assert (b.leaves() == ref["leaves"]).all()
    for i in {0, 1, L // 3, L - 1}:
        sib = b.prove(i)
        assert (sib == orc.merkle_prove(ref["digests"], L, cap_h, i)).all()
        assert orc.merkle_verify(b.leaves(i, 1)[0], i, ref["cap"], sib)
    # get_lde_values(index, step) bit-reverses index*step (oracle.rs:132-139)
    assert (b.get_lde_values(3 % (L // 8), 8) == ref["leaves"][bitrev(3 % (L // 8) * 8, lg + 3)]).all()
    b.free()


def test_commit_resident_input_matches_host_input(ctx, orc):
    vals = orc.rand_elems(77, (7, 1 << 12))
    d = ctx.upload(vals)
    b1 = PolynomialBatch.from_values(ctx, d, 3, False, 4, on_device=True, ncols=7, degree_log=12)
    b2 = PolynomialBatch.from_values(ctx, vals, 3, False, 4)
    assert (b1.merkle_cap.hashes == b2.merkle_cap.hashes).all()
    assert (ctx.download(d, vals.shape) == vals).all()  # caller's resident buffer is not clobbered
    ctx.free(d)


# ------------------------------------------------------------------ coset shard (multi-GPU commit)
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("lg,ncols", [(5, 3), (12, 20)])
def test_coset_shards_reassemble_the_commitment(ctx, orc, world, lg, ncols):
    """Every rank's shard, computed here one after the other on one GPU: the concatenated cap entries are the full
    cap, each shard's leaves are its slice of the full leaf matrix, and Merkle paths are the global ones."""
    from olavm_b200 import dist as odist

    vals = orc.rand_elems(900 + lg + world, (ncols, 1 << lg))
    ref = orc.commit(vals, is_coeffs=False, rate_bits=3, cap_height=4)
    L = 1 << (lg + 3)
    caps = []
    for rank in range(world):
        lo, hi = odist.coset_range(3, rank, world)
        b = PolynomialBatch._commit(ctx, vals, False, 3, 4, coset_first=lo, coset_count=hi - lo)
        caps.append(b.merkle_cap.hashes)
        assert (b.polynomials == ref["coeffs"]).all()
        per = L // world
        assert (b.leaves() == ref["leaves"][rank * per : (rank + 1) * per]).all()
        for g in {rank * per, rank * per + per // 3, (rank + 1) * per - 1}:
            owner, local = odist.leaf_owner(g, lg, 3, world)
            assert owner == rank
            sib = b.prove(local)
            assert (sib == orc.merkle_prove(ref["digests"], L, 4, g)).all()
            assert orc.merkle_verify(b.leaves(local, 1)[0], g, ref["cap"], sib)
        b.free()
    assert (np.concatenate(caps) == ref["cap"]).all()


def test_coset_shard_argument_errors(ctx, orc):
    vals = orc.rand_elems(5, (2, 16))
    for first, count in ((1, 2), (0, 3), (8, 1), (6, 4)):
        with pytest.raises(olavm_b200.OlaError):
            PolynomialBatch._commit(ctx, vals, False, 3, 4, coset_first=first, coset_count=count)
    with pytest.raises(olavm_b200.OlaError):  # cap coarser than the partition: 1 coset of 8 needs cap_height >= 3
        PolynomialBatch._commit(ctx, vals, False, 3, 2, coset_first=0, coset_count=1)


# ------------------------------------------------------------------ full-size properties (BASELINE configs)
def test_config2_shape_spot_and_linearity(ctx, orc):
    """200 columns x 2^20 rows, blowup 8 (BASELINE config #2 shape, reduced to 40 columns to bound test time):
    spot columns bit-exact vs the oracle, and linearity LDE(a + b) = LDE(a) + LDE(b) over the whole output."""
    lg, ncols = 20, 40
    n = 1 << lg
    vals = orc.rand_elems(2, (ncols, n))
    vals[ncols - 1] = (vals[0].astype(object) + vals[1].astype(object)) % P  # last column = col0 + col1
    d_out = ctx.alloc(ncols * n * 8)
    d_co = ctx.upload(vals)
    lib = ctx._lib
    ctx.check(lib.ola_ntt_inverse(ctx.handle, d_co, 1, ncols, lg))
    ctx.check(lib.ola_coset_lde(ctx.handle, d_co, d_out, 1, ncols, lg, 3, 7, 0))
    out = ctx.download(d_out, (ncols, n * 8))
    for c in (0, 17):
        co = orc.interpolate_poly(vals[c])
        ref = orc.evaluate_poly_with_offset(co, 7, 8)[bitrev_perm(lg + 3)]
        assert (out[c] == ref).all()
    s = out[0].astype(object) + out[1].astype(object)
    assert (np.array(s % P, dtype=np.uint64) == out[ncols - 1]).all()
    for p in (d_out, d_co):
        ctx.free(p)


def test_commit_large_roundtrip(ctx, orc):
    """2^18 rows x 16 columns: cap equals the oracle's, every queried path verifies."""
    vals = orc.rand_elems(4, (16, 1 << 18))
    b = PolynomialBatch.from_values(ctx, vals, 3, False, 4)
    ref = orc.commit(vals, rate_bits=3, cap_height=4, want_leaves=False, want_digests=False)
    assert (b.merkle_cap.hashes == ref["cap"]).all()
    L = 1 << 21
    for i in (0, 12345, L - 1):
        assert orc.merkle_verify(b.leaves(i, 1)[0], i, ref["cap"], b.prove(i))
    b.free()


# ------------------------------------------------------------------ register-tiled passes (ntt_tile.cuh)
@pytest.mark.parametrize("lg", [6, 7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 18, 19, 20, 21, 22])
@pytest.mark.parametrize("ncols", [5, 8])
def test_tiled_lde_leaf_order_many_columns(ctx, orc, lg, ncols):
    """Column groups of the tiled contiguous pass (full and partial groups of 8 / 4 lanes) at every pass split."""
    if lg >= 20 and ncols == 8:
        pytest.skip("covered by ncols=5 at this size")
    n = 1 << lg
    rb = 1 if lg >= 16 else 3
    c = orc.rand_elems(7000 + lg, (ncols, n))
    leaf = cfft.evaluate_poly_with_offset(ctx, c, 7, 1 << rb, natural_order=False)
    perm = bitrev_perm(lg + rb)
    for k in sorted({0, ncols // 2, ncols - 1}):
        ref = orc.evaluate_poly_with_offset(c[k], 7, 1 << rb)
        assert (leaf[k] == ref[perm]).all(), (lg, ncols, k)
    assert (leaf < np.uint64(P)).all()  # canonical outputs


@pytest.mark.parametrize("lg,ncols,is_coeffs", [(10, 3, 0), (17, 40, 0), (17, 33, 1), (13, 5, 0)])
def test_lde_batch_host_pipeline_matches_commit(ctx, orc, lg, ncols, is_coeffs):
    """ola_lde_batch (chunked, overlapped upload at the larger sizes) == the coefficients and leaves of ola_commit."""
    n, rb = 1 << lg, 3
    vals = orc.rand_elems(4242 + lg, (ncols, n))
    vals[0, :4] = [P - 1, 0, 1, P - 2]
    d_c, d_l = ctx.alloc(ncols * n), ctx.alloc(ncols * (n << rb))
    try:
        ctx.check(ctx._lib.ola_lde_batch(ctx.handle, olavm_b200._lib.hptr(vals), 0, ncols, lg, is_coeffs, rb, d_c, d_l))
        coeffs = ctx.download(d_c, (ncols, n))
        lde = ctx.download(d_l, (ncols, n << rb))
        # resident input, transformed in place, gives the same result
        d_v = ctx.upload(vals)
        ctx.check(ctx._lib.ola_lde_batch(ctx.handle, d_v, 1, ncols, lg, is_coeffs, rb, d_v, d_l))
        assert (ctx.download(d_v, (ncols, n)) == coeffs).all()
        assert (ctx.download(d_l, (ncols, n << rb)) == lde).all()
        ctx.free(d_v)
    finally:
        ctx.free(d_c)
        ctx.free(d_l)
    perm = bitrev_perm(lg + rb)
    for k in sorted({0, ncols // 2, ncols - 1}):
        co = vals[k] if is_coeffs else orc.interpolate_poly(vals[k])
        assert (coeffs[k] == co).all(), k
        assert (lde[k] == orc.evaluate_poly_with_offset(co, 7, 1 << rb)[perm]).all(), k
