/* ORACLE (test infrastructure, NOT product code) -- see oracle/gl.h header.
 *
 * CPU restatement of Poseidon-Goldilocks (width 12, x^7, 4+22+4 rounds) and of the sponge /
 * compression modes built on it.
 *
 * Follows:
 *   plonky2/plonky2/src/hash/poseidon.rs   constant_layer :476-487, sbox_monomial :521-527,
 *                                          mds_row_shf :168-189, mds_layer :236-257,
 *                                          partial_first_constant_layer :303-313,
 *                                          mds_partial_layer_init :332-358,
 *                                          mds_partial_layer_fast :392-421, full_rounds :566-574,
 *                                          partial_rounds :577-590, poseidon :593-604,
 *                                          partial_rounds_naive / poseidon_naive :607-628
 *   plonky2/plonky2/src/hash/hashing.rs    compress :66-74, hash_n_to_m_no_pad :84-106 (rate 8, overwrite mode)
 *   plonky2/plonky2/src/hash/poseidon_goldilocks.rs   KATs :293-314 (checked in tests/test_oracle.py)
 */
#include "oracle.h"
#include "poseidon_constants.h"
#include <string.h>

#define W 12
#define HALF_FULL 4
#define N_PARTIAL 22

static inline uint64_t sbox(uint64_t x) {
    uint64_t x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

static void constant_layer(uint64_t *s, int round_ctr) {
    for (int i = 0; i < W; i++) s[i] = gl_add(s[i], gl_canon(ORC_ALL_ROUND_CONSTANTS[i + W * round_ctr]));
}
static void sbox_layer(uint64_t *s) {
    for (int i = 0; i < W; i++) s[i] = sbox(s[i]);
}
/* poseidon.rs:168-189 + :236-257: out[r] = sum_i v[(i+r)%12]*CIRC[i] + v[r]*DIAG[r] */
static void mds_layer(uint64_t *s) {
    /* the same sums, accumulated the way poseidon.rs:236-257 does: 32-bit halves of the state in u64 accumulators (the
     * coefficients are < 2^6, 13 terms: no overflow), recombined as lo + (hi << 32) and reduced once */
    uint64_t lo[2 * W], hi[2 * W], out[W];
    for (int i = 0; i < W; i++) {
        lo[i] = lo[i + W] = s[i] & 0xFFFFFFFFULL;
        hi[i] = hi[i + W] = s[i] >> 32;
    }
    for (int r = 0; r < W; r++) {
        uint64_t al = lo[r] * ORC_MDS_MATRIX_DIAG[r], ah = hi[r] * ORC_MDS_MATRIX_DIAG[r];
        for (int i = 0; i < W; i++) {
            al += lo[i + r] * ORC_MDS_MATRIX_CIRC[i];
            ah += hi[i + r] * ORC_MDS_MATRIX_CIRC[i];
        }
        out[r] = gl_reduce128((u128)al + ((u128)ah << 32));
    }
    memcpy(s, out, sizeof(out));
}
static void full_rounds(uint64_t *s, int *round_ctr) {
    for (int r = 0; r < HALF_FULL; r++) {
        constant_layer(s, *round_ctr);
        sbox_layer(s);
        mds_layer(s);
        (*round_ctr)++;
    }
}
/* poseidon.rs:607-617 */
static void partial_rounds_naive(uint64_t *s, int *round_ctr) {
    for (int r = 0; r < N_PARTIAL; r++) {
        constant_layer(s, *round_ctr);
        s[0] = sbox(s[0]);
        mds_layer(s);
        (*round_ctr)++;
    }
}
/* poseidon.rs:577-590 */
static void partial_rounds_fast(uint64_t *s, int *round_ctr) {
    for (int i = 0; i < W; i++) s[i] = gl_add(s[i], gl_canon(ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
    { /* mds_partial_layer_init :332-358 */
        uint64_t res[W];
        memset(res, 0, sizeof(res));
        res[0] = s[0];
        for (int r = 1; r < W; r++)
            for (int c = 1; c < W; c++)
                res[c] = gl_add(res[c], gl_mul(s[r], gl_canon(ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r - 1][c - 1])));
        memcpy(s, res, sizeof(res));
    }
    for (int i = 0; i < N_PARTIAL; i++) {
        s[0] = sbox(s[0]);
        s[0] = gl_add(s[0], gl_canon(ORC_FAST_PARTIAL_ROUND_CONSTANTS[i]));
        /* mds_partial_layer_fast :392-421 */
        uint64_t d = gl_mul(s[0], ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0]);
        for (int k = 1; k < W; k++) d = gl_add(d, gl_mul(s[k], gl_canon(ORC_FAST_PARTIAL_ROUND_W_HATS[i][k - 1])));
        uint64_t res[W];
        res[0] = d;
        for (int k = 1; k < W; k++) res[k] = gl_add(s[k], gl_mul(s[0], gl_canon(ORC_FAST_PARTIAL_ROUND_VS[i][k - 1])));
        memcpy(s, res, sizeof(res));
    }
    *round_ctr += N_PARTIAL;
}

void orc_poseidon(uint64_t state[12]) {
    int rc = 0;
    for (int i = 0; i < W; i++) state[i] = gl_canon(state[i]);
    full_rounds(state, &rc);
    partial_rounds_fast(state, &rc);
    full_rounds(state, &rc);
}
/* One row of the Poseidon TABLE (builtins/poseidon/columns.rs:6-42) for a given permutation input: the witness
 * values trace generation records (core/src/vm hashing via calculate_poseidon_and_generate_intermediate_trace,
 * laid out by circuits/src/generation/poseidon.rs:18-80) are exactly the S-box inputs that PoseidonStark constrains
 * (poseidon_stark.rs:83-141): after the constant layer of full rounds 1..3 (first half) and 0..3 (second half), and
 * state[0] before each partial-round S-box in the fast formulation.  row[0..4) (the filters) are left 0.
 * Pinned by the reference's POSEIDON_ZERO_HASH_* / POSEIDON_1000_HASH_* tables (poseidon_utils.rs:11-287). */
void orc_poseidon_table_row(const uint64_t in[12], uint64_t row[134]) {
    uint64_t s[W];
    int rc = 0;
    memset(row, 0, 134 * sizeof(uint64_t));
    for (int i = 0; i < W; i++) row[4 + i] = s[i] = gl_canon(in[i]);
    for (int r = 0; r < HALF_FULL; r++) {
        constant_layer(s, rc);
        if (r != 0) memcpy(row + 28 + 12 * (r - 1), s, sizeof(s));
        sbox_layer(s);
        mds_layer(s);
        rc++;
    }
    for (int i = 0; i < W; i++) s[i] = gl_add(s[i], gl_canon(ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
    {
        uint64_t res[W];
        memset(res, 0, sizeof(res));
        res[0] = s[0];
        for (int r = 1; r < W; r++)
            for (int c = 1; c < W; c++)
                res[c] = gl_add(res[c], gl_mul(s[r], gl_canon(ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r - 1][c - 1])));
        memcpy(s, res, sizeof(res));
    }
    for (int i = 0; i < N_PARTIAL; i++) {
        row[64 + i] = s[0];
        s[0] = sbox(s[0]);
        if (i < N_PARTIAL - 1) s[0] = gl_add(s[0], gl_canon(ORC_FAST_PARTIAL_ROUND_CONSTANTS[i]));
        uint64_t d = gl_mul(s[0], ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0]);
        for (int k = 1; k < W; k++) d = gl_add(d, gl_mul(s[k], gl_canon(ORC_FAST_PARTIAL_ROUND_W_HATS[i][k - 1])));
        uint64_t res[W];
        res[0] = d;
        for (int k = 1; k < W; k++) res[k] = gl_add(s[k], gl_mul(s[0], gl_canon(ORC_FAST_PARTIAL_ROUND_VS[i][k - 1])));
        memcpy(s, res, sizeof(res));
    }
    rc += N_PARTIAL;
    for (int r = 0; r < HALF_FULL; r++) {
        constant_layer(s, rc);
        memcpy(row + 86 + 12 * r, s, sizeof(s));
        sbox_layer(s);
        mds_layer(s);
        rc++;
    }
    memcpy(row + 16, s, sizeof(s));
}

void orc_poseidon_naive(uint64_t state[12]) {
    int rc = 0;
    for (int i = 0; i < W; i++) state[i] = gl_canon(state[i]);
    full_rounds(state, &rc);
    partial_rounds_naive(state, &rc);
    full_rounds(state, &rc);
}

/* hashing.rs:84-108, num_outputs = 4 */
void orc_poseidon_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
    uint64_t st[W];
    memset(st, 0, sizeof(st));
    for (size_t off = 0; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < len; i++) st[i] = gl_canon(in[off + i]);
        orc_poseidon(st);
    }
    /* n == 0: no absorption, squeeze the zero state (hashing.rs:98-106) */
    memcpy(out, st, 4 * sizeof(uint64_t));
}

/* hashing.rs:66-74 */
void orc_poseidon_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[W];
    memset(st, 0, sizeof(st));
    memcpy(st, l, 32);
    memcpy(st + 4, r, 32);
    orc_poseidon(st);
    memcpy(out, st, 32);
}

/* C::Hasher of the GenericConfig in force (plonk/config.rs:115-122 PoseidonGoldilocksConfig, :153-161
 * Blake3GoldilocksConfig): 0 = PoseidonHash, 1 = Blake3_256<32>.  Process-wide, set between calls (test tooling). */
static int orc_hasher_id = 0;
void orc_set_hasher(int id) { orc_hasher_id = id; }
int orc_get_hasher(void) { return orc_hasher_id; }
void orc_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
    if (orc_hasher_id == 1)
        orc_blake3_hash_no_pad(in, n, out);
    else
        orc_poseidon_hash_no_pad(in, n, out);
}
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    if (orc_hasher_id == 1)
        orc_blake3_two_to_one(l, r, out);
    else
        orc_poseidon_two_to_one(l, r, out);
}
/* H::Permutation of the Challenger (challenger.rs:137-152): PoseidonPermutation or Blake3Permutation */
void orc_challenger_permute(uint64_t state[12]) {
    if (orc_hasher_id == 1)
        orc_blake3_permute(state);
    else
        orc_poseidon(state);
}

/* rows-major leaf hashing of a [nrows][ncols] matrix (MerkleTree::new_v2 leaf loop,
 * merkle_tree/mod.rs:186-201) */
void orc_hash_rows(const uint64_t *rows, size_t nrows, size_t ncols, uint64_t *digests) {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; r++) orc_hash_no_pad(rows + r * ncols, ncols, digests + 4 * r);
}
