/* ORACLE (test infrastructure, NOT product code) -- see oracle/gl.h header.
 *
 * CPU restatement of the polynomial-batch commitment used three times per table by the prover.
 *
 * Follows:
 *   plonky2/plonky2/src/fri/oracle.rs   from_values :45-64 (per-column ifft), from_coeffs :66-99
 *       (lde_values :101-129 = coset_fft_with_options(shift = 7, blowup = 2^rate_bits);
 *        transpose :84; reverse_index_bits_in_place :85; MerkleTree::new_v2 :86-90)
 *   plonky2/plonky2/src/util/mod.rs:20  transpose
 *   plonky2/util/src/lib.rs:190         reverse_index_bits_in_place
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* cols: column-major [ncols][n] (values on H, or coefficients when is_coeffs).
 * coeffs_out : [ncols][n]            natural-order coefficients          (PolynomialBatch.polynomials)
 * leaves_out : [n<<rate_bits][ncols] row-major, leaf r = LDE row bitrev(r) (merkle_tree.leaves)
 * digests_out/cap_out: as orc_merkle_new_v2.  Any *_out may be NULL except cap_out. */
int orc_commit(const uint64_t *cols, size_t ncols, size_t n, int is_coeffs, uint32_t rate_bits, uint32_t cap_height,
               uint64_t *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out) {
    size_t blowup = (size_t)1 << rate_bits, L = n * blowup;
    uint32_t lgL = orc_log2_strict(L);
    uint64_t *coeffs = coeffs_out ? coeffs_out : (uint64_t *)malloc(ncols * n * 8);
    for (size_t i = 0; i < ncols * n; i++) coeffs[i] = gl_canon(cols[i]);
    if (!is_coeffs) orc_ifft_batch(coeffs, ncols, n);
    uint64_t *lde = (uint64_t *)malloc(ncols * L * 8);
    orc_lde_batch(coeffs, ncols, n, GL_GEN, blowup, lde);
    uint64_t *leaves = leaves_out ? leaves_out : (uint64_t *)malloc(ncols * L * 8);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < L; r++) {
        size_t src = orc_bitrev(r, lgL);
        for (size_t c = 0; c < ncols; c++) leaves[r * ncols + c] = lde[c * L + src];
    }
    free(lde);
    uint64_t *dig = digests_out;
    size_t num_digests = 2 * (L - ((size_t)1 << cap_height));
    if (!dig) dig = (uint64_t *)malloc((num_digests ? num_digests : 1) * 32);
    int rc = orc_merkle_new_v2(leaves, L, ncols, cap_height, dig, cap_out);
    if (!digests_out) free(dig);
    if (!leaves_out) free(leaves);
    if (!coeffs_out) free(coeffs);
    return rc;
}
