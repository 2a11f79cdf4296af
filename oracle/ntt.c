/* ORACLE (test infrastructure, NOT product code) -- see oracle/gl.h header.
 *
 * CPU restatement of the reference's "cfft" (winterfell-derived) FFT over Goldilocks and of the
 * legacy plonky2 radix-2 FFT used as a cross-check.
 *
 * Follows:
 *   plonky2/field/src/cfft/mod.rs        get_twiddles :233-250, get_inv_twiddles :252-271,
 *                                        permute_index :282-290
 *   plonky2/field/src/cfft/serial.rs     evaluate_poly :9-15, evaluate_poly_with_offset :20-50,
 *                                        interpolate_poly :52-63, interpolate_poly_with_offset :65-79,
 *                                        permute :81-89, fft_in_place :91-127 (recursive; restated
 *                                        iteratively below: same butterflies, same twiddle indices)
 *   plonky2/field/src/fft.rs             fft_classic :182 (natural-order radix-2, cross-check only)
 *   plonky2/field/src/polynomial/mod.rs  ifft :60-65, coset_ifft :69-74, coset_fft_with_options :302
 *
 * The concurrent variants (cfft/concurrent.rs) compute the same exact field values with a
 * sqrt(n) x sqrt(n) decomposition; exact arithmetic => identical output, so only the serial
 * dataflow is restated.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* cfft/mod.rs:233-250: root.powers().take(n/2) then bit-reverse permuted */
void orc_get_twiddles(uint64_t *tw, size_t n, int inverse) {
    uint32_t lg = orc_log2_strict(n);
    uint64_t root = gl_root_of_unity((int)lg);
    if (inverse) root = gl_pow(root, (uint64_t)n - 1); /* mod.rs:266 */
    size_t half = n / 2;
    if (half == 0) return;
    uint64_t *tmp = (uint64_t *)malloc(half * sizeof(uint64_t));
    uint64_t acc = 1;
    for (size_t i = 0; i < half; i++) {
        tmp[i] = acc;
        acc = gl_mul(acc, root);
    }
    uint32_t hb = orc_log2_strict(half);
    for (size_t i = 0; i < half; i++) tw[i] = tmp[orc_bitrev(i, hb)];
    free(tmp);
}

/* serial.rs:81-89 */
void orc_permute(uint64_t *v, size_t n) {
    uint32_t lg = orc_log2_strict(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = orc_bitrev(i, lg);
        if (j > i) {
            uint64_t t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    }
}

/* serial.rs:91-127 fft_in_place, restated literally (same recursion, same MAX_LOOP batching,
 * same butterfly order).  Natural-order input, bit-reversed output; the caller permutes. */
static void orc_fft_rec(uint64_t *values, size_t len, const uint64_t *tw, size_t count, size_t stride, size_t offset);
void orc_fft_in_place(uint64_t *v, size_t n, const uint64_t *tw) { orc_fft_rec(v, n, tw, 1, 1, 0); }

#define ORC_MAX_LOOP 256
static void orc_fft_rec(uint64_t *values, size_t len, const uint64_t *tw, size_t count, size_t stride, size_t offset) {
    size_t size = len / stride;
    if (size > 2) {
        if (stride == count && count < ORC_MAX_LOOP) {
            orc_fft_rec(values, len, tw, 2 * count, 2 * stride, offset);
        } else {
            orc_fft_rec(values, len, tw, count, 2 * stride, offset);
            orc_fft_rec(values, len, tw, count, 2 * stride, offset + stride);
        }
    }
    if (size < 2) return;
    for (size_t o = offset; o < offset + count; o++) { /* butterfly, serial.rs:129-139 */
        size_t i = o, j = o + stride;
        uint64_t t = values[i];
        values[i] = gl_add(t, values[j]);
        values[j] = gl_sub(t, values[j]);
    }
    size_t last = offset + size * stride;
    size_t idx = 0;
    for (size_t o = offset; o < last; o += 2 * stride, idx++) {
        if (idx == 0) continue; /* .skip(1) */
        for (size_t j0 = o; j0 < o + count; j0++) { /* butterfly_twiddle, serial.rs:141-151 */
            size_t i = j0, j = j0 + stride;
            uint64_t t = values[i];
            values[j] = gl_mul(values[j], tw[idx]);
            values[i] = gl_add(t, values[j]);
            values[j] = gl_sub(t, values[j]);
        }
    }
}

/* serial.rs:9-15 : coefficients (natural) -> evaluations on H (natural) */
void orc_evaluate_poly(uint64_t *p, size_t n) {
    if (n < 2) return;
    uint64_t *tw = (uint64_t *)malloc((n / 2) * sizeof(uint64_t));
    orc_get_twiddles(tw, n, 0);
    orc_fft_in_place(p, n, tw);
    orc_permute(p, n);
    free(tw);
}

/* serial.rs:52-63 : evaluations on H (natural) -> coefficients (natural) */
void orc_interpolate_poly(uint64_t *v, size_t n) {
    if (n < 2) return;
    uint64_t *tw = (uint64_t *)malloc((n / 2) * sizeof(uint64_t));
    orc_get_twiddles(tw, n, 1);
    orc_fft_in_place(v, n, tw);
    uint64_t inv_len = gl_inv((uint64_t)n % GL_P);
    for (size_t i = 0; i < n; i++) v[i] = gl_mul(v[i], inv_len);
    orc_permute(v, n);
    free(tw);
}

/* serial.rs:20-50 : out[k] = p(offset * g_L^k), L = n*blowup, natural order */
void orc_evaluate_poly_with_offset(const uint64_t *p, size_t n, uint64_t domain_offset, size_t blowup, uint64_t *out) {
    size_t domain = n * blowup;
    uint64_t g = gl_root_of_unity((int)orc_log2_strict(domain));
    uint64_t *tw = (uint64_t *)malloc((n / 2 + 1) * sizeof(uint64_t));
    orc_get_twiddles(tw, n, 0);
    uint32_t bb = orc_log2_strict(blowup);
    for (size_t i = 0; i < blowup; i++) {
        uint64_t *chunk = out + i * n;
        size_t idx = blowup == 1 ? 0 : orc_bitrev(i, bb);
        uint64_t offset = gl_mul(gl_pow(g, idx), domain_offset);
        uint64_t factor = 1;
        for (size_t j = 0; j < n; j++) {
            chunk[j] = gl_mul(p[j], factor);
            factor = gl_mul(factor, offset);
        }
        if (n >= 2) orc_fft_in_place(chunk, n, tw);
    }
    orc_permute(out, domain);
    free(tw);
}

/* serial.rs:65-79 : evaluations on offset*H (natural) -> coefficients (natural) */
void orc_interpolate_poly_with_offset(uint64_t *v, size_t n, uint64_t domain_offset) {
    if (n >= 2) {
        uint64_t *tw = (uint64_t *)malloc((n / 2) * sizeof(uint64_t));
        orc_get_twiddles(tw, n, 1);
        orc_fft_in_place(v, n, tw);
        orc_permute(v, n);
        free(tw);
    }
    uint64_t off_inv = gl_inv(domain_offset);
    uint64_t offset = gl_inv((uint64_t)n % GL_P);
    for (size_t i = 0; i < n; i++) {
        v[i] = gl_mul(v[i], offset);
        offset = gl_mul(offset, off_inv);
    }
}

/* fft.rs:182 fft_classic semantics (bit-reverse, then radix-2 DIT with natural-order roots);
 * cross-check of the cfft restatement (BASELINE config #1 names plonky2::field::fft). */
void orc_fft_classic(uint64_t *v, size_t n) {
    uint32_t lg = orc_log2_strict(n);
    orc_permute(v, n);
    for (uint32_t s = 1; s <= lg; s++) {
        size_t m = (size_t)1 << s, half = m / 2;
        uint64_t wm = gl_root_of_unity((int)s);
        for (size_t k = 0; k < n; k += m) {
            uint64_t w = 1;
            for (size_t j = 0; j < half; j++) {
                uint64_t t = gl_mul(w, v[k + j + half]);
                uint64_t u = v[k + j];
                v[k + j] = gl_add(u, t);
                v[k + j + half] = gl_sub(u, t);
                w = gl_mul(w, wm);
            }
        }
    }
}

/* naive O(n^2) evaluation: the reference's own tests check the FFTs against this
 * (polynomial/mod.rs:495-540, fft.rs:219-251) */
uint64_t orc_poly_eval(const uint64_t *coeffs, size_t n, uint64_t x) {
    uint64_t acc = 0;
    for (size_t i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), coeffs[i]);
    return acc;
}

/* ---- batch wrappers over column-major [ncols][n] blocks (OpenMP over columns mirrors the
 * reference's rayon par_iter over polynomials, oracle.rs:56-60, :117-129) ---- */
void orc_ifft_batch(uint64_t *cols, size_t ncols, size_t n) {
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < ncols; c++) orc_interpolate_poly(cols + c * n, n);
}
void orc_lde_batch(const uint64_t *coeffs, size_t ncols, size_t n, uint64_t shift, size_t blowup, uint64_t *out) {
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < ncols; c++)
        orc_evaluate_poly_with_offset(coeffs + c * n, n, shift, blowup, out + c * n * blowup);
}
