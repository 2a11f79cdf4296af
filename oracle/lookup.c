/* ORACLE (test infrastructure, NOT product code): the lookup-argument column builder and the RangeCheck table generator.
 *
 *   orc_permuted_cols       circuits/src/stark/lookup.rs:68-131   permuted_cols (the Halo2-style permuted input / table pair)
 *   orc_generate_rc_trace   circuits/src/generation/builtin.rs:249-316   generate_rc_trace
 *                           core/src/trace/trace.rs:401-425      insert_rangecheck (limbs = the two 16-bit halves of val)
 *                           circuits/src/builtins/rangecheck/columns.rs:27-44   column order, RANGE_CHECK_U16_SIZE = 2^16
 *
 * A sequential restatement: sort both columns, then the reference's merge walk with its LIFO list of unused table values
 * and FIFO list of unfilled positions, statement by statement. */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* lookup.rs:68-131 */
void orc_permuted_cols(const uint64_t *inputs, const uint64_t *table, size_t n, uint64_t *permuted_inputs, uint64_t *permuted_table) {
    uint64_t *si = permuted_inputs;                       /* sorted_inputs is returned as the permuted inputs (:130) */
    uint64_t *st = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *unused_vals = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    size_t *unused_inds = (size_t *)malloc((n + 1) * sizeof(size_t));
    size_t nv = 0, ni = 0, i = 0, j = 0;
    for (size_t k = 0; k < n; ++k) {                      /* to_canonical before comparing (:80-89) */
        si[k] = gl_canon(inputs[k]);
        st[k] = gl_canon(table[k]);
    }
    qsort(si, n, sizeof(uint64_t), cmp_u64);
    qsort(st, n, sizeof(uint64_t), cmp_u64);
    memset(permuted_table, 0, n * sizeof(uint64_t));
    while (j < n && i < n) {                              /* :96-117 */
        const uint64_t input_val = si[i], table_val = st[j];
        if (input_val > table_val) {
            unused_vals[nv++] = st[j];
            j++;
        } else if (input_val < table_val) {
            if (nv > 0)
                permuted_table[i] = unused_vals[--nv];    /* Vec::pop: the most recently skipped table value */
            else
                unused_inds[ni++] = i;
            i++;
        } else {
            permuted_table[i] = st[j];
            i++;
            j++;
        }
    }
    for (size_t jj = j; jj < n; ++jj) unused_vals[nv++] = st[jj];   /* :120-122 */
    for (size_t ii = i; ii < n; ++ii) unused_inds[ni++] = ii;       /* :123-125 */
    /* zip_eq (:126-128): both lists have the same length, filled front to front */
    for (size_t k = 0; k < ni && k < nv; ++k) permuted_table[unused_inds[k]] = unused_vals[k];
    free(st);
    free(unused_vals);
    free(unused_inds);
}

/* builtin.rs:249-316.  vals[nrows] with kinds[nrows] in {0: cpu, 1: memory sort, 2: memory region, 3: comparison} (the
 * filter column set to 1 for that row, trace.rs:401-425; rangecheck/columns.rs:27-30).  out is column-major [12][n] with
 * n = max(next_power_of_two(nrows), 2^16); returns n. */
size_t orc_generate_rc_trace(const uint64_t *vals, const uint8_t *kinds, size_t nrows, uint64_t *out, size_t out_cap_rows) {
    size_t n = nrows > 65536 ? nrows : 65536;             /* max(trace_len, RANGE_CHECK_U16_SIZE) (:253) */
    size_t p = 2;
    while (p < n) p <<= 1;                                /* next_power_of_two (:254-262) */
    n = p;
    if (out == NULL || out_cap_rows < n) return n;
    memset(out, 0, 12 * n * sizeof(uint64_t));
    for (size_t i = 0; i < nrows; ++i) {                  /* :264-276 */
        out[(size_t)kinds[i] * n + i] = 1;                /* CPU_FILTER 0, MEMORY_SORT_FILTER 1, MEMORY_REGION_FILTER 2, CMP_FILTER 3 */
        const uint64_t v = gl_canon(vals[i]);
        out[4 * n + i] = v;                               /* VAL */
        out[5 * n + i] = v & 0xFFFF;                      /* LIMB_LO  (split_u16_limbs_from_field) */
        out[6 * n + i] = v >> 16;                         /* LIMB_HI */
    }
    for (size_t i = 0; i < n; ++i) out[9 * n + i] = i < 65536 ? i : 65535;   /* FIX_RANGE_CHECK_U16, padded with its last value (:278-293) */
    orc_permuted_cols(out + 5 * n, out + 9 * n, n, out + 7 * n, out + 10 * n);  /* LIMB_LO_PERMUTED, FIX_..._PERMUTED_LO (:295-301) */
    orc_permuted_cols(out + 6 * n, out + 9 * n, n, out + 8 * n, out + 11 * n);  /* LIMB_HI_PERMUTED, FIX_..._PERMUTED_HI (:303-309) */
    return n;
}

/* ---- generate_bitwise_trace (circuits/src/generation/builtin.rs:35-206) and generate_cmp_trace (:208-247) -------------------
 * Column order: builtins/bitwise/columns.rs:23-62 (FILTER 0, TAG 1, OP0 2, OP1 3, RES 4, OP0_LIMBS 5..9, OP1_LIMBS 9..13,
 * RES_LIMBS 13..17, the three permuted limb ranges 17..29, COMPRESS_LIMBS 29..33, COMPRESS_PERMUTED 33..37,
 * FIX_RANGE_CHECK_U8 37, its twelve permuted copies 38..50, FIX_TAG 50, FIX_BITWSIE_OP0/OP1/RES 51..54, FIX_COMPRESS 54,
 * FIX_COMPRESS_PERMUTED 55..59; COL_NUM_BITWISE = 59).
 *
 * Followed statement by statement, INCLUDING the reference's indexing of the fourth limb: it writes op0_3 to
 * trace[OP0_LIMBS.end] (= OP1_LIMBS.start, overwritten by op1_0 two statements later), op1_3 to trace[OP1_LIMBS.end]
 * (= RES_LIMBS.start, overwritten by res_0) and res_3 to trace[RES_LIMBS.end] (= OP0_LIMBS_PERMUTED.start, overwritten by
 * the permuted inputs), so columns 8, 12 and 16 -- limb 3 of op0, op1, res -- stay zero (builtin.rs:66, :71, :76).  A drop-in
 * generator has to produce the table the reference produces; operands below 2^24 give a table that satisfies the AIR. */
uint64_t orc_compress_challenge(const uint64_t *const *cols, uint32_t ncols, size_t n); /* stark_api.cpp */

/* tags[nrows] (c.opcode: the tag the row carries), op0 / op1 / res[nrows].  out = [59][n], n = max(next_power_of_two(nrows),
 * 2^18) (3 * 2^16 fixed rows, builtin.rs:39-53); returns n; *beta_out = the compress challenge. */
size_t orc_generate_bitwise_trace(const uint64_t *tags, const uint64_t *op0, const uint64_t *op1, const uint64_t *res, size_t nrows, uint64_t *out,
                                  size_t out_cap_rows, uint64_t *beta_out) {
    size_t n = nrows > 3 * 65536 ? nrows : 3 * 65536;
    size_t p = 2;
    while (p < n) p <<= 1;
    n = p;
    if (out == NULL || out_cap_rows < n) return n;
    memset(out, 0, 59 * n * sizeof(uint64_t));
#define T(c, i) out[(size_t)(c) * n + (i)]
    for (size_t i = 0; i < nrows; ++i) { /* :55-77 */
        const uint64_t a = gl_canon(op0[i]), b = gl_canon(op1[i]), r = gl_canon(res[i]);
        T(0, i) = 1;
        T(1, i) = tags[i];
        T(2, i) = a;
        T(3, i) = b;
        T(4, i) = r;
        T(5, i) = a & 255;          /* OP0_LIMBS.start     (split_limbs_from_field, core/src/utils.rs:9-16) */
        T(6, i) = (a >> 8) & 255;   /* OP0_LIMBS.start + 1 */
        T(7, i) = (a >> 16) & 255;  /* OP0_LIMBS.start + 2 */
        T(9, i) = (a >> 24) & 255;  /* OP0_LIMBS.end (!) */
        T(9, i) = b & 255;          /* OP1_LIMBS.start */
        T(10, i) = (b >> 8) & 255;
        T(11, i) = (b >> 16) & 255;
        T(13, i) = (b >> 24) & 255; /* OP1_LIMBS.end (!) */
        T(13, i) = r & 255;         /* RES_LIMBS.start */
        T(14, i) = (r >> 8) & 255;
        T(15, i) = (r >> 16) & 255;
        T(17, i) = (r >> 24) & 255; /* RES_LIMBS.end (!): overwritten by the permuted inputs below */
    }
    size_t index = 0; /* :82-117: the fixed tables; Opcode::AND = 18, OR = 17, XOR = 16 (core/src/program/instruction.rs:44-46) */
    for (size_t x = 0; x < 256; ++x) {
        T(37, x) = x;
        for (size_t y = 0; y < 256; ++y) {
            T(51, index) = x, T(52, index) = y, T(53, index) = x & y, T(50, index) = 1ull << 18;
            T(51, 65536 + index) = x, T(52, 65536 + index) = y, T(53, 65536 + index) = x | y, T(50, 65536 + index) = 1ull << 17;
            T(51, 2 * 65536 + index) = x, T(52, 2 * 65536 + index) = y, T(53, 2 * 65536 + index) = x ^ y, T(50, 2 * 65536 + index) = 1ull << 16;
            index++;
        }
    }
    const uint64_t *cols[12]; /* :120-131 */
    for (int k = 0; k < 4; ++k) cols[k] = out + (size_t)(5 + k) * n, cols[4 + k] = out + (size_t)(9 + k) * n, cols[8 + k] = out + (size_t)(13 + k) * n;
    const uint64_t beta = orc_compress_challenge(cols, 12, n);
    const uint64_t b2 = gl_mul(beta, beta), b3 = gl_mul(b2, beta);
    for (size_t i = 0; i < n; ++i) { /* :133-159 */
        for (int k = 0; k < 4; ++k)
            T(29 + k, i) = gl_add(gl_add(gl_add(gl_canon(T(1, i)), gl_mul(T(5 + k, i), beta)), gl_mul(T(9 + k, i), b2)), gl_mul(T(13 + k, i), b3));
        T(54, i) = gl_add(gl_add(gl_add(T(50, i), gl_mul(T(51, i), beta)), gl_mul(T(52, i), b2)), gl_mul(T(53, i), b3));
    }
    for (int k = 0; k < 4; ++k) { /* :162-195 */
        orc_permuted_cols(&T(5 + k, 0), &T(37, 0), n, &T(17 + k, 0), &T(38 + k, 0));
        orc_permuted_cols(&T(9 + k, 0), &T(37, 0), n, &T(21 + k, 0), &T(38 + 4 + k, 0));
        orc_permuted_cols(&T(13 + k, 0), &T(37, 0), n, &T(25 + k, 0), &T(38 + 8 + k, 0));
        orc_permuted_cols(&T(29 + k, 0), &T(54, 0), n, &T(33 + k, 0), &T(55 + k, 0));
    }
#undef T
    if (beta_out) *beta_out = beta;
    return n;
}

/* generate_cmp_trace (builtin.rs:208-247): rows (op0, op1, gte, abs_diff, abs_diff_inv, filter_looking_rc) as the executor
 * recorded them (cmp/columns.rs:16-22); padding rows (1, 0, 1, 1, 1, 0).  out = [6][n], n = max(next_power_of_two(nrows), 2). */
size_t orc_generate_cmp_trace(const uint64_t *cells /* [nrows][6] */, size_t nrows, uint64_t *out, size_t out_cap_rows) {
    size_t n = 2;
    while (n < nrows) n <<= 1;
    if (out == NULL || out_cap_rows < n) return n;
    memset(out, 0, 6 * n * sizeof(uint64_t));
    for (size_t i = 0; i < nrows; ++i)
        for (int c = 0; c < 6; ++c) out[(size_t)c * n + i] = gl_canon(cells[i * 6 + c]);
    for (size_t i = nrows; i < n; ++i) out[0 * n + i] = 1, out[2 * n + i] = 1, out[3 * n + i] = 1, out[4 * n + i] = 1;
    return n;
}
