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
