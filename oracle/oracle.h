/* ORACLE (test infrastructure, NOT product code) -- public declarations of the CPU restatement.
 * See the header of each .c file for the reference file:line each function follows. */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include "gl.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ntt.c */
void orc_get_twiddles(uint64_t *tw, size_t n, int inverse);
void orc_permute(uint64_t *v, size_t n);
void orc_fft_in_place(uint64_t *v, size_t n, const uint64_t *tw);
void orc_evaluate_poly(uint64_t *p, size_t n);
void orc_interpolate_poly(uint64_t *v, size_t n);
void orc_evaluate_poly_with_offset(const uint64_t *p, size_t n, uint64_t domain_offset, size_t blowup, uint64_t *out);
void orc_interpolate_poly_with_offset(uint64_t *v, size_t n, uint64_t domain_offset);
void orc_fft_classic(uint64_t *v, size_t n);
uint64_t orc_poly_eval(const uint64_t *coeffs, size_t n, uint64_t x);
void orc_ifft_batch(uint64_t *cols, size_t ncols, size_t n);
void orc_lde_batch(const uint64_t *coeffs, size_t ncols, size_t n, uint64_t shift, size_t blowup, uint64_t *out);

/* poseidon.c */
void orc_poseidon(uint64_t state[12]);
void orc_poseidon_naive(uint64_t state[12]);
/* one row of the Poseidon table (134 columns, filters 0) for a permutation input (generation/poseidon.rs:18-80) */
void orc_poseidon_table_row(const uint64_t in[12], uint64_t row[134]);
void orc_poseidon_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void orc_poseidon_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
/* the Hasher in force (0 PoseidonHash, 1 Blake3_256<32>): leaf / node hashing and the challenger's permutation */
void orc_set_hasher(int id);
int orc_get_hasher(void);
void orc_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void orc_challenger_permute(uint64_t state[12]);
void orc_hash_rows(const uint64_t *rows, size_t nrows, size_t ncols, uint64_t *digests);

/* lookup.c */
void orc_permuted_cols(const uint64_t *inputs, const uint64_t *table, size_t n, uint64_t *permuted_inputs, uint64_t *permuted_table);
size_t orc_generate_rc_trace(const uint64_t *vals, const uint8_t *kinds, size_t nrows, uint64_t *out, size_t out_cap_rows);
size_t orc_generate_bitwise_trace(const uint64_t *tags, const uint64_t *op0, const uint64_t *op1, const uint64_t *res, size_t nrows, uint64_t *out,
                                  size_t out_cap_rows, uint64_t *beta_out);
size_t orc_generate_cmp_trace(const uint64_t *cells, size_t nrows, uint64_t *out, size_t out_cap_rows);

/* generation_cpu.c */
void orc_generate_cpu_trace(const uint64_t *steps, size_t nrows, size_t n, uint64_t *out);
void orc_generate_memory_trace(const uint64_t *cells, size_t ncells, size_t n, uint64_t *out);
size_t orc_generate_prog_trace(const uint64_t *steps, size_t nsteps, const uint64_t *prog_rows, size_t m, const uint64_t roots[8], uint64_t *out,
                               size_t out_cap_rows, uint64_t *beta_out);

/* generation_small.c: the five small tables (records in struct field order, see the file).  Each returns the reference's row count and
 * fills out[ncols][out_rows] when out_rows >= that count. */
size_t orc_generate_poseidon_chunk_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows);
size_t orc_generate_storage_access_trace(const uint64_t *rows, size_t n_access, size_t n_prog, uint64_t *out, size_t out_rows);
size_t orc_generate_tape_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows);
size_t orc_generate_sccall_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows);
size_t orc_generate_prog_chunk_trace(const uint64_t *prog_rows, size_t m, uint64_t *out, size_t out_rows);

/* blake3.c */
void orc_blake3(const uint8_t *in, size_t len, uint8_t out[32]);
void orc_blake3_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void orc_blake3_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void orc_blake3_permute(uint64_t state[12]);
void orc_bytes_hash_to_fields(const uint64_t h[4], uint64_t out[5]);

/* merkle.c */
void orc_build_merkle_nodes(const uint64_t *leaf_hashes, size_t nleaves, uint64_t *nodes);
int orc_merkle_new_v2(const uint64_t *rows, size_t nrows, size_t ncols, uint32_t cap_height, uint64_t *digests_out,
                      uint64_t *cap_out);
int orc_merkle_prove(const uint64_t *digests, size_t nrows, uint32_t cap_height, size_t leaf_index, uint64_t *siblings_out);
int orc_merkle_verify(const uint64_t *leaf, size_t ncols, size_t leaf_index, const uint64_t *cap, const uint64_t *siblings,
                      size_t nsib);

/* pcs.c : PolynomialBatch::from_values / from_coeffs (fri/oracle.rs:45-99) */
int orc_commit(const uint64_t *cols, size_t ncols, size_t n, int is_coeffs, uint32_t rate_bits, uint32_t cap_height,
               uint64_t *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out);

#ifdef __cplusplus
}
#endif
#endif
