// C::Hasher of the context: PoseidonHash (PoseidonGoldilocksConfig, plonky2/plonky2/src/plonk/config.rs:115-122) or
// Blake3_256<32> (Blake3GoldilocksConfig, :153-161), chosen by ola_set_hasher.  The commitment code (batch.cu, fri.cu,
// api.cu) calls these; C::InnerHasher -- the proof-of-work hash -- is Poseidon in both configs and is not dispatched.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct ola_ctx;
namespace ola {
namespace blake3 {
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests);
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols, uint64_t* d_digests);
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop);
// digests[i] = hash_no_pad(flatten(vals[arity i .. arity (i + 1)))) of a planar [2][len] extension vector
void fri_leaves(ola_ctx* ctx, const uint64_t* d_vals, size_t len, int arity, size_t leaf_first, size_t nleaves, uint64_t* d_digests);
}  // namespace blake3
namespace hasher {
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests);
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols, uint64_t* d_digests);
// reduce heap-ordered nodes from the leaf digests at [nleaves, 2 nleaves) down to level `stop` (its first index)
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop);
}  // namespace hasher
}  // namespace ola
