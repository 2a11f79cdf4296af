// Device-resident PolynomialBatch (internal).
#pragma once
#include <memory>

#include "common.h"

struct ola_batch {
    size_t ncols = 0;
    uint32_t log_n = 0, rate_bits = 0, cap_height = 0;
    // coset shard: this batch holds LDE cosets [coset_first, coset_first + 2^shard_bits) = a contiguous
    // leaf range = whole cap subtrees of the global Merkle tree.  A full batch has shard_bits == rate_bits.
    uint32_t shard_bits = 0, coset_first = 0;
    uint64_t* d_coeffs = nullptr;  // [ncols][n]
    uint64_t* d_lde = nullptr;     // [ncols][n << shard_bits]  leaf order
    uint64_t* d_nodes = nullptr;   // [2 * (n << shard_bits)][4] heap order (the global subtree over this leaf range)
    uint32_t leaf_bits() const { return log_n + shard_bits; }
    uint32_t local_cap_height() const { return cap_height - (rate_bits - shard_bits); }  // cap level inside the local subtree
};

namespace ola {
void dev_alloc(uint64_t** p, size_t n_u64);
void dev_free(void* p);
void set_alloc_stream(cudaStream_t s);
// dst[i] = canonical(src[i]); dst may alias src
void canon_copy(ola_ctx* ctx, uint64_t* dst, const uint64_t* src, size_t n);
// grow-only per-context workspace of at least n_u64 elements (synchronises the stream when it has to grow)
uint64_t* ctx_scratch(ola_ctx* ctx, size_t n_u64);
void gather_rows(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t ncols, size_t first, size_t count,
                 uint64_t* out_host);
ola_batch* batch_commit(ola_ctx* ctx, const uint64_t* cols, bool on_device, size_t ncols, uint32_t log_n, bool is_coeffs,
                        uint32_t rate_bits, uint32_t cap_height, int coset_first = 0, int coset_count = -1);
// PolynomialBatch::from_values / from_coeffs up to and including lde_values (oracle.rs:45-60, :101-129): coefficients and
// the leaf-order LDE into caller-owned device buffers, no Merkle tree.  Host input is uploaded in column chunks whose
// transforms overlap the next chunk's copy.
void lde_batch(ola_ctx* ctx, const uint64_t* cols, bool on_device, size_t ncols, uint32_t log_n, bool is_coeffs, uint32_t rate_bits,
               uint64_t* d_coeffs, uint64_t* d_lde);
// multi-GPU prover: column-sharded iNTT + all-gather of coefficients, coset-sharded LDE / hashing / tree (device input)
ola_batch* batch_commit_dist(ola_ctx* ctx, const uint64_t* d_cols, size_t ncols, uint32_t log_n, bool is_coeffs, uint32_t rate_bits,
                             uint32_t cap_height);
void batch_release(ola_batch* b);
void batch_get_cap(ola_ctx* ctx, const ola_batch* b, uint64_t* cap_host);
void batch_get_leaves(ola_ctx* ctx, const ola_batch* b, size_t first, size_t count, uint64_t* out_host);
int batch_prove_leaf(ola_ctx* ctx, const ola_batch* b, size_t leaf, uint64_t* sib_host);
}  // namespace ola
