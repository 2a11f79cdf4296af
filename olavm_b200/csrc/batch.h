// Device-resident PolynomialBatch (internal).
#pragma once
#include <memory>

#include "common.h"

struct ola_batch {
    size_t ncols = 0;
    uint32_t log_n = 0, rate_bits = 0, cap_height = 0;
    uint64_t* d_coeffs = nullptr;  // [ncols][n]
    uint64_t* d_lde = nullptr;     // [ncols][n << rate_bits]  leaf order
    uint64_t* d_nodes = nullptr;   // [2 * (n << rate_bits)][4] heap order
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
                        uint32_t rate_bits, uint32_t cap_height);
void batch_release(ola_batch* b);
void batch_get_cap(ola_ctx* ctx, const ola_batch* b, uint64_t* cap_host);
void batch_get_leaves(ola_ctx* ctx, const ola_batch* b, size_t first, size_t count, uint64_t* out_host);
int batch_prove_leaf(ola_ctx* ctx, const ola_batch* b, size_t leaf, uint64_t* sib_host);
}  // namespace ola
