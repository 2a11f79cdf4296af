// BLAKE3 kernels of the Blake3GoldilocksConfig commitment: leaf hashing, Merkle level reduction, FRI layer leaves --
// and the dispatch on the context's hasher that the commitment code calls (namespace ola::hasher).
//
// Replaces, for C::Hasher = Blake3_256<32> (plonky2/plonky2/src/plonk/config.rs:153-161):
//   plonky2/plonky2/src/hash/merkle_tree/mod.rs   new_v2 leaf loop :186-201 (H::hash_no_pad per row),
//       build_merkle_nodes :311-337 / merkle_tree/concurrent.rs:14-66 (H::two_to_one per level)
//   plonky2/plonky2/src/fri/prover.rs:72-121      fri_committed_trees leaves (16 extension values per leaf)
// with H from plonky2/plonky2/src/hash/blake3.rs:201-234.
//
// Layout is the Poseidon path's: the LDE is read COLUMN-major (thread r reads element r of every column: one coalesced
// 256-byte segment per warp and column), digests are 4 u64 (here: the 32 digest bytes, little-endian words, NOT field
// elements -- never reduced), nodes are heap-ordered.  Algorithmic bytes: 8 * ncols read + 32 written per leaf, 96 per
// node.  ~900 ALU instructions per 64-byte block: 94 columns = 12 blocks = ~11k instructions per leaf against ~330k for
// the 12 Poseidon permutations of the same leaf.
#include "blake3.cuh"
#include "common.h"
#include "gl.cuh"
#include "poseidon.cuh"

namespace ola {
namespace blake3 {

struct ColLoad {
    const uint64_t* base;
    size_t stride, r;
    __device__ __forceinline__ uint64_t operator()(size_t c) const { return gl::canon(base[c * stride + r]); }
};
struct RowLoad {
    const uint64_t* row;
    __device__ __forceinline__ uint64_t operator()(size_t c) const { return gl::canon(row[c]); }
};
struct FriLoad {  // flatten(values[arity i .. arity (i + 1))): (c0, c1) interleaved, planar source
    const uint64_t *c0, *c1;
    __device__ __forceinline__ uint64_t operator()(size_t j) const { return gl::canon((j & 1) ? c1[j >> 1] : c0[j >> 1]); }
};

template <bool ROWMAJOR, bool MULTI>
__global__ void __launch_bounds__(256) hash_rows_kernel(const uint64_t* __restrict__ base, size_t col_stride, size_t nrows, size_t ncols,
                                                        uint64_t* __restrict__ digests) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint64_t h[4];
    if (ROWMAJOR)
        hash_u64s<MULTI>(ncols, RowLoad{base + r * ncols}, h);
    else
        hash_u64s<MULTI>(ncols, ColLoad{base, col_stride, r}, h);
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + 4 * r);
    d[0] = make_ulonglong2(h[0], h[1]);
    d[1] = make_ulonglong2(h[2], h[3]);
}

// nodes[i] = two_to_one(nodes[2i], nodes[2i+1]) for i in [first, first + count)
__global__ void __launch_bounds__(256) merkle_level_kernel(uint64_t* nodes, size_t first, size_t count) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const size_t i = first + k;
    const ulonglong2* ch = reinterpret_cast<const ulonglong2*>(nodes + 8 * i);
    const ulonglong2 a = ch[0], b = ch[1], c = ch[2], d = ch[3];
    const uint64_t l[4] = {a.x, a.y, b.x, b.y}, r[4] = {c.x, c.y, d.x, d.y};
    uint64_t h[4];
    two_to_one(l, r, h);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(nodes + 4 * i);
    o[0] = make_ulonglong2(h[0], h[1]);
    o[1] = make_ulonglong2(h[2], h[3]);
}

__global__ void __launch_bounds__(256) fri_leaves_kernel(const uint64_t* __restrict__ vals /* [2][len] */, size_t len, int arity, size_t leaf_first,
                                                         size_t leaf_count, uint64_t* __restrict__ digests) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= leaf_count) return;
    uint64_t h[4];
    hash_u64s<false>((size_t)2 * arity, FriLoad{vals + (leaf_first + i) * arity, vals + len + (leaf_first + i) * arity}, h);
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + 4 * i);
    d[0] = make_ulonglong2(h[0], h[1]);
    d[1] = make_ulonglong2(h[2], h[3]);
}

static void check_width(size_t ncols) {
    OLA_CHECK(ncols <= MAX_U64S, OLA_ERR_INVALID_ARG, "BLAKE3 leaf wider than 2048 elements (16 chunks)");
}
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests) {
    if (!nrows) return;
    check_width(ncols);
    const unsigned blocks = (unsigned)((nrows + 255) / 256);
    {
        Launch lz(ctx, "blake3_leaves_rowmajor");
        if (ncols <= 128)
            hash_rows_kernel<true, false><<<blocks, 256, 0, ctx->stream>>>(d_rows, 0, nrows, ncols, d_digests);
        else
            hash_rows_kernel<true, true><<<blocks, 256, 0, ctx->stream>>>(d_rows, 0, nrows, ncols, d_digests);
    }
    check_launch("blake3 hash_rows_kernel<row>");
}
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols, uint64_t* d_digests) {
    if (!nrows) return;
    check_width(ncols);
    const unsigned blocks = (unsigned)((nrows + 255) / 256);
    {
        Launch lz(ctx, "blake3_leaves");
        if (ncols <= 128)
            hash_rows_kernel<false, false><<<blocks, 256, 0, ctx->stream>>>(d_cols, col_stride, nrows, ncols, d_digests);
        else
            hash_rows_kernel<false, true><<<blocks, 256, 0, ctx->stream>>>(d_cols, col_stride, nrows, ncols, d_digests);
    }
    check_launch("blake3 hash_rows_kernel<col>");
}
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop) {
    if (stop < 1) stop = 1;
    for (size_t first = nleaves / 2; first >= stop && first >= 1; first /= 2) {
        {
            Launch lz(ctx, "blake3_merkle_level");
            merkle_level_kernel<<<(unsigned)((first + 255) / 256), 256, 0, ctx->stream>>>(d_nodes, first, first);
        }
        check_launch("blake3 merkle_level_kernel");
        if (first == 1) break;
    }
}
void fri_leaves(ola_ctx* ctx, const uint64_t* d_vals, size_t len, int arity, size_t leaf_first, size_t nleaves, uint64_t* d_digests) {
    if (!nleaves) return;
    OLA_CHECK(2 * (size_t)arity <= 128, OLA_ERR_INVALID_ARG, "FRI arity too large for a one-chunk leaf");
    {
        Launch lz(ctx, "blake3_fri_leaves");
        fri_leaves_kernel<<<(unsigned)((nleaves + 255) / 256), 256, 0, ctx->stream>>>(d_vals, len, arity, leaf_first, nleaves, d_digests);
    }
    check_launch("blake3 fri_leaves_kernel");
}

}  // namespace blake3

// ---- C::Hasher of the context (ola_set_hasher): what the commitment code calls ----
namespace hasher {
inline bool is_blake3(const ola_ctx* ctx) { return ctx->hasher == OLA_HASH_BLAKE3; }
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests) {
    if (is_blake3(ctx))
        blake3::hash_rows_rowmajor(ctx, d_rows, nrows, ncols, d_digests);
    else
        poseidon::hash_rows_rowmajor(ctx, d_rows, nrows, ncols, d_digests);
}
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols, uint64_t* d_digests) {
    if (is_blake3(ctx))
        blake3::hash_rows_colmajor(ctx, d_cols, col_stride, nrows, ncols, d_digests);
    else
        poseidon::hash_rows_colmajor(ctx, d_cols, col_stride, nrows, ncols, d_digests);
}
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop) {
    if (is_blake3(ctx))
        blake3::merkle_levels(ctx, d_nodes, nleaves, stop);
    else
        poseidon::merkle_levels(ctx, d_nodes, nleaves, stop);
}
}  // namespace hasher
}  // namespace ola
