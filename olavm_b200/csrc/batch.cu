// PolynomialBatch on the device: commit = iNTT -> coset LDE -> Poseidon leaf sponge -> Merkle levels.
//
// Replaces plonky2/plonky2/src/fri/oracle.rs:45-99 (PolynomialBatch::from_values / from_coeffs), the
// per-row accessors :132-164 and MerkleTree::{get, prove} (hash/merkle_tree/mod.rs:268-308).
//
// HBM layout of a batch (everything stays resident between prover phases):
//   coeffs [ncols][n]        natural-order coefficients                 (PolynomialBatch.polynomials)
//   lde    [ncols][8n]       COLUMN-major, "leaf order": lde[c][r] = p_c(7 * g^bitrev(r)); rows
//                            [i*n, (i+1)*n) are exactly coset 7*g^bitrev3(i)*H_n, i.e. 2 of the 16 cap
//                            subtrees -> natural multi-GPU shard
//   nodes  [2*8n][4]         heap order: node 1 = root, children 2i / 2i+1, leaf digests at [8n, 16n)
// The reference materialises row-major `leaves: Vec<Vec<F>>` plus an interleaved `digests` vector
// (merkle_tree/mod.rs:40-58); both are derivable views of the arrays above, produced on demand by
// ola_batch_get_leaves / ola_batch_prove_leaf.
#include "batch.h"

#include "gl.cuh"
#include "ntt.h"
#include "hasher.h"
#include "poseidon.cuh"

namespace ola {

__global__ void canon_copy_kernel(uint64_t* dst, const uint64_t* src, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = gl::canon(src[i]);
}

// out[k][c] = lde[c][first + k]
__global__ void gather_rows_kernel(const uint64_t* lde, size_t col_stride, size_t ncols, size_t first, size_t count,
                                   uint64_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * ncols) return;
    size_t k = i / ncols, c = i % ncols;
    out[i] = lde[c * col_stride + first + k];
}

// siblings bottom-up: level j sibling = nodes[((nleaves + leaf) >> j) ^ 1]
__global__ void gather_path_kernel(const uint64_t* nodes, size_t nleaves, size_t leaf, int nsib, uint64_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsib * 4) return;
    int j = i / 4, w = i % 4;
    size_t idx = ((nleaves + leaf) >> j) ^ 1;
    out[i] = nodes[idx * 4 + w];
}

void canon_copy(ola_ctx* ctx, uint64_t* dst, const uint64_t* src, size_t n) {
    if (!n) return;
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    canon_copy_kernel<<<blocks, 256, 0, ctx->stream>>>(dst, src, n);
    check_launch("canon_copy_kernel");
    count_launch(ctx);
}

// Stream-ordered device allocator: every buffer of the library comes from the device's default memory pool on the
// calling context's stream (set by the C-ABI entry guard), so repeated commits / proofs reuse the pooled memory
// instead of paying cudaMalloc/cudaFree (which also synchronise the device) for multi-GB buffers each time.
static thread_local cudaStream_t t_alloc_stream = nullptr;
void set_alloc_stream(cudaStream_t s) { t_alloc_stream = s; }
void dev_alloc(uint64_t** p, size_t n_u64) {
    cudaError_t e = cudaMallocAsync((void**)p, std::max<size_t>(n_u64, 1) * sizeof(uint64_t), t_alloc_stream);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        throw Error(OLA_ERR_OOM, "cudaMallocAsync: out of device memory (" + std::to_string(n_u64 * 8) + " bytes)");
    }
    OLA_CUDA(e);
}
void dev_free(void* p) {
    if (p) cudaFreeAsync(p, t_alloc_stream);
}

uint64_t* ctx_scratch(ola_ctx* ctx, size_t n_u64) {
    if (ctx->scratch_elems < n_u64) {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) ola::dev_free(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_elems = 0;
        dev_alloc(&ctx->scratch, n_u64);
        ctx->scratch_elems = n_u64;
    }
    return ctx->scratch;
}

void gather_rows(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t ncols, size_t first, size_t count,
                 uint64_t* out_host) {
    if (!count || !ncols) return;
    size_t total = count * ncols;
    uint64_t* tmp = ctx_scratch(ctx, total);
    {
        Launch lz(ctx, "gather_rows");
        gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_cols, col_stride, ncols, first, count, tmp);
    }
    check_launch("gather_rows_kernel");
    OLA_CUDA(cudaMemcpyAsync(out_host, tmp, total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
}

// coset LDE of coefficient columns [c0, c0 + cnt) of b (shift 7, blowup 2^rate_bits, cosets [coset_first, +coset_count))
// straight into leaf order (oracle.rs:101-129 + :84-85)
static void lde_columns(ola_ctx* ctx, ola_batch* b, size_t c0, size_t cnt, int coset_first, int coset_count) {
    const size_t n = (size_t)1 << b->log_n, L = n << b->shard_bits;
    ntt::FwdDesc d;
    d.src = b->d_coeffs + c0 * n;
    d.src_col_stride = n;
    d.dst = b->d_lde + c0 * L;
    d.dst_col_stride = L;
    d.dst_coset_stride = n;
    d.ncols = cnt;
    d.log_n = (int)b->log_n;
    d.coset_bits = (int)b->rate_bits;
    d.coset_first = coset_first;
    d.coset_count = coset_count;
    d.shift = gl::GEN;
    d.tag_strided = "lde_strided";
    d.tag_contig = "lde_contig";
    ntt::forward(ctx, d);
}
// MerkleTree::new_v2 over the LDE: leaf digests and the level reduction down to the cap
static void hash_commit(ola_ctx* ctx, ola_batch* b) {
    const size_t n = (size_t)1 << b->log_n, L = n << b->shard_bits;
    hasher::hash_rows_colmajor(ctx, b->d_lde, L, L, b->ncols, b->d_nodes + 4 * L);
    hasher::merkle_levels(ctx, b->d_nodes, L, (size_t)1 << b->local_cap_height());
}
static void finish_commit(ola_ctx* ctx, ola_batch* b, int coset_first, int coset_count) {
    lde_columns(ctx, b, 0, b->ncols, coset_first, coset_count);
    hash_commit(ctx, b);
}

// values on H (natural) -> coefficients (natural): forward network with inverse roots, x 1/n
// (PolynomialValues::ifft, polynomial/mod.rs:60-65).  `work` is clobbered (it may be src).
static void intt_columns(ola_ctx* ctx, const uint64_t* src, uint64_t* work, uint64_t* dst, size_t ncols, uint32_t log_n) {
    const size_t n = (size_t)1 << log_n;
    ntt::FwdDesc d;
    d.src = src;
    d.src_col_stride = n;
    d.work = work;
    d.work_col_stride = n;
    d.dst = dst;
    d.dst_col_stride = n;
    d.ncols = ncols;
    d.log_n = (int)log_n;
    d.inverse_roots = true;
    d.natural_output = true;
    d.apply_scale = true;
    d.scale = gl::inv(((uint64_t)1 << log_n) % gl::P);
    d.tag_strided = "intt_strided";
    d.tag_contig = "intt_contig";
    ntt::forward(ctx, d);
}

// Host columns -> coefficient columns in HBM, chunk by chunk through two staging buffers: the H2D copy of chunk k+1 (on
// the context's copy stream, one PCIe direction) overlaps the transforms of chunk k (context stream).  after_chunk(c0,
// cnt) is called (work enqueued on the context stream) once columns [c0, c0 + cnt) of d_coeffs hold coefficients.
template <typename F>
static void ingest_host_columns(ola_ctx* ctx, const uint64_t* h_cols, size_t ncols, uint32_t log_n, bool is_coeffs, uint64_t* d_coeffs,
                                F&& after_chunk) {
    const size_t n = (size_t)1 << log_n;
    // ~64 MB chunks, at least four of them when the batch is big enough to be worth pipelining
    size_t cc = std::max<size_t>(1, ((size_t)64 << 20) / (n * 8));
    if (ncols * n * 8 >= ((size_t)32 << 20)) cc = std::min(cc, std::max<size_t>(1, (ncols + 3) / 4));
    cc = std::min(cc, ncols);
    ensure_copy_stream(ctx);
    uint64_t* stage = nullptr;
    dev_alloc(&stage, 2 * cc * n);
    try {
        // the staging buffers exist (in stream order) once this event has fired; earlier work of the copy stream is done
        OLA_CUDA(cudaEventRecord(ctx->ev_free[0], ctx->stream));
        OLA_CUDA(cudaEventRecord(ctx->ev_free[1], ctx->stream));
        size_t k = 0;
        for (size_t c0 = 0; c0 < ncols; c0 += cc, ++k) {
            const size_t cnt = std::min(cc, ncols - c0);
            const int bsel = (int)(k & 1);
            uint64_t* sb = stage + (size_t)bsel * cc * n;
            OLA_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[bsel], 0));
            OLA_CUDA(cudaMemcpyAsync(sb, h_cols + c0 * n, cnt * n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            OLA_CUDA(cudaEventRecord(ctx->ev_ready[bsel], ctx->copy_stream));
            OLA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[bsel], 0));
            if (is_coeffs)
                canon_copy(ctx, d_coeffs + c0 * n, sb, cnt * n);
            else
                intt_columns(ctx, sb, sb, d_coeffs + c0 * n, cnt, log_n);
            OLA_CUDA(cudaEventRecord(ctx->ev_free[bsel], ctx->stream));
            after_chunk(c0, cnt);
        }
    } catch (...) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        dev_free(stage);
        throw;
    }
    dev_free(stage);  // stream-ordered after the last transform that reads it
}

void lde_batch(ola_ctx* ctx, const uint64_t* cols, bool on_device, size_t ncols, uint32_t log_n, bool is_coeffs, uint32_t rate_bits,
               uint64_t* d_coeffs, uint64_t* d_lde) {
    OLA_CHECK(cols != nullptr && ncols > 0, OLA_ERR_INVALID_ARG, "lde_batch: empty batch");
    OLA_CHECK(log_n + rate_bits <= 32, OLA_ERR_INVALID_ARG, "lde_batch: LDE size exceeds the field's two-adicity (2^32)");
    OLA_CHECK((1u << rate_bits) <= (unsigned)ntt::MAX_COSETS, OLA_ERR_INVALID_ARG, "lde_batch: rate_bits too large");
    const size_t n = (size_t)1 << log_n;
    ola_batch view;  // borrowed buffers, never released
    view.ncols = ncols;
    view.log_n = log_n;
    view.rate_bits = rate_bits;
    view.shard_bits = rate_bits;
    view.d_coeffs = d_coeffs;
    view.d_lde = d_lde;
    if (!on_device) {
        ingest_host_columns(ctx, cols, ncols, log_n, is_coeffs, d_coeffs, [&](size_t c0, size_t cnt) { lde_columns(ctx, &view, c0, cnt, 0, -1); });
        return;
    }
    if (is_coeffs) {
        canon_copy(ctx, d_coeffs, cols, ncols * n);
    } else {
        // in place when the caller passes its own value buffer as d_coeffs (PolynomialValues::ifft consumes its input)
        const bool multipass = ntt::plan_passes((int)log_n).size() > 1;
        uint64_t* work = (cols == d_coeffs) ? (multipass ? ctx_scratch(ctx, ncols * n) : nullptr) : d_lde;
        intt_columns(ctx, cols, work, d_coeffs, ncols, log_n);
    }
    lde_columns(ctx, &view, 0, ncols, 0, -1);
}

ola_batch* batch_commit(ola_ctx* ctx, const uint64_t* cols, bool on_device, size_t ncols, uint32_t log_n, bool is_coeffs,
                        uint32_t rate_bits, uint32_t cap_height, int coset_first, int coset_count) {
    OLA_CHECK(cols != nullptr && ncols > 0, OLA_ERR_INVALID_ARG, "commit: empty batch");
    OLA_CHECK(log_n + rate_bits <= 32, OLA_ERR_INVALID_ARG, "commit: LDE size exceeds the field's two-adicity (2^32)");
    OLA_CHECK(cap_height <= log_n + rate_bits, OLA_ERR_INVALID_ARG, "commit: cap height should be at most log2(leaves)");
    OLA_CHECK((1u << rate_bits) <= (unsigned)ntt::MAX_COSETS, OLA_ERR_INVALID_ARG, "commit: rate_bits too large");
    if (coset_count < 0) {
        coset_first = 0;
        coset_count = 1 << rate_bits;
    }
    uint32_t shard_bits = 0;
    while ((1 << shard_bits) < coset_count) ++shard_bits;
    OLA_CHECK((1 << shard_bits) == coset_count && coset_count <= (1 << rate_bits) && coset_first >= 0 && coset_first % coset_count == 0 &&
                  coset_first + coset_count <= (1 << rate_bits),
              OLA_ERR_INVALID_ARG, "commit: a coset shard is an aligned power-of-two range of the 2^rate_bits cosets");
    OLA_CHECK(cap_height >= rate_bits - shard_bits, OLA_ERR_INVALID_ARG, "commit: the cap must be at least as fine as the coset partition");
    std::unique_ptr<ola_batch, void (*)(ola_batch*)> b(new ola_batch(), [](ola_batch* p) {
        batch_release(p);
        delete p;
    });
    b->ncols = ncols;
    b->log_n = log_n;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->shard_bits = shard_bits;
    b->coset_first = (uint32_t)coset_first;
    const size_t n = (size_t)1 << log_n, L = n << shard_bits;
    dev_alloc(&b->d_coeffs, ncols * n);
    dev_alloc(&b->d_lde, ncols * L);
    dev_alloc(&b->d_nodes, 2 * L * 4);

    if (!on_device) {
        // pipelined upload: each chunk's LDE follows its coefficients while the next chunk is still in flight
        ola_batch* bp = b.get();
        ingest_host_columns(ctx, cols, ncols, log_n, is_coeffs, b->d_coeffs,
                            [&](size_t c0, size_t cnt) { lde_columns(ctx, bp, c0, cnt, coset_first, coset_count); });
        hash_commit(ctx, bp);
        return b.release();
    }
    if (is_coeffs) {
        canon_copy(ctx, b->d_coeffs, cols, ncols * n);
    } else {
        // never clobber a caller-owned device buffer: the (still unused) LDE buffer is the work area when it is big
        // enough, a second scratch region otherwise
        uint64_t* tmp_work = nullptr;
        uint64_t* work = b->d_lde;
        if (shard_bits < 1) {
            dev_alloc(&tmp_work, ncols * n);
            work = tmp_work;
        }
        try {
            intt_columns(ctx, cols, work, b->d_coeffs, ncols, log_n);
        } catch (...) {
            if (tmp_work) ola::dev_free(tmp_work);
            throw;
        }
        if (tmp_work) {
            OLA_CUDA(cudaStreamSynchronize(ctx->stream));
            ola::dev_free(tmp_work);
        }
    }
    finish_commit(ctx, b.get(), coset_first, coset_count);
    return b.release();
}

// Column-sharded coefficients + coset-sharded commitment for the multi-GPU prover (ola_set_comm): rank r turns columns
// [r*per, (r+1)*per) of the (replicated, device-resident) values into coefficients and one all-gather replicates them
// (openings and the FRI composition read every coefficient column on every rank); the LDE, leaf hashing and subtree
// reduction then cover this rank's cosets only.
ola_batch* batch_commit_dist(ola_ctx* ctx, const uint64_t* d_cols, size_t ncols, uint32_t log_n, bool is_coeffs, uint32_t rate_bits,
                             uint32_t cap_height) {
    OLA_CHECK(ctx->world > 1 && ((1 << rate_bits) % ctx->world) == 0, OLA_ERR_INTERNAL, "batch_commit_dist needs a communicator");
    OLA_CHECK(d_cols != nullptr && ncols > 0, OLA_ERR_INVALID_ARG, "commit: empty batch");
    OLA_CHECK(log_n + rate_bits <= 32, OLA_ERR_INVALID_ARG, "commit: LDE size exceeds the field's two-adicity (2^32)");
    const int coset_count = (1 << rate_bits) / ctx->world, coset_first = ctx->rank * coset_count;
    uint32_t shard_bits = 0;
    while ((1 << shard_bits) < coset_count) ++shard_bits;
    OLA_CHECK(cap_height >= rate_bits - shard_bits && cap_height <= log_n + rate_bits, OLA_ERR_INVALID_ARG, "commit: cap height out of range");
    std::unique_ptr<ola_batch, void (*)(ola_batch*)> b(new ola_batch(), [](ola_batch* p) {
        batch_release(p);
        delete p;
    });
    b->ncols = ncols;
    b->log_n = log_n;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->shard_bits = shard_bits;
    b->coset_first = (uint32_t)coset_first;
    const size_t n = (size_t)1 << log_n, L = n << shard_bits;
    const size_t per = (ncols + ctx->world - 1) / ctx->world;  // columns per rank (the last ranks may own fewer / none)
    dev_alloc(&b->d_coeffs, per * ctx->world * n);              // padded: the all-gather needs equal contributions
    dev_alloc(&b->d_lde, ncols * L);
    dev_alloc(&b->d_nodes, 2 * L * 4);
    if (is_coeffs) {
        canon_copy(ctx, b->d_coeffs, d_cols, ncols * n);
    } else {
        const size_t lo = std::min(ncols, (size_t)ctx->rank * per), hi = std::min(ncols, lo + per);
        uint64_t* tmp = nullptr;
        dev_alloc(&tmp, 2 * per * n);  // [send | work]
        try {
            if (hi - lo < per) OLA_CUDA(cudaMemsetAsync(tmp, 0, per * n * 8, ctx->stream));
            if (hi > lo) {
                ntt::FwdDesc d;
                d.src = d_cols + lo * n;
                d.src_col_stride = n;
                d.work = tmp + per * n;
                d.work_col_stride = n;
                d.dst = tmp;
                d.dst_col_stride = n;
                d.ncols = hi - lo;
                d.log_n = (int)log_n;
                d.inverse_roots = true;
                d.natural_output = true;
                d.apply_scale = true;
                d.scale = gl::inv(((uint64_t)1 << log_n) % gl::P);
                d.tag_strided = "intt_strided";
                d.tag_contig = "intt_contig";
                ntt::forward(ctx, d);
            }
            comm_allgather(ctx, tmp, b->d_coeffs, per * n * 8);
            OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            ola::dev_free(tmp);
            throw;
        }
        ola::dev_free(tmp);
    }
    finish_commit(ctx, b.get(), coset_first, coset_count);
    return b.release();
}

void batch_release(ola_batch* b) {
    if (!b) return;
    if (b->d_coeffs) ola::dev_free(b->d_coeffs);
    if (b->d_lde) ola::dev_free(b->d_lde);
    if (b->d_nodes) ola::dev_free(b->d_nodes);
    b->d_coeffs = b->d_lde = b->d_nodes = nullptr;
}

void batch_get_cap(ola_ctx* ctx, const ola_batch* b, uint64_t* cap_host) {
    const size_t L = (size_t)1 << b->leaf_bits(), ncap = (size_t)1 << b->local_cap_height();
    // cap = nodes[2^h .. 2^(h+1)); when the tree is all cap these are the leaf digests (mod.rs:216-225).
    // For a coset shard these are the global cap entries [coset_first, +2^shard_bits) * 2^(cap_height - rate_bits).
    OLA_CUDA(cudaMemcpyAsync(cap_host, b->d_nodes + 4 * ncap, ncap * 32, cudaMemcpyDeviceToHost, ctx->stream));
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    (void)L;
}

void batch_get_leaves(ola_ctx* ctx, const ola_batch* b, size_t first, size_t count, uint64_t* out_host) {
    const size_t L = (size_t)1 << b->leaf_bits();
    OLA_CHECK(first + count <= L, OLA_ERR_INVALID_ARG, "get_leaves: leaf index out of range");
    gather_rows(ctx, b->d_lde, L, b->ncols, first, count, out_host);
}

int batch_prove_leaf(ola_ctx* ctx, const ola_batch* b, size_t leaf, uint64_t* sib_host) {
    const uint32_t lg = b->leaf_bits();
    const size_t L = (size_t)1 << lg;
    OLA_CHECK(leaf < L, OLA_ERR_INVALID_ARG, "prove_leaf: leaf index out of range");
    const int nsib = (int)lg - (int)b->local_cap_height();
    if (nsib <= 0) return 0;
    uint64_t* tmp = nullptr;
    dev_alloc(&tmp, (size_t)nsib * 4);
    gather_path_kernel<<<(nsib * 4 + 63) / 64, 64, 0, ctx->stream>>>(b->d_nodes, L, leaf, nsib, tmp);
    count_launch(ctx);
    cudaError_t e = cudaMemcpyAsync(sib_host, tmp, (size_t)nsib * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ola::dev_free(tmp);
    OLA_CUDA(e);
    return nsib;
}

}  // namespace ola
