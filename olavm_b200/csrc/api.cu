// extern "C" surface of libola_gpu (declared in include/ola_gpu.h).  Every entry point converts C++
// exceptions into error codes; nothing propagates across the ABI.
#include <algorithm>
#include <cstring>
#include <memory>

#include "batch.h"
#include "common.h"
#include "gl.cuh"
#include "ntt.h"
#include "hasher.h"
#include "poseidon.cuh"
#include "stark.h"
#include "verify.h"

namespace ola {
namespace nccl {
void release(ola_ctx* ctx);  // nccl_comm.cu
}
namespace generation {
void poseidon_trace(ola_ctx* ctx, const uint64_t* d_inputs, const uint64_t* d_filters, size_t nrows, uint32_t log_n, uint64_t* d_out);
}
}  // namespace ola

namespace {

template <typename F>
int guarded(ola_ctx* ctx, F&& f) {
    try {
        if (ctx) {
            // a context is bound to its device: callers may come from any host thread / after switching devices
            int cur = -1;
            if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device) OLA_CUDA(cudaSetDevice(ctx->device));
            ola::set_alloc_stream(ctx->stream);
        }
        f();
        return OLA_OK;
    } catch (const ola::Error& e) {
        if (ctx) ctx->last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->last_error = "host allocation failed";
        return OLA_ERR_OOM;
    } catch (const std::exception& e) {
        if (ctx) ctx->last_error = e.what();
        return OLA_ERR_INTERNAL;
    } catch (...) {
        if (ctx) ctx->last_error = "unknown error";
        return OLA_ERR_INTERNAL;
    }
}

// RAII device scratch
struct DevBuf {
    uint64_t* p = nullptr;
    explicit DevBuf(size_t n) { ola::dev_alloc(&p, n); }
    ~DevBuf() {
        if (p) ola::dev_free(p);
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

thread_local std::string g_init_error;

void to_device(ola_ctx* ctx, uint64_t* dst, const uint64_t* src, size_t n) {
    OLA_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyHostToDevice, ctx->stream));
}
void to_host(ola_ctx* ctx, uint64_t* dst, const uint64_t* src, size_t n) {
    OLA_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
}

// natural -> natural transform of [ncols][n] (forward or inverse roots), data on device
void ntt_natural(ola_ctx* ctx, uint64_t* d_data, size_t ncols, uint32_t log_n, bool inverse) {
    const size_t n = (size_t)1 << log_n;
    const bool multipass = ola::ntt::plan_passes((int)log_n).size() > 1;
    ola::ntt::FwdDesc d;
    d.src = d_data;
    d.src_col_stride = n;
    d.work = multipass ? ola::ctx_scratch(ctx, ncols * n) : nullptr;
    d.work_col_stride = n;
    d.dst = d_data;
    d.dst_col_stride = n;
    d.ncols = ncols;
    d.log_n = (int)log_n;
    d.inverse_roots = inverse;
    d.natural_output = true;
    d.apply_scale = inverse;
    d.scale = inverse ? gl::inv(((uint64_t)1 << log_n) % gl::P) : 1;
    d.tag_strided = inverse ? "intt_strided" : "ntt_strided";
    d.tag_contig = inverse ? "intt_contig" : "ntt_contig";
    ola::ntt::forward(ctx, d);
}

}  // namespace

extern "C" {

int ola_gpu_init(int device, ola_ctx** out) {
    if (!out) return OLA_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return OLA_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OLA_ERR_NO_DEVICE;
    if (prop.major != 10) return OLA_ERR_NO_DEVICE;  // built for sm_100a only; no fallback path exists
    ola_ctx* ctx = new (std::nothrow) ola_ctx();
    if (!ctx) return OLA_ERR_OOM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    int rc = guarded(ctx, [&] {
        OLA_CUDA(cudaSetDevice(device));
        OLA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ola::set_alloc_stream(ctx->stream);
        cudaMemPool_t pool;
        OLA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;  // keep freed blocks in the pool until the context is destroyed
        OLA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        ola::poseidon::init_constants();
        ola::ntt::init_twiddles(ctx);
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    });
    if (rc != OLA_OK) {
        ola_gpu_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return OLA_OK;
}

void ola_gpu_destroy(ola_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ola::nccl::release(ctx);
    ola::ntt::free_twiddles(ctx);
    ola::set_alloc_stream(ctx->stream);
    if (ctx->scratch) ola::dev_free(ctx->scratch);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 2; ++i) {
            if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
            if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
        }
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        cudaStreamDestroy(ctx->stream);
    }
    ola::set_alloc_stream(nullptr);
    delete ctx;
}

const char* ola_gpu_last_error(const ola_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int ola_gpu_sync(ola_ctx* ctx) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { OLA_CUDA(cudaStreamSynchronize(ctx->stream)); });
}
uint64_t ola_gpu_kernel_launches(const ola_ctx* ctx) { return ctx ? ctx->kernel_launches : 0; }
void* ola_gpu_stream(ola_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int ola_profile_begin(ola_ctx* ctx) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->prof_totals.clear();
        ctx->profiling = true;
    });
}
int ola_profile_end(ola_ctx* ctx, char* json_out, size_t cap) {
    if (!ctx || (!json_out && cap)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->profiling = false;
        for (auto& r : ctx->prof_pending) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.start, r.stop);
            auto& t = ctx->prof_totals[r.name];
            t.first += ms;
            t.second += 1;
            cudaEventDestroy(r.start);
            cudaEventDestroy(r.stop);
        }
        ctx->prof_pending.clear();
        std::string js = "{";
        bool first = true;
        for (auto& kv : ctx->prof_totals) {
            char buf[256];
            snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.6f, \"launches\": %llu}", first ? "" : ", ", kv.first.c_str(),
                     kv.second.first, (unsigned long long)kv.second.second);
            js += buf;
            first = false;
        }
        js += "}";
        OLA_CHECK(js.size() + 1 <= cap, OLA_ERR_INVALID_ARG, "profile buffer too small");
        memcpy(json_out, js.c_str(), js.size() + 1);
    });
}

int ola_dev_alloc(ola_ctx* ctx, size_t n_u64, uint64_t** dptr) {
    if (!ctx || !dptr) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        ola::dev_alloc(dptr, n_u64);
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));  // usable from any stream once this returns
    });
}
int ola_dev_free(ola_ctx* ctx, uint64_t* dptr) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        if (dptr) ola::dev_free(dptr);
    });
}
int ola_dev_upload(ola_ctx* ctx, uint64_t* dst_dev, const uint64_t* src_host, size_t n_u64) {
    if (!ctx || (!dst_dev && n_u64) || (!src_host && n_u64)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        to_device(ctx, dst_dev, src_host, n_u64);
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int ola_dev_download(ola_ctx* ctx, uint64_t* dst_host, const uint64_t* src_dev, size_t n_u64) {
    if (!ctx || (!dst_host && n_u64) || (!src_dev && n_u64)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { to_host(ctx, dst_host, src_dev, n_u64); });
}

int ola_dev_copy(ola_ctx* ctx, uint64_t* dst_dev, const uint64_t* src_dev, size_t n_u64) {
    if (!ctx || (!dst_dev && n_u64) || (!src_dev && n_u64)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { OLA_CUDA(cudaMemcpyAsync(dst_dev, src_dev, n_u64 * 8, cudaMemcpyDeviceToDevice, ctx->stream)); });
}
int ola_dev_gather_rows(ola_ctx* ctx, const uint64_t* cols_dev, size_t col_stride, size_t ncols, size_t first_row,
                        size_t count, uint64_t* out_host) {
    if (!ctx || !cols_dev || (!out_host && count)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { ola::gather_rows(ctx, cols_dev, col_stride, ncols, first_row, count, out_host); });
}

static int ntt_entry(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n, bool inverse) {
    if (!ctx || !data) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(log_n <= 32, OLA_ERR_INVALID_ARG, "multiplicative subgroup of that size does not exist (two-adicity 32)");
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ntt_natural(ctx, data, ncols, log_n, inverse);
        } else {
            DevBuf d(ncols * n);
            to_device(ctx, d.p, data, ncols * n);
            ntt_natural(ctx, d.p, ncols, log_n, inverse);
            to_host(ctx, data, d.p, ncols * n);
        }
    });
}
int ola_ntt_forward(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n) {
    return ntt_entry(ctx, data, on_device, ncols, log_n, false);
}
int ola_ntt_inverse(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n) {
    return ntt_entry(ctx, data, on_device, ncols, log_n, true);
}

int ola_coset_lde(ola_ctx* ctx, const uint64_t* coeffs, uint64_t* out, int on_device, size_t ncols, uint32_t log_n,
                  uint32_t rate_bits, uint64_t shift, int natural_order) {
    if (!ctx || !coeffs || !out) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(log_n + rate_bits <= 32, OLA_ERR_INVALID_ARG,
                  "multiplicative subgroup of that size does not exist (two-adicity 32)");
        OLA_CHECK((1u << rate_bits) <= (unsigned)ola::ntt::MAX_COSETS, OLA_ERR_INVALID_ARG, "blowup factor too large");
        OLA_CHECK(gl::canon(shift) != 0, OLA_ERR_INVALID_ARG, "domain offset cannot be zero");
        const size_t n = (size_t)1 << log_n, L = n << rate_bits;
        std::unique_ptr<DevBuf> d_in, d_out, d_work;
        const uint64_t* src = coeffs;
        uint64_t* dst = out;
        if (!on_device) {
            d_in.reset(new DevBuf(ncols * n));
            d_out.reset(new DevBuf(ncols * L));
            to_device(ctx, d_in->p, coeffs, ncols * n);
            src = d_in->p;
            dst = d_out->p;
        }
        ola::ntt::FwdDesc d;
        d.src = src;
        d.src_col_stride = n;
        d.dst = dst;
        d.dst_col_stride = L;
        d.dst_coset_stride = n;
        d.ncols = ncols;
        d.log_n = (int)log_n;
        d.coset_bits = (int)rate_bits;
        d.shift = gl::canon(shift);
        d.natural_output = natural_order != 0;
        d.tag_strided = "lde_strided";
        d.tag_contig = "lde_contig";
        if (natural_order && ola::ntt::plan_passes((int)log_n).size() > 1) {
            d_work.reset(new DevBuf(ncols * L));
            d.work = d_work->p;
            d.work_col_stride = L;
            d.work_coset_stride = n;
        }
        ola::ntt::forward(ctx, d);
        if (!on_device)
            to_host(ctx, out, dst, ncols * L);
        else if (d_work)
            OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int ola_lde_batch(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs, uint32_t rate_bits,
                  uint64_t* coeffs_out_dev, uint64_t* lde_out_dev) {
    if (!ctx || !cols || !coeffs_out_dev || !lde_out_dev) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { ola::lde_batch(ctx, cols, on_device != 0, ncols, log_n, is_coeffs != 0, rate_bits, coeffs_out_dev, lde_out_dev); });
}

int ola_coset_intt(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n, uint64_t shift) {
    if (!ctx || !data) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(log_n <= 32, OLA_ERR_INVALID_ARG, "multiplicative subgroup of that size does not exist (two-adicity 32)");
        OLA_CHECK(gl::canon(shift) != 0, OLA_ERR_INVALID_ARG, "domain offset cannot be zero");
        const size_t n = (size_t)1 << log_n;
        std::unique_ptr<DevBuf> d;
        uint64_t* p = data;
        if (!on_device) {
            d.reset(new DevBuf(ncols * n));
            to_device(ctx, d->p, data, ncols * n);
            p = d->p;
        }
        ntt_natural(ctx, p, ncols, log_n, true);
        // coefficients of p(shift * x) -> coefficients of p: c_j *= shift^-j   (cfft/serial.rs:73-78)
        ola::ntt::scale_powers(ctx, p, n, ncols, n, 1, gl::inv(gl::canon(shift)));
        if (!on_device) to_host(ctx, data, p, ncols * n);
    });
}

int ola_poseidon_permute(ola_ctx* ctx, uint64_t* states, int on_device, size_t nstates) {
    if (!ctx || (!states && nstates)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        if (on_device) {
            ola::poseidon::permute_states(ctx, states, nstates);
        } else {
            DevBuf d(nstates * 12);
            to_device(ctx, d.p, states, nstates * 12);
            ola::poseidon::permute_states(ctx, d.p, nstates);
            to_host(ctx, states, d.p, nstates * 12);
        }
    });
}

int ola_hash_rows(ola_ctx* ctx, const uint64_t* rows, uint64_t* digests, int on_device, size_t nrows, size_t ncols) {
    if (!ctx || !digests || (!rows && nrows * ncols)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        if (on_device) {
            ola::hasher::hash_rows_rowmajor(ctx, rows, nrows, ncols, digests);
        } else {
            DevBuf d(nrows * ncols), o(nrows * 4);
            to_device(ctx, d.p, rows, nrows * ncols);
            ola::hasher::hash_rows_rowmajor(ctx, d.p, nrows, ncols, o.p);
            to_host(ctx, digests, o.p, nrows * 4);
        }
    });
}

int ola_merkle_rows(ola_ctx* ctx, const uint64_t* rows, int on_device, size_t nrows, size_t ncols, uint32_t cap_height,
                    uint64_t* cap_out_host, uint64_t* nodes_out_host) {
    if (!ctx || !rows || !cap_out_host) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(nrows && (nrows & (nrows - 1)) == 0, OLA_ERR_INVALID_ARG, "number of leaves must be a power of two");
        OLA_CHECK(((size_t)1 << cap_height) <= nrows, OLA_ERR_INVALID_ARG, "cap height should be at most log2(leaves.len())");
        std::unique_ptr<DevBuf> d;
        const uint64_t* src = rows;
        if (!on_device) {
            d.reset(new DevBuf(nrows * ncols));
            to_device(ctx, d->p, rows, nrows * ncols);
            src = d->p;
        }
        DevBuf nodes(2 * nrows * 4);
        OLA_CUDA(cudaMemsetAsync(nodes.p, 0, 2 * nrows * 32, ctx->stream));
        ola::hasher::hash_rows_rowmajor(ctx, src, nrows, ncols, nodes.p + 4 * nrows);
        ola::hasher::merkle_levels(ctx, nodes.p, nrows, nodes_out_host ? 1 : ((size_t)1 << cap_height));
        const size_t ncap = (size_t)1 << cap_height;
        to_host(ctx, cap_out_host, nodes.p + 4 * ncap, ncap * 4);
        if (nodes_out_host) to_host(ctx, nodes_out_host, nodes.p, 2 * nrows * 4);
    });
}

static int commit_impl(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs, uint32_t rate_bits,
                       uint32_t cap_height, int coset_first, int coset_count, ola_batch** out, uint64_t* cap_out_host) {
    if (!ctx || !out) return OLA_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        ola_batch* b = ola::batch_commit(ctx, cols, on_device != 0, ncols, log_n, is_coeffs != 0, rate_bits, cap_height, coset_first, coset_count);
        try {
            if (cap_out_host)
                ola::batch_get_cap(ctx, b, cap_out_host);
            else
                OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            ola::batch_release(b);
            delete b;
            throw;
        }
        *out = b;
    });
}
int ola_commit(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs, uint32_t rate_bits,
               uint32_t cap_height, ola_batch** out, uint64_t* cap_out_host) {
    return commit_impl(ctx, cols, on_device, ncols, log_n, is_coeffs, rate_bits, cap_height, 0, -1, out, cap_out_host);
}
int ola_commit_shard(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs, uint32_t rate_bits,
                     uint32_t cap_height, uint32_t coset_first, uint32_t coset_count, ola_batch** out, uint64_t* cap_slots_out_host) {
    if (coset_count == 0 || coset_count > 64 || coset_first > 64) return OLA_ERR_INVALID_ARG;
    return commit_impl(ctx, cols, on_device, ncols, log_n, is_coeffs, rate_bits, cap_height, (int)coset_first, (int)coset_count, out,
                       cap_slots_out_host);
}

int ola_prove(ola_ctx* ctx, const int* table_ids, uint32_t ntables, const uint64_t* const* traces, int on_device, const uint32_t* log_ns,
              const uint64_t* compress_challenges, int check_quotient_degree, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!ctx || !table_ids || !traces || !log_ns || !proof_len || (!proof_out && proof_cap)) return OLA_ERR_INVALID_ARG;
    *proof_len = 0;
    return guarded(ctx, [&] {
        ola::stark::Config cfg;
        cfg.check_quotient_degree = check_quotient_degree != 0;
        std::vector<int> ids(table_ids, table_ids + ntables);
        std::vector<const uint64_t*> tr(traces, traces + ntables);
        std::vector<uint32_t> lg(log_ns, log_ns + ntables);
        std::vector<uint64_t> cc;
        if (compress_challenges) cc.assign(compress_challenges, compress_challenges + ntables);
        std::vector<uint8_t> bytes = ola::stark::prove_all(ctx, ids, tr, on_device != 0, lg, cc, cfg);
        *proof_len = bytes.size();
        OLA_CHECK(bytes.size() <= proof_cap, OLA_ERR_INVALID_ARG, "proof buffer too small (needed size returned in proof_len)");
        memcpy(proof_out, bytes.data(), bytes.size());
    });
}
// ---- ola_prove with the transcript kept by the caller (ola_prove_session_*) ----
}  // extern "C"
#include <thread>
struct ola_session {
    ola_ctx* ctx = nullptr;
    ola::stark::TranscriptHost host;
    std::thread worker;
    std::vector<uint8_t> proof;
    int rc = OLA_OK;
    std::string error;
    bool finished = false;  // the worker has published DONE / FAILED
    std::vector<int> ids;
    std::vector<const uint64_t*> traces;
    std::vector<uint32_t> log_ns;
    std::vector<uint64_t> cc;
    bool on_device = false;
    ola::stark::Config cfg;
};
extern "C" {
int ola_prove_session_begin(ola_ctx* ctx, const int* table_ids, uint32_t ntables, const uint64_t* const* traces, int on_device, const uint32_t* log_ns,
                            const uint64_t* compress_challenges, int check_quotient_degree, ola_session** out) {
    if (!ctx || !table_ids || !traces || !log_ns || !out) return OLA_ERR_INVALID_ARG;
    if (ctx->world != 1) {
        ctx->last_error = "ola_prove_session_*: single-GPU contexts only";
        return OLA_ERR_INVALID_ARG;
    }
    ola_session* s = nullptr;
    try {
        s = new ola_session();
    } catch (...) {
        return OLA_ERR_OOM;
    }
    s->ctx = ctx;
    s->ids.assign(table_ids, table_ids + ntables);
    s->traces.assign(traces, traces + ntables);
    s->log_ns.assign(log_ns, log_ns + ntables);
    if (compress_challenges) s->cc.assign(compress_challenges, compress_challenges + ntables);
    s->on_device = on_device != 0;
    s->cfg.check_quotient_degree = check_quotient_degree != 0;
    s->worker = std::thread([s] {
        using TH = ola::stark::TranscriptHost;
        TH::Kind last = TH::DONE;
        s->rc = guarded(s->ctx, [&] { s->proof = ola::stark::prove_all(s->ctx, s->ids, s->traces, s->on_device, s->log_ns, s->cc, s->cfg, &s->host); });
        if (s->rc != OLA_OK) {
            s->error = s->ctx->last_error;
            last = TH::FAILED;
        }
        std::lock_guard<std::mutex> lk(s->host.m);
        s->host.pending = last;
        s->finished = true;
        s->host.cv.notify_all();
    });
    *out = s;
    return OLA_OK;
}
int ola_prove_session_next(ola_session* s, ola_transcript_event* ev) {
    if (!s || !ev) return OLA_ERR_INVALID_ARG;
    using TH = ola::stark::TranscriptHost;
    std::unique_lock<std::mutex> lk(s->host.m);
    // the previous OBSERVE / COMPACT event is consumed by asking for the next one
    if (!s->finished && s->host.pending != TH::NONE && s->host.pending != TH::CHALLENGE && s->host.delivered) {
        s->host.consumed = true;
        s->host.delivered = false;
        s->host.pending = TH::NONE;
        s->host.cv.notify_all();
    }
    if (!s->finished && s->host.pending == TH::CHALLENGE && s->host.delivered) return OLA_ERR_INVALID_ARG;  // supply the challenges first
    s->host.cv.wait(lk, [&] { return s->finished || (s->host.pending != TH::NONE && !s->host.delivered); });
    ev->kind = (int)s->host.pending;
    ev->stage = s->host.stage;
    ev->table = s->host.table;
    ev->elems = s->host.pending == TH::OBSERVE ? s->host.elems.data() : nullptr;
    ev->count = s->host.pending == TH::OBSERVE ? s->host.elems.size() : (s->host.pending == TH::CHALLENGE ? s->host.want : 0);
    if (!s->finished) s->host.delivered = true;
    return s->host.pending == TH::FAILED ? s->rc : OLA_OK;
}
int ola_prove_session_supply(ola_session* s, const uint64_t* challenges, size_t count) {
    if (!s || (!challenges && count)) return OLA_ERR_INVALID_ARG;
    using TH = ola::stark::TranscriptHost;
    std::lock_guard<std::mutex> lk(s->host.m);
    if (s->finished || s->host.pending != TH::CHALLENGE || !s->host.delivered || count != s->host.want) return OLA_ERR_INVALID_ARG;
    s->host.elems.assign(challenges, challenges + count);
    s->host.supplied = true;
    s->host.delivered = false;
    s->host.pending = TH::NONE;
    s->host.cv.notify_all();
    return OLA_OK;
}
int ola_prove_session_finish(ola_session* s, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!s) return OLA_ERR_INVALID_ARG;
    {
        std::lock_guard<std::mutex> lk(s->host.m);
        if (!s->finished) {  // abandoned: unblock the prover, which unwinds
            s->host.abort = true;
            s->host.cv.notify_all();
        }
    }
    if (s->worker.joinable()) s->worker.join();
    int rc = s->rc;
    if (rc == OLA_OK && s->host.abort) rc = OLA_ERR_INVALID_ARG;
    if (proof_len) *proof_len = s->proof.size();
    if (rc == OLA_OK) {
        if (!proof_out || s->proof.size() > proof_cap)
            rc = OLA_ERR_INVALID_ARG;
        else
            memcpy(proof_out, s->proof.data(), s->proof.size());
    } else if (!s->error.empty()) {
        s->ctx->last_error = s->error;
    }
    delete s;
    return rc;
}

int ola_set_comm(ola_ctx* ctx, int rank, int world, ola_allgather_fn allgather, ola_allreduce_u64_fn allreduce_sum, void* user) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(world >= 1 && world <= 8 && (8 % world) == 0 && rank >= 0 && rank < world, OLA_ERR_INVALID_ARG,
                  "set_comm: world must divide the blowup (8) and 0 <= rank < world");
        OLA_CHECK(world == 1 || (allgather && allreduce_sum), OLA_ERR_INVALID_ARG, "set_comm: both collectives are required");
        ctx->rank = rank;
        ctx->world = world;
        ctx->comm_allgather = allgather;
        ctx->comm_allreduce = allreduce_sum;
        ctx->comm_user = user;
    });
}
int ola_set_hasher(ola_ctx* ctx, int hasher) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        OLA_CHECK(hasher == OLA_HASH_POSEIDON || hasher == OLA_HASH_BLAKE3, OLA_ERR_INVALID_ARG, "unknown hasher id");
        ctx->hasher = hasher;
    });
}
int ola_get_hasher(const ola_ctx* ctx) { return ctx ? ctx->hasher : OLA_ERR_INVALID_ARG; }
static int verify_impl(int hasher, const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err, size_t errcap, bool full_system) {
    auto put = [&](const std::string& m) {
        if (err && errcap) snprintf(err, errcap, "%s", m.c_str());
    };
    if (!table_ids || !proof) {
        put("null argument");
        return OLA_ERR_INVALID_ARG;
    }
    if (hasher != OLA_HASH_POSEIDON && hasher != OLA_HASH_BLAKE3) {
        put("unknown hasher id");
        return OLA_ERR_INVALID_ARG;
    }
    for (uint32_t i = 0; i < ntables; ++i)
        if (table_ids[i] < 0 || table_ids[i] >= ola::stark::T_NUM || (i > 0 && table_ids[i] <= table_ids[i - 1])) {
            put("table ids must be distinct ids of the Table enum in ascending order");
            return OLA_ERR_INVALID_ARG;
        }
    if (full_system && ntables != (uint32_t)ola::stark::T_NUM) {
        // verify_proof (verifier.rs:32-212) is fixed at the 12 tables and checks every cross-table lookup; a subset would
        // silently skip the lookups whose other side is missing
        put("verify_proof covers the full 12-table system (ids 0..11); ola_verify_subsystem_cfg verifies a subsystem, with weaker guarantees");
        return OLA_ERR_INVALID_ARG;
    }
    try {
        const std::string e = ola::stark::verify::verify_all(proof, proof_len, std::vector<int>(table_ids, table_ids + ntables), hasher);
        put(e);
        return e.empty() ? OLA_OK : OLA_ERR_INVALID_ARG;
    } catch (const std::exception& ex) {
        put(ex.what());
        return OLA_ERR_INVALID_ARG;
    } catch (...) {
        put("unknown error");
        return OLA_ERR_INTERNAL;
    }
}
int ola_verify(const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err, size_t errcap) {
    return verify_impl(OLA_HASH_POSEIDON, table_ids, ntables, proof, proof_len, err, errcap, true);
}
int ola_verify_cfg(int hasher, const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err, size_t errcap) {
    return verify_impl(hasher, table_ids, ntables, proof, proof_len, err, errcap, true);
}
int ola_verify_subsystem_cfg(int hasher, const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err, size_t errcap) {
    return verify_impl(hasher, table_ids, ntables, proof, proof_len, err, errcap, false);
}
// ---- trace-generation tail (generation.cu) ----
int ola_generate_poseidon_trace(ola_ctx* ctx, const uint64_t* inputs, const uint64_t* filters, size_t nrows, uint32_t log_n, uint64_t* out,
                                int on_device) {
    if (!ctx || (!inputs && nrows) || !out || log_n > 28 || nrows > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::generation::poseidon_trace(ctx, inputs, filters, nrows, log_n, out);
            return;
        }
        DevBuf d_in(std::max<size_t>(nrows * 12, 1)), d_f(std::max<size_t>(nrows * 4, 1)), d_out(134 * n);
        if (nrows) to_device(ctx, d_in.p, inputs, nrows * 12);
        if (nrows && filters) to_device(ctx, d_f.p, filters, nrows * 4);
        ola::generation::poseidon_trace(ctx, d_in.p, filters ? d_f.p : nullptr, nrows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 134 * n);
    });
}
int ola_permuted_cols(ola_ctx* ctx, const uint64_t* inputs, const uint64_t* table, size_t n, uint64_t* permuted_inputs, uint64_t* permuted_table,
                      int on_device) {
    if (!ctx || !inputs || !table || !permuted_inputs || !permuted_table) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        if (on_device) {
            ola::lookup::permuted_cols(ctx, inputs, table, n, permuted_inputs, permuted_table);
            return;
        }
        OLA_CHECK(n >= 2 && (n & (n - 1)) == 0 && n <= ((size_t)1 << 24), OLA_ERR_INVALID_ARG, "permuted_cols: the column length must be a power of two in [2, 2^24]");
        DevBuf d_in(n), d_tab(n), d_pi(n), d_pt(n);
        to_device(ctx, d_in.p, inputs, n);
        to_device(ctx, d_tab.p, table, n);
        ola::lookup::permuted_cols(ctx, d_in.p, d_tab.p, n, d_pi.p, d_pt.p);
        to_host(ctx, permuted_inputs, d_pi.p, n);
        to_host(ctx, permuted_table, d_pt.p, n);
    });
}
int ola_generate_rangecheck_trace(ola_ctx* ctx, const uint64_t* vals, const uint64_t* kinds, size_t nrows, uint32_t log_n, uint64_t* out,
                                  int on_device) {
    if (!ctx || ((!vals || !kinds) && nrows) || !out || log_n < 16 || log_n > 24 || nrows > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::lookup::rangecheck_trace(ctx, vals, kinds, nrows, log_n, out);
            return;
        }
        for (size_t i = 0; i < nrows; ++i) OLA_CHECK(kinds[i] <= 3, OLA_ERR_INVALID_ARG, "RangeCheck row kind must be 0 (cpu), 1 (memory sort), 2 (memory region) or 3 (comparison)");
        DevBuf d_v(std::max<size_t>(nrows, 1)), d_k(std::max<size_t>(nrows, 1)), d_out(12 * n);
        if (nrows) to_device(ctx, d_v.p, vals, nrows);
        if (nrows) to_device(ctx, d_k.p, kinds, nrows);
        ola::lookup::rangecheck_trace(ctx, d_v.p, d_k.p, nrows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 12 * n);
    });
}
int ola_generate_bitwise_trace(ola_ctx* ctx, const uint64_t* tags, const uint64_t* op0, const uint64_t* op1, const uint64_t* res, size_t nrows,
                               uint32_t log_n, uint64_t* out, uint64_t* beta_out, int on_device) {
    if (!ctx || ((!tags || !op0 || !op1 || !res) && nrows) || !out || !beta_out || log_n < 18 || log_n > 24 || nrows > ((size_t)1 << log_n))
        return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            *beta_out = ola::lookup::bitwise_trace(ctx, tags, op0, op1, res, nrows, log_n, out);
            return;
        }
        const size_t m = std::max<size_t>(nrows, 1);
        DevBuf d_t(m), d_a(m), d_b(m), d_r(m), d_out(59 * n);
        if (nrows) {
            to_device(ctx, d_t.p, tags, nrows);
            to_device(ctx, d_a.p, op0, nrows);
            to_device(ctx, d_b.p, op1, nrows);
            to_device(ctx, d_r.p, res, nrows);
        }
        *beta_out = ola::lookup::bitwise_trace(ctx, d_t.p, d_a.p, d_b.p, d_r.p, nrows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 59 * n);
    });
}
int ola_generate_cmp_trace(ola_ctx* ctx, const uint64_t* cells, size_t nrows, uint32_t log_n, uint64_t* out, int on_device) {
    if (!ctx || (!cells && nrows) || !out || log_n < 1 || log_n > 28 || nrows > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::lookup::cmp_trace(ctx, cells, nrows, log_n, out);
            return;
        }
        DevBuf d_c(std::max<size_t>(nrows * 6, 1)), d_out(6 * n);
        if (nrows) to_device(ctx, d_c.p, cells, nrows * 6);
        ola::lookup::cmp_trace(ctx, d_c.p, nrows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 6 * n);
    });
}
int ola_generate_cpu_trace(ola_ctx* ctx, const uint64_t* steps, size_t nrows, uint32_t log_n, uint64_t* out, int on_device) {
    if (!ctx || (!steps && nrows) || !out || log_n > 26 || nrows > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::lookup::cpu_trace(ctx, steps, nrows, log_n, out);
            return;
        }
        DevBuf d_s(std::max<size_t>(nrows * 66, 1)), d_out(94 * n);
        if (nrows) to_device(ctx, d_s.p, steps, nrows * 66);
        ola::lookup::cpu_trace(ctx, d_s.p, nrows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 94 * n);
    });
}
int ola_generate_memory_trace(ola_ctx* ctx, const uint64_t* cells, size_t ncells, uint32_t log_n, uint64_t* out, int on_device) {
    if (!ctx || (!cells && ncells) || !out || log_n < 1 || log_n > 27 || ncells > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::lookup::memory_trace(ctx, cells, ncells, log_n, out);
            return;
        }
        DevBuf d_c(std::max<size_t>(ncells * 15, 1)), d_out(29 * n);
        if (ncells) to_device(ctx, d_c.p, cells, ncells * 15);
        ola::lookup::memory_trace(ctx, d_c.p, ncells, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 29 * n);
    });
}
int ola_generate_program_trace(ola_ctx* ctx, const uint64_t* steps, size_t nsteps, const uint64_t* prog_rows, size_t nprog_rows, const uint64_t* roots,
                               uint32_t log_n, uint64_t* out, uint64_t* beta_out, int on_device) {
    if (!ctx || (!steps && nsteps) || (!prog_rows && nprog_rows) || !roots || !out || !beta_out || log_n < 1 || log_n > 24) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            *beta_out = ola::lookup::program_trace(ctx, steps, nsteps, prog_rows, nprog_rows, roots, log_n, out);
            return;
        }
        DevBuf d_s(std::max<size_t>(nsteps * 66, 1)), d_p(std::max<size_t>(nprog_rows * 6, 1)), d_out(18 * n);
        if (nsteps) to_device(ctx, d_s.p, steps, nsteps * 66);
        if (nprog_rows) to_device(ctx, d_p.p, prog_rows, nprog_rows * 6);
        *beta_out = ola::lookup::program_trace(ctx, d_s.p, nsteps, d_p.p, nprog_rows, roots, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 18 * n);
    });
}
extern "C++" {
// the five small tables: record rows in, one column-major table out (gen_tables.cu)
template <class F>
static int generate_small(ola_ctx* ctx, const uint64_t* rows, size_t nrows, size_t rec, uint32_t log_n, int ncols, uint64_t* out, int on_device, F&& run) {
    if (!ctx || (!rows && nrows) || !out || log_n < 1 || log_n > 28 || nrows > ((size_t)1 << log_n)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            run(rows, out);
            return;
        }
        DevBuf d_r(std::max<size_t>(nrows * rec, 1)), d_out((size_t)ncols * n);
        if (nrows) to_device(ctx, d_r.p, rows, nrows * rec);
        run(d_r.p, d_out.p);
        to_host(ctx, out, d_out.p, (size_t)ncols * n);
    });
}
}
int ola_generate_poseidon_chunk_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device) {
    return generate_small(ctx, rows, nrows, 32, log_n, 53, out, on_device,
                          [&](const uint64_t* r, uint64_t* o) { ola::lookup::poseidon_chunk_trace(ctx, r, nrows, log_n, o); });
}
int ola_generate_storage_access_trace(ola_ctx* ctx, const uint64_t* rows, size_t n_access, size_t n_prog_reads, uint32_t log_n, uint64_t* out,
                                      int on_device) {
    if (n_access + n_prog_reads < n_access) return OLA_ERR_INVALID_ARG;
    return generate_small(ctx, rows, n_access + n_prog_reads, 38, log_n, 48, out, on_device,
                          [&](const uint64_t* r, uint64_t* o) { ola::lookup::storage_access_trace(ctx, r, n_access, n_prog_reads, log_n, o); });
}
int ola_generate_tape_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device) {
    return generate_small(ctx, rows, nrows, 5, log_n, 6, out, on_device,
                          [&](const uint64_t* r, uint64_t* o) { ola::lookup::tape_trace(ctx, r, nrows, log_n, o); });
}
int ola_generate_sccall_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device) {
    return generate_small(ctx, rows, nrows, 24, log_n, 26, out, on_device,
                          [&](const uint64_t* r, uint64_t* o) { ola::lookup::sccall_trace(ctx, r, nrows, log_n, o); });
}
int ola_generate_prog_chunk_trace(ola_ctx* ctx, const uint64_t* prog_rows, size_t nprog_rows, uint32_t log_n, uint64_t* out, int on_device) {
    if (!ctx || (!prog_rows && nprog_rows) || !out || log_n < 1 || log_n > 28 || nprog_rows > ((size_t)1 << 31)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        const size_t n = (size_t)1 << log_n;
        if (on_device) {
            ola::lookup::prog_chunk_trace(ctx, prog_rows, nprog_rows, log_n, out);
            return;
        }
        DevBuf d_r(std::max<size_t>(nprog_rows * 6, 1)), d_out(40 * n);
        if (nprog_rows) to_device(ctx, d_r.p, prog_rows, nprog_rows * 6);
        ola::lookup::prog_chunk_trace(ctx, d_r.p, nprog_rows, log_n, d_out.p);
        to_host(ctx, out, d_out.p, 40 * n);
    });
}
int ola_compress_challenge(const uint64_t* const* cols, uint32_t ncols, size_t n, uint64_t* beta_out) {
    if ((!cols && ncols) || !beta_out) return OLA_ERR_INVALID_ARG;
    try {
        ola::stark::Challenger ch(OLA_HASH_POSEIDON);
        for (uint32_t c = 0; c < ncols; ++c) {
            if (!cols[c] && n) return OLA_ERR_INVALID_ARG;
            for (size_t i = 0; i < n; ++i) ch.observe(cols[c][i]);
        }
        *beta_out = ch.get_challenge();
        return OLA_OK;
    } catch (...) {
        return OLA_ERR_INTERNAL;
    }
}

int ola_air_constraints(int table_id, const uint64_t* lv, const uint64_t* nv, uint64_t compress_challenge, uint64_t* vals_out, int* kinds_out, int cap) {
    if (!lv || !nv || (cap > 0 && (!vals_out || !kinds_out))) return OLA_ERR_INVALID_ARG;
    try {
        if (!ola::stark::table_available(table_id)) return OLA_ERR_INVALID_ARG;
        ola::stark::TableInfo t = ola::stark::table_info(table_id);
        t.compress_challenge = gl::canon(compress_challenge);
        using ola::stark::verify::XE;
        std::vector<XE> l((size_t)t.columns), n((size_t)t.columns);
        for (int c = 0; c < t.columns; ++c) {
            l[c] = XE((ola::stark::F)lv[c]);
            n[c] = XE((ola::stark::F)nv[c]);
        }
        ola::stark::verify::XRow lrow{l.data()}, nrow{n.data()};
        ola::stark::verify::RecordingConsumer rc;
        ola::stark::verify::eval_table_t(t, lrow, nrow, rc);
        const int k = (int)rc.vals.size();
        for (int i = 0; i < k && i < cap; ++i) {
            vals_out[i] = gl::canon(rc.vals[i].v.c0);  // base-field rows: the extension component stays 0
            kinds_out[i] = rc.kinds[i];
        }
        return k;
    } catch (...) {
        return OLA_ERR_INTERNAL;
    }
}

int ola_table_columns(int table_id) {
    try {
        return ola::stark::table_available(table_id) ? ola::stark::table_info(table_id).columns : -1;
    } catch (...) {
        return -1;
    }
}

int ola_batch_free(ola_ctx* ctx, ola_batch* b) {
    if (!ctx) return OLA_ERR_INVALID_ARG;
    if (!b) return OLA_OK;
    return guarded(ctx, [&] {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        ola::batch_release(b);
        delete b;
    });
}
size_t ola_batch_ncols(const ola_batch* b) { return b ? b->ncols : 0; }
uint32_t ola_batch_degree_log(const ola_batch* b) { return b ? b->log_n : 0; }
uint32_t ola_batch_rate_bits(const ola_batch* b) { return b ? b->rate_bits : 0; }
const uint64_t* ola_batch_coeffs_dev(const ola_batch* b) { return b ? b->d_coeffs : nullptr; }
const uint64_t* ola_batch_lde_dev(const ola_batch* b) { return b ? b->d_lde : nullptr; }
const uint64_t* ola_batch_nodes_dev(const ola_batch* b) { return b ? b->d_nodes : nullptr; }

int ola_batch_get_coeffs(ola_ctx* ctx, const ola_batch* b, uint64_t* out_host) {
    if (!ctx || !b || !out_host) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { to_host(ctx, out_host, b->d_coeffs, b->ncols << b->log_n); });
}
int ola_batch_get_cap(ola_ctx* ctx, const ola_batch* b, uint64_t* cap_out_host) {
    if (!ctx || !b || !cap_out_host) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { ola::batch_get_cap(ctx, b, cap_out_host); });
}
int ola_batch_get_leaves(ola_ctx* ctx, const ola_batch* b, size_t leaf_index, size_t count, uint64_t* out_host) {
    if (!ctx || !b || (!out_host && count)) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { ola::batch_get_leaves(ctx, b, leaf_index, count, out_host); });
}
int ola_batch_prove_leaf(ola_ctx* ctx, const ola_batch* b, size_t leaf_index, uint64_t* siblings_out_host) {
    if (!ctx || !b || !siblings_out_host) return OLA_ERR_INVALID_ARG;
    int n = 0;
    int rc = guarded(ctx, [&] { n = ola::batch_prove_leaf(ctx, b, leaf_index, siblings_out_host); });
    return rc == OLA_OK ? n : rc;
}

// ---- Trace JSON ingest and the `ola prove` flow ----
}  // extern "C"
#include "trace_json.h"
struct ola_trace {
    ola::tracejson::Records rec;
};
namespace {
uint32_t log_rows(size_t filled, size_t at_least) {  // the generators' "next power of two, at least `at_least`"
    uint32_t lg = 0;
    while (((size_t)1 << lg) < std::max(filled, at_least)) ++lg;
    return lg;
}
// the row count generate_traces gives each table (circuits/src/generation/*.rs, the first lines of every generator)
uint32_t trace_table_log_rows(const ola::tracejson::Records& r, int table) {
    using namespace ola::tracejson;
    switch (table) {
        case 0: return log_rows(r.n(REC_STEP), 1);                        // cpu.rs:13-18
        case 1: return log_rows(r.n(REC_MEMORY), 2);                      // memory.rs:11-20
        case 2: return log_rows(r.n(REC_BW_TAG), (size_t)3 << 16);        // builtin.rs:39-52 (BITWISE_U8_SIZE = 3 * 2^16)
        case 3: return log_rows(r.n(REC_CMP), 2);                         // builtin.rs:209-218
        case 4: return log_rows(r.n(REC_RC_VAL), (size_t)1 << 16);        // builtin.rs:252-262
        case 5: return log_rows(r.n(REC_PSDN_INPUT), 2);                  // poseidon.rs:6-15
        case 6: return log_rows(r.n(REC_PCHUNK), 2);
        case 7: return log_rows(r.n(REC_STORAGE), 2);
        case 8: return log_rows(r.n(REC_TAPE), 2);
        case 9: return log_rows(r.n(REC_SCCALL), 2);
        case 10: {  // prog.rs:30-55: max(words fetched by the executed lines, words of all programs)
            size_t exec_len = 0;
            const uint64_t* steps = r.rows(REC_STEP);
            for (size_t i = 0; i < r.n(REC_STEP); ++i) {
                const uint64_t* s = steps + i * 66;
                if (s[13] != 0) continue;
                exec_len += (s[26] == 1 || s[27] == (1ull << 22) || s[27] == (1ull << 21)) ? 2 : 1;
            }
            return log_rows(std::max(exec_len, r.n(REC_PROG_ROW)), 2);
        }
        case 11: {
            size_t lines = 0;
            const uint64_t* pr = r.rows(REC_PROG_ROW);
            for (size_t i = 0; i < r.n(REC_PROG_ROW); ++i) lines += pr[i * 6 + 4] % 8 == 0;
            return log_rows(lines, 2);
        }
    }
    return 0;
}
struct Upload {  // a record array in device memory for the duration of one generator call
    DevBuf d;
    Upload(ola_ctx* ctx, const ola::tracejson::Records& r, int kind) : d(std::max<size_t>(r.u64s(kind), 1)) {
        if (r.u64s(kind)) to_device(ctx, d.p, r.rows(kind), r.u64s(kind));
    }
};
// The Bitwise table's compress challenge computed on a second host thread (host-only work: lookup.cu bitwise_beta)
struct BitwiseLate {
    std::vector<uint64_t> limbs;
    std::thread worker;
    uint64_t beta = 0;
    bool failed = false;
    ~BitwiseLate() {
        if (worker.joinable()) worker.join();
    }
};
// generate_traces (circuits/src/generation/mod.rs:79-213): twelve column-major tables in device memory, tables[t] of
// ola_table_columns(t) << log_ns[t] u64 (owned by the caller: ola_dev_free), and the two compress challenges.  With `defer` the
// Bitwise table is left at bitwise_trace_begin (its challenge is being computed by defer->worker; cc[2] stays 0) and the caller
// completes it with bitwise_trace_finish after joining the worker.
void generate_all(ola_ctx* ctx, const ola::tracejson::Records& r, uint64_t** tables, uint32_t* log_ns, uint64_t* cc, BitwiseLate* defer = nullptr) {
    namespace L = ola::lookup;
    using namespace ola::tracejson;
    OLA_CHECK(r.n(REC_RC_KIND) == r.n(REC_RC_VAL) && r.n(REC_BW_OP0) == r.n(REC_BW_TAG) && r.n(REC_BW_OP1) == r.n(REC_BW_TAG) &&
                  r.n(REC_BW_RES) == r.n(REC_BW_TAG) && r.n(REC_PSDN_FILTER) == r.n(REC_PSDN_INPUT) && r.n_storage_access <= r.n(REC_STORAGE),
              OLA_ERR_INVALID_ARG, "trace records: the parallel arrays of one table differ in length");
    for (int t = 0; t < 12; ++t) {
        tables[t] = nullptr;
        cc[t] = 0;
        log_ns[t] = trace_table_log_rows(r, t);
    }
    try {
        for (int t = 0; t < 12; ++t) ola::dev_alloc(&tables[t], (size_t)ola::stark::table_info(t).columns << log_ns[t]);
        {  // first: with `defer` its sequential host transcript then runs beside everything below (and beside the commitments)
            Upload a(ctx, r, REC_BW_TAG), b(ctx, r, REC_BW_OP0), c(ctx, r, REC_BW_OP1), d(ctx, r, REC_BW_RES);
            if (defer) {
                defer->limbs = L::bitwise_trace_begin(ctx, a.d.p, b.d.p, c.d.p, d.d.p, r.n(REC_BW_TAG), log_ns[2], tables[2]);
                defer->worker = std::thread([defer] {
                    try {
                        defer->beta = L::bitwise_beta(defer->limbs);
                    } catch (...) {
                        defer->failed = true;
                    }
                });
            } else {
                cc[2] = L::bitwise_trace(ctx, a.d.p, b.d.p, c.d.p, d.d.p, r.n(REC_BW_TAG), log_ns[2], tables[2]);
            }
        }
        {
            Upload s(ctx, r, REC_STEP), p(ctx, r, REC_PROG_ROW);  // the Step records serve the CPU and the Program table
            L::cpu_trace(ctx, s.d.p, r.n(REC_STEP), log_ns[0], tables[0]);
            cc[10] = L::program_trace(ctx, s.d.p, r.n(REC_STEP), p.d.p, r.n(REC_PROG_ROW), r.roots, log_ns[10], tables[10]);
            L::prog_chunk_trace(ctx, p.d.p, r.n(REC_PROG_ROW), log_ns[11], tables[11]);
            OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        { Upload u(ctx, r, REC_MEMORY); L::memory_trace(ctx, u.d.p, r.n(REC_MEMORY), log_ns[1], tables[1]); }
        { Upload u(ctx, r, REC_CMP); L::cmp_trace(ctx, u.d.p, r.n(REC_CMP), log_ns[3], tables[3]); }
        { Upload v(ctx, r, REC_RC_VAL), k(ctx, r, REC_RC_KIND); L::rangecheck_trace(ctx, v.d.p, k.d.p, r.n(REC_RC_VAL), log_ns[4], tables[4]); }
        {
            Upload i(ctx, r, REC_PSDN_INPUT), f(ctx, r, REC_PSDN_FILTER);
            ola::generation::poseidon_trace(ctx, i.d.p, f.d.p, r.n(REC_PSDN_INPUT), log_ns[5], tables[5]);
        }
        { Upload u(ctx, r, REC_PCHUNK); L::poseidon_chunk_trace(ctx, u.d.p, r.n(REC_PCHUNK), log_ns[6], tables[6]); }
        {
            Upload u(ctx, r, REC_STORAGE);
            L::storage_access_trace(ctx, u.d.p, r.n_storage_access, r.n(REC_STORAGE) - r.n_storage_access, log_ns[7], tables[7]);
        }
        { Upload u(ctx, r, REC_TAPE); L::tape_trace(ctx, u.d.p, r.n(REC_TAPE), log_ns[8], tables[8]); }
        { Upload u(ctx, r, REC_SCCALL); L::sccall_trace(ctx, u.d.p, r.n(REC_SCCALL), log_ns[9], tables[9]); }
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        cudaStreamSynchronize(ctx->stream);
        for (int t = 0; t < 12; ++t)
            if (tables[t]) ola::dev_free(tables[t]), tables[t] = nullptr;
        throw;
    }
}
}  // namespace
extern "C" {
int ola_trace_from_json(const char* json, size_t len, ola_trace** out, char* err, size_t errcap) {
    if (err && errcap) err[0] = 0;
    if (!json || !out) return OLA_ERR_INVALID_ARG;
    *out = nullptr;
    try {
        std::unique_ptr<ola_trace> t(new ola_trace());
        ola::tracejson::parse(json, len, t->rec);
        *out = t.release();
        return OLA_OK;
    } catch (const std::exception& e) {
        if (err && errcap) {
            strncpy(err, e.what(), errcap - 1);
            err[errcap - 1] = 0;
        }
        return OLA_ERR_INVALID_ARG;
    } catch (...) {
        return OLA_ERR_INTERNAL;
    }
}
void ola_trace_free(ola_trace* t) { delete t; }
int ola_trace_new(ola_trace** out) {
    if (!out) return OLA_ERR_INVALID_ARG;
    try {
        *out = new ola_trace();
        return OLA_OK;
    } catch (...) {
        return OLA_ERR_OOM;
    }
}
int ola_trace_set_records(ola_trace* t, int kind, const uint64_t* rows, size_t nrows) {
    if (!t || (!rows && nrows && kind != OLA_REC_STORAGE_ACCESS_COUNT)) return OLA_ERR_INVALID_ARG;
    ola::tracejson::Records& r = t->rec;
    if (kind == OLA_REC_ROOTS) {
        if (nrows != 1) return OLA_ERR_INVALID_ARG;
        memcpy(r.roots, rows, sizeof r.roots);
        return OLA_OK;
    }
    if (kind == OLA_REC_STORAGE_ACCESS_COUNT) {
        r.n_storage_access = nrows;
        return OLA_OK;
    }
    if (kind < 0 || kind >= ola::tracejson::REC_KINDS) return OLA_ERR_INVALID_ARG;
    r.own[kind].clear();
    r.ptr[kind] = nrows ? rows : nullptr;  // borrowed: the caller keeps the array alive while the trace is in use
    r.count[kind] = nrows;
    return OLA_OK;
}
int ola_trace_records(const ola_trace* t, int kind, const uint64_t** rows, size_t* nrows, uint32_t* rec_u64) {
    if (!t || !rows || !nrows || !rec_u64) return OLA_ERR_INVALID_ARG;
    const ola::tracejson::Records& r = t->rec;
    if (kind == OLA_REC_ROOTS) {
        *rows = r.roots, *nrows = 1, *rec_u64 = 8;
        return OLA_OK;
    }
    if (kind == OLA_REC_STORAGE_ACCESS_COUNT) {
        *rows = nullptr, *nrows = r.n_storage_access, *rec_u64 = 0;
        return OLA_OK;
    }
    if (kind < 0 || kind >= ola::tracejson::REC_KINDS) return OLA_ERR_INVALID_ARG;
    *rows = r.rows(kind), *nrows = r.n(kind), *rec_u64 = ola::tracejson::kRecWidth[kind];
    return OLA_OK;
}
int ola_trace_table_log_rows(const ola_trace* t, int table_id) {
    if (!t || table_id < 0 || table_id > 11) return OLA_ERR_INVALID_ARG;
    return (int)trace_table_log_rows(t->rec, table_id);
}
int ola_generate_traces(ola_ctx* ctx, const ola_trace* t, uint64_t** tables_dev, uint32_t* log_ns, uint64_t* compress_challenges) {
    if (!ctx || !t || !tables_dev || !log_ns || !compress_challenges) return OLA_ERR_INVALID_ARG;
    return guarded(ctx, [&] { generate_all(ctx, t->rec, tables_dev, log_ns, compress_challenges); });
}
int ola_prove_trace(ola_ctx* ctx, const ola_trace* t, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!ctx || !t || !proof_len || (!proof_out && proof_cap)) return OLA_ERR_INVALID_ARG;
    *proof_len = 0;
    return guarded(ctx, [&] {
        uint64_t* tables[12];
        uint32_t log_ns[12];
        uint64_t cc[12];
        BitwiseLate bw;  // joined by its destructor on every path
        generate_all(ctx, t->rec, tables, log_ns, cc, &bw);
        const ola::stark::LateTable late{2, [&]() -> uint64_t {
                                             bw.worker.join();
                                             OLA_CHECK(!bw.failed, OLA_ERR_INTERNAL, "the Bitwise compress challenge could not be computed");
                                             ola::lookup::bitwise_trace_finish(ctx, log_ns[2], bw.beta, tables[2]);
                                             return bw.beta;
                                         }};
        std::vector<uint8_t> bytes;
        try {
            ola::stark::Config cfg;  // StarkConfig::standard_fast_config(), the degree check on (client/src/main.rs:192)
            std::vector<int> ids = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
            std::vector<const uint64_t*> tr(tables, tables + 12);
            std::vector<uint32_t> lg(log_ns, log_ns + 12);
            std::vector<uint64_t> c(cc, cc + 12);
            bytes = ola::stark::prove_all(ctx, ids, tr, true, lg, c, cfg, nullptr, &late);
        } catch (...) {
            cudaStreamSynchronize(ctx->stream);
            for (int i = 0; i < 12; ++i) ola::dev_free(tables[i]);
            throw;
        }
        for (int i = 0; i < 12; ++i) ola::dev_free(tables[i]);
        *proof_len = bytes.size();
        OLA_CHECK(bytes.size() <= proof_cap, OLA_ERR_INVALID_ARG, "proof buffer too small (needed size returned in proof_len)");
        memcpy(proof_out, bytes.data(), bytes.size());
    });
}

}  // extern "C"
