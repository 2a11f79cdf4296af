// Internal shared declarations of libola_gpu (context, error handling, device buffers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ola_gpu.h"

namespace ola {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define OLA_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ola::Error(OLA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                               __FILE__ + ":" + std::to_string(__LINE__) + ")");    \
    } while (0)

#define OLA_CHECK(cond, code, msg)                         \
    do {                                                   \
        if (!(cond)) throw ola::Error((code), (msg));      \
    } while (0)

// Device twiddle material, built once per context (the analogue of the reference's
// `twiddle_map: BTreeMap<usize, Vec<F>>`, fri/oracle.rs:51, prover.rs:106 -- but size-independent).
struct Twiddles {
    // Three-level power tables of Omega = omega_{2^32} (TA[x] = Omega^x, x < 2^11; TB[x] = Omega^(x<<11),
    // x < 2^11; TC[x] = Omega^(x<<22), x < 2^10), forward ([0]) and inverse ([1]).
    uint64_t* pw[2] = {nullptr, nullptr};  // 5120 entries each: TA | TB | TC
    // Bit-reversed small root table BRS[i] = omega_{2^(k+1)}^{bitrev_k(i)} for i < 2^k <= 2^11
    // (prefix property: one table serves every k), forward and inverse.
    uint64_t* brs[2] = {nullptr, nullptr};  // 2048 entries each
};

}  // namespace ola

namespace ola {
// Per-launch CUDA-event tracing (the device-side analogue of the reference's TimingTree,
// plonky2/plonky2/src/util/timing.rs): off by default, enabled by ola_profile_begin().
struct ProfRecord {
    std::string name;
    cudaEvent_t start, stop;
};
}  // namespace ola

struct ola_ctx {
    bool profiling = false;
    std::vector<ola::ProfRecord> prof_pending;
    std::map<std::string, std::pair<double, uint64_t>> prof_totals;  // name -> (ms, launches)
    int device = 0;
    int hasher = 0;  // OLA_HASH_POSEIDON / OLA_HASH_BLAKE3: C::Hasher of the commitments and the transcript (ola_set_hasher)
    cudaStream_t stream = nullptr;
    // second stream + events for uploads that overlap the transforms of the previous chunk (batch.cu ingest_host_columns)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    int sm_count = 148;
    ola::Twiddles tw;
    std::string last_error;
    uint64_t kernel_launches = 0;  // counted by every launcher (bench.py "gpu_launches")
    uint64_t* scratch = nullptr;   // grow-only device workspace (multi-pass transform intermediates)
    size_t scratch_elems = 0;
    // coset-shard communicator (ola_set_comm); world == 1: single GPU
    int rank = 0, world = 1;
    ola_allgather_fn comm_allgather = nullptr;
    ola_allreduce_u64_fn comm_allreduce = nullptr;
    void* comm_user = nullptr;
    void* nccl_state = nullptr;   // ola_set_comm_nccl: the library's own NCCL communicator (nccl_comm.cu)
    uint64_t comm_bytes = 0;      // bytes this rank received through the communicator (ola_comm_bytes)
};

namespace ola {
// RAII scope around one kernel launch: counts it, and when profiling brackets it with events on ctx->stream.
struct Launch {
    ola_ctx* ctx;
    cudaEvent_t start = nullptr, stop = nullptr;
    const char* name;
    Launch(ola_ctx* c, const char* n) : ctx(c), name(n) {
        ctx->kernel_launches += 1;
        if (ctx->profiling) {
            cudaEventCreate(&start);
            cudaEventCreate(&stop);
            cudaEventRecord(start, ctx->stream);
        }
    }
    ~Launch() {
        if (start) {
            cudaEventRecord(stop, ctx->stream);
            ctx->prof_pending.push_back({name, start, stop});
        }
    }
};
inline void count_launch(ola_ctx* ctx, uint64_t n = 1) { ctx->kernel_launches += n; }
// the context's copy stream and its double-buffering events, created on first use
inline void ensure_copy_stream(ola_ctx* ctx) {
    if (ctx->copy_stream) return;
    cudaStream_t s = nullptr;
    OLA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        OLA_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
        OLA_CUDA(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
    }
    ctx->copy_stream = s;
}
inline void comm_allgather(ola_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    OLA_CHECK(ctx->comm_allgather != nullptr, OLA_ERR_INTERNAL, "no communicator");
    ctx->comm_bytes += bytes_per_rank * (size_t)ctx->world;
    Launch lz(ctx, "comm_allgather");  // event-timed like a kernel: transfer time plus the wait for the slowest rank
    OLA_CHECK(ctx->comm_allgather(ctx->comm_user, send, recv, bytes_per_rank, (void*)ctx->stream) == 0, OLA_ERR_INTERNAL, "all-gather callback failed");
}
inline void comm_allreduce(ola_ctx* ctx, void* buf, size_t count) {
    OLA_CHECK(ctx->comm_allreduce != nullptr, OLA_ERR_INTERNAL, "no communicator");
    ctx->comm_bytes += count * 8;
    Launch lz(ctx, "comm_allreduce");
    OLA_CHECK(ctx->comm_allreduce(ctx->comm_user, buf, count, (void*)ctx->stream) == 0, OLA_ERR_INTERNAL, "all-reduce callback failed");
}
inline void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(OLA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace ola
