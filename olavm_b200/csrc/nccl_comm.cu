// Native NCCL communicator for the coset-sharded prover: the library binds libnccl itself (dlopen, no
// link-time dependency) and issues ncclAllGather / ncclAllReduce on the context's stream -- no host-language callback on
// the proving path.  The host only moves the 128-byte ncclUniqueId from rank 0 to the other ranks (any side channel:
// torch.distributed, MPI, a file, the Rust host's own RPC).
#include <dlfcn.h>

#include <cstring>

#include "common.h"

namespace ola {
namespace nccl {

struct UniqueId {
    char internal[128];  // NCCL_UNIQUE_ID_BYTES
};
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(void**, int, UniqueId, int);
typedef int (*CommDestroy_t)(void*);
typedef const char* (*GetErrorString_t)(int);
typedef int (*AllGather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommSplit_t)(void*, int, int, void**, void*);
enum { kUint8 = 1, kUint64 = 5, kSum = 0 };  // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct Api {
    void* handle = nullptr;
    GetUniqueId_t get_unique_id = nullptr;
    CommInitRank_t comm_init_rank = nullptr;
    CommDestroy_t comm_destroy = nullptr;
    GetErrorString_t error_string = nullptr;
    AllGather_t all_gather = nullptr;
    AllReduce_t all_reduce = nullptr;
    CommSplit_t comm_split = nullptr;  // optional (NCCL >= 2.18)
};

static bool load_api(const char* path, Api& api, std::string& err) {
    void* h = nullptr;
    if (path && *path) {
        h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    } else {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already loaded (torch's), if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) {
        err = std::string("cannot load libnccl: ") + (dlerror() ? dlerror() : "not found");
        return false;
    }
    api.handle = h;
    api.get_unique_id = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
    api.error_string = (GetErrorString_t)dlsym(h, "ncclGetErrorString");
    api.all_gather = (AllGather_t)dlsym(h, "ncclAllGather");
    api.all_reduce = (AllReduce_t)dlsym(h, "ncclAllReduce");
    api.comm_split = (CommSplit_t)dlsym(h, "ncclCommSplit");
    if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather || !api.all_reduce) {
        err = "libnccl lacks a required symbol";
        return false;
    }
    return true;
}

struct Comm {
    Api api;
    void* comm = nullptr;
    void* comm2 = nullptr;  // a second communicator over the same ranks, for collectives on the copy stream that overlap
                            // the proving stream's (collectives of ONE communicator are serialised by NCCL)
};

static int cb_allgather(void* user, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
    Comm* c = (Comm*)user;
    return c->api.all_gather(send, recv, bytes_per_rank, kUint8, c->comm, (cudaStream_t)stream);
}
static int cb_allreduce(void* user, void* buf, size_t count, void* stream) {
    Comm* c = (Comm*)user;
    return c->api.all_reduce(buf, buf, count, kUint64, kSum, c->comm, (cudaStream_t)stream);
}

// all-gather on the second communicator, on `stream`; false when there is none (the caller falls back to the first)
bool allgather_side(ola_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
    Comm* c = (Comm*)ctx->nccl_state;
    if (!c || !c->comm2) return false;
    ctx->comm_bytes += bytes_per_rank * (size_t)ctx->world;
    if (c->api.all_gather(send, recv, bytes_per_rank, kUint8, c->comm2, stream) != 0) throw ola::Error(OLA_ERR_INTERNAL, "ncclAllGather (side communicator) failed");
    return true;
}
bool has_side_comm(const ola_ctx* ctx) {
    const Comm* c = (const Comm*)ctx->nccl_state;
    return c && c->comm2;
}

void release(ola_ctx* ctx) {
    Comm* c = (Comm*)ctx->nccl_state;
    if (!c) return;
    if (c->comm2) c->api.comm_destroy(c->comm2);
    if (c->comm) c->api.comm_destroy(c->comm);
    delete c;
    ctx->nccl_state = nullptr;
}

}  // namespace nccl
}  // namespace ola

extern "C" {

int ola_nccl_unique_id(const char* libnccl_path, uint8_t id_out[128]) {
    if (!id_out) return OLA_ERR_INVALID_ARG;
    ola::nccl::Api api;
    std::string err;
    if (!ola::nccl::load_api(libnccl_path, api, err)) return OLA_ERR_INTERNAL;
    ola::nccl::UniqueId id;
    if (api.get_unique_id(&id) != 0) return OLA_ERR_INTERNAL;
    memcpy(id_out, id.internal, 128);
    return OLA_OK;
}

int ola_set_comm_nccl(ola_ctx* ctx, const char* libnccl_path, int rank, int world, const uint8_t id[128]) {
    if (!ctx || !id) return OLA_ERR_INVALID_ARG;
    try {
        if (!(world >= 1 && world <= 8 && (8 % world) == 0 && rank >= 0 && rank < world))
            throw ola::Error(OLA_ERR_INVALID_ARG, "set_comm_nccl: world must divide the blowup (8) and 0 <= rank < world");
        OLA_CUDA(cudaSetDevice(ctx->device));
        ola::nccl::release(ctx);
        auto* c = new ola::nccl::Comm();
        std::string err;
        if (!ola::nccl::load_api(libnccl_path, c->api, err)) {
            delete c;
            throw ola::Error(OLA_ERR_INTERNAL, err);
        }
        ola::nccl::UniqueId uid;
        memcpy(uid.internal, id, 128);
        const int rc = c->api.comm_init_rank(&c->comm, world, uid, rank);
        if (rc != 0) {
            std::string m = std::string("ncclCommInitRank: ") + (c->api.error_string ? c->api.error_string(rc) : "error");
            delete c;
            throw ola::Error(OLA_ERR_INTERNAL, m);
        }
        // OLA_NCCL_SIDE_COMM=0 disables the second communicator (no overlap of the trace all-gathers with the commits)
        const char* side = getenv("OLA_NCCL_SIDE_COMM");
        if (world > 1 && c->api.comm_split && !(side && side[0] == '0')) {
            if (c->api.comm_split(c->comm, 0, rank, &c->comm2, nullptr) != 0) c->comm2 = nullptr;
        }
        ctx->nccl_state = c;
        ctx->rank = rank;
        ctx->world = world;
        ctx->comm_allgather = ola::nccl::cb_allgather;
        ctx->comm_allreduce = ola::nccl::cb_allreduce;
        ctx->comm_user = c;
        return OLA_OK;
    } catch (const ola::Error& e) {
        ctx->last_error = e.what();
        return e.code;
    } catch (...) {
        ctx->last_error = "set_comm_nccl failed";
        return OLA_ERR_INTERNAL;
    }
}

uint64_t ola_comm_bytes(const ola_ctx* ctx) { return ctx ? ctx->comm_bytes : 0; }

}  // extern "C"
