// K9 -- per-ICP-iteration all-reduce of the 30 normal-equation accumulators across the ranks of
// one node (the reference has no multi-GPU code).  One process per GPU; the
// host program (torch.distributed) broadcasts the ncclUniqueId; libnccl.so.2 (the one torch bundles)
// is dlopen'ed so the library has no link-time NCCL dependency.  The all-reduce is enqueued on the
// context's stream between the correspondence+reduction kernel and icp_step_kernel: no host sync.
#include <dlfcn.h>

#include "internal.cuh"

namespace pls {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

bool load_nccl(const char* path, NcclApi& api, std::string& err) {
    api.handle = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) {
        err = std::string("dlopen(libnccl) failed: ") + dlerror();
        return false;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
        err = "libnccl is missing a required symbol";
        return false;
    }
    return true;
}
}  // namespace

struct Comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
};

int comm_rank(pls_context* ctx) { return ctx->comm ? ctx->comm->rank : 0; }
int comm_size(pls_context* ctx) { return ctx->comm ? ctx->comm->size : 1; }

void comm_allreduce_sums(pls_context* ctx, double* sums_dev) {
    Comm* c = ctx->comm;
    if (!c || c->size <= 1) return;
    ncclResult_t r = c->api.AllReduce(sums_dev, sums_dev, NACC, kNcclDouble, kNcclSum, c->comm, ctx->stream);
    if (r != 0)
        throw Error{PLS_E_COMM, std::string("ncclAllReduce: ") + (c->api.GetErrorString ? c->api.GetErrorString(r) : "error")};
}

void comm_free(pls_context* ctx) {
    if (!ctx->comm) return;
    if (ctx->comm->comm) ctx->comm->api.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

}  // namespace pls

using namespace pls;

extern "C" {

int pls_comm_unique_id(const char* nccl_library, void* out_id_128_bytes) {
    if (!out_id_128_bytes) return PLS_E_INVALID;
    NcclApi api;
    std::string err;
    if (!load_nccl(nccl_library, api, err)) {
        fprintf(stderr, "pls_comm_unique_id: %s\n", err.c_str());
        return PLS_E_COMM;
    }
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != 0) return PLS_E_COMM;
    memcpy(out_id_128_bytes, id.internal, 128);
    return PLS_OK;
}

int pls_comm_init(pls_context* ctx, int num_ranks, int rank, const void* nccl_unique_id, const char* nccl_library) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(num_ranks >= 1 && rank >= 0 && rank < num_ranks && nccl_unique_id, "pls_comm_init: bad arguments");
    comm_free(ctx);
    if (num_ranks == 1) return PLS_OK;
    Comm* c = new Comm();
    std::string err;
    if (!load_nccl(nccl_library, c->api, err)) {
        delete c;
        throw pls::Error{PLS_E_COMM, err};
    }
    ncclUniqueId id;
    memcpy(id.internal, nccl_unique_id, 128);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    ncclResult_t r = c->api.CommInitRank(&c->comm, num_ranks, id, rank);
    if (r != 0) {
        std::string msg = std::string("ncclCommInitRank: ") + (c->api.GetErrorString ? c->api.GetErrorString(r) : "error");
        delete c;
        throw pls::Error{PLS_E_COMM, msg};
    }
    c->rank = rank;
    c->size = num_ranks;
    ctx->comm = c;
    PLS_API_END(ctx)
}

int pls_comm_destroy(pls_context* ctx) {
    PLS_API_BEGIN(ctx)
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    comm_free(ctx);
    PLS_API_END(ctx)
}

}  // extern "C"
