// K9 -- per-ICP-iteration all-reduce of the 30 normal-equation accumulators across the ranks of
// one node (the reference has no multi-GPU code).  One process per GPU; the
// host program (torch.distributed) broadcasts the ncclUniqueId; libnccl.so.2 (the one torch bundles)
// is dlopen'ed so the library has no link-time NCCL dependency.  The all-reduce is enqueued on the
// context's stream between the correspondence+reduction kernel and icp_step_kernel: no host sync.
#include <dlfcn.h>

#include <vector>

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

constexpr size_t kP2PSlotBytes = 256;  // sizeof(P2PSlot): 30 doubles + seq + pad

struct Comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    // one-shot peer-to-peer mode
    bool p2p = false;
    void* xchg = nullptr;            // this rank's exchange buffer [2][size] slots (written by the peers)
    std::vector<void*> peer_ptrs;    // every rank's exchange buffer mapped into this process (own = xchg)
    void* peer_table_dev = nullptr;  // device copy of peer_ptrs
    unsigned long long* seq_dev = nullptr;  // number of exchanges this rank has EXECUTED (device side: a launch that finds
                                            // the frame converged does not count) -- the round number of the next one
    double* allreduce_buf = nullptr;        // NCCL mode: the sums in flight (kept apart from the FrameResult)
};

bool comm_is_p2p(pls_context* ctx) { return ctx->comm && ctx->comm->p2p; }
void* comm_p2p_peers(pls_context* ctx) { return ctx->comm->peer_table_dev; }
unsigned long long* comm_p2p_seq(pls_context* ctx) { return ctx->comm->seq_dev; }
double* comm_allreduce_buffer(pls_context* ctx) {
    Comm* c = ctx->comm;
    if (!c->allreduce_buf) {
        PLS_CUDA(cudaMalloc(&c->allreduce_buf, NACC * sizeof(double)));
        PLS_CUDA(cudaMemset(c->allreduce_buf, 0, NACC * sizeof(double)));
    }
    return c->allreduce_buf;
}

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
    if (ctx->comm->p2p) {
        for (int r = 0; r < (int)ctx->comm->peer_ptrs.size(); ++r)
            if (r != ctx->comm->rank && ctx->comm->peer_ptrs[r]) cudaIpcCloseMemHandle(ctx->comm->peer_ptrs[r]);
        if (ctx->comm->peer_table_dev) cudaFree(ctx->comm->peer_table_dev);
        if (ctx->comm->seq_dev) cudaFree(ctx->comm->seq_dev);
        if (ctx->comm->xchg) cudaFree(ctx->comm->xchg);
    }
    if (ctx->comm->allreduce_buf) cudaFree(ctx->comm->allreduce_buf);
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

// ---- one-shot P2P mode: step 1, every rank allocates its exchange buffer and exports an IPC handle; the buffer waits
// in the context (p2p_pending_xchg) for step 2

int pls_comm_p2p_handle(pls_context* ctx, int num_ranks, void* out_handle_64_bytes) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(num_ranks >= 2 && num_ranks <= 16 && out_handle_64_bytes, "pls_comm_p2p_handle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* buf = nullptr;
    const size_t bytes = 2 * (size_t)num_ranks * kP2PSlotBytes;
    PLS_CUDA(cudaMalloc(&buf, bytes));
    PLS_CUDA(cudaMemset(buf, 0, bytes));
    cudaIpcMemHandle_t h;
    PLS_CUDA(cudaIpcGetMemHandle(&h, buf));
    memcpy(out_handle_64_bytes, &h, 64);
    if (ctx->p2p_pending_xchg) cudaFree(ctx->p2p_pending_xchg);
    ctx->p2p_pending_xchg = buf;
    PLS_API_END(ctx)
}

// step 2 (after the host program all-gathered the handles): map every peer's buffer
int pls_comm_p2p_init(pls_context* ctx, int num_ranks, int rank, const void* all_handles) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(num_ranks >= 2 && rank >= 0 && rank < num_ranks && all_handles && ctx->p2p_pending_xchg,
                "pls_comm_p2p_init: call pls_comm_p2p_handle first");
    sync_all(ctx);
    comm_free(ctx);
    Comm* c = new Comm();
    c->p2p = true;
    c->rank = rank;
    c->size = num_ranks;
    c->xchg = ctx->p2p_pending_xchg;
    ctx->p2p_pending_xchg = nullptr;
    c->peer_ptrs.assign(num_ranks, nullptr);
    for (int r = 0; r < num_ranks; ++r) {
        if (r == rank) {
            c->peer_ptrs[r] = c->xchg;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + 64 * (size_t)r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer_ptrs[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            std::string msg = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
            cudaGetLastError();
            for (int o = 0; o < r; ++o)  // unmap what was mapped so far, release this rank's own buffer
                if (o != rank && c->peer_ptrs[o]) cudaIpcCloseMemHandle(c->peer_ptrs[o]);
            cudaFree(c->xchg);
            delete c;
            throw pls::Error{PLS_E_COMM, msg};
        }
    }
    PLS_CUDA(cudaMalloc(&c->seq_dev, sizeof(unsigned long long)));
    PLS_CUDA(cudaMemset(c->seq_dev, 0, sizeof(unsigned long long)));
    PLS_CUDA(cudaMalloc(&c->peer_table_dev, num_ranks * sizeof(void*)));
    PLS_CUDA(cudaMemcpy(c->peer_table_dev, c->peer_ptrs.data(), num_ranks * sizeof(void*), cudaMemcpyHostToDevice));
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
