// comm.cu — NCCL plumbing for the sharded path (one process per GPU, NVLink 5 / NVSwitch).
//
// The 2^N state vector is sharded by its top log2(world) spin bits.  Two collectives exist:
//   exchange_shards : for every top bit j, swap the whole local shard with rank ^ (1 << j)
//                     (ncclSend/ncclRecv grouped; NVSwitch makes every partner equidistant)
//   allreduce_sum   : fp64 sums of dot products / norms / the reorth coefficient vector
// NCCL is resolved with dlopen so that a single-GPU process needs no NCCL at all and a torch process
// reuses the libnccl.so.2 torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace dsea {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* load_nccl() {
    static NcclApi api;
    if (api.handle) return &api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return nullptr;
    }
#define DSEA_SYM(field, name)                                   \
    *(void**)(&api.field) = dlsym(h, name);                     \
    if (!api.field) {                                           \
        set_error("libnccl is missing symbol %s", name);        \
        return nullptr;                                         \
    }
    DSEA_SYM(GetUniqueId, "ncclGetUniqueId")
    DSEA_SYM(CommInitRank, "ncclCommInitRank")
    DSEA_SYM(CommDestroy, "ncclCommDestroy")
    DSEA_SYM(AllReduce, "ncclAllReduce")
    DSEA_SYM(Send, "ncclSend")
    DSEA_SYM(Recv, "ncclRecv")
    DSEA_SYM(AllGather, "ncclAllGather")
    DSEA_SYM(GroupStart, "ncclGroupStart")
    DSEA_SYM(GroupEnd, "ncclGroupEnd")
    DSEA_SYM(GetErrorString, "ncclGetErrorString")
#undef DSEA_SYM
    api.handle = h;
    return &api;
}

#define DSEA_NCCL(api, expr)                                                                         \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            set_error("%s:%d NCCL error: %s", __FILE__, __LINE__, (api)->GetErrorString(_r));         \
            return DSEA_ERR_NCCL;                                                                    \
        }                                                                                            \
    } while (0)

int comm_unique_id(void* id128) {
    NcclApi* api = load_nccl();
    if (!api) return DSEA_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    DSEA_NCCL(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return DSEA_OK;
}

int comm_init(dsea_ctx* ctx, const void* id128) {
    if (ctx->world == 1) return DSEA_OK;
    NcclApi* api = load_nccl();
    if (!api) return DSEA_ERR_NCCL;
    DSEA_ARG(id128 != nullptr, "world > 1 needs an ncclUniqueId");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    DSEA_NCCL(api, api->CommInitRank(&comm, ctx->world, id, ctx->rank));
    ctx->nccl = api;
    ctx->nccl_comm = comm;
    return DSEA_OK;
}

int comm_destroy(dsea_ctx* ctx) {
    p2p_teardown(ctx);
    if (ctx->ipc_scratch) cudaFree(ctx->ipc_scratch);
    ctx->ipc_scratch = nullptr;
    if (ctx->nccl_comm && ctx->nccl) {
        ctx->nccl->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return DSEA_OK;
}

int allreduce_sum(dsea_ctx* ctx, double* buf, int64_t count, cudaStream_t st) {
    if (ctx->world == 1) return DSEA_OK;
    NcclApi* api = ctx->nccl;
    DSEA_NCCL(api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, st));
    ctx->fresh_collective = true;
    return DSEA_OK;
}

// A 1-double allreduce used purely as a cross-rank, stream-ordered barrier.
int comm_barrier(dsea_ctx* ctx, cudaStream_t st) { return allreduce_sum(ctx, ctx->scal + S_TMP1, 1, st); }

// ---- peer-memory arena ---------------------------------------------------------------------------
// Every rank cudaMalloc's an arena of log2(world) slots, publishes its cudaIpcMemHandle with an NCCL
// allgather and maps its partners' arenas.  Afterwards a kernel on rank r can store its shard straight
// into slot j of rank r ^ (1 << j) over NVLink, which lets the exchange ride on the kernel that produces
// the vector (reorth pass 2) instead of being a separate NCCL send/recv after it.
int p2p_teardown(dsea_ctx* ctx) {
    if (!ctx->arena) return DSEA_OK;
    cudaDeviceSynchronize();
    for (int j = 0; j < kMaxRemote; ++j) {
        if (ctx->peer_base[j]) cudaIpcCloseMemHandle(ctx->peer_base[j]);
        ctx->peer_base[j] = nullptr;
        ctx->peer_slot[j] = nullptr;
    }
    if (ctx->world > 1 && ctx->nccl_comm) {       // nobody may still have this arena mapped when it is freed
        comm_barrier(ctx, ctx->comm_stream);
        cudaStreamSynchronize(ctx->comm_stream);
    }
    cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_stride = 0;
    ctx->p2p_ok = false;
    return DSEA_OK;
}

int p2p_setup(dsea_ctx* ctx, int64_t n_loc) {
    if (ctx->world == 1 || ctx->p2p_disabled) return DSEA_OK;
    const int64_t stride = (n_loc + 15) & ~(int64_t)15;
    if (ctx->arena && ctx->arena_stride >= stride) return DSEA_OK;
    DSEA_TRY(p2p_teardown(ctx));
    NcclApi* api = ctx->nccl;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    cudaStream_t st = ctx->comm_stream;
    double ok = 1.0;
    if (cudaMalloc(&ctx->arena, (size_t)ctx->log2world * stride * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        ctx->arena = nullptr;
        ok = 0.0;
    }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok != 0.0 && cudaIpcGetMemHandle(&mine, ctx->arena) != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    if (!ctx->ipc_scratch) DSEA_CUDA(cudaMalloc(&ctx->ipc_scratch, 64 * 256));
    char* dbuf = (char*)ctx->ipc_scratch;
    DSEA_CUDA(cudaMemcpyAsync(dbuf + 64 * ctx->rank, &mine, 64, cudaMemcpyHostToDevice, st));
    DSEA_NCCL(api, api->AllGather(dbuf + 64 * ctx->rank, dbuf, 64, ncclChar, comm, st));
    cudaIpcMemHandle_t all[256];
    DSEA_CUDA(cudaMemcpyAsync(all, dbuf, 64 * (size_t)ctx->world, cudaMemcpyDeviceToHost, st));
    DSEA_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < ctx->log2world && ok != 0.0; ++j) {
        const int peer = ctx->rank ^ (1 << j);
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, all[peer], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0.0;
            break;
        }
        ctx->peer_base[j] = base;
        ctx->peer_slot[j] = (double*)base + (int64_t)j * stride;
    }
    // all ranks must agree on the mode
    double* flag = ctx->scal + S_TMP1;
    DSEA_CUDA(cudaMemcpyAsync(flag, &ok, sizeof(double), cudaMemcpyHostToDevice, st));
    DSEA_NCCL(api, api->AllReduce(flag, flag, 1, ncclDouble, ncclMin, comm, st));
    double agreed = 0.0;
    DSEA_CUDA(cudaMemcpyAsync(&agreed, flag, sizeof(double), cudaMemcpyDeviceToHost, st));
    DSEA_CUDA(cudaStreamSynchronize(st));
    if (agreed != 0.0) {
        ctx->arena_stride = stride;
        ctx->p2p_ok = true;
        ctx->fresh_collective = true;
    } else {
        for (int j = 0; j < kMaxRemote; ++j) {
            if (ctx->peer_base[j]) cudaIpcCloseMemHandle(ctx->peer_base[j]);
            ctx->peer_base[j] = nullptr;
            ctx->peer_slot[j] = nullptr;
        }
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->p2p_ok = false;
    }
    return DSEA_OK;
}

int exchange_shards(dsea_ctx* ctx, const double* send, double* recv_base, int64_t n_loc, cudaStream_t st) {
    if (ctx->world == 1) return DSEA_OK;
    NcclApi* api = ctx->nccl;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    DSEA_NCCL(api, api->GroupStart());
    for (int j = 0; j < ctx->log2world; ++j) {
        const int peer = ctx->rank ^ (1 << j);
        DSEA_NCCL(api, api->Send(send, (size_t)n_loc, ncclDouble, peer, comm, st));
        DSEA_NCCL(api, api->Recv(recv_base + (int64_t)j * n_loc, (size_t)n_loc, ncclDouble, peer, comm, st));
    }
    DSEA_NCCL(api, api->GroupEnd());
    return DSEA_OK;
}

}  // namespace dsea
