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

static void mailbox_teardown(dsea_ctx* ctx);

int comm_destroy(dsea_ctx* ctx) {
    p2p_teardown(ctx);
    mailbox_teardown(ctx);
    if (ctx->ipc_scratch) cudaFree(ctx->ipc_scratch);
    ctx->ipc_scratch = nullptr;
    if (ctx->nccl_comm && ctx->nccl) {
        ctx->nccl->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return DSEA_OK;
}

// ---- small all-reduce over peer memory ---------------------------------------------------------------
// Every reduction in the Lanczos / CG loops is a handful of doubles (<= k).  Instead of a finalize kernel
// followed by an NCCL all-reduce, ONE kernel sums the per-CTA partials, stores the rank's value straight
// into every peer's mailbox over NVLink (value, then sequence number with release semantics), waits for
// the peers' values (acquire) and adds them in rank order — so every rank obtains bit-identical sums,
// and the call also acts as a barrier.  Two mailbox parities make back-to-back calls safe: a rank can
// only reach call s+2 after every peer has consumed call s (their value for s+1 is sent afterwards).
struct MailSlot {
    double val;
    unsigned long long seq;
};
struct MailParams {
    MailSlot* peer[kMaxMailRanks];
    const MailSlot* mine;
    int world, rank;
    unsigned long long seq;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(128)
finalize_allreduce_kernel(const double* __restrict__ partials, int nblocks, int ncols, double* __restrict__ out,
                          const MailParams mp) {
    pdl_prologue();
    __shared__ double red[32];
    __shared__ double vals[kMaxMailRanks];
    const int col = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[(size_t)b * ncols + col];
    s = block_sum(s, red);
    if (threadIdx.x == 0) vals[mp.rank] = s;
    __syncthreads();
    const int par = (int)(mp.seq & 1ull);
    const int t = threadIdx.x;
    if (t < mp.world && t != mp.rank) {
        MailSlot* dst = mp.peer[t] + ((size_t)(mp.rank * 2 + par) * kMaxK + col);
        dst->val = vals[mp.rank];
        st_release_sys(&dst->seq, mp.seq);
        const MailSlot* src = mp.mine + ((size_t)(t * 2 + par) * kMaxK + col);
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(&src->seq) != mp.seq) {
            if (clock64() - t0 > 240000000000ll) { ok = false; break; }     // ~2 min: a peer died; do not hang the GPU
        }
        vals[t] = ok ? *reinterpret_cast<const volatile double*>(&src->val) : __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int r = 0; r < mp.world; ++r) tot += vals[r];
        out[col] = tot;
    }
}

static int mail_launch(dsea_ctx* ctx, const double* partials, int nblocks, int ncols, double* out, cudaStream_t st) {
    MailParams mp;
    for (int r = 0; r < kMaxMailRanks; ++r) mp.peer[r] = (MailSlot*)ctx->mail_peer[r];
    mp.mine = (const MailSlot*)ctx->mail_local;
    mp.world = ctx->world;
    mp.rank = ctx->rank;
    mp.seq = ++ctx->mail_seq;
    launch_k(ctx, finalize_allreduce_kernel, dim3(ncols), dim3(128), 0, st, partials, nblocks, ncols, out, mp);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    ctx->fresh_collective = true;
    return DSEA_OK;
}

int finalize_reduce(dsea_ctx* ctx, int nblocks, int ncols, double* out, cudaStream_t st, const double* src) {
    if (!src) src = ctx->partials;
    if (ctx->world > 1 && ctx->mail_ok && ncols <= kMaxK) return mail_launch(ctx, src, nblocks, ncols, out, st);
    DSEA_TRY(finalize_partials(ctx, nblocks, ncols, out, st, src));
    return allreduce_sum(ctx, out, ncols, st);
}

int allreduce_sum(dsea_ctx* ctx, double* buf, int64_t count, cudaStream_t st) {
    if (ctx->world == 1) return DSEA_OK;
    if (ctx->mail_ok && count <= kMaxK) return mail_launch(ctx, buf, 1, (int)count, buf, st);
    NcclApi* api = ctx->nccl;
    DSEA_NCCL(api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, st));
    ctx->fresh_collective = true;
    return DSEA_OK;
}

// A 1-double allreduce used purely as a cross-rank, stream-ordered barrier.
int comm_barrier(dsea_ctx* ctx, cudaStream_t st) {
    if (ctx->world == 1) return DSEA_OK;
    DSEA_CUDA(cudaMemsetAsync(ctx->scal + S_TMP1, 0, sizeof(double), st));
    return allreduce_sum(ctx, ctx->scal + S_TMP1, 1, st);
}

// Maps every peer's mailbox (collective; called once when the context is created).
int mailbox_setup(dsea_ctx* ctx) {
    if (ctx->world == 1 || ctx->world > kMaxMailRanks || ctx->mail_disabled) return DSEA_OK;
    NcclApi* api = ctx->nccl;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    cudaStream_t st = ctx->comm_stream;
    const size_t bytes = (size_t)ctx->world * 2 * kMaxK * sizeof(MailSlot);
    double ok = 1.0;
    if (cudaMalloc(&ctx->mail_local, bytes) != cudaSuccess) {
        cudaGetLastError();
        ctx->mail_local = nullptr;
        ok = 0.0;
    } else {
        DSEA_CUDA(cudaMemset(ctx->mail_local, 0, bytes));
    }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok != 0.0 && cudaIpcGetMemHandle(&mine, ctx->mail_local) != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
    }
    if (!ctx->ipc_scratch) DSEA_CUDA(cudaMalloc(&ctx->ipc_scratch, 64 * 256));
    char* dbuf = (char*)ctx->ipc_scratch;
    DSEA_CUDA(cudaMemcpyAsync(dbuf + 64 * ctx->rank, &mine, 64, cudaMemcpyHostToDevice, st));
    DSEA_NCCL(api, api->AllGather(dbuf + 64 * ctx->rank, dbuf, 64, ncclChar, comm, st));
    cudaIpcMemHandle_t all[256];
    DSEA_CUDA(cudaMemcpyAsync(all, dbuf, 64 * (size_t)ctx->world, cudaMemcpyDeviceToHost, st));
    DSEA_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < ctx->world && ok != 0.0; ++r) {
        if (r == ctx->rank) continue;
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0.0;
            break;
        }
        ctx->mail_peer[r] = base;
    }
    double* flag = ctx->scal + S_TMP1;
    DSEA_CUDA(cudaMemcpyAsync(flag, &ok, sizeof(double), cudaMemcpyHostToDevice, st));
    DSEA_NCCL(api, api->AllReduce(flag, flag, 1, ncclDouble, ncclMin, comm, st));
    double agreed = 0.0;
    DSEA_CUDA(cudaMemcpyAsync(&agreed, flag, sizeof(double), cudaMemcpyDeviceToHost, st));
    DSEA_CUDA(cudaStreamSynchronize(st));
    ctx->mail_ok = (agreed != 0.0);
    ctx->mail_seq = 0;
    if (!ctx->mail_ok) {
        for (int r = 0; r < kMaxMailRanks; ++r) {
            if (ctx->mail_peer[r]) cudaIpcCloseMemHandle(ctx->mail_peer[r]);
            ctx->mail_peer[r] = nullptr;
        }
        if (ctx->mail_local) cudaFree(ctx->mail_local);
        ctx->mail_local = nullptr;
    }
    return DSEA_OK;
}

static void mailbox_teardown(dsea_ctx* ctx) {
    if (!ctx->mail_local) return;
    cudaDeviceSynchronize();
    NcclApi* api = ctx->nccl;
    ctx->mail_ok = false;                      // the closing barrier below goes through NCCL
    for (int r = 0; r < kMaxMailRanks; ++r) {
        if (ctx->mail_peer[r]) cudaIpcCloseMemHandle(ctx->mail_peer[r]);
        ctx->mail_peer[r] = nullptr;
    }
    if (api && ctx->nccl_comm) {
        api->AllReduce(ctx->scal + S_TMP1, ctx->scal + S_TMP1, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm,
                       ctx->comm_stream);
        cudaStreamSynchronize(ctx->comm_stream);
    }
    cudaFree(ctx->mail_local);
    ctx->mail_local = nullptr;
}

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
