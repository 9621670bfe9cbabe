// common.cuh — shared declarations for libdsea (sm_100a).  Internal; the public ABI is include/dsea.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dsea.h"

namespace dsea {

void set_error(const char* fmt, ...);

#define DSEA_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            dsea::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                            cudaGetErrorString(_e));                                            \
            return DSEA_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define DSEA_TRY(expr)               \
    do {                             \
        int _s = (expr);             \
        if (_s != DSEA_OK) return _s; \
    } while (0)

#define DSEA_ARG(cond, msg)                                                  \
    do {                                                                     \
        if (!(cond)) {                                                       \
            dsea::set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg); \
            return DSEA_ERR_ARG;                                             \
        }                                                                    \
    } while (0)

constexpr int kMaxRemote = 8;        // log2(max world) remote (top) spin bits
constexpr int kMaxMailRanks = 32;    // peer-memory small all-reduce is used up to this many ranks
constexpr int kMaxSweeps = 40;      // worst case: one new spin bit per sweep (tiny test tiles)
constexpr int kMaxPartialBlocks = 4096;   // upper bound on CTAs writing reduction partials
constexpr int kMaxK = 2048;               // max Lanczos vectors
constexpr int64_t kPartialDoubles = 4 << 20;   // 32 MB of per-CTA partial sums (>= kMaxPartialBlocks * 1024)
// The last 8192 doubles hold the partial sums of a matvec's dot epilogue.  On one GPU the consumer kernels sum
// them in their own prologue (fixed sequential order => bit-identical in every thread) instead of waiting for a
// separate finalize launch; the main region [0, kDotPartialsOffset) stays free for the consumer's own partials.
constexpr int64_t kDotPartialsOffset = kPartialDoubles - 8192;

// ---- device scalars kept in ctx->scal (all double) -------------------------------------------
enum ScalarSlot {
    S_DOT = 0,        // generic dot result
    S_BETA2,          // |r|^2 of the current Lanczos step
    S_INVBETA,        // 1 / beta
    S_RR,             // CG: r.r (current)
    S_RR_NEW,         // CG: r.r (next)
    S_DAD,            // CG: d.Ad
    S_ALPHA,          // CG: step length
    S_BETA,           // CG: direction update coefficient
    S_DONE,           // CG: 1.0 when converged / finished (kernels become no-ops)
    S_ITERS,          // CG: iterations performed
    S_RNORM,          // CG: |r|
    S_KEFF,           // Lanczos: effective k (breakdown truncation), as double
    S_BREAK,          // Lanczos: 1.0 if breakdown occurred
    S_EPS,            // CG: tolerance
    S_MAXIT,          // CG: iteration cap
    S_TMP0,
    S_TMP1,
    S_ALPHA_L,        // Lanczos: alpha_i = q_i . A q_i of the current step
    S_BETAPREV,       // Lanczos: beta_{i-1}
    S_COUNT = 32
};

// peer pointers handed to producing kernels by value
struct PeerPtrs {
    double* p[kMaxRemote];
    int n;
};

// Three-term recurrence applied in the prologue of reorth pass 1: r0 = u - (*alpha) qi - (*beta) qim1,
// written to r0_out (qim1 may be NULL for the first step).
struct Recurrence {
    const void* qi;            // basis columns: double or float according to the kernel's basis element type
    const void* qim1;
    const double* alpha;
    const double* beta;
    double* r0_out;
    const double* alpha_partials = nullptr;   // n_alpha > 0: alpha = sum of these (deferred matvec dot epilogue)
    int n_alpha = 0;
};

struct NcclApi;   // resolved with dlopen at context creation (comm.cu)
struct Profiler;  // optional per-kernel CUDA-event timing (api.cu)

enum ProfKind { PK_MATVEC = 0, PK_REORTH_DOTS, PK_REORTH_UPDATE, PK_RITZ, PK_CG_UPDATE, PK_NORMALISE, PK_TRIDIAG,
                PK_ADJOINT, PK_NOOP, PK_EXCHANGE, PK_COUNT };

}  // namespace dsea

struct dsea_ctx {
    int device = 0;
    int rank = 0;
    int world = 1;
    int log2world = 0;
    int num_sms = 148;
    void* nccl_comm = nullptr;          // ncclComm_t
    dsea::NcclApi* nccl = nullptr;
    cudaStream_t comm_stream = nullptr; // exchanges of top-bit shards run here, overlapped with the local sweep
    cudaEvent_t ev_ready = nullptr, ev_comm = nullptr, ev_poll[2] = {nullptr, nullptr};
    double* scal = nullptr;             // S_COUNT device scalars
    double* partials = nullptr;         // kPartialDoubles doubles of per-CTA reduction partials
    double* cvec = nullptr;             // kMaxK doubles: reorth coefficients / Ritz coefficients
    double* yvec = nullptr;             // 2*kMaxK doubles: tridiagonal eigenvectors (min, max)
    double* tri_work = nullptr;         // tridiagonal solver scratch
    double* pinned = nullptr;           // small pinned host buffer for polling
    unsigned int* counters = nullptr;   // device counters (last-block patterns)
    int64_t launches = 0;
    const double* guard = nullptr;      // device flag consulted by operator kernels (set during CG)
    dsea::Profiler* prof = nullptr;     // non-null while per-kernel timing is enabled
    // Peer-memory exchange arena (CUDA IPC over NVLink): slot j receives the shard of rank ^ (1 << j),
    // written there directly by the PRODUCING kernel of that rank (fused compute + exchange).
    double* arena = nullptr;
    int64_t arena_stride = 0;           // doubles per slot
    double* peer_slot[dsea::kMaxRemote] = {};   // partner j's slot j: where this rank's shard goes
    void* peer_base[dsea::kMaxRemote] = {};
    void* ipc_scratch = nullptr;
    bool p2p_ok = false;
    bool p2p_disabled = false;
    bool fresh_collective = true;       // a collective completed since the arena was last read (WAR guard)
    // Mailbox for small all-reduces over peer memory (every rank maps every other rank's mailbox).
    void* mail_local = nullptr;
    void* mail_peer[dsea::kMaxMailRanks] = {};
    bool mail_ok = false;
    bool mail_disabled = false;
    unsigned long long mail_seq = 0;
    // options
    int tfim_tile_bits = 13;
    int tfim_run_bits = 0;              // 0 = auto
    int tfim_pipeline = 1;              // persistent double-buffered sweep kernel for full 2^13 tiles
    int tfim_tma = 1;                   // stage contiguous tiles with TMA bulk copies (UBLKCP + mbarrier) instead of LDGSTS
    int tfim_pipe_threads = 512;        // 256 / 512 / 1024 threads: 16 / 8 / 4 pairs per thread (5 / 4 / 3 register-resident tile bits)
    int tfim_stage = 1;                 // strided sweeps stage their epilogue operands in thread-private shared-memory slots
    int tfim_generic_min_operands = 4;  // last sweep: direct bits + remote bits from which the generic kernel is used
    int tfim_unroll = 1;                // compile-time flip-bit range in the pipelined kernel (fully unrolled LDS loop)
    int tfim_direct = 1;                // top local bits beyond two sweeps by direct (L2-served) loads instead of a third sweep
    int tfim_fuse_scale = 1;            // fold the Lanczos normalisation q = r / beta into the first matvec sweep
    int tfim_l2_prefetch = 0;           // prefetch.global.L2 the next tile's epilogue operands
    int tfim_pipe_adjoint = 1;          // adjoint contraction (K6) through the pipelined kernel
    int tfim_pipe_remote = 1;           // last sweep of a sharded matvec through the pipelined kernel
    int cg_fuse_push = 1;               // CG direction update stores d into the partners' arenas
    int cg_check_every = 16;
    int reorth_ctas_per_sm = 8;         // persistent CTAs per SM for the reorth GEMVs (measured best of 2..8)
    double polish_eps = 1e-10;          // absolute CG tolerance of the Jacobi-Davidson polish (fp32 basis)
    int64_t last_polish_iters = 0;
    int pdl_staged = 0;                 // experiment knob: bit 0 = staged sweep launched with the PDL attribute, bit 1 = it triggers its dependents early
    int pdl = 0;                        // programmatic dependent launch for every kernel (set to 1 by ctx_create when world == 1)
    int fuse_small = 1;                 // one GPU: consumers sum matvec / norm partials themselves (no finalize launches)
    int pending_dot_n = 0;              // > 0: the last matvec left this many dot partials at partials + kDotPartialsOffset
    int pending_norm_n = 0;             // > 0: the last reorth pass 2 left this many |r|^2 partials at partials
    int basis_fp32 = 0;                 // opt-in: Lanczos basis also kept as an fp32 shadow that the reorth passes stream
};

struct dsea_op {
    int kind = 0;
    dsea_ctx* ctx = nullptr;
    int64_t n_loc = 0;
    // TFIM
    int N = 0;
    int L = 0;                  // local bits = N - log2(world)
    // CSR
    int64_t nnz = 0;
    const int64_t* rowptr = nullptr;
    const int64_t* colidx = nullptr;
    const double* vals = nullptr;
    // dense
    const double* A = nullptr;
    int64_t ld = 0;
};

namespace dsea {

// ---- internal kernels' host launchers (each returns a DSEA status) ----------------------------
// tfim.cu
// `exchange` (sharded runs with the peer arena): XCH_PUSH = publish v with barrier + push kernel + barrier;
// XCH_PREPUSHED = the kernel that produced v already stored it into the partners' arenas and a collective has
// completed since (Lanczos: reorth pass 2 followed by the beta^2 reduction); XCH_PREPUSHED_BARRIER = stored by the
// producer, but no collective followed: one barrier is issued before the last sweep (CG direction update).
// The arena copy may be unscaled: the remote terms are multiplied by *remote_scale.
// `in_scale` / `q_out`: the logical input is (*in_scale) * v and is also written to q_out (fused normalisation of
// the new Lanczos vector); only when tfim_can_fuse_scale().
enum { XCH_PUSH = 0, XCH_PREPUSHED = 1, XCH_PREPUSHED_BARRIER = 2 };
int tfim_apply(dsea_ctx* ctx, const dsea_op* op, const double* g, const double* shift, const double* v,
               double* u, const double* dotw, double* dot_out, double* work, cudaStream_t st,
               int exchange = XCH_PUSH, const double* remote_scale = nullptr, const double* in_scale = nullptr,
               double* q_out = nullptr, bool round_remote = false, bool defer_dot = false);
bool tfim_can_fuse_scale(const dsea_ctx* ctx, const dsea_op* op);
int tfim_dHdg(dsea_ctx* ctx, const dsea_op* op, const double* v, double* u, double* work, cudaStream_t st);
int tfim_adjoint(dsea_ctx* ctx, const dsea_op* op, const double* v1, const double* v2, double* out,
                 double* work, cudaStream_t st);
// spmv.cu
int csr_apply(dsea_ctx* ctx, const dsea_op* op, const double* pdiag, const double* shift, const double* v,
              double* u, double* dot_out, cudaStream_t st);
int dense_apply(dsea_ctx* ctx, const dsea_op* op, const double* shift, const double* v, double* u,
                double* dot_out, cudaStream_t st);
int hadamard(dsea_ctx* ctx, int64_t n, const double* a, const double* b, double* out, cudaStream_t st);
int outer(dsea_ctx* ctx, int64_t n, double scale, const double* a, const double* b, double* out, cudaStream_t st);
// reorth.cu
int reorth_dots(dsea_ctx* ctx, int64_t n, int64_t ldq, int ncols, const double* Q, const double* u, double* c_out,
                cudaStream_t st, const Recurrence* rec = nullptr);   // c_out = Q^T (u [- recurrence terms])  (allreduced)
// `defer_norm`: leave the |r|^2 partials in ctx->partials (ctx->pending_norm_n) for the consumer to sum
int reorth_update(dsea_ctx* ctx, int64_t n, int64_t ldq, int ncols, const double* Q, const double* u,
                  const double* c, double sign, double* r_out, double* norm2_out,
                  cudaStream_t st, const PeerPtrs* peers = nullptr, bool defer_norm = false);   // r = u + sign * Q c  (u may be NULL)
int scale_by_inv_sqrt(dsea_ctx* ctx, int64_t n, double* x, const double* norm2, cudaStream_t st);
// fp32 shadow basis (opt-in "basis_fp32"): the same two passes over float columns with fp64 accumulation, and the
// normalisation that rounds the new vector to fp32 (q32 = fl32(r * *scale), q64 = the same values widened)
int reorth_dots_f32(dsea_ctx* ctx, int64_t n, int64_t ldq, int ncols, const float* Q, const double* u, double* c_out,
                    cudaStream_t st, const Recurrence* rec = nullptr);
int reorth_update_f32(dsea_ctx* ctx, int64_t n, int64_t ldq, int ncols, const float* Q, const double* u,
                      const double* c, double sign, double* r_out, double* norm2_out, cudaStream_t st,
                      const PeerPtrs* peers = nullptr, bool defer_norm = false);
int scale_round_store(dsea_ctx* ctx, int64_t n, const double* r, const double* scale, double* q64, float* q32,
                      cudaStream_t st);
// blas1.cu
int dot(dsea_ctx* ctx, int64_t n, const double* a, const double* b, double* out, cudaStream_t st);
int axpby(dsea_ctx* ctx, int64_t n, const double* a, const double* x, const double* b, double* y, cudaStream_t st);
int project(dsea_ctx* ctx, int64_t n, const double* psi, const double* b, double* out, cudaStream_t st);
int randn(dsea_ctx* ctx, int64_t n, uint64_t seed, uint64_t sid, uint64_t offset, double* out, cudaStream_t st);
int finalize_partials(dsea_ctx* ctx, int nblocks, int ncols, double* out, cudaStream_t st, const double* src = nullptr);
// second reduction stage + cross-rank sum in one step (fused peer-memory kernel when available)
int finalize_reduce(dsea_ctx* ctx, int nblocks, int ncols, double* out, cudaStream_t st,
                    const double* src = nullptr);   // comm.cu; src defaults to ctx->partials
// tridiag.cu
int tridiag_extreme(dsea_ctx* ctx, int k, int which, const double* alpha, const double* beta, const double* keff,
                    double* evals, double* y_min, double* y_max, cudaStream_t st);
// cg.cu
int cg_setup(dsea_ctx* ctx, double eps, int64_t maxit, cudaStream_t st);
// `peers`: when non-null the kernels that write the search direction d also store it into the partners' arenas
int cg_init(dsea_ctx* ctx, int64_t n, const double* b, const double* Ax, double* r, double* d, cudaStream_t st,
            const PeerPtrs* peers = nullptr);
// `n_dad` > 0: d.Ad is the sum of that many partials at ctx->partials + kDotPartialsOffset (deferred matvec epilogue)
int cg_iterate(dsea_ctx* ctx, int64_t n, double* x, double* r, double* d, const double* Ad, cudaStream_t st,
               const PeerPtrs* peers = nullptr, int n_dad = 0);
// comm.cu
int comm_init(dsea_ctx* ctx, const void* id);
int comm_destroy(dsea_ctx* ctx);
int comm_unique_id(void* id128);
int allreduce_sum(dsea_ctx* ctx, double* buf, int64_t count, cudaStream_t st);
int exchange_shards(dsea_ctx* ctx, const double* send, double* recv_base, int64_t n_loc, cudaStream_t st);
int mailbox_setup(dsea_ctx* ctx);                 // collective, at context creation
int p2p_setup(dsea_ctx* ctx, int64_t n_loc);      // collective; falls back to NCCL send/recv if IPC is unavailable
int p2p_teardown(dsea_ctx* ctx);
int comm_barrier(dsea_ctx* ctx, cudaStream_t st);
int push_to_peers(dsea_ctx* ctx, const double* v, int64_t n, cudaStream_t st);   // blas1.cu
inline PeerPtrs peer_ptrs(const dsea_ctx* ctx) {
    PeerPtrs pp;
    pp.n = ctx->p2p_ok ? ctx->log2world : 0;
    for (int j = 0; j < kMaxRemote; ++j) pp.p[j] = j < pp.n ? ctx->peer_slot[j] : nullptr;
    return pp;
}

inline void count_launch(dsea_ctx* ctx, int n = 1) { ctx->launches += n; }

#ifdef __CUDACC__
// Every kernel of the library is launched through launch_k: with "pdl" on (default on one GPU) the launch carries the
// programmatic-stream-serialization attribute, so its CTAs may be scheduled while the previous kernel of the stream is
// still draining; each kernel starts with pdl_prologue() = griddepcontrol.launch_dependents (let MY successor be
// scheduled early too) + griddepcontrol.wait (block until the predecessor grid has completed and its writes are
// visible) BEFORE its first global-memory access.  That hides the 2-3 us launch latency between the ~2500 dependent
// kernels of a solve without changing any ordering.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_opt(const dsea_ctx* ctx, bool allow_pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                                size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (ctx && ctx->pdl && allow_pdl) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(const dsea_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t st, Args&&... args) {
    return launch_k_opt(ctx, true, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
#endif

// Per-kernel timing with CUDA events on the launching stream (no-ops unless dsea_profile_enable()).
int prof_begin(dsea_ctx* ctx, int kind, double algorithmic_bytes, cudaStream_t st);
void prof_end(dsea_ctx* ctx, int token, cudaStream_t st);
void prof_guard_key(dsea_ctx* ctx, int64_t key);      // tags the following records as guarded step `key` (-1: none)
void prof_guard_next_phase(dsea_ctx* ctx);            // key += 1 (direction update of the same CG iteration)

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 doubles.  Fixed order => deterministic.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();            // protect `red` from a previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

// L1 hazard (found the hard way: CG stopped converging with PDL on).  The per-launch L1 invalidation of a
// programmatically launched grid happens when ITS CTAs are scheduled, i.e. possibly while the predecessor is still
// running; a line the predecessor's CTAs load afterwards (the scalar block, a partial-sum array) can then survive in
// that SM's L1 and be served, stale, to this grid after griddepcontrol.wait.  A gpu-scope fence compiles to
// MEMBAR + CCTL.IVALL, which invalidates the SM's L1 after the wait, restoring the usual launch-boundary semantics.
__device__ __forceinline__ void pdl_prologue(bool trigger = true) {
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __threadfence();
}

// Sum of n partials computed redundantly by every (full) warp with a fixed lane-strided + xor-shuffle order: every
// thread of every CTA that calls it obtains the same bits.  Must be called by all 32 lanes of the warp.
__device__ __forceinline__ double sum_partials_seq(const double* __restrict__ p, int n) {
    double s = 0.0;
    for (int i = threadIdx.x & 31; i < n; i += 32) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

__device__ __forceinline__ double2 ldg2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void stg2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// streaming (evict-first) 16-byte load for data touched once per kernel
__device__ __forceinline__ double2 ldg2_stream(const double* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

#endif  // __CUDACC__

}  // namespace dsea
