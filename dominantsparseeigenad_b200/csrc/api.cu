// api.cu — extern "C" boundary of libdsea.so (declared in include/dsea.h) and the host-side
// orchestration of the device-resident Lanczos and CG loops.  No kernel lives here except the few
// single-thread bookkeeping kernels that keep the loops free of host synchronisation.
#include <stdarg.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace dsea {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct ProfRec {
    cudaEvent_t a, b;
    int kind;
    double bytes;
    int64_t guard_key;             // >= 0: launched under the CG done flag as step `guard_key` of the running solve
};
struct Profiler {
    std::vector<ProfRec> recs;     // event pool, reused across collections
    size_t used = 0;
    int64_t guard_key = -1;        // set by dsea_cg around the launches it guards (2*iteration + phase)
};

int prof_begin(dsea_ctx* ctx, int kind, double bytes, cudaStream_t st) {
    Profiler* p = ctx->prof;
    if (!p) return -1;
    if (p->used == p->recs.size()) {
        ProfRec r;
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
        p->recs.push_back(r);
    }
    ProfRec& r = p->recs[p->used];
    r.kind = kind;
    r.bytes = bytes;
    r.guard_key = p->guard_key;
    cudaEventRecord(r.a, st);
    return (int)p->used++;
}

void prof_end(dsea_ctx* ctx, int token, cudaStream_t st) {
    if (token < 0 || !ctx->prof) return;
    cudaEventRecord(ctx->prof->recs[token].b, st);
}

void prof_guard_key(dsea_ctx* ctx, int64_t key) {
    if (ctx->prof) ctx->prof->guard_key = key;
}

// Launches guarded by the CG done flag exit immediately once it is set.  After the solve the executed
// iteration count is known, so every record issued for a later step is re-labelled PK_NOOP: it keeps its
// (tiny) elapsed time but is credited no bytes and does not count as a launch of its kernel.
void prof_retire_guarded(dsea_ctx* ctx, size_t first, int64_t first_noop_key) {
    Profiler* p = ctx->prof;
    if (!p) return;
    for (size_t i = first; i < p->used; ++i) {
        ProfRec& r = p->recs[i];
        if (r.guard_key >= 0 && r.guard_key >= first_noop_key) {
            r.kind = PK_NOOP;
            r.bytes = 0.0;
        }
        r.guard_key = -1;
    }
}

void prof_guard_next_phase(dsea_ctx* ctx) {
    if (ctx->prof && ctx->prof->guard_key >= 0) ctx->prof->guard_key += 1;
}

size_t prof_mark(const dsea_ctx* ctx) { return ctx->prof ? ctx->prof->used : 0; }

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

static int apply_op(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift, const double* v,
                    double* u, double* dot_out, double* work, cudaStream_t st, int exchange = XCH_PUSH,
                    const double* remote_scale = nullptr, const double* in_scale = nullptr, double* q_out = nullptr,
                    bool round_remote = false, bool defer_dot = false) {
    ctx->pending_dot_n = 0;
    switch (op->kind) {
        case DSEA_OP_TFIM:
            DSEA_ARG(param != nullptr, "TFIM operator needs the device scalar g");
            return tfim_apply(ctx, op, param, shift, v, u, nullptr, dot_out, work, st, exchange, remote_scale, in_scale,
                              q_out, round_remote, defer_dot);
        case DSEA_OP_CSR:
            return csr_apply(ctx, op, param, shift, v, u, dot_out, st);
        case DSEA_OP_DENSE:
            return dense_apply(ctx, op, shift, v, u, dot_out, st);
    }
    set_error("unknown operator kind %d", op->kind);
    return DSEA_ERR_ARG;
}

// ---- Lanczos bookkeeping (single thread) -------------------------------------------------------
__global__ void lanczos_reset_kernel(double* scal) {
    pdl_prologue();
    scal[S_KEFF] = 0.0;
    scal[S_BREAK] = 0.0;
}

// alpha[i] = q_i . A q_i (Lanczos.py:55,72); beta[i] = |r| after re-orthogonalisation (:69); flags breakdown once.
// n_alpha / n_beta2 > 0 (one GPU): the kernel first sums the deferred partials itself — alpha with the same sequential
// sum reorth pass 1 used in its prologue (bit-identical), |r|^2 with a fixed-order block reduction — which replaces two
// finalize launches per Lanczos step.
__global__ void __launch_bounds__(256)
lanczos_record_kernel(double* scal, double* alpha, double* beta, int i, int has_beta, const double* __restrict__ alpha_partials,
                      int n_alpha, const double* __restrict__ beta2_partials, int n_beta2) {
    pdl_prologue();
    __shared__ double red[32];
    if (n_beta2 > 0) {
        double s = 0.0;
        for (int j = threadIdx.x; j < n_beta2; j += blockDim.x) s += beta2_partials[j];
        s = block_sum(s, red);
        if (threadIdx.x == 0) scal[S_BETA2] = s;
    }
    const double a_sum = n_alpha > 0 ? sum_partials_seq(alpha_partials, n_alpha) : 0.0;   // whole warps take part
    if (threadIdx.x != 0) return;
    if (n_alpha > 0) scal[S_ALPHA_L] = a_sum;
    alpha[i] = scal[S_ALPHA_L];
    if (has_beta) {
        const double b2 = scal[S_BETA2];
        const double b = b2 > 0.0 ? sqrt(b2) : 0.0;
        beta[i] = b;
        scal[S_BETAPREV] = b;
        scal[S_INVBETA] = (b > 0.0 && isfinite(b)) ? 1.0 / b : 0.0;
        if (!(b > 0.0) || !isfinite(b)) {
            if (scal[S_BREAK] == 0.0) {
                scal[S_BREAK] = 1.0;
                scal[S_KEFF] = (double)(i + 1);
            }
            scal[S_BETA2] = 0.0;      // makes the normalisation write a zero column
        }
    }
}

static int lanczos_start_impl(dsea_ctx* ctx, int64_t n, double* Q, cudaStream_t st) {
    launch_k(ctx, lanczos_reset_kernel, dim3(1), dim3(1), 0, st, ctx->scal);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    DSEA_TRY(dot(ctx, n, Q, Q, ctx->scal + S_BETA2, st));
    return scale_by_inv_sqrt(ctx, n, Q, ctx->scal + S_BETA2, st);          // Lanczos.py:53
}

// One Lanczos step exactly as the reference orders it (Lanczos.py:61-75):
//   r0 = u - alpha_i q_i - beta_{i-1} q_{i-1}      (formed in the prologue of pass 1, written to Q[:, i+1])
//   c  = Q[:, :i+1]^T r0 ; r = r0 - Q[:, :i+1] c    (one classical Gram-Schmidt sweep, two fused passes)
//   beta_i = |r| ; q_{i+1} = r / beta_i
// Removing the two large components with the recurrence scalars BEFORE the sweep is what keeps the
// orthogonality defect from being re-amplified by |alpha| / beta every step (a sweep applied to u
// itself is unstable once beta becomes small, e.g. for k close to the Krylov dimension).
// `alpha_ready`: scal[S_ALPHA_L] already holds q_i . u (epilogue of the matvec); otherwise it is computed here.
// `push_next`: pass 2 also stores the new (un-normalised) vector into the partners' arenas, so the next
// matvec finds its remote shards already in place (the beta^2 reduction that follows is the barrier).
// `normalise`: scale the new column by 1 / beta here (K3); false when the next matvec's first sweep does it (fused).
static int lanczos_step_impl(dsea_ctx* ctx, int64_t n, int64_t ldq, int k, int i, double* Q, const double* u,
                             double* alpha, double* beta, cudaStream_t st, bool alpha_ready, bool push_next = false,
                             bool normalise = true) {
    const int m = i + 1;
    const double* qi = Q + (int64_t)i * ldq;
    const int n_alpha = alpha_ready ? ctx->pending_dot_n : 0;      // > 0: the matvec deferred its dot epilogue to us
    ctx->pending_dot_n = 0;
    if (!alpha_ready) DSEA_TRY(dot(ctx, n, qi, u, ctx->scal + S_ALPHA_L, st));
    const bool more = (i < k - 1);
    int n_beta2 = 0;
    if (more) {
        double* qnext = Q + (int64_t)m * ldq;
        Recurrence rec;
        rec.qi = qi;
        rec.qim1 = i > 0 ? Q + (int64_t)(i - 1) * ldq : nullptr;
        rec.alpha = ctx->scal + S_ALPHA_L;
        rec.beta = ctx->scal + S_BETAPREV;
        rec.r0_out = qnext;
        rec.alpha_partials = ctx->partials + kDotPartialsOffset;
        rec.n_alpha = n_alpha;
        DSEA_TRY(reorth_dots(ctx, n, ldq, m, Q, u, ctx->cvec, st, &rec));
        PeerPtrs pp = peer_ptrs(ctx);
        DSEA_TRY(reorth_update(ctx, n, ldq, m, Q, qnext, ctx->cvec, -1.0, qnext, ctx->scal + S_BETA2, st,
                               push_next ? &pp : nullptr, /*defer_norm=*/true));
        n_beta2 = ctx->pending_norm_n;
        ctx->pending_norm_n = 0;
    }
    launch_k(ctx, lanczos_record_kernel, dim3(1), dim3((n_alpha > 0 || n_beta2 > 0) ? 256 : 1), 0, st, 
        ctx->scal, alpha, beta, i, more ? 1 : 0, ctx->partials + kDotPartialsOffset, n_alpha, ctx->partials, n_beta2);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    if (more && normalise) DSEA_TRY(scale_by_inv_sqrt(ctx, n, Q + (int64_t)m * ldq, ctx->scal + S_BETA2, st));   // :70,75
    return DSEA_OK;
}

static int lanczos_ritz_impl(dsea_ctx* ctx, int64_t n, int64_t ldq, int k, int which, const double* Q,
                             const double* alpha, const double* beta, double* evals, double* evec_min,
                             double* evec_max, int64_t* info_host, cudaStream_t st) {
    double* ymin = ctx->yvec;
    double* ymax = ctx->yvec + kMaxK;
    DSEA_TRY(tridiag_extreme(ctx, k, which, alpha, beta, ctx->scal + S_KEFF, evals, ymin, ymax, st));
    if ((which == DSEA_MIN || which == DSEA_BOTH) && evec_min)
        DSEA_TRY(reorth_update(ctx, n, ldq, k, Q, nullptr, ymin, 1.0, evec_min, nullptr, st));           // :99 (one column)
    if ((which == DSEA_MAX || which == DSEA_BOTH) && evec_max)
        DSEA_TRY(reorth_update(ctx, n, ldq, k, Q, nullptr, ymax, 1.0, evec_max, nullptr, st));
    if (info_host) {
        DSEA_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->scal + S_KEFF, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        DSEA_CUDA(cudaStreamSynchronize(st));
        const int64_t ke = (int64_t)ctx->pinned[0];
        info_host[0] = ke > 0 ? ke : k;
        info_host[1] = (int64_t)ctx->pinned[1];
    }
    return DSEA_OK;
}

static inline int64_t col_stride(int64_t n) { return (n + 15) & ~(int64_t)15; }

// ---- Arnoldi bookkeeping (single thread): H[0..i, i] = c1 + c2, H[i+1, i] = |r| -----------------
__global__ void arnoldi_record_kernel(double* scal, const double* c1, const double* c2, double* hcol, int i) {
    pdl_prologue();
    for (int j = 0; j <= i; ++j) hcol[j] = c1[j] + c2[j];
    const double b2 = scal[S_BETA2];
    const double b = b2 > 0.0 ? sqrt(b2) : 0.0;
    hcol[i + 1] = b;
    if (!(b > 0.0) || !isfinite(b)) scal[S_BETA2] = 0.0;      // invariant subspace: next column is zero
}

}  // namespace dsea

using namespace dsea;

extern "C" {

const char* dsea_last_error(void) { return g_err; }
int dsea_version(void) { return 100; }

int dsea_nccl_unique_id(void* id128_host) { return comm_unique_id(id128_host); }

int dsea_ctx_create(int device, int rank, int world, const void* nccl_id_host, dsea_ctx** out) {
    DSEA_ARG(out != nullptr, "out is NULL");
    DSEA_ARG(world >= 1 && (world & (world - 1)) == 0 && world <= (1 << kMaxRemote), "world must be a power of two");
    DSEA_ARG(rank >= 0 && rank < world, "rank out of range");
    DSEA_CUDA(cudaSetDevice(device));
    dsea_ctx* ctx = new (std::nothrow) dsea_ctx();
    DSEA_ARG(ctx != nullptr, "out of host memory");
    ctx->device = device;
    ctx->rank = rank;
    ctx->world = world;
    while ((1 << ctx->log2world) < world) ++ctx->log2world;
    ctx->pdl = (world == 1) ? 1 : 0;        // programmatic dependent launch: on for one GPU (validated there), opt-in when sharded
    cudaDeviceProp prop;
    DSEA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("libdsea is built for sm_100a (B200) only; device %d is sm_%d%d", device, prop.major, prop.minor);
        delete ctx;
        return DSEA_ERR_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount;
    DSEA_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    DSEA_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    DSEA_CUDA(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    DSEA_CUDA(cudaEventCreateWithFlags(&ctx->ev_poll[0], cudaEventDisableTiming));
    DSEA_CUDA(cudaEventCreateWithFlags(&ctx->ev_poll[1], cudaEventDisableTiming));
    DSEA_CUDA(cudaMalloc(&ctx->scal, S_COUNT * sizeof(double)));
    DSEA_CUDA(cudaMemset(ctx->scal, 0, S_COUNT * sizeof(double)));
    DSEA_CUDA(cudaMalloc(&ctx->partials, kPartialDoubles * sizeof(double)));
    DSEA_CUDA(cudaMalloc(&ctx->cvec, kMaxK * sizeof(double)));
    DSEA_CUDA(cudaMalloc(&ctx->yvec, 2 * kMaxK * sizeof(double)));
    DSEA_CUDA(cudaMalloc(&ctx->tri_work, 10 * kMaxK * sizeof(double)));
    DSEA_CUDA(cudaMallocHost(&ctx->pinned, 64 * sizeof(double)));
    int s = comm_init(ctx, nccl_id_host);
    if (s != DSEA_OK) return s;
    s = mailbox_setup(ctx);
    if (s != DSEA_OK) return s;
    *out = ctx;
    return DSEA_OK;
}

int dsea_ctx_destroy(dsea_ctx* ctx) {
    if (!ctx) return DSEA_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    comm_destroy(ctx);
    cudaFree(ctx->scal);
    cudaFree(ctx->partials);
    cudaFree(ctx->cvec);
    cudaFree(ctx->yvec);
    cudaFree(ctx->tri_work);
    cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev_ready);
    cudaEventDestroy(ctx->ev_comm);
    cudaEventDestroy(ctx->ev_poll[0]);
    cudaEventDestroy(ctx->ev_poll[1]);
    cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return DSEA_OK;
}

int dsea_ctx_p2p(const dsea_ctx* ctx) { return ctx->p2p_ok ? 1 : 0; }
int dsea_ctx_rank(const dsea_ctx* ctx) { return ctx->rank; }
int dsea_ctx_world(const dsea_ctx* ctx) { return ctx->world; }
int64_t dsea_launch_count(const dsea_ctx* ctx) { return ctx->launches; }

int dsea_profile_enable(dsea_ctx* ctx, int on) {
    DSEA_ARG(ctx != nullptr, "NULL ctx");
    if (on && !ctx->prof) ctx->prof = new (std::nothrow) Profiler();
    if (on && ctx->prof) ctx->prof->used = 0;
    if (!on && ctx->prof) {
        for (auto& r : ctx->prof->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        delete ctx->prof;
        ctx->prof = nullptr;
    }
    return DSEA_OK;
}

int dsea_profile_collect(dsea_ctx* ctx, int nkinds, double* ms_host, double* bytes_host, int64_t* count_host) {
    DSEA_ARG(ctx && ms_host && bytes_host && count_host, "NULL argument");
    for (int i = 0; i < nkinds; ++i) { ms_host[i] = 0.0; bytes_host[i] = 0.0; count_host[i] = 0; }
    if (!ctx->prof) return DSEA_OK;
    DSEA_CUDA(cudaDeviceSynchronize());
    for (size_t i = 0; i < ctx->prof->used; ++i) {
        const ProfRec& r = ctx->prof->recs[i];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        if (r.kind < nkinds) { ms_host[r.kind] += ms; bytes_host[r.kind] += r.bytes; count_host[r.kind] += 1; }
    }
    ctx->prof->used = 0;
    return DSEA_OK;
}

int dsea_ctx_set_option(dsea_ctx* ctx, const char* key, int64_t value) {
    DSEA_ARG(ctx && key, "NULL argument");
    if (!strcmp(key, "tfim_tile_bits")) ctx->tfim_tile_bits = (int)value;
    else if (!strcmp(key, "tfim_run_bits")) ctx->tfim_run_bits = (int)value;
    else if (!strcmp(key, "cg_check_every")) ctx->cg_check_every = value < 1 ? 1 : (int)value;
    else if (!strcmp(key, "reorth_ctas_per_sm")) ctx->reorth_ctas_per_sm = value < 1 ? 1 : (value > 16 ? 16 : (int)value);
    else if (!strcmp(key, "basis_fp32")) ctx->basis_fp32 = (value != 0);
    else if (!strcmp(key, "fuse_small")) ctx->fuse_small = (value != 0);
    else if (!strcmp(key, "pdl")) ctx->pdl = (value != 0);
    else if (!strcmp(key, "pdl_staged")) ctx->pdl_staged = (int)value;
    else if (!strcmp(key, "polish_eps_1e15")) ctx->polish_eps = 1e-15 * (double)(value < 1 ? 1 : value);
    else if (!strcmp(key, "p2p")) ctx->p2p_disabled = (value == 0);
    else if (!strcmp(key, "tfim_pipeline")) ctx->tfim_pipeline = (value != 0);
    else if (!strcmp(key, "tfim_tma")) ctx->tfim_tma = (value != 0);
    else if (!strcmp(key, "tfim_pipe_threads")) ctx->tfim_pipe_threads = (value == 256 || value == 1024) ? (int)value : 512;
    else if (!strcmp(key, "tfim_unroll")) ctx->tfim_unroll = (value != 0);
    else if (!strcmp(key, "tfim_stage")) ctx->tfim_stage = (value != 0);
    else if (!strcmp(key, "tfim_generic_min_operands")) ctx->tfim_generic_min_operands = (int)value;
    else if (!strcmp(key, "tfim_direct")) ctx->tfim_direct = (value != 0);
    else if (!strcmp(key, "tfim_fuse_scale")) ctx->tfim_fuse_scale = (value != 0);
    else if (!strcmp(key, "tfim_l2_prefetch")) ctx->tfim_l2_prefetch = (value != 0);
    else if (!strcmp(key, "tfim_pipe_adjoint")) ctx->tfim_pipe_adjoint = (value != 0);
    else if (!strcmp(key, "tfim_pipe_remote")) ctx->tfim_pipe_remote = (value != 0);
    else if (!strcmp(key, "cg_fuse_push")) ctx->cg_fuse_push = (value != 0);
    else if (!strcmp(key, "mailbox")) { if (value == 0) ctx->mail_ok = false; }
    else {
        set_error("unknown option %s", key);
        return DSEA_ERR_ARG;
    }
    return DSEA_OK;
}

// ---- operators ------------------------------------------------------------------------------------
int dsea_op_tfim(dsea_ctx* ctx, int N, dsea_op** out) {
    DSEA_ARG(ctx && out, "NULL argument");
    DSEA_ARG(N >= 2 && N <= 40, "TFIM needs 2 <= N <= 40");
    DSEA_ARG(N - ctx->log2world >= 1, "too few spins for this many ranks");
    dsea_op* op = new (std::nothrow) dsea_op();
    DSEA_ARG(op != nullptr, "out of host memory");
    op->kind = DSEA_OP_TFIM;
    op->ctx = ctx;
    op->N = N;
    op->L = N - ctx->log2world;
    op->n_loc = (int64_t)1 << op->L;
    int s = p2p_setup(ctx, op->n_loc);       // collective: every rank creates its operator at the same point
    if (s != DSEA_OK) {
        delete op;
        return s;
    }
    *out = op;
    return DSEA_OK;
}

int dsea_op_csr(dsea_ctx* ctx, int64_t n, int64_t nnz, const int64_t* rowptr, const int64_t* colidx,
                const double* vals, dsea_op** out) {
    DSEA_ARG(ctx && out && rowptr && (nnz == 0 || (colidx && vals)), "NULL argument");
    DSEA_ARG(ctx->world == 1, "CSR operators are single-GPU");
    dsea_op* op = new (std::nothrow) dsea_op();
    DSEA_ARG(op != nullptr, "out of host memory");
    op->kind = DSEA_OP_CSR;
    op->ctx = ctx;
    op->n_loc = n;
    op->nnz = nnz;
    op->rowptr = rowptr;
    op->colidx = colidx;
    op->vals = vals;
    *out = op;
    return DSEA_OK;
}

int dsea_op_dense(dsea_ctx* ctx, int64_t n, int64_t ld, const double* A, dsea_op** out) {
    DSEA_ARG(ctx && out && A, "NULL argument");
    DSEA_ARG(ctx->world == 1, "dense operators are single-GPU");
    DSEA_ARG(ld >= n, "ld < n");
    dsea_op* op = new (std::nothrow) dsea_op();
    DSEA_ARG(op != nullptr, "out of host memory");
    op->kind = DSEA_OP_DENSE;
    op->ctx = ctx;
    op->n_loc = n;
    op->A = A;
    op->ld = ld;
    *out = op;
    return DSEA_OK;
}

int dsea_op_destroy(dsea_op* op) {
    delete op;
    return DSEA_OK;
}

int64_t dsea_op_local_dim(const dsea_op* op) { return op->n_loc; }
int64_t dsea_op_work_doubles(const dsea_op* op) {
    if (op->kind != DSEA_OP_TFIM) return 0;
    if (op->ctx->p2p_ok && op->ctx->arena_stride >= op->n_loc) return 0;     // partners write into the IPC arena
    return (int64_t)op->ctx->log2world * op->n_loc;                          // NCCL recv buffers
}
int64_t dsea_col_stride(int64_t n_loc) { return col_stride(n_loc); }

int dsea_matvec(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift, const double* v,
                double* u, double* dot_out, double* work, void* stream) {
    DSEA_ARG(ctx && op && v && u, "NULL argument");
    DSEA_ARG(aligned16(v) && aligned16(u) && aligned16(work), "vectors must be 16-byte aligned");
    DSEA_ARG(v != u, "in-place matvec is not supported");
    return apply_op(ctx, op, param, shift, v, u, dot_out, work, (cudaStream_t)stream);
}

int dsea_tfim_dHdg(dsea_ctx* ctx, const dsea_op* op, const double* v, double* u, double* work, void* stream) {
    DSEA_ARG(ctx && op && v && u && op->kind == DSEA_OP_TFIM, "dHdg needs a TFIM operator");
    DSEA_ARG(aligned16(v) && aligned16(u) && aligned16(work) && v != u, "vectors must be distinct and 16-byte aligned");
    return tfim_dHdg(ctx, op, v, u, work, (cudaStream_t)stream);
}

int dsea_adjoint(dsea_ctx* ctx, const dsea_op* op, const double* v1, const double* v2, double* out, double* work,
                 void* stream) {
    DSEA_ARG(ctx && op && v1 && v2 && out, "NULL argument");
    DSEA_ARG(aligned16(v1) && aligned16(v2) && aligned16(work), "vectors must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (op->kind) {
        case DSEA_OP_TFIM:
            return tfim_adjoint(ctx, op, v1, v2, out, work, st);
        case DSEA_OP_CSR:
            return hadamard(ctx, op->n_loc, v1, v2, out, st);
        case DSEA_OP_DENSE:
            return outer(ctx, op->n_loc, 1.0, v1, v2, out, st);
    }
    set_error("unknown operator kind %d", op->kind);
    return DSEA_ERR_ARG;
}

// ---- CG core (used by dsea_cg and by the eigenpair polish) ----------------------------------------------------
static int cg_poll(dsea_ctx* ctx, int slot, cudaStream_t st) {
    DSEA_CUDA(cudaMemcpyAsync(ctx->pinned + 8 * slot, ctx->scal + S_DONE, 3 * sizeof(double),
                              cudaMemcpyDeviceToHost, st));          // {done, iters, rnorm}
    DSEA_CUDA(cudaEventRecord(ctx->ev_poll[slot], st));
    return DSEA_OK;
}

// `proj` (may be NULL): after every operator application the result is projected off `proj` (unit vector), i.e. the
// system solved is P (A - shift) P x = b with P = 1 - proj proj^T — the Jacobi-Davidson correction equation used by
// the eigenpair polish.  b and x0 must be orthogonal to proj.
static int cg_solve_impl(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift, const double* b,
                         double* x, double* work, double eps, int64_t maxit, const double* proj, int64_t* iters_host,
                         cudaStream_t st) {
    const int64_t n = op->n_loc, ld = col_stride(n);
    double* r = work;
    double* d = work + ld;
    double* Ad = work + 2 * ld;
    double* opwork = work + 3 * ld;
    if (maxit <= 0) maxit = n;                                                         // CG.py:32
    DSEA_TRY(cg_setup(ctx, eps, maxit, st));
    DSEA_TRY(apply_op(ctx, op, param, shift, x, Ad, nullptr, opwork, st));             // CG.py:27
    if (proj) DSEA_TRY(project(ctx, n, proj, Ad, Ad, st));
    // sharded TFIM: the kernels that write the search direction d also store it into the partners' arenas, so the
    // matvec of d needs no push pass; a single barrier before its last sweep orders the stores (cg_fuse_push)
    const bool fuse_push = op->kind == DSEA_OP_TFIM && ctx->cg_fuse_push && ctx->p2p_ok && ctx->arena_stride >= n;
    PeerPtrs pp = peer_ptrs(ctx);
    if (fuse_push && !ctx->fresh_collective) DSEA_TRY(comm_barrier(ctx, st));          // partners are done reading x's shards
    DSEA_TRY(cg_init(ctx, n, b, Ad, r, d, st, fuse_push ? &pp : nullptr));
    ctx->guard = ctx->scal + S_DONE;
    const size_t prof_first = prof_mark(ctx);
    int status = DSEA_OK;
    int64_t issued = 0;
    int slot = 0;
    bool have_prev = false;
    bool finished = false;
    while (!finished) {
        const int64_t chunk = ctx->cg_check_every;
        for (int64_t q = 0; q < chunk && status == DSEA_OK; ++q) {
            prof_guard_key(ctx, 2 * (issued + q));
            status = apply_op(ctx, op, param, shift, d, Ad, ctx->scal + S_DAD, opwork, st,
                              fuse_push ? XCH_PREPUSHED_BARRIER : XCH_PUSH, nullptr, nullptr, nullptr, false,
                              /*defer_dot=*/true);                                             // one matvec / iteration
            const int n_dad = ctx->pending_dot_n;                                              // > 0: summed by the update kernel
            ctx->pending_dot_n = 0;
            if (status == DSEA_OK && proj) status = project(ctx, n, proj, Ad, Ad, st);         // d.Ad is unchanged: d is orthogonal to proj
            if (status == DSEA_OK) status = cg_iterate(ctx, n, x, r, d, Ad, st, fuse_push ? &pp : nullptr, n_dad);
        }
        if (status != DSEA_OK) break;
        issued += chunk;
        status = cg_poll(ctx, slot, st);
        if (status != DSEA_OK) break;
        if (have_prev) {      // look at the PREVIOUS chunk's flag while this chunk runs
            cudaError_t e = cudaEventSynchronize(ctx->ev_poll[slot ^ 1]);
            if (e != cudaSuccess) { set_error("CG poll failed: %s", cudaGetErrorString(e)); status = DSEA_ERR_CUDA; break; }
            if (ctx->pinned[8 * (slot ^ 1)] != 0.0) finished = true;
        }
        have_prev = true;
        slot ^= 1;
        if (issued >= maxit) finished = true;
    }
    ctx->guard = nullptr;
    prof_guard_key(ctx, -1);
    if (status != DSEA_OK) return status;
    DSEA_TRY(cg_poll(ctx, slot, st));
    DSEA_CUDA(cudaStreamSynchronize(st));
    const double done = ctx->pinned[8 * slot], iters = ctx->pinned[8 * slot + 1], rnorm = ctx->pinned[8 * slot + 2];
    if (iters_host) *iters_host = (int64_t)iters;
    // iteration `iters - 1` set the flag in its scalar kernel: its direction update (key 2*(iters-1)+1) and
    // everything issued afterwards were no-ops
    prof_retire_guarded(ctx, prof_first, done != 0.0 ? 2 * (int64_t)iters - 1 : INT64_MAX);
    if (done != 1.0) {        // 2: iteration cap or NaN residual; 0: loop left without the flag (cannot happen)
        set_error("CG stopped after %lld iterations with |r| = %.3e >= eps = %.3e (CG.py:32 returns silently here)",
                  (long long)iters, rnorm, eps);
        return DSEA_ERR_NOCONV;
    }
    return DSEA_OK;
}


// ---- Lanczos ----------------------------------------------------------------------------------------
static inline int64_t col_stride_f32(int64_t n) { return (n + 31) & ~(int64_t)31; }      // floats; 128-byte columns

int64_t dsea_lanczos_work_doubles(const dsea_op* op) {
    const int64_t nvec = (op->ctx->basis_fp32 && op->kind == DSEA_OP_TFIM) ? 9 : 1;      // fp32 basis: u, r, q + polish (b, delta, r, d, Ad, spare)
    return nvec * col_stride(op->n_loc) + dsea_op_work_doubles(op);
}

int64_t dsea_lanczos_basis_doubles(const dsea_op* op, int k) {
    const int64_t n = op->n_loc;
    if (op->ctx->basis_fp32 && op->kind == DSEA_OP_TFIM) {
        const int64_t need = ((int64_t)k * col_stride_f32(n) + 1) / 2;                   // k float columns
        return need > col_stride(n) ? need : col_stride(n);                              // ... and room for the fp64 q0 on entry
    }
    return (int64_t)k * col_stride(n);
}

__global__ void polish_scalars_kernel(double* scal) {
    pdl_prologue();
    // S_TMP0 = 1 / |x| from S_BETA2 = x.x
    const double nn = scal[S_BETA2];
    scal[S_TMP0] = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
}

__global__ void rayleigh_residual_kernel(const double* __restrict__ x, const double* __restrict__ Ax,
                                         const double* __restrict__ theta, double* __restrict__ b, int64_t n) {
    pdl_prologue();
    // b = theta x - A x   (right-hand side of the correction equation, orthogonal to x by construction)
    const double th = *theta;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = th * x[i] - Ax[i];
}

// Jacobi-Davidson polish of an approximate ground state x (any norm), in fp64:
//   x <- x / |x| ; theta = x.Ax ; solve P (A - theta) P delta = theta x - A x with P = 1 - x x^T (projected CG) ;
//   x <- (x + delta) / |x + delta| ; theta <- x.Ax.
// One step is quadratically convergent: the fp32-basis Ritz vector (eigenvector error ~1e-6, residual ~3e-7) comes out
// with residual ~1e-11 after ~40 CG iterations (measured on the CPU prototype at N = 12 ... 14).
// work: 7 column strides (Ax, b, delta, CG r / d / Ad) + operator work.
static int lanczos_polish(dsea_ctx* ctx, const dsea_op* op, const double* param, double* x, double* theta_out,
                          double* work, double* opwork, cudaStream_t st) {
    const int64_t n = op->n_loc, ld = col_stride(n);
    double* Ax = work;
    double* b = work + ld;
    double* delta = work + 2 * ld;
    double* cgwork = work + 3 * ld;               // r, d, Ad (+ operator work right behind, see dsea_lanczos_work_doubles)
    (void)opwork;
    const int grid = ctx->num_sms * 8;
    for (int pass = 0; pass < 2; ++pass) {
        DSEA_TRY(dot(ctx, n, x, x, ctx->scal + S_BETA2, st));
        DSEA_TRY(scale_by_inv_sqrt(ctx, n, x, ctx->scal + S_BETA2, st));
        DSEA_TRY(apply_op(ctx, op, param, nullptr, x, Ax, theta_out, cgwork + 3 * ld, st));      // theta = x . A x
        if (pass == 1) break;
        launch_k(ctx, rayleigh_residual_kernel, dim3(grid), dim3(256), 0, st, x, Ax, theta_out, b, n);
        count_launch(ctx);
        DSEA_CUDA(cudaGetLastError());
        DSEA_CUDA(cudaMemsetAsync(delta, 0, (size_t)n * sizeof(double), st));
        int64_t iters = 0;
        const int s = cg_solve_impl(ctx, op, param, theta_out, b, delta, cgwork, ctx->polish_eps, 2000, x, &iters, st);
        if (s != DSEA_OK && s != DSEA_ERR_NOCONV) return s;
        ctx->last_polish_iters = iters;
        DSEA_TRY(axpby(ctx, n, nullptr, delta, nullptr, x, st));                                  // x += delta
    }
    return DSEA_OK;
}

// Lanczos with the fp32 shadow basis (opt-in "basis_fp32", TFIM operators, extreme = "min").
// Every stored Lanczos vector is q~ = fl32(r / beta), and that rounded vector IS the Lanczos vector: the matvec input
// (fp64 copy `qcur`), the recurrence terms (read back from the float columns) and the remote shards (the partners
// round the same r / beta) all use it, so c = Q~^T r0 is exact with respect to the stored basis and only the rounding
// of each new vector (6e-8) perturbs the Lanczos relation.  The Ritz VALUE of T is then only good to ~1e-9, so the
// eigenvalue returned is the Rayleigh quotient of the polished Ritz vector (lanczos_polish).
static int lanczos_fp32_impl(dsea_ctx* ctx, const dsea_op* op, const double* param, int k, double* Qraw, double* work,
                             double* alpha, double* beta, double* evals, double* evec_min, int64_t* info_host,
                             cudaStream_t st) {
    const int64_t n = op->n_loc, ld = col_stride(n), ldq = col_stride_f32(n);
    float* Q = (float*)Qraw;
    double* u = work;
    double* rvec = work + ld;
    double* qcur = work + 2 * ld;
    double* polish_work = work + 3 * ld;          // 6 strides
    double* opwork = work + 9 * ld;
    // q0 arrives as n doubles at the start of the basis buffer, which column 0 (n floats) overlaps: move it out first
    DSEA_CUDA(cudaMemcpyAsync(rvec, Qraw, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    launch_k(ctx, lanczos_reset_kernel, dim3(1), dim3(1), 0, st, ctx->scal);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    DSEA_TRY(dot(ctx, n, rvec, rvec, ctx->scal + S_BETA2, st));
    launch_k(ctx, polish_scalars_kernel, dim3(1), dim3(1), 0, st, ctx->scal);
    count_launch(ctx);
    DSEA_TRY(scale_round_store(ctx, n, rvec, ctx->scal + S_TMP0, qcur, Q, st));               // Lanczos.py:53
    const bool fuse_push = ctx->p2p_ok && ctx->arena_stride >= n;
    PeerPtrs pp = peer_ptrs(ctx);
    for (int i = 0; i < k; ++i) {
        const bool pre = fuse_push && i > 0;
        DSEA_TRY(apply_op(ctx, op, param, nullptr, qcur, u, ctx->scal + S_ALPHA_L, opwork, st,
                          pre ? XCH_PREPUSHED : XCH_PUSH, ctx->scal + S_INVBETA, nullptr, nullptr, /*round_remote=*/true,
                          /*defer_dot=*/true));
        const int n_alpha = ctx->pending_dot_n;
        ctx->pending_dot_n = 0;
        const int m = i + 1;
        const bool more = (i < k - 1);
        int n_beta2 = 0;
        if (more) {
            Recurrence rec;
            rec.qi = Q + (int64_t)i * ldq;
            rec.qim1 = i > 0 ? Q + (int64_t)(i - 1) * ldq : nullptr;
            rec.alpha = ctx->scal + S_ALPHA_L;
            rec.beta = ctx->scal + S_BETAPREV;
            rec.r0_out = rvec;
            rec.alpha_partials = ctx->partials + kDotPartialsOffset;
            rec.n_alpha = n_alpha;
            DSEA_TRY(reorth_dots_f32(ctx, n, ldq, m, Q, u, ctx->cvec, st, &rec));
            DSEA_TRY(reorth_update_f32(ctx, n, ldq, m, Q, rvec, ctx->cvec, -1.0, rvec, ctx->scal + S_BETA2, st,
                                       fuse_push ? &pp : nullptr, /*defer_norm=*/true));
            n_beta2 = ctx->pending_norm_n;
            ctx->pending_norm_n = 0;
        }
        launch_k(ctx, lanczos_record_kernel, dim3(1), dim3((n_alpha > 0 || n_beta2 > 0) ? 256 : 1), 0, st, 
            ctx->scal, alpha, beta, i, more ? 1 : 0, ctx->partials + kDotPartialsOffset, n_alpha, ctx->partials, n_beta2);
        count_launch(ctx);
        DSEA_CUDA(cudaGetLastError());
        if (more) DSEA_TRY(scale_round_store(ctx, n, rvec, ctx->scal + S_INVBETA, qcur, Q + (int64_t)m * ldq, st));
    }
    double* ymin = ctx->yvec;
    DSEA_TRY(tridiag_extreme(ctx, k, DSEA_MIN, alpha, beta, ctx->scal + S_KEFF, evals, ymin, ctx->yvec + kMaxK, st));
    DSEA_TRY(reorth_update_f32(ctx, n, ldq, k, Q, nullptr, ymin, 1.0, evec_min, nullptr, st));
    DSEA_TRY(lanczos_polish(ctx, op, param, evec_min, evals, polish_work, opwork, st));
    if (info_host) {
        DSEA_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->scal + S_KEFF, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        DSEA_CUDA(cudaStreamSynchronize(st));
        const int64_t ke = (int64_t)ctx->pinned[0];
        info_host[0] = ke > 0 ? ke : k;
        info_host[1] = (int64_t)ctx->pinned[1];
    }
    return DSEA_OK;
}

int dsea_lanczos(dsea_ctx* ctx, const dsea_op* op, const double* param, int k, int which, double* Q, double* work,
                 double* alpha, double* beta, double* evals, double* evec_min, double* evec_max,
                 int64_t* info_host, void* stream) {
    DSEA_ARG(ctx && op && Q && work && alpha && beta && evals, "NULL argument");
    DSEA_ARG(k >= 1 && k <= kMaxK, "k out of range");
    DSEA_ARG(aligned16(Q) && aligned16(work) && aligned16(evec_min) && aligned16(evec_max), "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = op->n_loc, ldq = col_stride(n);
    if (ctx->basis_fp32 && op->kind == DSEA_OP_TFIM) {
        DSEA_ARG(which == DSEA_MIN && evec_min != nullptr, "the fp32 shadow basis supports extreme = 'min' only");
        return lanczos_fp32_impl(ctx, op, param, k, Q, work, alpha, beta, evals, evec_min, info_host, st);
    }
    double* u = work;
    double* opwork = work + ldq;
    DSEA_TRY(lanczos_start_impl(ctx, n, Q, st));
    const bool fuse_push = op->kind == DSEA_OP_TFIM && ctx->p2p_ok && ctx->arena_stride >= n;
    // K3 fused into K1: pass 2 leaves the un-normalised r in column i+1; the first sweep of the next matvec reads it,
    // multiplies by 1 / beta (device scalar), writes q back in place and produces u from it (Lanczos.py:70-71).
    const bool fuse_scale = op->kind == DSEA_OP_TFIM && tfim_can_fuse_scale(ctx, op);
    for (int i = 0; i < k; ++i) {
        const bool pre = fuse_push && i > 0;
        double* qi = Q + (int64_t)i * ldq;
        const bool fs = fuse_scale && i > 0;
        DSEA_TRY(apply_op(ctx, op, param, nullptr, qi, u, ctx->scal + S_ALPHA_L, opwork, st,
                          pre ? XCH_PREPUSHED : XCH_PUSH, ctx->scal + S_INVBETA, fs ? ctx->scal + S_INVBETA : nullptr,
                          fs ? qi : nullptr, false, /*defer_dot=*/true));          // Lanczos.py:54-55,71-72 (alpha in the epilogue)
        DSEA_TRY(lanczos_step_impl(ctx, n, ldq, k, i, Q, u, alpha, beta, st, true, fuse_push, !fuse_scale));
    }
    return lanczos_ritz_impl(ctx, n, ldq, k, which, Q, alpha, beta, evals, evec_min, evec_max, info_host, st);
}

int dsea_lanczos_start(dsea_ctx* ctx, int64_t n_loc, double* Q, void* stream) {
    DSEA_ARG(ctx && Q && aligned16(Q), "bad Q");
    return lanczos_start_impl(ctx, n_loc, Q, (cudaStream_t)stream);
}

int dsea_lanczos_step(dsea_ctx* ctx, int64_t n_loc, int k, int i, double* Q, const double* u, double* alpha,
                      double* beta, void* stream) {
    DSEA_ARG(ctx && Q && u && alpha && beta, "NULL argument");
    DSEA_ARG(i >= 0 && i < k && k <= kMaxK, "step index out of range");
    DSEA_ARG(aligned16(Q) && aligned16(u), "buffers must be 16-byte aligned");
    return lanczos_step_impl(ctx, n_loc, col_stride(n_loc), k, i, Q, u, alpha, beta, (cudaStream_t)stream, false);
}

int dsea_lanczos_ritz(dsea_ctx* ctx, int64_t n_loc, int k, int which, const double* Q, const double* alpha,
                      const double* beta, double* evals, double* evec_min, double* evec_max, int64_t* info_host,
                      void* stream) {
    DSEA_ARG(ctx && Q && alpha && beta && evals, "NULL argument");
    DSEA_ARG(k >= 1 && k <= kMaxK, "k out of range");
    DSEA_ARG(aligned16(Q) && aligned16(evec_min) && aligned16(evec_max), "buffers must be 16-byte aligned");
    return lanczos_ritz_impl(ctx, n_loc, col_stride(n_loc), k, which, Q, alpha, beta, evals, evec_min, evec_max,
                             info_host, (cudaStream_t)stream);
}

// ---- Arnoldi (non-symmetric family, eig.py) ---------------------------------------------------------
int dsea_arnoldi_start(dsea_ctx* ctx, int64_t n_loc, double* Q, double* norm_out, void* stream) {
    DSEA_ARG(ctx && Q && aligned16(Q), "bad Q");
    cudaStream_t st = (cudaStream_t)stream;
    DSEA_TRY(dot(ctx, n_loc, Q, Q, ctx->scal + S_BETA2, st));
    if (norm_out) DSEA_CUDA(cudaMemcpyAsync(norm_out, ctx->scal + S_BETA2, sizeof(double), cudaMemcpyDeviceToDevice, st));
    return scale_by_inv_sqrt(ctx, n_loc, Q, ctx->scal + S_BETA2, st);
}

int dsea_arnoldi_step(dsea_ctx* ctx, int64_t n_loc, int m, int i, double* Q, const double* u, double* H,
                      void* stream) {
    DSEA_ARG(ctx && Q && u && H, "NULL argument");
    DSEA_ARG(i >= 0 && i < m && m < kMaxK, "Arnoldi step index out of range");
    DSEA_ARG(aligned16(Q) && aligned16(u), "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldq = col_stride(n_loc);
    const int cols = i + 1;
    double* qnext = Q + (int64_t)cols * ldq;
    double* c1 = ctx->cvec;
    double* c2 = ctx->yvec;
    // classical Gram-Schmidt with one re-orthogonalisation sweep (CGS2, as ARPACK's DGKS refinement)
    DSEA_TRY(reorth_dots(ctx, n_loc, ldq, cols, Q, u, c1, st));
    DSEA_TRY(reorth_update(ctx, n_loc, ldq, cols, Q, u, c1, -1.0, qnext, nullptr, st));
    DSEA_TRY(reorth_dots(ctx, n_loc, ldq, cols, Q, qnext, c2, st));
    DSEA_TRY(reorth_update(ctx, n_loc, ldq, cols, Q, qnext, c2, -1.0, qnext, ctx->scal + S_BETA2, st));
    launch_k(ctx, arnoldi_record_kernel, dim3(1), dim3(1), 0, st, ctx->scal, c1, c2, H + (int64_t)i * (m + 1), i);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return scale_by_inv_sqrt(ctx, n_loc, qnext, ctx->scal + S_BETA2, st);
}

int dsea_combine(dsea_ctx* ctx, int64_t n_loc, int m, const double* Q, const double* coef, const double* add,
                 double* out, void* stream) {
    DSEA_ARG(ctx && Q && coef && out, "NULL argument");
    DSEA_ARG(m >= 1 && m <= kMaxK, "m out of range");
    DSEA_ARG(aligned16(Q) && aligned16(out) && aligned16(add), "buffers must be 16-byte aligned");
    return reorth_update(ctx, n_loc, col_stride(n_loc), m, Q, add, coef, 1.0, out, nullptr, (cudaStream_t)stream);
}

// ---- CG -----------------------------------------------------------------------------------------------
int64_t dsea_cg_work_doubles(const dsea_op* op) { return 3 * col_stride(op->n_loc) + dsea_op_work_doubles(op); }

int dsea_cg(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift, const double* b, double* x,
            double* work, double eps, int64_t maxit, int64_t* iters_host, void* stream) {
    DSEA_ARG(ctx && op && b && x && work, "NULL argument");
    DSEA_ARG(aligned16(b) && aligned16(x) && aligned16(work), "buffers must be 16-byte aligned");
    return cg_solve_impl(ctx, op, param, shift, b, x, work, eps, maxit, nullptr, iters_host, (cudaStream_t)stream);
}

int dsea_cg_init(dsea_ctx* ctx, int64_t n_loc, const double* b, const double* Ax, double* r, double* d,
                 double* state_host, void* stream) {
    DSEA_ARG(ctx && b && Ax && r && d, "NULL argument");
    DSEA_ARG(aligned16(b) && aligned16(Ax) && aligned16(r) && aligned16(d), "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DSEA_TRY(cg_setup(ctx, state_host ? state_host[0] : 1e-7, state_host ? (int64_t)state_host[1] : n_loc, st));
    DSEA_TRY(cg_init(ctx, n_loc, b, Ax, r, d, st));
    if (state_host) {
        DSEA_TRY(cg_poll(ctx, 0, st));
        DSEA_CUDA(cudaStreamSynchronize(st));
        state_host[0] = ctx->pinned[2];
        state_host[1] = ctx->pinned[1];
        state_host[2] = ctx->pinned[0];
    }
    return DSEA_OK;
}

int dsea_cg_update(dsea_ctx* ctx, int64_t n_loc, double* x, double* r, double* d, const double* Ad,
                   double* state_host, void* stream) {
    DSEA_ARG(ctx && x && r && d && Ad, "NULL argument");
    DSEA_ARG(aligned16(x) && aligned16(r) && aligned16(d) && aligned16(Ad), "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DSEA_TRY(dot(ctx, n_loc, d, Ad, ctx->scal + S_DAD, st));
    DSEA_TRY(cg_iterate(ctx, n_loc, x, r, d, Ad, st));
    if (state_host) {
        DSEA_TRY(cg_poll(ctx, 0, st));
        DSEA_CUDA(cudaStreamSynchronize(st));
        state_host[0] = ctx->pinned[2];
        state_host[1] = ctx->pinned[1];
        state_host[2] = ctx->pinned[0];
    }
    return DSEA_OK;
}

// ---- level 1 ----------------------------------------------------------------------------------------------
int dsea_dot(dsea_ctx* ctx, int64_t n_loc, const double* a, const double* b, double* out, void* stream) {
    DSEA_ARG(ctx && a && b && out && aligned16(a) && aligned16(b), "bad argument");
    return dot(ctx, n_loc, a, b, out, (cudaStream_t)stream);
}

int dsea_project(dsea_ctx* ctx, int64_t n_loc, const double* psi, const double* b, double* out, void* stream) {
    DSEA_ARG(ctx && psi && b && out && aligned16(psi) && aligned16(b) && aligned16(out), "bad argument");
    return project(ctx, n_loc, psi, b, out, (cudaStream_t)stream);
}

int dsea_axpby(dsea_ctx* ctx, int64_t n_loc, const double* a, const double* x, const double* b, double* y,
               void* stream) {
    DSEA_ARG(ctx && x && y && aligned16(x) && aligned16(y), "bad argument");
    return axpby(ctx, n_loc, a, x, b, y, (cudaStream_t)stream);
}

int dsea_outer(dsea_ctx* ctx, int64_t n, double scale, const double* a, const double* b, double* out, void* stream) {
    DSEA_ARG(ctx && a && b && out, "bad argument");
    return outer(ctx, n, scale, a, b, out, (cudaStream_t)stream);
}

int dsea_randn(dsea_ctx* ctx, int64_t n_loc, uint64_t seed, uint64_t stream_id, double* out, void* stream) {
    DSEA_ARG(ctx && out, "bad argument");
    return randn(ctx, n_loc, seed, stream_id, (uint64_t)ctx->rank * (uint64_t)n_loc, out, (cudaStream_t)stream);
}

}  // extern "C"
