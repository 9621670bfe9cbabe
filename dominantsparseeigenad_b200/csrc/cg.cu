// cg.cu — conjugate gradients for (A - E0) x = b on the subspace orthogonal to the ground state (K7).
// Replaces CG.CG_torch (CG.py:3-41) with
//   * ONE operator application per iteration (the reference applies it twice, CG.py:34 and :40);
//   * fused vector updates: {x += a d, r -= a Ad, |r|^2} in one pass (48 n bytes) and d = r + b d (24 n bytes);
//   * d.Ad folded into the matvec epilogue;
//   * all scalars (alpha, beta, r.r, iteration count, convergence flag) resident in device memory: the
//     reference's two `.item()` host syncs per iteration (CG.py:28,35) become one poll of a pinned flag
//     every `cg_check_every` iterations; once the flag is set every later kernel returns immediately.
// Stopping rule is the reference's: |r|_2 < eps absolute (eps = 1e-7, CG.py:25), at most maxit steps.
#include "common.cuh"

namespace dsea {

constexpr int kCgThreads = 256;

static inline int cg_grid(const dsea_ctx* ctx, int64_t n) {
    int64_t want = (n + (int64_t)kCgThreads * 8 - 1) / ((int64_t)kCgThreads * 8);
    int64_t cap = (int64_t)ctx->num_sms * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

__global__ void cg_setup_kernel(double* scal, double eps, double maxit) {
    pdl_prologue();
    scal[S_DONE] = 0.0;
    scal[S_ITERS] = 0.0;
    scal[S_EPS] = eps;
    scal[S_MAXIT] = maxit;
    scal[S_RR] = 0.0;
    scal[S_RR_NEW] = 0.0;
    scal[S_RNORM] = 0.0;
}

// r = b - Ax ; d = r ; partial r.r
__global__ void __launch_bounds__(kCgThreads)
cg_init_kernel(const double* __restrict__ b, const double* __restrict__ Ax, double* __restrict__ r,
               double* __restrict__ d, int64_t n, double* __restrict__ partials, const PeerPtrs peers) {
    pdl_prologue();
    __shared__ double red[32];
    double s = 0.0;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 bv = ldg2(b + 2 * i), av = ldg2(Ax + 2 * i);
        const double2 rv = make_double2(bv.x - av.x, bv.y - av.y);
        stg2(r + 2 * i, rv);
        stg2(d + 2 * i, rv);
        for (int pj = 0; pj < peers.n; ++pj) stg2(peers.p[pj] + 2 * i, rv);      // fused exchange of d (NVLink stores)
        s += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double rv = b[n - 1] - Ax[n - 1];
        r[n - 1] = rv;
        d[n - 1] = rv;
        for (int pj = 0; pj < peers.n; ++pj) peers.p[pj][n - 1] = rv;
        s += rv * rv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// after the initial residual: |r0| < eps  =>  done (CG.py:28-29)
__global__ void cg_first_check_kernel(double* scal) {
    pdl_prologue();
    const double rn = sqrt(scal[S_RR]);
    scal[S_RNORM] = rn;
    if (rn < scal[S_EPS]) scal[S_DONE] = 1.0;
}

// x += alpha d ; r -= alpha Ad ; partial |r|^2      with alpha = rr / dAd      (CG.py:31,33,34,40)
__global__ void __launch_bounds__(kCgThreads)
cg_update_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d,
                    const double* __restrict__ Ad, int64_t n, const double* __restrict__ scal,
                    double* __restrict__ partials, const double* __restrict__ dad_partials, int n_dad) {
    pdl_prologue();
    __shared__ double red[32];
    if (scal[S_DONE] != 0.0) return;
    const double dad = n_dad > 0 ? sum_partials_seq(dad_partials, n_dad) : scal[S_DAD];   // deferred matvec epilogue
    const double alpha = scal[S_RR] / dad;
    double s = 0.0;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 dv = ldg2(d + 2 * i), av = ldg2(Ad + 2 * i);
        double2 xv = ldg2(x + 2 * i), rv = ldg2(r + 2 * i);
        xv.x += alpha * dv.x;
        xv.y += alpha * dv.y;
        rv.x -= alpha * av.x;
        rv.y -= alpha * av.y;
        stg2(x + 2 * i, xv);
        stg2(r + 2 * i, rv);
        s += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        x[n - 1] += alpha * d[n - 1];
        const double rv = r[n - 1] - alpha * Ad[n - 1];
        r[n - 1] = rv;
        s += rv * rv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// scalar bookkeeping of one iteration                                              (CG.py:35-38)
// n_rr > 0: one CTA first sums the |r|^2 partials of the update kernel itself (fixed order), replacing the separate
// finalize launch; otherwise S_RR_NEW was reduced (and, when sharded, all-reduced) before.
__global__ void __launch_bounds__(256) cg_scalar_kernel(double* scal, const double* __restrict__ rr_partials, int n_rr) {
    pdl_prologue();
    __shared__ double red[32];
    if (scal[S_DONE] != 0.0) return;
    if (n_rr > 0) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n_rr; i += blockDim.x) s += rr_partials[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) scal[S_RR_NEW] = s;
    }
    if (threadIdx.x != 0) return;
    const double rr_new = scal[S_RR_NEW];
    const double it = scal[S_ITERS] + 1.0;
    scal[S_ITERS] = it;
    const double rn = sqrt(rr_new);
    scal[S_RNORM] = rn;
    if (rn < scal[S_EPS]) {                  // converged (CG.py:35)
        scal[S_DONE] = 1.0;
        return;
    }
    if (it >= scal[S_MAXIT] || !(rn == rn)) {   // iteration cap (CG.py:32) or NaN: reported as DSEA_ERR_NOCONV
        scal[S_DONE] = 2.0;
        return;
    }
    scal[S_BETA] = rr_new / scal[S_RR];
    scal[S_RR] = rr_new;
}

// d = r + beta d                                                                   (CG.py:39)
__global__ void __launch_bounds__(kCgThreads)
cg_update_d_kernel(double* __restrict__ d, const double* __restrict__ r, int64_t n, const double* __restrict__ scal,
                   const PeerPtrs peers) {
    pdl_prologue();
    if (scal[S_DONE] != 0.0) return;
    const double beta = scal[S_BETA];
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 rv = ldg2(r + 2 * i);
        double2 dv = ldg2(d + 2 * i);
        dv.x = rv.x + beta * dv.x;
        dv.y = rv.y + beta * dv.y;
        stg2(d + 2 * i, dv);
        for (int pj = 0; pj < peers.n; ++pj) stg2(peers.p[pj] + 2 * i, dv);      // fused exchange: the next matvec's remote shards
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double dv = r[n - 1] + beta * d[n - 1];
        d[n - 1] = dv;
        for (int pj = 0; pj < peers.n; ++pj) peers.p[pj][n - 1] = dv;
    }
}

int cg_setup(dsea_ctx* ctx, double eps, int64_t maxit, cudaStream_t st) {
    launch_k(ctx, cg_setup_kernel, dim3(1), dim3(1), 0, st, ctx->scal, eps, (double)maxit);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

int cg_init(dsea_ctx* ctx, int64_t n, const double* b, const double* Ax, double* r, double* d, cudaStream_t st,
            const PeerPtrs* peers) {
    const int grid = cg_grid(ctx, n);
    PeerPtrs pp;
    pp.n = 0;
    if (peers) pp = *peers;
    launch_k(ctx, cg_init_kernel, dim3(grid), dim3(kCgThreads), 0, st, b, Ax, r, d, n, ctx->partials, pp);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    DSEA_TRY(finalize_reduce(ctx, grid, 1, ctx->scal + S_RR, st));
    launch_k(ctx, cg_first_check_kernel, dim3(1), dim3(1), 0, st, ctx->scal);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// one iteration AFTER Ad and d.Ad (scal[S_DAD]) are available
int cg_iterate(dsea_ctx* ctx, int64_t n, double* x, double* r, double* d, const double* Ad, cudaStream_t st,
               const PeerPtrs* peers, int n_dad) {
    const int grid = cg_grid(ctx, n);
    PeerPtrs pp;
    pp.n = 0;
    if (peers) pp = *peers;
    int tok = prof_begin(ctx, PK_CG_UPDATE, 48.0 * (double)n, st);
    const bool fused = ctx->world == 1 && ctx->fuse_small;
    launch_k(ctx, cg_update_xr_kernel, dim3(grid), dim3(kCgThreads), 0, st, x, r, d, Ad, n, ctx->scal, ctx->partials,
                                                     ctx->partials + kDotPartialsOffset, n_dad);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    if (fused) {
        launch_k(ctx, cg_scalar_kernel, dim3(1), dim3(256), 0, st, ctx->scal, ctx->partials, grid);
    } else {
        DSEA_TRY(finalize_reduce(ctx, grid, 1, ctx->scal + S_RR_NEW, st));
        launch_k(ctx, cg_scalar_kernel, dim3(1), dim3(1), 0, st, ctx->scal, nullptr, 0);
    }
    prof_guard_next_phase(ctx);
    tok = prof_begin(ctx, PK_CG_UPDATE, (24.0 + 8.0 * pp.n) * (double)n, st);
    launch_k(ctx, cg_update_d_kernel, dim3(grid), dim3(kCgThreads), 0, st, d, r, n, ctx->scal, pp);
    prof_end(ctx, tok, st);
    count_launch(ctx, 2);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

}  // namespace dsea
