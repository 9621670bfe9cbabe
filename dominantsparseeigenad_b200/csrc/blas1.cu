// blas1.cu — streaming level-1 kernels: deterministic two-stage reductions, projection (K8),
// fused CG vector updates (K7), axpby, Philox normal generator.  All fp64, 16-byte vector accesses,
// grid sized as a small multiple of the SM count with grid-stride loops.
#include "common.cuh"

namespace dsea {

constexpr int kThreads = 256;

static inline int stream_grid(const dsea_ctx* ctx, int64_t n, int per_thread = 8) {
    int64_t want = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    int64_t cap = (int64_t)ctx->num_sms * 8;
    if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ---- second stage of every reduction: out[col] = sum_b partials[b*ncols + col], fixed order ----
__global__ void __launch_bounds__(128) finalize_kernel(const double* __restrict__ partials, int nblocks, int ncols,
                                                       double* __restrict__ out) {
    pdl_prologue();
    __shared__ double red[32];
    const int col = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[(size_t)b * ncols + col];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[col] = s;
}

int finalize_partials(dsea_ctx* ctx, int nblocks, int ncols, double* out, cudaStream_t st, const double* src) {
    launch_k(ctx, finalize_kernel, dim3(ncols), dim3(128), 0, st, src ? src : ctx->partials, nblocks, ncols, out);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// ---- dot ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       int64_t n, double* __restrict__ partials) {
    pdl_prologue();
    __shared__ double red[32];
    double s = 0.0;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 x = ldg2(a + 2 * i), y = ldg2(b + 2 * i);
        s += x.x * y.x + x.y * y.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s += a[n - 1] * b[n - 1];
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

int dot(dsea_ctx* ctx, int64_t n, const double* a, const double* b, double* out, cudaStream_t st) {
    const int grid = stream_grid(ctx, n);
    launch_k(ctx, dot_kernel, dim3(grid), dim3(kThreads), 0, st, a, b, n, ctx->partials);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return finalize_reduce(ctx, grid, 1, out, st);
}

// ---- y = a x + b y -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) axpby_kernel(const double* __restrict__ pa, const double* __restrict__ x,
                                                         const double* __restrict__ pb, double* __restrict__ y,
                                                         int64_t n) {
    pdl_prologue();
    const double a = pa ? *pa : 1.0, b = pb ? *pb : 1.0;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 xv = ldg2(x + 2 * i);
        double2 yv = ldg2(y + 2 * i);
        yv.x = a * xv.x + b * yv.x;
        yv.y = a * xv.y + b * yv.y;
        stg2(y + 2 * i, yv);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = a * x[n - 1] + b * y[n - 1];
}

int axpby(dsea_ctx* ctx, int64_t n, const double* a, const double* x, const double* b, double* y, cudaStream_t st) {
    launch_k(ctx, axpby_kernel, dim3(stream_grid(ctx, n)), dim3(kThreads), 0, st, a, x, b, y, n);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// ---- projection out = b - (psi.b) psi  (K8): dot pass, then fused axpy pass ----------------------
__global__ void __launch_bounds__(kThreads) project_apply_kernel(const double* __restrict__ psi,
                                                                 const double* __restrict__ b,
                                                                 const double* __restrict__ pdot,
                                                                 double* __restrict__ out, int64_t n) {
    pdl_prologue();
    const double d = *pdot;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 p = ldg2(psi + 2 * i), v = ldg2(b + 2 * i);
        stg2(out + 2 * i, make_double2(v.x - d * p.x, v.y - d * p.y));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = b[n - 1] - d * psi[n - 1];
}

int project(dsea_ctx* ctx, int64_t n, const double* psi, const double* b, double* out, cudaStream_t st) {
    double* d = ctx->scal + S_TMP0;
    DSEA_TRY(dot(ctx, n, psi, b, d, st));
    launch_k(ctx, project_apply_kernel, dim3(stream_grid(ctx, n)), dim3(kThreads), 0, st, psi, b, d, out, n);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// ---- x *= 1/sqrt(norm2)  (K3: normalise a new Lanczos vector in place) ----------------------------
__global__ void __launch_bounds__(kThreads) scale_inv_sqrt_kernel(double* __restrict__ x,
                                                                  const double* __restrict__ norm2, int64_t n) {
    pdl_prologue();
    const double nn = *norm2;
    const double s = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;     // breakdown (|r| == 0) leaves a zero column
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v = ldg2(x + 2 * i);
        v.x *= s;
        v.y *= s;
        stg2(x + 2 * i, v);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] *= s;
}

int scale_by_inv_sqrt(dsea_ctx* ctx, int64_t n, double* x, const double* norm2, cudaStream_t st) {
    const int tok = prof_begin(ctx, PK_NORMALISE, 16.0 * (double)n, st);
    launch_k(ctx, scale_inv_sqrt_kernel, dim3(stream_grid(ctx, n)), dim3(kThreads), 0, st, x, norm2, n);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// ---- publish a shard to the partners' arenas (NVLink stores), used when no producing kernel can ----
__global__ void __launch_bounds__(kThreads) push_kernel(const double* __restrict__ v, int64_t n, PeerPtrs peers) {
    pdl_prologue();
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 x = ldg2(v + 2 * i);
        for (int j = 0; j < peers.n; ++j) stg2(peers.p[j] + 2 * i, x);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        for (int j = 0; j < peers.n; ++j) peers.p[j][n - 1] = v[n - 1];
}

int push_to_peers(dsea_ctx* ctx, const double* v, int64_t n, cudaStream_t st) {
    launch_k(ctx, push_kernel, dim3(stream_grid(ctx, n)), dim3(kThreads), 0, st, v, n, peer_ptrs(ctx));
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// ---- Philox4x32-10 standard normals ----------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
}

__global__ void __launch_bounds__(kThreads) randn_kernel(double* __restrict__ out, int64_t n, uint64_t seed,
                                                         uint64_t sid, uint64_t offset) {
    pdl_prologue();
    const int64_t n2 = (n + 1) >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const uint64_t ctr = (offset >> 1) + (uint64_t)i;      // one Philox block per PAIR of outputs
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)sid, (uint32_t)(sid >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        // two uniforms in (0,1] / [0,1) with 53 and 32+ bits, Box-Muller
        const uint64_t a = ((uint64_t)c[0] << 32) | c[1];
        const uint64_t b = ((uint64_t)c[2] << 32) | c[3];
        const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
        const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0, 1)
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        out[2 * i] = rad * cs;
        if (2 * i + 1 < n) out[2 * i + 1] = rad * sn;
    }
}

int randn(dsea_ctx* ctx, int64_t n, uint64_t seed, uint64_t sid, uint64_t offset, double* out, cudaStream_t st) {
    launch_k(ctx, randn_kernel, dim3(stream_grid(ctx, n, 4)), dim3(kThreads), 0, st, out, n, seed, sid, offset);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

}  // namespace dsea
