// reorth.cu — full re-orthogonalisation as a fused two-pass tall-skinny GEMV (K2a/K2b) and the
// Ritz-vector GEMV (K5).  Replaces `r -= Q[:, :i] (Q[:, :i]^T r)` (Lanczos.py:66) and
// `Qk @ eigvectors` (Lanczos.py:99).
//
// Layout: the basis is COLUMN-contiguous (column j at Q + j*ldq, ldq a multiple of 16 doubles), the
// opposite of the reference's row-major (n, k) tensor whose column views have stride k.
//
//   pass 1 (reorth_dots)   c = Q[:, :m]^T u        reads 8 n (m + 1) bytes
//   pass 2 (reorth_update) r = u - Q[:, :m] c      reads 8 n (m + 1), writes 8 n; |r|^2 in the epilogue
//
// Both are HBM-streaming kernels: a CTA owns 1024-row tiles (4 rows per thread as two 16 B vectors),
// keeps the u / r tile in registers and streams the m column segments past it with >= 16 independent
// 16 B loads in flight per thread.  Pass 2 walks the tiles in the opposite order to pass 1 so that it
// starts on the part of Q that pass 1 left in the 126 MB L2.
// Reductions are deterministic: per-warp shared-memory accumulators -> per-CTA partials (fixed order)
// -> one finalize CTA per column (fixed order) -> NCCL allreduce when sharded.
#include "common.cuh"

namespace dsea {

constexpr int kRThreads = 256;
constexpr int kRowsPerThread = 4;
constexpr int kTileRows = kRThreads * kRowsPerThread;   // 1024
constexpr int kJB = 8;                                   // columns per register chunk in pass 1

// Sum kJB=8 per-lane values across the warp with 9 shuffles; afterwards every lane holds the warp
// total of value index ((lane >> 2) & 7).
__device__ __forceinline__ double warp_reduce8(double (&v)[8], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double send = up ? v[i] : v[i + 4];
            const double keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double send = up ? v[i] : v[i + 2];
            const double keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// ---- pass 1: partials[cta][j] = sum over the CTA's rows of Q[row, j] * u[row] -------------------
__global__ void __launch_bounds__(kRThreads, 2)
reorth_dots_kernel(const double* __restrict__ Q, int64_t ldq, const double* __restrict__ u, int64_t n, int m,
                   double* __restrict__ partials, const Recurrence rec) {
    extern __shared__ double wacc[];                 // [8 warps][m] per-warp accumulators
    // three-term recurrence folded into the prologue: r0 = u - alpha q_i - beta q_{i-1}  (Lanczos.py:61)
    const double ra = rec.qi ? *rec.alpha : 0.0;
    const double rb = rec.qim1 ? *rec.beta : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < 8 * m; j += kRThreads) wacc[j] = 0.0;
    __syncthreads();
    double* my = wacc + warp * m;
    const int64_t ntiles = (n + kTileRows - 1) / kTileRows;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * kTileRows + 2 * threadIdx.x;          // rows r0, r0+1
        const int64_t r1 = r0 + kTileRows / 2;                        // rows r1, r1+1
        const bool full = (t + 1) * kTileRows <= n;
        double2 u0, u1;
        if (full) {
            u0 = ldg2(u + r0);
            u1 = ldg2(u + r1);
        } else {
            u0.x = r0 < n ? u[r0] : 0.0;
            u0.y = r0 + 1 < n ? u[r0 + 1] : 0.0;
            u1.x = r1 < n ? u[r1] : 0.0;
            u1.y = r1 + 1 < n ? u[r1 + 1] : 0.0;
        }
        if (rec.qi) {
            if (full) {
                const double2 q0 = ldg2(rec.qi + r0), q1 = ldg2(rec.qi + r1);
                u0.x -= ra * q0.x; u0.y -= ra * q0.y; u1.x -= ra * q1.x; u1.y -= ra * q1.y;
                if (rec.qim1) {
                    const double2 p0 = ldg2(rec.qim1 + r0), p1 = ldg2(rec.qim1 + r1);
                    u0.x -= rb * p0.x; u0.y -= rb * p0.y; u1.x -= rb * p1.x; u1.y -= rb * p1.y;
                }
                stg2(rec.r0_out + r0, u0);
                stg2(rec.r0_out + r1, u1);
            } else {
                const int64_t rows[4] = {r0, r0 + 1, r1, r1 + 1};
                double* uv[4] = {&u0.x, &u0.y, &u1.x, &u1.y};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rows[q] < n) {
                        *uv[q] -= ra * rec.qi[rows[q]];
                        if (rec.qim1) *uv[q] -= rb * rec.qim1[rows[q]];
                        rec.r0_out[rows[q]] = *uv[q];
                    }
            }
        }
        for (int j0 = 0; j0 < m; j0 += kJB) {
            double acc[kJB];
            if (full && j0 + kJB <= m) {
                double2 a[kJB], b[kJB];
#pragma unroll
                for (int j = 0; j < kJB; ++j) {
                    const double* col = Q + (int64_t)(j0 + j) * ldq;
                    a[j] = ldg2_stream(col + r0);
                    b[j] = ldg2_stream(col + r1);
                }
#pragma unroll
                for (int j = 0; j < kJB; ++j)
                    acc[j] = a[j].x * u0.x + a[j].y * u0.y + b[j].x * u1.x + b[j].y * u1.y;
            } else {
#pragma unroll
                for (int j = 0; j < kJB; ++j) {
                    acc[j] = 0.0;
                    if (j0 + j < m) {
                        const double* col = Q + (int64_t)(j0 + j) * ldq;
                        if (r0 < n) acc[j] += col[r0] * u0.x;
                        if (r0 + 1 < n) acc[j] += col[r0 + 1] * u0.y;
                        if (r1 < n) acc[j] += col[r1] * u1.x;
                        if (r1 + 1 < n) acc[j] += col[r1 + 1] * u1.y;
                    }
                }
            }
            const double tot = warp_reduce8(acc, lane);
            const int j = j0 + ((lane >> 2) & 7);
            if ((lane & 3) == 0 && j < m) my[j] += tot;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += kRThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += wacc[w * m + j];
        partials[(size_t)blockIdx.x * m + j] = s;
    }
}

// ---- pass 2: r = u + sign * Q[:, :m] c ;  partial |r|^2 ---------------------------------------------
__global__ void __launch_bounds__(kRThreads, 2)
reorth_update_kernel(const double* __restrict__ Q, int64_t ldq, const double* __restrict__ u,
                     const double* __restrict__ c, double sign, int64_t n, int m, double* __restrict__ r,
                     double* __restrict__ partials, const PeerPtrs peers) {
    extern __shared__ double cs[];                   // m coefficients (pre-multiplied by sign)
    __shared__ double red[32];
    for (int j = threadIdx.x; j < m; j += kRThreads) cs[j] = sign * c[j];
    __syncthreads();
    const int64_t ntiles = (n + kTileRows - 1) / kTileRows;
    double nrm = 0.0;

    for (int64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
        const int64_t t = ntiles - 1 - tt;                            // reverse of pass 1 (L2 reuse)
        const int64_t r0 = t * kTileRows + 2 * threadIdx.x;
        const int64_t r1 = r0 + kTileRows / 2;
        const bool full = (t + 1) * kTileRows <= n;
        double2 x0 = make_double2(0.0, 0.0), x1 = make_double2(0.0, 0.0);
        if (full) {
            if (u) {
                x0 = ldg2(u + r0);
                x1 = ldg2(u + r1);
            }
            int j = m;                                                // columns also in reverse
            for (; j >= 8; j -= 8) {
                double2 a[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double* col = Q + (int64_t)(j - 1 - q) * ldq;
                    a[q] = ldg2_stream(col + r0);
                    b[q] = ldg2_stream(col + r1);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double cj = cs[j - 1 - q];
                    x0.x += cj * a[q].x;
                    x0.y += cj * a[q].y;
                    x1.x += cj * b[q].x;
                    x1.y += cj * b[q].y;
                }
            }
            for (; j >= 1; --j) {
                const double* col = Q + (int64_t)(j - 1) * ldq;
                const double2 a = ldg2_stream(col + r0), b = ldg2_stream(col + r1);
                const double cj = cs[j - 1];
                x0.x += cj * a.x;
                x0.y += cj * a.y;
                x1.x += cj * b.x;
                x1.y += cj * b.y;
            }
            stg2(r + r0, x0);
            stg2(r + r1, x1);
            for (int pj = 0; pj < peers.n; ++pj) {       // fused exchange: NVLink stores into the partners' arenas
                stg2(peers.p[pj] + r0, x0);
                stg2(peers.p[pj] + r1, x1);
            }
        } else {
            const int64_t rows[4] = {r0, r0 + 1, r1, r1 + 1};
            double xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xv[q] = (u && rows[q] < n) ? u[rows[q]] : 0.0;
            for (int j = 0; j < m; ++j) {
                const double* col = Q + (int64_t)j * ldq;
                const double cj = cs[j];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rows[q] < n) xv[q] += cj * col[rows[q]];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (rows[q] < n) {
                    r[rows[q]] = xv[q];
                    for (int pj = 0; pj < peers.n; ++pj) peers.p[pj][rows[q]] = xv[q];
                }
            x0 = make_double2(xv[0], xv[1]);
            x1 = make_double2(xv[2], xv[3]);
        }
        nrm += x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y;
    }
    if (partials) {
        const double tot = block_sum(nrm, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = tot;
    }
}

// `m` = partial sums each CTA writes (pass 1: one per column): grid * m must fit ctx->partials.
static inline int reorth_grid(const dsea_ctx* ctx, int64_t n, int m = 1) {
    const int64_t ntiles = (n + kTileRows - 1) / kTileRows;
    int64_t cap = (int64_t)ctx->num_sms * ctx->reorth_ctas_per_sm;
    if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
    if (cap * m > kPartialDoubles) cap = kPartialDoubles / m;
    if (cap < 1) cap = 1;
    return (int)(ntiles < cap ? ntiles : cap);
}

int reorth_dots(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const double* Q, const double* u, double* c_out,
                cudaStream_t st, const Recurrence* rec) {
    const int grid = reorth_grid(ctx, n, m);
    const size_t smem = (size_t)8 * m * sizeof(double);
    DSEA_ARG(smem <= 200 * 1024, "too many Lanczos vectors for the reorth accumulators");
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        DSEA_CUDA(cudaFuncSetAttribute(reorth_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        smem_set = 200 * 1024;
    }
    const int tok =
        prof_begin(ctx, PK_REORTH_DOTS, 8.0 * (double)n * (m + 1 + (rec ? 3 : 0)), st);
    Recurrence rc;
    rc.qi = rc.qim1 = nullptr;
    rc.alpha = rc.beta = nullptr;
    rc.r0_out = nullptr;
    if (rec) rc = *rec;
    reorth_dots_kernel<<<grid, kRThreads, smem, st>>>(Q, ldq, u, n, m, ctx->partials, rc);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return finalize_reduce(ctx, grid, m, c_out, st);
}

int reorth_update(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const double* Q, const double* u, const double* c,
                  double sign, double* r_out, double* norm2_out, cudaStream_t st, const PeerPtrs* peers) {
    const int grid = reorth_grid(ctx, n);
    const size_t smem = (size_t)m * sizeof(double);
    const int tok =
        prof_begin(ctx, u ? PK_REORTH_UPDATE : PK_RITZ, 8.0 * (double)n * (m + (u ? 2 : 1)), st);
    PeerPtrs pp;
    pp.n = 0;
    if (peers) pp = *peers;
    reorth_update_kernel<<<grid, kRThreads, smem, st>>>(Q, ldq, u, c, sign, n, m, r_out,
                                                       norm2_out ? ctx->partials : nullptr, pp);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    if (norm2_out) {
        DSEA_TRY(finalize_reduce(ctx, grid, 1, norm2_out, st));
    }
    return DSEA_OK;
}

}  // namespace dsea
