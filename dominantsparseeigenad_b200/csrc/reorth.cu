// reorth.cu — full re-orthogonalisation as a fused two-pass tall-skinny GEMV (K2a/K2b) and the
// Ritz-vector GEMV (K5).  Replaces `r -= Q[:, :i] (Q[:, :i]^T r)` (Lanczos.py:66) and
// `Qk @ eigvectors` (Lanczos.py:99).
//
// Layout: the basis is COLUMN-contiguous (column j at Q + j*ldq), the opposite of the reference's
// row-major (n, k) tensor whose column views have stride k.
//
//   pass 1 (reorth_dots)   c = Q[:, :m]^T u        reads s n m + 8 n bytes        (s = bytes per basis element)
//   pass 2 (reorth_update) r = u - Q[:, :m] c      reads s n m + 8 n, writes 8 n; |r|^2 in the epilogue
//
// Both are HBM-streaming kernels: a CTA owns tiles of 256 x 2 x (16 / s) rows (two 16-byte basis vectors per
// thread and column), keeps the u / r tile in registers as fp64 and streams the m column segments past it
// with >= 16 independent 16 B loads in flight per thread.  Pass 2 walks the tiles in the opposite order to
// pass 1 so that it starts on the part of Q that pass 1 left in the 126 MB L2.
// Reductions are deterministic: per-warp shared-memory accumulators -> per-CTA partials (fixed order)
// -> one finalize CTA per column (fixed order) -> cross-rank sum when sharded.
//
// Basis element type.  double (s = 8) is the reference's precision (Lanczos.py:43,49).  float (s = 4) is the
// opt-in shadow basis ("basis_fp32"): the stored Lanczos vectors are rounded to fp32 — and those rounded values
// ARE the Lanczos vectors (the matvec input and the recurrence use them too, see api.cu) — while every
// accumulation stays fp64.  It halves the dominant HBM traffic and the basis footprint (N = 30, k = 200 on 8 GPUs:
// 107 GB per GPU instead of 215 GB); the eigenpair is then polished in fp64 (api.cu, lanczos_polish).
#include "common.cuh"

namespace dsea {

constexpr int kRThreads = 256;
constexpr int kJB = 8;                                   // columns per register chunk in pass 1

template <typename QT> struct QTraits;
template <> struct QTraits<double> {
    static constexpr int R = 2;                          // rows per 16-byte vector
    typedef double2 V;
    static __device__ __forceinline__ V load(const double* p) { return ldg2_stream(p); }
    static __device__ __forceinline__ void widen(const V& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
};
template <> struct QTraits<float> {
    static constexpr int R = 4;
    typedef float4 V;
    static __device__ __forceinline__ V load(const float* p) {
        V r;
        asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "l"(p));
        return r;
    }
    static __device__ __forceinline__ void widen(const V& v, double (&o)[4]) {
        o[0] = (double)v.x; o[1] = (double)v.y; o[2] = (double)v.z; o[3] = (double)v.w;
    }
};

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

// Three-term recurrence applied in the prologue of pass 1 (typed view of `Recurrence`).
template <typename QT>
struct RecT {
    const QT* qi;
    const QT* qim1;
    const double* alpha;
    const double* beta;
    double* r0_out;
    const double* alpha_partials;
    int n_alpha;
    int skip_cols;      // 1 or 2: qi (and qim1) ARE the last columns of Q[:, :m]; their dots come from the prologue's registers
};

// ---- pass 1: partials[cta][j] = sum over the CTA's rows of Q[row, j] * u[row] -------------------
template <typename QT>
__global__ void __launch_bounds__(kRThreads, 2)
reorth_dots_kernel(const QT* __restrict__ Q, int64_t ldq, const double* __restrict__ u, int64_t n, int m,
                   double* __restrict__ partials, const RecT<QT> rec) {
    pdl_prologue();
    typedef QTraits<QT> TR;
    constexpr int R = TR::R;
    constexpr int kTileRows = kRThreads * 2 * R;
    extern __shared__ double wacc[];                 // [8 warps][m] per-warp accumulators
    // three-term recurrence folded into the prologue: r0 = u - alpha q_i - beta q_{i-1}  (Lanczos.py:61)
    const double ra = rec.qi ? (rec.n_alpha > 0 ? sum_partials_seq(rec.alpha_partials, rec.n_alpha) : *rec.alpha) : 0.0;
    const double rb = rec.qim1 ? *rec.beta : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < 8 * m; j += kRThreads) wacc[j] = 0.0;
    __syncthreads();
    double* my = wacc + warp * m;
    const int64_t ntiles = (n + kTileRows - 1) / kTileRows;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t rA = t * kTileRows + R * threadIdx.x;           // rows rA .. rA+R-1
        const int64_t rB = rA + kTileRows / 2;                        // rows rB .. rB+R-1
        const bool full = (t + 1) * kTileRows <= n;
        double ua[R], ub[R];
        if (full) {
#pragma unroll
            for (int e = 0; e < R; e += 2) {
                const double2 x = ldg2(u + rA + e), y = ldg2(u + rB + e);
                ua[e] = x.x; ua[e + 1] = x.y; ub[e] = y.x; ub[e + 1] = y.y;
            }
        } else {
#pragma unroll
            for (int e = 0; e < R; ++e) {
                ua[e] = rA + e < n ? u[rA + e] : 0.0;
                ub[e] = rB + e < n ? u[rB + e] : 0.0;
            }
        }
        if (rec.qi) {
            double di = 0.0, dm = 0.0;                               // q_i . r0 and q_{i-1} . r0 of this thread's rows
            if (full) {
                double qa[R], qb[R], pa[R], pb[R];
                TR::widen(TR::load(rec.qi + rA), qa);
                TR::widen(TR::load(rec.qi + rB), qb);
#pragma unroll
                for (int e = 0; e < R; ++e) { ua[e] -= ra * qa[e]; ub[e] -= ra * qb[e]; }
                if (rec.qim1) {
                    TR::widen(TR::load(rec.qim1 + rA), pa);
                    TR::widen(TR::load(rec.qim1 + rB), pb);
#pragma unroll
                    for (int e = 0; e < R; ++e) { ua[e] -= rb * pa[e]; ub[e] -= rb * pb[e]; }
                }
#pragma unroll
                for (int e = 0; e < R; e += 2) {
                    stg2(rec.r0_out + rA + e, make_double2(ua[e], ua[e + 1]));
                    stg2(rec.r0_out + rB + e, make_double2(ub[e], ub[e + 1]));
                }
                if (rec.skip_cols > 0) {
#pragma unroll
                    for (int e = 0; e < R; ++e) di += qa[e] * ua[e];
#pragma unroll
                    for (int e = 0; e < R; ++e) di += qb[e] * ub[e];
                    if (rec.skip_cols > 1) {
#pragma unroll
                        for (int e = 0; e < R; ++e) dm += pa[e] * ua[e];
#pragma unroll
                        for (int e = 0; e < R; ++e) dm += pb[e] * ub[e];
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < R; ++e) {
                    if (rA + e < n) {
                        const double qv = (double)rec.qi[rA + e], pv = rec.qim1 ? (double)rec.qim1[rA + e] : 0.0;
                        ua[e] -= ra * qv;
                        if (rec.qim1) ua[e] -= rb * pv;
                        rec.r0_out[rA + e] = ua[e];
                        di += qv * ua[e];
                        dm += pv * ua[e];
                    }
                    if (rB + e < n) {
                        const double qv = (double)rec.qi[rB + e], pv = rec.qim1 ? (double)rec.qim1[rB + e] : 0.0;
                        ub[e] -= ra * qv;
                        if (rec.qim1) ub[e] -= rb * pv;
                        rec.r0_out[rB + e] = ub[e];
                        di += qv * ub[e];
                        dm += pv * ub[e];
                    }
                }
            }
            if (rec.skip_cols > 0) {                                  // columns m-1 (and m-2) are not streamed again below
                di = warp_sum(di);
                if (rec.skip_cols > 1) dm = warp_sum(dm);
                if (lane == 0) {
                    my[m - 1] += di;
                    if (rec.skip_cols > 1) my[m - 2] += dm;
                }
            }
        }
        const int mloop = m - rec.skip_cols;
        for (int j0 = 0; j0 < mloop; j0 += kJB) {
            double acc[kJB];
            if (full && j0 + kJB <= mloop) {
                typename TR::V a[kJB], b[kJB];
#pragma unroll
                for (int j = 0; j < kJB; ++j) {
                    const QT* col = Q + (int64_t)(j0 + j) * ldq;
                    a[j] = TR::load(col + rA);
                    b[j] = TR::load(col + rB);
                }
#pragma unroll
                for (int j = 0; j < kJB; ++j) {
                    double wa[R], wb[R];
                    TR::widen(a[j], wa);
                    TR::widen(b[j], wb);
                    double s = wa[0] * ua[0];
#pragma unroll
                    for (int e = 1; e < R; ++e) s += wa[e] * ua[e];
#pragma unroll
                    for (int e = 0; e < R; ++e) s += wb[e] * ub[e];
                    acc[j] = s;
                }
            } else {
#pragma unroll
                for (int j = 0; j < kJB; ++j) {
                    acc[j] = 0.0;
                    if (j0 + j < mloop) {
                        const QT* col = Q + (int64_t)(j0 + j) * ldq;
#pragma unroll
                        for (int e = 0; e < R; ++e) {
                            if (rA + e < n) acc[j] += (double)col[rA + e] * ua[e];
                            if (rB + e < n) acc[j] += (double)col[rB + e] * ub[e];
                        }
                    }
                }
            }
            const double tot = warp_reduce8(acc, lane);
            const int j = j0 + ((lane >> 2) & 7);
            if ((lane & 3) == 0 && j < mloop) my[j] += tot;
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
template <typename QT>
__global__ void __launch_bounds__(kRThreads, 2)
reorth_update_kernel(const QT* __restrict__ Q, int64_t ldq, const double* __restrict__ u,
                     const double* __restrict__ c, double sign, int64_t n, int m, double* __restrict__ r,
                     double* __restrict__ partials, const PeerPtrs peers) {
    pdl_prologue();
    typedef QTraits<QT> TR;
    constexpr int R = TR::R;
    constexpr int kTileRows = kRThreads * 2 * R;
    extern __shared__ double cs[];                   // m coefficients (pre-multiplied by sign)
    __shared__ double red[32];
    for (int j = threadIdx.x; j < m; j += kRThreads) cs[j] = sign * c[j];
    __syncthreads();
    const int64_t ntiles = (n + kTileRows - 1) / kTileRows;
    double nrm = 0.0;

    for (int64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
        const int64_t t = ntiles - 1 - tt;                            // reverse of pass 1 (L2 reuse)
        const int64_t rA = t * kTileRows + R * threadIdx.x;
        const int64_t rB = rA + kTileRows / 2;
        const bool full = (t + 1) * kTileRows <= n;
        double xa[R], xb[R];
#pragma unroll
        for (int e = 0; e < R; ++e) xa[e] = xb[e] = 0.0;
        if (full) {
            if (u) {
#pragma unroll
                for (int e = 0; e < R; e += 2) {
                    const double2 x = ldg2(u + rA + e), y = ldg2(u + rB + e);
                    xa[e] = x.x; xa[e + 1] = x.y; xb[e] = y.x; xb[e + 1] = y.y;
                }
            }
            int j = m;                                                // columns also in reverse
            for (; j >= 8; j -= 8) {
                typename TR::V a[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const QT* col = Q + (int64_t)(j - 1 - q) * ldq;
                    a[q] = TR::load(col + rA);
                    b[q] = TR::load(col + rB);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double cj = cs[j - 1 - q];
                    double wa[R], wb[R];
                    TR::widen(a[q], wa);
                    TR::widen(b[q], wb);
#pragma unroll
                    for (int e = 0; e < R; ++e) { xa[e] += cj * wa[e]; xb[e] += cj * wb[e]; }
                }
            }
            for (; j >= 1; --j) {
                const QT* col = Q + (int64_t)(j - 1) * ldq;
                double wa[R], wb[R];
                TR::widen(TR::load(col + rA), wa);
                TR::widen(TR::load(col + rB), wb);
                const double cj = cs[j - 1];
#pragma unroll
                for (int e = 0; e < R; ++e) { xa[e] += cj * wa[e]; xb[e] += cj * wb[e]; }
            }
#pragma unroll
            for (int e = 0; e < R; e += 2) {
                const double2 va = make_double2(xa[e], xa[e + 1]), vb = make_double2(xb[e], xb[e + 1]);
                stg2(r + rA + e, va);
                stg2(r + rB + e, vb);
                for (int pj = 0; pj < peers.n; ++pj) {   // fused exchange: NVLink stores into the partners' arenas
                    stg2(peers.p[pj] + rA + e, va);
                    stg2(peers.p[pj] + rB + e, vb);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < R; ++e) {
                if (u && rA + e < n) xa[e] = u[rA + e];
                if (u && rB + e < n) xb[e] = u[rB + e];
            }
            for (int j = 0; j < m; ++j) {
                const QT* col = Q + (int64_t)j * ldq;
                const double cj = cs[j];
#pragma unroll
                for (int e = 0; e < R; ++e) {
                    if (rA + e < n) xa[e] += cj * (double)col[rA + e];
                    if (rB + e < n) xb[e] += cj * (double)col[rB + e];
                }
            }
#pragma unroll
            for (int e = 0; e < R; ++e) {
                if (rA + e < n) {
                    r[rA + e] = xa[e];
                    for (int pj = 0; pj < peers.n; ++pj) peers.p[pj][rA + e] = xa[e];
                } else {
                    xa[e] = 0.0;
                }
                if (rB + e < n) {
                    r[rB + e] = xb[e];
                    for (int pj = 0; pj < peers.n; ++pj) peers.p[pj][rB + e] = xb[e];
                } else {
                    xb[e] = 0.0;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < R; ++e) nrm += xa[e] * xa[e] + xb[e] * xb[e];
    }
    if (partials) {
        const double tot = block_sum(nrm, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = tot;
    }
}

// `m` = partial sums each CTA writes (pass 1: one per column): grid * m must fit ctx->partials.
static inline int reorth_grid(const dsea_ctx* ctx, int64_t n, int tile_rows, int m = 1) {
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    int64_t cap = (int64_t)ctx->num_sms * ctx->reorth_ctas_per_sm;
    if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
    if (cap * m > kDotPartialsOffset) cap = kDotPartialsOffset / m;     // never reach into the matvec's dot-partials region
    if (cap < 1) cap = 1;
    return (int)(ntiles < cap ? ntiles : cap);
}

template <typename QT>
static int reorth_dots_t(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const QT* Q, const double* u, double* c_out,
                         cudaStream_t st, const Recurrence* rec) {
    const int grid = reorth_grid(ctx, n, kRThreads * 2 * QTraits<QT>::R, m);
    const size_t smem = (size_t)8 * m * sizeof(double);
    DSEA_ARG(smem <= 200 * 1024, "too many Lanczos vectors for the reorth accumulators");
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        DSEA_CUDA(cudaFuncSetAttribute(reorth_dots_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        smem_set = 200 * 1024;
    }
    const double s = (double)sizeof(QT);
    RecT<QT> rc;
    rc.qi = rc.qim1 = nullptr;
    rc.alpha = rc.beta = nullptr;
    rc.r0_out = nullptr;
    rc.alpha_partials = nullptr;
    rc.n_alpha = 0;
    rc.skip_cols = 0;
    if (rec) {
        // Lanczos: q_i and q_{i-1} are the last two stored columns; the prologue already holds them in registers
        if (m >= 1 && rec->qi == (const void*)(Q + (int64_t)(m - 1) * ldq)) {
            rc.skip_cols = 1;
            if (m >= 2 && rec->qim1 == (const void*)(Q + (int64_t)(m - 2) * ldq)) rc.skip_cols = 2;
        }
        rc.alpha_partials = rec->alpha_partials;
        rc.n_alpha = rec->n_alpha;
        rc.qi = (const QT*)rec->qi;
        rc.qim1 = (const QT*)rec->qim1;
        rc.alpha = rec->alpha;
        rc.beta = rec->beta;
        rc.r0_out = rec->r0_out;
    }
    // bytes: u (8) + the m columns (s each; q_i / q_{i-1} are read ONCE, in the prologue) [+ r0 written (8) + recurrence
    // operands that are not stored columns]
    const int tok = prof_begin(ctx, PK_REORTH_DOTS,
                               (double)n * (s * m + 8.0 + (rec ? 8.0 + s * (2 - rc.skip_cols) * (rc.qim1 ? 1.0 : 0.5) : 0.0)), st);
    launch_k(ctx, reorth_dots_kernel<QT>, dim3(grid), dim3(kRThreads), smem, st, Q, ldq, u, n, m, ctx->partials, rc);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return finalize_reduce(ctx, grid, m, c_out, st);
}

template <typename QT>
static int reorth_update_t(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const QT* Q, const double* u, const double* c,
                           double sign, double* r_out, double* norm2_out, cudaStream_t st, const PeerPtrs* peers,
                           bool defer_norm) {
    const int grid = reorth_grid(ctx, n, kRThreads * 2 * QTraits<QT>::R);
    const size_t smem = (size_t)m * sizeof(double);
    const double s = (double)sizeof(QT);
    const int tok = prof_begin(ctx, u ? PK_REORTH_UPDATE : PK_RITZ, (double)n * (s * m + (u ? 16.0 : 8.0)), st);
    PeerPtrs pp;
    pp.n = 0;
    if (peers) pp = *peers;
    launch_k(ctx, reorth_update_kernel<QT>, dim3(grid), dim3(kRThreads), smem, st, Q, ldq, u, c, sign, n, m, r_out,
                                                           norm2_out ? ctx->partials : nullptr, pp);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    ctx->pending_norm_n = 0;
    if (norm2_out) {
        if (defer_norm && ctx->world == 1 && ctx->fuse_small) ctx->pending_norm_n = grid;
        else DSEA_TRY(finalize_reduce(ctx, grid, 1, norm2_out, st));
    }
    return DSEA_OK;
}

int reorth_dots(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const double* Q, const double* u, double* c_out,
                cudaStream_t st, const Recurrence* rec) {
    return reorth_dots_t<double>(ctx, n, ldq, m, Q, u, c_out, st, rec);
}

int reorth_update(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const double* Q, const double* u, const double* c,
                  double sign, double* r_out, double* norm2_out, cudaStream_t st, const PeerPtrs* peers, bool defer_norm) {
    return reorth_update_t<double>(ctx, n, ldq, m, Q, u, c, sign, r_out, norm2_out, st, peers, defer_norm);
}

int reorth_dots_f32(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const float* Q, const double* u, double* c_out,
                    cudaStream_t st, const Recurrence* rec) {
    return reorth_dots_t<float>(ctx, n, ldq, m, Q, u, c_out, st, rec);
}

int reorth_update_f32(dsea_ctx* ctx, int64_t n, int64_t ldq, int m, const float* Q, const double* u, const double* c,
                      double sign, double* r_out, double* norm2_out, cudaStream_t st, const PeerPtrs* peers,
                      bool defer_norm) {
    return reorth_update_t<float>(ctx, n, ldq, m, Q, u, c, sign, r_out, norm2_out, st, peers, defer_norm);
}

// ---- K3 for the fp32 shadow basis: q = fl32(r * s) stored twice ---------------------------------------------------
// q32 receives the rounded vector (the basis column), q64 the same values widened (the next matvec's input).
__global__ void __launch_bounds__(256) scale_round_kernel(const double* __restrict__ r, const double* __restrict__ ps,
                                                          double* __restrict__ q64, float* __restrict__ q32, int64_t n) {
    pdl_prologue();
    const double s = *ps;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const double2 a = ldg2(r + 4 * i), b = ldg2(r + 4 * i + 2);
        float4 f;
        f.x = (float)(a.x * s); f.y = (float)(a.y * s); f.z = (float)(b.x * s); f.w = (float)(b.y * s);
        *reinterpret_cast<float4*>(q32 + 4 * i) = f;
        stg2(q64 + 4 * i, make_double2((double)f.x, (double)f.y));
        stg2(q64 + 4 * i + 2, make_double2((double)f.z, (double)f.w));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = n4 * 4; i < n; ++i) {
            const float f = (float)(r[i] * s);
            q32[i] = f;
            q64[i] = (double)f;
        }
    }
}

int scale_round_store(dsea_ctx* ctx, int64_t n, const double* r, const double* scale, double* q64, float* q32,
                      cudaStream_t st) {
    int64_t want = (n / 4 + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap ? want : cap);
    const int tok = prof_begin(ctx, PK_NORMALISE, 20.0 * (double)n, st);
    launch_k(ctx, scale_round_kernel, dim3(grid), dim3(256), 0, st, r, scale, q64, q32, n);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

}  // namespace dsea
