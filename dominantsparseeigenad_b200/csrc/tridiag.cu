// tridiag.cu — extreme eigenpair(s) of the k x k Lanczos tridiagonal in ONE CTA (K4).
//
// Replaces `torch.symeig(T)` on a densified T (Lanczos.py:76,98), which costs O(k^3), a host sync and
// k eigenvectors when only the first/last is used (Lanczos.py:100-105).
//   eigenvalue : parallel multisection on the Sturm count (256 shifts per round, ~8 rounds to 1 ulp)
//   eigenvector: inverse iteration on (T - theta I) with a partially pivoted tridiagonal LU (one lane
//                per requested end of the spectrum; O(k) per sweep)
// The effective size k_eff (Lanczos breakdown) is read from device memory so no host sync is needed.
#include <float.h>

#include "common.cuh"

namespace dsea {

constexpr int kTriThreads = 256;

__device__ __forceinline__ int sturm_count(const double* a, const double* b2, int k, double x, double pivmin) {
    // number of eigenvalues <= x  (LAPACK dlaebz recurrence)
    int cnt = 0;
    double d = a[0] - x;
    if (fabs(d) < pivmin) d = -pivmin;
    cnt += (d <= 0.0);
    for (int j = 1; j < k; ++j) {
        d = (a[j] - x) - b2[j - 1] / d;
        if (fabs(d) < pivmin) d = -pivmin;
        cnt += (d <= 0.0);
    }
    return cnt;
}

// Finds the m-th smallest eigenvalue (1-based) inside [lo, hi]; all threads return the same value.
__device__ double multisection(const double* a, const double* b2, int k, int m, double lo, double hi, double pivmin,
                               int* cnt_s, double* bnd_s) {
    const int t = threadIdx.x;
    for (int round = 0; round < 24; ++round) {
        const double w = hi - lo;
        const double tol = 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin;
        if (!(w > tol)) break;
        const double xt = lo + w * ((double)(t + 1) / (double)(kTriThreads + 1));
        cnt_s[t] = sturm_count(a, b2, k, xt, pivmin);
        __syncthreads();
        if (t == 0) {
            double nlo = lo, nhi = hi;
            int first = kTriThreads;
            for (int q = 0; q < kTriThreads; ++q)
                if (cnt_s[q] >= m) { first = q; break; }
            if (first < kTriThreads) nhi = lo + w * ((double)(first + 1) / (double)(kTriThreads + 1));
            if (first > 0) nlo = lo + w * ((double)first / (double)(kTriThreads + 1));
            if (!(nlo < nhi)) { nlo = lo; nhi = lo; }      // rounding collapsed the bracket: stop
            bnd_s[0] = nlo;
            bnd_s[1] = nhi;
        }
        __syncthreads();
        const double nlo = bnd_s[0], nhi = bnd_s[1];
        __syncthreads();
        if (nlo == lo && nhi == hi) break;
        lo = nlo;
        hi = nhi;
    }
    return 0.5 * (lo + hi);
}

// Inverse iteration for the eigenvector of eigenvalue theta.  Single thread; scratch w has 5k doubles.
__device__ void inverse_iteration(const double* a, const double* b, int k, double theta, double tnorm, double* y,
                                  double* w) {
    if (k == 1) { y[0] = 1.0; return; }
    double* d = w;            // U diagonal
    double* du = w + k;       // U first super-diagonal
    double* du2 = w + 2 * k;  // U second super-diagonal
    double* l = w + 3 * k;    // multipliers
    double* pv = w + 4 * k;   // 1.0 if rows i, i+1 were swapped
    const double tiny = fmax(DBL_EPSILON * tnorm, DBL_MIN * 1e16);
    for (int i = 0; i < k; ++i) {
        d[i] = a[i] - theta;
        du[i] = (i < k - 1) ? b[i] : 0.0;
        du2[i] = 0.0;
    }
    for (int i = 0; i < k - 1; ++i) {
        const double sub = b[i];
        if (fabs(d[i]) >= fabs(sub)) {
            double piv = d[i];
            if (fabs(piv) < tiny) { piv = (piv < 0.0 ? -tiny : tiny); d[i] = piv; }
            const double f = sub / piv;
            l[i] = f;
            d[i + 1] -= f * du[i];
            pv[i] = 0.0;
        } else {
            const double f = d[i] / sub;
            l[i] = f;
            d[i] = sub;
            const double tmp = du[i];
            du[i] = d[i + 1];
            d[i + 1] = tmp - f * d[i + 1];
            if (i < k - 2) {
                du2[i] = du[i + 1];
                du[i + 1] = -f * du[i + 1];
            }
            pv[i] = 1.0;
        }
    }
    if (fabs(d[k - 1]) < tiny) d[k - 1] = (d[k - 1] < 0.0 ? -tiny : tiny);

    // deterministic pseudo-random start vector
    uint32_t lcg = 0x9E3779B9u;
    for (int i = 0; i < k; ++i) {
        lcg = lcg * 1664525u + 1013904223u;
        y[i] = 0.5 + (double)(lcg >> 8) * (1.0 / 16777216.0);
    }
    for (int it = 0; it < 5; ++it) {
        // forward: L^{-1} P y
        for (int i = 0; i < k - 1; ++i) {
            if (pv[i] == 0.0) {
                y[i + 1] -= l[i] * y[i];
            } else {
                const double tmp = y[i];
                y[i] = y[i + 1];
                y[i + 1] = tmp - l[i] * y[i];
            }
        }
        // backward: U^{-1}
        y[k - 1] /= d[k - 1];
        if (k >= 2) y[k - 2] = (y[k - 2] - du[k - 2] * y[k - 1]) / d[k - 2];
        for (int i = k - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / d[i];
        // normalise (scale by max first to stay clear of overflow)
        double mx = 0.0;
        for (int i = 0; i < k; ++i) mx = fmax(mx, fabs(y[i]));
        if (!(mx > 0.0) || !isfinite(mx)) {     // pathological: fall back to unit vector
            for (int i = 0; i < k; ++i) y[i] = (i == 0);
            mx = 1.0;
        }
        double s = 0.0;
        for (int i = 0; i < k; ++i) { y[i] /= mx; s += y[i] * y[i]; }
        s = 1.0 / sqrt(s);
        for (int i = 0; i < k; ++i) y[i] *= s;
    }
    // fix the sign so the result is reproducible: largest-magnitude component positive
    int im = 0;
    for (int i = 1; i < k; ++i)
        if (fabs(y[i]) > fabs(y[im])) im = i;
    if (y[im] < 0.0)
        for (int i = 0; i < k; ++i) y[i] = -y[i];
}

__global__ void __launch_bounds__(kTriThreads)
tridiag_kernel(int kmax, int which, const double* __restrict__ alpha, const double* __restrict__ beta,
               const double* __restrict__ keff_ptr, double* __restrict__ evals, double* __restrict__ y_min,
               double* __restrict__ y_max, double* __restrict__ work) {
    pdl_prologue();
    extern __shared__ double sm[];
    __shared__ int cnt_s[kTriThreads];
    __shared__ double bnd_s[2];
    __shared__ double red[2 * 32];
    int k = kmax;
    if (keff_ptr) {
        const int ke = (int)(*keff_ptr);
        if (ke > 0 && ke < k) k = ke;
    }
    double* a = sm;
    double* b = sm + kmax;
    double* b2 = sm + 2 * kmax;
    const int t = threadIdx.x;
    double gl = DBL_MAX, gu = -DBL_MAX, bmax = 0.0;
    for (int j = t; j < k; j += kTriThreads) {
        const double aj = alpha[j];
        const double bj = (j < k - 1) ? beta[j] : 0.0;
        a[j] = aj;
        b[j] = bj;
        b2[j] = bj * bj;
        const double bl = (j > 0) ? fabs(beta[j - 1]) : 0.0;
        const double rad = bl + fabs(bj);
        gl = fmin(gl, aj - rad);
        gu = fmax(gu, aj + rad);
        bmax = fmax(bmax, bj * bj);
    }
    // block min / max
    for (int o = 16; o > 0; o >>= 1) {
        gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
        bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    }
    if ((t & 31) == 0) { red[t >> 5] = gl; red[8 + (t >> 5)] = gu; red[16 + (t >> 5)] = bmax; }
    __syncthreads();
    gl = red[0]; gu = red[8]; bmax = red[16];
    for (int q = 1; q < kTriThreads / 32; ++q) {
        gl = fmin(gl, red[q]);
        gu = fmax(gu, red[8 + q]);
        bmax = fmax(bmax, red[16 + q]);
    }
    const double tnorm = fmax(fabs(gl), fabs(gu));
    const double pivmin = DBL_MIN * fmax(1.0, bmax);
    const double pad = 2.1 * tnorm * DBL_EPSILON * k + 4.2 * pivmin;
    gl -= pad;
    gu += pad;
    __syncthreads();

    double th_min = 0.0, th_max = 0.0;
    if (which == DSEA_MIN || which == DSEA_BOTH) th_min = multisection(a, b2, k, 1, gl, gu, pivmin, cnt_s, bnd_s);
    if (which == DSEA_MAX || which == DSEA_BOTH) th_max = multisection(a, b2, k, k, gl, gu, pivmin, cnt_s, bnd_s);
    __syncthreads();
    if (t == 0 && (which == DSEA_MIN || which == DSEA_BOTH)) {
        evals[0] = th_min;
        inverse_iteration(a, b, k, th_min, tnorm, y_min, work);
        for (int j = k; j < kmax; ++j) y_min[j] = 0.0;
    }
    if (t == 32 && (which == DSEA_MAX || which == DSEA_BOTH)) {
        evals[1] = th_max;
        inverse_iteration(a, b, k, th_max, tnorm, y_max, work + 5 * kmax);
        for (int j = k; j < kmax; ++j) y_max[j] = 0.0;
    }
}

int tridiag_extreme(dsea_ctx* ctx, int k, int which, const double* alpha, const double* beta, const double* keff,
                    double* evals, double* y_min, double* y_max, cudaStream_t st) {
    DSEA_ARG(k >= 1 && k <= kMaxK, "k out of range for the tridiagonal solver");
    const size_t smem = (size_t)3 * k * sizeof(double);
    static bool attr = false;
    if (!attr) {
        DSEA_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(3 * kMaxK * sizeof(double))));
        attr = true;
    }
    const int tok = prof_begin(ctx, PK_TRIDIAG, 16.0 * k, st);
    launch_k(ctx, tridiag_kernel, dim3(1), dim3(kTriThreads), smem, st, k, which, alpha, beta, keff, evals, y_min, y_max, ctx->tri_work);
    prof_end(ctx, tok, st);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

}  // namespace dsea
