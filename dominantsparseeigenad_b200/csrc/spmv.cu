// spmv.cu — explicit-matrix operators: CSR (+ parameter diagonal) SpMV (K1s), dense symmetric GEMV (K1d),
// and the element-wise / rank-1 adjoints that go with them.
//   CSR  : u = CSR v + p o v - shift v      (Schrodinger1D.Hsparse, schrodinger1D.py:18-27; adjoint v1 o v2, :29-34)
//   dense: u = A v - shift v                (Lanczos.py:48, CG.py:23); adjoint scale * a b^T (symeig.py:29, CG.py:69)
// Each kernel optionally folds the v.u dot product (needed by CG) into its epilogue.
#include "common.cuh"

namespace dsea {

// ---- CSR, one thread per row (short rows: stencils, banded matrices) ----------------------------
__global__ void __launch_bounds__(256)
csr_row_per_thread_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ colidx,
                          const double* __restrict__ vals, const double* __restrict__ pdiag,
                          const double* __restrict__ shift, const double* __restrict__ v, double* __restrict__ u,
                          double* __restrict__ partials, const double* __restrict__ guard) {
    pdl_prologue();
    __shared__ double red[32];
    if (guard && *guard != 0.0) return;
    const double sh = shift ? *shift : 0.0;
    double part = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
        double s = 0.0;
        const int64_t e = rowptr[row + 1];
        for (int64_t q = rowptr[row]; q < e; ++q) s += vals[q] * v[colidx[q]];
        const double x = v[row];
        s += ((pdiag ? pdiag[row] : 0.0) - sh) * x;
        u[row] = s;
        part += x * s;
    }
    if (partials) {
        part = block_sum(part, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = part;
    }
}

// ---- CSR, one warp per row (long rows) ------------------------------------------------------------
__global__ void __launch_bounds__(256)
csr_row_per_warp_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ colidx,
                        const double* __restrict__ vals, const double* __restrict__ pdiag,
                        const double* __restrict__ shift, const double* __restrict__ v, double* __restrict__ u,
                        double* __restrict__ partials, const double* __restrict__ guard) {
    pdl_prologue();
    __shared__ double red[32];
    if (guard && *guard != 0.0) return;
    const double sh = shift ? *shift : 0.0;
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (int64_t row = wid; row < n; row += nw) {
        double s = 0.0;
        const int64_t e = rowptr[row + 1];
        for (int64_t q = rowptr[row] + lane; q < e; q += 32) s += vals[q] * v[colidx[q]];
        s = warp_sum(s);
        if (lane == 0) {
            const double x = v[row];
            s += ((pdiag ? pdiag[row] : 0.0) - sh) * x;
            u[row] = s;
            part += x * s;
        }
    }
    if (partials) {
        part = block_sum(part, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = part;
    }
}

int csr_apply(dsea_ctx* ctx, const dsea_op* op, const double* pdiag, const double* shift, const double* v,
              double* u, double* dot_out, cudaStream_t st) {
    const int64_t n = op->n_loc;
    const double avg = n > 0 ? (double)op->nnz / (double)n : 0.0;
    int64_t cap = (int64_t)ctx->num_sms * 8;
    int grid;
    if (avg <= 12.0) {
        int64_t want = (n + 255) / 256;
        grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        launch_k(ctx, csr_row_per_thread_kernel, dim3(grid), dim3(256), 0, st, n, op->rowptr, op->colidx, op->vals, pdiag, shift, v, u,
                                                        dot_out ? ctx->partials : nullptr, ctx->guard);
    } else {
        int64_t want = (n + 7) / 8;
        grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        launch_k(ctx, csr_row_per_warp_kernel, dim3(grid), dim3(256), 0, st, n, op->rowptr, op->colidx, op->vals, pdiag, shift, v, u,
                                                      dot_out ? ctx->partials : nullptr, ctx->guard);
    }
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    if (dot_out) DSEA_TRY(finalize_partials(ctx, grid, 1, dot_out, st));
    return DSEA_OK;
}

// ---- dense symmetric GEMV: one warp per row, coalesced along the row --------------------------------
__global__ void __launch_bounds__(256)
dense_gemv_kernel(int64_t n, int64_t ld, const double* __restrict__ A, const double* __restrict__ shift,
                  const double* __restrict__ v, double* __restrict__ u, double* __restrict__ partials,
                  const double* __restrict__ guard) {
    pdl_prologue();
    __shared__ double red[32];
    if (guard && *guard != 0.0) return;
    const double sh = shift ? *shift : 0.0;
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (int64_t row = wid; row < n; row += nw) {
        const double* a = A + row * ld;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t j = lane;
        for (; j + 96 < n; j += 128) {
            s0 += a[j] * v[j];
            s1 += a[j + 32] * v[j + 32];
            s2 += a[j + 64] * v[j + 64];
            s3 += a[j + 96] * v[j + 96];
        }
        for (; j < n; j += 32) s0 += a[j] * v[j];
        double s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) {
            const double x = v[row];
            s -= sh * x;
            u[row] = s;
            part += x * s;
        }
    }
    if (partials) {
        part = block_sum(part, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = part;
    }
}

int dense_apply(dsea_ctx* ctx, const dsea_op* op, const double* shift, const double* v, double* u, double* dot_out,
                cudaStream_t st) {
    const int64_t n = op->n_loc;
    int64_t want = (n + 7) / 8;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    launch_k(ctx, dense_gemv_kernel, dim3(grid), dim3(256), 0, st, n, op->ld, op->A, shift, v, u, dot_out ? ctx->partials : nullptr, ctx->guard);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    if (dot_out) DSEA_TRY(finalize_partials(ctx, grid, 1, dot_out, st));
    return DSEA_OK;
}

// ---- adjoints for explicit matrices ------------------------------------------------------------------
__global__ void __launch_bounds__(256) hadamard_kernel(int64_t n, const double* __restrict__ a,
                                                       const double* __restrict__ b, double* __restrict__ out) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] * b[i];
}

int hadamard(dsea_ctx* ctx, int64_t n, const double* a, const double* b, double* out, cudaStream_t st) {
    int64_t want = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    launch_k(ctx, hadamard_kernel, dim3((int)(want < cap ? (want < 1 ? 1 : want) : cap)), dim3(256), 0, st, n, a, b, out);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

__global__ void __launch_bounds__(256) outer_kernel(int64_t n, double scale, const double* __restrict__ a,
                                                    const double* __restrict__ b, double* __restrict__ out) {
    pdl_prologue();
    const int64_t total = n * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / n, c = i - r * n;
        out[i] = scale * a[r] * b[c];
    }
}

int outer(dsea_ctx* ctx, int64_t n, double scale, const double* a, const double* b, double* out, cudaStream_t st) {
    int64_t want = (n * n + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    launch_k(ctx, outer_kernel, dim3((int)(want < cap ? (want < 1 ? 1 : want) : cap)), dim3(256), 0, st, n, scale, a, b, out);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

}  // namespace dsea
