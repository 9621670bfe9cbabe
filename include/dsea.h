/*
 * dsea.h — C ABI of libdsea.so, the B200 (sm_100a) dominant-eigenpair solver.
 *
 * This is the drop-in boundary for the hot path of buwantaiji/DominantSparseEigenAD
 * (SURVEY.md section 8b).  The reference is pure Python/PyTorch and has no FFI of its own;
 * each entry point below names the reference function it replaces (file:line under the
 * reference checkout).  The Python package `dominantsparseeigenad_b200` binds these with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every `double*` / `int*` data pointer is a DEVICE pointer unless the name ends in
 *     `_host`; the library borrows it for the duration of the call and keeps no reference;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising unless documented;
 *   - all arithmetic is IEEE fp64; vectors are rank-local shards of n_loc contiguous doubles;
 *   - return value: 0 on success, negative on error (message via dsea_last_error()).
 *   - NOT thread-safe per context (the reference is not either: symeig.py:66, CG.py:116).
 */
#ifndef DSEA_H_
#define DSEA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsea_ctx dsea_ctx;   /* device, stream-ordered scratch, optional NCCL communicator */
typedef struct dsea_op dsea_op;     /* linear-operator descriptor (TFIM / CSR+diag / dense)        */

enum { DSEA_OK = 0, DSEA_ERR_CUDA = -1, DSEA_ERR_ARG = -2, DSEA_ERR_NCCL = -3, DSEA_ERR_NOCONV = -4 };
enum { DSEA_MIN = 0, DSEA_MAX = 1, DSEA_BOTH = 2 };            /* Lanczos.py:82-85 `extreme`   */
enum { DSEA_OP_TFIM = 1, DSEA_OP_CSR = 2, DSEA_OP_DENSE = 3 };

const char* dsea_last_error(void);
int dsea_version(void);

/* ---- context ------------------------------------------------------------------------------
 * One per process (one process per GPU).  world==1: no communicator.  world>1: `nccl_id_host`
 * is the 128-byte ncclUniqueId obtained on rank 0 with dsea_nccl_unique_id() and broadcast by
 * the host (torch.distributed store).  The reference has no distributed mode. */
int dsea_nccl_unique_id(void* id128_host);
int dsea_ctx_create(int device, int rank, int world, const void* nccl_id_host, dsea_ctx** out);
int dsea_ctx_destroy(dsea_ctx* ctx);
int dsea_ctx_rank(const dsea_ctx* ctx);
/* 1 when the top-bit exchange runs as peer (NVLink) stores fused into the producing kernels through a
 * CUDA-IPC arena, 0 when it runs as NCCL send/recv (world==1, option "p2p"=0, or IPC unavailable). */
int dsea_ctx_p2p(const dsea_ctx* ctx);
int dsea_ctx_world(const dsea_ctx* ctx);
/* kernel-launch counter (every __global__ launch the library issues), for bench accounting */
int64_t dsea_launch_count(const dsea_ctx* ctx);
/* Per-kernel timing with CUDA events on the launching stream, for the roofline report.  Kinds:
 * 0 matvec, 1 reorth pass 1 (Q^T u), 2 reorth pass 2 (u - Q c), 3 Ritz GEMV, 4 CG vector updates,
 * 5 normalise, 6 tridiagonal solve, 7 adjoint contraction, 8 no-op (launches issued after the CG convergence
 * flag was set: they exit immediately, are credited no bytes and are not counted as launches of their kernel),
 * 9 exchange (barrier + push kernel + barrier of a sharded matvec whose input no producing kernel published).
 * collect() synchronises the device, sums elapsed ms / algorithmic bytes / launches per kind since the last
 * collect, and resets. */
int dsea_profile_enable(dsea_ctx* ctx, int on);
int dsea_profile_collect(dsea_ctx* ctx, int nkinds, double* ms_host, double* bytes_host, int64_t* count_host);
/* tuning knobs: "tfim_tile_bits" (<=14), "tfim_run_bits", "cg_check_every", "reorth_ctas_per_sm" (<=16),
 * "tfim_pipeline" (0: generic sweep kernel only), "tfim_tma" (0: stage contiguous tiles with LDGSTS
 * instead of TMA bulk copies), "tfim_pipe_threads" (512 | 256: 4 or 5 register-resident tile bits),
 * "tfim_direct" (0: a third sweep instead of direct loads for the top local bits), "tfim_fuse_scale" (0:
 * separate normalisation pass per Lanczos step), "tfim_l2_prefetch", "tfim_pipe_adjoint", "tfim_pipe_remote",
 * "cg_fuse_push" (0: separate push pass per CG iteration when sharded), "fuse_small" (0: separate finalize launches
 * instead of consumers summing partials; one GPU), "pdl" (programmatic dependent launch; default 1 on one GPU, 0 sharded), "basis_fp32" (1: opt-in fp32 shadow
 * of the Lanczos basis for the re-orthogonalisation passes), "p2p" (0 disables the peer-memory exchange; set
 * before creating operators), "mailbox" (0: small all-reduces through NCCL instead of the peer-memory mailbox) */
int dsea_ctx_set_option(dsea_ctx* ctx, const char* key, int64_t value);

/* ---- operators ----------------------------------------------------------------------------
 * TFIM:  H(g) = -sum_i (g sx_i + sz_i sz_{i+1}), N spins, periodic; flip indices s^(1<<i) and the
 *        diagonal -(N - 2 popc(s ^ rotl_N s)) are generated by bit arithmetic on the GLOBAL index
 *        (rank << (N - log2 world)) | s_loc.  Replaces TFIM._flips_basis / _diags / H / pHpg
 *        (examples/TFIM/TFIM.py:39-51, 58-65, 91-98).  Parameter: scalar g (device pointer).
 * CSR:   A(p) = CSR + diag(p); p (n doubles, may be NULL) is the parameter.  Covers
 *        Schrodinger1D.Hsparse (examples/schrodinger1D.py:18-27).  Single GPU only.
 * DENSE: symmetric row-major n x n with leading dimension ld (Lanczos.py:48, CG.py:23).          */
int dsea_op_tfim(dsea_ctx* ctx, int N, dsea_op** out);
int dsea_op_csr(dsea_ctx* ctx, int64_t n, int64_t nnz, const int64_t* rowptr, const int64_t* colidx,
                const double* vals, dsea_op** out);
int dsea_op_dense(dsea_ctx* ctx, int64_t n, int64_t ld, const double* A, dsea_op** out);
int dsea_op_destroy(dsea_op* op);
int64_t dsea_op_local_dim(const dsea_op* op);    /* n_loc: rows owned by this rank               */
int64_t dsea_op_work_doubles(const dsea_op* op); /* scratch doubles a matvec needs (recv buffers) */
/* Column stride (in doubles) of every multi-vector buffer: n_loc rounded up to 16 doubles (128 B),
 * so that each Lanczos vector starts on a cache-line boundary. */
int64_t dsea_col_stride(int64_t n_loc);

/* host-callable bit maps, for the bit-exact tests against TFIM.py:39-51 */
int64_t dsea_tfim_flip_index(int N, int64_t s, int i);
double dsea_tfim_diag(int N, int64_t s);
/* The sweep schedule the TFIM kernels use for `local_bits` spin bits per rank with tiles of at most
 * 2^tile_bits doubles (run_bits = 0: automatic, runs of >= 128 bytes): rows {T, c, hshift, b0, ndirect} = tile
 * bits, contiguous low bits, global position of tile bit c, first tile bit the sweep is responsible for, and the
 * number of local bits ABOVE the tile's strided range (positions hshift + T - c ...) that the sweep serves with
 * direct, L2-resident global loads instead of a further sweep (allow_direct != 0; at most 4).  Room for 40 rows.
 * Returns the sweep count. */
int dsea_tfim_plan(int local_bits, int tile_bits, int run_bits, int allow_direct, int* out5x40);

/* u = A(param) v - shift*v.   `shift` may be NULL (0).  `dot_out` (device, may be NULL) receives
 * v.u reduced over all ranks.  `work`: dsea_op_work_doubles() doubles (may be NULL when 0).
 * Replaces the user callable A(v) (Lanczos.py:54,71; CG.py:27,31,34,40,120).                     */
int dsea_matvec(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift,
                const double* v, double* u, double* dot_out, double* work, void* stream);

/* TFIM only: u = (dH/dg) v = -sum_i v[s ^ (1<<i)].  Replaces TFIM.pHpg (TFIM.py:58-65); needed as a
 * differentiable building block of the adjoint contraction for second derivatives.              */
int dsea_tfim_dHdg(dsea_ctx* ctx, const dsea_op* op, const double* v, double* u, double* work, void* stream);

/* out[0] = v1^T (dA/dparam) v2 for scalar parameters (TFIM: -sum_s v1[s] sum_i v2[s^(1<<i)]),
 * or out[0..n) = v1 o v2 for CSR+diag.  Replaces Hadjoint_to_gadjoint (TFIM.py:100-101) and
 * Hadjoint_to_padjoint (schrodinger1D.py:29-34).                                                */
int dsea_adjoint(dsea_ctx* ctx, const dsea_op* op, const double* v1, const double* v2, double* out,
                 double* work, void* stream);

/* ---- Lanczos ------------------------------------------------------------------------------
 * k-step Lanczos with full re-orthogonalisation, tridiagonal eigensolve and Ritz vector(s), all
 * device resident, no host synchronisation.  Replaces Lanczos.Lanczos + symeigLanczos
 * (Lanczos.py:3-105).
 *   Q      k * ldq doubles, ldq = dsea_col_stride(n_loc): COLUMN-contiguous basis (column j at
 *          Q + j*ldq); on entry column 0 holds the start vector q0 (any norm; Lanczos.py:52 draws
 *          randn) — it is normalised in place;
 *   work   dsea_lanczos_work_doubles() doubles;
 *   alpha  k doubles out, beta k doubles out (beta[k-1] unused);
 *   evals  2 doubles out {min, max}; evec_min / evec_max n_loc doubles out (NULL if not wanted);
 *   info_host (may be NULL): 2 int64 {k_effective, breakdown_flag}; passing it synchronises.   */
int64_t dsea_lanczos_work_doubles(const dsea_op* op);
/* Size (in doubles) of the basis buffer `Q` for k vectors: k * dsea_col_stride(n_loc) in the reference's precision.
 * With the opt-in option "basis_fp32" = 1 (TFIM operators, which = DSEA_MIN only) the buffer holds k FLOAT columns
 * (half the bytes: N = 30, k = 200 on 8 GPUs needs 107 GB per GPU instead of 215 GB, and the re-orthogonalisation
 * passes stream half the traffic).  The stored vectors are the fp32-rounded Lanczos vectors, all accumulation stays
 * fp64, and the returned eigenpair is polished in fp64 by one Jacobi-Davidson step (projected CG on
 * P (A - theta) P delta = theta x - A x): evals[0] is the Rayleigh quotient of the polished vector.  Replaces the
 * (n, k) fp64 allocation of Lanczos.py:49.  q0 is still passed as n_loc doubles at the start of `Q`. */
int64_t dsea_lanczos_basis_doubles(const dsea_op* op, int k);
int dsea_lanczos(dsea_ctx* ctx, const dsea_op* op, const double* param, int k, int which,
                 double* Q, double* work, double* alpha, double* beta, double* evals,
                 double* evec_min, double* evec_max, int64_t* info_host, void* stream);

/* Building blocks of the same loop for operators that live in Python (callback A(v)):
 *   start:  normalise column 0 of Q in place;
 *   step i: given u = A q_i:  alpha[i] = q_i . u;  if i < k-1:  r0 = u - alpha[i] q_i - beta[i-1] q_{i-1},
 *           r = r0 - Q[:, :i+1] (Q[:, :i+1]^T r0),  beta[i] = |r|,  Q[:, i+1] = r / beta[i]
 *           — the reference's order of operations (Lanczos.py:61-75): recurrence first, then one
 *           classical Gram-Schmidt sweep;
 *   ritz:   tridiagonal eigensolve + Ritz vector(s)                            (Lanczos.py:98-105) */
int dsea_lanczos_start(dsea_ctx* ctx, int64_t n_loc, double* Q, void* stream);
int dsea_lanczos_step(dsea_ctx* ctx, int64_t n_loc, int k, int i, double* Q, const double* u,
                      double* alpha, double* beta, void* stream);
int dsea_lanczos_ritz(dsea_ctx* ctx, int64_t n_loc, int k, int which, const double* Q,
                      const double* alpha, const double* beta, double* evals, double* evec_min,
                      double* evec_max, int64_t* info_host, void* stream);

/* ---- Arnoldi (non-symmetric family) -------------------------------------------------------------
 * Building blocks for the DominantEig primitives (eig.py:5-152), whose reference implementation calls
 * ARPACK (scipy.sparse.linalg.eigs, two calls: A and A^T) and scipy GMRES.  The Krylov basis is built
 * with the same fused two-pass GEMV kernels as the Lanczos loop, with a second Gram-Schmidt sweep.
 *   start:  normalise column 0 of Q in place; norm_out (device, may be NULL) receives |q0|^2;
 *   step i: given u = A q_i: h = Q[:, :i+1]^T u (two sweeps), Q[:, i+1] = (u - Q h)/|.|,
 *           H[0..i+1, i] written to the column-major (m+1) x m device array H;
 *   combine: out = add + sum_j coef[j] Q[:, j]  (Ritz vector / GMRES update; `add` may be NULL).   */
int dsea_arnoldi_start(dsea_ctx* ctx, int64_t n_loc, double* Q, double* norm_out, void* stream);
int dsea_arnoldi_step(dsea_ctx* ctx, int64_t n_loc, int m, int i, double* Q, const double* u, double* H,
                      void* stream);
int dsea_combine(dsea_ctx* ctx, int64_t n_loc, int m, const double* Q, const double* coef, const double* add,
                 double* out, void* stream);

/* ---- CG -----------------------------------------------------------------------------------
 * Solves (A(param) - shift) x = b by conjugate gradients with the reference's stopping rule
 * |r|_2 < eps (absolute), at most maxit iterations, ONE operator application per iteration.
 * x holds x0 on entry and the solution on exit.  work: dsea_cg_work_doubles() doubles.
 * iters_host (may be NULL) receives the iteration count.  The call synchronises `stream`
 * (the convergence flag is polled every "cg_check_every" iterations from pinned memory).
 * Replaces CG.CG_torch (CG.py:3-41).                                                            */
int64_t dsea_cg_work_doubles(const dsea_op* op);
int dsea_cg(dsea_ctx* ctx, const dsea_op* op, const double* param, const double* shift,
            const double* b, double* x, double* work, double eps, int64_t maxit,
            int64_t* iters_host, void* stream);

/* Fused pieces of one CG iteration for Python-side operators (callback A(v)):
 *   init:   r = b - Ax (Ax supplied), d = r, rr = r.r                          (CG.py:27-30)
 *   update: alpha = rr / (d.Ad); x += alpha d; r -= alpha Ad; rr' = r.r;
 *           beta = rr'/rr; d = r + beta d                                      (CG.py:31-40)
 *   state_host (3 doubles, may be NULL): dsea_cg_init reads {eps, maxit} from it; both calls then
 *   write {|r|, iterations, done} to it, which synchronises `stream`.                            */
int dsea_cg_init(dsea_ctx* ctx, int64_t n_loc, const double* b, const double* Ax, double* r,
                 double* d, double* state_host, void* stream);
int dsea_cg_update(dsea_ctx* ctx, int64_t n_loc, double* x, double* r, double* d, const double* Ad,
                   double* state_host, void* stream);

/* ---- level-1 helpers (all reductions are deterministic two-stage sums + NCCL allreduce) ---- */
/* out[0] = a . b                                                     (torch.matmul(v, w))        */
int dsea_dot(dsea_ctx* ctx, int64_t n_loc, const double* a, const double* b, double* out, void* stream);
/* out = b - (psi . b) psi         (symeig.py:27,80; CG.py:59,67,122,132)                         */
int dsea_project(dsea_ctx* ctx, int64_t n_loc, const double* psi, const double* b, double* out,
                 void* stream);
/* y = a*x + b*y with a, b device scalars (NULL => 1)                                             */
int dsea_axpby(dsea_ctx* ctx, int64_t n_loc, const double* a, const double* x, const double* b,
               double* y, void* stream);
/* out[i,j] = scale * a[i] * b[j], row-major n x n  (symeig.py:29, CG.py:69)                      */
int dsea_outer(dsea_ctx* ctx, int64_t n, double scale, const double* a, const double* b, double* out,
               void* stream);
/* standard normal doubles from Philox4x32-10, counter = global element index (rank offset
 * applied), so the sharded vector equals the single-GPU one.  (Lanczos.py:52, CG.py:58,121)      */
int dsea_randn(dsea_ctx* ctx, int64_t n_loc, uint64_t seed, uint64_t stream_id, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSEA_H_ */
