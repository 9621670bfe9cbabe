// tfim.cu — matrix-free transverse-field Ising operator on the 2^N spin basis (K1, K6).
//
//   u[s] = diag(s) v[s] - g * sum_{i<N} v[s ^ (1<<i)]  (- shift * v[s])
//   diag(s) = -(N - 2 popc(s ^ rotl_N(s)))             (bit-exact restatement of TFIM.py:39-46)
//   flip index s ^ (1<<i)                              (TFIM.py:48-51)
//
// No index table exists: the reference's (2^N, N) int64 table (60 GB at N=28) is replaced by bit
// arithmetic on the global index (rank << L) | s_loc.
//
// Memory plan.  A CTA stages a TILE of 2^T doubles in shared memory and serves every flip whose
// bit lies inside the tile from there, so each vector element is read from HBM once per SWEEP:
//   sweep 0 : tile = 2^T contiguous doubles            -> handles spin bits [0, T)
//   sweep j : tile = 2^c contiguous x 2^h strided runs -> handles h spin bits starting at `hshift`
//             (run stride 2^hshift doubles; c >= 2 keeps every global access a full 32 B sector,
//             c >= 4 a full 128 B line)
//   top log2(world) bits : the shard of rank ^ (1<<j) sits in slot j of this rank's peer arena, stored
//             there over NVLink by the kernel that produced it (reorth pass 2, or push_kernel); the
//             last sweep adds the slots.  Fallback: grouped ncclSend/ncclRecv on a side stream,
//             overlapped with the local sweeps.
// HBM bytes per element: 16 (first sweep: read v, write u) + 24 per extra sweep (read v, read u,
// write u) + 8 per remote bit.  The dot-product epilogue (v.u for CG, or w.u) rides on the last sweep.
#include "common.cuh"

namespace dsea {

struct Sweep {
    int T;        // tile bits
    int c;        // contiguous low bits of the tile
    int hshift;   // global bit position of tile bit c
    int b0;       // first tile bit this sweep is responsible for
};

struct SweepParams {
    const double* v;
    const double* uin;
    double* uout;
    const double* w;        // dot partner (may alias v); nullptr = no dot
    const double* g;
    const double* shift;
    const double* recv;     // nrecv buffers, `recv_stride` doubles apart (NCCL recv buffers or the IPC arena)
    const double* remote_scale;   // device scalar multiplying the remote terms (nullptr = 1)
    uint64_t recv_stride;
    const double* guard;    // if non-null and != 0 the kernel is a no-op (CG converged)
    double* partials;
    uint64_t rank_off;      // rank << L
    uint64_t n_loc;
    uint64_t ntiles;
    int nrecv;
    int N, T, c, hshift, b0;
    int no_diag;            // 1: drop the diagonal (u = -g * flip sum), used for dH/dg
    int use_tma;            // 1: stage contiguous tiles with cp.async.bulk + mbarrier (pipelined kernel only)
};

__device__ __forceinline__ double tfim_diag_dev(uint64_t s, int N, uint64_t mask) {
    const uint64_t rot = ((s << 1) | (s >> (N - 1))) & mask;
    return -(double)(N - 2 * __popcll(s ^ rot));
}

constexpr int kSweepThreads = 512;
enum { MODE_FIRST = 0, MODE_ACCUM = 1, MODE_ADJ = 2 };

template <int MODE>
__global__ void __launch_bounds__(kSweepThreads, 2) tfim_sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(16) double tile[];
    __shared__ double red[32];
    if (p.guard && *p.guard != 0.0) return;
    const int T = p.T, c = p.c;
    const uint32_t cmask = (1u << c) - 1u;
    const int half = 1 << (T - 1);
    const int midbits = p.hshift - c;
    const uint64_t nmask = (p.N >= 64) ? ~0ull : ((1ull << p.N) - 1ull);
    const double g = (MODE == MODE_ADJ || p.g == nullptr) ? 1.0 : *p.g;
    const double shift = (MODE == MODE_FIRST && p.shift) ? *p.shift : 0.0;
    const double rscale = p.remote_scale ? *p.remote_scale : 1.0;
    double part = 0.0;

    for (uint64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const uint64_t t_mid = t & ((1ull << midbits) - 1ull), t_up = t >> midbits;
        const uint64_t base = (t_mid << c) | (t_up << (p.hshift + T - c));
        // ---- stage the tile (coalesced 16 B loads; runs of 2^c doubles) ----
#pragma unroll 4
        for (int e2 = threadIdx.x; e2 < half; e2 += kSweepThreads) {
            const uint32_t e = 2u * e2;
            const uint64_t gi = base | (e & cmask) | ((uint64_t)(e >> c) << p.hshift);
            *reinterpret_cast<double2*>(&tile[e]) = ldg2(p.v + gi);
        }
        __syncthreads();
        // ---- neighbour sums from shared memory ----
#pragma unroll 2
        for (int e2 = threadIdx.x; e2 < half; e2 += kSweepThreads) {
            const uint32_t e = 2u * e2;
            const uint64_t gi = base | (e & cmask) | ((uint64_t)(e >> c) << p.hshift);
            const double2 x = *reinterpret_cast<const double2*>(&tile[e]);
            double a0 = 0.0, a1 = 0.0;
            int b = p.b0;
            if (b == 0) { a0 = x.y; a1 = x.x; b = 1; }        // bit 0: the pair partner
            for (; b < T; ++b) {
                const double2 y = *reinterpret_cast<const double2*>(&tile[e ^ (1u << b)]);
                a0 += y.x;
                a1 += y.y;
            }
            if (p.nrecv > 0) {                                  // top (remote) spin bits
                double r0 = 0.0, r1 = 0.0;
                for (int j = 0; j < p.nrecv; ++j) {
                    const double2 y = ldg2(p.recv + (uint64_t)j * p.recv_stride + gi);
                    r0 += y.x;
                    r1 += y.y;
                }
                a0 += rscale * r0;
                a1 += rscale * r1;
            }
            if (MODE == MODE_ADJ) {
                const double2 wv = ldg2(p.w + gi);
                part -= wv.x * a0 + wv.y * a1;
            } else {
                double2 o;
                if (MODE == MODE_FIRST) {
                    const uint64_t s = p.rank_off | gi;
                    const double d0 = p.no_diag ? 0.0 : tfim_diag_dev(s, p.N, nmask);
                    const double d1 = p.no_diag ? 0.0 : tfim_diag_dev(s | 1ull, p.N, nmask);
                    o.x = (d0 - shift) * x.x - g * a0;
                    o.y = (d1 - shift) * x.y - g * a1;
                } else {
                    const double2 ui = ldg2(p.uin + gi);
                    o.x = ui.x - g * a0;
                    o.y = ui.y - g * a1;
                }
                stg2(p.uout + gi, o);
                if (p.w) {
                    const double2 wv = (p.w == p.v) ? x : ldg2(p.w + gi);
                    part += wv.x * o.x + wv.y * o.y;
                }
            }
        }
        __syncthreads();
    }
    if (p.partials) {
        const double tot = block_sum(part, red);
        if (threadIdx.x == 0) p.partials[blockIdx.x] = tot;
    }
}

// ---- fast path: persistent, double-buffered, register-blocked sweep for full 2^13 tiles -------------
// One CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Tile i+1 is fetched with cp.async
// (LDGSTS, 16 B) into the second 64 KB buffer while tile i is being reduced, so the HBM stream never
// waits for the shared-memory phase.  A thread owns 8 pairs = 16 amplitudes whose tile bits {0,10,11,12}
// vary: flips of those four bits are register moves; only tile bits 1..9 are served by shared memory
// (9 conflict-free LDS.128 per pair instead of 12), which brings the crossbar traffic of a tile below
// its HBM time.
constexpr int kPipeT = 13;
constexpr int kPipeThreads = 512;
constexpr int kPipePairs = (1 << (kPipeT - 1)) / kPipeThreads;   // 8

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: used for CONTIGUOUS tiles, where one
// elected thread moves the whole 64 KB tile with four instructions instead of 4096 LDGSTS.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kPipeThreads, 1) tfim_sweep_pipe_kernel(const SweepParams p) {
    extern __shared__ __align__(128) double bufs[];              // 2 x 2^13 doubles
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar[2];
    if (p.guard && *p.guard != 0.0) return;
    constexpr int T = kPipeT;
    const bool tma = p.use_tma != 0;                             // contiguous tiles only (first sweep)
    uint32_t phase[2] = {0u, 0u};
    if (tma) {
        if (threadIdx.x == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const int c = p.c;
    const uint32_t cmask = (1u << c) - 1u;
    const int midbits = p.hshift - c;
    const uint64_t nmask = (p.N >= 64) ? ~0ull : ((1ull << p.N) - 1ull);
    const double g = (MODE == MODE_ADJ || p.g == nullptr) ? 1.0 : *p.g;
    const double shift = (MODE == MODE_FIRST && p.shift) ? *p.shift : 0.0;
    const double rscale = p.remote_scale ? *p.remote_scale : 1.0;
    const int bstart = p.b0 > 1 ? p.b0 : 1;
    double part = 0.0;

    uint32_t e[kPipePairs];                                      // tile offsets of this thread's pairs
#pragma unroll
    for (int j = 0; j < kPipePairs; ++j) e[j] = 2u * (threadIdx.x + kPipeThreads * j);

    auto tile_base = [&](uint64_t t) -> uint64_t {
        const uint64_t t_mid = t & ((1ull << midbits) - 1ull), t_up = t >> midbits;
        return (t_mid << c) | (t_up << (p.hshift + T - c));
    };
    auto gidx = [&](uint64_t base, uint32_t ee) -> uint64_t {
        return base | (ee & cmask) | ((uint64_t)(ee >> c) << p.hshift);
    };
    auto prefetch = [&](uint64_t t, double* buf, int st) {
        const uint64_t base = tile_base(t);
        if (tma) {
            if (threadIdx.x == 0) {
                constexpr uint32_t kChunk = (uint32_t)(sizeof(double) << T) / 4;           // 16 KB
                mbar_expect_tx(&mbar[st], 4 * kChunk);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    tma_bulk_g2s(buf + q * (kChunk / 8), p.v + base + q * (kChunk / 8), kChunk, &mbar[st]);
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < kPipePairs; ++j) cp_async16(buf + e[j], p.v + gidx(base, e[j]));
        cp_async_commit();
    };

    uint64_t t = blockIdx.x;
    int stage = 0;
    if (t < p.ntiles) prefetch(t, bufs, 0);
    for (; t < p.ntiles; t += gridDim.x, stage ^= 1) {
        const double* buf = bufs + ((size_t)stage << T);
        const uint64_t tn = t + gridDim.x;
        const uint64_t base = tile_base(t);
        if (tn < p.ntiles) prefetch(tn, bufs + ((size_t)(stage ^ 1) << T), stage ^ 1);
        double2 ui[kPipePairs];
        if (MODE == MODE_ACCUM) {                                 // issue these HBM loads before waiting
#pragma unroll
            for (int j = 0; j < kPipePairs; ++j) ui[j] = ldg2(p.uin + gidx(base, e[j]));
        }
        if (tma) {
            mbar_wait(&mbar[stage], phase[stage]);                // the bulk copy's bytes have landed
            phase[stage] ^= 1u;
        } else {
            if (tn < p.ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
        }

        double2 x[kPipePairs], a[kPipePairs];
#pragma unroll
        for (int j = 0; j < kPipePairs; ++j) x[j] = *reinterpret_cast<const double2*>(buf + e[j]);
#pragma unroll
        for (int j = 0; j < kPipePairs; ++j) {                    // register-resident flips
            a[j] = (p.b0 == 0) ? make_double2(x[j].y, x[j].x) : make_double2(0.0, 0.0);   // tile bit 0
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) {                      // tile bits 10, 11, 12
                a[j].x += x[j ^ (1 << jb)].x;
                a[j].y += x[j ^ (1 << jb)].y;
            }
        }
        for (int b = bstart; b < 10; ++b) {                       // tile bits 1..9 from shared memory
#pragma unroll
            for (int j = 0; j < kPipePairs; ++j) {
                const double2 y = *reinterpret_cast<const double2*>(buf + (e[j] ^ (1u << b)));
                a[j].x += y.x;
                a[j].y += y.y;
            }
        }
#pragma unroll
        for (int j = 0; j < kPipePairs; ++j) {
            const uint64_t gi = gidx(base, e[j]);
            double a0 = a[j].x, a1 = a[j].y;
            if (p.nrecv > 0) {                                    // top (remote) spin bits
                double r0 = 0.0, r1 = 0.0;
                for (int q = 0; q < p.nrecv; ++q) {
                    const double2 y = ldg2(p.recv + (uint64_t)q * p.recv_stride + gi);
                    r0 += y.x;
                    r1 += y.y;
                }
                a0 += rscale * r0;
                a1 += rscale * r1;
            }
            if (MODE == MODE_ADJ) {
                const double2 wv = ldg2(p.w + gi);
                part -= wv.x * a0 + wv.y * a1;
            } else {
                double2 o;
                if (MODE == MODE_FIRST) {
                    const uint64_t s = p.rank_off | gi;
                    const double d0 = p.no_diag ? 0.0 : tfim_diag_dev(s, p.N, nmask);
                    const double d1 = p.no_diag ? 0.0 : tfim_diag_dev(s | 1ull, p.N, nmask);
                    o.x = (d0 - shift) * x[j].x - g * a0;
                    o.y = (d1 - shift) * x[j].y - g * a1;
                } else {
                    o.x = ui[j].x - g * a0;
                    o.y = ui[j].y - g * a1;
                }
                stg2(p.uout + gi, o);
                if (p.w) {
                    const double2 wv = (p.w == p.v) ? x[j] : ldg2(p.w + gi);
                    part += wv.x * o.x + wv.y * o.y;
                }
            }
        }
        __syncthreads();          // everyone is done with `buf` before the next prefetch overwrites it
    }
    if (p.partials) {
        const double tot = block_sum(part, red);
        if (threadIdx.x == 0) p.partials[blockIdx.x] = tot;
    }
}

// Sweep schedule for L local bits with tiles of at most Tmax bits.
static int plan_sweeps(int L, int Tmax, int run_bits, Sweep* out) {
    int n = 0;
    const int T1 = L < Tmax ? L : Tmax;
    out[n++] = Sweep{T1, T1, T1, 0};
    int rem = L - T1, pos = T1;
    if (rem > 0) {
        int c = run_bits;
        if (c <= 0) {
            // fewest sweeps subject to c >= 2 (32 B sectors); then the longest runs that keep that count
            const int nmin = (rem + (Tmax - 2) - 1) / (Tmax - 2);
            c = 2;
            for (int cc = 3; cc <= 8 && cc < Tmax; ++cc)
                if ((rem + (Tmax - cc) - 1) / (Tmax - cc) == nmin) c = cc;
        }
        if (c > T1) c = T1;
        if (c > Tmax - 1) c = Tmax - 1;      // every strided sweep must handle at least one new bit
        if (c < 1) c = 1;
        int nsw = (rem + (Tmax - c) - 1) / (Tmax - c);
        while (rem > 0 && n < kMaxSweeps) {
            const int h = (rem + nsw - 1) / nsw;
            out[n++] = Sweep{c + h, c, pos, c};
            pos += h;
            rem -= h;
            --nsw;
        }
        if (rem > 0) return -1;
    }
    return n;
}

template <int MODE>
static int launch_sweep(dsea_ctx* ctx, const SweepParams& p, int grid, bool pipe, cudaStream_t st) {
    if (pipe) {
        const size_t smem2 = (sizeof(double) << kPipeT) * 2;
        static bool pipe_attr_done[3] = {false, false, false};
        if (!pipe_attr_done[MODE]) {
            DSEA_CUDA(cudaFuncSetAttribute(tfim_sweep_pipe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem2));
            pipe_attr_done[MODE] = true;
        }
        tfim_sweep_pipe_kernel<MODE><<<grid, kPipeThreads, smem2, st>>>(p);
        count_launch(ctx);
        DSEA_CUDA(cudaGetLastError());
        return DSEA_OK;
    }
    const size_t smem = sizeof(double) << p.T;
    static bool attr_done[3] = {false, false, false};
    if (!attr_done[MODE]) {
        DSEA_CUDA(cudaFuncSetAttribute(tfim_sweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(double) << 14)));
        attr_done[MODE] = true;
    }
    tfim_sweep_kernel<MODE><<<grid, kSweepThreads, smem, st>>>(p);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// Shared driver: mode_adj=false -> u = H v (- shift v), optional dot with `dotw`;
//                mode_adj=true  -> out = -sum_s w[s] * sum_i v[s^(1<<i)].
static int tfim_run(dsea_ctx* ctx, const dsea_op* op, bool mode_adj, const double* g, const double* shift,
                    const double* v, double* u, const double* w, double* dot_out, double* work,
                    cudaStream_t st, bool prepushed = false, const double* remote_scale = nullptr) {
    Sweep sw[kMaxSweeps];
    const int L = op->L;
    int Tmax = ctx->tfim_tile_bits;
    if (Tmax > 14) Tmax = 14;
    if (Tmax < 3) Tmax = 3;
    const int ns = plan_sweeps(L, Tmax, ctx->tfim_run_bits, sw);
    DSEA_ARG(ns > 0, "TFIM sweep plan failed");
    const int nrecv = ctx->log2world;
    const bool p2p = nrecv > 0 && ctx->p2p_ok && ctx->arena_stride >= (int64_t)op->n_loc;
    DSEA_ARG(nrecv == 0 || p2p || work != nullptr, "sharded TFIM matvec needs a work buffer");
    DSEA_ARG(!prepushed || p2p, "prepushed input without a peer arena");

    if (p2p) {         // partners' shards arrive in the arena by peer stores
        if (!prepushed) {
            if (!ctx->fresh_collective) DSEA_TRY(comm_barrier(ctx, st));   // partners finished reading the arena
            DSEA_TRY(push_to_peers(ctx, v, op->n_loc, st));
            DSEA_TRY(comm_barrier(ctx, st));                               // partners' stores have landed
        }
    } else if (nrecv > 0) {   // top-bit shards travel on the side stream while the local sweeps run
        DSEA_CUDA(cudaEventRecord(ctx->ev_ready, st));
        DSEA_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
        DSEA_TRY(exchange_shards(ctx, v, work, op->n_loc, ctx->comm_stream));
        DSEA_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    }
    const bool want_dot = mode_adj || (dot_out != nullptr);
    int total_partials = 0;
    const int tok = prof_begin(ctx, mode_adj ? PK_ADJOINT : PK_MATVEC, 16.0 * (double)op->n_loc, st);
    for (int j = 0; j < ns; ++j) {
        const bool last = (j == ns - 1);
        SweepParams p;
        p.v = v;
        p.uin = u;
        p.uout = u;
        p.g = g;
        p.shift = shift;
        p.recv = p2p ? ctx->arena : work;
        p.recv_stride = p2p ? (uint64_t)ctx->arena_stride : (uint64_t)op->n_loc;
        p.remote_scale = (p2p && prepushed) ? remote_scale : nullptr;
        p.guard = ctx->guard;
        p.no_diag = (!mode_adj && g == nullptr) ? 1 : 0;
        p.use_tma = (ctx->tfim_tma && sw[j].c == sw[j].T && (((uintptr_t)v) & 127u) == 0) ? 1 : 0;
        p.nrecv = last ? nrecv : 0;
        p.rank_off = (uint64_t)ctx->rank << L;
        p.n_loc = (uint64_t)op->n_loc;
        p.N = op->N;
        p.T = sw[j].T;
        p.c = sw[j].c;
        p.hshift = sw[j].hshift;
        p.b0 = sw[j].b0;
        p.ntiles = 1ull << (L - sw[j].T);
        // full 2^13 tiles take the persistent double-buffered kernel (one CTA per SM)
        // (the adjoint reduction keeps the 2-CTA/SM generic kernel: it has no output stream to overlap)
        // and so does a sweep that adds remote shards: its extra HBM reads sit in the epilogue, where the
        // 2-CTA/SM kernel hides their latency better — measured 81.7 vs 99.3 ms per solve at P=2)
        const bool pipe = ctx->tfim_pipeline && !mode_adj && p.nrecv == 0 && p.T == kPipeT && (p.b0 == 0 || p.c <= 10) &&
                          p.ntiles >= 2;
        int grid = pipe ? (int)(p.ntiles < (uint64_t)ctx->num_sms ? p.ntiles : (uint64_t)ctx->num_sms)
                        : (int)(p.ntiles < 2048 ? p.ntiles : 2048);
        if (last && nrecv > 0 && !p2p) DSEA_CUDA(cudaStreamWaitEvent(st, ctx->ev_comm, 0));
        if (mode_adj) {
            p.w = w;
            p.partials = ctx->partials + total_partials;   // every sweep contributes partial sums
            total_partials += grid;
            DSEA_TRY(launch_sweep<MODE_ADJ>(ctx, p, grid, pipe, st));
        } else {
            p.w = (last && want_dot) ? w : nullptr;
            p.partials = (last && want_dot) ? ctx->partials : nullptr;
            if (last && want_dot) total_partials = grid;
            if (j == 0) DSEA_TRY(launch_sweep<MODE_FIRST>(ctx, p, grid, pipe, st));
            else DSEA_TRY(launch_sweep<MODE_ACCUM>(ctx, p, grid, pipe, st));
        }
    }
    prof_end(ctx, tok, st);
    if (p2p) ctx->fresh_collective = false;    // the arena was just read: the next push needs a collective first
    if (want_dot) {
        DSEA_TRY(finalize_reduce(ctx, total_partials, 1, dot_out, st));
    }
    return DSEA_OK;
}

int tfim_apply(dsea_ctx* ctx, const dsea_op* op, const double* g, const double* shift, const double* v,
               double* u, const double* dotw, double* dot_out, double* work, cudaStream_t st, bool prepushed,
               const double* remote_scale) {
    return tfim_run(ctx, op, false, g, shift, v, u, dotw ? dotw : v, dot_out, work, st, prepushed, remote_scale);
}

// u = (dH/dg) v: the same sweeps with g = 1 and the diagonal dropped (g == nullptr selects this).
int tfim_dHdg(dsea_ctx* ctx, const dsea_op* op, const double* v, double* u, double* work, cudaStream_t st) {
    return tfim_run(ctx, op, false, nullptr, nullptr, v, u, v, nullptr, work, st);
}

int tfim_adjoint(dsea_ctx* ctx, const dsea_op* op, const double* v1, const double* v2, double* out,
                 double* work, cudaStream_t st) {
    return tfim_run(ctx, op, true, nullptr, nullptr, v2, nullptr, v1, out, work, st);
}

}  // namespace dsea

extern "C" int64_t dsea_tfim_flip_index(int N, int64_t s, int i) {
    (void)N;
    return s ^ ((int64_t)1 << i);
}

extern "C" double dsea_tfim_diag(int N, int64_t s) {
    const uint64_t mask = (N >= 64) ? ~0ull : ((1ull << N) - 1ull);
    const uint64_t us = (uint64_t)s;
    const uint64_t rot = ((us << 1) | (us >> (N - 1))) & mask;
    return -(double)(N - 2 * __builtin_popcountll(us ^ rot));
}

// Host-callable view of the sweep schedule (for the CPU tests of the index logic): writes up to 40 rows
// {T, c, hshift, b0} and returns the number of sweeps (negative on failure).
extern "C" int dsea_tfim_plan(int local_bits, int tile_bits, int run_bits, int* out4x40) {
    dsea::Sweep sw[dsea::kMaxSweeps];
    if (tile_bits > 14) tile_bits = 14;
    if (tile_bits < 3) tile_bits = 3;
    const int n = dsea::plan_sweeps(local_bits, tile_bits, run_bits, sw);
    for (int j = 0; j < n && j < dsea::kMaxSweeps; ++j) {
        out4x40[4 * j + 0] = sw[j].T;
        out4x40[4 * j + 1] = sw[j].c;
        out4x40[4 * j + 2] = sw[j].hshift;
        out4x40[4 * j + 3] = sw[j].b0;
    }
    return n;
}
