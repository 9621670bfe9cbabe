// tfim.cu — matrix-free transverse-field Ising operator on the 2^N spin basis (K1, K6).
//
//   u[s] = diag(s) v[s] - g * sum_{i<N} v[s ^ (1<<i)]  (- shift * v[s])
//   diag(s) = -(N - 2 popc(s ^ rotl_N(s)))             (bit-exact restatement of TFIM.py:39-46)
//   flip index s ^ (1<<i)                              (TFIM.py:48-51)
//
// No index table exists: the reference's (2^N, N) int64 table (60 GB at N=28) is replaced by bit
// arithmetic on the global index (rank << L) | s_loc.
//
// Memory plan.  A CTA stages a TILE of 2^T doubles in shared memory and serves every flip whose bit lies
// inside the tile from there, so each vector element is read from HBM once per SWEEP:
//   sweep 0 : tile = 2^T contiguous doubles            -> handles spin bits [0, T)
//   sweep j : tile = 2^c contiguous x 2^h strided runs -> handles h spin bits starting at `hshift`
//             (run stride 2^hshift doubles).  Runs are at least 2^4 doubles = one 128-byte line, so a warp's
//             16-byte accesses touch 4 lines per instruction exactly like a contiguous stream (round 1 used
//             32-byte runs: 16 lines per instruction, which made the sweep L1TEX-wavefront bound).
//   direct  : with 13-bit tiles two sweeps reach 13 + 9 = 22 local bits.  Up to kMaxDirect further (top)
//             local bits are NOT given a third sweep (24 more bytes per element of HBM traffic); the last
//             sweep reads the partner elements v[s ^ (1<<b)] straight from global memory.  The tile order
//             puts the 2^ndirect tiles that are each other's partners on neighbouring CTAs in the same
//             iteration, so those reads are served by L2 (each tile is fetched from HBM once, by its owner).
//   top log2(world) bits : the shard of rank ^ (1<<j) sits in slot j of this rank's peer arena, stored
//             there over NVLink by the kernel that produced it (reorth pass 2, the CG direction update, or
//             push_kernel); the last sweep adds the slots.  Fallback: grouped ncclSend/ncclRecv on a side
//             stream, overlapped with the local sweeps.
// HBM bytes per element: 16 (first sweep: read v, write u) + 24 per extra sweep (read v, read u, write u)
// + 8 per remote bit.  The dot-product epilogue (v.u for CG, or w.u) rides on the last sweep.
// Fused normalisation (K3): the first sweep can take its input as in_scale * v and write the scaled vector
// back (q = r / beta), which removes the separate 16 B/element normalisation pass of every Lanczos step.
#include "common.cuh"

namespace dsea {

constexpr int kMaxDirect = 4;   // measured: from 5 direct bits on, a third sweep is faster (L = 27: 1.70 vs 1.76 ms)

struct Sweep {
    int T;        // tile bits
    int c;        // contiguous low bits of the tile
    int hshift;   // global bit position of tile bit c
    int b0;       // first tile bit this sweep is responsible for
    int ndirect;  // local bits above the tile's strided range served by direct global loads
};

struct SweepParams {
    const double* v;
    const double* uin;
    double* uout;
    const double* w;        // dot partner (may alias v); nullptr = no dot
    const double* g;
    const double* shift;
    const double* in_scale; // MODE_FIRST: the logical input is (*in_scale) * v   (nullptr = 1)
    double* q_out;          // MODE_FIRST: if non-null the scaled input is written here (may alias v)
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
    int upbits;             // address bits above the tile's strided range: L - (hshift + T - c)
    int ndirect;            // 0 or upbits
    int fast_up;            // 1: the tile index's LOW bits select the upper address bits (partner tiles adjacent)
    int no_diag;            // 1: drop the diagonal (u = -g * flip sum), used for dH/dg
    int use_tma;            // 1: stage contiguous tiles with cp.async.bulk + mbarrier (pipelined kernel only)
    int l2_prefetch;        // 1: prefetch.global.L2 the next tile's epilogue operands
    int round_remote;       // 1: remote terms are fl32(remote_scale * y): the partner's fp32-rounded Lanczos vector
    int pdl_trigger;        // staged kernel: 1 = let the next kernel's CTAs be scheduled while this one runs
};

// diag(s) = -(N - 2 popc(s ^ rotl_N(s))), an integer in [-N, N].  POPC and I2F run on the quarter-rate pipe (16 per clock
// per SM), and the 64-bit forms cost two POPC each: for N <= 32 the bit arithmetic is done in 32 bits, and the integer is
// turned into a double with the 2^52 trick (one LOP3 + one DADD on the fp64 pipe) instead of I2F.F64 — bit-identical.
__device__ __forceinline__ double small_int_to_double(int k) {
    return __hiloint2double(0x43300000, (int)((unsigned)k ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}
__device__ __forceinline__ double tfim_diag_dev(uint64_t s, int N, uint64_t mask) {
    if (N <= 32) {
        const uint32_t s32 = (uint32_t)s, m32 = (uint32_t)mask;
        const uint32_t rot = ((s32 << 1) | (s32 >> (N - 1))) & m32;
        return small_int_to_double(2 * __popc(s32 ^ rot) - N);
    }
    const uint64_t rot = ((s << 1) | (s >> (N - 1))) & mask;
    return small_int_to_double(2 * __popcll(s ^ rot) - N);
}

// tile index -> offset of the tile's element 0
__device__ __forceinline__ uint64_t tile_base(const SweepParams& p, uint64_t t) {
    const int midbits = p.hshift - p.c;
    uint64_t t_mid, t_up;
    if (p.fast_up) {
        t_up = t & ((1ull << p.upbits) - 1ull);
        t_mid = t >> p.upbits;
    } else {
        t_mid = t & ((1ull << midbits) - 1ull);
        t_up = t >> midbits;
    }
    return (t_mid << p.c) | (t_up << (p.hshift + p.T - p.c));
}

constexpr int kSweepThreads = 512;
enum { MODE_FIRST = 0, MODE_ACCUM = 1, MODE_ADJ = 2 };

// ---- generic sweep: any tile size / plan; 2 CTAs per SM, one pair per thread at a time ---------------------
template <int MODE>
__global__ void __launch_bounds__(kSweepThreads, 2) tfim_sweep_kernel(const SweepParams p) {
    pdl_prologue();
    extern __shared__ __align__(16) double tile[];
    __shared__ double red[32];
    if (p.guard && *p.guard != 0.0) return;
    const int T = p.T, c = p.c;
    const uint32_t cmask = (1u << c) - 1u;
    const int half = 1 << (T - 1);
    const int dpos = p.hshift + T - c;
    const uint64_t nmask = (p.N >= 64) ? ~0ull : ((1ull << p.N) - 1ull);
    const double g = (MODE == MODE_ADJ || p.g == nullptr) ? 1.0 : *p.g;
    const double shift = (MODE == MODE_FIRST && p.shift) ? *p.shift : 0.0;
    const double rscale = p.remote_scale ? *p.remote_scale : 1.0;
    double part = 0.0;

    for (uint64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const uint64_t base = tile_base(p, t);
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
            for (int d = 0; d < p.ndirect; ++d) {               // top local bits: partner tiles, through L2
                const double2 y = ldg2(p.v + (gi ^ (1ull << (dpos + d))));
                a0 += y.x;
                a1 += y.y;
            }
            if (p.nrecv > 0) {                                  // top (remote) spin bits
                double r0 = 0.0, r1 = 0.0;
                for (int j = 0; j < p.nrecv; ++j) {
                    const double2 y = ldg2(p.recv + (uint64_t)j * p.recv_stride + gi);
                    if (p.round_remote) {
                        a0 += (double)(float)(rscale * y.x);
                        a1 += (double)(float)(rscale * y.y);
                    } else {
                        r0 += y.x;
                        r1 += y.y;
                    }
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
// One CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Tile i+1 is fetched asynchronously into
// the second 64 KB buffer while tile i is being reduced: contiguous tiles by TMA bulk copies (UBLKCP, one
// elected thread, completion on an mbarrier), strided tiles by cp.async (LDGSTS, 16 B, runs >= 128 B).
// A thread owns PAIRS = 4096 / THREADS pairs whose tile bit 0 and top log2(PAIRS) tile bits vary: flips of
// those bits are register moves; only the remaining tile bits are served by conflict-free LDS.128.
constexpr int kPipeT = 13;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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

// LB: compile-time first shared-memory flip bit (1: contiguous first sweep; 4: strided sweep with 128-byte runs;
// 0: taken from the plan at run time).  With LB known the flip loop is fully unrolled, so the LDS of the next bit
// are in flight while the adds of the current one retire.
template <int MODE, int THREADS, int LB>
__global__ void __launch_bounds__(THREADS, 1) tfim_sweep_pipe_kernel(const SweepParams p) {
    pdl_prologue();
    constexpr int T = kPipeT;
    constexpr int PAIRS = (1 << (T - 1)) / THREADS;            // 16 / 8 / 4 pairs per thread (256 / 512 / 1024 threads)
    constexpr int RBITS = (PAIRS == 16) ? 4 : (PAIRS == 8 ? 3 : 2);   // register-resident top tile bits
    constexpr int RB0 = T - RBITS;                             // first of them
    extern __shared__ __align__(128) double bufs[];            // 2 x 2^13 doubles
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar[2];
    if (p.guard && *p.guard != 0.0) return;
    const bool tma = p.use_tma != 0;                           // contiguous tiles only (first sweep)
    uint32_t phase0 = 0u, phase1 = 0u;
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
    const int dpos = p.hshift + T - c;
    const uint64_t nmask = (p.N >= 64) ? ~0ull : ((1ull << p.N) - 1ull);
    const double g = (MODE == MODE_ADJ || p.g == nullptr) ? 1.0 : *p.g;
    const double shift = (MODE == MODE_FIRST && p.shift) ? *p.shift : 0.0;
    const double scl = (MODE == MODE_FIRST && p.in_scale) ? *p.in_scale : 1.0;
    const double rscale = p.remote_scale ? *p.remote_scale : 1.0;
    const int bstart = LB > 0 ? LB : (p.b0 > 1 ? p.b0 : 1);
    const bool fold_diag = (MODE == MODE_FIRST) && !p.no_diag && g != 0.0;
    const double inv_g = fold_diag ? 1.0 / g : 0.0;
    double part = 0.0;

    auto eoff = [&](int j) -> uint32_t { return 2u * (threadIdx.x + THREADS * j); };   // tile offset of pair j
    auto gidx = [&](uint64_t base, uint32_t ee) -> uint64_t {
        return base | (ee & cmask) | ((uint64_t)(ee >> c) << p.hshift);
    };
    auto prefetch = [&](uint64_t t, double* buf, int st) {
        const uint64_t base = tile_base(p, t);
        if (tma) {
            if (threadIdx.x == 0) {
                constexpr uint32_t kChunk = (uint32_t)(sizeof(double) << T) / 4;           // 16 KB
                mbar_expect_tx(&mbar[st], 4 * kChunk);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    tma_bulk_g2s(buf + q * (kChunk / 8), p.v + base + q * (kChunk / 8), kChunk, &mbar[st]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) cp_async16(buf + eoff(j), p.v + gidx(base, eoff(j)));
            cp_async_commit();
        }
        if (p.l2_prefetch && c >= 4) {                          // the next tile's epilogue operands, one 128 B line each
            for (uint32_t l = threadIdx.x; l < (1u << (T - 4)); l += THREADS) {
                const uint64_t gl = gidx(base, 16u * l);
                if (MODE == MODE_ACCUM) prefetch_l2(p.uin + gl);
                for (int d = 0; d < p.ndirect; ++d) prefetch_l2(p.v + (gl ^ (1ull << (dpos + d))));
                for (int q = 0; q < p.nrecv; ++q) prefetch_l2(p.recv + (uint64_t)q * p.recv_stride + gl);
                if (p.w && p.w != p.v) prefetch_l2(p.w + gl);
            }
        }
    };

    uint64_t t = blockIdx.x;
    int stage = 0;
    if (t < p.ntiles) prefetch(t, bufs, 0);
    for (; t < p.ntiles; t += gridDim.x, stage ^= 1) {
        const double* buf = bufs + ((size_t)stage << T);
        const uint64_t tn = t + gridDim.x;
        const uint64_t base = tile_base(p, t);
        if (tn < p.ntiles) prefetch(tn, bufs + ((size_t)(stage ^ 1) << T), stage ^ 1);
        if (tma) {
            if (stage == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1u; }    // the bulk copy's bytes have landed
            else { mbar_wait(&mbar[1], phase1); phase1 ^= 1u; }
        } else {
            if (tn < p.ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
        }

        double2 a[PAIRS];
        {
            double2 x[PAIRS];
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) x[j] = *reinterpret_cast<const double2*>(buf + eoff(j));
#pragma unroll
            for (int j = 0; j < PAIRS; ++j)                       // tile bit 0: the pair partner
                a[j] = (p.b0 == 0) ? make_double2(x[j].y, x[j].x) : make_double2(0.0, 0.0);
#pragma unroll
            for (int jb = 0; jb < RBITS; ++jb) {                  // register-resident flips: tile bits RB0 .. 12
                if (LB > 0 || RB0 + jb >= p.b0) {
#pragma unroll
                    for (int j = 0; j < PAIRS; ++j) {
                        a[j].x += x[j ^ (1 << jb)].x;
                        a[j].y += x[j ^ (1 << jb)].y;
                    }
                }
            }
            if (MODE == MODE_FIRST && fold_diag) {
                // diagonal term folded into the accumulators while x is still in registers: its integer / popc /
                // conversion work then runs in the shadow of the LDS phase instead of after it, and the epilogue
                // shrinks to o = -g * scale * a.   u = (d - shift) x - g S  =  -g (S - ((d - shift) / g) x)
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    const uint64_t s = p.rank_off | gidx(base, eoff(j));
                    const double d0 = tfim_diag_dev(s, p.N, nmask), d1 = tfim_diag_dev(s | 1ull, p.N, nmask);
                    a[j].x -= ((d0 - shift) * inv_g) * x[j].x;
                    a[j].y -= ((d1 - shift) * inv_g) * x[j].y;
                }
            }
        }                                                         // x is dead here: re-read from the tile when needed
        if (LB > 0) {
#pragma unroll
            for (int b = LB; b < RB0; ++b) {                      // the other tile bits from shared memory
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    const double2 y = *reinterpret_cast<const double2*>(buf + (eoff(j) ^ (1u << b)));
                    a[j].x += y.x;
                    a[j].y += y.y;
                }
            }
        } else {
            for (int b = bstart; b < RB0; ++b) {
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    const double2 y = *reinterpret_cast<const double2*>(buf + (eoff(j) ^ (1u << b)));
                    a[j].x += y.x;
                    a[j].y += y.y;
                }
            }
        }
        // ---- epilogue: operands that live in global memory (L2-resident partner tiles, u, w, arena slots) ----
        for (int d = 0; d < p.ndirect; ++d) {                     // top local bits: partner tiles
            const uint64_t dbit = 1ull << (dpos + d);
            double2 y[PAIRS];
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) y[j] = ldg2(p.v + (gidx(base, eoff(j)) ^ dbit));
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) {
                a[j].x += y[j].x;
                a[j].y += y[j].y;
            }
        }
        for (int q = 0; q < p.nrecv; ++q) {                       // top (remote) spin bits
            const double* slot = p.recv + (uint64_t)q * p.recv_stride;
            double2 y[PAIRS];
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) y[j] = ldg2(slot + gidx(base, eoff(j)));
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) {
                const double t0 = rscale * y[j].x, t1 = rscale * y[j].y;
                a[j].x += p.round_remote ? (double)(float)t0 : t0;
                a[j].y += p.round_remote ? (double)(float)t1 : t1;
            }
        }
        if (MODE == MODE_ADJ) {
            double2 wv[PAIRS];
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) wv[j] = ldg2(p.w + gidx(base, eoff(j)));
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) part -= wv[j].x * a[j].x + wv[j].y * a[j].y;
        } else {
            const bool dot_self = p.w != nullptr && p.w == p.v;
            double2 o[PAIRS];
            if (MODE == MODE_FIRST && (fold_diag || p.no_diag)) {
                const double mg = -g * scl;
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    o[j].x = mg * a[j].x;
                    o[j].y = mg * a[j].y;
                }
                if (p.q_out || dot_self) {
#pragma unroll
                    for (int j = 0; j < PAIRS; ++j) {
                        double2 x = *reinterpret_cast<const double2*>(buf + eoff(j));
                        x.x *= scl;                               // the logical input (for q_out and the dot)
                        x.y *= scl;
                        if (p.q_out) stg2(p.q_out + gidx(base, eoff(j)), x);
                        if (dot_self) part += x.x * o[j].x + x.y * o[j].y;
                    }
                }
            } else if (MODE == MODE_FIRST) {                      // g == 0: u = (d - shift) x exactly
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    double2 x = *reinterpret_cast<const double2*>(buf + eoff(j));
                    const uint64_t gi = gidx(base, eoff(j));
                    const uint64_t s = p.rank_off | gi;
                    const double d0 = tfim_diag_dev(s, p.N, nmask);
                    const double d1 = tfim_diag_dev(s | 1ull, p.N, nmask);
                    o[j].x = scl * ((d0 - shift) * x.x - g * a[j].x);
                    o[j].y = scl * ((d1 - shift) * x.y - g * a[j].y);
                    x.x *= scl;
                    x.y *= scl;
                    if (p.q_out) stg2(p.q_out + gi, x);
                    if (dot_self) part += x.x * o[j].x + x.y * o[j].y;
                }
            } else {
                double2 ui[PAIRS];
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) ui[j] = ldg2(p.uin + gidx(base, eoff(j)));
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    o[j].x = ui[j].x - g * a[j].x;
                    o[j].y = ui[j].y - g * a[j].y;
                }
                if (dot_self) {
#pragma unroll
                    for (int j = 0; j < PAIRS; ++j) {
                        const double2 x = *reinterpret_cast<const double2*>(buf + eoff(j));
                        part += x.x * o[j].x + x.y * o[j].y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) stg2(p.uout + gidx(base, eoff(j)), o[j]);
            if (p.w && !dot_self) {
                double2 wv[PAIRS];
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) wv[j] = ldg2(p.w + gidx(base, eoff(j)));
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) part += wv[j].x * o[j].x + wv[j].y * o[j].y;
            }
        }
        __syncthreads();          // everyone is done with `buf` before the next prefetch overwrites it
    }
    if (p.partials) {
        const double tot = block_sum(part, red);
        if (threadIdx.x == 0) p.partials[blockIdx.x] = tot;
    }
}

// ---- staged strided sweep ----------------------------------------------------------------------------------------
// ncu's source page of the pipelined strided sweep shows 43 % of its samples in `stall_long_sb`: after the flip phase the
// epilogue waits, round after round, for operands that live in global memory — the partner tiles of the direct bits, the
// arena slots of the remote bits, and u (or w for the adjoint).  There are no registers to load them early (x and the
// accumulators of 2^13 elements fill the file).  This kernel gives every thread PRIVATE shared-memory slots instead:
// the operands of one SUB-TILE (PAIRS / S of the thread's pairs) are fetched with cp.async while the flip phase of that
// sub-tile runs, then read back by the same thread (no barrier needed), and the slots are refilled for the next
// sub-tile.  Costs +1 LDS.128 per operand and pair (the flip phase stays well below the HBM time of the sweep).
template <int MODE, int S, int LB>
__global__ void __launch_bounds__(512, 1) tfim_sweep_staged_kernel(const SweepParams p) {
    pdl_prologue(p.pdl_trigger != 0);
    constexpr int T = kPipeT, THREADS = 512, PAIRS = 8, PPS = PAIRS / S, RBITS = 3, RB0 = T - RBITS;
    extern __shared__ __align__(128) double bufs[];            // 2 tiles of 2^13 doubles + the staging slots
    __shared__ double red[32];
    if (p.guard && *p.guard != 0.0) return;
    double* stage = bufs + 2 * (1 << T);
    const int c = p.c;
    const uint32_t cmask = (1u << c) - 1u;
    const int dpos = p.hshift + T - c;
    const double g = (MODE == MODE_ADJ || p.g == nullptr) ? 1.0 : *p.g;
    const double rscale = p.remote_scale ? *p.remote_scale : 1.0;
    const int bstart = LB > 0 ? LB : (p.b0 > 1 ? p.b0 : 1);
    const int nops = p.ndirect + p.nrecv + 1;                  // ... + u (ACCUM) or w (ADJ), always last
    const double* last_op = (MODE == MODE_ACCUM) ? p.uin : p.w;
    const bool dot_self = (MODE == MODE_ACCUM) && p.w != nullptr && p.w == p.v;
    double part = 0.0;

    auto eoff = [&](int j) -> uint32_t { return 2u * (threadIdx.x + THREADS * j); };
    auto gidx = [&](uint64_t base, uint32_t ee) -> uint64_t {
        return base | (ee & cmask) | ((uint64_t)(ee >> c) << p.hshift);
    };
    auto slot = [&](int o, int jj) -> double* { return stage + 2 * ((size_t)(o * PPS + jj) * THREADS + threadIdx.x); };
    auto issue_stage = [&](uint64_t base, int sub) {
        for (int o = 0; o < nops; ++o) {
#pragma unroll
            for (int jj = 0; jj < PPS; ++jj) {
                const uint64_t gi = gidx(base, eoff(sub * PPS + jj));
                const double* src;
                if (o < p.ndirect) src = p.v + (gi ^ (1ull << (dpos + o)));
                else if (o < p.ndirect + p.nrecv) src = p.recv + (uint64_t)(o - p.ndirect) * p.recv_stride + gi;
                else src = last_op + gi;
                cp_async16(slot(o, jj), src);
            }
        }
        cp_async_commit();
    };
    auto prefetch_tile = [&](uint64_t t, double* buf) {
        const uint64_t base = tile_base(p, t);
#pragma unroll
        for (int j = 0; j < PAIRS; ++j) cp_async16(buf + eoff(j), p.v + gidx(base, eoff(j)));
        cp_async_commit();
    };

    uint64_t t = blockIdx.x;
    int st = 0;
    if (t < p.ntiles) prefetch_tile(t, bufs);
    for (; t < p.ntiles; t += gridDim.x, st ^= 1) {
        const double* buf = bufs + ((size_t)st << T);
        const uint64_t tn = t + gridDim.x;
        const uint64_t base = tile_base(p, t);
        const bool more = tn < p.ntiles;
        issue_stage(base, 0);                                     // group S0 (slots were consumed last iteration)
        if (more) prefetch_tile(tn, bufs + ((size_t)(st ^ 1) << T));   // group P
        if (more) cp_async_wait<2>(); else cp_async_wait<1>();    // this tile's data is the oldest group
        __syncthreads();

        double2 a[PAIRS];
        {
            double2 x[PAIRS];
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) x[j] = *reinterpret_cast<const double2*>(buf + eoff(j));
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) a[j] = make_double2(0.0, 0.0);       // strided sweeps never own tile bit 0
#pragma unroll
            for (int jb = 0; jb < RBITS; ++jb) {
                if (LB > 0 || RB0 + jb >= p.b0) {
#pragma unroll
                    for (int j = 0; j < PAIRS; ++j) {
                        a[j].x += x[j ^ (1 << jb)].x;
                        a[j].y += x[j ^ (1 << jb)].y;
                    }
                }
            }
        }
#pragma unroll
        for (int sub = 0; sub < S; ++sub) {
            // flips of this sub-tile's pairs from the shared-memory tile (overlaps the staging copies in flight)
            if (LB > 0) {
#pragma unroll
                for (int b = LB; b < RB0; ++b) {
#pragma unroll
                    for (int jj = 0; jj < PPS; ++jj) {
                        const double2 y = *reinterpret_cast<const double2*>(buf + (eoff(sub * PPS + jj) ^ (1u << b)));
                        a[sub * PPS + jj].x += y.x;
                        a[sub * PPS + jj].y += y.y;
                    }
                }
            } else {
                for (int b = bstart; b < RB0; ++b) {
#pragma unroll
                    for (int jj = 0; jj < PPS; ++jj) {
                        const double2 y = *reinterpret_cast<const double2*>(buf + (eoff(sub * PPS + jj) ^ (1u << b)));
                        a[sub * PPS + jj].x += y.x;
                        a[sub * PPS + jj].y += y.y;
                    }
                }
            }
            // the staged operands of this sub-tile have landed (the tile prefetch may still be in flight behind S0)
            if (sub == 0 && more) cp_async_wait<1>(); else cp_async_wait<0>();
            for (int o = 0; o < p.ndirect; ++o) {
#pragma unroll
                for (int jj = 0; jj < PPS; ++jj) {
                    const double2 y = *reinterpret_cast<const double2*>(slot(o, jj));
                    a[sub * PPS + jj].x += y.x;
                    a[sub * PPS + jj].y += y.y;
                }
            }
            for (int o = p.ndirect; o < p.ndirect + p.nrecv; ++o) {
#pragma unroll
                for (int jj = 0; jj < PPS; ++jj) {
                    const double2 y = *reinterpret_cast<const double2*>(slot(o, jj));
                    const double t0 = rscale * y.x, t1 = rscale * y.y;
                    a[sub * PPS + jj].x += p.round_remote ? (double)(float)t0 : t0;
                    a[sub * PPS + jj].y += p.round_remote ? (double)(float)t1 : t1;
                }
            }
#pragma unroll
            for (int jj = 0; jj < PPS; ++jj) {
                const int j = sub * PPS + jj;
                const double2 z = *reinterpret_cast<const double2*>(slot(nops - 1, jj));   // u, or w for the adjoint
                if (MODE == MODE_ADJ) {
                    part -= z.x * a[j].x + z.y * a[j].y;
                } else {
                    const double2 o = make_double2(z.x - g * a[j].x, z.y - g * a[j].y);
                    stg2(p.uout + gidx(base, eoff(j)), o);
                    if (dot_self) {
                        const double2 x = *reinterpret_cast<const double2*>(buf + eoff(j));
                        part += x.x * o.x + x.y * o.y;
                    } else if (p.w) {
                        const double2 wv = ldg2(p.w + gidx(base, eoff(j)));
                        part += wv.x * o.x + wv.y * o.y;
                    }
                }
            }
            if (sub + 1 < S) issue_stage(base, sub + 1);          // refill the slots this thread just consumed
        }
        __syncthreads();          // everyone is done with `buf` before the next prefetch overwrites it
    }
    if (p.partials) {
        const double tot = block_sum(part, red);
        if (threadIdx.x == 0) p.partials[blockIdx.x] = tot;
    }
}

// Sweep schedule for L local bits with tiles of at most Tmax bits.
//   run_bits > 0 : every strided sweep uses runs of exactly 2^run_bits doubles (tests / tuning), no direct bits;
//   run_bits = 0 : automatic.  Runs are at least `cmin` bits long (4 = one 128-byte line for production tiles, 2 for
//                  the tiny tiles the tests use); the fewest sweeps that reach every bit are used, each with the
//                  longest runs that keep that count.  With `allow_direct`, up to kMaxDirect top bits left over
//                  by TWO sweeps become direct loads of the last sweep instead of a third sweep.
static int plan_sweeps(int L, int Tmax, int run_bits, bool allow_direct, Sweep* out) {
    int n = 0;
    const int T1 = L < Tmax ? L : Tmax;
    out[n++] = Sweep{T1, T1, T1, 0, 0};
    int rem = L - T1, pos = T1;
    if (rem <= 0) return n;
    if (run_bits > 0) {
        int c = run_bits;
        if (c > T1) c = T1;
        if (c > Tmax - 1) c = Tmax - 1;      // every strided sweep must handle at least one new bit
        if (c < 1) c = 1;
        int nsw = (rem + (Tmax - c) - 1) / (Tmax - c);
        while (rem > 0 && n < kMaxSweeps) {
            const int h = (rem + nsw - 1) / nsw;
            out[n++] = Sweep{c + h, c, pos, c, 0};
            pos += h;
            rem -= h;
            --nsw;
        }
        return rem > 0 ? -1 : n;
    }
    int cmin = Tmax >= 10 ? 4 : 2;
    if (cmin > Tmax - 1) cmin = Tmax - 1;
    if (cmin < 1) cmin = 1;
    const int hmax = Tmax - cmin;
    if (allow_direct && Tmax >= 10 && rem > hmax && rem - hmax <= kMaxDirect) {
        out[n++] = Sweep{Tmax, cmin, pos, cmin, rem - hmax};
        return n;
    }
    int nsw = (rem + hmax - 1) / hmax;
    while (rem > 0 && n < kMaxSweeps) {
        const int h = (rem + nsw - 1) / nsw;
        int c = Tmax - h;                    // longest runs that still fit h new bits into one tile
        if (c > T1) c = T1;
        out[n++] = Sweep{c + h, c, pos, c, 0};
        pos += h;
        rem -= h;
        --nsw;
    }
    return rem > 0 ? -1 : n;
}

template <int MODE, int THREADS, int LB>
static int launch_pipe(dsea_ctx* ctx, const SweepParams& p, int grid, cudaStream_t st) {
    const size_t smem2 = (sizeof(double) << kPipeT) * 2;
    static bool done = false;
    if (!done) {
        DSEA_CUDA(cudaFuncSetAttribute(tfim_sweep_pipe_kernel<MODE, THREADS, LB>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        done = true;
    }
    launch_k(ctx, tfim_sweep_pipe_kernel<MODE, THREADS, LB>, dim3(grid), dim3(THREADS), smem2, st, p);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

template <int MODE, int S, int LB>
static int launch_staged(dsea_ctx* ctx, const SweepParams& p, int grid, int nops, cudaStream_t st) {
    const size_t smem = (sizeof(double) << kPipeT) * 2 + (size_t)nops * (8 / S) * 512 * 16;
    static size_t set = 0;
    if (smem > set) {
        DSEA_CUDA(cudaFuncSetAttribute(tfim_sweep_staged_kernel<MODE, S, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        set = smem;
    }
    launch_k_opt(ctx, (ctx->pdl_staged & 1) != 0, tfim_sweep_staged_kernel<MODE, S, LB>, dim3(grid), dim3(512), smem, st, p);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

// Sub-tile count of the staged kernel for `nops` global operands per element (0: does not fit in shared memory).
static inline int staged_subtiles(int nops) { return nops <= 3 ? 2 : (nops <= 6 ? 4 : 0); }

template <int MODE>
static int launch_staged_any(dsea_ctx* ctx, const SweepParams& p, int grid, int nops, cudaStream_t st) {
    const bool lb4 = ctx->tfim_unroll && p.b0 == 4 && p.c == 4;
    if (staged_subtiles(nops) == 2) {
        return lb4 ? launch_staged<MODE, 2, 4>(ctx, p, grid, nops, st) : launch_staged<MODE, 2, 0>(ctx, p, grid, nops, st);
    }
    return lb4 ? launch_staged<MODE, 4, 4>(ctx, p, grid, nops, st) : launch_staged<MODE, 4, 0>(ctx, p, grid, nops, st);
}

template <int MODE, int THREADS>
static int launch_pipe_lb(dsea_ctx* ctx, const SweepParams& p, int grid, cudaStream_t st) {
    if (ctx->tfim_unroll && p.b0 == 0) return launch_pipe<MODE, THREADS, 1>(ctx, p, grid, st);
    if (ctx->tfim_unroll && p.b0 == 4 && p.c == 4) return launch_pipe<MODE, THREADS, 4>(ctx, p, grid, st);
    return launch_pipe<MODE, THREADS, 0>(ctx, p, grid, st);
}

template <int MODE>
static int launch_sweep(dsea_ctx* ctx, const SweepParams& p, int grid, bool pipe, cudaStream_t st) {
    if (pipe) {
        if (ctx->tfim_pipe_threads == 256) return launch_pipe_lb<MODE, 256>(ctx, p, grid, st);
        if (ctx->tfim_pipe_threads == 1024) return launch_pipe_lb<MODE, 1024>(ctx, p, grid, st);
        return launch_pipe_lb<MODE, 512>(ctx, p, grid, st);
    }
    const size_t smem = sizeof(double) << p.T;
    static bool attr_done[3] = {false, false, false};
    if (!attr_done[MODE]) {
        DSEA_CUDA(cudaFuncSetAttribute(tfim_sweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(double) << 14)));
        attr_done[MODE] = true;
    }
    launch_k(ctx, tfim_sweep_kernel<MODE>, dim3(grid), dim3(kSweepThreads), smem, st, p);
    count_launch(ctx);
    DSEA_CUDA(cudaGetLastError());
    return DSEA_OK;
}

static inline int clamp_tile_bits(int Tmax) { return Tmax > 14 ? 14 : (Tmax < 3 ? 3 : Tmax); }

static inline bool pipe_eligible(const dsea_ctx* ctx, const Sweep& s, uint64_t ntiles) {
    return ctx->tfim_pipeline && s.T == kPipeT && ntiles >= 2;
}

// True when the FIRST sweep of this operator runs in the pipelined kernel, i.e. tfim_apply can take
// `in_scale` / `q_out` (fused normalisation of the Lanczos vector).
bool tfim_can_fuse_scale(const dsea_ctx* ctx, const dsea_op* op) {
    Sweep sw[kMaxSweeps];
    const int ns = plan_sweeps(op->L, clamp_tile_bits(ctx->tfim_tile_bits), ctx->tfim_run_bits, ctx->tfim_direct != 0, sw);
    return ns > 0 && ctx->tfim_fuse_scale && pipe_eligible(ctx, sw[0], 1ull << (op->L - sw[0].T));
}

// Shared driver: mode_adj=false -> u = H (in_scale v) (- shift ...), optional dot with `dotw`;
//                mode_adj=true  -> out = -sum_s w[s] * sum_i v[s^(1<<i)].
// `exchange`: how the top-bit shards of v reach this rank when the peer arena is in use:
//   XCH_PUSH       nobody has published v yet: barrier, push kernel, barrier (all attributed to PK_EXCHANGE);
//   XCH_PREPUSHED  the kernel that produced v stored it into the partners' arenas and a collective has completed since;
//   XCH_PREPUSHED_BARRIER  as above but no collective followed the producer: one barrier before the last sweep.
static int tfim_run(dsea_ctx* ctx, const dsea_op* op, bool mode_adj, const double* g, const double* shift,
                    const double* v, double* u, const double* w, double* dot_out, double* work, cudaStream_t st,
                    int exchange, const double* remote_scale, const double* in_scale, double* q_out,
                    bool round_remote = false, bool defer_dot = false) {
    Sweep sw[kMaxSweeps];
    const int L = op->L;
    const int Tmax = clamp_tile_bits(ctx->tfim_tile_bits);
    const int ns = plan_sweeps(L, Tmax, ctx->tfim_run_bits, ctx->tfim_direct != 0, sw);
    DSEA_ARG(ns > 0, "TFIM sweep plan failed");
    const int nrecv = ctx->log2world;
    const bool p2p = nrecv > 0 && ctx->p2p_ok && ctx->arena_stride >= (int64_t)op->n_loc;
    const bool prepushed = exchange != XCH_PUSH;
    DSEA_ARG(nrecv == 0 || p2p || work != nullptr, "sharded TFIM matvec needs a work buffer");
    DSEA_ARG(!prepushed || p2p, "prepushed input without a peer arena");
    DSEA_ARG((in_scale == nullptr && q_out == nullptr) ||
                 (!mode_adj && pipe_eligible(ctx, sw[0], 1ull << (L - sw[0].T))),
             "fused input scaling needs the pipelined first sweep");

    if (p2p) {         // partners' shards arrive in the arena by peer stores
        if (!prepushed) {
            const int xt = prof_begin(ctx, PK_EXCHANGE, 8.0 * (double)op->n_loc * nrecv, st);
            if (!ctx->fresh_collective) DSEA_TRY(comm_barrier(ctx, st));   // partners finished reading the arena
            DSEA_TRY(push_to_peers(ctx, v, op->n_loc, st));
            DSEA_TRY(comm_barrier(ctx, st));                               // partners' stores have landed
            prof_end(ctx, xt, st);
        }
    } else if (nrecv > 0) {   // top-bit shards travel on the side stream while the local sweeps run
        DSEA_CUDA(cudaEventRecord(ctx->ev_ready, st));
        DSEA_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
        DSEA_TRY(exchange_shards(ctx, v, work, op->n_loc, ctx->comm_stream));
        DSEA_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    }
    const bool want_dot = mode_adj || (dot_out != nullptr);
    int total_partials = 0;
    int tok = prof_begin(ctx, mode_adj ? PK_ADJOINT : PK_MATVEC, 16.0 * (double)op->n_loc, st);
    for (int j = 0; j < ns; ++j) {
        const bool last = (j == ns - 1);
        if (last && p2p && exchange == XCH_PREPUSHED_BARRIER) {
            // the local sweeps above overlapped the partners' stores; wait for them only now
            prof_end(ctx, tok, st);
            const int xt = prof_begin(ctx, PK_EXCHANGE, 0.0, st);
            DSEA_TRY(comm_barrier(ctx, st));
            prof_end(ctx, xt, st);
            tok = prof_begin(ctx, mode_adj ? PK_ADJOINT : PK_MATVEC, 0.0, st);
        }
        SweepParams p;
        p.v = v;
        p.uin = u;
        p.uout = u;
        p.g = g;
        p.shift = shift;
        p.in_scale = (j == 0) ? in_scale : nullptr;
        p.q_out = (j == 0) ? q_out : nullptr;
        p.recv = p2p ? ctx->arena : work;
        p.recv_stride = p2p ? (uint64_t)ctx->arena_stride : (uint64_t)op->n_loc;
        p.remote_scale = (p2p && prepushed) ? remote_scale : nullptr;
        p.guard = ctx->guard;
        p.no_diag = (!mode_adj && g == nullptr) ? 1 : 0;
        p.use_tma = (ctx->tfim_tma && sw[j].c == sw[j].T && (((uintptr_t)v) & 127u) == 0) ? 1 : 0;
        p.l2_prefetch = ctx->tfim_l2_prefetch;
        p.round_remote = (round_remote && p2p && prepushed) ? 1 : 0;
        p.pdl_trigger = (ctx->pdl_staged & 2) ? 1 : 0;
        p.nrecv = last ? nrecv : 0;
        p.rank_off = (uint64_t)ctx->rank << L;
        p.n_loc = (uint64_t)op->n_loc;
        p.N = op->N;
        p.T = sw[j].T;
        p.c = sw[j].c;
        p.hshift = sw[j].hshift;
        p.b0 = sw[j].b0;
        p.ndirect = sw[j].ndirect;
        p.upbits = L - (sw[j].hshift + sw[j].T - sw[j].c);
        p.fast_up = sw[j].ndirect > 0 ? 1 : 0;
        p.ntiles = 1ull << (L - sw[j].T);
        // full 2^13 tiles take the persistent double-buffered kernel (one CTA per SM)
        // ... except a last sweep with many global-memory operands per element (direct partner tiles + arena slots): its
        // epilogue is a chain of dependent load rounds, which the 2-CTA/SM generic kernel hides better (measured at
        // L = 26, 4 direct bits: 0.77 vs 0.84 ms per matvec; at L = 24 / 25 with 2 / 3 the pipelined kernel wins)
        // (the adjoint reduction has no output stream and stays pipelined: 0.48 vs 0.63 ms at L = 26)
        const bool pipe = pipe_eligible(ctx, sw[j], p.ntiles) && !(mode_adj && !ctx->tfim_pipe_adjoint) &&
                          !(p.nrecv > 0 && !ctx->tfim_pipe_remote) &&
                          !(!mode_adj && p.ndirect + p.nrecv >= ctx->tfim_generic_min_operands);
        // strided sweeps with global operands in their epilogue take the staged kernel when the slots fit
        const int nops = p.ndirect + p.nrecv + 1;
        // (not the adjoint reduction: it has no output stream and measured 0.26 -> 0.28 ms slower at L = 25 when staged)
        const bool staged = ctx->tfim_stage && pipe_eligible(ctx, sw[j], p.ntiles) && j > 0 && sw[j].c < sw[j].T &&
                            sw[j].c >= 1 && staged_subtiles(nops) > 0 && !mode_adj;
        int grid = (pipe || staged) ? (int)(p.ntiles < (uint64_t)ctx->num_sms ? p.ntiles : (uint64_t)ctx->num_sms)
                                    : (int)(p.ntiles < 2048 ? p.ntiles : 2048);
        if (last && nrecv > 0 && !p2p) DSEA_CUDA(cudaStreamWaitEvent(st, ctx->ev_comm, 0));
        if (mode_adj) {
            p.w = w;
            p.partials = ctx->partials + total_partials;   // every sweep contributes partial sums (main region: up to 40 sweeps x 2048)
            total_partials += grid;
            if (staged) DSEA_TRY(launch_staged_any<MODE_ADJ>(ctx, p, grid, nops, st));
            else DSEA_TRY(launch_sweep<MODE_ADJ>(ctx, p, grid, pipe, st));
        } else {
            p.w = (last && want_dot) ? w : nullptr;
            p.partials = (last && want_dot) ? ctx->partials + kDotPartialsOffset : nullptr;
            if (last && want_dot) total_partials = grid;
            if (j == 0) DSEA_TRY(launch_sweep<MODE_FIRST>(ctx, p, grid, pipe, st));
            else if (staged) DSEA_TRY(launch_staged_any<MODE_ACCUM>(ctx, p, grid, nops, st));
            else DSEA_TRY(launch_sweep<MODE_ACCUM>(ctx, p, grid, pipe, st));
        }
    }
    prof_end(ctx, tok, st);
    if (p2p) ctx->fresh_collective = false;    // the arena was just read: the next push needs a collective first
    ctx->pending_dot_n = 0;
    if (want_dot) {
        if (defer_dot && !mode_adj && ctx->world == 1 && ctx->fuse_small) {
            ctx->pending_dot_n = total_partials;      // the consumer kernel sums them in its prologue
        } else {
            DSEA_TRY(finalize_reduce(ctx, total_partials, 1, dot_out, st,
                                     mode_adj ? ctx->partials : ctx->partials + kDotPartialsOffset));
        }
    }
    return DSEA_OK;
}

int tfim_apply(dsea_ctx* ctx, const dsea_op* op, const double* g, const double* shift, const double* v,
               double* u, const double* dotw, double* dot_out, double* work, cudaStream_t st, int exchange,
               const double* remote_scale, const double* in_scale, double* q_out, bool round_remote, bool defer_dot) {
    // with a fused input scale the dot partner is the SCALED input, which the kernel holds in registers (w == v)
    return tfim_run(ctx, op, false, g, shift, v, u, dotw ? dotw : v, dot_out, work, st, exchange, remote_scale,
                    in_scale, q_out, round_remote, defer_dot);
}

// u = (dH/dg) v: the same sweeps with g = 1 and the diagonal dropped (g == nullptr selects this).
int tfim_dHdg(dsea_ctx* ctx, const dsea_op* op, const double* v, double* u, double* work, cudaStream_t st) {
    return tfim_run(ctx, op, false, nullptr, nullptr, v, u, v, nullptr, work, st, XCH_PUSH, nullptr, nullptr, nullptr);
}

int tfim_adjoint(dsea_ctx* ctx, const dsea_op* op, const double* v1, const double* v2, double* out,
                 double* work, cudaStream_t st) {
    return tfim_run(ctx, op, true, nullptr, nullptr, v2, nullptr, v1, out, work, st, XCH_PUSH, nullptr, nullptr,
                    nullptr);
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
// {T, c, hshift, b0, ndirect} and returns the number of sweeps (negative on failure).  `allow_direct` selects the
// production plan (direct loads for up to 5 top bits) or the pure shared-memory plan.
extern "C" int dsea_tfim_plan(int local_bits, int tile_bits, int run_bits, int allow_direct, int* out5x40) {
    dsea::Sweep sw[dsea::kMaxSweeps];
    const int n = dsea::plan_sweeps(local_bits, dsea::clamp_tile_bits(tile_bits), run_bits, allow_direct != 0, sw);
    for (int j = 0; j < n && j < dsea::kMaxSweeps; ++j) {
        out5x40[5 * j + 0] = sw[j].T;
        out5x40[5 * j + 1] = sw[j].c;
        out5x40[5 * j + 2] = sw[j].hshift;
        out5x40[5 * j + 3] = sw[j].b0;
        out5x40[5 * j + 4] = sw[j].ndirect;
    }
    return n;
}
