// Micro-benchmark: do warp shuffles and shared-memory loads share one per-SM bandwidth budget on B200?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o shfl_vs_lds shfl_vs_lds.cu && ./shfl_vs_lds
// Three kernels with the same loop structure (512 threads x 148 CTAs, 1 CTA/SM):
//   lds : 8 LDS.128 per iteration                    (512 B per warp-instruction = 4 wavefronts)
//   shfl: 32 SHFL.BFLY.32 per iteration              (128 B per warp-instruction)  -> same bytes as 8 LDS.128
//   both: 8 LDS.128 + 32 SHFL per iteration
// If `both` takes ~ max(lds, shfl) the two paths are independent; if ~ lds + shfl they share the crossbar.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += 512) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc[8];
    float f[32];
    for (int j = 0; j < 8; ++j) acc[j] = make_double2(0, 0);
    for (int j = 0; j < 32; ++j) f[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
        if (MODE != 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double2 y = sm[(threadIdx.x + 512 * j) ^ (1 + (it & 255))];
                acc[j].x += y.x;
                acc[j].y += y.y;
            }
        }
        if (MODE != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __shfl_xor_sync(0xffffffffu, f[j], 1 + (it & 15));
        }
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
    for (int j = 0; j < 32; ++j) s += f[j];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

template <int MODE>
float run(double* out, int iters) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<148, 512, 65536>>>(out, 10);
    cudaEventRecord(a);
    k<MODE><<<148, 512, 65536>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 512 * 8);
    const int iters = 20000;
    const float t0 = run<0>(out, iters), t1 = run<1>(out, iters), t2 = run<2>(out, iters);
    // per SM per iteration: 16 warps x 8 LDS.128 = 128 warp-instr x 4 wavefronts = 512 wavefronts; shuffles: 16 x 32 = 512 warp-instr
    printf("{\"lds_ms\": %.3f, \"shfl_ms\": %.3f, \"both_ms\": %.3f, \"lds_wavefronts_per_us_per_sm\": %.1f, "
           "\"shfl_instr_per_us_per_sm\": %.1f, \"both_over_sum\": %.3f, \"both_over_max\": %.3f}\n",
           t0, t1, t2, 512.0 * iters / (t0 * 1e3), 512.0 * iters / (t1 * 1e3), t2 / (t0 + t1), t2 / (t0 > t1 ? t0 : t1));
    return 0;
}
