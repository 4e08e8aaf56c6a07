// Micro-benchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_ffma2 microbench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 x = make_float2(a, a * 0.5f), y = make_float2(b, b * 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {
                acc[i].x = fmaf(x.x, acc[i].x, y.x);
                acc[i].y = fmaf(x.y, acc[i].y, y.y);
            } else {
                acc[i] = __ffma2_rn(x, acc[i], y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 512 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512}) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0)
                    k<0><<<148 * (2048 / threads), threads>>>(out, iters, 0.999f, 0.001f);
                else
                    k<1><<<148 * (2048 / threads), threads>>>(out, iters, 0.999f, 0.001f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 1) {
                    const double fma = (double)148 * 2048 * iters * 32.0;
                    printf("threads/block %d mode %s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", threads, mode ? "FFMA2" : "FFMA ", ms,
                           fma / ms / 1e9, 2 * fma / ms / 1e9);
                }
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
