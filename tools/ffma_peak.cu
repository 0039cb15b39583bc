// FP32 FFMA peak of the device (the roofline denominator of the sphere-tracing kernel, which no
// driver-written file provides).  Every thread runs 16 independent fused multiply-add chains; a launch fills all SMs with
// 2048 threads each.  Prints one JSON line: {"ffma_tflops": best-of-N, "sm_count": ..., "clock_mhz": ...}.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/build/ffma_peak tools/ffma_peak.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int CHAINS = 16, ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3f + c;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fmaf(x[c], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains alive, never true in practice
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int blocks = prop.multiProcessorCount * 8 * 4;  // 8 resident blocks of 256 threads per SM, 4 waves
    float* out = nullptr;
    cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 12; ++rep) {
        cudaEventRecord(e0);
        k_ffma<<<blocks, 256>>>(out, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * CHAINS * ITERS * (double)blocks * 256;
        if (rep >= 2 && flops / (ms * 1e-3) > best) best = flops / (ms * 1e-3);
    }
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("{\"ffma_tflops\": %.2f, \"sm_count\": %d, \"clock_mhz\": %.0f, \"nominal_tflops\": %.2f, \"how\": \"16 independent FFMA chains x 4096 iterations per thread, %d blocks x 256 threads, best of 10 after 2 warm-ups, CUDA events\"}\n",
           best / 1e12, prop.multiProcessorCount, clock_khz / 1e3, prop.multiProcessorCount * 128 * 2.0 * clock_khz * 1e3 / 1e12, blocks);
    return 0;
}
