// FP32 pipe microbenchmark for sm_100a: scalar FFMA chains vs packed FFMA2 (fma.rn.f32x2) chains.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_peak fp32_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) chains(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
    float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (PACKED) {
                acc[k] = __ffma2_rn(acc[k], x, y);
            } else {
                acc[k].x = __fmaf_rn(acc[k].x, x.x, y.x);
                acc[k].y = __fmaf_rn(acc[k].y, x.y, y.y);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; ++packed) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (packed) chains<1><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
            else chains<0><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double flop = 2.0 * 16 * (double)iters * blocks * threads;
        printf("%s: %.3f ms  %.2f TFLOP/s  (%d SMs, max clock %.0f MHz, nominal peak %.2f TFLOP/s)\n", packed ? "FFMA2 (f32x2)" : "FFMA (scalar)",
               best, flop / best / 1e9, sms, khz / 1e3, sms * 128 * 2.0 * khz / 1e9);
    }
    return 0;
}
