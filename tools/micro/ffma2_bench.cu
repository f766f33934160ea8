// Micro-benchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, float x, float y, int iters) {
    float2 acc[16];
    float2 b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = make_float2(y + i, y - i);
    const float2 a2 = make_float2(x, x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {  // two scalar FFMAs with three distinct register sources
                acc[i].x = fmaf(x, b[i & 3].x, acc[i].x);
                acc[i].y = fmaf(x, b[i & 3].y, acc[i].y);
            } else {
                acc[i] = fma2(a2, b[i & 3], acc[i]);
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
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            else k<1><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * (double)iters * 32;
        printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2 (f32x2)" : "FFMA  (scalar)", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
    return 0;
}
