// Micro-benchmark: issue rate of the legacy warp-level tensor-core path on sm_100a.
//   mma.sync.m16n8k8  tf32 (the instruction of dense_nn_mma_kernel / hidden_bwd_mma_kernel, 1024 MACs each)
//   mma.sync.m16n8k16 bf16 (2048 MACs each) for comparison
// with 4 / 8 / 16 warps per SM and 4 / 8 independent accumulator chains per warp.
// Build on the GPU box (not shipped as a binary):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/mma_sync_rate tools/micro/mma_sync_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int CHAINS>
__global__ void rate_kernel(float* out, int iters, uint32_t seed) {
    float acc[CHAINS][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = seed + threadIdx.x * 7u + i;
    b[0] = seed ^ 0x3f800000u;
    b[1] = seed ^ 0x3f000000u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0) mma_tf32(acc[c], a, b);
            else mma_bf16(acc[c], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) s += acc[c][e];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every MMA of the inner loop reads its own A and B registers: no operand-collector reuse between consecutive instructions
template <int CHAINS>
__global__ void rate_distinct_kernel(float* out, int iters, uint32_t seed) {
    float acc[CHAINS][4];
    uint32_t a[CHAINS][4], b[CHAINS][2];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[c][i] = seed + threadIdx.x * 7u + i + 13u * c;
        b[c][0] = (seed ^ 0x3f800000u) + c;
        b[c][1] = (seed ^ 0x3f000000u) + 3u * c;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma_tf32(acc[c], a[c], b[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) s += acc[c][e];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
static void run_distinct(int warps, float* out, int sms, double clk_ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_distinct_kernel<CHAINS><<<sms, warps * 32>>>(out, 100, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_distinct_kernel<CHAINS><<<sms, warps * 32>>>(out, iters, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * warps * CHAINS * iters;
    printf("{\"kind\": \"tf32_m16n8k8_distinct_operands\", \"warps_per_sm\": %d, \"chains\": %d, \"ms\": %.4f, "
           "\"mma_per_us_per_sm\": %.1f, \"mma_per_clk_per_sm\": %.4f, \"tflops\": %.1f}\n",
           warps, CHAINS, ms, mmas / sms / (ms * 1e3), mmas / sms / (ms * 1e-3 * clk_ghz * 1e9), 2.0 * mmas * 1024.0 / (ms * 1e-3) / 1e12);
}

template <int KIND, int CHAINS>
static void run(const char* name, int warps, float* out, int sms, double clk_ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<KIND, CHAINS><<<sms, warps * 32>>>(out, 100, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<KIND, CHAINS><<<sms, warps * 32>>>(out, iters, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * warps * CHAINS * iters;
    const double macs = mmas * (KIND == 0 ? 1024.0 : 2048.0);
    printf("{\"kind\": \"%s\", \"warps_per_sm\": %d, \"chains\": %d, \"ms\": %.4f, \"mma_per_us_per_sm\": %.1f, "
           "\"mma_per_clk_per_sm\": %.4f, \"mac_per_clk_per_sm\": %.1f, \"tflops\": %.1f}\n",
           name, warps, CHAINS, ms, mmas / sms / (ms * 1e3), mmas / sms / (ms * 1e-3 * clk_ghz * 1e9),
           macs / sms / (ms * 1e-3 * clk_ghz * 1e9), 2.0 * macs / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    const int sms = p.multiProcessorCount;
    float* out;
    cudaMalloc(&out, (size_t)sms * 1024 * sizeof(float));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, sms, ghz);
    run<0, 4>("tf32_m16n8k8", 4, out, sms, ghz);
    run<0, 8>("tf32_m16n8k8", 4, out, sms, ghz);
    run<0, 4>("tf32_m16n8k8", 8, out, sms, ghz);
    run<0, 8>("tf32_m16n8k8", 8, out, sms, ghz);
    run<0, 8>("tf32_m16n8k8", 16, out, sms, ghz);
    run_distinct<6>(4, out, sms, ghz);
    run_distinct<6>(8, out, sms, ghz);
    run_distinct<6>(16, out, sms, ghz);
    run<1, 4>("bf16_m16n8k16", 4, out, sms, ghz);
    run<1, 8>("bf16_m16n8k16", 8, out, sms, ghz);
    run<1, 8>("bf16_m16n8k16", 16, out, sms, ghz);
    cudaFree(out);
    return 0;
}
