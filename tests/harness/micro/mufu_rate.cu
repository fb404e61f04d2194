// Microbenchmark: per-SM throughput of MUFU.EX2 (f32) vs MUFU.EX2.F16 (ex2.approx.f16x2 = two per instruction).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048, kIlp = 8;

__global__ void k_f32(float* out, long long* cyc, float seed) {
    float x[kIlp];
    for (int i = 0; i < kIlp; i++) x[i] = seed + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kIlp; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < kIlp; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_f16x2(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x[kIlp];
    for (int i = 0; i < kIlp; i++) x[i] = seed + 0x00010001u * (threadIdx.x & 15) + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kIlp; i++) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < kIlp; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const int blocks = 148, threads = 256;   // 8 warps per SM, 2 per sub-partition
    float* of; uint32_t* oh; long long* cyc;
    cudaMalloc(&of, blocks * threads * 4); cudaMalloc(&oh, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    long long h[148];
    for (int rep = 0; rep < 2; rep++) {
        k_f32<<<blocks, threads>>>(of, cyc, -0.3f);
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double ops = (double)kIters * kIlp * threads;
        printf("f32   : %lld cycles, %.2f ex2 results / clk / SM\n", h[0], ops / h[0]);
        k_f16x2<<<blocks, threads>>>(oh, cyc, 0xB4CDB4CDu);
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("f16x2 : %lld cycles, %.2f ex2 results / clk / SM (two per instruction)\n", h[0], 2 * ops / h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
