// Microbenchmark: would FOUR softmax warps per SM sub-partition -- two threads per S row, 64 columns each, exponentials
// against the row's current reference (no cross-warp exchange on the S -> P path) -- lift the per-sub-partition softmax
// throughput that bounds the kernel (two tiles per ~2150 cycles with two full-row warps, tests/harness/micro/softmax_bench.cu)?
// 16 warps: warp w serves TMEM lanes 32*(w%4).., tile (w/4)/2, column half (w/4)%2.  Each iteration refills the warp's
// half of S, runs the half-row softmax (chunk max + exp2 + pack + tcgen05.st + arrive) and idles `gap` cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -maxrregcount=112 -o build/softmax_half_bench tests/harness/micro/softmax_half_bench.cu
#include <cstdio>
#include <cstdlib>

#include "../../../flash_attention_cuda_b200/csrc/fa_fwd_sm100.cuh"

using namespace fa;

template <int kPoly>
__device__ __forceinline__ void softmax_half(const Params& p, uint32_t tS, uint32_t bar_p, float& m_ref, float& m_seen, float& l_run) {
    uint32_t a[32], b[32];
    tmem_ld_x32(tS, a);
    tmem_wait_ld();
    tmem_ld_x32(tS + 32, b);
    const float mx0 = max_chunk(a);
    const float m_use = (m_ref == -INFINITY) ? mx0 : m_ref;
    const float neg = -((m_use == -INFINITY) ? 0.0f : m_use) * p.scale_log2;
    const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t sum_a = 0ull, sum_b = 0ull;
    uint32_t pk[32];
    exp_half<kPoly, false, 32>(a, pk, scale2, neg2, sum_a, sum_b);
    tmem_wait_ld();
    const float mx1 = max_chunk(b);
    exp_half<kPoly, false, 32>(b, pk + 16, scale2, neg2, sum_a, sum_b);
    tmem_st_x32(tS, pk);                  // this warp's piece of P over its own S columns
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(bar_p);
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
    m_ref = m_use;
    m_seen = fmaxf(m_seen, fmaxf(mx0, mx1));
}

template <int kPoly>
__global__ void __launch_bounds__(512, 1)
bench(Params p, unsigned long long* out, int iters, int active_mask, int offset1, int gap) {
    __shared__ alignas(8) unsigned long long bars[16];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) mbar_init(smem_u32(&bars[i]), 4);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int grp = warp >> 2, t = grp >> 1, half = grp & 1;
    unsigned long long total = 0, calls = 0;
    if ((active_mask >> t) & 1) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + (t ? 128 : 0) + 64 * half;
        const uint32_t bar_p = smem_u32(&bars[4 * t + half]);
        float m_ref = -INFINITY, m_seen = -INFINITY, l_run = 0.f;
        if (t == 1) {
            const long long s0 = clock64();
            while (clock64() - s0 < offset1) __nanosleep(32);
        }
        for (int it = 0; it < iters; it++) {
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    uint32_t h = (uint32_t)(it * 131 + (2 * half + c) * 32 + i) * 2654435761u + (uint32_t)threadIdx.x * 40503u;
                    v[i] = __float_as_uint(((h >> 9) & 0x7fff) * (4.0f / 32768.0f) - 2.0f);
                }
                if (c == 0) v[0] = __float_as_uint(2.0f);
                tmem_st_x32(tS + 32 * c, v);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            tc_fence_after();
            const long long t0 = clock64();
            softmax_half<kPoly>(p, tS, bar_p, m_ref, m_seen, l_run);
            const long long t1 = clock64();
            if (it >= 4) { total += (unsigned long long)(t1 - t0); ++calls; }
            while (clock64() - t1 < gap) __nanosleep(32);
        }
        if (lane == 0) { out[warp * 2] = total; out[warp * 2 + 1] = calls; }
        if (l_run == 123.456f || m_seen == 77.f) out[63] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int kPoly>
void run(const char* name, int mask, int offset1, int gap) {
    Params p;
    memset(&p, 0, sizeof p);
    p.scale = 1.0f / sqrtf(128.f);
    p.scale_log2 = p.scale * 1.4426950408889634f;
    unsigned long long* d;
    cudaMalloc(&d, 64 * 8);
    cudaMemset(d, 0, 64 * 8);
    bench<kPoly><<<1, 512>>>(p, d, 400, mask, offset1, gap);
    unsigned long long h[64];
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("  poly %d  %-5s gap %4d:", kPoly, name, gap);
    for (int w = 0; w < 16; w += 4)
        if (h[w * 2 + 1]) printf("  tile %d half %d %6.0f cyc", w / 8, (w / 4) & 1, (double)h[w * 2] / h[w * 2 + 1]);
    printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main(int argc, char** argv) {
    const int gap = argc > 1 ? atoi(argv[1]) : 935;
    printf("half-row softmax microbenchmark (two threads per row, 4 warps per sub-partition when both tiles run)\n");
    run<1>("solo", 1, 0, gap);        // one tile: 2 warps per sub-partition, each half a row
    run<1>("dual", 3, 700, gap);      // both tiles, tile 1 half a period later
    run<1>("lock", 3, 0, gap);
    run<1>("solo", 1, 0, 0);
    run<1>("lock", 3, 0, 0);          // back to back: pure throughput, 4 half-tiles per sub-partition per iteration
    run<0>("lock", 3, 0, 0);
    run<2>("lock", 3, 0, 0);
    run<3>("lock", 3, 0, 0);
    return 0;
}
