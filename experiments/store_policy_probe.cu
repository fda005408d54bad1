// Probe (NOT product code): which store policy / run length lets a shifted copy approach the 7.19 TB/s fill ceiling?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/bin/store_policy_probe experiments/store_policy_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int POL> __device__ __forceinline__ void st4(float4* p, float4 v) {
    if (POL == 0) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (POL == 1) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (POL == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (POL == 3) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (POL == 4) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// persistent: warp takes SEG consecutive runs of 32*VPT vectors of one row; aligned shift only (r % 4 == 0): the
// realignment shuffles are not what is being measured
template <int POL, int VPT, int SEG>
__global__ void __maxnreg__(64) copy_persist(const float4* __restrict__ src, float4* __restrict__ dst, const int* __restrict__ rq,
                                              const int* __restrict__ mi, int rows_per_mix, int T4, int rows, unsigned* ctr) {
    const int lane = threadIdx.x & 31;
    const int runs = T4 / (32 * VPT);           // T4 chosen as a multiple
    const int segs = runs / SEG;
    const unsigned total = (unsigned)rows * segs;
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(ctr, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= total) break;
        const int row = it / segs, seg = it - row * segs;
        const int n = row / rows_per_mix, c = row - n * rows_per_mix;
        const float4* s = src + ((size_t)mi[n] * rows_per_mix + c) * T4;
        float4* d = dst + (size_t)row * T4;
        const int q = rq[row];
        for (int run = seg * SEG; run < (seg + 1) * SEG; ++run) {
            const int v0 = run * 32 * VPT + lane;
            float4 a[VPT];
#pragma unroll
            for (int v = 0; v < VPT; ++v) { int i = v0 + 32 * v + q; if (i >= T4) i -= T4; a[v] = ld4(s + i); }
#pragma unroll
            for (int v = 0; v < VPT; ++v) st4<POL>(d + v0 + 32 * v, a[v]);
        }
    }
}

__global__ void fill_kernel(float4* dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

template <int POL, int VPT, int SEG>
float run(const float4* src, float4* dst, const int* rq, const int* mi, int M, int T4, int rows, unsigned* ctr, int ctas, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaMemsetAsync(ctr, 0, 4));
        cudaEventRecord(e0);
        copy_persist<POL, VPT, SEG><<<ctas, threads>>>(src, dst, rq, mi, M, T4, rows, ctr);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const int B = 64, M = 7, T = 144384 /* multiple of 4*32*12*... */, N = 1152;
    const int T4 = T / 4;   // 36096 = 32 * 12 * 94
    const int rows = N * M;
    float4 *src, *dst; int *rq, *mi; unsigned* ctr;
    CK(cudaMalloc(&src, (size_t)B * M * T4 * 16));
    CK(cudaMalloc(&dst, (size_t)rows * T4 * 16));
    CK(cudaMalloc(&rq, rows * 4)); CK(cudaMalloc(&mi, N * 4)); CK(cudaMalloc(&ctr, 4));
    CK(cudaMemset(src, 0, (size_t)B * M * T4 * 16));
    std::vector<int> hq(rows), hm(N);
    for (int i = 0; i < rows; ++i) hq[i] = (i % M == 0) ? 0 : (rand() % 150);
    for (int n = 0; n < N; ++n) hm[n] = n * B / N;
    CK(cudaMemcpy(rq, hq.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(mi, hm.data(), N * 4, cudaMemcpyHostToDevice));
    const double gb = (double)rows * T4 * 16 / 1e9;
    {   // fill ceiling
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0); fill_kernel<<<148 * 8, 256>>>(dst, (size_t)rows * T4); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("fill: %.3f ms %.0f GB/s\n", best, gb / best * 1e3);
    }
#define RUN(POL, VPT, SEG, CTAS, TH) { float ms = run<POL, VPT, SEG>(src, dst, rq, mi, M, T4, rows, ctr, CTAS, TH); \
    printf("policy %d vpt %d seg %d ctas %d threads %d: %.3f ms %.0f GB/s\n", POL, VPT, SEG, CTAS, TH, ms, gb / ms * 1e3); }
    RUN(0, 12, 2, 148, 512) RUN(1, 12, 2, 148, 512) RUN(2, 12, 2, 148, 512) RUN(3, 12, 2, 148, 512) RUN(4, 12, 2, 148, 512)
    RUN(0, 12, 1, 148, 512) RUN(0, 12, 47, 148, 512) RUN(1, 12, 47, 148, 512)
    RUN(0, 12, 2, 296, 512) RUN(0, 6, 2, 296, 512) RUN(0, 12, 2, 148, 1024) RUN(0, 12, 2, 148, 256) RUN(0, 12, 2, 296, 256)
    RUN(1, 12, 2, 296, 512) RUN(3, 12, 2, 296, 512)
    return 0;
}
