// shift_grouped_probe.cu -- does staging the source once per (mixture, mic, chunk) in shared memory and writing all of
// that mixture's patches from it beat the direct kernel (every patch re-reads its source from L2)?
//
// direct : grid (chunks, M, N), 256 threads x 4 vectors, two aligned 16 B global loads + register funnel per vector
//          (the product kernel, shift_stack.cu).
// grouped: grid (chunks, M, runs); a run = consecutive patches of one mixture.  The CTA loads the window
//          [t0 + rmin, t0 + 4096 + rmax) of the source row into shared memory once (aligned 16 B loads), then every
//          patch of the run is two aligned LDS.128 + funnel per vector and one streaming 16 B store.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o experiments/bin/shift_grouped_probe experiments/shift_grouped_probe.cu
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 256, kVpt = 4, kChunk = kThreads * kVpt * 4;   // 4096 floats per CTA and row
constexpr int kMaxSpread = 2048;                                         // rmax - rmin the staged window can hold
constexpr int kWin = kChunk + kMaxSpread + 8;

template <int O>
__device__ __forceinline__ float4 funnel(float4 a, float4 c) {
    if (O == 0) return a;
    if (O == 1) return make_float4(a.y, a.z, a.w, c.x);
    if (O == 2) return make_float4(a.z, a.w, c.x, c.y);
    return make_float4(a.w, c.x, c.y, c.z);
}

template <int O>
__device__ __forceinline__ void direct_chunk(const float* __restrict__ src, float4* __restrict__ dst, int r, int T) {
    const int T4 = T >> 2;
    const int v0 = blockIdx.x * (kThreads * kVpt) + threadIdx.x;
    float4 val[kVpt];
#pragma unroll
    for (int v = 0; v < kVpt; ++v) {
        const int t4 = v0 + v * kThreads;
        if (t4 < T4) {
            int s = 4 * t4 + r;
            if (s >= T) s -= T;
            const int base = s - O;
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + base));
            float4 c = a;
            if (O) c = __ldg(reinterpret_cast<const float4*>(src + ((base + 4 < T) ? base + 4 : 0)));
            val[v] = funnel<O>(a, c);
        }
    }
#pragma unroll
    for (int v = 0; v < kVpt; ++v) {
        const int t4 = v0 + v * kThreads;
        if (t4 < T4) __stcs(dst + t4, val[v]);
    }
}

__global__ void __launch_bounds__(kThreads) direct_kernel(const float* __restrict__ mix, const int* __restrict__ shifts,
                                                          const int* __restrict__ mi, int M, int T, float* __restrict__ out) {
    const int c = blockIdx.y, n = blockIdx.z, row = n * M + c;
    const float* src = mix + ((size_t)mi[n] * M + c) * (size_t)T;
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * T);
    int r = shifts[row] % T;
    if (r < 0) r += T;
    switch (r & 3) {
        case 0: direct_chunk<0>(src, dst, r, T); break;
        case 1: direct_chunk<1>(src, dst, r, T); break;
        case 2: direct_chunk<2>(src, dst, r, T); break;
        default: direct_chunk<3>(src, dst, r, T); break;
    }
}

template <int O>
__device__ __forceinline__ void staged_patch(const float* __restrict__ win, float4* __restrict__ dst, int o, int T4) {
    // o = offset (floats, multiple-of-4 part already removed into O) of this patch's chunk inside the window
    const int v0 = blockIdx.x * (kThreads * kVpt) + threadIdx.x;
    const float4* w4 = reinterpret_cast<const float4*>(win + o);
#pragma unroll
    for (int v = 0; v < kVpt; ++v) {
        const int j = threadIdx.x + v * kThreads;
        const int t4 = v0 + v * kThreads;
        if (t4 < T4) {
            const float4 a = w4[j];
            float4 c = a;
            if (O) c = w4[j + 1];
            __stcs(dst + t4, funnel<O>(a, c));
        }
    }
}

__global__ void __launch_bounds__(kThreads) grouped_kernel(const float* __restrict__ mix, const int* __restrict__ shifts,
                                                           const int* __restrict__ run_start, int M, int T,
                                                           const int* __restrict__ mi, float* __restrict__ out) {
    __shared__ __align__(16) float win[kWin];
    __shared__ int s_min, s_max;
    const int c = blockIdx.y, run = blockIdx.z;
    const int n0 = run_start[run], n1 = run_start[run + 1];
    const int t0 = blockIdx.x * kChunk;
    const float* src = mix + ((size_t)mi[n0] * M + c) * (size_t)T;
    // shifts as signed offsets in (-T/2, T/2]: the spread of one mixture's patches is a few hundred samples
    if (threadIdx.x == 0) { s_min = INT_MAX; s_max = INT_MIN; }
    __syncthreads();
    for (int n = n0 + threadIdx.x; n < n1; n += kThreads) {
        int r = shifts[n * M + c] % T;
        if (r > T / 2) r -= T;
        if (r <= -T / 2) r += T;
        atomicMin(&s_min, r);
        atomicMax(&s_max, r);
    }
    __syncthreads();
    const int rmin = s_min, rmax = s_max;
    const int len = min(kChunk, T - t0);
    const int a0 = (t0 + rmin) & ~3;                               // window start, aligned (may be negative)
    const int nload = ((t0 + len + rmax + 3) - a0 + 3) >> 2;       // float4s to stage (+1 for the funnel's second vector)
    if (rmax - rmin <= kMaxSpread) {
        for (int i = threadIdx.x; i <= nload && 4 * i + 4 <= kWin; i += kThreads) {
            int s = a0 + 4 * i;
            s %= T;
            if (s < 0) s += T;                                     // T % 4 == 0: the vector never straddles the wrap
            reinterpret_cast<float4*>(win)[i] = __ldg(reinterpret_cast<const float4*>(src + s));
        }
        __syncthreads();
        const int T4 = T >> 2;
        for (int n = n0; n < n1; ++n) {
            int r = shifts[n * M + c] % T;
            if (r > T / 2) r -= T;
            if (r <= -T / 2) r += T;
            const int off = t0 + r - a0;                           // >= 0
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)n * M + c) * T);
            switch (off & 3) {
                case 0: staged_patch<0>(win, dst, off & ~3, T4); break;
                case 1: staged_patch<1>(win, dst, off & ~3, T4); break;
                case 2: staged_patch<2>(win, dst, off & ~3, T4); break;
                default: staged_patch<3>(win, dst, off & ~3, T4); break;
            }
        }
    }
}

int main(int argc, char** argv) {
    const int N = 128, B = argc > 1 ? atoi(argv[1]) : 4, M = 7, T = 144000;
    const int chunks = (T + kChunk - 1) / kChunk;
    std::vector<float> mix((size_t)B * M * T);
    uint32_t st = 12345;
    for (auto& v : mix) { st = st * 1664525u + 1013904223u; v = (float)(int)(st >> 8) * (1.f / 16777216.f) - 0.5f; }
    std::vector<int> shifts((size_t)N * M), mi(N), runs;
    for (int n = 0; n < N; ++n) {
        mi[n] = n * B / N;                                           // sorted by mixture, N / B patches each
        if (n == 0 || mi[n] != mi[n - 1]) runs.push_back(n);
        shifts[(size_t)n * M] = 0;
        for (int c = 1; c < M; ++c) { st = st * 1664525u + 1013904223u; shifts[(size_t)n * M + c] = (int)(st >> 8) % 701 - 350; }
    }
    runs.push_back(N);
    const int R = (int)runs.size() - 1;
    float *d_mix, *d_out, *d_out2;
    int *d_sh, *d_mi, *d_runs;
    const size_t out_bytes = (size_t)N * M * T * 4;
    CK(cudaMalloc(&d_mix, mix.size() * 4)); CK(cudaMalloc(&d_out, out_bytes)); CK(cudaMalloc(&d_out2, out_bytes));
    CK(cudaMalloc(&d_sh, shifts.size() * 4)); CK(cudaMalloc(&d_mi, N * 4)); CK(cudaMalloc(&d_runs, runs.size() * 4));
    CK(cudaMemcpy(d_mix, mix.data(), mix.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sh, shifts.data(), shifts.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mi, mi.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_runs, runs.data(), runs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, out_bytes)); CK(cudaMemset(d_out2, 0xee, out_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double mb = out_bytes / 1e6;
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        for (int w = 0; w < 3; ++w) direct_kernel<<<dim3(chunks, M, N), kThreads>>>(d_mix, d_sh, d_mi, M, T, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 20; ++w) direct_kernel<<<dim3(chunks, M, N), kThreads>>>(d_mix, d_sh, d_mi, M, T, (w & 1) ? d_out2 : d_out);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("direct : %.1f us  %.0f GB/s\n", ms / 20 * 1e3, mb / (ms / 20));
        for (int w = 0; w < 3; ++w) grouped_kernel<<<dim3(chunks, M, R), kThreads>>>(d_mix, d_sh, d_runs, M, T, d_mi, d_out2);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 20; ++w) grouped_kernel<<<dim3(chunks, M, R), kThreads>>>(d_mix, d_sh, d_runs, M, T, d_mi, (w & 1) ? d_out : d_out2);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("grouped: %.1f us  %.0f GB/s  (%d runs of ~%d patches)\n", ms / 20 * 1e3, mb / (ms / 20), R, N / R);
    }
    // correctness: both kernels against each other and a few rows against the host
    direct_kernel<<<dim3(chunks, M, N), kThreads>>>(d_mix, d_sh, d_mi, M, T, d_out);
    grouped_kernel<<<dim3(chunks, M, R), kThreads>>>(d_mix, d_sh, d_runs, M, T, d_mi, d_out2);
    CK(cudaDeviceSynchronize());
    std::vector<float> a((size_t)N * M * T), b((size_t)N * M * T);
    CK(cudaMemcpy(a.data(), d_out, out_bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), d_out2, out_bytes, cudaMemcpyDeviceToHost));
    printf("grouped == direct: %d\n", (int)(memcmp(a.data(), b.data(), out_bytes) == 0));
    size_t bad = 0;
    for (int n = 0; n < N; n += 17)
        for (int c = 0; c < M; ++c) {
            long long r = shifts[(size_t)n * M + c] % T; if (r < 0) r += T;
            for (int t = 0; t < T; ++t)
                if (b[((size_t)n * M + c) * T + t] != mix[((size_t)mi[n] * M + c) * T + (t + r) % T]) ++bad;
        }
    printf("host check mismatches: %zu\n", bad);
    return 0;
}
