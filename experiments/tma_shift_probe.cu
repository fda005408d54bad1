// tma_shift_probe.cu -- can the shift-stack be done by the copy engine alone?
//
// One persistent CTA of one warp per SM.  Every 16 KB tile of an output row is fetched from the (L2-resident)
// mixture by 1-D tensor TMA loads (boxes of 256 floats at an ARBITRARY element coordinate: the source of a shifted
// row is misaligned by r mod 4 samples, which plain 16-byte bulk copies cannot express) and written back with one
// bulk shared->global store.  Tiles that straddle the circular wrap read from a small "seam" buffer holding
// [x[T-4096..T), x[0..4096)] of every source row, so they are contiguous as well.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tma_shift_probe experiments/tma_shift_probe.cu
// run  : timeout 120 ./tma_shift_probe [N=128] [B=32] [stages=8]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kTile = 4096;   // floats per tile
constexpr int kBox = 256;     // floats per TMA box (the per-dimension box limit)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void* dst, const CUtensorMap* map, int c0, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void seam_kernel(const float* __restrict__ mix, int rows, int T, float* __restrict__ seam) {
    const int row = blockIdx.y;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 2 * kTile; k += gridDim.x * blockDim.x)
        seam[(size_t)row * 2 * kTile + k] = mix[(size_t)row * T + (k < kTile ? T - kTile + k : k - kTile)];
}

template <int S, bool BULK>
__global__ void __launch_bounds__(32) tma_shift_kernel(const float* __restrict__ mix, const float* __restrict__ seam,
                                                       const __grid_constant__ CUtensorMap mix_map,
                                                       const __grid_constant__ CUtensorMap seam_map,
                                                       const int* __restrict__ shifts, const int* __restrict__ mix_index,
                                                       int N, int M, int T, int tiles_per_row, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    __shared__ uint64_t full[S];
    constexpr int L = S / 2;   // loads in flight; the other half of the ring is draining through stores
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long total = (long long)N * M * tiles_per_row;
    const long long first = blockIdx.x, step = gridDim.x;
    const long long K = first < total ? (total - first + step - 1) / step : 0;

    for (long long k = 0; k < K + L; ++k) {
        if (k >= L) {                                   // consume tile j: loads landed -> one bulk store
            const long long j = k - L;
            const int s = (int)(j % S);
            const uint32_t parity = (uint32_t)((j / S) & 1);
            for (unsigned spin = 0; !mbar_try_wait(&full[s], parity); ++spin)
                if (spin > 20000000u) __trap();          // never hang the box on a protocol mistake
            if (lane == 0) {
                const long long i = first + j * step;
                const long long row = i / tiles_per_row;
                const int t0 = (int)(i % tiles_per_row) * kTile;
                const int len = min(kTile, T - t0);
                bulk_store(out + row * T + t0, stage + (size_t)s * kTile, (uint32_t)len * 4u);
                bulk_commit();
            }
        }
        if (k < K) {                                    // produce tile k into stage k % S
            const int s = (int)(k % S);
            if (lane == 0) bulk_wait_read<S - L>();     // the store that last used this stage has read it
            __syncwarp();
            const long long i = first + k * step;
            const long long row = i / tiles_per_row;
            const int n = (int)(row / M), c = (int)(row % M);
            const int t0 = (int)(i % tiles_per_row) * kTile;
            const int len = min(kTile, T - t0);
            int r = shifts[n * M + c];
            if (r <= -T || r >= T) r %= T;
            if (r < 0) r += T;
            int src = t0 + r;
            if (src >= T) src -= T;
            const long long srow = (long long)mix_index[n] * M + c;
            const int nbox = (len + kBox - 1) / kBox;
            const bool wraps = src + len > T;
            const long long c0 = wraps ? srow * (2 * kTile) + (src - (T - kTile)) : srow * T + src;
            if (BULK) {                                 // aligned sources only: one 16 KB bulk copy per tile
                if (lane == 0) {
                    mbar_expect_tx(&full[s], (uint32_t)len * 4u);
                    bulk_load(stage + (size_t)s * kTile, (wraps ? seam : mix) + c0, (uint32_t)len * 4u, &full[s]);
                }
            } else {
                if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)nbox * kBox * 4u);
                __syncwarp();
                const CUtensorMap* map = wraps ? &seam_map : &mix_map;
                if (lane < nbox)
                    tma_load_1d(stage + (size_t)s * kTile + lane * kBox, map, (int)(c0 + lane * kBox), &full[s]);
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map_1d(EncodeFn enc, void* base, uint64_t elems) {
    CUtensorMap m;
    cuuint64_t dims[1] = {elems};
    cuuint64_t strides[1] = {0};
    cuuint32_t box[1] = {kBox};
    cuuint32_t estr[1] = {1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

template <int S, bool BULK>
static float run(const float* mix, const float* seam, const CUtensorMap& mm, const CUtensorMap& sm, const int* shifts, const int* mi, int N, int M, int T,
                 float* out, int grid, int iters) {
    const int tiles_per_row = (T + kTile - 1) / kTile;
    const size_t smem = (size_t)S * kTile * 4;
    CK(cudaFuncSetAttribute(tma_shift_kernel<S, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) tma_shift_kernel<S, BULK><<<grid, 32, smem>>>(mix, seam, mm, sm, shifts, mi, N, M, T, tiles_per_row, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int w = 0; w < iters; ++w) tma_shift_kernel<S, BULK><<<grid, 32, smem>>>(mix, seam, mm, sm, shifts, mi, N, M, T, tiles_per_row, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters * 1e3f;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128, B = argc > 2 ? atoi(argv[2]) : 32;
    const int M = 7, T = 144000;
    const int align = argc > 3 ? atoi(argv[3]) : 1;   // shifts forced to multiples of this
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeFn enc = (EncodeFn)fn;

    const size_t rows = (size_t)B * M;
    std::vector<float> mix(rows * T);
    uint32_t st = 12345;
    for (auto& v : mix) { st = st * 1664525u + 1013904223u; v = (float)(int)(st >> 8) * (1.f / 16777216.f) - 0.5f; }
    std::vector<int> shifts((size_t)N * M), mi(N);
    for (int n = 0; n < N; ++n) {
        mi[n] = n % B;
        shifts[(size_t)n * M] = 0;
        for (int c = 1; c < M; ++c) { st = st * 1664525u + 1013904223u; shifts[(size_t)n * M + c] = (int)(st >> 8) % 701 - 350; }
    }
    shifts[1] = T - 1; shifts[2] = -T; shifts[3] = 3 * T + 5; shifts[4] = -(T - 4096); shifts[5] = 4095; shifts[6] = -4097;
    for (auto& v : shifts) v = v / align * align;
    printf("align=%d\n", align); fflush(stdout);

    float *d_mix, *d_seam, *d_out;
    int *d_sh, *d_mi;
    CK(cudaMalloc(&d_mix, rows * T * 4));
    CK(cudaMalloc(&d_seam, rows * 2 * kTile * 4));
    CK(cudaMalloc(&d_out, (size_t)N * M * T * 4));
    CK(cudaMalloc(&d_sh, shifts.size() * 4));
    CK(cudaMalloc(&d_mi, mi.size() * 4));
    CK(cudaMemcpy(d_mix, mix.data(), rows * T * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sh, shifts.data(), shifts.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mi, mi.data(), mi.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, (size_t)N * M * T * 4));
    seam_kernel<<<dim3(8, (unsigned)rows), 256>>>(d_mix, (int)rows, T, d_seam);
    CK(cudaDeviceSynchronize());
    printf("seam ok\n"); fflush(stdout);
    CUtensorMap mm = make_map_1d(enc, d_mix, rows * T), sm = make_map_1d(enc, d_seam, rows * 2 * kTile);

    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double mb = (double)N * M * T * 4 / 1e6;
    float us = run<8, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
    printf("S=8  grid=%d: %.1f us  %.0f GB/s written\n", sms, us, mb / us * 1e3);

    std::vector<float> got((size_t)N * M * T);
    CK(cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < M; ++c) {
            long long r = shifts[(size_t)n * M + c] % T; if (r < 0) r += T;
            const float* src = &mix[((size_t)mi[n] * M + c) * T];
            const float* g = &got[((size_t)n * M + c) * T];
            for (int t = 0; t < T; ++t) {
                const float want = src[(t + r) % T];
                if (memcmp(&want, &g[t], 4) != 0) { if (bad < 5) printf("mismatch n=%d c=%d t=%d r=%lld\n", n, c, t, r); ++bad; }
            }
        }
    printf("mismatches: %zu of %zu\n", bad, got.size());

    us = run<4, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
    printf("S=4  grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
    us = run<6, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
    printf("S=6  grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
    us = run<12, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
    printf("S=12 grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
    us = run<6, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, 2 * sms, 20);
    printf("S=6  grid=%d: %.1f us  %.0f GB/s\n", 2 * sms, us, mb / us * 1e3);
    us = run<4, false>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, 3 * sms, 20);
    printf("S=4  grid=%d: %.1f us  %.0f GB/s\n", 3 * sms, us, mb / us * 1e3);
    if (align % 4 == 0) {
        CK(cudaMemset(d_out, 0xff, (size_t)N * M * T * 4));
        us = run<8, true>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
        printf("bulk S=8  grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
        std::vector<float> got2((size_t)N * M * T);
        CK(cudaMemcpy(got2.data(), d_out, got2.size() * 4, cudaMemcpyDeviceToHost));
        printf("bulk equals tensor path: %d\n", (int)(memcmp(got2.data(), got.data(), got.size() * 4) == 0));
        us = run<4, true>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
        printf("bulk S=4  grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
        us = run<12, true>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, sms, 20);
        printf("bulk S=12 grid=%d: %.1f us  %.0f GB/s\n", sms, us, mb / us * 1e3);
        us = run<6, true>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, 2 * sms, 20);
        printf("bulk S=6  grid=%d: %.1f us  %.0f GB/s\n", 2 * sms, us, mb / us * 1e3);
        us = run<2, true>(d_mix, d_seam, mm, sm, d_sh, d_mi, N, M, T, d_out, 4 * sms, 20);
        printf("bulk S=2  grid=%d: %.1f us  %.0f GB/s\n", 4 * sms, us, mb / us * 1e3);
    }
    return bad ? 2 : 0;
}
