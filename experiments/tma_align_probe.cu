// tma_align_probe.cu -- does a tiled TMA load accept an inner coordinate that is not a multiple of 16 bytes?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, float* out, int* status) {
    __shared__ __align__(128) float buf[256];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                     ::"r"(smem_u32(buf)), "l"(&map), "r"(c0), "r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        for (unsigned spin = 0; spin < 4000000u && !ok; ++spin)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        *status = ok;
        if (ok) for (int i = 0; i < 256; ++i) out[i] = buf[i];
    }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const int n = 4096;
    float h[n]; for (int i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *o; int* st;
    CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&o, 1024)); CK(cudaMalloc(&st, 4));
    CK(cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice));
    const CUtensorMapDataType types[3] = {CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_DATA_TYPE_UINT32, CU_TENSOR_MAP_DATA_TYPE_UINT8};
    for (int ty = 0; ty < 3; ++ty) {
        CUtensorMap m;
        const int esz = ty == 2 ? 1 : 4;
        cuuint64_t dims[1] = {(cuuint64_t)(n * 4 / esz)}; cuuint64_t strides[1] = {0};
        cuuint32_t box[1] = {(cuuint32_t)(ty == 2 ? 256 : 256)}, es[1] = {1};
        CUresult r = enc(&m, types[ty], 1, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("type %d encode rc=%d\n", ty, (int)r);
        if (r) continue;
        const int coords[] = {0, 4, 8, 1, 2, 3, 5, -4, -1};
        for (int c : coords) {
            CK(cudaMemset(st, 0xff, 4));
            probe<<<1, 32>>>(m, c, o, st);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  type %d coord %d: CUDA error %s\n", ty, c, cudaGetErrorString(e)); return 0; }
            int s; float first[2];
            CK(cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(first, o, 8, cudaMemcpyDeviceToHost));
            printf("  type %d coord %d: landed=%d first=%g second=%g\n", ty, c, s, first[0], first[1]);
        }
    }
    return 0;
}
