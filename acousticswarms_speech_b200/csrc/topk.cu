// topk.cu -- per-mixture MAX_POWER and K best hypercubes of the SRP map.
//
// Reference: MAX_POWER = amax(SRP_map) (sep/Traditional_SP/SRP_Prunning.py:432).  The reference prunes
// by thresholded peak picking (:500-544), not by a top-K; the top-K here is the exchange format of
// the multi-GPU merge (each rank scores a slice of the hypercubes) and a pre-filter for pruning.
//
// One CTA per mixture.  Keys are 64-bit (order-preserving value bits << 32 | ~index) so every key is
// unique and ties resolve to the lower index.  The K-th largest key is found by an 8-pass MSB radix
// select (256-bin shared histogram, warp suffix scan), the K survivors are compacted into shared
// memory and ordered by a bitonic sort.
#include <math_constants.h>

#include "common.cuh"

namespace asw {
namespace {

constexpr int kThreads = 1024;
constexpr int kMaxK = 1024;

__device__ __forceinline__ unsigned long long make_key(float v, int idx) {
    unsigned int u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
}
__device__ __forceinline__ float key_value(unsigned long long k) {
    unsigned int u = (unsigned int)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ int key_index(unsigned long long k) {
    return (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull));
}

__global__ void __launch_bounds__(kThreads) topk_kernel(const float* __restrict__ map, int G, int K, int idx_offset,
                                                         float* __restrict__ val, int32_t* __restrict__ idx) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_keys[kMaxK];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k, s_count;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const float* m = map + (size_t)b * G;
    const int Ke = min(K, G);

    if (tid == 0) {
        s_prefix = 0ull;
        s_k = Ke;
        s_count = 0;
    }
    for (int pass = 0; pass < 8 && Ke > 0; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
        for (int g = tid; g < G; g += kThreads) {
            const unsigned long long key = make_key(m[g], g);
            if ((key & himask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns bins [8l, 8l+8); find the bin where the count from the top reaches k
            unsigned int c[8], tot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                c[i] = hist[8 * tid + i];
                tot += c[i];
            }
            unsigned int above = 0;  // total of higher lanes
            unsigned int run = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int o = __shfl_down_sync(0xffffffffu, run, d);
                if (tid + d < 32) run += o;
            }
            above = run - tot;  // suffix sum excluding this lane
            const unsigned int k = (unsigned int)s_k;
            if (above < k && above + tot >= k) {
                unsigned int cum = above;
                for (int i = 7; i >= 0; --i) {
                    if (cum + c[i] >= k) {
                        s_prefix = prefix | ((unsigned long long)(8 * tid + i) << shift);
                        s_k = (int)(k - cum);
                        break;
                    }
                    cum += c[i];
                }
            }
        }
        __syncthreads();
    }
    // compact the survivors: key >= K-th largest key (keys are unique, so exactly Ke of them)
    const unsigned long long kth = s_prefix;
    for (int i = tid; i < kMaxK; i += kThreads) s_keys[i] = 0ull;
    __syncthreads();
    if (Ke > 0) {
        for (int g = tid; g < G; g += kThreads) {
            const unsigned long long key = make_key(m[g], g);
            if (key >= kth) {
                const int slot = atomicAdd(&s_count, 1);
                if (slot < kMaxK) s_keys[slot] = key;
            }
        }
    }
    __syncthreads();
    // bitonic sort, descending, over the next power of two >= Ke
    int n = 1;
    while (n < Ke) n <<= 1;
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n >> 1); t += kThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = s_keys[lo], c = s_keys[hi];
                if ((a < c) == desc) {
                    s_keys[lo] = c;
                    s_keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += kThreads) {
        float v = -CUDART_INF_F;
        int id = -1;
        if (i < Ke) {
            v = key_value(s_keys[i]);
            id = key_index(s_keys[i]) + idx_offset;
        }
        val[(size_t)b * K + i] = v;
        idx[(size_t)b * K + i] = id;
    }
}

}  // namespace

int launch_topk(const float* map, int B, int G, int K, int idx_offset, float* val, int32_t* idx, cudaStream_t s) {
    if (K < 1 || K > kMaxK) {
        set_error("topk: K=%d outside [1, %d]", K, kMaxK);
        return ASW_ERR_ARG;
    }
    ASW_CARVE_ONCE(topk_kernel);
    topk_kernel<<<B, kThreads, 0, s>>>(map, G, K, idx_offset, val, idx);
    ASW_LAUNCH_CHECK("topk_kernel");
    return ASW_OK;
}

}  // namespace asw
