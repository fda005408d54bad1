// powers.cu -- per-patch output statistics after the separator: de-mean, power and max_avg_power.
//
// Replaces the per-row numpy loops of binary_search_baseline (sep/helpers/local_utils_3d.py:339-360) and
// Spotform_Small_Patch_Parallel (sep/Mic_Array.py:288-296):
//     x      = x - np.mean(x)
//     power  = np.sum(x ** 2)
//     power2 = max_avg_power(x)[0]            (local_utils_3d.py:13-17: 12000-tap box RMS, zero padded to the right)
// One CTA per row.  The box sums are a running window: S[0] = sum of the first W squares, S[i+1] = S[i] + q[i+W] - q[i],
// evaluated tile by tile with a block-wide scan of the differences (coalesced, two reads per sample) and carried in
// double, which is what scipy's uniform_filter1d does serially.  Squares are rounded to float32 first, like x ** 2.
#include "common.cuh"

namespace asw {

constexpr int kPowThreads = 512;
constexpr int kPowWarps = kPowThreads / 32;
constexpr int kPowItems = 4;      // window starts per thread and tile

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block, result broadcast to every thread.  `red` holds kPowWarps doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kPowWarps; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(kPowThreads)
patch_powers_kernel(float* __restrict__ x_all, int T, int W, int demean, float* __restrict__ mean_out,
                    float* __restrict__ power_out, float* __restrict__ maxavg_out, int* __restrict__ argmax_out) {
    __shared__ double red[kPowWarps];
    __shared__ double best_v[kPowWarps];
    __shared__ int best_i[kPowWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* x = x_all + (size_t)blockIdx.x * T;

    double s = 0.0;
    for (int t = tid; t < T; t += kPowThreads) s += (double)x[t];
    const float mean = (float)(block_sum(s, red) / (double)T);

    const int head_len = W < T ? W : T;
    double tot = 0.0, head = 0.0;
    for (int t = tid; t < T; t += kPowThreads) {
        const float xm = x[t] - mean;
        const double q = (double)__fmul_rn(xm, xm);
        tot += q;
        if (t < head_len) head += q;
    }
    tot = block_sum(tot, red);
    head = block_sum(head, red);

    // Window starts 0 .. L-1; beyond T-W the zero padding only removes terms, so the maximum is not there.
    const int L = T > W ? T - W + 1 : 1;
    double carry = head;       // S[i0]
    double bv = -1.0;
    int bi = 0;
    for (int i0 = 0; i0 < L; i0 += kPowThreads * kPowItems) {
        const int i = i0 + tid * kPowItems;          // this thread's kPowItems consecutive window starts
        double d[kPowItems];
        double inc = 0.0;
#pragma unroll
        for (int k = 0; k < kPowItems; ++k) {
            d[k] = 0.0;
            if (i + k + W < T) {
                const float a = x[i + k + W] - mean, b = x[i + k] - mean;
                d[k] = (double)__fmul_rn(a, a) - (double)__fmul_rn(b, b);
            }
            inc += d[k];
        }
        const double mine = inc;                     // inclusive scan of the per-thread totals over the tile
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        __syncthreads();
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        double base = 0.0, tile = 0.0;
#pragma unroll
        for (int w = 0; w < kPowWarps; ++w) {
            const double r = red[w];
            if (w < warp) base += r;
            tile += r;
        }
        double Si = carry + base + (inc - mine);
#pragma unroll
        for (int k = 0; k < kPowItems; ++k) {
            if (i + k < L && Si > bv) { bv = Si; bi = i + k; }   // ascending starts per thread: keeps the first maximum
            Si += d[k];
        }
        carry += tile;
    }
    // arg-max over the block, lowest index on ties (np.argmax)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { best_v[warp] = bv; best_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kPowWarps; ++w)
            if (best_v[w] > bv || (best_v[w] == bv && best_i[w] < bi)) { bv = best_v[w]; bi = best_i[w]; }
        mean_out[blockIdx.x] = mean;
        power_out[blockIdx.x] = (float)tot;
        maxavg_out[blockIdx.x] = sqrtf(fabsf((float)(bv / (double)W)));
        if (argmax_out) argmax_out[blockIdx.x] = bi;
    }
    if (demean) {
        __syncthreads();
        for (int t = tid; t < T; t += kPowThreads) x[t] = x[t] - mean;
    }
}

int launch_patch_powers(float* x, int N, int T, int W, int demean, float* mean_out, float* power_out,
                        float* maxavg_out, int* argmax_out, cudaStream_t s) {
    if (N == 0) return ASW_OK;
    patch_powers_kernel<<<N, kPowThreads, 0, s>>>(x, T, W, demean, mean_out, power_out, maxavg_out, argmax_out);
    ASW_LAUNCH_CHECK("patch_powers_kernel");
    return ASW_OK;
}

}  // namespace asw
