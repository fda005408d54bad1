// shift_stack.cu -- per-hypercube circular channel shift, stacked into the separator's input layout.
//
// Reference: the loop of DataParallelSpotModel.shift_and_sep (sep/training/JointModel/network.py:75-83)
// with roll_by_gather (:12-25):  data[n] = gather(mix, 1, (t - shifts) mod T), shifts = -round([0,*off]),
// i.e.   out[n][c][t] = mix[c][(t + r[n][c]) mod T].
// The reference builds an int64 (M, T) index tensor per patch and calls torch.gather; here one
// launch writes the whole (N, M, T) batch.  The kernel is HBM-write bound: 4*N*M*T bytes out, the
// (M, T) source stays L2-resident.  Each thread produces 16 B per store from two aligned 16 B loads
// (the source is misaligned by r mod 4) and only the one vector per row that straddles the wrap
// point takes the scalar path.
//
// The fused variant also applies normalize_input (sep/training/SpeakerLocalization/network.py:28-40):
// x = round(x * 2^15) / 2^15; ref = mean over mics; (x - mean_t(ref)) / std_t(ref) (unbiased).
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace asw {
namespace {

constexpr int kThreads = 256;
constexpr int kVpt = 4;  // 16-byte vectors per thread

__device__ __forceinline__ float quant16(float x) { return rintf(x * 32768.f) * (1.f / 32768.f); }

// round(x 2^15) without the quarter-rate FRND: for |x| < 128 the sum x 2^15 + 1.5 2^23 lands where the float spacing
// is 1, so the FMA's round-to-nearest-even IS rintf(x 2^15) (x 2^15 itself is exact), and subtracting the constant
// is exact.  (r 2^-15 - mean) is one more FMA with a single rounding: the same bits as fl(q - mean), q = r 2^-15.
constexpr float kRoundMagic = 12582912.f;
__device__ __forceinline__ float round15_small(float x) { return fmaf(x, 32768.f, kRoundMagic) - kRoundMagic; }
__device__ __forceinline__ float4 norm4(float4 x, float mean, float inv_sd) {
    const float big = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
    float4 r;
    if (big < 128.f) {
        r = make_float4(round15_small(x.x), round15_small(x.y), round15_small(x.z), round15_small(x.w));
    } else {   // also NaN
        r = make_float4(rintf(x.x * 32768.f), rintf(x.y * 32768.f), rintf(x.z * 32768.f), rintf(x.w * 32768.f));
    }
    return make_float4(fmaf(r.x, 1.f / 32768.f, -mean) * inv_sd, fmaf(r.y, 1.f / 32768.f, -mean) * inv_sd,
                       fmaf(r.z, 1.f / 32768.f, -mean) * inv_sd, fmaf(r.w, 1.f / 32768.f, -mean) * inv_sd);
}

__device__ __forceinline__ int reduce_shift(int r, int T) {
    // python-style r mod T; the common case |r| < T needs no division (the XU pipe was the limiter)
    if (r <= -T || r >= T) r %= T;
    return r < 0 ? r + T : r;
}

// 4 consecutive source samples starting at circular position s = base + O (0 <= s < T, T % 4 == 0,
// O = s & 3 is the same for a whole row because T and the output position are multiples of 4).
template <int O>
__device__ __forceinline__ float4 load4_circ(const float* __restrict__ src, int s, int T) {
    if (O == 0) return __ldg(reinterpret_cast<const float4*>(src + s));   // aligned, never straddles
    const int base = s - O;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + base));
    const int nb = (base + 4 < T) ? base + 4 : 0;                          // wrap of the second vector
    const float4 c = __ldg(reinterpret_cast<const float4*>(src + nb));
    if (O == 1) return make_float4(a.y, a.z, a.w, c.x);
    if (O == 2) return make_float4(a.z, a.w, c.x, c.y);
    return make_float4(a.w, c.x, c.y, c.z);
}

template <bool NORM, int O>
__device__ __forceinline__ void shift_row_chunk(const float* __restrict__ src, float4* __restrict__ dst, int r, int T,
                                                float mean, float sd) {
    const int T4 = T >> 2;
    const int v0 = blockIdx.x * (kThreads * kVpt) + threadIdx.x;
    float4 val[kVpt];
#pragma unroll
    for (int v = 0; v < kVpt; ++v) {
        const int t4 = v0 + v * kThreads;
        if (t4 < T4) {
            int s = 4 * t4 + r;
            if (s >= T) s -= T;
            val[v] = load4_circ<O>(src, s, T);
        }
    }
#pragma unroll
    for (int v = 0; v < kVpt; ++v) {
        const int t4 = v0 + v * kThreads;
        if (t4 < T4) {
            float4 x = val[v];
            // sd holds 1 / std here: one multiply instead of an IEEE division per sample (<= 1.5 ulp)
            if (NORM) x = norm4(x, mean, sd);
            __stcs(dst + t4, x);
        }
    }
}

// grid = (chunks of the row, mic c, patch n): no integer division anywhere.
template <bool NORM>
__global__ void __launch_bounds__(kThreads) shift_stack_vec_kernel(const float* __restrict__ mix,
                                                                    const int32_t* __restrict__ shifts,
                                                                    const int32_t* __restrict__ mix_index, int B, int M,
                                                                    int T, float* __restrict__ out,
                                                                    const double* __restrict__ work,
                                                                    float* __restrict__ means, float* __restrict__ stds,
                                                                    const int32_t* __restrict__ n_valid, int n_base) {
    const int c = blockIdx.y, n = blockIdx.z;
    const int row = n * M + c;
    // A CTA lives ~2.4 us: the three table reads are issued together (the table is allocated to its capacity, so
    // reading a row beyond the count is harmless) instead of waiting for the count before fetching the rest.
    const int nv = n_valid ? __ldg(n_valid) : INT_MAX;
    const int mi = mix_index ? __ldg(mix_index + n) : 0;
    const int r_raw = __ldg(shifts + row);
    if (n_base + n >= nv) return;
    if (mi < 0 || mi >= B) return;                         // stale / foreign table row: never read outside mix
    const float* src = mix + ((size_t)mi * M + c) * (size_t)T;
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * T);
    const int r = reduce_shift(r_raw, T);
    float mean = 0.f, sd = 1.f;
    if (NORM) {   // {mean, 1 / std} finalised once per patch by the statistics kernel (no fp64 divide / sqrt per CTA)
        const float2 ms = __ldg(reinterpret_cast<const float2*>(work + 2 * n));
        mean = ms.x;
        sd = ms.y;                                        // shift_row_chunk multiplies
    }
    switch (r & 3) {
        case 0: shift_row_chunk<NORM, 0>(src, dst, r, T, mean, sd); break;
        case 1: shift_row_chunk<NORM, 1>(src, dst, r, T, mean, sd); break;
        case 2: shift_row_chunk<NORM, 2>(src, dst, r, T, mean, sd); break;
        default: shift_row_chunk<NORM, 3>(src, dst, r, T, mean, sd); break;
    }
}

// Persistent variant (ASW_STACK=persist, opt-in; the default stays the kernel above): ONE resident CTA per SM, 16 warps
// x 64 registers = half of the register file, a quarter of the thread slots, no shared memory.  It was written to let
// scoring CTAs of another stream share the SMs with the HBM-bound copy; what the B200 measurements say
// (profiles/r02_notes.md section 8, experiments/coresidency_probe.py):
//  * alone it writes as fast as 6 x 256 threads per SM (0.779 vs 0.794 ms per 1152 patches): the bytes in flight, not
//    the thread count, are what saturates HBM;
//  * a grid larger than the machine never shares SMs with another grid (CTAs of the second grid are dispatched only
//    once the first is fully dispatched), a persistent grid does, but only when every co-resident kernel asks for the
//    SAME shared-memory carve-out (ASW_CARVE): an SM changes its L1 / shared split only when empty;
//  * the loads in flight live in L1 lines, so the carve-out the STFT needs (164 KB) costs this kernel 7 % of its
//    bandwidth, and the STFT next to a saturated HBM pipe runs at a third of its speed: the pair gains 0.13 ms of
//    1.36 ms, the pipelined step nothing.  Hence opt-in.
// A warp owns a contiguous run of 32 * VPT output vectors; lane l loads the ALIGNED source vectors
// src4[(run + 32 v + l + r / 4) mod T/4] (one LDG.128 per output vector instead of two), and the r mod 4 samples it
// misses come from its neighbour's registers by shuffle (lane 31 takes them from lane 0's next vector; one broadcast
// load supplies the vector after the run).  4 * (VPT + 1) data registers hold 16 * VPT bytes in flight per thread.
// Warps fetch work items (4 consecutive runs of one row) from an atomic counter, one item ahead of their use.
__device__ __forceinline__ float4 ldg_nc_v4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int O>
__device__ __forceinline__ float4 funnel_next(float4 a, float4 nxt_for_lane0, int lane) {
    if (O == 0) return a;
    const int from = (lane + 1) & 31;
    const bool l0 = lane == 0;
    const float cx = __shfl_sync(0xffffffffu, l0 ? nxt_for_lane0.x : a.x, from);
    if (O == 1) return make_float4(a.y, a.z, a.w, cx);
    const float cy = __shfl_sync(0xffffffffu, l0 ? nxt_for_lane0.y : a.y, from);
    if (O == 2) return make_float4(a.z, a.w, cx, cy);
    const float cz = __shfl_sync(0xffffffffu, l0 ? nxt_for_lane0.z : a.z, from);
    return make_float4(a.w, cx, cy, cz);
}

// A run that neither wraps around the end of the source row nor passes the end of the output row: every address is
// one base plus an immediate.  p = src4 + s0 + lane, q = dst4 + run0 + lane.
template <bool NORM, int O, int VPT>
__device__ __forceinline__ void shift_run_fast(const float4* __restrict__ p, float4* __restrict__ q, int lane, float mean,
                                               float sd) {
    float4 a[VPT + 1];
    // volatile asm keeps every load of the run ahead of the first store (ptxas otherwise interleaves them to save
    // registers, and the bytes in flight per thread are the point of this kernel)
#pragma unroll
    for (int v = 0; v < VPT; ++v) a[v] = ldg_nc_v4(p + 32 * v);
    if (O != 0) a[VPT] = ldg_nc_v4(p + 32 * VPT - lane);      // the vector after the run (lane 31's neighbour): broadcast
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        float4 x = funnel_next<O>(a[v], a[v + 1], lane);
        if (NORM) x = norm4(x, mean, sd);
        __stcs(q + 32 * v, x);
    }
}

// The two runs per row that do wrap or end early (of ~100): one vector at a time, two aligned loads each.
template <bool NORM>
__device__ __noinline__ void shift_run_slow(const float4* __restrict__ src4, float4* __restrict__ dst4, int run0, int n_vec,
                                            int rq, int o, int T4, float mean, float sd) {
    const int lane = threadIdx.x & 31;
    for (int t4 = run0 + lane; t4 < run0 + n_vec && t4 < T4; t4 += 32) {
        int s = t4 + rq;
        if (s >= T4) s -= T4;
        const float4 a = __ldg(src4 + s);
        const float4 c = __ldg(src4 + (s + 1 == T4 ? 0 : s + 1));
        float4 x = a;
        if (o == 1) x = make_float4(a.y, a.z, a.w, c.x);
        if (o == 2) x = make_float4(a.z, a.w, c.x, c.y);
        if (o == 3) x = make_float4(a.w, c.x, c.y, c.z);
        if (NORM) x = norm4(x, mean, sd);
        __stcs(dst4 + t4, x);
    }
}

// Work counters of the persistent kernel: {next item, finished CTAs} per launch slot; the last CTA to finish clears
// its slot, the host hands slots out round-robin (64 launches of one device can be in flight at a time).
constexpr int kCtrSlots = 64;
__device__ unsigned int g_stack_ctr[kCtrSlots][2];
constexpr int kSegRuns = 4;            // consecutive runs of one row per work item (24 KB at VPT = 12)

// what is fetched one item ahead of its use (two dependent table reads behind the atomic)
struct RowMeta {
    int mi, r_raw;
    float mean, sd;
};

template <bool NORM, int THREADS, int VPT>
__global__ void __maxnreg__(64) shift_stack_persist_kernel(
    const float* __restrict__ mix, const int32_t* __restrict__ shifts, const int32_t* __restrict__ mix_index, int B, int M,
    int T, float* __restrict__ out, const double* __restrict__ work, const int32_t* __restrict__ n_valid, int n_base, int N,
    int slot) {
    const int lane = threadIdx.x & 31;
    const int T4 = T >> 2;
    const int runs_per_row = (T4 + 32 * VPT - 1) / (32 * VPT);
    const int segs_per_row = (runs_per_row + kSegRuns - 1) / kSegRuns;
    int nv = N;
    if (n_valid) {
        const int left = __ldg(n_valid) - n_base;
        nv = left < nv ? left : nv;
    }
    const unsigned total = nv > 0 ? (unsigned)nv * (unsigned)M * (unsigned)segs_per_row : 0u;
    unsigned int* ctr = g_stack_ctr[slot];

    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0) v = atomicAdd(ctr, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    auto fetch = [&](unsigned item, RowMeta& m) {
        m.mi = -1;
        if (item >= total) return;
        const int row = (int)(item / (unsigned)segs_per_row);
        const int n = row / M;
        m.mi = mix_index ? __ldg(mix_index + n) : 0;
        m.r_raw = __ldg(shifts + row);
        if (NORM) {
            const float2 ms = __ldg(reinterpret_cast<const float2*>(work + 2 * n));
            m.mean = ms.x;
            m.sd = ms.y;
        } else {
            m.mean = 0.f;
            m.sd = 1.f;
        }
    };

    unsigned item = grab();
    RowMeta cur;
    fetch(item, cur);
    while (item < total) {
        const unsigned item_n = grab();
        RowMeta nxt;
        fetch(item_n, nxt);
        if (cur.mi >= 0 && cur.mi < B) {                       // stale / foreign table row: never read outside mix
            const int row = (int)(item / (unsigned)segs_per_row);
            const int seg = (int)(item - (unsigned)row * (unsigned)segs_per_row);
            const int c = row % M;
            const float4* src4 = reinterpret_cast<const float4*>(mix + ((size_t)cur.mi * M + c) * (size_t)T);
            float4* dst4 = reinterpret_cast<float4*>(out + (size_t)row * T);
            const int r = reduce_shift(cur.r_raw, T);
            const int rq = r >> 2;
            const int run_end = min(runs_per_row, (seg + 1) * kSegRuns);
            const int o = r & 3;
            for (int run = seg * kSegRuns; run < run_end; ++run) {
                const int v0 = run * (32 * VPT);
                int s0 = v0 + rq;
                if (s0 >= T4) s0 -= T4;
                if (s0 + 32 * VPT + 1 <= T4 && v0 + 32 * VPT <= T4) {
                    const float4* p = src4 + s0 + lane;
                    float4* q = dst4 + v0 + lane;
                    switch (o) {
                        case 0: shift_run_fast<NORM, 0, VPT>(p, q, lane, cur.mean, cur.sd); break;
                        case 1: shift_run_fast<NORM, 1, VPT>(p, q, lane, cur.mean, cur.sd); break;
                        case 2: shift_run_fast<NORM, 2, VPT>(p, q, lane, cur.mean, cur.sd); break;
                        default: shift_run_fast<NORM, 3, VPT>(p, q, lane, cur.mean, cur.sd); break;
                    }
                } else {
                    shift_run_slow<NORM>(src4, dst4, v0, 32 * VPT, rq, o, T4, cur.mean, cur.sd);
                }
            }
        }
        cur = nxt;
        item = item_n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
            ctr[0] = 0;
            ctr[1] = 0;
            __threadfence();
        }
    }
}

// Generic fallback for T % 4 != 0 or unaligned buffers.
template <bool NORM>
__global__ void __launch_bounds__(kThreads) shift_stack_scalar_kernel(const float* __restrict__ mix,
                                                                       const int32_t* __restrict__ shifts,
                                                                       const int32_t* __restrict__ mix_index, int B,
                                                                       int M, int T, float* __restrict__ out,
                                                                       const double* __restrict__ work,
                                                                       float* __restrict__ means,
                                                                       float* __restrict__ stds,
                                                                       const int32_t* __restrict__ n_valid, int n_base) {
    const int c = blockIdx.y, n = blockIdx.z;
    if (n_valid && n_base + n >= *n_valid) return;
    const int row = n * M + c;
    const int mi = mix_index ? mix_index[n] : 0;
    if (mi < 0 || mi >= B) return;
    const float* src = mix + ((size_t)mi * M + c) * (size_t)T;
    float* dst = out + (size_t)row * T;
    const int r = reduce_shift(shifts[row], T);
    float mean = 0.f, sd = 1.f;
    if (NORM) {
        const float2 ms = *reinterpret_cast<const float2*>(work + 2 * n);
        mean = ms.x;
        sd = ms.y;                                        // 1 / std
    }
    for (int t = blockIdx.x * (kThreads * kVpt * 4) + threadIdx.x, e = min(T, (int)(blockIdx.x + 1) * (kThreads * kVpt * 4));
         t < e; t += kThreads) {
        int s = t + r;
        if (s >= T) s -= T;
        float x = __ldg(src + s);
        if (NORM) x = (quant16(x) - mean) * sd;
        dst[t] = x;
    }
}

// Pass 1 of the fused variant: per patch, sum and sum of squares over t of the mic-averaged,
// re-quantised, shifted signal (accumulated in double).
// Vector path (T % 4 == 0, 16-byte aligned rows): a thread takes 4 consecutive samples of every mic with two aligned
// 16-byte loads and the same register funnel as the stack kernel; the per-sample arithmetic (mic order of the float
// sum, the 1/M scale) is unchanged.
template <int O>
__device__ __forceinline__ void add_quant4(float4& ref, const float* __restrict__ row, int s, int T) {
    const float4 v = load4_circ<O>(row, s, T);
    ref.x += quant16(v.x);
    ref.y += quant16(v.y);
    ref.z += quant16(v.z);
    ref.w += quant16(v.w);
}

// One CTA per patch and a fixed reduction order (thread -> warp -> CTA), so the statistics -- and with them every
// normalised sample -- are reproducible bit for bit from run to run; atomics across CTAs would not be.
constexpr int kStatThreads = 1024;

// Both moments -> what the stack pass needs: means[n], stds[n] for the caller and {mean, 1 / std} as two floats at
// the start of the patch's work slot.
__device__ __forceinline__ void finalise_stats(double S, double SS, int T, int n, double* __restrict__ work,
                                               float* __restrict__ means, float* __restrict__ stds) {
    const double mu = S / (double)T;
    const double var = (SS - S * mu) / (double)(T - 1);
    const float sd = (float)sqrt(var > 0.0 ? var : 0.0);
    const float mean = (float)mu;
    means[n] = mean;
    stds[n] = sd;
    *reinterpret_cast<float2*>(work + 2 * n) = make_float2(mean, 1.0f / sd);
}

// With `tables` (asw_corr_tables, xcorr.cu) the two moments come from M + M + P look-ups per patch:
//     sum ref = 1/M sum_c S_c,   sum ref^2 = 1/M^2 (sum_c E_c + 2 sum_{c<c'} R_cc'(r_c' - r_c)),
// and the CTA returns before touching the audio.  A patch whose lag falls outside the table, or whose centred second
// moment is small against the fp32 round-off of the table entries (bounded through (sum_c sqrt E_c)^2 >= M^2 sum ref^2),
// takes the exact pass below instead, so the result never depends on the table's range.
constexpr double kTableCondition = 0.02;

template <bool VEC>
__global__ void __launch_bounds__(kStatThreads) shift_ref_stats_kernel(const float* __restrict__ mix,
                                                                        const int32_t* __restrict__ shifts,
                                                                        const int32_t* __restrict__ mix_index, int B, int M,
                                                                        int T, double* __restrict__ work,
                                                                        float* __restrict__ means,
                                                                        float* __restrict__ stds,
                                                                        const double* __restrict__ tables,
                                                                        int table_stride, int L,
                                                                        const int32_t* __restrict__ n_valid, int n_base) {
    __shared__ int s_r[kMaxMics];
    __shared__ double s_red[2][kStatThreads / 32];
    __shared__ int s_done;
    const int n = blockIdx.x;
    if (n_valid && n_base + n >= *n_valid) return;             // beyond the device-resident patch count
    const int mi = mix_index ? mix_index[n] : 0;
    if (mi < 0 || mi >= B) {                                   // stale table row: leave a recognisable result, read nothing
        if (threadIdx.x == 0) {
            means[n] = 0.f;
            stds[n] = 0.f;
            *reinterpret_cast<float2*>(work + 2 * n) = make_float2(0.f, 0.f);
        }
        return;
    }
    const float* src = mix + (size_t)mi * M * (size_t)T;
    if (threadIdx.x < M) s_r[threadIdx.x] = reduce_shift(shifts[n * M + threadIdx.x], T);
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    if (tables) {
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const double* tb = tables + (size_t)mi * table_stride;
            const int P = M * (M - 1) / 2;
            double sS = 0.0, sE = 0.0, sA = 0.0, sR = 0.0;
            int bad = 0;
            if (lane < M) {
                sS = tb[lane];
                sE = tb[M + lane];
                sA = sqrt(sE);
            }
            for (int p = lane; p < P; p += 32) {                 // pair p -> (i, j), i < j row-major (M <= 32: <= 31 steps)
                int i = 0, rem = p;
                while (rem >= M - 1 - i) {
                    rem -= M - 1 - i;
                    ++i;
                }
                const int j = i + 1 + rem;
                int d = s_r[j] - s_r[i];                         // both in [0, T)
                if (2 * d > T) d -= T;
                if (2 * d < -T) d += T;
                if (d < -L || d > L) bad = 1;
                else sR += tb[2 * M + (size_t)p * (2 * L + 1) + (d + L)];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sS += __shfl_xor_sync(0xffffffffu, sS, o);
                sE += __shfl_xor_sync(0xffffffffu, sE, o);
                sA += __shfl_xor_sync(0xffffffffu, sA, o);
                sR += __shfl_xor_sync(0xffffffffu, sR, o);
                bad |= __shfl_xor_sync(0xffffffffu, bad, o);
            }
            const double inv_m = 1.0 / (double)M;
            const double S = sS * inv_m, SS = (sE + 2.0 * sR) * inv_m * inv_m;
            const double centred = SS - S * S / (double)T;
            const double bound = sA * sA * inv_m * inv_m;
            if (!bad && centred > kTableCondition * bound) {
                if (lane == 0) {
                    finalise_stats(S, SS, T, n, work, means, stds);
                    s_done = 1;
                }
            }
        }
        __syncthreads();
        if (s_done) return;
    }
    const float inv_m = 1.f / (float)M;
    double sum = 0.0, sq = 0.0;
    if (VEC) {
        for (int t = 4 * threadIdx.x; t < T; t += 4 * kStatThreads) {
            float4 ref = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < M; ++c) {
                int s = t + s_r[c];
                if (s >= T) s -= T;
                const float* row = src + (size_t)c * T;
                switch (s & 3) {                              // uniform over the CTA: s_r[c] is, t is a multiple of 4
                    case 0: add_quant4<0>(ref, row, s, T); break;
                    case 1: add_quant4<1>(ref, row, s, T); break;
                    case 2: add_quant4<2>(ref, row, s, T); break;
                    default: add_quant4<3>(ref, row, s, T); break;
                }
            }
            const float r4[4] = {ref.x * inv_m, ref.y * inv_m, ref.z * inv_m, ref.w * inv_m};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sum += (double)r4[k];
                sq += (double)r4[k] * (double)r4[k];
            }
        }
    } else {
        for (int t = threadIdx.x; t < T; t += kStatThreads) {
            float ref = 0.f;
            for (int c = 0; c < M; ++c) {
                int s = t + s_r[c];
                if (s >= T) s -= T;
                ref += quant16(__ldg(src + (size_t)c * T + s));
            }
            ref *= inv_m;
            sum += (double)ref;
            sq += (double)ref * (double)ref;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, d);
        sq += __shfl_down_sync(0xffffffffu, sq, d);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_red[0][warp] = sum;
        s_red[1][warp] = sq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < kStatThreads / 32; ++i) {
            a += s_red[0][i];
            c += s_red[1][i];
        }
        finalise_stats(a, c, T, n, work, means, stds);
    }
}

// ---- statistics of a table whose rows are grouped by mixture ------------------------------------------------------------
// The coarse stage stacks ~35 patches per mixture: the per-mixture correlation tables cost as much as the exact pass
// they replace (13 - 18 us per mixture), and the exact pass reads a mixture's M T samples once PER PATCH through L2.
// Here a CTA owns (mixture, time tile): it quantises the tile of all channels (plus a halo of kGrpHalo samples) into
// shared memory ONCE and serves every patch of the mixture from it.  Both moments are sums of integers --
//     s[t] = sum_c k_c[(t + d_c) mod T],  k = rint(x 2^15),  d_c = r_c - r_0 (a common shift leaves a circular sum alone),
//     sum ref = sum_t s / (M 2^15),   sum ref^2 = sum_t s^2 / (M 2^15)^2  (s^2 < 2^38, the total < 2^56: exact in int64)
// so any order of accumulation (warp reductions, shared and global atomics) gives the same bits, and the statistics are
// those of the unrounded mic average (the per-patch pass rounds ref = fl(sum / M) to float first: 1e-8 relative apart).
constexpr int kGrpThreads = 512, kGrpHalo = 512, kGrpTile = 2048, kGrpChunk = 128, kGrpMaxM = 8;

__global__ void group_ranges_kernel(const int32_t* __restrict__ mix_index, const int32_t* __restrict__ n_valid, int n_base,
                                    int N, int B, int32_t* __restrict__ start, unsigned long long* __restrict__ acc) {
    const int nv = n_valid ? max(0, min(N, *n_valid - n_base)) : N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) {
        acc[2 * i] = 0ull;
        acc[2 * i + 1] = 0ull;
    }
    if (i <= B) {                       // first row whose mixture index is >= i (rows are non-decreasing in mixture)
        int lo = 0, hi = nv;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (mix_index[mid] < i) lo = mid + 1;
            else hi = mid;
        }
        start[i] = lo;
    }
}

template <int M>
__global__ void __launch_bounds__(kGrpThreads, 2) shift_stats_grouped_kernel(const float* __restrict__ mix,
                                                                            const int32_t* __restrict__ shifts,
                                                                            const int32_t* __restrict__ start, int T,
                                                                            unsigned long long* __restrict__ acc) {
    extern __shared__ int s_q[];                               // [M][kGrpTile + 2 kGrpHalo] quantised samples
    __shared__ int s_d[kGrpChunk][M];                          // relative shifts of the chunk's patches
    __shared__ unsigned s_acc[kGrpChunk][4];                  // CTA sums: sum s (two's complement) and three 20-bit limbs of sum s^2
    __shared__ int s_far[kGrpChunk];                           // some |d_c| exceeds the halo: read that patch from global
    constexpr int kW = kGrpTile + 2 * kGrpHalo;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int n0 = start[b], n1 = start[b + 1];
    if (n0 >= n1) return;
    const int t0 = blockIdx.x * kGrpTile;
    const int len = min(kGrpTile, T - t0);
    const float* src = mix + (size_t)b * M * (size_t)T;
    if (T >= kW) {                                             // the tile wraps at most once: no integer division
        for (int c = 0; c < M; ++c) {
            const float* row = src + (size_t)c * T;
            int* dst = s_q + c * kW;
#pragma unroll 6
            for (int i = tid; i < len + 2 * kGrpHalo; i += kGrpThreads) {
                int tt = t0 - kGrpHalo + i;
                tt += tt < 0 ? T : 0;
                tt -= tt >= T ? T : 0;
                dst[i] = __float2int_rn(__ldg(row + tt) * 32768.f);
            }
        }
    } else {
        for (int c = 0; c < M; ++c)
            for (int i = tid; i < len + 2 * kGrpHalo; i += kGrpThreads) {
                int tt = (t0 - kGrpHalo + i) % T;
                if (tt < 0) tt += T;
                s_q[c * kW + i] = __float2int_rn(__ldg(src + (size_t)c * T + tt) * 32768.f);
            }
    }
    for (int c0 = n0; c0 < n1; c0 += kGrpChunk) {
        const int np = min(kGrpChunk, n1 - c0);
        __syncthreads();                                       // tile ready / previous chunk flushed
        for (int i = tid; i < np; i += kGrpThreads) {
            const int32_t* sh = shifts + (size_t)(c0 + i) * M;
            const int r0 = reduce_shift(sh[0], T);
            int far = 0;
#pragma unroll
            for (int c = 0; c < M; ++c) {
                int d = reduce_shift(sh[c], T) - r0;           // in (-T, T)
                if (2 * d > T) d -= T;
                if (2 * d < -T) d += T;
                s_d[i][c] = d;
                far |= (d < -kGrpHalo || d > kGrpHalo);
            }
            s_far[i] = far;
            s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0u;
        }
        __syncthreads();
        for (int pi = 0; pi < np; ++pi) {
            int a1 = 0;
            long long a2 = 0;
            if (!s_far[pi]) {
                int base[M];
#pragma unroll
                for (int c = 0; c < M; ++c) base[c] = c * kW + kGrpHalo + s_d[pi][c] + tid;
#pragma unroll
                for (int k = 0; k < kGrpTile / kGrpThreads; ++k) {
                    if (tid + k * kGrpThreads < len) {
                        int sum = 0;
#pragma unroll
                        for (int c = 0; c < M; ++c) sum += s_q[base[c] + k * kGrpThreads];
                        a1 += sum;
                        a2 += (long long)sum * sum;
                    }
                }
            } else {
                for (int k = 0; k < kGrpTile / kGrpThreads; ++k) {
                    const int i = tid + k * kGrpThreads;
                    if (i < len) {
                        int sum = 0;
                        for (int c = 0; c < M; ++c) {
                            int tt = (t0 + i + s_d[pi][c]) % T;
                            if (tt < 0) tt += T;
                            sum += __float2int_rn(__ldg(src + (size_t)c * T + tt) * 32768.f);
                        }
                        a1 += sum;
                        a2 += (long long)sum * sum;
                    }
                }
            }
            // warp sums: |a1| < 2^22 per thread; a2 < 2^40 per thread, split into 20-bit limbs for the 32-bit redux
            const int w1 = __reduce_add_sync(0xffffffffu, a1);
            const unsigned long long u2 = (unsigned long long)a2;
            const unsigned l0 = __reduce_add_sync(0xffffffffu, (unsigned)(u2 & 0xfffffull));
            const unsigned l1 = __reduce_add_sync(0xffffffffu, (unsigned)((u2 >> 20) & 0xfffffull));
            const unsigned l2 = __reduce_add_sync(0xffffffffu, (unsigned)(u2 >> 40));
            if ((tid & 31) == 0) {        // native 32-bit shared atomics (a 64-bit shared atomicAdd is a CAS loop): the CTA
                atomicAdd(&s_acc[pi][0], (unsigned)w1);   // sums stay below 2^31: |sum s| < 2^29, limbs < 16 warps x 2^25
                atomicAdd(&s_acc[pi][1], l0);
                atomicAdd(&s_acc[pi][2], l1);
                atomicAdd(&s_acc[pi][3], l2);
            }
        }
        __syncthreads();
        for (int i = tid; i < np; i += kGrpThreads) {
            atomicAdd(acc + 2 * (size_t)(c0 + i), (unsigned long long)(long long)(int)s_acc[i][0]);
            atomicAdd(acc + 2 * (size_t)(c0 + i) + 1, (unsigned long long)s_acc[i][1] + ((unsigned long long)s_acc[i][2] << 20) +
                                                          ((unsigned long long)s_acc[i][3] << 40));
        }
    }
}

__global__ void group_finalise_kernel(const int32_t* __restrict__ mix_index, const int32_t* __restrict__ n_valid, int n_base,
                                      int N, int B, int M, int T, double* __restrict__ work, float* __restrict__ means,
                                      float* __restrict__ stds) {
    const int nv = n_valid ? max(0, min(N, *n_valid - n_base)) : N;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nv) return;
    const int mi = mix_index[n];
    if (mi < 0 || mi >= B) {                                   // stale table row: the per-patch pass's recognisable result
        means[n] = 0.f;
        stds[n] = 0.f;
        *reinterpret_cast<float2*>(work + 2 * n) = make_float2(0.f, 0.f);
        return;
    }
    const unsigned long long* acc = reinterpret_cast<const unsigned long long*>(work);
    const long long a1 = (long long)acc[2 * n];
    const unsigned long long a2 = acc[2 * n + 1];
    const double scale = 1.0 / (32768.0 * (double)M);
    finalise_stats((double)a1 * scale, (double)a2 * scale * scale, T, n, work, means, stds);
}

template <int M>
int launch_grouped_t(const float* mix, const int32_t* shifts, const int32_t* start, int B, int T, double* work,
                     cudaStream_t s) {
    const size_t smem = (size_t)M * (kGrpTile + 2 * kGrpHalo) * sizeof(int);
    static PerDeviceOnce attr_once;
    if (attr_once.need())
        ASW_CUDA_CHECK(cudaFuncSetAttribute(shift_stats_grouped_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((T + kGrpTile - 1) / kGrpTile, B);
    shift_stats_grouped_kernel<M><<<grid, kGrpThreads, smem, s>>>(mix, shifts, start, T,
                                                                  reinterpret_cast<unsigned long long*>(work));
    ASW_LAUNCH_CHECK("shift_stats_grouped_kernel");
    return ASW_OK;
}

// int16 PCM -> float32 in [-1, 1): x / 32768, the conversion soundfile/librosa apply when the reference loads
// its PCM_16 wav files (sep/helpers/utils.py read_audio_file); exact in float32.
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out,
                                                           size_t n) {
    const size_t n8 = n >> 3;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
        const short* s = reinterpret_cast<const short*>(&v);
        float4 a, b;
        a.x = (float)s[0] * (1.f / 32768.f);
        a.y = (float)s[1] * (1.f / 32768.f);
        a.z = (float)s[2] * (1.f / 32768.f);
        a.w = (float)s[3] * (1.f / 32768.f);
        b.x = (float)s[4] * (1.f / 32768.f);
        b.y = (float)s[5] * (1.f / 32768.f);
        b.z = (float)s[6] * (1.f / 32768.f);
        b.w = (float)s[7] * (1.f / 32768.f);
        reinterpret_cast<float4*>(out)[2 * i] = a;
        reinterpret_cast<float4*>(out)[2 * i + 1] = b;
    }
    for (size_t i = (n8 << 3) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (float)in[i] * (1.f / 32768.f);
}

constexpr int kPersistThreads = 512, kPersistVpt = 12;   // 16 warps x 64 registers = half of an SM's register file
// resident CTAs of the persistent kernel (0 = the 256-thread kernel with one CTA per 16 KB of a row)
int stack_persist_ctas() {
    static const int v = [] {
        const char* e = getenv("ASW_STACK");
        if (!e || strncmp(e, "persist", 7) != 0) return 0;
        const int n = e[7] == ':' ? atoi(e + 8) : 0;
        return n > 0 ? n : kNumSms;
    }();
    return v;
}

template <bool NORM>
int launch_rows(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M, int T, float* out,
                const double* work, float* means, float* stds, cudaStream_t s, const int32_t* n_valid = nullptr,
                int n_base = 0) {
    const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(mix) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int per_cta = kThreads * kVpt * 4;  // samples per CTA
    for (int n0 = 0; n0 < N; n0 += 65535) {   // gridDim.z limit
        const int nn = (N - n0 < 65535) ? N - n0 : 65535;
        dim3 grid((unsigned)((T + per_cta - 1) / per_cta), (unsigned)M, (unsigned)nn);
        const int32_t* sh = shifts + (size_t)n0 * M;
        const int32_t* mi = mix_index ? mix_index + n0 : nullptr;
        float* o = out + (size_t)n0 * M * T;
        const double* wk = work ? work + 2 * (size_t)n0 : nullptr;
        float* mu = means ? means + n0 : nullptr;
        float* sd = stds ? stds + n0 : nullptr;
        const int pctas = vec ? stack_persist_ctas() : 0;
        constexpr int pvpt = kPersistVpt;
        const long long items = (long long)nn * M * (((T / 4 + 32 * pvpt - 1) / (32 * pvpt) + kSegRuns - 1) / kSegRuns);
        if (pctas > 0 && T / 4 >= 32 * pvpt && items < (1ll << 31)) {
            // ASW_STACK=persist[:ctas]: one resident CTA per SM that leaves half of the register file to other streams
            static std::atomic<unsigned> next_slot{0};
            const int slot = (int)(next_slot.fetch_add(1) % kCtrSlots);
            const int grid = (int)(items < pctas ? items : pctas);
ASW_CARVE_ONCE((shift_stack_persist_kernel<NORM, kPersistThreads, kPersistVpt>));
            shift_stack_persist_kernel<NORM, kPersistThreads, kPersistVpt><<<grid, kPersistThreads, 0, s>>>(
                mix, sh, mi, B, M, T, o, wk, n_valid, n_base + n0, nn, slot);
            ASW_LAUNCH_CHECK("shift_stack_persist_kernel");
        } else if (vec) {
            ASW_CARVE_ONCE(shift_stack_vec_kernel<NORM>);
            shift_stack_vec_kernel<NORM><<<grid, kThreads, 0, s>>>(mix, sh, mi, B, M, T, o, wk, mu, sd, n_valid, n_base + n0);
            ASW_LAUNCH_CHECK("shift_stack_vec_kernel");
        } else {
            shift_stack_scalar_kernel<NORM><<<grid, kThreads, 0, s>>>(mix, sh, mi, B, M, T, o, wk, mu, sd, n_valid,
                                                                      n_base + n0);
            ASW_LAUNCH_CHECK("shift_stack_scalar_kernel");
        }
    }
    return ASW_OK;
}

}  // namespace

int launch_shift_stack(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M, int T,
                       float* out, cudaStream_t s) {
    if (N == 0) return ASW_OK;
    return launch_rows<false>(mix, shifts, mix_index, N, B, M, T, out, nullptr, nullptr, nullptr, s);
}

int launch_pcm16_to_f32(const short* in, float* out, size_t n, cudaStream_t s) {
    if (n == 0) return ASW_OK;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("pcm16_to_f32: buffers must be 16-byte aligned");
        return ASW_ERR_ARG;
    }
    ASW_CARVE_ONCE(pcm16_to_f32_kernel);
    pcm16_to_f32_kernel<<<kNumSms * 8, 256, 0, s>>>(in, out, n);
    ASW_LAUNCH_CHECK("pcm16_to_f32_kernel");
    return ASW_OK;
}

int launch_shift_stack_counted(const float* mix, const int32_t* shifts, const int32_t* mix_index, const int32_t* n_valid,
                               int n_base, int N, int B, int M, int T, float* out, cudaStream_t s) {
    if (N == 0) return ASW_OK;
    return launch_rows<false>(mix, shifts + (size_t)n_base * M, mix_index + n_base, N, B, M, T, out, nullptr, nullptr,
                              nullptr, s, n_valid, n_base);
}

int launch_shift_stack_norm(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M,
                            int T, float* out, float* means, float* stds, double* work, const double* tables,
                            int table_stride, int max_lag, const int32_t* n_valid, int n_base, cudaStream_t s) {
    if (N == 0) return ASW_OK;
    shifts += (size_t)n_base * M;
    if (mix_index) mix_index += n_base;
    if ((T % 4 == 0) && ((reinterpret_cast<uintptr_t>(mix) & 15) == 0))
        shift_ref_stats_kernel<true><<<N, kStatThreads, 0, s>>>(mix, shifts, mix_index, B, M, T, work, means, stds, tables,
                                                                table_stride, max_lag, n_valid, n_base);
    else
        shift_ref_stats_kernel<false><<<N, kStatThreads, 0, s>>>(mix, shifts, mix_index, B, M, T, work, means, stds, tables,
                                                                 table_stride, max_lag, n_valid, n_base);
    ASW_LAUNCH_CHECK("shift_ref_stats_kernel");
    return launch_rows<true>(mix, shifts, mix_index, N, B, M, T, out, work, means, stds, s, n_valid, n_base);
}

// Fused shift-stack + normalize_input for a table grouped by mixture (see shift_stats_grouped_kernel); M > 8 or a
// null mix_index fall back to the per-patch pass.
int launch_shift_stack_norm_grouped(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M,
                                    int T, float* out, float* means, float* stds, double* work, int32_t* ranges,
                                    const int32_t* n_valid, int n_base, cudaStream_t s) {
    if (N == 0) return ASW_OK;
    if (M > kGrpMaxM || M < 2 || !mix_index || !ranges)
        return launch_shift_stack_norm(mix, shifts, mix_index, N, B, M, T, out, means, stds, work, nullptr, 0, 0, n_valid,
                                       n_base, s);
    const int32_t* sh = shifts + (size_t)n_base * M;
    const int32_t* mi = mix_index + n_base;
    const int rows = N > B + 1 ? N : B + 1;
    group_ranges_kernel<<<(rows + 255) / 256, 256, 0, s>>>(mi, n_valid, n_base, N, B, ranges,
                                                           reinterpret_cast<unsigned long long*>(work));
    ASW_LAUNCH_CHECK("group_ranges_kernel");
    int rc = ASW_OK;
    switch (M) {
        case 2: rc = launch_grouped_t<2>(mix, sh, ranges, B, T, work, s); break;
        case 3: rc = launch_grouped_t<3>(mix, sh, ranges, B, T, work, s); break;
        case 4: rc = launch_grouped_t<4>(mix, sh, ranges, B, T, work, s); break;
        case 5: rc = launch_grouped_t<5>(mix, sh, ranges, B, T, work, s); break;
        case 6: rc = launch_grouped_t<6>(mix, sh, ranges, B, T, work, s); break;
        case 7: rc = launch_grouped_t<7>(mix, sh, ranges, B, T, work, s); break;
        default: rc = launch_grouped_t<8>(mix, sh, ranges, B, T, work, s); break;
    }
    if (rc != ASW_OK) return rc;
    group_finalise_kernel<<<(N + 255) / 256, 256, 0, s>>>(mi, n_valid, n_base, N, B, M, T, work, means, stds);
    ASW_LAUNCH_CHECK("group_finalise_kernel");
    return launch_rows<true>(mix, sh, mi, N, B, M, T, out, work, means, stds, s, n_valid, n_base);
}

}  // namespace asw
