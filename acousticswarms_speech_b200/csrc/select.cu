// select.cu -- greedy selection of coarse hypercube patches from the SRP peaks, on the device.
//
// Reference: SRP_PHAT.local_source_adaptive (sep/Traditional_SP/SRP_Prunning.py:547-643) with its helpers
// hyperbola_area_sample (:30-39) and hyperbola_area_init / hyperbola_offset (:19-61):
//   peaks sorted by descending power (:560); for each not-yet-covered peak (:571):
//     * a width-8 cube around the peak cluster's quantised TDoA vector is trimmed against every patch
//       accepted so far (:576-602): all dimensions share the same occupied prefix, so the result is one
//       scalar `cut` = min over overlapping patches of (8 + delta1), or "discard" (:594-595, :610-612);
//     * new centre_i = round(centre_i + (cut - 8) / 2) (:615), width_i = cut (:614);
//     * every peak inside the closed box centre +- 4.1 is marked covered (:623-624);
//     * the patch is kept only if some 1 cm voxel of the room lies inside the box centre_new +- (cut+0.2)/2
//       in TDoA space: first the 5 cm volume is probed (:45-47), then the 1 cm volume restricted to the
//       bounding box of the 5 cm hits (:49-60); empty -> the patch is dropped but the covered marks stay.
// The reference spends 0.2 s of its 0.23 s here scanning the whole 5 cm volume in numpy per patch.  Device
// version: one CTA per mixture walks the peaks in order; the 5 cm volume is stored bucketed by the integer
// parts of its first two TDoA coordinates (voxel records contiguous per bucket row), so the probe touches
// ~1 % of the voxels with coalesced loads; the "any 1 cm voxel inside"
// decision first tries the 1 cm voxels that coincide with the 5 cm hits (always inside the bounding box)
// and only if none is inside scans the reference's 1 cm cut exactly.  Comparisons are in double on the
// same float64 volumes the host code uses, so decisions are identical.  The member points of a kept
// patch (`area_points`) are not produced here: the host builds them lazily for the few patches that
// survive the network stage and get subdivided.
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"

struct asw_select {
    int device = 0, G = 0, D = 0, W = 8;
    int n5 = 0, Nx5 = 0, Ny5 = 0, Ny1 = 0, Nx1 = 0, Nz = 0;
    double ax0 = 0, ax1 = 0, ay0 = 0, ay1 = 0;
    int32_t* d_cl_off = nullptr;  // [G][D]
    double* d_off5 = nullptr;     // [n5][D] voxel records ordered by bucket (floor(o0), floor(o1))
    double* d_off1s = nullptr;    // [n5][D] 1 cm offsets at the voxel coinciding with each 5 cm voxel (NaN: none)
    int32_t* d_vox5 = nullptr;    // [n5] iy * Nx5 + ix of the ordered voxels
    int32_t* d_bstart = nullptr;  // [NB0 * NB1 + 1] first voxel of every bucket
    int b0min = 0, b1min = 0, NB0 = 1, NB1 = 1;
    double* d_xx5 = nullptr;      // [Nx5]
    double* d_yy5 = nullptr;      // [Ny5]
    double* d_off1 = nullptr;     // [Ny1][Nx1][Nz][D]
    unsigned char* d_box = nullptr;  // [G] is the untrimmed box of cluster g non-empty
};

namespace asw {
namespace {

constexpr int kSelThreads = 512;
constexpr int kMaxPeaks = 1024;
constexpr int kMaxPatches = 128;
constexpr int kMaxD = 31;
constexpr int kMaxRows = 32;    // first-coordinate buckets one probe box can touch (width <= 8.2 -> 10)

struct SelectParams {
    const int32_t* peaks;   // [B][max_peaks]
    const int32_t* count;   // [B]
    const float* map;       // [B][G]
    int max_peaks, G, D, W;
    const int32_t* cl_off;
    const double* off5;
    const double* off1s;
    const int32_t* vox5;
    const int32_t* bstart;
    int b0min, b1min, NB0, NB1;
    int n5, Nx5, Ny5;
    const double* xx5;
    const double* yy5;
    double ax0, ax1, ay0, ay1;
    const double* off1;
    int Ny1, Nx1, Nz;
    const unsigned char* box_table;   // [G] untrimmed box non-empty (null while it is being built)
    int32_t* out_count;     // [B]
    int32_t* out_off;       // [B][max_patches][D]
    int32_t* out_width;     // [B][max_patches]
    int32_t* out_peak;      // [B][max_patches] cluster id of the peak the patch grew from
    int max_patches;
};

struct ProbeFlags {
    int cut, cnt5, hit1, ixmin, ixmax, iymin, iymax, pad;
};

struct ProbeShared {
    int rstart[kMaxRows], rpre[kMaxRows + 1];
    int cutbox[4];
    ProbeFlags f;
};

// "Does the closed TDoA box [lo, hi] contain a 1 cm voxel of the room?" -- hyperbola_area_init (:41-61) reduced to
// its decision: 5 cm probe over the bucket rows the box touches, the coinciding 1 cm voxels first, the exact 1 cm
// cut only if none of them is inside.  Block-wide (all threads call it, contains barriers); lo/hi are in shared
// memory and already visible.  Returns 1 = non-empty, 0 = empty / no 5 cm voxel.
__device__ int area_nonempty(const SelectParams& p, const double* s_lo, const double* s_hi, ProbeShared* ps) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    ProbeFlags* f = &ps->f;
    if (warp == 0) {
        int b0lo = max((int)floor(s_lo[0]) - p.b0min, 0);
        int b0hi = min((int)floor(s_hi[0]) - p.b0min, p.NB0 - 1);
        if (b0hi - b0lo + 1 > kMaxRows) b0hi = b0lo + kMaxRows - 1;   // cannot happen for widths <= 8.2
        int b1lo = 0, b1hi = 0;
        if (D >= 2) {
            b1lo = max((int)floor(s_lo[1]) - p.b1min, 0);
            b1hi = min((int)floor(s_hi[1]) - p.b1min, p.NB1 - 1);
        }
        int len = 0, start = 0;
        if (b0lo + lane <= b0hi && b1lo <= b1hi) {
            start = p.bstart[(b0lo + lane) * p.NB1 + b1lo];
            len = p.bstart[(b0lo + lane) * p.NB1 + b1hi + 1] - start;
        }
        int inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        ps->rstart[lane] = start;
        ps->rpre[lane] = inc - len;
        if (lane == 31) ps->rpre[32] = inc;
        if (lane == 0) {
            f->cnt5 = 0;
            f->hit1 = 0;
            f->ixmin = INT_MAX;
            f->iymin = INT_MAX;
            f->ixmax = -1;
            f->iymax = -1;
        }
    }
    __syncthreads();
    {
        int cnt = 0, ixmin = INT_MAX, ixmax = -1, iymin = INT_MAX, iymax = -1, hit = 0;
        const int total = ps->rpre[32];
        for (int c = tid; c < total; c += blockDim.x) {
            int r = 0;
            while (r < 31 && c >= ps->rpre[r + 1]) ++r;
            const int k = ps->rstart[r] + (c - ps->rpre[r]);
            const double* o = p.off5 + (size_t)k * D;
            bool in = true;
            for (int i = 0; i < D; ++i) {
                const double v = o[i];
                in = in && (v >= s_lo[i]) && (v <= s_hi[i]);
            }
            if (in) {
                const int v = p.vox5[k];
                const int iy = v / p.Nx5, ix = v - iy * p.Nx5;
                ++cnt;
                ixmin = min(ixmin, ix);
                ixmax = max(ixmax, ix);
                iymin = min(iymin, iy);
                iymax = max(iymax, iy);
                if (!hit) {                                   // one inside 1 cm voxel is all the decision needs
                    const double* o1 = p.off1s + (size_t)k * D;
                    bool in1 = true;                          // NaN (no coinciding voxel) compares false
                    for (int i = 0; i < D; ++i) {
                        const double v1 = o1[i];
                        in1 = in1 && (v1 >= s_lo[i]) && (v1 <= s_hi[i]);
                    }
                    hit = in1 ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            hit |= __shfl_xor_sync(0xffffffffu, hit, d);
            ixmin = min(ixmin, __shfl_xor_sync(0xffffffffu, ixmin, d));
            ixmax = max(ixmax, __shfl_xor_sync(0xffffffffu, ixmax, d));
            iymin = min(iymin, __shfl_xor_sync(0xffffffffu, iymin, d));
            iymax = max(iymax, __shfl_xor_sync(0xffffffffu, iymax, d));
        }
        if (lane == 0 && cnt > 0) {
            atomicAdd(&f->cnt5, cnt);
            if (hit) f->hit1 = 1;
            atomicMin(&f->ixmin, ixmin);
            atomicMax(&f->ixmax, ixmax);
            atomicMin(&f->iymin, iymin);
            atomicMax(&f->iymax, iymax);
        }
    }
    __syncthreads();
    if (f->cnt5 == 0) return 0;                               // init_area is None (:46-47)
    if (f->hit1 != 0) return 1;
    // exact scan of the reference's 1 cm cut (:49-60); rare: no coinciding 1 cm voxel was inside
    if (tid == 0) {
        double x0 = p.xx5[f->ixmin] - 0.05, x1 = p.xx5[f->ixmax] + 0.05;
        x0 = fmax(p.ax0, x0);
        x1 = fmin(p.ax1, x1);
        double y0 = p.yy5[f->iymin] - 0.05, y1 = p.yy5[f->iymax] + 0.05;
        y0 = fmax(p.ay0, y0);
        y1 = fmin(p.ay1, y1);
        const int ix0 = (int)floor((x0 - p.ax0) / 0.01), ix1 = (int)ceil((x1 - p.ax0) / 0.01);
        const int iy0 = (int)floor((y0 - p.ay0) / 0.01), iy1 = (int)ceil((y1 - p.ay0) / 0.01);
        ps->cutbox[0] = max(0, min(ix0, p.Nx1));
        ps->cutbox[1] = max(0, min(ix1, p.Nx1));
        ps->cutbox[2] = max(0, min(iy0, p.Ny1));
        ps->cutbox[3] = max(0, min(iy1, p.Ny1));
    }
    __syncthreads();
    const int ix0 = ps->cutbox[0], nx = ps->cutbox[1] - ix0, iy0 = ps->cutbox[2], ny = ps->cutbox[3] - iy0;
    const long long total = (nx > 0 && ny > 0) ? (long long)nx * ny * p.Nz : 0;
    int hit = 0;
    for (long long k = tid; k < total && !hit; k += blockDim.x) {
        const int iz = (int)(k % p.Nz);
        const long long r2 = k / p.Nz;
        const int ix = ix0 + (int)(r2 % nx), iy = iy0 + (int)(r2 / nx);
        const double* o = p.off1 + (((size_t)iy * p.Nx1 + ix) * p.Nz + iz) * D;
        bool in1 = true;
        for (int i = 0; i < D; ++i) in1 = in1 && (o[i] >= s_lo[i]) && (o[i] <= s_hi[i]);
        hit = in1 ? 1 : 0;
    }
    if (hit) f->hit1 = 1;
    __syncthreads();
    return f->hit1;                                           // 0: empty init_area (:633-636)
}

// Per-cluster table for the common untrimmed case (cut == W, so the new centre is the cluster's own TDoA
// vector): is the box centre +- (W + 0.2)/2 non-empty?  A static property of the geometry, computed once at
// handle creation with the same code the per-mixture kernel uses for trimmed boxes.
__global__ void __launch_bounds__(128) box_table_kernel(SelectParams p, unsigned char* __restrict__ table) {
    __shared__ double s_lo[kMaxD], s_hi[kMaxD];
    __shared__ ProbeShared ps;
    const int g = blockIdx.x;
    if (threadIdx.x < p.D) {
        const double half = ((double)p.W + 0.2) / 2.0;
        const int c = p.cl_off[(size_t)g * p.D + threadIdx.x];
        s_lo[threadIdx.x] = (double)c - half;
        s_hi[threadIdx.x] = (double)c + half;
    }
    __syncthreads();
    const int r = area_nonempty(p, s_lo, s_hi, &ps);
    if (threadIdx.x == 0) table[g] = (unsigned char)r;
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(SelectParams p) {
    extern __shared__ int s_peakoff[];                       // [n][D] TDoA vectors of the peak clusters
    __shared__ unsigned long long keys[kMaxPeaks];
    __shared__ int s_vis[kMaxPeaks];
    __shared__ int s_poff[kMaxPatches][kMaxD];
    __shared__ int s_pw[kMaxPatches];
    __shared__ int s_centre[kMaxD], s_new[kMaxD];
    __shared__ double s_lo[kMaxD], s_hi[kMaxD];
    __shared__ int s_cut[2];                                 // double-buffered by processed-peak parity
    __shared__ ProbeShared ps;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int D = p.D, W = p.W;
    int n = p.count[b];
    if (n > p.max_peaks) n = p.max_peaks;
    if (n > kMaxPeaks) n = kMaxPeaks;
    const int32_t* ids = p.peaks + (size_t)b * p.max_peaks;
    const float* m = p.map + (size_t)b * p.G;

    // order: descending power, ties by first-seen position (:560); stage the peaks' TDoA vectors
    int sz = 1;
    while (sz < n) sz <<= 1;
    for (int i = tid; i < sz; i += kSelThreads) {
        unsigned long long k = ~0ull;
        if (i < n) k = ((unsigned long long)(~__float_as_uint(m[ids[i]])) << 32) | (unsigned)i;
        keys[i] = k;
        s_vis[i] = 0;
    }
    for (int i = tid; i < n * D; i += kSelThreads) {
        const int j = i / D;
        s_peakoff[i] = p.cl_off[(size_t)ids[j] * D + (i - j * D)];
    }
    __syncthreads();
    for (int size = 2; size <= sz; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (sz >> 1); t += kSelThreads) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == asc) {
                    keys[lo] = c;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }

    int npatch = 0, heavy = 0;
    for (int it = 0; it < n; ++it) {
        const int pid = (int)(keys[it] & 0xffffffffull);
        if (s_vis[pid] >= 1) continue;                       // uniform: written before the last barrier
        int* cutp = &s_cut[heavy & 1];
        ++heavy;
        // ---- phase A (warp 0 only): trim against the accepted patches (:580-602), new centre, box bounds
        if (warp == 0) {
            int cut = W;
            for (int q = lane; q < npatch; q += 32) {
                const double hw = (double)s_pw[q] / 2.0;
                double m1 = -CUDART_INF, m2 = CUDART_INF;
                for (int i = 0; i < D; ++i) {
                    const double delta = (double)(s_poff[q][i] - s_peakoff[pid * D + i]);
                    const double lo1 = (delta - hw) - (double)W / 2.0;   // range_low1 - range_high
                    const double hi1 = (delta + hw) + (double)W / 2.0;   // range_high1 - range_low
                    m1 = fmax(m1, lo1);
                    m2 = fmin(m2, hi1);
                }
                const int d1 = (int)rint(m1), d2 = (int)rint(m2);
                if (!(d1 >= 0 || d2 <= 0)) {
                    int c = W + d1;
                    if (c < 0) c = 0;
                    cut = min(cut, c);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cut = min(cut, __shfl_xor_sync(0xffffffffu, cut, d));
            if (lane < D) {
                const int c0 = s_peakoff[pid * D + lane];
                s_centre[lane] = c0;
                const int nw = (int)rint((double)c0 + (double)(cut - W) / 2.0);   // (:615)
                s_new[lane] = nw;
                const double half = ((double)cut + 0.2) / 2.0;   // width_list_new[0] + err_tolerance, halved (:27)
                s_lo[lane] = (double)nw - half;
                s_hi[lane] = (double)nw + half;
            }
            if (lane == 0) *cutp = cut;
        }
        __syncthreads();
        const int cut = *cutp;
        if (cut == 0) continue;                               // all_discard (:610-618)
        // ---- phase B (all warps): covered peaks (:623-624)
        {
            const double hb = ((double)W + 0.2) / 2.0;       // closed box centre +- 4.1 on integer vectors
            for (int j = tid; j < n; j += kSelThreads) {
                bool in = true;
                for (int i = 0; i < D; ++i) {
                    const double v = (double)s_peakoff[j * D + i];
                    in = in && (v >= (double)s_centre[i] - hb) && (v <= (double)s_centre[i] + hb);
                }
                if (in) s_vis[j] += 1;
            }
        }
        // ---- is there a 1 cm voxel inside the (possibly trimmed) box?  (:629-636)
        int ok;
        if (cut == W && p.box_table) {
            ok = p.box_table[ids[pid]];                       // untrimmed: static per cluster
            __syncthreads();                                  // s_vis visible before the next peak is examined
        } else {
            ok = area_nonempty(p, s_lo, s_hi, &ps);           // contains barriers (also publishes s_vis)
        }
        if (!ok) continue;
        // ---- accept (:637-639); written by warp 0, which is also the only reader before the next barrier
        if (npatch < kMaxPatches && npatch < p.max_patches && warp == 0) {
            if (lane < D) {
                s_poff[npatch][lane] = s_new[lane];
                p.out_off[((size_t)b * p.max_patches + npatch) * D + lane] = s_new[lane];
            }
            if (lane == 0) {
                s_pw[npatch] = cut;
                p.out_width[(size_t)b * p.max_patches + npatch] = cut;
                p.out_peak[(size_t)b * p.max_patches + npatch] = ids[pid];
            }
            __syncwarp();
        }
        ++npatch;
    }
    if (tid == 0) p.out_count[b] = npatch;
}

// Dense shift table from the per-mixture patch lists: shifts[n][0] = 0, shifts[n][c] = offset[c-1]
// (the offsets are already integers, so round_half_even(float32(.)) of network.py:81-82 is the identity).
__global__ void build_shift_table_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ off, int B,
                                         int max_patches, int D, int32_t* __restrict__ shifts,
                                         int32_t* __restrict__ mix_index, int32_t* __restrict__ n_total, int cap) {
    __shared__ int s_start[1025];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < B; ++b) {
            s_start[b] = acc;
            int c = cnt[b];
            if (c > max_patches) c = max_patches;
            acc += c;
        }
        s_start[B] = acc;
        *n_total = acc < cap ? acc : cap;
    }
    __syncthreads();
    const int M = D + 1;
    for (int b = 0; b < B; ++b) {
        const int s0 = s_start[b], c = s_start[b + 1] - s0;
        for (int i = threadIdx.x; i < c * M; i += blockDim.x) {
            const int q = i / M, ch = i - q * M;
            const int n = s0 + q;
            if (n >= cap) continue;
            shifts[(size_t)n * M + ch] = ch == 0 ? 0 : off[((size_t)b * max_patches + q) * D + ch - 1];
            if (ch == 0) mix_index[n] = b;
        }
    }
}

}  // namespace
}  // namespace asw

using namespace asw;

static void fill_geometry(const asw_select* h, SelectParams& p) {
    p.G = h->G;
    p.D = h->D;
    p.W = h->W;
    p.cl_off = h->d_cl_off;
    p.off5 = h->d_off5;
    p.off1s = h->d_off1s;
    p.vox5 = h->d_vox5;
    p.bstart = h->d_bstart;
    p.b0min = h->b0min;
    p.b1min = h->b1min;
    p.NB0 = h->NB0;
    p.NB1 = h->NB1;
    p.n5 = h->n5;
    p.Nx5 = h->Nx5;
    p.Ny5 = h->Ny5;
    p.xx5 = h->d_xx5;
    p.yy5 = h->d_yy5;
    p.ax0 = h->ax0;
    p.ax1 = h->ax1;
    p.ay0 = h->ay0;
    p.ay1 = h->ay1;
    p.off1 = h->d_off1;
    p.Ny1 = h->Ny1;
    p.Nx1 = h->Nx1;
    p.Nz = h->Nz;
    p.box_table = h->d_box;
}

extern "C" {

int asw_select_create(asw_select_t** out, int device, int G, int D, int W, const int32_t* cluster_offsets,
                      const double* off5_sorted, const double* off1_at5, const int32_t* vox5,
                      const int32_t* bucket_start, int b0min, int b1min, int NB0, int NB1, int n5, int Nx5, int Ny5,
                      const double* xx5, const double* yy5, const double* axis_range4, const double* off1, int Ny1,
                      int Nx1, int Nz) {
    if (!out || !cluster_offsets || !off5_sorted || !off1_at5 || !vox5 || !bucket_start || !xx5 || !yy5 ||
        !axis_range4 || !off1 || G < 1 || D < 1 || D > kMaxD || n5 < 1 || NB0 < 1 || NB1 < 1) {
        set_error("asw_select_create: null argument or unsupported shape (D=%d, limit %d)", D, kMaxD);
        return ASW_ERR_ARG;
    }
    *out = nullptr;
    ASW_CUDA_CHECK(cudaSetDevice(device));
    asw_select* h = new asw_select();
    h->device = device;
    h->G = G;
    h->D = D;
    h->W = W;
    h->n5 = n5;
    h->b0min = b0min;
    h->b1min = b1min;
    h->NB0 = NB0;
    h->NB1 = NB1;
    h->Nx5 = Nx5;
    h->Ny5 = Ny5;
    h->Ny1 = Ny1;
    h->Nx1 = Nx1;
    h->Nz = Nz;
    h->ax0 = axis_range4[0];
    h->ax1 = axis_range4[1];
    h->ay0 = axis_range4[2];
    h->ay1 = axis_range4[3];
    cudaError_t e = cudaSuccess;
    auto up = [&](auto** dst, const auto* src, size_t n) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dst, n * sizeof(**dst));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src, n * sizeof(**dst), cudaMemcpyHostToDevice);
    };
    up(&h->d_cl_off, cluster_offsets, (size_t)G * D);
    up(&h->d_off5, off5_sorted, (size_t)D * n5);
    up(&h->d_off1s, off1_at5, (size_t)D * n5);
    up(&h->d_vox5, vox5, (size_t)n5);
    up(&h->d_bstart, bucket_start, (size_t)NB0 * NB1 + 1);
    up(&h->d_xx5, xx5, (size_t)Nx5);
    up(&h->d_yy5, yy5, (size_t)Ny5);
    up(&h->d_off1, off1, (size_t)Ny1 * Nx1 * Nz * D);
    unsigned char* box = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&box, (size_t)G);
    if (e != cudaSuccess) {
        set_error("asw_select_create: %s", cudaGetErrorString(e));
        asw_select_destroy(h);
        return ASW_ERR_CUDA;
    }
    {
        SelectParams p{};
        fill_geometry(h, p);
        p.box_table = nullptr;
        box_table_kernel<<<G, 128>>>(p, box);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        count_launch();
        if (e != cudaSuccess) {
            set_error("asw_select_create: box table kernel: %s", cudaGetErrorString(e));
            cudaFree(box);
            asw_select_destroy(h);
            return ASW_ERR_CUDA;
        }
        h->d_box = box;
    }
    *out = h;
    return ASW_OK;
}

int asw_select_destroy(asw_select_t* h) {
    if (!h) return ASW_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_cl_off);
    cudaFree(h->d_off5);
    cudaFree(h->d_off1s);
    cudaFree(h->d_vox5);
    cudaFree(h->d_bstart);
    cudaFree(h->d_xx5);
    cudaFree(h->d_yy5);
    cudaFree(h->d_off1);
    cudaFree(h->d_box);
    delete h;
    return ASW_OK;
}

int asw_select_patches(asw_select_t* h, const float* map_dev, const int32_t* peaks_dev, int max_peaks,
                       const int32_t* count_dev, int B, int32_t* out_count_dev, int32_t* out_offsets_dev,
                       int32_t* out_width_dev, int32_t* out_peak_dev, int max_patches, void* stream) {
    if (!h || !map_dev || !peaks_dev || !count_dev || !out_count_dev || !out_offsets_dev || !out_width_dev ||
        !out_peak_dev || B < 1 || max_peaks < 1 || max_patches < 1) {
        set_error("asw_select_patches: null argument or bad shape");
        return ASW_ERR_ARG;
    }
    SelectParams p{};
    p.peaks = peaks_dev;
    p.count = count_dev;
    p.map = map_dev;
    p.max_peaks = max_peaks;
    fill_geometry(h, p);
    p.out_count = out_count_dev;
    p.out_off = out_offsets_dev;
    p.out_width = out_width_dev;
    p.out_peak = out_peak_dev;
    p.max_patches = max_patches < kMaxPatches ? max_patches : kMaxPatches;
    if (max_patches > kMaxPatches) {
        set_error("asw_select_patches: max_patches %d exceeds the kernel limit %d", max_patches, kMaxPatches);
        return ASW_ERR_ARG;
    }
    const int peak_words = (max_peaks < kMaxPeaks ? max_peaks : kMaxPeaks) * h->D;
    size_t smem = (size_t)peak_words * sizeof(int);
    if (smem > 160 * 1024) {
        set_error("asw_select_patches: %d peaks x %d dimensions do not fit the shared-memory stage", max_peaks, h->D);
        return ASW_ERR_RANGE;
    }
    if (smem > 16 * 1024) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    select_kernel<<<B, kSelThreads, smem, (cudaStream_t)stream>>>(p);
    ASW_LAUNCH_CHECK("select_kernel");
    return ASW_OK;
}

int asw_build_shift_table(const int32_t* count_dev, const int32_t* offsets_dev, int B, int max_patches, int D,
                          int32_t* shifts_dev, int32_t* mix_index_dev, int32_t* n_total_dev, int capacity,
                          void* stream) {
    if (!count_dev || !offsets_dev || !shifts_dev || !mix_index_dev || !n_total_dev || B < 1 || B > 1024 ||
        max_patches < 1 || D < 1 || capacity < 1) {
        set_error("asw_build_shift_table: null argument or bad shape (B <= 1024)");
        return ASW_ERR_ARG;
    }
    build_shift_table_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(count_dev, offsets_dev, B, max_patches, D, shifts_dev,
                                                                  mix_index_dev, n_total_dev, capacity);
    ASW_LAUNCH_CHECK("build_shift_table_kernel");
    return ASW_OK;
}

}  // extern "C"
