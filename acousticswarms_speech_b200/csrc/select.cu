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
// version: one CTA per mixture walks the peaks in order; the 5 cm volume is stored sorted by its first
// TDoA coordinate so a binary search bounds the scan to ~10 % of the voxels; the "any 1 cm voxel inside"
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
    double* d_off5 = nullptr;     // [D][n5] sorted by coordinate 0
    double* d_off1s = nullptr;    // [D][n5] 1 cm offsets at the voxel coinciding with each 5 cm voxel (NaN: none)
    int32_t* d_vox5 = nullptr;    // [n5] iy * Nx5 + ix of the sorted voxels
    double* d_xx5 = nullptr;      // [Nx5]
    double* d_yy5 = nullptr;      // [Ny5]
    double* d_off1 = nullptr;     // [Ny1][Nx1][Nz][D]
};

namespace asw {
namespace {

constexpr int kSelThreads = 512;
constexpr int kMaxPeaks = 1024;
constexpr int kMaxPatches = 128;
constexpr int kMaxD = 31;

struct SelectParams {
    const int32_t* peaks;   // [B][max_peaks]
    const int32_t* count;   // [B]
    const float* map;       // [B][G]
    int max_peaks, G, D, W;
    const int32_t* cl_off;
    const double* off5;
    const double* off1s;
    const int32_t* vox5;
    int n5, Nx5, Ny5;
    const double* xx5;
    const double* yy5;
    double ax0, ax1, ay0, ay1;
    const double* off1;
    int Ny1, Nx1, Nz;
    int32_t* out_count;     // [B]
    int32_t* out_off;       // [B][max_patches][D]
    int32_t* out_width;     // [B][max_patches]
    int32_t* out_peak;      // [B][max_patches] cluster id of the peak the patch grew from
    int max_patches;
};

__global__ void __launch_bounds__(kSelThreads) select_kernel(SelectParams p) {
    __shared__ unsigned long long keys[kMaxPeaks];
    __shared__ int s_vis[kMaxPeaks];
    __shared__ int s_poff[kMaxPatches][kMaxD];
    __shared__ int s_pw[kMaxPatches];
    __shared__ int s_centre[kMaxD], s_new[kMaxD];
    __shared__ double s_lo[kMaxD], s_hi[kMaxD];
    __shared__ int s_cut, s_cnt5, s_hit1, s_ixmin, s_ixmax, s_iymin, s_iymax, s_cutbox[4];
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int D = p.D, W = p.W;
    int n = p.count[b];
    if (n > p.max_peaks) n = p.max_peaks;
    if (n > kMaxPeaks) n = kMaxPeaks;
    const int32_t* ids = p.peaks + (size_t)b * p.max_peaks;
    const float* m = p.map + (size_t)b * p.G;

    // order: descending power, ties by first-seen position (:560)
    int sz = 1;
    while (sz < n) sz <<= 1;
    for (int i = tid; i < sz; i += kSelThreads) {
        unsigned long long k = ~0ull;
        if (i < n) k = ((unsigned long long)(~__float_as_uint(m[ids[i]])) << 32) | (unsigned)i;
        keys[i] = k;
        s_vis[i] = 0;
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

    int npatch = 0;
    for (int it = 0; it < n; ++it) {
        const int pid = (int)(keys[it] & 0xffffffffull);
        if (s_vis[pid] >= 1) continue;                       // uniform: shared value, no writes since last sync
        const int id = ids[pid];
        __syncthreads();
        if (tid < D) s_centre[tid] = p.cl_off[(size_t)id * D + tid];
        if (tid == 0) {
            s_cut = W;
            s_cnt5 = 0;
            s_hit1 = 0;
            s_ixmin = INT_MAX;
            s_iymin = INT_MAX;
            s_ixmax = -1;
            s_iymax = -1;
        }
        __syncthreads();
        // trim against the accepted patches (:580-602)
        for (int q = tid; q < npatch; q += kSelThreads) {
            const double hw = (double)s_pw[q] / 2.0;
            double m1 = -CUDART_INF, m2 = CUDART_INF;
            for (int i = 0; i < D; ++i) {
                const double delta = (double)(s_poff[q][i] - s_centre[i]);
                const double lo1 = (delta - hw) - (double)W / 2.0;   // range_low1 - range_high
                const double hi1 = (delta + hw) + (double)W / 2.0;   // range_high1 - range_low
                m1 = fmax(m1, lo1);
                m2 = fmin(m2, hi1);
            }
            const int d1 = (int)rint(m1), d2 = (int)rint(m2);
            if (!(d1 >= 0 || d2 <= 0)) {
                int c = W + d1;
                if (c < 0) c = 0;
                atomicMin(&s_cut, c);
            }
        }
        __syncthreads();
        const int cut = s_cut;
        if (cut == 0) continue;                               // all_discard (:610-618)
        if (tid < D) {
            const int nw = (int)rint((double)s_centre[tid] + (double)(cut - W) / 2.0);   // (:615)
            s_new[tid] = nw;
            const double half = ((double)cut + 0.2) / 2.0;   // width_list_new[0] + err_tolerance, halved (:27)
            s_lo[tid] = (double)nw - half;
            s_hi[tid] = (double)nw + half;
        }
        // covered peaks: closed box centre +- (W + 0.2)/2 on integer TDoA vectors == |diff| <= 4 for W = 8 (:623-624)
        {
            const double hb = ((double)W + 0.2) / 2.0;
            for (int j = tid; j < n; j += kSelThreads) {
                const int32_t* o = p.cl_off + (size_t)ids[j] * D;
                bool in = true;
                for (int i = 0; i < D && in; ++i) {
                    const double v = (double)o[i];
                    in = (v >= (double)s_centre[i] - hb) && (v <= (double)s_centre[i] + hb);
                }
                if (in) s_vis[j] += 1;
            }
        }
        __syncthreads();
        // 5 cm probe restricted by binary search on the sorted first coordinate
        int a = 0, e = p.n5;
        {
            const double lo0 = s_lo[0], hi0 = s_hi[0];
            int l = 0, r = p.n5;
            while (l < r) {
                const int mid = (l + r) >> 1;
                if (p.off5[mid] < lo0) l = mid + 1; else r = mid;
            }
            a = l;
            r = p.n5;
            while (l < r) {
                const int mid = (l + r) >> 1;
                if (p.off5[mid] <= hi0) l = mid + 1; else r = mid;
            }
            e = l;
        }
        for (int k = a + tid; k < e; k += kSelThreads) {
            bool in = true;
            for (int i = 1; i < D && in; ++i) {
                const double v = p.off5[(size_t)i * p.n5 + k];
                in = (v >= s_lo[i]) && (v <= s_hi[i]);
            }
            if (in) {
                const int v = p.vox5[k];
                const int iy = v / p.Nx5, ix = v - iy * p.Nx5;
                atomicAdd(&s_cnt5, 1);
                atomicMin(&s_ixmin, ix);
                atomicMax(&s_ixmax, ix);
                atomicMin(&s_iymin, iy);
                atomicMax(&s_iymax, iy);
                bool in1 = true;                              // the coinciding 1 cm voxel
                for (int i = 0; i < D && in1; ++i) {
                    const double v1 = p.off1s[(size_t)i * p.n5 + k];
                    in1 = (v1 >= s_lo[i]) && (v1 <= s_hi[i]);  // NaN compares false
                }
                if (in1) s_hit1 = 1;
            }
        }
        __syncthreads();
        if (s_cnt5 == 0) continue;                            // init_area is None (:46-47, :631-632)
        if (s_hit1 == 0) {
            // exact scan of the reference's 1 cm cut (:49-60)
            if (tid == 0) {
                double x0 = p.xx5[s_ixmin] - 0.05, x1 = p.xx5[s_ixmax] + 0.05;
                x0 = fmax(p.ax0, x0);
                x1 = fmin(p.ax1, x1);
                double y0 = p.yy5[s_iymin] - 0.05, y1 = p.yy5[s_iymax] + 0.05;
                y0 = fmax(p.ay0, y0);
                y1 = fmin(p.ay1, y1);
                int ix0 = (int)floor((x0 - p.ax0) / 0.01), ix1 = (int)ceil((x1 - p.ax0) / 0.01);
                int iy0 = (int)floor((y0 - p.ay0) / 0.01), iy1 = (int)ceil((y1 - p.ay0) / 0.01);
                s_cutbox[0] = max(0, min(ix0, p.Nx1));
                s_cutbox[1] = max(0, min(ix1, p.Nx1));
                s_cutbox[2] = max(0, min(iy0, p.Ny1));
                s_cutbox[3] = max(0, min(iy1, p.Ny1));
            }
            __syncthreads();
            const int ix0 = s_cutbox[0], nx = s_cutbox[1] - ix0, iy0 = s_cutbox[2], ny = s_cutbox[3] - iy0;
            const long long total = (nx > 0 && ny > 0) ? (long long)nx * ny * p.Nz : 0;
            for (long long k = tid; k < total; k += kSelThreads) {
                if ((k & 0x3fff) < kSelThreads && s_hit1) break;
                const int iz = (int)(k % p.Nz);
                const long long r2 = k / p.Nz;
                const int ix = ix0 + (int)(r2 % nx), iy = iy0 + (int)(r2 / nx);
                const double* o = p.off1 + (((size_t)iy * p.Nx1 + ix) * p.Nz + iz) * D;
                bool in1 = true;
                for (int i = 0; i < D && in1; ++i) in1 = (o[i] >= s_lo[i]) && (o[i] <= s_hi[i]);
                if (in1) s_hit1 = 1;
            }
            __syncthreads();
            if (s_hit1 == 0) continue;                        // empty init_area (:633-636)
        }
        // accept (:637-639)
        if (npatch < kMaxPatches && npatch < p.max_patches) {
            if (tid < D) {
                s_poff[npatch][tid] = s_new[tid];
                p.out_off[((size_t)b * p.max_patches + npatch) * D + tid] = s_new[tid];
            }
            if (tid == 0) {
                s_pw[npatch] = cut;
                p.out_width[(size_t)b * p.max_patches + npatch] = cut;
                p.out_peak[(size_t)b * p.max_patches + npatch] = id;
            }
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

extern "C" {

int asw_select_create(asw_select_t** out, int device, int G, int D, int W, const int32_t* cluster_offsets,
                      const double* off5_sorted, const double* off1_at5, const int32_t* vox5, int n5, int Nx5, int Ny5,
                      const double* xx5, const double* yy5, const double* axis_range4, const double* off1, int Ny1,
                      int Nx1, int Nz) {
    if (!out || !cluster_offsets || !off5_sorted || !off1_at5 || !vox5 || !xx5 || !yy5 || !axis_range4 || !off1 ||
        G < 1 || D < 1 || D > kMaxD || n5 < 1) {
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
    up(&h->d_xx5, xx5, (size_t)Nx5);
    up(&h->d_yy5, yy5, (size_t)Ny5);
    up(&h->d_off1, off1, (size_t)Ny1 * Nx1 * Nz * D);
    if (e != cudaSuccess) {
        set_error("asw_select_create: %s", cudaGetErrorString(e));
        asw_select_destroy(h);
        return ASW_ERR_CUDA;
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
    cudaFree(h->d_xx5);
    cudaFree(h->d_yy5);
    cudaFree(h->d_off1);
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
    p.G = h->G;
    p.D = h->D;
    p.W = h->W;
    p.cl_off = h->d_cl_off;
    p.off5 = h->d_off5;
    p.off1s = h->d_off1s;
    p.vox5 = h->d_vox5;
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
    p.out_count = out_count_dev;
    p.out_off = out_offsets_dev;
    p.out_width = out_width_dev;
    p.out_peak = out_peak_dev;
    p.max_patches = max_patches < kMaxPatches ? max_patches : kMaxPatches;
    if (max_patches > kMaxPatches) {
        set_error("asw_select_patches: max_patches %d exceeds the kernel limit %d", max_patches, kMaxPatches);
        return ASW_ERR_ARG;
    }
    select_kernel<<<B, kSelThreads, 0, (cudaStream_t)stream>>>(p);
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
