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
    double* d_grid1 = nullptr;    // [Nx1 + Ny1 + Nz] coordinates of the 1 cm volume (asw_select_set_grid1), optional
    unsigned char* d_box = nullptr;  // [G] is the untrimmed box of cluster g non-empty
    // subdivision workspace (grown on demand): per candidate one area list and two ping-pong node lists
    int32_t* d_lists = nullptr;
    size_t lists_cap = 0;
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

// 5 cm probe (:45-47): counts the 5 cm voxels inside the closed box [lo, hi], their index bounding box, and
// whether one of the coinciding 1 cm voxels is inside.  Block-wide, ends with a barrier; results in ps->f.
__device__ void probe5cm(const SelectParams& p, const double* s_lo, const double* s_hi, ProbeShared* ps) {
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
}

// The reference's 1 cm cut (:49-58): bounding box of the 5 cm hits +- 5 cm, clamped to the room, as index ranges
// into the 1 cm volume (python slice semantics).  One thread.
__device__ void cut_box(const SelectParams& p, ProbeShared* ps) {
    const ProbeFlags* f = &ps->f;
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

// "Does the closed TDoA box [lo, hi] contain a 1 cm voxel of the room?" -- hyperbola_area_init (:41-61) reduced to
// its decision: 5 cm probe, the coinciding 1 cm voxels first, the exact 1 cm cut only if none of them is inside.
// Block-wide (all threads call it, contains barriers); lo/hi are in shared memory and already visible.
// Returns 1 = non-empty, 0 = empty / no 5 cm voxel.
__device__ int area_nonempty(const SelectParams& p, const double* s_lo, const double* s_hi, ProbeShared* ps) {
    const int tid = threadIdx.x;
    const int D = p.D;
    ProbeFlags* f = &ps->f;
    probe5cm(p, s_lo, s_hi, ps);
    if (f->cnt5 == 0) return 0;                               // init_area is None (:46-47)
    if (f->hit1 != 0) return 1;
    // exact scan of the reference's 1 cm cut (:49-60); rare: no coinciding 1 cm voxel was inside
    if (tid == 0) cut_box(p, ps);
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

// ---------------------------------------------------------------------------------------------------------
// Subdivision of the kept coarse hypercubes: search_area / binary_area_divide_width
// (sep/helpers/local_utils_3d.py:212-335) with Patch.check_out (sep/Traditional_SP/Patch_3D.py:69-87).
// The reference recomputes the TDoA vector of every 1 cm member point (:220-224); those values are bit for bit
// Offset_1 at the member voxels (numpy's x ** 0.5 is sqrt), so the device works on voxel indices into Offset_1.
// A node's member set is always "root members inside an axis-aligned box" (the intersection of the closed
// half-boxes on the path), which is also what the host needs to rebuild a leaf's area_points on demand.
constexpr int kSubD = 8;            // TDoA dimensions supported on the device path (M <= 9)
constexpr int kSubThreads = 256;
constexpr int kMaxNodes = 128;      // nodes per level
constexpr int kListCap = 1 << 17;   // member voxels per candidate (and per level, duplicates included)

struct SubNode {
    int c[kSubD], w[kSubD];
    double lo[kSubD], hi[kSubD];    // membership box relative to the root's members
    int start, count;
};

struct SubParams {
    SelectParams g;                 // geometry
    const int32_t* centres;         // [n][D]
    const int32_t* widths;          // [n] coarse width (identical in every dimension)
    double ub[kSubD];               // upper_bound_pairwise (sep/Mic_Array.py:113-115)
    int32_t* lists;                 // [n][3][kListCap]: area list, ping, pong
    int max_leaves;
    int32_t* leaf_count;            // [n]
    int32_t* leaf_off;              // [n][max_leaves][D]
    int32_t* leaf_w;                // [n][max_leaves][D]
    int32_t* leaf_npts;             // [n][max_leaves]
    double* leaf_box;               // [n][max_leaves][2][D]
    double* leaf_centre;            // [n][max_leaves][3] mean position of the leaf's member voxels (optional)
    int32_t* root_members;          // [n][member_cap] 1 cm voxel indices inside the coarse patch, unordered (optional)
    int32_t* root_count;            // [n] how many there are (may exceed member_cap: then the list is truncated)
    int member_cap;
    const double* grid1;            // xx1 [Nx1], yy1 [Ny1], zz [Nz]
    int32_t* root_after;            // [n][2][D] root offsets / widths after check_out (the reference mutates the candidate)
    int32_t* status;                // [n] 0 ok, 1 member list overflow, 2 node overflow, 3 leaf overflow
};

__device__ __forceinline__ long long trunc_ll(double x) { return (long long)x; }   // numpy float -> int64 item assignment

__global__ void __launch_bounds__(kSubThreads) subdivide_kernel(SubParams q) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    SubNode* cur = reinterpret_cast<SubNode*>(s_raw);
    SubNode* nxt = cur + kMaxNodes;
    __shared__ double s_lo[kMaxD], s_hi[kMaxD];
    __shared__ ProbeShared ps;
    __shared__ int s_n, s_cnt[kSubD][2], s_pos[2], s_flag[4];
    __shared__ double s_blo[kSubD], s_bhi[kSubD], s_hlo[kSubD][2], s_hhi[kSubD][2];
    __shared__ int s_elig[kSubD], s_hc[kSubD][2], s_hw[kSubD];
    __shared__ double s_sum[kSubThreads / 32][3];
    const SelectParams& p = q.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cand = blockIdx.x, D = p.D;
    int32_t* area = q.lists + (size_t)cand * 3 * kListCap;
    int32_t* buf[2] = {area + kListCap, area + 2 * kListCap};
    const int wc = q.widths[cand];
    if (wc <= 0) {                    // empty slot of a padded candidate list (batched callers): nothing to subdivide
        if (tid == 0) {
            q.leaf_count[cand] = 0;
            q.status[cand] = 0;
            if (q.root_count) q.root_count[cand] = 0;
        }
        return;
    }

    // ---- member voxels of the coarse patch: hyperbola_area_init (SRP_Prunning.py:41-61), unordered
    if (tid < D) {
        const double half = ((double)wc + 0.2) / 2.0;
        const int c = q.centres[(size_t)cand * D + tid];
        s_lo[tid] = (double)c - half;
        s_hi[tid] = (double)c + half;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    probe5cm(p, s_lo, s_hi, &ps);
    int status = 0;
    if (ps.f.cnt5 > 0) {
        if (tid == 0) cut_box(p, &ps);
        __syncthreads();
        const int ix0 = ps.cutbox[0], nx = ps.cutbox[1] - ix0, iy0 = ps.cutbox[2], ny = ps.cutbox[3] - iy0;
        const long long total = (nx > 0 && ny > 0) ? (long long)nx * ny * p.Nz : 0;
        for (long long k0 = 0; k0 < total; k0 += kSubThreads) {
            const long long k = k0 + tid;
            bool in = false;
            int vox = 0;
            if (k < total) {
                const int iz = (int)(k % p.Nz);
                const long long r2 = k / p.Nz;
                const int ix = ix0 + (int)(r2 % nx), iy = iy0 + (int)(r2 / nx);
                vox = (iy * p.Nx1 + ix) * p.Nz + iz;
                const double* o = p.off1 + (size_t)vox * D;
                in = true;
                for (int i = 0; i < D; ++i) in = in && (o[i] >= s_lo[i]) && (o[i] <= s_hi[i]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (in) {
                const int pos = base + __popc(m & ((1u << lane) - 1));
                if (pos < kListCap) area[pos] = vox;
            }
        }
    }
    __syncthreads();
    int n_root = s_n;
    if (n_root > kListCap) {
        n_root = kListCap;
        status = 1;
    }

    if (q.root_members) {             // Patch.area_points of the candidate, as voxel indices (the host sorts them)
        int32_t* dstm = q.root_members + (size_t)cand * q.member_cap;
        for (int k = tid; k < n_root && k < q.member_cap; k += kSubThreads) dstm[k] = area[k];
        if (tid == 0) q.root_count[cand] = s_n;
    }

    // ---- level-synchronous walk of the split tree
    int n_cur = 1, n_leaf = 0;
    if (tid == 0) {
        for (int i = 0; i < D; ++i) {
            cur[0].c[i] = q.centres[(size_t)cand * D + i];
            cur[0].w[i] = wc;
            cur[0].lo[i] = s_lo[i];
            cur[0].hi[i] = s_hi[i];
        }
        cur[0].start = 0;
        cur[0].count = n_root;
    }
    __syncthreads();
    const int32_t* src = area;
    int level = 0;
    while (n_cur > 0) {
        int32_t* dst = buf[level & 1];
        int n_nxt = 0, dst_off = 0;
        for (int ni = 0; ni < n_cur; ++ni) {
            SubNode* nd = &cur[ni];
            if (tid == 0) {
                // Patch.check_out (Patch_3D.py:69-87), in place
                for (int i = 0; i < D; ++i) {
                    while (true) {
                        const int c = nd->c[i], w = nd->w[i];
                        if (fabs((double)c) <= q.ub[i] || w <= 4) break;
                        if ((double)c > q.ub[i]) nd->c[i] = (int)trunc_ll((double)c - (double)w / 4.0);
                        else if ((double)c < -q.ub[i]) nd->c[i] = (int)trunc_ll((double)c + (double)w / 4.0);
                        nd->w[i] = (int)trunc_ll((double)w / 2.0);
                    }
                }
                if (level == 0)
                    for (int i = 0; i < D; ++i) {
                        q.root_after[((size_t)cand * 2 + 0) * D + i] = nd->c[i];
                        q.root_after[((size_t)cand * 2 + 1) * D + i] = nd->w[i];
                    }
                int wmax = 0;
                for (int i = 0; i < D; ++i) wmax = max(wmax, nd->w[i]);
                // (:260-262) leaf: every dimension fine enough and few enough member points
                s_flag[0] = ((double)wmax / 2.0 <= 2.0 && nd->count <= 400) ? 1 : 0;
                for (int i = 0; i < D; ++i) {
                    const double c = (double)nd->c[i], w = (double)nd->w[i];
                    s_blo[i] = c - w / 2.0 - 1e-3;            // Patch.hyperbola_sample bounds (Patch_3D.py:40-47)
                    s_bhi[i] = c + w / 2.0 + 1e-3;
                    s_elig[i] = (w / 2.0 < 3.0) ? 0 : 1;      // (:271-272) MIN_WIDTH
                    const int hw = (int)trunc_ll(w / 2.0);    // (:279-280) float into int64
                    s_hw[i] = hw;
                    s_hc[i][0] = (int)trunc_ll(c - w / 4.0);  // (:275-278)
                    s_hc[i][1] = (int)trunc_ll(c + w / 4.0);
                    for (int sd = 0; sd < 2; ++sd) {
                        s_hlo[i][sd] = (double)s_hc[i][sd] - (double)hw / 2.0 - 1e-3;
                        s_hhi[i][sd] = (double)s_hc[i][sd] + (double)hw / 2.0 + 1e-3;
                    }
                    s_cnt[i][0] = 0;
                    s_cnt[i][1] = 0;
                }
            }
            __syncthreads();
            bool leaf = s_flag[0] != 0;
            int chosen = -1;
            if (!leaf) {
                // one pass: member counts of both halves along every eligible dimension
                int cnt[kSubD][2];
#pragma unroll
                for (int i = 0; i < kSubD; ++i) cnt[i][0] = cnt[i][1] = 0;
                for (int k = tid; k < nd->count; k += kSubThreads) {
                    const double* o = p.off1 + (size_t)src[nd->start + k] * D;
                    unsigned in = 0;
                    double v[kSubD];
#pragma unroll
                    for (int i = 0; i < kSubD; ++i)
                        if (i < D) {
                            v[i] = o[i];
                            if (v[i] >= s_blo[i] && v[i] <= s_bhi[i]) in |= 1u << i;
                        }
                    const unsigned all = (1u << D) - 1;
#pragma unroll
                    for (int i = 0; i < kSubD; ++i)
                        if (i < D && s_elig[i] && ((in | (1u << i)) == all)) {
                            if (v[i] >= s_hlo[i][0] && v[i] <= s_hhi[i][0]) ++cnt[i][0];
                            if (v[i] >= s_hlo[i][1] && v[i] <= s_hhi[i][1]) ++cnt[i][1];
                        }
                }
#pragma unroll
                for (int i = 0; i < kSubD; ++i)
                    if (i < D) {
#pragma unroll
                        for (int sd = 0; sd < 2; ++sd) {
                            int c = cnt[i][sd];
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                            if (lane == 0 && c) atomicAdd(&s_cnt[i][sd], c);
                        }
                    }
                __syncthreads();
                if (tid == 0) {
                    // the reference's choice (:264-331): most balanced split, dimensions still wider than 2 first
                    int best = -1, best_diff = 2500000, last = -1;
                    bool keep8 = false;
                    for (int i = 0; i < D; ++i) {
                        if (!s_elig[i]) continue;
                        last = i;
                        const int diff = abs(s_cnt[i][0] - s_cnt[i][1]);
                        if (s_hw[i] > 2) {
                            if (!keep8) {
                                best_diff = diff;
                                best = i;
                                keep8 = true;
                            } else if (diff < best_diff) {
                                best_diff = diff;
                                best = i;
                            }
                        } else if (!keep8 && diff < best_diff) {
                            best_diff = diff;
                            best = i;
                        }
                    }
                    // (:333-335) "min_patch is None or len(two_patches) == 0" -- two_patches of the LAST dimension tried
                    if (best < 0 || last < 0 || (s_cnt[last][0] == 0 && s_cnt[last][1] == 0)) s_flag[1] = -1;
                    else s_flag[1] = best;
                    s_pos[0] = 0;
                    s_pos[1] = 0;
                }
                __syncthreads();
                chosen = s_flag[1];
                if (chosen < 0) leaf = true;
            }
            if (leaf) {
                if (q.leaf_centre && n_leaf < q.max_leaves) {
                    // Patch.center_pos() of the leaf: mean of its member voxels' positions (Patch_3D.py, np.mean of
                    // area_points); summed per thread, per warp, then over the warps in a fixed order
                    double sx = 0.0, sy = 0.0, sz = 0.0;
                    for (int k = tid; k < nd->count; k += kSubThreads) {
                        const int vox = src[nd->start + k];
                        const int iz = vox % p.Nz, r2 = vox / p.Nz;
                        const int ix = r2 % p.Nx1, iy = r2 / p.Nx1;
                        sx += q.grid1[ix];
                        sy += q.grid1[p.Nx1 + iy];
                        sz += q.grid1[p.Nx1 + p.Ny1 + iz];
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, d);
                        sy += __shfl_xor_sync(0xffffffffu, sy, d);
                        sz += __shfl_xor_sync(0xffffffffu, sz, d);
                    }
                    if (lane == 0) {
                        s_sum[tid >> 5][0] = sx;
                        s_sum[tid >> 5][1] = sy;
                        s_sum[tid >> 5][2] = sz;
                    }
                    __syncthreads();
                    if (tid < 3) {
                        double t = 0.0;
                        for (int w = 0; w < kSubThreads / 32; ++w) t += s_sum[w][tid];
                        q.leaf_centre[((size_t)cand * q.max_leaves + n_leaf) * 3 + tid] =
                            nd->count > 0 ? t / (double)nd->count : nan("");
                    }
                }
                if (tid == 0) {
                    if (n_leaf < q.max_leaves) {
                        const size_t lb = (size_t)cand * q.max_leaves + n_leaf;
                        for (int i = 0; i < D; ++i) {
                            q.leaf_off[lb * D + i] = nd->c[i];
                            q.leaf_w[lb * D + i] = nd->w[i];
                            q.leaf_box[(lb * 2 + 0) * D + i] = nd->lo[i];
                            q.leaf_box[(lb * 2 + 1) * D + i] = nd->hi[i];
                        }
                        q.leaf_npts[lb] = nd->count;
                    }
                }
                if (n_leaf >= q.max_leaves) status = 3;
                ++n_leaf;
                __syncthreads();
                continue;
            }
            // split along `chosen`: children = the non-empty halves, in order; their members by a second pass
            const int size0 = s_cnt[chosen][0], size1 = s_cnt[chosen][1];
            const int off0 = dst_off, off1 = dst_off + size0;
            const int add = (size0 > 0 ? 1 : 0) + (size1 > 0 ? 1 : 0);
            if (off1 + size1 > kListCap) {
                status = 1;                                   // uniform: all of these are shared values
            } else if (n_nxt + add > kMaxNodes) {
                status = 2;
            } else {
                for (int k0 = 0; k0 < nd->count; k0 += kSubThreads) {
                    const int k = k0 + tid;
                    bool in0 = false, in1 = false;
                    int vox = 0;
                    if (k < nd->count) {
                        vox = src[nd->start + k];
                        const double* o = p.off1 + (size_t)vox * D;
                        bool others = true;
                        double vi = 0.0;
                        for (int i = 0; i < D; ++i) {
                            const double v = o[i];
                            if (i == chosen) vi = v;
                            else others = others && (v >= s_blo[i]) && (v <= s_bhi[i]);
                        }
                        in0 = others && vi >= s_hlo[chosen][0] && vi <= s_hhi[chosen][0];
                        in1 = others && vi >= s_hlo[chosen][1] && vi <= s_hhi[chosen][1];
                    }
                    // warp-aggregated append (a member within 1e-3 of the split plane goes to both halves)
                    const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
                    int b0 = 0, b1 = 0;
                    if (lane == 0) {
                        if (m0) b0 = atomicAdd(&s_pos[0], __popc(m0));
                        if (m1) b1 = atomicAdd(&s_pos[1], __popc(m1));
                    }
                    b0 = __shfl_sync(0xffffffffu, b0, 0);
                    b1 = __shfl_sync(0xffffffffu, b1, 0);
                    const unsigned below = (1u << lane) - 1;
                    if (in0) dst[off0 + b0 + __popc(m0 & below)] = vox;
                    if (in1) dst[off1 + b1 + __popc(m1 & below)] = vox;
                }
                if (tid == 0) {
                    int slot = n_nxt;
                    for (int sd = 0; sd < 2; ++sd) {
                        const int sz = sd == 0 ? size0 : size1;
                        if (sz == 0) continue;
                        SubNode* ch = &nxt[slot++];
                        for (int i = 0; i < D; ++i) {
                            ch->c[i] = nd->c[i];
                            ch->w[i] = nd->w[i];
                            double lo = s_blo[i], hi = s_bhi[i];
                            if (i == chosen) {
                                ch->c[i] = s_hc[i][sd];
                                ch->w[i] = s_hw[i];
                                lo = s_hlo[i][sd];
                                hi = s_hhi[i][sd];
                            }
                            ch->lo[i] = fmax(nd->lo[i], lo);
                            ch->hi[i] = fmin(nd->hi[i], hi);
                        }
                        ch->start = sd == 0 ? off0 : off1;
                        ch->count = sz;
                    }
                }
                n_nxt += add;
                dst_off = off1 + size1;
            }
            __syncthreads();
        }
        // next level
        SubNode* t = cur;
        cur = nxt;
        nxt = t;
        n_cur = n_nxt;
        src = dst;
        ++level;
        __syncthreads();
    }
    if (tid == 0) {
        q.leaf_count[cand] = n_leaf;
        q.status[cand] = status;
    }
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

// Fine-stage shift table (the patch-list assembly of Spotform_Small_Patch_Parallel, sep/Mic_Array.py:244-262): every
// candidate contributes its leaves followed by one centre patch at its own (checked-out) offsets.
// Kernel 1: exclusive prefix of the row counts (one CTA, chunks of 1024 candidates); kernel 2: one CTA per candidate.
__global__ void __launch_bounds__(1024) fine_table_prefix_kernel(const int32_t* __restrict__ leaf_count,
                                                                  const int32_t* __restrict__ widths, int n,
                                                                  int max_leaves, int32_t* __restrict__ cand_start,
                                                                  int32_t* __restrict__ n_total, int cap) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        int rows = 0;
        if (i < n && widths[i] > 0) rows = min(leaf_count[i], max_leaves) + 1;
        int incl = rows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += v;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int before = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - rows;
        if (i < n) cand_start[i] = before;
        __syncthreads();
        if (tid == 1023) s_carry = before + rows;
        __syncthreads();
    }
    if (tid == 0) {
        cand_start[n] = s_carry;
        *n_total = s_carry < cap ? s_carry : cap;
    }
}

__global__ void __launch_bounds__(128) fine_table_rows_kernel(const int32_t* __restrict__ leaf_count,
                                                               const int32_t* __restrict__ leaf_off,
                                                               const int32_t* __restrict__ root_after,
                                                               const int32_t* __restrict__ owner,
                                                               const int32_t* __restrict__ cand_start, int max_leaves,
                                                               int D, int32_t* __restrict__ shifts,
                                                               int32_t* __restrict__ mix_index,
                                                               int32_t* __restrict__ cand_index, int cap) {
    const int i = blockIdx.x;
    const int s0 = cand_start[i], rows = cand_start[i + 1] - s0;
    if (rows <= 0) return;
    const int M = D + 1, leaves = rows - 1;
    const int own = owner ? owner[i] : 0;
    for (int k = threadIdx.x; k < rows * M; k += blockDim.x) {
        const int q = k / M, ch = k - q * M;
        const int nrow = s0 + q;
        if (nrow >= cap) continue;
        int v = 0;
        if (ch > 0)
            v = q < leaves ? leaf_off[((size_t)i * max_leaves + q) * D + ch - 1] : root_after[(size_t)i * 2 * D + ch - 1];
        shifts[(size_t)nrow * M + ch] = v;
        if (ch == 0) {
            mix_index[nrow] = own;
            if (cand_index) cand_index[nrow] = i;
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
    cudaFree(h->d_grid1);
    cudaFree(h->d_yy5);
    cudaFree(h->d_off1);
    cudaFree(h->d_box);
    cudaFree(h->d_lists);
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
    DeviceGuard guard(h->device);
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

int asw_select_set_grid1(asw_select_t* h, const double* xx1, const double* yy1, const double* zz) {
    if (!h || !xx1 || !yy1 || !zz) {
        set_error("asw_select_set_grid1: null argument");
        return ASW_ERR_ARG;
    }
    ASW_CUDA_CHECK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->Nx1 + h->Ny1 + h->Nz;
    if (!h->d_grid1) ASW_CUDA_CHECK(cudaMalloc(&h->d_grid1, n * sizeof(double)));
    ASW_CUDA_CHECK(cudaMemcpy(h->d_grid1, xx1, sizeof(double) * h->Nx1, cudaMemcpyHostToDevice));
    ASW_CUDA_CHECK(cudaMemcpy(h->d_grid1 + h->Nx1, yy1, sizeof(double) * h->Ny1, cudaMemcpyHostToDevice));
    ASW_CUDA_CHECK(cudaMemcpy(h->d_grid1 + h->Nx1 + h->Ny1, zz, sizeof(double) * h->Nz, cudaMemcpyHostToDevice));
    return ASW_OK;
}

int asw_subdivide(asw_select_t* h, const int32_t* centres_dev, const int32_t* widths_dev, int n,
                  const double* upper_bound, int max_leaves, int32_t* leaf_count_dev, int32_t* leaf_off_dev,
                  int32_t* leaf_w_dev, int32_t* leaf_npts_dev, double* leaf_box_dev, double* leaf_centre_dev,
                  int32_t* root_after_dev, int32_t* status_dev, int32_t* root_members_dev, int member_cap,
                  int32_t* root_count_dev, void* stream) {
    if (root_members_dev && (!root_count_dev || member_cap < 1)) {
        set_error("asw_subdivide: root_members_dev needs root_count_dev and member_cap >= 1");
        return ASW_ERR_ARG;
    }
    if (leaf_centre_dev && (!h || !h->d_grid1)) {
        set_error("asw_subdivide: leaf centres need the 1 cm grid coordinates (asw_select_set_grid1)");
        return ASW_ERR_ARG;
    }
    if (!h || !centres_dev || !widths_dev || !upper_bound || !leaf_count_dev || !leaf_off_dev || !leaf_w_dev ||
        !leaf_npts_dev || !leaf_box_dev || !root_after_dev || !status_dev || n < 0 || max_leaves < 1) {
        set_error("asw_subdivide: null argument or bad shape");
        return ASW_ERR_ARG;
    }
    if (h->D > kSubD) {
        set_error("asw_subdivide: %d TDoA dimensions exceed the device path's limit of %d (use the host path)", h->D,
                  kSubD);
        return ASW_ERR_RANGE;
    }
    if (n == 0) return ASW_OK;
    DeviceGuard guard(h->device);
    const size_t need = (size_t)n * 3 * kListCap;
    if (need > h->lists_cap) {
        if (h->d_lists) cudaFree(h->d_lists);
        h->d_lists = nullptr;
        h->lists_cap = 0;
        cudaError_t e = cudaMalloc(&h->d_lists, need * sizeof(int32_t));
        if (e != cudaSuccess) {
            set_error("asw_subdivide: workspace of %zu bytes: %s", need * sizeof(int32_t), cudaGetErrorString(e));
            return ASW_ERR_ALLOC;
        }
        h->lists_cap = need;
    }
    SubParams q{};
    fill_geometry(h, q.g);
    q.centres = centres_dev;
    q.widths = widths_dev;
    for (int i = 0; i < h->D; ++i) q.ub[i] = upper_bound[i];
    q.lists = h->d_lists;
    q.max_leaves = max_leaves;
    q.leaf_count = leaf_count_dev;
    q.leaf_off = leaf_off_dev;
    q.leaf_w = leaf_w_dev;
    q.leaf_npts = leaf_npts_dev;
    q.leaf_box = leaf_box_dev;
    q.leaf_centre = leaf_centre_dev;
    q.root_members = root_members_dev;
    q.root_count = root_count_dev;
    q.member_cap = member_cap;
    q.grid1 = h->d_grid1;
    q.root_after = root_after_dev;
    q.status = status_dev;
    const size_t smem = 2 * (size_t)kMaxNodes * sizeof(SubNode);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(subdivide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    subdivide_kernel<<<n, kSubThreads, smem, (cudaStream_t)stream>>>(q);
    ASW_LAUNCH_CHECK("subdivide_kernel");
    return ASW_OK;
}

int asw_build_fine_table(const int32_t* leaf_count_dev, const int32_t* leaf_off_dev, const int32_t* root_after_dev,
                         const int32_t* widths_dev, const int32_t* owner_dev, int n, int max_leaves, int D,
                         int32_t* shifts_dev, int32_t* mix_index_dev, int32_t* cand_index_dev, int32_t* cand_start_dev,
                         int32_t* n_total_dev, int capacity, void* stream) {
    if (!leaf_count_dev || !leaf_off_dev || !root_after_dev || !widths_dev || !shifts_dev || !mix_index_dev ||
        !cand_start_dev || !n_total_dev || n < 1 || max_leaves < 1 || D < 1 || capacity < 1) {
        set_error("asw_build_fine_table: null argument or bad shape");
        return ASW_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    fine_table_prefix_kernel<<<1, 1024, 0, s>>>(leaf_count_dev, widths_dev, n, max_leaves, cand_start_dev, n_total_dev,
                                                capacity);
    ASW_LAUNCH_CHECK("fine_table_prefix_kernel");
    fine_table_rows_kernel<<<n, 128, 0, s>>>(leaf_count_dev, leaf_off_dev, root_after_dev, owner_dev, cand_start_dev,
                                             max_leaves, D, shifts_dev, mix_index_dev, cand_index_dev, capacity);
    ASW_LAUNCH_CHECK("fine_table_rows_kernel");
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
