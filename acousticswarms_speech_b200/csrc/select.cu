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
    float* d_cell = nullptr;      // [ncy][ncx][Nz][D][2] min / max of Offset_1 over each 5 x 5 x 1 block of 1 cm voxels
    int ncx = 0, ncy = 0;
    // subdivision workspace (grown on demand): per candidate one area list and two ping-pong node lists
    int32_t* d_lists = nullptr;
    size_t lists_cap = 0;
    int32_t* d_nodes_i = nullptr;   // node records of the level walk (see subdivide_kernel)
    size_t nodes_i_cap = 0;
    double* d_nodes_d = nullptr;
    size_t nodes_d_cap = 0;
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
    const float* cell;                // per-cell bounds of Offset_1 (see cell_bounds_kernel), ncy x ncx x Nz cells
    int ncx, ncy;
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

// Outward-rounded min / max of every TDoA coordinate over each 5 x 5 x 1 block of the 1 cm volume.  The member scan of
// asw_subdivide tests a block's 12 bounds before touching its 25 voxels: a coarse patch's cut (the axis-aligned bounding
// box of a slanted hyperbolic sliver, up to ~1e6 voxels) is mostly voxels far outside the TDoA box.  Exact by
// construction: a block is skipped only if no voxel of it can pass the test.
constexpr int kCell = 5;
__global__ void __launch_bounds__(128) cell_bounds_kernel(const double* __restrict__ off1, int Ny1, int Nx1, int Nz, int D,
                                                          int ncy, int ncx, float* __restrict__ out) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (long long)ncy * ncx * Nz) return;
    const int iz = (int)(cell % Nz);
    const int cx = (int)((cell / Nz) % ncx), cy = (int)(cell / ((long long)Nz * ncx));
    for (int i = 0; i < D; ++i) {
        double mn = CUDART_INF, mx = -CUDART_INF;
        bool bad = false;
        for (int dy = 0; dy < kCell; ++dy)
            for (int dx = 0; dx < kCell; ++dx) {
                const int iy = kCell * cy + dy, ix = kCell * cx + dx;
                if (iy >= Ny1 || ix >= Nx1) continue;
                const double v = off1[(((size_t)iy * Nx1 + ix) * Nz + iz) * D + i];
                if (!(v == v)) bad = true;
                mn = fmin(mn, v);
                mx = fmax(mx, v);
            }
        if (bad) {                                    // a NaN inside: never skip the block
            mn = -CUDART_INF;
            mx = CUDART_INF;
        }
        out[((size_t)cell * D + i) * 2] = __double2float_rd(mn);
        out[((size_t)cell * D + i) * 2 + 1] = __double2float_ru(mx);
    }
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
// The kernel is instantiated for three dimension buckets, KD = 8 / 16 / 32 (D <= 31 = what select_kernel supports).
//
// Work layout (round 2): the tree is walked level by level, and WITHIN a level every phase is parallel over what it
// touches -- node decisions over the nodes (one thread each), member tests over the flattened member lists of all
// nodes of the level (the lists of a level are contiguous, in node order).  The first version processed the nodes of a
// level one after the other with ~5 block barriers and a single-thread section each: a median candidate (3 k member
// voxels, ~45 nodes) took 1.2 ms of almost pure barrier latency, the largest (39 k) 5 ms, and a batch of 542
// candidates 5-7 ms whatever the occupancy.  Now a level costs 5 barriers in total.
// Node records live in global scratch (read through L1; two buffers per candidate, ping-pong by level); shared memory
// holds only the per-node counters.  Slot assignment (leaf numbers, child slots, list offsets) is one short serial
// pass per level, which is what keeps the reference's order: leaves in level order, within a level in node order,
// children as (lower half, upper half).
constexpr int kListCap = 1 << 17;   // member voxels per candidate (and per level, duplicates included)

template <int KD>
struct SubCfg {
    static constexpr int kNodes = KD <= 8 ? 128 : (KD <= 16 ? 256 : 384);   // nodes per level
    static constexpr int kNI = 2 * KD + 8;                                   // ints per node record
    static constexpr int kND = 8 * KD;                                       // doubles per node record
    // One candidate per SM with as many warps as the register file allows: a candidate's time is what its own warps
    // can issue (2 warps per scheduler, mostly waiting on dependent loads, gave IPC 0.25 and a 4 ms tail for the
    // candidates with ~4e4 member voxels whatever else ran on the GPU), so threads per CANDIDATE matter, not CTAs per SM.
    static constexpr int kThreads = KD <= 8 ? 1024 : 512;
    static constexpr int kCtasPerSm = 1;
    static constexpr size_t smem_bytes() {
        return sizeof(double) * kNodes * 3 + sizeof(int) * ((size_t)kNodes * KD * 2 + kNodes * 2 + 2 * (kNodes + 1));
    }
};
// int fields after c[KD], w[KD]
enum { kFStart = 0, kFCount, kFFlag, kFChosen, kFSize0, kFSize1, kFSlot, kFOff0 };
enum { kNodeSplit = 0, kNodeLeaf = 1, kNodeDropped = 2 };

struct SubParams {
    SelectParams g;                 // geometry
    const int32_t* centres;         // [n][D]
    const int32_t* widths;          // [n] coarse width (identical in every dimension); <= 0: empty slot
    double ub[kMaxD];               // upper_bound_pairwise (sep/Mic_Array.py:113-115)
    int32_t* lists;                 // [n][3][kListCap]: area list, ping, pong
    int32_t* nodes_i;               // [n][2][kNodes][kNI]
    double* nodes_d;                // [n][2][kNodes][kND]: lo[KD], hi[KD] (membership box relative to the root's members),
                                    // then [KD][6] the bounds of the member tests (DimBounds), written once per node
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

// The bounds every membership test of a node is made of (all double, evaluated exactly as the first version did):
//   node box   c -+ w/2 -+ 1e-3                      Patch.hyperbola_sample (Patch_3D.py:40-47) on the checked-out node
//   half boxes hc -+ hw/2 -+ 1e-3, hc = trunc(c -+ w/4), hw = trunc(w/2)    binary_area_divide_width (:275-281)
struct DimBounds {
    double blo, bhi, hlo0, hhi0, hlo1, hhi1;
    int hw, hc0, hc1, elig;
};
__device__ __forceinline__ DimBounds dim_bounds(int ci, int wi) {
    DimBounds b;
    const double c = (double)ci, w = (double)wi;
    b.blo = c - w / 2.0 - 1e-3;
    b.bhi = c + w / 2.0 + 1e-3;
    b.elig = (w / 2.0 < 3.0) ? 0 : 1;                 // (:271-272) MIN_WIDTH
    b.hw = (int)trunc_ll(w / 2.0);                    // (:279-280) float into int64
    b.hc0 = (int)trunc_ll(c - w / 4.0);               // (:275-278)
    b.hc1 = (int)trunc_ll(c + w / 4.0);
    b.hlo0 = (double)b.hc0 - (double)b.hw / 2.0 - 1e-3;
    b.hhi0 = (double)b.hc0 + (double)b.hw / 2.0 + 1e-3;
    b.hlo1 = (double)b.hc1 - (double)b.hw / 2.0 - 1e-3;
    b.hhi1 = (double)b.hc1 + (double)b.hw / 2.0 + 1e-3;
    return b;
}

template <int KD>
__global__ void __launch_bounds__(SubCfg<KD>::kThreads, SubCfg<KD>::kCtasPerSm) subdivide_kernel(SubParams q) {
    using C = SubCfg<KD>;
    constexpr int NN = C::kNodes, NI = C::kNI, ND = C::kND;
    constexpr int kSubThreads = C::kThreads;
    constexpr int kU = 4;                                                // members per thread in flight
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* s_lsum = reinterpret_cast<double*>(s_raw);                   // [NN][3] leaf position sums
    int* s_cnt = reinterpret_cast<int*>(s_lsum + NN * 3);                // [NN][KD][2]
    int* s_pos = s_cnt + NN * KD * 2;                                    // [NN][2] append cursors of the two children
    int* s_start = s_pos + NN * 2;                                       // [2][NN + 1] list boundaries, by level parity
    __shared__ double s_lo[kMaxD], s_hi[kMaxD];
    __shared__ ProbeShared ps;
    __shared__ int s_n, s_nc, s_state[4];                                // n_leaf, status, n_nxt, (unused)
    const SelectParams& p = q.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cand = blockIdx.x, D = p.D;
    int32_t* area = q.lists + (size_t)cand * 3 * kListCap;
    int32_t* buf[2] = {area + kListCap, area + 2 * kListCap};
    int32_t* nodes_i = q.nodes_i + (size_t)cand * 2 * NN * NI;
    double* nodes_d = q.nodes_d + (size_t)cand * 2 * NN * ND;
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
    int status0 = 0;
    if (ps.f.cnt5 > 0) {
        if (tid == 0) cut_box(p, &ps);
        __syncthreads();
        const int ix0 = ps.cutbox[0], nx = ps.cutbox[1] - ix0, iy0 = ps.cutbox[2], ny = ps.cutbox[3] - iy0;
        // Stage 1: the 5 x 5 x 1 blocks overlapping the cut whose offset bounds intersect the TDoA box (see
        // cell_bounds_kernel), compacted into the (still unused) ping list as packed (ix, iy, iz) of the block's corner.
        int32_t* cells = buf[0];
        if (tid == 0) s_nc = 0;
        __syncthreads();
        if (nx > 0 && ny > 0) {
            const int cx0 = ix0 / kCell, cx1 = (ix0 + nx - 1) / kCell, cy0 = iy0 / kCell, cy1 = (iy0 + ny - 1) / kCell;
            const int ncx = cx1 - cx0 + 1, rowc = ncx * p.Nz;              // one row of blocks: cx, iz
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int j0 = 0; j0 < rowc; j0 += kSubThreads) {
                    const int j = j0 + tid;
                    bool hit = j < rowc;
                    int cx = 0, iz = 0;
                    if (hit) {
                        cx = cx0 + j / p.Nz;
                        iz = j - (j / p.Nz) * p.Nz;
                        const float* cb = p.cell + (((size_t)cy * p.ncx + cx) * p.Nz + iz) * D * 2;
                        for (int i = 0; i < D; ++i)
                            hit = hit && ((double)cb[2 * i + 1] >= s_lo[i]) && ((double)cb[2 * i] <= s_hi[i]);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&s_nc, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (hit) {
                        const int pos = base + __popc(m & ((1u << lane) - 1));
                        if (pos < kListCap) cells[pos] = (kCell * cx) | ((kCell * cy) << 12) | (iz << 24);
                    }
                }
        }
        __syncthreads();
        // Stage 2: the voxels of those blocks that lie inside the cut, tested exactly as the reference does
        const int ncell = min(s_nc, kListCap);
        if (s_nc > kListCap) status0 = 1;
        for (int j0 = 0; j0 < ncell * (kCell * kCell); j0 += 2 * kSubThreads) {
            bool in[2];
            int vox[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + u * kSubThreads + tid;
                in[u] = j < ncell * (kCell * kCell);
                vox[u] = 0;
                if (in[u]) {
                    const int c = cells[j / (kCell * kCell)], l = j % (kCell * kCell);
                    const int ix = (c & 0xfff) + l % kCell, iy = ((c >> 12) & 0xfff) + l / kCell, iz = (c >> 24) & 0xff;
                    in[u] = ix >= ix0 && ix < ix0 + nx && iy >= iy0 && iy < iy0 + ny;
                    if (in[u]) vox[u] = (iy * p.Nx1 + ix) * p.Nz + iz;
                }
            }
            for (int i = 0; i < D; ++i) {
                const double v0 = p.off1[(size_t)vox[0] * D + i], v1 = p.off1[(size_t)vox[1] * D + i];
                in[0] = in[0] && (v0 >= s_lo[i]) && (v0 <= s_hi[i]);
                in[1] = in[1] && (v1 >= s_lo[i]) && (v1 <= s_hi[i]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, in[u]);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (in[u]) {
                    const int pos = base + __popc(m & ((1u << lane) - 1));
                    if (pos < kListCap) area[pos] = vox[u];
                }
            }
        }
    }
    __syncthreads();
    int n_root = s_n;
    if (n_root > kListCap) {
        n_root = kListCap;
        status0 = 1;
    }

    if (q.root_members) {             // Patch.area_points of the candidate, as voxel indices (the host sorts them)
        int32_t* dstm = q.root_members + (size_t)cand * q.member_cap;
        for (int k = tid; k < n_root && k < q.member_cap; k += kSubThreads) dstm[k] = area[k];
        if (tid == 0) q.root_count[cand] = s_n;
    }

    // ---- level-synchronous walk of the split tree
    if (tid == 0) {
        int32_t* r = nodes_i;
        double* rd = nodes_d;
        for (int i = 0; i < D; ++i) {
            r[i] = q.centres[(size_t)cand * D + i];
            r[KD + i] = wc;
            rd[i] = s_lo[i];
            rd[KD + i] = s_hi[i];
        }
        r[2 * KD + kFStart] = 0;
        r[2 * KD + kFCount] = n_root;
        s_start[0] = 0;
        s_start[1] = n_root;
        s_state[0] = 0;
        s_state[1] = status0;
        s_state[2] = 0;
    }
    __syncthreads();
    const int32_t* src = area;
    int n_cur = 1, level = 0;
    const unsigned all = (D >= 32) ? 0xffffffffu : ((1u << D) - 1);
    while (n_cur > 0) {
        int32_t* dst = buf[level & 1];
        int32_t* ni = nodes_i + (size_t)(level & 1) * NN * NI;
        double* nd = nodes_d + (size_t)(level & 1) * NN * ND;
        int32_t* ni_nxt = nodes_i + (size_t)((level + 1) & 1) * NN * NI;
        double* nd_nxt = nodes_d + (size_t)((level + 1) & 1) * NN * ND;
        const int* start = s_start + (level & 1) * (NN + 1);
        int* start_nxt = s_start + ((level + 1) & 1) * (NN + 1);
        const int total = start[n_cur];

        // ---- A1, one thread per node: Patch.check_out (Patch_3D.py:69-87) in place, the leaf test of :260-262
        for (int node = tid; node < n_cur; node += kSubThreads) {
            int32_t* r = ni + (size_t)node * NI;
            int wmax = 0;
            for (int i = 0; i < D; ++i) {
                int c = r[i], w = r[KD + i];
                while (true) {
                    if (fabs((double)c) <= q.ub[i] || w <= 4) break;
                    if ((double)c > q.ub[i]) c = (int)trunc_ll((double)c - (double)w / 4.0);
                    else if ((double)c < -q.ub[i]) c = (int)trunc_ll((double)c + (double)w / 4.0);
                    w = (int)trunc_ll((double)w / 2.0);
                }
                r[i] = c;
                r[KD + i] = w;
                wmax = max(wmax, w);
                {   // the member loops read these instead of recomputing them per member (int -> double conversions and
                    // ten fp64 operations per dimension were a fifth of a large candidate's instructions)
                    const DimBounds b = dim_bounds(c, w);
                    double* bd = nd + (size_t)node * ND + 2 * KD + 6 * i;
                    bd[0] = b.blo;
                    bd[1] = b.bhi;
                    bd[2] = b.hlo0;
                    bd[3] = b.hhi0;
                    bd[4] = b.hlo1;
                    bd[5] = b.hhi1;
                }
                if (level == 0) {
                    q.root_after[((size_t)cand * 2 + 0) * D + i] = c;
                    q.root_after[((size_t)cand * 2 + 1) * D + i] = w;
                }
                s_cnt[(node * KD + i) * 2] = 0;
                s_cnt[(node * KD + i) * 2 + 1] = 0;
            }
            r[2 * KD + kFFlag] = ((double)wmax / 2.0 <= 2.0 && r[2 * KD + kFCount] <= 400) ? kNodeLeaf : kNodeSplit;
            s_pos[node * 2] = 0;
            s_pos[node * 2 + 1] = 0;
            s_lsum[node * 3] = 0.0;
            s_lsum[node * 3 + 1] = 0.0;
            s_lsum[node * 3 + 2] = 0.0;
        }
        __syncthreads();

        // ---- A2, flat over the members of the level: per eligible dimension, how many members fall into each half
        // (a member counts for dimension i when it is inside the node's box along every OTHER dimension).  A thread
        // keeps kU members in flight: the voxel index and then its D offsets are dependent global loads (~1 us per
        // member when taken one at a time, which is what bounded the kernel).
        {
            int node = 0;
            for (int k0 = 0; k0 < total; k0 += kU * kSubThreads) {
                int nodeu[kU], vox[kU];
                bool valid[kU], work[kU];
                unsigned in[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int k = k0 + u * kSubThreads + tid;
                    valid[u] = k < total;
                    if (valid[u])
                        while (k >= start[node + 1]) ++node;
                    nodeu[u] = node;
                    vox[u] = valid[u] ? src[k] : 0;
                    in[u] = 0;
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) work[u] = valid[u] && ni[(size_t)nodeu[u] * NI + 2 * KD + kFFlag] == kNodeSplit;
                for (int i = 0; i < D; ++i) {
                    double v[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u) v[u] = p.off1[(size_t)vox[u] * D + i];       // independent loads
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const double* bd = nd + (size_t)nodeu[u] * ND + 2 * KD + 6 * i;
                        if (work[u] && v[u] >= bd[0] && v[u] <= bd[1]) in[u] |= 1u << i;
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int nu = nodeu[u];
                    const int32_t* r = ni + (size_t)nu * NI;
                    const double* o = p.off1 + (size_t)vox[u] * D;
                    // whole warp inside one node (the common case: lists are contiguous): vote, one atomic per tally
                    const int n0 = __shfl_sync(0xffffffffu, nu, 0);
                    const bool uniform = __all_sync(0xffffffffu, !valid[u] || nu == n0);
                    const unsigned missing = all & ~in[u];
                    const bool maybe = work[u] && !(missing & (missing - 1));   // outside along two dimensions: nowhere
                    if (uniform) {
                        const int32_t* r0 = ni + (size_t)n0 * NI;
                        if (r0[2 * KD + kFFlag] != kNodeSplit) continue;         // uniform
                        const double* bd0 = nd + (size_t)n0 * ND + 2 * KD;
                        for (int i = 0; i < D; ++i) {
                            if (r0[KD + i] < 6) continue;                        // not eligible (w / 2 < 3), uniform
                            bool h0 = false, h1 = false;
                            if (maybe && ((in[u] | (1u << i)) == all)) {
                                const double v = o[i];
                                h0 = v >= bd0[6 * i + 2] && v <= bd0[6 * i + 3];
                                h1 = v >= bd0[6 * i + 4] && v <= bd0[6 * i + 5];
                            }
                            const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                            if (lane == 0) {
                                if (m0) atomicAdd(&s_cnt[(n0 * KD + i) * 2], __popc(m0));
                                if (m1) atomicAdd(&s_cnt[(n0 * KD + i) * 2 + 1], __popc(m1));
                            }
                        }
                    } else if (maybe) {
                        const double* bd = nd + (size_t)nu * ND + 2 * KD;
                        for (int i = 0; i < D; ++i) {
                            if (r[KD + i] < 6 || ((in[u] | (1u << i)) != all)) continue;
                            const double v = o[i];
                            if (v >= bd[6 * i + 2] && v <= bd[6 * i + 3]) atomicAdd(&s_cnt[(nu * KD + i) * 2], 1);
                            if (v >= bd[6 * i + 4] && v <= bd[6 * i + 5]) atomicAdd(&s_cnt[(nu * KD + i) * 2 + 1], 1);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- B1, one thread per node: the reference's choice (:264-331): most balanced split, dimensions still wider
        // than 2 first; (:333-335) "min_patch is None or len(two_patches) == 0" -- two_patches of the LAST dimension tried
        for (int node = tid; node < n_cur; node += kSubThreads) {
            int32_t* r = ni + (size_t)node * NI;
            if (r[2 * KD + kFFlag] == kNodeLeaf) continue;
            int best = -1, best_diff = 2500000, last = -1;
            bool keep8 = false;
            for (int i = 0; i < D; ++i) {
                const DimBounds b = dim_bounds(r[i], r[KD + i]);
                if (!b.elig) continue;
                last = i;
                const int diff = abs(s_cnt[(node * KD + i) * 2] - s_cnt[(node * KD + i) * 2 + 1]);
                if (b.hw > 2) {
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
            if (best < 0 || last < 0 || (s_cnt[(node * KD + last) * 2] == 0 && s_cnt[(node * KD + last) * 2 + 1] == 0)) {
                r[2 * KD + kFFlag] = kNodeLeaf;
            } else {
                r[2 * KD + kFChosen] = best;
                r[2 * KD + kFSize0] = s_cnt[(node * KD + best) * 2];
                r[2 * KD + kFSize1] = s_cnt[(node * KD + best) * 2 + 1];
            }
        }
        __syncthreads();

        // ---- B2, one thread: leaf numbers, child slots and list offsets in node order (the reference's order)
        if (tid == 0) {
            int n_leaf = s_state[0], status = s_state[1], n_nxt = 0, dst_off = 0;
            for (int node = 0; node < n_cur; ++node) {
                int32_t* r = ni + (size_t)node * NI;
                if (r[2 * KD + kFFlag] == kNodeLeaf) {
                    r[2 * KD + kFSlot] = n_leaf;
                    if (n_leaf >= q.max_leaves) status = 3;
                    ++n_leaf;
                    continue;
                }
                const int size0 = r[2 * KD + kFSize0], size1 = r[2 * KD + kFSize1];
                const int add = (size0 > 0 ? 1 : 0) + (size1 > 0 ? 1 : 0);
                if (dst_off + size0 + size1 > kListCap) {
                    status = 1;
                    r[2 * KD + kFFlag] = kNodeDropped;
                } else if (n_nxt + add > NN) {
                    status = 2;
                    r[2 * KD + kFFlag] = kNodeDropped;
                } else {
                    r[2 * KD + kFSlot] = n_nxt;
                    r[2 * KD + kFOff0] = dst_off;
                    if (size0 > 0) start_nxt[n_nxt++] = dst_off;
                    if (size1 > 0) start_nxt[n_nxt++] = dst_off + size0;
                    dst_off += size0 + size1;
                }
            }
            start_nxt[n_nxt] = dst_off;
            s_state[0] = n_leaf;
            s_state[1] = status;
            s_state[2] = n_nxt;
        }
        __syncthreads();

        // ---- B3, one thread per node: leaf records / child node records
        for (int node = tid; node < n_cur; node += kSubThreads) {
            const int32_t* r = ni + (size_t)node * NI;
            const double* rd = nd + (size_t)node * ND;
            const int flag = r[2 * KD + kFFlag];
            if (flag == kNodeLeaf) {
                const int slot = r[2 * KD + kFSlot];
                if (slot < q.max_leaves) {
                    const size_t lb = (size_t)cand * q.max_leaves + slot;
                    for (int i = 0; i < D; ++i) {
                        q.leaf_off[lb * D + i] = r[i];
                        q.leaf_w[lb * D + i] = r[KD + i];
                        q.leaf_box[(lb * 2 + 0) * D + i] = rd[i];
                        q.leaf_box[(lb * 2 + 1) * D + i] = rd[KD + i];
                    }
                    q.leaf_npts[lb] = r[2 * KD + kFCount];
                }
            } else if (flag == kNodeSplit) {
                const int chosen = r[2 * KD + kFChosen];
                int slot = r[2 * KD + kFSlot];
                for (int sd = 0; sd < 2; ++sd) {
                    const int sz = r[2 * KD + (sd == 0 ? kFSize0 : kFSize1)];
                    if (sz == 0) continue;
                    int32_t* ch = ni_nxt + (size_t)slot * NI;
                    double* chd = nd_nxt + (size_t)slot * ND;
                    for (int i = 0; i < D; ++i) {
                        const DimBounds b = dim_bounds(r[i], r[KD + i]);
                        int c = r[i], w = r[KD + i];
                        double lo = b.blo, hi = b.bhi;
                        if (i == chosen) {
                            c = sd == 0 ? b.hc0 : b.hc1;
                            w = b.hw;
                            lo = sd == 0 ? b.hlo0 : b.hlo1;
                            hi = sd == 0 ? b.hhi0 : b.hhi1;
                        }
                        ch[i] = c;
                        ch[KD + i] = w;
                        chd[i] = fmax(rd[i], lo);
                        chd[KD + i] = fmin(rd[KD + i], hi);
                    }
                    ch[2 * KD + kFStart] = start_nxt[slot];
                    ch[2 * KD + kFCount] = sz;
                    ++slot;
                }
            }
        }
        // ---- C, flat over the members: children's member lists (a member within 1e-3 of the split plane goes to both
        // halves); position sums of the leaves' members for Patch.center_pos().  kU members in flight as above.
        {
            int node = 0;
            for (int k0 = 0; k0 < total; k0 += kU * kSubThreads) {
                int nodeu[kU], vox[kU], flag[kU], chosen[kU];
                bool valid[kU], others[kU];
                double vi[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int k = k0 + u * kSubThreads + tid;
                    valid[u] = k < total;
                    if (valid[u])
                        while (k >= start[node + 1]) ++node;
                    nodeu[u] = node;
                    vox[u] = valid[u] ? src[k] : 0;
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int32_t* r = ni + (size_t)nodeu[u] * NI;
                    flag[u] = valid[u] ? r[2 * KD + kFFlag] : kNodeDropped;
                    chosen[u] = flag[u] == kNodeSplit ? r[2 * KD + kFChosen] : -1;
                    others[u] = flag[u] == kNodeSplit;
                    vi[u] = 0.0;
                }
                bool any_split = false;
#pragma unroll
                for (int u = 0; u < kU; ++u) any_split = any_split || flag[u] == kNodeSplit;
                if (any_split)
                    for (int i = 0; i < D; ++i) {
                        double v[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) v[u] = p.off1[(size_t)vox[u] * D + i];   // independent loads
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            if (i == chosen[u]) {
                                vi[u] = v[u];
                            } else {
                                const double* bd = nd + (size_t)nodeu[u] * ND + 2 * KD + 6 * i;
                                others[u] = others[u] && (v[u] >= bd[0]) && (v[u] <= bd[1]);
                            }
                        }
                    }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int nu = nodeu[u];
                    const int32_t* r = ni + (size_t)nu * NI;
                    bool in0 = false, in1 = false;
                    if (flag[u] == kNodeSplit) {
                        const double* bc = nd + (size_t)nu * ND + 2 * KD + 6 * chosen[u];
                        in0 = others[u] && vi[u] >= bc[2] && vi[u] <= bc[3];
                        in1 = others[u] && vi[u] >= bc[4] && vi[u] <= bc[5];
                    }
                    double sx = 0.0, sy = 0.0, sz = 0.0;
                    const bool centre = flag[u] == kNodeLeaf && q.leaf_centre != nullptr;
                    if (centre) {
                        const int iz = vox[u] % p.Nz, r2 = vox[u] / p.Nz;
                        const int ix = r2 % p.Nx1, iy = r2 / p.Nx1;
                        sx = q.grid1[ix];
                        sy = q.grid1[p.Nx1 + iy];
                        sz = q.grid1[p.Nx1 + p.Ny1 + iz];
                    }
                    const int n0 = __shfl_sync(0xffffffffu, nu, 0);
                    const bool uniform = __all_sync(0xffffffffu, !valid[u] || nu == n0);
                    if (uniform) {
                        const int32_t* r0 = ni + (size_t)n0 * NI;
                        const int f0 = r0[2 * KD + kFFlag];
                        if (f0 == kNodeSplit) {                                  // warp-aggregated append
                            const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
                            int b0 = 0, b1 = 0;
                            if (lane == 0) {
                                if (m0) b0 = atomicAdd(&s_pos[n0 * 2], __popc(m0));
                                if (m1) b1 = atomicAdd(&s_pos[n0 * 2 + 1], __popc(m1));
                            }
                            b0 = __shfl_sync(0xffffffffu, b0, 0);
                            b1 = __shfl_sync(0xffffffffu, b1, 0);
                            const unsigned below = (1u << lane) - 1;
                            const int off0 = r0[2 * KD + kFOff0], off1 = off0 + r0[2 * KD + kFSize0];
                            if (in0) dst[off0 + b0 + __popc(m0 & below)] = vox[u];
                            if (in1) dst[off1 + b1 + __popc(m1 & below)] = vox[u];
                        } else if (f0 == kNodeLeaf && q.leaf_centre != nullptr) {
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) {
                                sx += __shfl_xor_sync(0xffffffffu, sx, d);
                                sy += __shfl_xor_sync(0xffffffffu, sy, d);
                                sz += __shfl_xor_sync(0xffffffffu, sz, d);
                            }
                            if (lane == 0) {
                                atomicAdd(&s_lsum[n0 * 3], sx);
                                atomicAdd(&s_lsum[n0 * 3 + 1], sy);
                                atomicAdd(&s_lsum[n0 * 3 + 2], sz);
                            }
                        }
                    } else {
                        if (flag[u] == kNodeSplit) {
                            const int off0 = r[2 * KD + kFOff0], off1 = off0 + r[2 * KD + kFSize0];
                            if (in0) dst[off0 + atomicAdd(&s_pos[nu * 2], 1)] = vox[u];
                            if (in1) dst[off1 + atomicAdd(&s_pos[nu * 2 + 1], 1)] = vox[u];
                        } else if (centre) {
                            atomicAdd(&s_lsum[nu * 3], sx);
                            atomicAdd(&s_lsum[nu * 3 + 1], sy);
                            atomicAdd(&s_lsum[nu * 3 + 2], sz);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- D, one thread per node: Patch.center_pos() of the leaves (np.mean of area_points; the sums above are
        // accumulated in double in arbitrary order: differences of a few ulp between runs)
        if (q.leaf_centre)
            for (int node = tid; node < n_cur; node += kSubThreads) {
                const int32_t* r = ni + (size_t)node * NI;
                const int slot = r[2 * KD + kFSlot], cnt = r[2 * KD + kFCount];
                if (r[2 * KD + kFFlag] != kNodeLeaf || slot >= q.max_leaves) continue;
                for (int a = 0; a < 3; ++a)
                    q.leaf_centre[((size_t)cand * q.max_leaves + slot) * 3 + a] =
                        cnt > 0 ? s_lsum[node * 3 + a] / (double)cnt : nan("");
            }
        n_cur = s_state[2];
        src = dst;
        ++level;
        __syncthreads();
    }
    if (tid == 0) {
        q.leaf_count[cand] = s_state[0];
        q.status[cand] = s_state[1];
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

template <typename T>
int grow_scratch(T** ptr, size_t* cap, size_t need, const char* what) {
    if (need <= *cap) return ASW_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need * sizeof(T));
    if (e != cudaSuccess) {
        set_error("asw_subdivide: %s of %zu bytes: %s", what, need * sizeof(T), cudaGetErrorString(e));
        return ASW_ERR_ALLOC;
    }
    *cap = need;
    return ASW_OK;
}

template <int KD>
int launch_subdivide(asw_select* h, SubParams& q, int n, cudaStream_t stream) {
    using C = SubCfg<KD>;
    int rc;
    if ((rc = grow_scratch(&h->d_nodes_i, &h->nodes_i_cap, (size_t)n * 2 * C::kNodes * C::kNI, "node scratch")) != ASW_OK) return rc;
    if ((rc = grow_scratch(&h->d_nodes_d, &h->nodes_d_cap, (size_t)n * 2 * C::kNodes * C::kND, "node scratch")) != ASW_OK) return rc;
    q.nodes_i = h->d_nodes_i;
    q.nodes_d = h->d_nodes_d;
    const size_t smem = C::smem_bytes();
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(subdivide_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // enough carve-out for the resident CTAs the register budget allows (the default fits fewer)
        int pct = (int)((C::kCtasPerSm * (smem + 4096) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
        if (pct > 100) pct = 100;
        ASW_CUDA_CHECK(cudaFuncSetAttribute(subdivide_kernel<KD>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    subdivide_kernel<KD><<<n, C::kThreads, smem, stream>>>(q);
    ASW_LAUNCH_CHECK("subdivide_kernel");
    return ASW_OK;
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
    p.cell = h->d_cell;
    p.ncx = h->ncx;
    p.ncy = h->ncy;
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
    if (Nx1 > 4095 || Ny1 > 4095 || Nz > 255) {
        set_error("asw_select_create: 1 cm volume %d x %d x %d exceeds the packed index range (4095 x 4095 x 255)", Nx1, Ny1, Nz);
        return ASW_ERR_RANGE;
    }
    DeviceGuard guard(device);      // the caller's current device is restored on return
    if (!guard.ok) {
        set_error("cannot make CUDA device %d current (no CUDA device, or a bad index)", device);
        return ASW_ERR_CUDA;
    }
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
    {
        h->ncx = (Nx1 + kCell - 1) / kCell;
        h->ncy = (Ny1 + kCell - 1) / kCell;
        const long long cells = (long long)h->ncx * h->ncy * Nz;
        e = cudaMalloc(&h->d_cell, sizeof(float) * 2 * (size_t)cells * D);
        if (e == cudaSuccess) {
            cell_bounds_kernel<<<(unsigned)((cells + 127) / 128), 128>>>(h->d_off1, Ny1, Nx1, Nz, D, h->ncy, h->ncx, h->d_cell);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            count_launch();
        }
        if (e != cudaSuccess) {
            set_error("asw_select_create: cell bounds: %s", cudaGetErrorString(e));
            asw_select_destroy(h);
            return ASW_ERR_CUDA;
        }
    }
    *out = h;
    return ASW_OK;
}

int asw_select_destroy(asw_select_t* h) {
    if (!h) return ASW_OK;
    DeviceGuard guard(h->device);
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
    cudaFree(h->d_cell);
    cudaFree(h->d_lists);
    cudaFree(h->d_nodes_i);
    cudaFree(h->d_nodes_d);
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
    ASW_CARVE_ONCE(select_kernel);
    select_kernel<<<B, kSelThreads, smem, (cudaStream_t)stream>>>(p);
    ASW_LAUNCH_CHECK("select_kernel");
    return ASW_OK;
}

int asw_select_set_grid1(asw_select_t* h, const double* xx1, const double* yy1, const double* zz) {
    if (!h || !xx1 || !yy1 || !zz) {
        set_error("asw_select_set_grid1: null argument");
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(h->device);
    if (!guard.ok) {
        set_error("cannot make device %d current", h->device);
        return ASW_ERR_CUDA;
    }
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
    if (h->D > kMaxD) {
        set_error("asw_subdivide: %d TDoA dimensions exceed the device path's limit of %d (use the host path)", h->D,
                  kMaxD);
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
    if (h->D <= 8) return launch_subdivide<8>(h, q, n, (cudaStream_t)stream);
    if (h->D <= 16) return launch_subdivide<16>(h, q, n, (cudaStream_t)stream);
    return launch_subdivide<32>(h, q, n, (cudaStream_t)stream);
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
    ASW_CARVE_ONCE(build_shift_table_kernel);
    build_shift_table_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(count_dev, offsets_dev, B, max_patches, D, shifts_dev,
                                                                  mix_index_dev, n_total_dev, capacity);
    ASW_LAUNCH_CHECK("build_shift_table_kernel");
    return ASW_OK;
}

}  // extern "C"
