// prune.cu -- peak picking of the SRP map on the device.
//
// Reference (sep/Traditional_SP/SRP_Prunning.py):
//   :347-357  fill_powermap_torch: POWER_MAP[voxel] = SRP_map[cluster(voxel)] for member voxels, 0 elsewhere
//   :432      MAX_POWER = amax(SRP_map)
//   :500-544  find_valid_peak_new: thr = clamp(0.15 MAX, 0.015, 0.05), thr2 = 4 thr; on the interior
//             [2:-2, 2:-2, 1:-1]:  cond1 = P > thr2 (1 + 1/d);
//             cond2 = P >= all 49 neighbours (dx,dy in [-2,2], dz in {-1,0} -- sic) and P > thr (0.9 + 1/d)
//                     and P <= thr2 (1 + 1/d);   peaks = voxels with cond1 | cond2, mapped to cluster ids,
//             duplicates dropped keeping the first occurrence in C order.
// The reference does this in numpy on the host after copying the map back and scattering it voxel by
// voxel in a Python loop.  Here the power volume is never materialised: a voxel's power is
// map[power_index[voxel]], comparisons are done in double on the float32 map values (exactly what the
// host code does with the same map), and the first-occurrence rule is an atomicMin on the voxel's C-order
// rank followed by a small per-mixture sort.
#include <limits.h>

#include "common.cuh"

struct asw_peaks {
    int device = 0, Lx = 0, Ly = 0, Lz = 0, G = 0;
    int32_t* d_index = nullptr;   // [Lx][Ly][Lz] cluster id or -1
    double* d_a1 = nullptr;       // [Lx][Ly] 0.9 + 1/dis
    double* d_a2 = nullptr;       // [Lx][Ly] 1 + 1/dis
    double thr[3] = {0.15, 0.015, 0.05};
    double ratio2 = 4.0;
    int* d_first = nullptr;       // [Bcap][G] first C-order rank of a peak voxel of the cluster
    int Bcap = 0;
};

namespace asw {
namespace {

constexpr int kMaxPeaksSort = 2048;

// MAX_POWER of every mixture; the same sweep resets the per-cluster "first peak rank" scratch of peak_flag_kernel, so a
// call never depends on how the previous call on this handle ended.
__global__ void __launch_bounds__(1024) map_max_kernel(const float* __restrict__ map, int G, float* __restrict__ mx,
                                                        int* __restrict__ first) {
    __shared__ float s[32];
    const float* m = map + (size_t)blockIdx.x * G;
    int* f = first + (size_t)blockIdx.x * G;
    float v = 0.f;   // the map is >= 0 by construction (:253)
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        v = fmaxf(v, m[g]);
        f[g] = INT_MAX;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = s[threadIdx.x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        if (threadIdx.x == 0) mx[blockIdx.x] = v;
    }
}

struct PeakParams {
    const float* map;
    const int32_t* index;
    const double* a1;
    const double* a2;
    const float* mx;
    int* first;
    int Lx, Ly, Lz, G;
    double t_ratio, t_lo, t_hi, ratio2;
};

__device__ __forceinline__ double voxel_power(const PeakParams& p, const float* m, int ix, int iy, int iz) {
    const int id = p.index[((size_t)ix * p.Ly + iy) * p.Lz + iz];
    return id >= 0 ? (double)m[id] : 0.0;
}

__global__ void __launch_bounds__(256) peak_flag_kernel(PeakParams p) {
    const int b = blockIdx.y;
    const int nx = p.Lx - 4, ny = p.Ly - 4, nz = p.Lz - 2;
    const int rank = blockIdx.x * blockDim.x + threadIdx.x;   // C-order rank inside the interior volume
    if (rank >= nx * ny * nz) return;
    const int iz = rank % nz + 1;
    const int iy = (rank / nz) % ny + 2;
    const int ix = rank / (nz * ny) + 2;
    const float* m = p.map + (size_t)b * p.G;
    const int id = p.index[((size_t)ix * p.Ly + iy) * p.Lz + iz];
    if (id < 0) return;                       // power 0 can never exceed a positive threshold
    const double v = (double)m[id];
    double thr = __dmul_rn(p.t_ratio, (double)p.mx[b]);
    if (thr < p.t_lo) thr = p.t_lo;
    else if (thr > p.t_hi) thr = p.t_hi;
    const double thr2 = __dmul_rn(thr, p.ratio2);
    const double t1 = __dmul_rn(thr, p.a1[ix * p.Ly + iy]);
    const double t2 = __dmul_rn(thr2, p.a2[ix * p.Ly + iy]);
    bool peak = v > t2;
    if (!peak && v > t1) {
        peak = true;
        for (int dx = -2; dx <= 2 && peak; ++dx)
            for (int dy = -2; dy <= 2 && peak; ++dy)
                for (int dz = -1; dz <= 0; ++dz) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    if (!(v >= voxel_power(p, m, ix + dx, iy + dy, iz + dz))) {
                        peak = false;
                        break;
                    }
                }
    }
    if (peak) atomicMin(p.first + (size_t)b * p.G + id, rank);
}

// one CTA per mixture: gather (rank, id) of flagged clusters, sort by rank, emit ids; reset the scratch
__global__ void __launch_bounds__(1024) peak_collect_kernel(int* __restrict__ first, int G, int32_t* __restrict__ peaks,
                                                             int max_peaks, int32_t* __restrict__ count) {
    __shared__ unsigned long long keys[kMaxPeaksSort];
    __shared__ int n;
    const int b = blockIdx.x;
    int* f = first + (size_t)b * G;
    if (threadIdx.x == 0) n = 0;
    for (int i = threadIdx.x; i < kMaxPeaksSort; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const int r = f[g];
        if (r != INT_MAX) {
            const int slot = atomicAdd(&n, 1);
            if (slot < kMaxPeaksSort) keys[slot] = ((unsigned long long)(unsigned)r << 32) | (unsigned)g;
            f[g] = INT_MAX;
        }
    }
    __syncthreads();
    const int total = n;
    const int m = total < kMaxPeaksSort ? total : kMaxPeaksSort;
    int sz = 1;
    while (sz < m) sz <<= 1;
    for (int size = 2; size <= sz; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (sz >> 1); t += blockDim.x) {
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
    for (int i = threadIdx.x; i < max_peaks; i += blockDim.x)
        peaks[(size_t)b * max_peaks + i] = (i < m) ? (int32_t)(keys[i] & 0xffffffffull) : -1;
    if (threadIdx.x == 0) count[b] = total;
}


}  // namespace
}  // namespace asw

using namespace asw;

extern "C" {

int asw_peaks_create(asw_peaks_t** out, int device, int Lx, int Ly, int Lz, int G, const int32_t* power_index,
                     const double* dis_matrix, const double* threshold3, double ratio2) {
    if (!out || !power_index || !dis_matrix || !threshold3 || Lx < 5 || Ly < 5 || Lz < 3 || G < 1) {
        set_error("asw_peaks_create: null argument or volume smaller than the 5x5x2 neighbourhood");
        return ASW_ERR_ARG;
    }
    *out = nullptr;
    DeviceGuard guard(device);      // the caller's current device is restored on return
    if (!guard.ok) {
        set_error("cannot make CUDA device %d current (no CUDA device, or a bad index)", device);
        return ASW_ERR_CUDA;
    }
    asw_peaks* h = new asw_peaks();
    h->device = device;
    h->Lx = Lx;
    h->Ly = Ly;
    h->Lz = Lz;
    h->G = G;
    for (int i = 0; i < 3; ++i) h->thr[i] = threshold3[i];
    h->ratio2 = ratio2;
    const size_t nv = (size_t)Lx * Ly * Lz, nc = (size_t)Lx * Ly;
    double* a1 = new double[nc];
    double* a2 = new double[nc];
    for (size_t i = 0; i < nc; ++i) {
        const double inv = 1.0 / dis_matrix[i];      // same two roundings as numpy's 0.9 + 1/d
        a1[i] = 0.9 + inv;
        a2[i] = 1.0 + inv;
    }
    cudaError_t e = cudaMalloc(&h->d_index, nv * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_a1, nc * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_a2, nc * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_index, power_index, nv * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_a1, a1, nc * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_a2, a2, nc * sizeof(double), cudaMemcpyHostToDevice);
    delete[] a1;
    delete[] a2;
    if (e != cudaSuccess) {
        set_error("asw_peaks_create: %s", cudaGetErrorString(e));
        asw_peaks_destroy(h);
        return ASW_ERR_CUDA;
    }
    *out = h;
    return ASW_OK;
}

int asw_peaks_destroy(asw_peaks_t* h) {
    if (!h) return ASW_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->d_index);
    cudaFree(h->d_a1);
    cudaFree(h->d_a2);
    cudaFree(h->d_first);
    delete h;
    return ASW_OK;
}

int asw_peaks_find(asw_peaks_t* h, const float* map_dev, int B, int32_t* peaks_dev, int max_peaks, int32_t* count_dev,
                   float* max_power_dev, void* stream) {
    if (!h || !map_dev || !peaks_dev || !count_dev || !max_power_dev || B < 1 || max_peaks < 1) {
        set_error("asw_peaks_find: null argument or bad shape");
        return ASW_ERR_ARG;
    }
    if (B > 65535) {
        set_error("asw_peaks_find: B=%d exceeds the grid limit 65535; split the batch", B);
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (B > h->Bcap) {
        if (h->d_first) cudaFree(h->d_first);
        h->d_first = nullptr;
        h->Bcap = 0;
        ASW_CUDA_CHECK(cudaMalloc(&h->d_first, (size_t)B * h->G * sizeof(int)));
        h->Bcap = B;
    }
    ASW_CARVE_ONCE(map_max_kernel);
    ASW_CARVE_ONCE(peak_flag_kernel);
    ASW_CARVE_ONCE(peak_collect_kernel);
    map_max_kernel<<<B, 1024, 0, s>>>(map_dev, h->G, max_power_dev, h->d_first);
    ASW_LAUNCH_CHECK("map_max_kernel");
    PeakParams p{};
    p.map = map_dev;
    p.index = h->d_index;
    p.a1 = h->d_a1;
    p.a2 = h->d_a2;
    p.mx = max_power_dev;
    p.first = h->d_first;
    p.Lx = h->Lx;
    p.Ly = h->Ly;
    p.Lz = h->Lz;
    p.G = h->G;
    p.t_ratio = h->thr[0];
    p.t_lo = h->thr[1];
    p.t_hi = h->thr[2];
    p.ratio2 = h->ratio2;
    const int n_int = (h->Lx - 4) * (h->Ly - 4) * (h->Lz - 2);
    dim3 grid((n_int + 255) / 256, B);
    peak_flag_kernel<<<grid, 256, 0, s>>>(p);
    ASW_LAUNCH_CHECK("peak_flag_kernel");
    peak_collect_kernel<<<B, 1024, 0, s>>>(h->d_first, h->G, peaks_dev, max_peaks, count_dev);
    ASW_LAUNCH_CHECK("peak_collect_kernel");
    return ASW_OK;
}

}  // extern "C"
