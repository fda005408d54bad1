// xcorr.cu -- per-mixture circular cross-correlation tables of the re-quantised channels: what the fused
// shift-stack + normalize_input needs to know about a patch BEFORE it writes it, without reading the patch.
//
// normalize_input (sep/training/SpeakerLocalization/network.py:28-40) divides every patch by the standard
// deviation of its mic-average ref[t] = 1/M sum_c q_c[(t + r_c) mod T], q = round(x 2^15) / 2^15.  Both moments
// of ref are functions of per-MIXTURE quantities:
//     sum_t ref    = 1/M   sum_c S_c                                   S_c = sum_t q_c[t]          (shift invariant)
//     sum_t ref^2  = 1/M^2 (sum_c E_c + 2 sum_{c<c'} R_cc'(r_c' - r_c))  E_c = sum_t q_c[t]^2
//     R_cc'(l)     = sum_t q_c[t] q_c'[(t + l) mod T]                  (circular cross-correlation)
// and a hypercube patch only ever asks for |l| <= (array aperture in samples + hypercube width).  So one table
// per mixture, S_c, E_c and R_cc'(l) for |l| <= L, replaces the statistics pass that re-read all M T samples of
// every patch (shift_ref_stats_kernel: 100 us of the 195 us a fused 128-patch batch took in round 1).
//
// R is computed by overlap-save in the frequency domain with the register warp-FFT of fft32.cuh:
//   block s of channel c (role A): a_s[k] = q_c[s Nb + k], k < Nb = 2048 - 2L, zero padded to 2048
//   block s of channel c' (role B): b_s[k] = q_c'[(s Nb - L + k) mod T], k < 2048
//   sum_k a_s[k] b_s[k + j] for j in [0, 2L] never wraps inside the 2048-point circular correlation, so
//   R_cc'(j - L) = IFFT_2048( sum_s conj(A_s) B_s )[j].
//   K1 xcorr_fft_kernel    one warp per (mixture, block, role/channel): quantise, FFT-2048 (packed 1024-point complex),
//                          real-input split for ALL bins -> spectra in global memory
//   K2 xcorr_pair_kernel   thread = bin, 4 x 8 pair tiles in registers, blocks summed in groups -> partial cross-spectra
//   K3 xcorr_inverse_kernel one warp per (mixture, pair): sum the groups, inverse real FFT, lags -L..L as float64
//   S_c and E_c (float64, exact: 16-bit values, 30-bit products) ride along: every sample belongs to exactly one
//   role-A block (the last channel: to the centre of one role-B block), K1 leaves per-block sums, K3 adds them up.
// The butterflies run in fp32: the table entries carry ~3e-7 of sqrt(E_c E_c') of round-off, i.e. a few 1e-7
// relative on the standard deviation; the patch statistics kernel (shift_stack.cu) falls back to the exact pass
// for any patch whose lag is outside the table or whose variance is small against that round-off.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fft32.cuh"

struct asw_corr {
    int device = 0, M = 0, P = 0, L = 0, Nb = 0;
    float2* d_tw1024 = nullptr;   // [1024] exp(-2 pi i t / 1024)
    float2* d_tw2048 = nullptr;   // [1024] exp(-2 pi i k / 2048)
    float2* d_spec = nullptr;     // [chunk][nblk][2(M-1)][1024] spectra of one chunk of mixtures
    size_t spec_cap = 0;
    float2* d_part = nullptr;     // [chunk][groups][P][1024] partial cross-spectra
    size_t part_cap = 0;
    double* d_sums = nullptr;     // [chunk][nblk][M][2] per-block sum q, sum q^2
    size_t sums_cap = 0;
};

namespace asw {
namespace {

constexpr int kXWarps = 4;                     // FFT warps per CTA
constexpr int kTile = 2 * 32 * 33;             // floats of one warp's transpose tile
constexpr int kBlocksPerGroup = 8;             // blocks summed by one pair CTA
constexpr int kPairThreads = 256;
constexpr int kTI = 4, kTJ = 8, kMaxTiles = 24;
constexpr size_t kSpecBudget = 512u << 20;     // spectra workspace per chunk of mixtures (see asw_corr_tables)

__device__ __forceinline__ float quant16(float x) { return rintf(x * 32768.f) * (1.f / 32768.f); }

struct XP {
    const float* mix;        // [B][M][T], first mixture of the chunk
    float2* spec;
    float2* part;
    double* sums;            // [Bc][nblk][M][2]
    double* tables;          // [B][stride], first mixture of the chunk
    const float2* tw1024;
    const float2* tw2048;
    int Bc, M, T, L, Nb, nblk, ngrp, U2, P, stride;
    int vec;                 // rows are 8-byte aligned and T is even: float2 loads
};

// K1: one warp = one 2048-sample block of one channel in one role.
__global__ void __launch_bounds__(32 * kXWarps) xcorr_fft_kernel(XP p) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = smem + warp * kTile;
    const long long wid = (long long)blockIdx.x * kXWarps + warp;
    const long long total = (long long)p.Bc * p.nblk * p.U2;
    if (wid >= total) return;                                  // whole warp
    const int u = (int)(wid % p.U2);
    const int s = (int)((wid / p.U2) % p.nblk);
    const int b = (int)(wid / ((long long)p.U2 * p.nblk));
    const bool roleB = u >= p.M - 1;
    const int c = roleB ? u - (p.M - 1) + 1 : u;
    const float* row = p.mix + ((size_t)b * p.M + c) * (size_t)p.T;
    const int T = p.T;

    c64 v[32];
    double sum = 0.0, sq = 0.0;                                 // this block's share of S_c and E_c
    if (!roleB) {
        const int start = s * p.Nb;
        const int nvalid = min(p.Nb, T - start);                // samples of this block that exist
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int i = 2 * (32 * q + lane);
            float x0 = 0.f, x1 = 0.f;
            if (i + 1 < nvalid && p.vec) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(row + start + i));
                x0 = t.x;
                x1 = t.y;
            } else {
                if (i < nvalid) x0 = __ldg(row + start + i);
                if (i + 1 < nvalid) x1 = __ldg(row + start + i + 1);
            }
            x0 = quant16(x0);
            x1 = quant16(x1);
            v[q] = pk(x0, x1);
            if (2 * 32 * q < p.Nb) {                            // beyond Nb the block is zero padding
                sum += (double)x0 + (double)x1;
                sq += (double)x0 * (double)x0 + (double)x1 * (double)x1;
            }
        }
    } else {
        const int start = s * p.Nb - p.L;                       // may be negative (first block) or run past T (last)
        const bool last_ch = c == p.M - 1;                      // the last channel has no role-A blocks
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int k = 2 * (32 * q + lane);
            int i = start + k;
            if (i < 0) i += T;
            if (i >= T) i -= T;                                 // T >= 4096: one wrap is enough
            float x0, x1;
            if (p.vec) {                                        // i even, T even: the pair never straddles the wrap
                const float2 t = __ldg(reinterpret_cast<const float2*>(row + i));
                x0 = t.x;
                x1 = t.y;
            } else {
                x0 = __ldg(row + i);
                x1 = __ldg(row + (i + 1 == T ? 0 : i + 1));
            }
            x0 = quant16(x0);
            x1 = quant16(x1);
            v[q] = pk(x0, x1);
            if (last_ch) {                                      // centre [L, L + Nb) of the block = samples s Nb .. of the signal
                const int g0 = s * p.Nb + k - p.L;
                if (k >= p.L && k < p.L + p.Nb && g0 < T) {
                    sum += (double)x0;
                    sq += (double)x0 * (double)x0;
                }
                if (k + 1 >= p.L && k + 1 < p.L + p.Nb && g0 + 1 < T) {
                    sum += (double)x1;
                    sq += (double)x1 * (double)x1;
                }
            }
        }
    }
    if (!roleB || c == p.M - 1) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            sq += __shfl_xor_sync(0xffffffffu, sq, d);
        }
        if (lane == 0) {
            double* o = p.sums + (((size_t)b * p.nblk + s) * p.M + c) * 2;
            o[0] = sum;
            o[1] = sq;
        }
    }
    const float2 w1 = __ldg(p.tw1024 + lane);
    const float2 w8 = __ldg(p.tw1024 + ((8 * lane) & 1023));
    const float2 w16 = __ldg(p.tw1024 + ((16 * lane) & 1023));
    const float2 w24 = __ldg(p.tw1024 + ((24 * lane) & 1023));
    dft32(v);
    twiddle_and_transpose(v, tile, lane, w1, w8, w16, w24);
    __syncwarp();
    load_transposed(v, tile, lane);
    dft32(v);                                                   // Z[lane + 32 k2] at v[bitrev5(k2)]

    // real-input split for every bin k = lane + 32 k2 (stft_cc_warp.cu does the same for its 198 bins):
    // X[k] = e + W_2048^k o,  e = (Z[k] + conj Z[1024-k]) / 2,  o = -i (Z[k] - conj Z[1024-k]) / 2.
    // Bin 0 carries (X[0], X[1024]), both real.
    float2* out = p.spec + (size_t)wid * kNc;
    const int partner = (32 - lane) & 31;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const int k = lane + 32 * k2;
        const float2 zk = upk(v[bitrev5(k2)]);
        const float2 give = upk(v[bitrev5(31 - k2)]);
        float2 zc = make_float2(__shfl_sync(0xffffffffu, give.x, partner), __shfl_sync(0xffffffffu, give.y, partner));
        if (lane == 0) zc = upk(v[bitrev5((32 - k2) & 31)]);
        const float2 post = __ldg(p.tw2048 + k);
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
        float2 X = cadd(e, cmul(post, o));
        if (k == 0) X = make_float2(zk.x + zk.y, zk.x - zk.y);
        out[k] = X;
    }
}

struct Tiles {
    int n;
    unsigned char i0[kMaxTiles], j0[kMaxTiles];
};

// K2: conj(A_i) B_j summed over the blocks of one group; thread = bin, the tile's 32 accumulators in registers.
__global__ void __launch_bounds__(kPairThreads, 2) xcorr_pair_kernel(XP p, Tiles tiles) {   // <= 128 registers: 2 CTAs per SM
    const int k = blockIdx.x * kPairThreads + threadIdx.x;      // bin 0..1023
    const int grp = blockIdx.y;
    const int t = blockIdx.z % tiles.n, b = blockIdx.z / tiles.n;
    const int i0 = tiles.i0[t], j0 = tiles.j0[t];
    const int M = p.M;
    float2 acc[kTI][kTJ];
#pragma unroll
    for (int a = 0; a < kTI; ++a)
#pragma unroll
        for (int c = 0; c < kTJ; ++c) acc[a][c] = make_float2(0.f, 0.f);
    const int s0 = grp * kBlocksPerGroup, s1 = min(p.nblk, s0 + kBlocksPerGroup);
    for (int s = s0; s < s1; ++s) {
        const float2* sp = p.spec + ((size_t)b * p.nblk + s) * p.U2 * kNc + k;
        float2 A[kTI], Bv[kTJ];
#pragma unroll
        for (int a = 0; a < kTI; ++a)                            // role A of channel i: u = i (i <= M-2)
            A[a] = (i0 + a < M - 1) ? sp[(size_t)(i0 + a) * kNc] : make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < kTJ; ++c)                            // role B of channel j: u = (M-1) + j - 1 (j >= 1)
            Bv[c] = (j0 + c >= 1 && j0 + c < M) ? sp[(size_t)(M - 2 + j0 + c) * kNc] : make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < kTI; ++a)
#pragma unroll
            for (int c = 0; c < kTJ; ++c)
                if (i0 + a < j0 + c && j0 + c < M) {
                    if (k == 0) {                                // packed (X[0], X[1024]): two real products
                        acc[a][c].x = fmaf(A[a].x, Bv[c].x, acc[a][c].x);
                        acc[a][c].y = fmaf(A[a].y, Bv[c].y, acc[a][c].y);
                    } else {                                     // conj(A) B
                        acc[a][c].x += fmaf(A[a].x, Bv[c].x, A[a].y * Bv[c].y);
                        acc[a][c].y += fmaf(A[a].x, Bv[c].y, -A[a].y * Bv[c].x);
                    }
                }
    }
    float2* out = p.part + (((size_t)b * p.ngrp + grp) * p.P) * kNc + k;
#pragma unroll
    for (int a = 0; a < kTI; ++a)
#pragma unroll
        for (int c = 0; c < kTJ; ++c) {
            const int i = i0 + a, j = j0 + c;
            if (i < j && j < M) out[(size_t)(i * M - i * (i + 1) / 2 + (j - i - 1)) * kNc] = acc[a][c];
        }
}

// K3: one CTA per (mixture, pair): every thread sums one bin over the block groups, then warp 0 runs the inverse real
// FFT of length 2048 (the two-for-one packing of gcc.cu's gcc_fft_kernel over the full spectrum) and keeps lags -L..L.
// CTAs P .. P+M-1 of a mixture add up the per-block channel sums instead.
constexpr int kInvThreads = 512;               // 2 bins per thread; 1024 threads would cap the FFT warp at 64 registers
__global__ void __launch_bounds__(kInvThreads) xcorr_inverse_kernel(XP p) {
    __shared__ __align__(16) float tile[kTile];
    __shared__ float2 s_x[kNc];
    const int lane = threadIdx.x & 31;
    const int pr = blockIdx.x, b = blockIdx.y;
    if (pr >= p.P) {                                             // S_c and E_c from the per-block sums
        if (threadIdx.x >= 32) return;
        const int c = pr - p.P;
        double a = 0.0, e = 0.0;
        for (int s = lane; s < p.nblk; s += 32) {
            const double* o = p.sums + (((size_t)b * p.nblk + s) * p.M + c) * 2;
            a += o[0];
            e += o[1];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, d);
            e += __shfl_xor_sync(0xffffffffu, e, d);
        }
        if (lane == 0) {
            p.tables[(size_t)b * p.stride + c] = a;
            p.tables[(size_t)b * p.stride + p.M + c] = e;
        }
        return;
    }
    {
        const int k = threadIdx.x;
        const float2* src = p.part + (((size_t)b * p.ngrp) * p.P + pr) * kNc + k;
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
#pragma unroll 6
        for (int g = 0; g < p.ngrp; ++g) {
            const float2 v0 = src[(size_t)g * p.P * kNc], v1 = src[(size_t)g * p.P * kNc + kInvThreads];
            s0.x += v0.x;
            s0.y += v0.y;
            s1.x += v1.x;
            s1.y += v1.y;
        }
        s_x[k] = s0;
        s_x[k + kInvThreads] = s1;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    // Z[k] = (X[k] (1 + i t_k) + conj(X[1024-k]) (1 - i t_k)) / 2,  t_k = exp(i pi k / 1024);  v = conj Z
    c64 v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        const int k = 32 * q + lane;
        float2 z;
        if (k == 0) {
            const float2 x = s_x[0];                             // (X[0], X[1024])
            z = make_float2(0.5f * (x.x + x.y), 0.5f * (x.x - x.y));
        } else {
            const float2 x = s_x[k], y = s_x[kNc - k];
            const float2 tw = __ldg(p.tw2048 + k);               // exp(-i pi k / 1024) = conj t_k
            const float tc = tw.x, ts = -tw.y;
            const float2 a = cmul(x, make_float2(1.f - ts, tc));
            const float2 c2 = cmul(make_float2(y.x, -y.y), make_float2(1.f + ts, -tc));
            z = make_float2(0.5f * (a.x + c2.x), 0.5f * (a.y + c2.y));
        }
        v[q] = pk(z.x, -z.y);
    }
    const float2 w1 = __ldg(p.tw1024 + lane);
    const float2 w8 = __ldg(p.tw1024 + ((8 * lane) & 1023));
    const float2 w16 = __ldg(p.tw1024 + ((16 * lane) & 1023));
    const float2 w24 = __ldg(p.tw1024 + ((24 * lane) & 1023));
    dft32(v);
    twiddle_and_transpose(v, tile, lane, w1, w8, w16, w24);
    __syncwarp();
    load_transposed(v, tile, lane);
    dft32(v);                                                    // D[lane + 32 k2] at v[bitrev5(k2)]; z[n] = conj(D[n]) / 1024
    double* tab = p.tables + (size_t)b * p.stride + 2 * p.M + (size_t)pr * (2 * p.L + 1);
    const int nl = 2 * p.L + 1;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const float2 o = upk(v[bitrev5(k2)]);
        const int j = 2 * (lane + 32 * k2);                      // c(j) = Re z[n], c(j + 1) = Im z[n];  lag = j - L
        if (j < nl) tab[j] = (double)o.x * (1.0 / 1024.0);
        if (j + 1 < nl) tab[j + 1] = -(double)o.y * (1.0 / 1024.0);
    }
}

template <typename T>
int grow(T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return ASW_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need * sizeof(T));
    if (e != cudaSuccess) {
        set_error("asw_corr: device allocation of %zu bytes failed: %s", need * sizeof(T), cudaGetErrorString(e));
        return ASW_ERR_ALLOC;
    }
    *cap = need;
    return ASW_OK;
}

}  // namespace
}  // namespace asw

using namespace asw;

extern "C" {

int asw_corr_create(asw_corr_t** out, int device, int M, int max_lag) {
    if (!out) {
        set_error("asw_corr_create: null argument");
        return ASW_ERR_ARG;
    }
    *out = nullptr;
    if (M < 2 || M > kMaxMics || max_lag < 1 || max_lag > 512) {
        set_error("asw_corr_create: M=%d (2..%d) / max_lag=%d (1..512) unsupported", M, kMaxMics, max_lag);
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(device);      // the caller's current device is restored on return
    if (!guard.ok) {
        set_error("cannot make CUDA device %d current (no CUDA device, or a bad index)", device);
        return ASW_ERR_CUDA;
    }
    asw_corr* h = new asw_corr();
    h->device = device;
    h->M = M;
    h->P = M * (M - 1) / 2;
    h->L = max_lag;
    h->Nb = kNfft - 2 * max_lag;
    std::vector<float2> tw(kNc), tw2(kNc);
    for (int t = 0; t < kNc; ++t) {
        const double a = -2.0 * M_PI * (double)t / (double)kNc, a2 = -2.0 * M_PI * (double)t / (double)kNfft;
        tw[t] = make_float2((float)cos(a), (float)sin(a));
        tw2[t] = make_float2((float)cos(a2), (float)sin(a2));
    }
    cudaError_t e = cudaMalloc(&h->d_tw1024, sizeof(float2) * kNc);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_tw2048, sizeof(float2) * kNc);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tw1024, tw.data(), sizeof(float2) * kNc, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tw2048, tw2.data(), sizeof(float2) * kNc, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("asw_corr_create: %s", cudaGetErrorString(e));
        asw_corr_destroy(h);
        return ASW_ERR_CUDA;
    }
    *out = h;
    return ASW_OK;
}

int asw_corr_destroy(asw_corr_t* h) {
    if (!h) return ASW_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->d_tw1024);
    cudaFree(h->d_tw2048);
    cudaFree(h->d_spec);
    cudaFree(h->d_part);
    cudaFree(h->d_sums);
    delete h;
    return ASW_OK;
}

int asw_corr_table_len(const asw_corr_t* h) {
    if (!h) return 0;
    return 2 * h->M + h->P * (2 * h->L + 1);
}

int asw_corr_tables(asw_corr_t* h, const float* mix_dev, int B, int T, double* tables_dev, void* stream) {
    if (!h || !mix_dev || !tables_dev || B < 1) {
        set_error("asw_corr_tables: null handle/buffer or B < 1");
        return ASW_ERR_ARG;
    }
    if (T < 2 * kNfft) {
        set_error("asw_corr_tables: T=%d is shorter than %d samples (use asw_shift_stack_norm, the exact pass)", T,
                  2 * kNfft);
        return ASW_ERR_RANGE;
    }
    DeviceGuard guard(h->device);
    if (!guard.ok) {
        set_error("asw_corr_tables: cannot make device %d current", h->device);
        return ASW_ERR_CUDA;
    }
    cudaStream_t s = (cudaStream_t)stream;
    XP p{};
    p.tw1024 = h->d_tw1024;
    p.tw2048 = h->d_tw2048;
    p.M = h->M;
    p.T = T;
    p.L = h->L;
    p.Nb = h->Nb;
    p.nblk = (T + h->Nb - 1) / h->Nb;
    p.ngrp = (p.nblk + kBlocksPerGroup - 1) / kBlocksPerGroup;
    p.U2 = 2 * (h->M - 1);
    p.P = h->P;
    p.stride = asw_corr_table_len(h);
    p.vec = (T % 2 == 0) && ((reinterpret_cast<uintptr_t>(mix_dev) & 7) == 0);
    // Chunking: the spectra (14 MB per mixture at 7 mics / 3 s) are written by K1 and read once by K2.  Chunks small
    // enough to stay in L2 (3 mixtures) were measured slower than one large launch per 36 mixtures: 11 x 3 small
    // launches with tails against 3, while the extra HBM traffic is ~5 us per mixture.
    const size_t spec_per = (size_t)p.nblk * p.U2 * kNc;          // float2 per mixture
    int chunk = (int)(kSpecBudget / (spec_per * sizeof(float2)));
    Tiles tiles{};
    for (int i0 = 0; i0 < h->M - 1; i0 += kTI)
        for (int j0 = i0; j0 < h->M; j0 += kTJ) {
            tiles.i0[tiles.n] = (unsigned char)i0;
            tiles.j0[tiles.n] = (unsigned char)j0;
            ++tiles.n;
        }
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    if (chunk > 65535 / tiles.n) chunk = 65535 / tiles.n;        // gridDim.z of the pair kernel, gridDim.y of the inverse
    int rc;
    if ((rc = grow(&h->d_spec, &h->spec_cap, spec_per * chunk)) != ASW_OK) return rc;
    if ((rc = grow(&h->d_part, &h->part_cap, (size_t)chunk * p.ngrp * p.P * kNc)) != ASW_OK) return rc;
    if ((rc = grow(&h->d_sums, &h->sums_cap, (size_t)chunk * p.nblk * h->M * 2)) != ASW_OK) return rc;
    p.spec = h->d_spec;
    p.part = h->d_part;
    p.sums = h->d_sums;

    static PerDeviceOnce attr_once;
    const size_t smem_fft = (size_t)kXWarps * kTile * sizeof(float);
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(xcorr_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft));
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        p.Bc = (B - b0 < chunk) ? B - b0 : chunk;
        p.mix = mix_dev + (size_t)b0 * h->M * T;
        p.tables = tables_dev + (size_t)b0 * p.stride;
        const long long nw = (long long)p.Bc * p.nblk * p.U2;
        xcorr_fft_kernel<<<(unsigned)((nw + kXWarps - 1) / kXWarps), 32 * kXWarps, smem_fft, s>>>(p);
        ASW_LAUNCH_CHECK("xcorr_fft_kernel");
        xcorr_pair_kernel<<<dim3(kNc / kPairThreads, p.ngrp, p.Bc * tiles.n), kPairThreads, 0, s>>>(p, tiles);
        ASW_LAUNCH_CHECK("xcorr_pair_kernel");
        xcorr_inverse_kernel<<<dim3(p.P + h->M, p.Bc), kInvThreads, 0, s>>>(p);
        ASW_LAUNCH_CHECK("xcorr_inverse_kernel");
    }
    return ASW_OK;
}

}  // extern "C"
