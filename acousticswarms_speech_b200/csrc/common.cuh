// common.cuh -- shared helpers for libasw (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdio.h>

#include "../../include/asw.h"

namespace asw {

constexpr int kNfft = 2048;          // STFT frame (sep/helpers/constants.py:27)
constexpr int kNc = kNfft / 2;       // complex FFT length after even/odd packing
constexpr int kHop = kNfft / 4;      // SRP_Prunning.py:406
constexpr int kMaxMics = 32;
constexpr int kFracBits = 20;        // lag position fixed point: Q12.20
constexpr int kMaxEntries = 1 << (32 - kFracBits);
constexpr int kNumSms = 148;         // B200

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ASW_CUDA_CHECK(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            asw::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                     \
            return ASW_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define ASW_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            asw::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return ASW_ERR_CUDA;                                                    \
        }                                                                           \
        asw::count_launch();                                                        \
    } while (0)

// Makes a handle's device current for the duration of an entry point (and restores the caller's).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
        if (cur != dev) {
            ok = cudaSetDevice(dev) == cudaSuccess;
            if (ok) prev = cur;
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// cudaFuncSetAttribute is per device: true the first time a kernel is configured on the current device.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        return !(mask.fetch_or(bit) & bit);
    }
};

// One shared-memory carve-out for every kernel of the pipelined path (ASW_CARVE=<percent of the maximum>): an SM changes its L1 / shared
// split only when it is empty, so CTAs of kernels that prefer different splits never share an SM and streams that
// should overlap take turns instead.
int carve_all();
#define ASW_CARVE_ONCE(kernel)                                                                              \
    do {                                                                                                    \
        static asw::PerDeviceOnce _carve_once;                                                              \
        if (asw::carve_all() && _carve_once.need())                                                         \
            ASW_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,     \
                                                asw::carve_all()));                                         \
    } while (0)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// ---- kernel launchers (defined in the .cu files) ---------------------------------------------

struct StftCcParams {
    const float* mix;      // [B][M][T]
    float2* cc_part;       // [B][Nw][NG][P][F]  (bin fastest: coalesced for the writer and for gcc.cu)
    float2* px_out;        // [B][Nw][Nf][M][F] PHAT-normalised spectra (split path only, else null)
    const float2* tw1024;  // [1024]  exp(-2 pi i t / 1024)
    const float2* twpost;  // [F]     exp(-2 pi i k / 2048), k = bin0 + f
    int B, M, T, Nw, step, Nf, NG, FG, bin0, F, P;
    int win_len;           // samples of one analysis window; frame n holds min(nfft, win_len - n*hop) of them, zeros after
    float tol;
};
int launch_stft_cc(const StftCcParams& p, cudaStream_t s);          // generic (any M <= 32)
bool stft_cc_warp_supported(const StftCcParams& p);                  // fast path: M <= 8, bins in [1, 224)
int stft_cc_warp_ctas_per_sm(int M);
int launch_stft_cc_warp(const StftCcParams& p, cudaStream_t s);
bool stft_split_supported(const StftCcParams& p);                    // FFT + PHAT to global, then pair products: any M <= 32
int launch_stft_split(const StftCcParams& p, cudaStream_t s);

struct GccParams {
    const float2* cc_part;  // [B][Nw][NG][P][F]
    float* gcc;             // [B][tab_len * Nw]  pair-major: pair p at Nw*off[p], then [Nw][npad[p]]
    float2* cc_out;         // optional [B][Nw][F][P] summed + 1/Nf scaled (parity tap), may be null
    const int* lag_lo;      // [P] first lag (integer samples) of pair p's table
    const int* n_entries;   // [P] valid entries
    const int* npad;        // [P] entries padded to a multiple of 4
    const int* off;         // [P] float offset of the pair segment (per window unit)
    const float* fir;       // [(U-1)][12] Lagrange upsampling weights, nodes -5..+6, fraction fr/U
    const float2* tw1024;   // [1024] forward twiddles (FFT path), may be null
    const float2* twpost;   // [F]    exp(-i pi k / 1024), k = bin0 + f (FFT path), may be null
    int max_int_lags;       // largest integer-lag count (n_int) over the pairs
    int B, Nw, NG, F, P, bin0, U, tab_len;
    float inv_nf, scale;    // 1/Nf ; 1/(F*P)
};
int launch_gcc(const GccParams& p, cudaStream_t s);
int launch_patch_powers(float* x, int N, int T, int W, int demean, float* mean_out, float* power_out,
                        float* maxavg_out, int* argmax_out, cudaStream_t s);

struct SrpGatherParams {
    const float* gcc;       // as above
    const uint32_t* pos;    // [P][Gpad] Q12.20 position inside the pair's table, in slot order
    const int* perm;        // [Gpad] slot -> hypercube index (-1 = padding)
    const int* npad;        // [P]
    const int* off;         // [P]
    // staging plan of this (tile, windows-per-chunk) combination: a CTA stages, per pair, only the table entries its
    // hypercubes' 4-tap windows touch
    const int* rng_lo;      // [ntiles][P] first staged entry (multiple of 4)
    const int* rng_n;       // [ntiles][P] staged entries (multiple of 4)
    const int* tile_grp;    // [ntiles + 1] offset of each tile's group list in grp_flat
    const int* grp_flat;    // per tile: pair boundaries of its staging groups (n_groups + 1 entries)
    float* map;             // [B][G]
    int B, G, Gpad, P, Nw, tab_len;
    int tile;               // hypercubes per CTA
    int stage_floats;       // floats per stage buffer (two stages)
};
int launch_srp_gather(const SrpGatherParams& p, cudaStream_t s);      // p.tile / p.stage_floats from the plan
int srp_gather_windows_per_chunk();
int srp_gather_smem_budget();
int srp_gather_choose_tile(int G, int B, int P, int tab_len);

int launch_topk(const float* map, int B, int G, int K, int idx_offset, float* val, int32_t* idx, cudaStream_t s);

int launch_shift_stack(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M,
                       int T, float* out, cudaStream_t s);
int launch_pcm16_to_f32(const short* in, float* out, size_t n, cudaStream_t s);
int launch_shift_stack_counted(const float* mix, const int32_t* shifts, const int32_t* mix_index, const int32_t* n_valid,
                               int n_base, int N, int B, int M, int T, float* out, cudaStream_t s);
int launch_shift_stack_norm(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B,
                            int M, int T, float* out, float* means, float* stds, double* work, const double* tables,
                            int table_stride, int max_lag, const int32_t* n_valid, int n_base, cudaStream_t s);
int launch_shift_stack_norm_grouped(const float* mix, const int32_t* shifts, const int32_t* mix_index, int N, int B, int M,
                                    int T, float* out, float* means, float* stds, double* work, int32_t* ranges,
                                    const int32_t* n_valid, int n_base, cudaStream_t s);

}  // namespace asw
