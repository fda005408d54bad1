// gcc.cu -- band-limited GCC-PHAT lag tables from the pair cross-spectra.
//
// The reference never forms GCC curves: it contracts CC with a (G, F, P) steering table
// (sep/Traditional_SP/SRP_Prunning.py:228, :428-429).  Because
//     tab[g, f, p] = exp(+2 pi i k_f tau[g, p] / nfft),      tau = fs (d_i - d_j) / C   (:375-379)
// that contraction is, per pair, the band-limited inverse transform of CC[:, p] evaluated at the
// fractional lag tau[g, p]:
//     R_p(l) = 1/(F P) * Re sum_k CC[k, p] exp(+2 pi i k l / nfft).
// This kernel tabulates R_p on the U-times oversampled lag grid covering the lags the geometry can
// produce (plus interpolation margin); srp_gather.cu then interpolates.  One CTA per
// (pair, window, mixture), two stages:
//   A. R_p at the INTEGER lags (plus a 6-sample margin) into shared memory: each entry is a 4-way
//      split Horner evaluation of the polynomial in z = exp(2 pi i l / nfft), with exact integer
//      phase reduction for the seeds;
//   B. x U upsampling with a fixed 12-tap Lagrange interpolator (weights from the host).  R_p is
//      band-limited to bin1/nfft < 0.1 cycles/sample, so the 12-tap error (1.6e-6 of the map's max,
//      measured) is the same as evaluating every fractional lag directly, at 1/U of the work -- ncu
//      showed the direct version 87 % issue-bound.
#include "common.cuh"
#include "fft32.cuh"

namespace asw {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBins = 256;

__device__ __forceinline__ float2 cis_turns(int idx, int NU) {
    int r = idx % NU;
    if (r < 0) r += NU;
    float s, c;
    sincospif((float)r * (2.0f / (float)NU), &s, &c);
    return make_float2(c, s);
}

constexpr int kTaps = 12;       // upsampling interpolator length
constexpr int kMargin = 5;      // nodes -5 .. +6 around the integer lag below the target
constexpr int kMaxInt = kMaxEntries + kTaps + 4;

__global__ void __launch_bounds__(kThreads) gcc_kernel(GccParams p) {
    __shared__ float2 s_c[kMaxBins + 4];
    __shared__ float s_r1[kMaxInt];
    __shared__ float s_fir[7 * kTaps];
    const int tid = threadIdx.x;
    const int pr = blockIdx.x, w = blockIdx.y, b = blockIdx.z;
    const int F = p.F;
    const int Fpad = (F + 3) & ~3;
    const int U = p.U;

    for (int f = tid; f < Fpad; f += kThreads) {
        float2 s = make_float2(0.f, 0.f);
        if (f < F) {
            const float2* src = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG) * (size_t)p.P + pr) * F + f;
            for (int g = 0; g < p.NG; ++g) {
                const float2 v = src[(size_t)g * F * p.P];
                s.x += v.x;
                s.y += v.y;
            }
            s.x *= p.inv_nf;
            s.y *= p.inv_nf;
            if (p.cc_out) p.cc_out[(((size_t)b * p.Nw + w) * F + f) * p.P + pr] = s;
        }
        s_c[f] = s;
    }
    for (int i = tid; i < (U - 1) * kTaps; i += kThreads) s_fir[i] = p.fir[i];
    __syncthreads();

    const int lo = p.lag_lo[pr], n = p.n_entries[pr], npd = p.npad[pr];
    float* out = p.gcc + (size_t)b * p.tab_len * p.Nw + (size_t)p.Nw * p.off[pr] + (size_t)w * npd;
    const int nch = Fpad >> 2;
    const int n_int = (n - 1) / U + 1 + kTaps;   // integer lags lo - kMargin ... hi + 6

    // stage A: integer lags
    for (int j = tid; j < n_int; j += kThreads) {
        const int L = lo - kMargin + j;
        const float2 z = cis_turns(L, kNfft);
        const float2 z4 = cis_turns(4 * L, kNfft);
        const float2 zk = cis_turns(p.bin0 * L, kNfft);
        float2 a0 = s_c[4 * (nch - 1)], a1 = s_c[4 * (nch - 1) + 1];
        float2 a2 = s_c[4 * (nch - 1) + 2], a3 = s_c[4 * (nch - 1) + 3];
        for (int m = nch - 2; m >= 0; --m) {
            a0 = cadd(cmul(a0, z4), s_c[4 * m]);
            a1 = cadd(cmul(a1, z4), s_c[4 * m + 1]);
            a2 = cadd(cmul(a2, z4), s_c[4 * m + 2]);
            a3 = cadd(cmul(a3, z4), s_c[4 * m + 3]);
        }
        float2 t = cadd(a2, cmul(z, a3));
        t = cadd(a1, cmul(z, t));
        t = cadd(a0, cmul(z, t));
        s_r1[j] = (zk.x * t.x - zk.y * t.y) * p.scale;
    }
    __syncthreads();

    // stage B: x U upsampling; entry i is lag lo + i / U
    for (int i = tid; i < npd; i += kThreads) {
        float val = 0.f;
        if (i < n) {
            const int j = i / U, fr = i - j * U;
            const float* r = s_r1 + j;              // r[kMargin] is the integer lag at or below the target
            if (fr == 0) {
                val = r[kMargin];
            } else {
                const float* wt = s_fir + (fr - 1) * kTaps;
#pragma unroll
                for (int t = 0; t < kTaps; ++t) val = fmaf(wt[t], r[t], val);
            }
        }
        out[i] = val;
    }
}

// Stage A by FFT (one warp per curve): the integer lags of R_p are the inverse real FFT of length 2048 of the
// one-sided spectrum CC[:, p].  Packed as the usual two-for-one trick run backwards,
//     Z[k] = (X[k] (1 + i t_k) + conj(X[N-k]) (1 + i conj(t_{N-k}))) / 2,   t_k = exp(i pi k / 1024),  N = 1024,
//     z = IDFT_N(Z) = conj(DFT_N(conj Z)),   r[2n] = Re z[n],   r[2n+1] = Im z[n],
// so the warp-FFT of stft_cc_warp.cu (two register DFT-32 passes around one shared transpose) does the work:
// ~2 k warp-instructions per curve instead of ~16 k for the Horner evaluation of every lag (ncu: the Horner
// version was 87 % issue-bound).  R_p is periodic in 2048 samples, so lags index the result modulo 2048.
constexpr int kFftWarps = 4;
constexpr int kTileFloats = 2 * 32 * 33;

__global__ void __launch_bounds__(32 * kFftWarps, 5) gcc_fft_kernel(GccParams p) {   // 5 CTAs/SM: <= 96 registers
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ float s_fir[7 * kTaps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = s_dyn + warp * (kTileFloats + 2 * kMaxBins);
    float2* s_x = reinterpret_cast<float2*>(tile + kTileFloats);          // [F] spectrum of this curve
    const int pr = blockIdx.x * kFftWarps + warp, w = blockIdx.y, b = blockIdx.z;
    const int F = p.F, U = p.U;
    for (int i = threadIdx.x; i < (U - 1) * kTaps; i += blockDim.x) s_fir[i] = p.fir[i];
    __syncthreads();
    if (pr >= p.P) return;                                               // whole warp exits together

    {
        // frame-group partial sums of this curve's spectrum: all bins of a lane (<= 8) are loaded together for every
        // group, four groups in flight (the first version walked bins x groups one dependent load at a time: a third of
        // the kernel's samples sat on those loads); same order of additions per bin
        constexpr int kBinsPerLane = kMaxBins / 32;
        const float2* src = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG) * (size_t)p.P + pr) * F;
        const size_t gstride = (size_t)F * p.P;
        float2 sum[kBinsPerLane];
#pragma unroll
        for (int i = 0; i < kBinsPerLane; ++i) sum[i] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int g = 0; g < p.NG; ++g) {
#pragma unroll
            for (int i = 0; i < kBinsPerLane; ++i) {
                const int f = lane + 32 * i;
                if (f < F) {
                    const float2 v = src[(size_t)g * gstride + f];
                    sum[i].x += v.x;
                    sum[i].y += v.y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kBinsPerLane; ++i) {
            const int f = lane + 32 * i;
            if (f < F) {
                const float2 sc = make_float2(sum[i].x * p.inv_nf, sum[i].y * p.inv_nf);
                if (p.cc_out) p.cc_out[(((size_t)b * p.Nw + w) * F + f) * p.P + pr] = sc;
                s_x[f] = sc;
            }
        }
    }
    __syncwarp();

    // v[q] = conj(Z[32 q + lane]), packed (re, im) for the FFMA2 / FADD2 butterflies of fft32.cuh
    c64 v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        const int k = 32 * q + lane;
        float2 z = make_float2(0.f, 0.f);
        const int f1 = k - p.bin0;
        if (f1 >= 0 && f1 < F) {                                         // X[k] (1 + i t_k) / 2,  t_k = conj(twpost)
            const float2 x = s_x[f1], t = __ldg(p.twpost + f1);          // twpost = exp(-i pi k / 1024)
            const float2 m = make_float2(1.f + t.y, t.x);                // 1 + i conj(twpost) = (1 + sin, cos)... see below
            z = cadd(z, cmul(x, m));
        }
        const int kk = (kNc - k) & (kNc - 1);
        const int f2 = kk - p.bin0;
        if (f2 >= 0 && f2 < F && k != 0) {                               // conj(X[N-k]) (1 + i conj(t_{N-k})) / 2
            const float2 x = s_x[f2], t = __ldg(p.twpost + f2);
            const float2 xc = make_float2(x.x, -x.y);
            const float2 m = make_float2(1.f - t.y, t.x);                // 1 + i twpost
            z = cadd(z, cmul(xc, m));
        }
        v[q] = pk(0.5f * z.x, -0.5f * z.y);
    }
    const float2 w1 = __ldg(p.tw1024 + lane);
    const float2 w8 = __ldg(p.tw1024 + ((8 * lane) & 1023));
    const float2 w16 = __ldg(p.tw1024 + ((16 * lane) & 1023));
    const float2 w24 = __ldg(p.tw1024 + ((24 * lane) & 1023));
    dft32(v);
    twiddle_and_transpose(v, tile, lane, w1, w8, w16, w24);
    __syncwarp();
    load_transposed(v, tile, lane);
    __syncwarp();
    dft32(v);   // DFT(conj Z)[lane + 32 k2] at v[bitrev5(k2)];  z[n] = conj of it

    const int lo = p.lag_lo[pr], n = p.n_entries[pr], npd = p.npad[pr];
    const int n_int = (n - 1) / U + 1 + kTaps;                           // integer lags lo - kMargin ... (<= 2048 here)
    float* s_r1 = tile;                                                  // the tile is free again
    const int base = lo - kMargin;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const float2 o = upk(v[bitrev5(k2)]);
        const int l0 = 2 * (lane + 32 * k2);                             // lag of Re z[n]; Im z[n] is lag l0 + 1
        int j = (l0 - base) & (kNfft - 1);                               // R_p has period 2048
        if (j < n_int) s_r1[j] = o.x * p.scale;
        j = (l0 + 1 - base) & (kNfft - 1);
        if (j < n_int) s_r1[j] = -o.y * p.scale;
    }
    __syncwarp();

    float* out = p.gcc + (size_t)b * p.tab_len * p.Nw + (size_t)p.Nw * p.off[pr] + (size_t)w * npd;
    if (U == 4) {
        // a lane per INTEGER lag: 12 shared loads feed the lag's four table entries (the first version did a lane per
        // entry: 24 shared loads each), the weights stay in registers, one 16-byte store; same fma chains, same bits
        float wt[3][kTaps];
#pragma unroll
        for (int fr = 0; fr < 3; ++fr)
#pragma unroll
            for (int t = 0; t < kTaps; ++t) wt[fr][t] = s_fir[fr * kTaps + t];
        for (int j = lane; 4 * j < npd; j += 32) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * j < n) {
                float r[kTaps];
#pragma unroll
                for (int t = 0; t < kTaps; ++t) r[t] = s_r1[j + t];
                float v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
                for (int t = 0; t < kTaps; ++t) {
                    v1 = fmaf(wt[0][t], r[t], v1);
                    v2 = fmaf(wt[1][t], r[t], v2);
                    v3 = fmaf(wt[2][t], r[t], v3);
                }
                o.x = r[kMargin];
                o.y = 4 * j + 1 < n ? v1 : 0.f;
                o.z = 4 * j + 2 < n ? v2 : 0.f;
                o.w = 4 * j + 3 < n ? v3 : 0.f;
            }
            *reinterpret_cast<float4*>(out + 4 * j) = o;    // npad, off and tab_len are multiples of 4 entries
        }
        return;
    }
    for (int i = lane; i < npd; i += 32) {
        float val = 0.f;
        if (i < n) {
            const int j = i / U, fr = i - j * U;
            const float* r = s_r1 + j;
            if (fr == 0) {
                val = r[kMargin];
            } else {
                const float* wt = s_fir + (fr - 1) * kTaps;
#pragma unroll
                for (int t = 0; t < kTaps; ++t) val = fmaf(wt[t], r[t], val);
            }
        }
        out[i] = val;
    }
}

}  // namespace

int launch_gcc(const GccParams& p, cudaStream_t s) {
    if (p.F > kMaxBins) {
        set_error("gcc: %d scored bins exceed the kernel limit of %d", p.F, kMaxBins);
        return ASW_ERR_RANGE;
    }
    if (p.U < 1 || p.U > 8) {
        set_error("gcc: oversampling %d unsupported", p.U);
        return ASW_ERR_ARG;
    }
    // FFT path: every pair's integer-lag range must fit one period and the transpose tile
    if (p.tw1024 && p.twpost && p.max_int_lags <= kTileFloats && p.max_int_lags <= kNfft && p.bin0 + p.F <= kNc) {
        dim3 grid((p.P + kFftWarps - 1) / kFftWarps, p.Nw, p.B);
        const size_t smem = (size_t)kFftWarps * (kTileFloats + 2 * kMaxBins) * sizeof(float);
        static PerDeviceOnce attr_once;
        if (attr_once.need()) {
            ASW_CUDA_CHECK(cudaFuncSetAttribute(gcc_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        ASW_CARVE_ONCE(gcc_fft_kernel);
        gcc_fft_kernel<<<grid, 32 * kFftWarps, smem, s>>>(p);
        ASW_LAUNCH_CHECK("gcc_fft_kernel");
        return ASW_OK;
    }
    dim3 grid(p.P, p.Nw, p.B);
    gcc_kernel<<<grid, kThreads, 0, s>>>(p);
    ASW_LAUNCH_CHECK("gcc_kernel");
    return ASW_OK;
}

}  // namespace asw
