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
            const float2* src = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG) * (size_t)F + f) * p.P + pr;
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
    dim3 grid(p.P, p.Nw, p.B);
    gcc_kernel<<<grid, kThreads, 0, s>>>(p);
    ASW_LAUNCH_CHECK("gcc_kernel");
    return ASW_OK;
}

}  // namespace asw
