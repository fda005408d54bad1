// stft_cc.cu -- fused analysis-window framing + rectangular-window STFT (2048-point real FFT as a
// 1024-point complex Stockham radix-4 autosort FFT + split) + per-channel PHAT + pair cross-spectra.
//
// Reference arithmetic (sep/Traditional_SP/SRP_Prunning.py):
//   :403      seg = signal[:, j*step : j*step + window]
//   :404-409  X[m] = stft.analysis(seg[m], 2048, 512).T           (rectangular window, A1)
//   :414-416  pX = X / max(|X|, tol)
//   :421-426  CC[k, (i,j)] = (1/Nf) sum_n pX[i,k,n] conj(pX[j,k,n]),  i<j row-major, k in [bin0, bin1)
// Only the scored bins are ever formed.  One CTA owns one (mixture, window, frame group) and all M
// mics of it: the per-frame pX tile lives in shared memory, the pair sums live in registers (M <= 8)
// and are written once per CTA as a partial sum cc_part[b][w][group][p][f]; the 1/Nf scaling and
// the sum over groups are folded into the consumer (gcc.cu), so the result is deterministic.
#include "common.cuh"

namespace asw {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void radix4(float2& v0, float2& v1, float2& v2, float2& v3) {
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), d = csub(v1, v3);
    const float2 a3 = make_float2(d.y, -d.x);  // -i * (v1 - v3)
    v0 = cadd(a0, a2);
    v1 = cadd(a1, a3);
    v2 = csub(a0, a2);
    v3 = csub(a1, a3);
}

// One Stockham radix-4 pass over 1024 points with NS = 4^pass sub-transform length.
// Thread j reads in[j + 256 r], twiddles by exp(-2 pi i r (j mod NS) / (4 NS)), writes autosorted.
template <int NS>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out,
                                              const float2* __restrict__ tw, int j) {
    float2 v0 = in[j], v1 = in[j + 256], v2 = in[j + 512], v3 = in[j + 768];
    const int kk = j & (NS - 1);
    const int s = kk * (256 / NS);
    v1 = cmul(v1, tw[s]);
    v2 = cmul(v2, tw[2 * s]);
    v3 = cmul(v3, tw[3 * s]);
    radix4(v0, v1, v2, v3);
    const int d = ((j - kk) << 2) + kk;
    out[d] = v0;
    out[d + NS] = v1;
    out[d + 2 * NS] = v2;
    out[d + 3 * NS] = v3;
}

template <int MT>
__global__ void __launch_bounds__(kThreads) stft_cc_kernel(StftCcParams p) {
    __shared__ float2 bufA[kNc];
    __shared__ float2 bufB[kNc];
    __shared__ float2 s_tw[kNc];
    extern __shared__ float2 s_px[];  // [M][F]

    const int j = threadIdx.x;
    const int grp = blockIdx.x, w = blockIdx.y, b = blockIdx.z;
    const int M = (MT > 0) ? MT : p.M;
    const int F = p.F;

    for (int i = j; i < kNc; i += kThreads) s_tw[i] = p.tw1024[i];
    const float2 post = (j < F) ? p.twpost[j] : make_float2(0.f, 0.f);
    __syncthreads();

    constexpr int PT = (MT > 0) ? MT * (MT - 1) / 2 : 1;
    float2 acc[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) acc[q] = make_float2(0.f, 0.f);

    float2* cc_out = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG + grp) * (size_t)F) * p.P;
    const int n0 = grp * p.FG;
    const int n1 = min(p.Nf, n0 + p.FG);

    for (int n = n0; n < n1; ++n) {
        for (int m = 0; m < M; ++m) {
            const float* x = p.mix + ((size_t)b * p.M + m) * (size_t)p.T + (size_t)w * p.step + (size_t)n * kHop;
            // pass 0 straight from global: z[t] = x[2t] + i x[2t+1]
            float2 v0, v1, v2, v3;
            const int valid = p.win_len - n * kHop;   // < nfft only for the ragged last frame of the tail-padded mode
            if (valid < kNfft) {
                auto ld = [&](int i) { return i < valid ? __ldg(x + i) : 0.f; };
                v0 = make_float2(ld(2 * j), ld(2 * j + 1));
                v1 = make_float2(ld(2 * (j + 256)), ld(2 * (j + 256) + 1));
                v2 = make_float2(ld(2 * (j + 512)), ld(2 * (j + 512) + 1));
                v3 = make_float2(ld(2 * (j + 768)), ld(2 * (j + 768) + 1));
            } else if ((reinterpret_cast<uintptr_t>(x) & 7) == 0) {
                const float2* z = reinterpret_cast<const float2*>(x);
                v0 = __ldg(z + j);
                v1 = __ldg(z + j + 256);
                v2 = __ldg(z + j + 512);
                v3 = __ldg(z + j + 768);
            } else {
                v0 = make_float2(__ldg(x + 2 * j), __ldg(x + 2 * j + 1));
                v1 = make_float2(__ldg(x + 2 * (j + 256)), __ldg(x + 2 * (j + 256) + 1));
                v2 = make_float2(__ldg(x + 2 * (j + 512)), __ldg(x + 2 * (j + 512) + 1));
                v3 = make_float2(__ldg(x + 2 * (j + 768)), __ldg(x + 2 * (j + 768) + 1));
            }
            radix4(v0, v1, v2, v3);
            bufA[4 * j] = v0;
            bufA[4 * j + 1] = v1;
            bufA[4 * j + 2] = v2;
            bufA[4 * j + 3] = v3;
            __syncthreads();
            stockham_pass<4>(bufA, bufB, s_tw, j);
            __syncthreads();
            stockham_pass<16>(bufB, bufA, s_tw, j);
            __syncthreads();
            stockham_pass<64>(bufA, bufB, s_tw, j);
            __syncthreads();
            stockham_pass<256>(bufB, bufA, s_tw, j);
            __syncthreads();
            // split the packed transform into the real-input spectrum, scored bins only, then PHAT
            if (j < F) {
                const int k = p.bin0 + j;
                const float2 zk = bufA[k];
                const float2 zc = bufA[(kNc - k) & (kNc - 1)];
                const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));   // (Zk + conj Zn)/2
                const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));  // (Zk - conj Zn)/(2i)
                float2 X = cadd(e, cmul(post, o));
                float mag = sqrtf(fmaf(X.x, X.x, X.y * X.y));
                mag = fmaxf(mag, p.tol);
                const float inv = 1.0f / mag;
                s_px[m * F + j] = make_float2(X.x * inv, X.y * inv);
            }
            __syncthreads();  // bufA is rewritten by the next pass 0; s_px[m] is published
        }
        if (j < F) {
            if (MT > 0) {
                float2 a[(MT > 0) ? MT : 1];
#pragma unroll
                for (int m = 0; m < MT; ++m) a[m] = s_px[m * F + j];
                int q = 0;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int jj = i + 1; jj < MT; ++jj) {
                        acc[q].x += fmaf(a[i].x, a[jj].x, a[i].y * a[jj].y);
                        acc[q].y += fmaf(a[i].y, a[jj].x, -a[i].x * a[jj].y);
                        ++q;
                    }
            } else {
                // generic mic count: this CTA exclusively owns cc_out, accumulate there
                int q = 0;
                for (int i = 0; i < M; ++i) {
                    const float2 ai = s_px[i * F + j];
                    for (int jj = i + 1; jj < M; ++jj) {
                        const float2 aj = s_px[jj * F + j];
                        float2 c = make_float2(fmaf(ai.x, aj.x, ai.y * aj.y), fmaf(ai.y, aj.x, -ai.x * aj.y));
                        float2* dst = cc_out + (size_t)q * F + j;
                        if (n > n0) {
                            const float2 old = *dst;
                            c.x += old.x;
                            c.y += old.y;
                        }
                        *dst = c;
                        ++q;
                    }
                }
            }
        }
    }
    if (MT > 0 && j < F) {
#pragma unroll
        for (int q = 0; q < PT; ++q) cc_out[(size_t)q * F + j] = acc[q];
    }
    if (MT == 0 && n1 <= n0 && j < F) {
        for (int q = 0; q < p.P; ++q) cc_out[(size_t)q * F + j] = make_float2(0.f, 0.f);
    }
}

template <int MT>
int launch_t(const StftCcParams& p, cudaStream_t s) {
    dim3 grid(p.NG, p.Nw, p.B);
    const size_t smem = (size_t)p.M * p.F * sizeof(float2);
    if (smem > 20 * 1024) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    }
    stft_cc_kernel<MT><<<grid, kThreads, smem, s>>>(p);
    ASW_LAUNCH_CHECK("stft_cc_kernel");
    return ASW_OK;
}

}  // namespace

int launch_stft_cc(const StftCcParams& p, cudaStream_t s) {
    if (p.F > kThreads) {
        set_error("stft_cc: %d scored bins exceed the kernel limit of %d", p.F, kThreads);
        return ASW_ERR_RANGE;
    }
    switch (p.M) {
        case 2: return launch_t<2>(p, s);
        case 3: return launch_t<3>(p, s);
        case 4: return launch_t<4>(p, s);
        case 5: return launch_t<5>(p, s);
        case 6: return launch_t<6>(p, s);
        case 7: return launch_t<7>(p, s);
        case 8: return launch_t<8>(p, s);
        default: return launch_t<0>(p, s);
    }
}

}  // namespace asw
