// stft_cc_warp.cu -- fast path of the fused STFT + PHAT + pair cross-spectra (M <= 8 mics, scored
// bins inside [1, 224)).  Same arithmetic as stft_cc.cu / the reference
// (sep/Traditional_SP/SRP_Prunning.py:403-426); different machine mapping:
//
//   * one WARP computes one 1024-point complex FFT (the even/odd packed 2048-sample frame) entirely
//     in registers: lane t holds z[32 q + t], q = 0..31, does a radix-32 DFT over q, applies the
//     W_1024^{t k1} twiddles, transposes through a padded per-warp shared tile (conflict-free, only
//     __syncwarp), and does the second radix-32 DFT over t.  Lane k1 then owns Z[k1 + 32 k2].
//   * only Z[k] for k < 224 and k > 800 feed the scored bins, so the second DFT is pruned to 14 of
//     its 32 outputs (dead-code elimination of the unrolled butterflies does the pruning).
//   * the real-input split needs Z[k] and Z[1024 - k]: lane (32 - k1) & 31 holds the partner, fetched
//     with warp shuffles.
//   * warp m of the CTA handles mic m; after the M FFTs of a frame one __syncthreads publishes the
//     PHAT-normalised bins, and thread k accumulates the P pair products of bin k in registers.
//
// ncu on the first version (radix-4 x 5 through shared memory, 256 threads per FFT) showed 92 % L1/
// shared throughput with 60 % of the shared wavefronts being bank-conflict replays; this version moves
// 8x less data through shared memory per FFT and has no conflicts.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fft32.cuh"

namespace asw {
namespace {

constexpr int kK2 = 7;  // second-pass outputs kept on each side: bins k1 + 32 k2, k2 < 7  (k < 224)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Asynchronously copy one 2048-sample frame (8 KB, contiguous) into this warp's tile, raw layout.
__device__ __forceinline__ void prefetch_frame(float* tile, const float* x, int lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(x);
    if ((a & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) cp_async16(tile + 4 * (lane + 32 * j), x + 4 * (lane + 32 * j));
    } else if ((a & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) cp_async8(tile + 2 * (lane + 32 * j), x + 2 * (lane + 32 * j));
    } else {
#pragma unroll 8
        for (int j = 0; j < 64; ++j) cp_async4(tile + (lane + 32 * j), x + (lane + 32 * j));
    }
    cp_async_commit();
}

// The ragged last frame of the tail-padded frame mode: only `valid` (< 2048) samples belong to the analysis
// window, the rest of the frame is zeros (the samples that follow in memory belong to the next window).
__device__ __forceinline__ void prefetch_partial_frame(float* tile, const float* x, int lane, int valid) {
    for (int i = lane; i < kNfft; i += 32) {
        if (i < valid) cp_async4(tile + i, x + i);
        else tile[i] = 0.f;
    }
    cp_async_commit();
}

__device__ __forceinline__ void prefetch_frame_n(float* tile, const float* xw, int n, int win_len, int lane) {
    const int valid = win_len - n * kHop;
    if (valid >= kNfft) prefetch_frame(tile, xw + (size_t)n * kHop, lane);
    else prefetch_partial_frame(tile, xw + (size_t)n * kHop, lane, valid);
}

template <int M>
struct Cfg {
    static constexpr int kThreads = 32 * M;
    static constexpr int kP = M * (M - 1) / 2;
    static constexpr int kNb = (200 + kThreads - 1) / kThreads;  // bins per thread in the pair-product phase
    static constexpr int kFmax = kThreads * kNb;
};

// SPLIT = false (M <= 8): warp m = mic m, pair products accumulated in registers, one partial cross-spectrum per CTA.
// SPLIT = true  (any M <= 32): the CTA's M_ = 8 warps take 8 mics of a chunk (blockIdx.z = b * chunks + chunk), the
// PHAT-normalised spectra go to global memory as pX[b][w][frame][mic][bin] and pair_products_kernel forms the
// cross-spectra: M(M-1)/2 accumulators per bin do not fit next to an FFT in registers beyond 8 mics.
template <int M, bool SPLIT>
// no __launch_bounds__ here: it would override the per-file -maxrregcount=128 (see build.py)
__global__ void stft_cc_warp_kernel(StftCcParams p) {
    using C = Cfg<M>;
    extern __shared__ __align__(16) float smem[];
    // per-warp transpose tile: packed complex [32][33] (8448 B, also the landing buffer of the next raw frame); then pX double buffer [2][M][F]
    float* tile = smem + (threadIdx.x >> 5) * (2 * 32 * 33);
    float2* s_px = reinterpret_cast<float2*>(smem + M * (2 * 32 * 33));
    const int lane = threadIdx.x & 31;
    const int tid = threadIdx.x;
    const int grp = blockIdx.x, w = blockIdx.y;
    const int mchunks = SPLIT ? (p.M + M - 1) / M : 1;
    const int b = SPLIT ? blockIdx.z / mchunks : blockIdx.z;
    const int m = (SPLIT ? (blockIdx.z % mchunks) * M : 0) + (threadIdx.x >> 5);  // this warp's mic
    if (SPLIT && m >= p.M) return;   // whole warp; the split path has no block-wide barrier
    const int F = p.F;
    const float tol2 = p.tol * p.tol;

    // per-lane twiddle seeds W_1024^{t j}, j = 1, 8, 16, 24 (t = lane)
    const float2 w1 = __ldg(p.tw1024 + lane);
    const float2 w8 = __ldg(p.tw1024 + ((8 * lane) & 1023));
    const float2 w16 = __ldg(p.tw1024 + ((16 * lane) & 1023));
    const float2 w24 = __ldg(p.tw1024 + ((24 * lane) & 1023));

    c64 acc[SPLIT ? 1 : C::kNb][SPLIT ? 1 : C::kP];   // packed (re, im): the pair products run on FFMA2 / FADD2
#pragma unroll
    for (int i = 0; i < (SPLIT ? 1 : C::kNb); ++i)
#pragma unroll
        for (int q = 0; q < (SPLIT ? 1 : C::kP); ++q) acc[i][q] = pk(0.f, 0.f);

    const int n0 = grp * p.FG;
    const int n1 = min(p.Nf, n0 + p.FG);
    const float* xm = p.mix + ((size_t)b * p.M + m) * (size_t)p.T + (size_t)w * p.step;

    // The tile is idle between the transposed read of frame n and the transposed write of frame n+1:
    // it doubles as the landing buffer of the cp.async prefetch of frame n+1 (raw layout, 8 KB), so
    // the global-load latency overlaps the second DFT pass, the split/PHAT and the pair products.
    if (n0 < n1) prefetch_frame_n(tile, xm, n0, p.win_len, lane);

    for (int n = n0; n < n1; ++n) {
        float2* px = SPLIT ? p.px_out + ((((size_t)b * p.Nw + w) * p.Nf + n) * p.M) * F
                           : s_px + (size_t)(n & 1) * M * F;
        {
            c64 v[32];
            cp_async_wait_all();
            __syncwarp();
            {
                const c64* raw = reinterpret_cast<const c64*>(tile);
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = raw[32 * q + lane];   // z[t] = x[2t] + i x[2t+1]
            }
            __syncwarp();  // every lane has its raw samples before the tile is overwritten
            dft32(v);  // Y[k1] at v[bitrev5(k1)]
            twiddle_and_transpose(v, tile, lane, w1, w8, w16, w24);
            __syncwarp();
            load_transposed(v, tile, lane);
            __syncwarp();  // the tile is free again: start fetching the next frame into it
            if (n + 1 < n1) prefetch_frame_n(tile, xm, n + 1, p.win_len, lane);
            dft32(v);      // Z[k1 + 32 k2] at v[bitrev5(k2)], k1 = lane
            // real-input split for bins k = lane + 32 k2, k2 < kK2, then PHAT
            const int partner = (32 - lane) & 31;
#pragma unroll
            for (int k2 = 0; k2 < kK2; ++k2) {
                const int k = lane + 32 * k2;
                const float2 zk = upk(v[bitrev5(k2)]);
                // partner value Z[1024 - k]: lane' = (32 - lane) & 31, k2' = 31 - k2 (lane != 0) or 32 - k2 (lane == 0)
                const float2 give = upk(v[bitrev5(31 - k2)]);
                float2 zc = make_float2(__shfl_sync(0xffffffffu, give.x, partner),
                                        __shfl_sync(0xffffffffu, give.y, partner));
                if (lane == 0) zc = upk(v[bitrev5((32 - k2) & 31)]);
                const int f = k - p.bin0;
                if (f >= 0 && f < F) {
                    const float2 post = __ldg(p.twpost + f);
                    const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                    const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
                    const float2 X = cadd(e, cmul(post, o));
                    // X / max(|X|, tol) as X * rsqrt(max(|X|^2, tol^2)): one MUFU.RSQ (2^-22 relative) instead of the
                    // IEEE sqrt + divide sequences with their slow-path branches
                    const float n2 = fmaxf(fmaf(X.x, X.x, X.y * X.y), tol2);
                    float inv;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(n2));
                    px[m * F + f] = make_float2(X.x * inv, X.y * inv);
                }
            }
        }
        if (SPLIT) continue;
        __syncthreads();  // all M mics of frame n are in px
#pragma unroll
        for (int i = 0; i < C::kNb; ++i) {
            const int f = tid + i * C::kThreads;
            if (f < F) {
                float2 a[M];
#pragma unroll
                for (int mm = 0; mm < M; ++mm) a[mm] = px[mm * F + f];
                int q = 0;
#pragma unroll
                for (int ii = 0; ii < M; ++ii) {
                    const c64 ai = pk(a[ii].x, a[ii].y), ai_rot = pk(a[ii].y, -a[ii].x);
#pragma unroll
                    for (int jj = ii + 1; jj < M; ++jj) {
                        // a_i conj(a_j) = (fma(ai.x, aj.x, ai.y aj.y), fma(ai.y, aj.x, -ai.x aj.y))
                        const c64 u = fma2(ai, pk(a[jj].x, a[jj].x), mul2(ai_rot, pk(a[jj].y, a[jj].y)));
                        acc[i][q] = add2(acc[i][q], u);
                        ++q;
                    }
                }
            }
        }
        // no second barrier: frame n+1 writes the other px buffer, and frame n+2 reuses this one only
        // after the barrier of frame n+1, which every thread reaches after finishing these reads
    }
    if (SPLIT) return;
    float2* cc_out = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG + grp) * (size_t)F) * p.P;
#pragma unroll
    for (int i = 0; i < C::kNb; ++i) {
        const int f = tid + i * C::kThreads;
        if (f < F) {
#pragma unroll
            for (int q = 0; q < C::kP; ++q) cc_out[(size_t)q * F + f] = upk(acc[i][q]);   // [pair][bin]: coalesced
        }
    }
}

// ---- round-robin variant of the fused kernel (M <= 8) ----------------------------------------------------------------
// ncu's source view of stft_cc_warp_kernel<7, false> (profiles/r02_notes.md section 10): 15 % of all warp samples sit at
// the __syncthreads between the seven FFTs of a frame and its pair products.  Two causes: seven warps per CTA spread
// over four sub-partitions as 2/2/2/1 (two resident CTAs: 4/4/3/3), so the FFTs of the fuller partitions finish last
// every frame; and the barrier makes every warp wait for the slowest FFT of the CURRENT frame.  Here
//   * the CTA has eight warps whatever M is and the (frame, mic) FFTs of its frame group are dealt round-robin,
//     task t = frame * M + mic to warp t mod 8: four warps per sub-partition, 8/M frames per round;
//   * the block barrier is replaced by two pairs of alternating mbarriers.  `full`: a warp arrives (non-blocking) once
//     its spectrum of round r is in the ring and waits for round r - 1 only AFTER the FFT of round r, just before the
//     pair products of the frames that round r - 1 completed.  `read`: a warp arrives when its pair products of round
//     r are done, and waits for round r - 1 just before it writes the spectrum of round r into the ring.  Either way a
//     warp blocks only if some other warp is a whole FFT behind;
//   * the spectra go through a ring of K frames, K the smallest depth for which a frame's slot is rewritten two rounds
//     after the round that completed it (M = 7: 4), i.e. after the round in which it was read.
// The FFTs and the order of the frame sums are those of stft_cc_warp_kernel: same bits.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rr_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void rr_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rr_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

constexpr int kRrWarps = 8;
__host__ __device__ constexpr int rr_ring(int M) {   // smallest K with floor(M (f + K) / 8) >= floor((M f + M - 1) / 8) + 2 for all f
    return M == 2 ? 8 : M == 3 ? 6 : M == 8 ? 2 : 4;
}

template <int M>
__global__ void stft_cc_rr_kernel(StftCcParams p) {
    constexpr int kP = M * (M - 1) / 2;
    constexpr int K = rr_ring(M);
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t s_full[2], s_read[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    float* tile = smem + warp * (2 * 32 * 33);
    c64* s_px = reinterpret_cast<c64*>(smem + kRrWarps * (2 * 32 * 33));   // [K][M][F]
    float2* s_post = reinterpret_cast<float2*>(s_px + (size_t)K * M * p.F);   // [F] exp(-i pi k / 1024) of the scored bins
    float2* s_seed = s_post + p.F;                                           // [4][32] W_1024^{lane j}, j = 1, 8, 16, 24
    const int grp = blockIdx.x, w = blockIdx.y, b = blockIdx.z;
    const int F = p.F;
    if (tid == 0) {
        rr_mbar_init(&s_full[0], kRrWarps);
        rr_mbar_init(&s_full[1], kRrWarps);
        rr_mbar_init(&s_read[0], kRrWarps);
        rr_mbar_init(&s_read[1], kRrWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // twiddles live in shared memory, not in registers and not behind __ldg: the register file is full (64 data +
    // 2 (M choose 2) sums) and the L1 left next to two CTAs' shared memory is 28 KB
    for (int i = tid; i < F; i += 32 * kRrWarps) s_post[i] = __ldg(p.twpost + i);
    if (tid < 128) {
        const int j = tid >> 5, mult = j == 0 ? 1 : 8 * j;
        s_seed[tid] = __ldg(p.tw1024 + ((mult * (tid & 31)) & 1023));
    }
    __syncthreads();

    c64 acc[kP];
#pragma unroll
    for (int q = 0; q < kP; ++q) acc[q] = pk(0.f, 0.f);

    const int n0 = grp * p.FG;
    const int nfr = min(p.Nf, n0 + p.FG) - n0;
    const int total = nfr * M, rounds = (total + kRrWarps - 1) / kRrWarps;
    const float* xw = p.mix + (size_t)b * p.M * (size_t)p.T + (size_t)w * p.step;
    auto prefetch_task = [&](int t) {
        const int fr = t / M, m = t - fr * M;
        prefetch_frame_n(tile, xw + (size_t)m * p.T, n0 + fr, p.win_len, lane);
    };

    if (warp < total) prefetch_task(warp);
    for (int r = 0; r <= rounds; ++r) {     // round `rounds`: only the pair products of the last round's frames
        const int t = r * kRrWarps + warp;
        if (t < total) {
            c64 v[32];
            cp_async_wait_all();
            __syncwarp();
            {
                const c64* raw = reinterpret_cast<const c64*>(tile);
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = raw[32 * q + lane];
            }
            __syncwarp();
            dft32(v);
            {
                // opaque per round: otherwise the 28 products of the twiddle recurrence (a function of the lane only) are
                // hoisted out of the round loop and kept in LOCAL memory -- 224 bytes per thread that the 28 KB of L1
                // left next to two CTAs' shared memory cannot hold
                float2 a1 = s_seed[lane], a8 = s_seed[32 + lane], a16 = s_seed[64 + lane], a24 = s_seed[96 + lane];
                asm volatile("" : "+f"(a1.x), "+f"(a1.y), "+f"(a8.x), "+f"(a8.y), "+f"(a16.x), "+f"(a16.y), "+f"(a24.x), "+f"(a24.y));
                twiddle_and_transpose(v, tile, lane, a1, a8, a16, a24);
            }
            __syncwarp();
            load_transposed(v, tile, lane);
            __syncwarp();
            if (t + kRrWarps < total) prefetch_task(t + kRrWarps);
            dft32(v);
            // the ring slot of this frame was last read in round <= r - 1 (rr_ring): every warp has left those reads
            if (r >= 1) rr_mbar_wait(&s_read[(r - 1) & 1], (uint32_t)(((r - 1) >> 1) & 1));
            const int fr = t / M, m = t - fr * M;
            c64* px = s_px + ((size_t)(fr % K) * M + m) * F;
            const float tol2 = p.tol * p.tol;
            const int partner = (32 - lane) & 31;
            // opaque per round: otherwise the seven (bin index, twiddle address, slot address) triples are hoisted out
            // of the round loop and spilled around every FFT (the register file is full: 64 data + 2 (M choose 2) sums)
            int bin0 = p.bin0;
            asm volatile("" : "+r"(bin0));
#pragma unroll
            for (int k2 = 0; k2 < kK2; ++k2) {
                const int k = lane + 32 * k2;
                const float2 zk = upk(v[bitrev5(k2)]);
                const float2 give = upk(v[bitrev5(31 - k2)]);
                float2 zc = make_float2(__shfl_sync(0xffffffffu, give.x, partner),
                                        __shfl_sync(0xffffffffu, give.y, partner));
                if (lane == 0) zc = upk(v[bitrev5((32 - k2) & 31)]);
                const int f = k - bin0;
                if (f >= 0 && f < F) {
                    const float2 post = s_post[f];
                    const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                    const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
                    const float2 Xf = cadd(e, cmul(post, o));
                    const float n2 = fmaxf(fmaf(Xf.x, Xf.x, Xf.y * Xf.y), tol2);
                    float inv;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(n2));
                    px[f] = pk(Xf.x * inv, Xf.y * inv);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && r < rounds) rr_mbar_arrive(&s_full[r & 1]);   // this warp's spectrum of round r is in the ring
        if (r >= 1) {
            // round r - 1 is complete for every warp (they have had one FFT of slack to get there)
            rr_mbar_wait(&s_full[(r - 1) & 1], (uint32_t)(((r - 1) >> 1) & 1));
            // frames completed by round r - 1 and not by round r - 2 (recomputed, not carried: no register to spare)
            const int cnt = min(nfr, (kRrWarps * r) / M);
            for (int fdone = min(nfr, (kRrWarps * (r - 1)) / M); fdone < cnt; ++fdone) {
                if (tid < F) {
                    const c64* px = s_px + (size_t)(fdone % K) * M * F + tid;
                    float2 a[M];
#pragma unroll
                    for (int mm = 0; mm < M; ++mm) a[mm] = upk(px[mm * F]);
                    int q = 0;
#pragma unroll
                    for (int ii = 0; ii < M; ++ii) {
                        const c64 ai = pk(a[ii].x, a[ii].y), ai_rot = pk(a[ii].y, -a[ii].x);
#pragma unroll
                        for (int jj = ii + 1; jj < M; ++jj) {
                            const c64 u = fma2(ai, pk(a[jj].x, a[jj].x), mul2(ai_rot, pk(a[jj].y, a[jj].y)));
                            acc[q] = add2(acc[q], u);
                            ++q;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && r < rounds) rr_mbar_arrive(&s_read[r & 1]);   // this warp's reads of round r are done
    }
    float2* cc_out = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG + grp) * (size_t)F) * p.P;
    if (tid < F) {
#pragma unroll
        for (int q = 0; q < kP; ++q) cc_out[(size_t)q * F + tid] = upk(acc[q]);
    }
}

// Cross-spectra from the stored spectra (split path).  The M(M-1)/2 pairs are covered by tiles of 4 first mics x 8
// second mics (i in [i0, i0+4), j in [j0, j0+8), j0 = i0, i0+8, ...): a thread owns one bin, keeps the tile's 32
// accumulators in registers and loads 12 spectra per frame.  CTA = (frame group, window, mixture x tile).
// Same products, same fused adds and the same frame order as the register path, so M <= 8 gives the same bits either way.
constexpr int kPairThreads = 256;
constexpr int kTileI = 4, kTileJ = 8, kMaxPairTiles = 24;   // M = 32: 8 i-blocks x <= 4 j-blocks, 20 tiles

struct PairTiles {
    int n;
    unsigned char i0[kMaxPairTiles], j0[kMaxPairTiles];
};

__global__ void __launch_bounds__(kPairThreads) pair_products_kernel(StftCcParams p, PairTiles tiles) {
    const int grp = blockIdx.x, w = blockIdx.y;
    const int t = blockIdx.z % tiles.n, b = blockIdx.z / tiles.n;
    const int i0 = tiles.i0[t], j0 = tiles.j0[t];
    const int f = threadIdx.x, F = p.F, M = p.M;
    if (f >= F) return;
    c64 acc[kTileI][kTileJ];
#pragma unroll
    for (int a = 0; a < kTileI; ++a)
#pragma unroll
        for (int c = 0; c < kTileJ; ++c) acc[a][c] = pk(0.f, 0.f);
    const int n0 = grp * p.FG, n1 = min(p.Nf, n0 + p.FG);
    for (int n = n0; n < n1; ++n) {
        const float2* px = p.px_out + ((((size_t)b * p.Nw + w) * p.Nf + n) * M) * F + f;
        c64 ai[kTileI], ai_rot[kTileI];
        float2 aj[kTileJ];
#pragma unroll
        for (int a = 0; a < kTileI; ++a) {
            const float2 v = (i0 + a < M) ? px[(size_t)(i0 + a) * F] : make_float2(0.f, 0.f);
            ai[a] = pk(v.x, v.y);
            ai_rot[a] = pk(v.y, -v.x);
        }
#pragma unroll
        for (int c = 0; c < kTileJ; ++c) aj[c] = (j0 + c < M) ? px[(size_t)(j0 + c) * F] : make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < kTileI; ++a)
#pragma unroll
            for (int c = 0; c < kTileJ; ++c)
                if (i0 + a < j0 + c && j0 + c < M) {
                    const c64 u = fma2(ai[a], pk(aj[c].x, aj[c].x), mul2(ai_rot[a], pk(aj[c].y, aj[c].y)));
                    acc[a][c] = add2(acc[a][c], u);
                }
    }
    float2* cc_out = p.cc_part + ((((size_t)b * p.Nw + w) * p.NG + grp) * (size_t)F) * p.P;
#pragma unroll
    for (int a = 0; a < kTileI; ++a)
#pragma unroll
        for (int c = 0; c < kTileJ; ++c) {
            const int i = i0 + a, j = j0 + c;
            if (i < j && j < M) cc_out[(size_t)(i * M - i * (i + 1) / 2 + (j - i - 1)) * F + f] = upk(acc[a][c]);
        }
}

}  // namespace
int stft_cc_warp_ctas_per_sm(int M);
namespace {

template <int M>
size_t warp_smem_bytes(int F) {
    return (size_t)M * (2 * 32 * 33) * sizeof(float) + (size_t)2 * M * F * sizeof(float2);
}

// ASW_STFT_CTAS=1: at most one CTA per SM (the request is padded past half of the shared memory), which leaves the
// other half of the register file to a shift-stack CTA of another stream (shift_stack.cu, deep variant).
int stft_cta_cap() {
    static const int v = [] { const char* e = getenv("ASW_STFT_CTAS"); return e ? atoi(e) : 0; }();
    return v;
}

template <int M>
int launch_t(const StftCcParams& p, cudaStream_t s) {
    size_t smem = warp_smem_bytes<M>(p.F);
    const size_t half = 114 * 1024;
    if (stft_cta_cap() == 1 && smem < half) smem = half;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        const size_t most = stft_cta_cap() == 1 && warp_smem_bytes<M>(200) < half ? half : warp_smem_bytes<M>(200);
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_warp_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)most));
        // without this the driver may pick a carve-out that fits only one CTA (ncu: occupancy limit 1)
        // just enough shared memory for the resident CTAs, the rest stays L1
        const int ctas = stft_cta_cap() == 1 ? 1 : stft_cc_warp_ctas_per_sm(M);
        int pct = (int)((ctas * (most + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
        if (pct > 100) pct = 100;
        if (carve_all()) pct = carve_all();
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_warp_kernel<M, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    dim3 grid(p.NG, p.Nw, p.B);
    stft_cc_warp_kernel<M, false><<<grid, 32 * M, smem, s>>>(p);
    ASW_LAUNCH_CHECK("stft_cc_warp_kernel");
    return ASW_OK;
}

template <int M>
size_t rr_smem_bytes(int F) {
    return (size_t)kRrWarps * (2 * 32 * 33) * sizeof(float) + (size_t)rr_ring(M) * M * F * sizeof(float2) +
           (size_t)(F + 128) * sizeof(float2);
}

template <int M>
int launch_rr_t(const StftCcParams& p, cudaStream_t s) {
    const size_t smem = rr_smem_bytes<M>(p.F);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        const size_t most = rr_smem_bytes<M>(200);
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_rr_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
        // two CTAs per SM (16 warps at 128 registers fill the register file): shared memory for both, the rest stays L1
        int pct = (int)((2 * (most + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
        if (pct > 100) pct = 100;
        if (carve_all()) pct = carve_all();
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_rr_kernel<M>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    dim3 grid(p.NG, p.Nw, p.B);
    stft_cc_rr_kernel<M><<<grid, 32 * kRrWarps, smem, s>>>(p);
    ASW_LAUNCH_CHECK("stft_cc_rr_kernel");
    return ASW_OK;
}

}  // namespace

bool stft_cc_warp_supported(const StftCcParams& p) {
    return p.M >= 2 && p.M <= 8 && p.F <= 200 && p.bin0 >= 1 && p.bin0 + p.F <= 32 * kK2;
}

// resident CTAs per SM of the fast kernel (drives the frame-group choice in api.cu)
int stft_cc_warp_ctas_per_sm(int M) {
    // registers are allocated per SM sub-partition (16K each): 128 regs/thread = 4 warps per partition
    const int by_regs = 16 / M;
    const int by_smem = (int)((227 * 1024) / (M * (2 * 32 * 33) * sizeof(float) + 2 * M * 200 * sizeof(float2) + 1024));
    const int n = by_regs < by_smem ? by_regs : by_smem;
    return n < 1 ? 1 : n;
}

bool stft_split_supported(const StftCcParams& p) {
    return p.M >= 2 && p.M <= kMaxMics && p.F <= 200 && p.F <= kPairThreads && p.bin0 >= 1 && p.bin0 + p.F <= 32 * kK2;
}

// FFT + PHAT of 8 mics per CTA to p.px_out, then the pair products: any M <= 32.
int launch_stft_split(const StftCcParams& p, cudaStream_t s) {
    constexpr int W = 8;
    if (!p.px_out) {
        set_error("stft split path: no spectrum workspace");
        return ASW_ERR_ARG;
    }
    const size_t smem = (size_t)W * (2 * 32 * 33) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_warp_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int pct = (int)((2 * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
        ASW_CUDA_CHECK(cudaFuncSetAttribute(stft_cc_warp_kernel<W, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    const int mchunks = (p.M + W - 1) / W;
    PairTiles tiles{};
    for (int i0 = 0; i0 < p.M - 1; i0 += kTileI)
        for (int j0 = i0; j0 < p.M; j0 += kTileJ) {
            tiles.i0[tiles.n] = (unsigned char)i0;
            tiles.j0[tiles.n] = (unsigned char)j0;
            ++tiles.n;
        }
    if ((long long)p.B * mchunks > 65535 || (long long)p.B * tiles.n > 65535) {
        set_error("stft split path: B=%d x M=%d exceeds the grid limit; split the batch", p.B, p.M);
        return ASW_ERR_ARG;
    }
    stft_cc_warp_kernel<W, true><<<dim3(p.NG, p.Nw, p.B * mchunks), 32 * W, smem, s>>>(p);
    ASW_LAUNCH_CHECK("stft_cc_warp_kernel(split)");
    pair_products_kernel<<<dim3(p.NG, p.Nw, p.B * tiles.n), kPairThreads, 0, s>>>(p, tiles);
    ASW_LAUNCH_CHECK("pair_products_kernel");
    return ASW_OK;
}

int launch_stft_cc_warp(const StftCcParams& p, cudaStream_t s) {
    // Round-robin kernel for M <= 7 (C2, 64 mixtures of 7 mics: 455 -> 410 us; M = 2 .. 6: 5 - 15 % faster); M = 8 deals
    // one frame per round either way and its 28 pair sums leave the round-robin bookkeeping no registers (measured
    // 2 % slower), so it keeps the first kernel.  ASW_STFT=classic / rr force one kernel for A/B measurements and
    // the bit-identity test.
    static const int force = [] {
        const char* e = getenv("ASW_STFT");
        return !e ? 0 : strcmp(e, "classic") == 0 ? 1 : strcmp(e, "rr") == 0 ? 2 : 0;
    }();
    const bool rr = force == 2 || (force == 0 && p.M <= 7);
    if (rr && stft_cta_cap() != 1) {
        switch (p.M) {
            case 2: return launch_rr_t<2>(p, s);
            case 3: return launch_rr_t<3>(p, s);
            case 4: return launch_rr_t<4>(p, s);
            case 5: return launch_rr_t<5>(p, s);
            case 6: return launch_rr_t<6>(p, s);
            case 7: return launch_rr_t<7>(p, s);
            case 8: return launch_rr_t<8>(p, s);
            default: break;
        }
    }
    switch (p.M) {
        case 2: return launch_t<2>(p, s);
        case 3: return launch_t<3>(p, s);
        case 4: return launch_t<4>(p, s);
        case 5: return launch_t<5>(p, s);
        case 6: return launch_t<6>(p, s);
        case 7: return launch_t<7>(p, s);
        case 8: return launch_t<8>(p, s);
        default: set_error("stft_cc_warp: unsupported mic count %d", p.M); return ASW_ERR_ARG;
    }
}

}  // namespace asw
