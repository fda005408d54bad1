// srp_gather.cu -- steered response power of every hypercube by gathering the pair GCC tables at
// the hypercube's fractional pair lags, summed over pairs, max over analysis windows.
//
// Reference arithmetic (sep/Traditional_SP/SRP_Prunning.py):
//   :428-429  map_w[g] = sum_{f,p} Re(CC[f,p] tab[g,f,p]) / F / P    == sum_p R_p(tau[g,p])  (see gcc.cu)
//   :253,:430 SRP_map = maximum(SRP_map, map_w), SRP_map starts at zeros  (so the map is >= 0)
//
// Staging is pair-major and double buffered: a group of pairs x up to kWc windows of GCC table is copied to one of
// two shared-memory stages with cp.async while the threads gather from the other one.  Each thread takes its
// hypercubes' fixed-point lag positions for the staged pairs (one coalesced load, reused by every window and computed
// once into 4-tap Lagrange weights), and gathers 4 adjacent table entries per (hypercube, pair, window).  The
// per-window sums stay in registers; the max over windows is taken at the end.  Hypercubes are visited in the k-d
// order built by asw_srp_create (a warp's 32 hypercubes are neighbours in every pair's table), results are written
// back through the slot -> hypercube permutation.  The kernel is shared-memory-bandwidth bound (16 B per gather).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace asw {
namespace {

constexpr int kMaxThreads = 1024;        // 32 warps hide the shared-memory gather latency (ncu: short_scoreboard)
constexpr int kWc = 8;                    // windows per staging chunk
constexpr int kSmemBudget = 100 * 1024;   // bytes of GCC table per stage; two stages: one is gathered from while the
                                          // next is being filled by cp.async

template <int GPT, int kThreads>
__global__ void __launch_bounds__(kThreads, kMaxThreads / kThreads) srp_gather_kernel(SrpGatherParams p) {
    extern __shared__ __align__(16) float s_tab[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int b = blockIdx.y, t = blockIdx.x;
    const int g_base = t * p.tile;            // p.tile <= kThreads * GPT hypercubes per CTA
    bool on[GPT];
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
        const int sl = gi * kThreads + tid;
        on[gi] = sl < p.tile && g_base + sl < p.G;
    }
    const float* gcc_b = p.gcc + (size_t)b * p.tab_len * p.Nw;
    const int* rlo = p.rng_lo + (size_t)t * p.P;
    const int* rn = p.rng_n + (size_t)t * p.P;
    const int* gb = p.grp_flat + p.tile_grp[t];
    const int n_groups = p.tile_grp[t + 1] - p.tile_grp[t] - 1;

    float best[GPT];
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) best[gi] = 0.f;

    for (int w0 = 0; w0 < p.Nw; w0 += kWc) {
        const int wc = min(kWc, p.Nw - w0);
        float acc[kWc][GPT];
#pragma unroll
        for (int w = 0; w < kWc; ++w)
#pragma unroll
            for (int gi = 0; gi < GPT; ++gi) acc[w][gi] = 0.f;

        // asynchronous copy of one group of pair tables into a stage buffer: for every pair the rows of this chunk's
        // windows, restricted to the entries [lo, lo + n) this tile touches; one (pair, window) row per warp
        auto stage = [&](int grp, float* buf) {
            const int q0 = gb[grp], q1 = gb[grp + 1];
            int so = 0, row = 0;
            for (int pp = q0; pp < q1; ++pp) {
                const int npd = p.npad[pp], n = rn[pp], lo = rlo[pp];
                const float* src = gcc_b + (size_t)p.Nw * p.off[pp] + (size_t)w0 * npd + lo;
                for (int w = 0; w < wc; ++w, ++row) {
                    if (row % kWarps != warp) continue;
                    const float* sr = src + (size_t)w * npd;
                    float* ds = buf + so + w * n;
                    for (int k = 4 * lane; k < n; k += 128)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                                         (unsigned)__cvta_generic_to_shared(ds + k)),
                                     "l"(sr + k));
                }
                so += wc * n;
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        __syncthreads();  // the previous chunk's last stage is no longer read
        stage(0, s_tab);
        for (int grp = 0; grp < n_groups; ++grp) {
            const int p0 = gb[grp], p1 = gb[grp + 1];
            const float* s_cur = s_tab + (grp & 1) * p.stage_floats;
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");   // this thread's part of group `grp` has landed
            __syncthreads();  // everybody's part has, and everybody is done gathering from the other stage
            if (grp + 1 < n_groups) stage(grp + 1, s_tab + ((grp + 1) & 1) * p.stage_floats);
            int sm_off = 0;
            for (int pp = p0; pp < p1; ++pp) {
                const int n = rn[pp], lo = rlo[pp];
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) {
                    if (!on[gi]) continue;
                    const uint32_t q = __ldg(p.pos + (size_t)pp * p.Gpad + g_base + gi * kThreads + tid);
                    const int i0 = (int)(q >> kFracBits);
                    const float f = (float)(q & ((1u << kFracBits) - 1)) * (1.0f / (float)(1 << kFracBits));
                    // 4-tap Lagrange weights for nodes -1, 0, 1, 2
                    const float fm1 = f - 1.f, fm2 = f - 2.f, fp1 = f + 1.f;
                    const float c0 = -(1.f / 6.f) * f * fm1 * fm2;
                    const float c1 = 0.5f * fp1 * fm1 * fm2;
                    const float c2 = -0.5f * fp1 * f * fm2;
                    const float c3 = (1.f / 6.f) * fp1 * f * fm1;
                    const float* tp = s_cur + sm_off + (i0 - lo) - 1;
#pragma unroll
                    for (int w = 0; w < kWc; ++w) {
                        if (w < wc) {
                            float v = c0 * tp[0];
                            v = fmaf(c1, tp[1], v);
                            v = fmaf(c2, tp[2], v);
                            v = fmaf(c3, tp[3], v);
                            acc[w][gi] += v;
                            tp += n;
                        }
                    }
                }
                sm_off += wc * n;
            }
        }
#pragma unroll
        for (int w = 0; w < kWc; ++w)
            if (w < wc) {
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) best[gi] = fmaxf(best[gi], acc[w][gi]);
            }
    }
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
        const int slot = g_base + gi * kThreads + tid;
        if (on[gi]) p.map[(size_t)b * p.G + p.perm[slot]] = best[gi];
    }
}

// ---- round 2: persistent CTAs, bulk-copy (TMA) staging, mbarrier pipeline ----------------------------------------------
// The staged rows are contiguous, 16-byte aligned 1-D ranges -- the one place on this path where "TMA where tiles are
// regular" applies.  One CTA per SM walks a flat sequence of stages (tile, window chunk, pair group) through the same
// two shared-memory buffers:
//   * warp 0 issues one cp.async.bulk per (pair, window) row of the NEXT stage (a lane per row) and arms the stage's
//     `full` mbarrier with the byte count; nobody executes per-lane copy instructions or waits for "its" copies;
//   * every warp waits on `full`, gathers, and arrives on the stage's `empty` mbarrier; only the issuing warp ever
//     waits on `empty`, so there is no block-wide barrier in the loop and fast warps run ahead into the next stage;
//   * because the sequence continues across tiles, the first fill of a tile overlaps the last gathers of the previous
//     one (the first version paid that latency once per CTA, twice per SM at the C2 size).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct StageCursor {      // position in a CTA's flat stage sequence
    int item, w0, grp;
};

template <int GPT, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) srp_gather_bulk_kernel(SrpGatherParams p, int ntiles, int n_items) {
    extern __shared__ __align__(128) float s_tab[];
    __shared__ uint64_t s_full[2], s_empty[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], kWarps);
        mbar_init(&s_empty[1], kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto groups_of = [&](int item) { const int t = item % ntiles; return p.tile_grp[t + 1] - p.tile_grp[t] - 1; };
    auto advance = [&](StageCursor c) {      // next stage of this CTA; item >= n_items: past the end
        if (c.grp + 1 < groups_of(c.item)) {
            ++c.grp;
        } else if (c.w0 + kWc < p.Nw) {
            c.w0 += kWc;
            c.grp = 0;
        } else {
            c.item += gridDim.x;
            c.w0 = 0;
            c.grp = 0;
        }
        return c;
    };
    // warp 0: one bulk copy per (pair, window) row of the stage, a lane per row; lane 0 arms the barrier
    auto issue = [&](StageCursor c, int bufi) {
        const int b = c.item / ntiles, t = c.item % ntiles;
        const float* gcc_b = p.gcc + (size_t)b * p.tab_len * p.Nw;
        const int* rlo = p.rng_lo + (size_t)t * p.P;
        const int* rn = p.rng_n + (size_t)t * p.P;
        const int* gb = p.grp_flat + p.tile_grp[t];
        const int wc = min(kWc, p.Nw - c.w0);
        float* buf = s_tab + bufi * p.stage_floats;
        int so = 0, row = 0;
        for (int pp = gb[c.grp]; pp < gb[c.grp + 1]; ++pp) {
            const int npd = p.npad[pp], n = rn[pp], lo = rlo[pp];
            const float* src = gcc_b + (size_t)p.Nw * p.off[pp] + (size_t)c.w0 * npd + lo;
            for (int w = 0; w < wc; ++w, ++row)
                if ((row & 31) == lane) bulk_load(buf + so + w * n, src + (size_t)w * npd, (uint32_t)n * 4u, &s_full[bufi]);
            so += wc * n;
        }
        if (lane == 0) mbar_expect_tx(&s_full[bufi], (uint32_t)so * 4u);
    };

    StageCursor cur{(int)blockIdx.x, 0, 0};
    if (cur.item >= n_items) return;
    if (warp == 0) issue(cur, 0);

    bool on[GPT];
    float best[GPT];
    float acc[kWc][GPT];
    int s = 0;
    while (cur.item < n_items) {
        const StageCursor nxt = advance(cur);
        const int bufi = s & 1;
        if (warp == 0 && nxt.item < n_items) {
            // the other buffer was last read by stage s - 1: every warp has arrived on its `empty` barrier before the refill
            if (s >= 1) mbar_wait(&s_empty[bufi ^ 1], (uint32_t)(((s - 1) >> 1) & 1));
            issue(nxt, bufi ^ 1);
        }
        const int b = cur.item / ntiles, t = cur.item % ntiles;
        const int g_base = t * p.tile;
        const int* rlo = p.rng_lo + (size_t)t * p.P;
        const int* rn = p.rng_n + (size_t)t * p.P;
        const int* gb = p.grp_flat + p.tile_grp[t];
        const int n_groups = p.tile_grp[t + 1] - p.tile_grp[t] - 1;
        const int wc = min(kWc, p.Nw - cur.w0);
        if (cur.grp == 0) {
            if (cur.w0 == 0) {
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) {
                    const int sl = gi * kThreads + tid;
                    on[gi] = sl < p.tile && g_base + sl < p.G;
                    best[gi] = 0.f;
                }
            }
#pragma unroll
            for (int w = 0; w < kWc; ++w)
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) acc[w][gi] = 0.f;
        }
        const float* s_cur = s_tab + bufi * p.stage_floats;
        // the lag positions of the first pair are fetched before the wait for the stage, those of pair pp + 1 while
        // pair pp is gathered: ncu showed the dependent global load at the head of every pair as the top stall
        // (long_scoreboard 6.1 warps per issue) once the block barriers were gone
        const int pp_end = gb[cur.grp + 1];
        uint32_t qn[GPT];
#pragma unroll
        for (int gi = 0; gi < GPT; ++gi)
            qn[gi] = on[gi] ? __ldg(p.pos + (size_t)gb[cur.grp] * p.Gpad + g_base + gi * kThreads + tid) : 0u;
        mbar_wait(&s_full[bufi], (uint32_t)((s >> 1) & 1));
        int sm_off = 0;
        for (int pp = gb[cur.grp]; pp < pp_end; ++pp) {
            const int n = rn[pp], lo = rlo[pp];
            uint32_t qc[GPT];
#pragma unroll
            for (int gi = 0; gi < GPT; ++gi) {
                qc[gi] = qn[gi];
                if (pp + 1 < pp_end && on[gi]) qn[gi] = __ldg(p.pos + (size_t)(pp + 1) * p.Gpad + g_base + gi * kThreads + tid);
            }
#pragma unroll
            for (int gi = 0; gi < GPT; ++gi) {
                if (!on[gi]) continue;
                const uint32_t q = qc[gi];
                const int i0 = (int)(q >> kFracBits);
                const float f = (float)(q & ((1u << kFracBits) - 1)) * (1.0f / (float)(1 << kFracBits));
                // 4-tap Lagrange weights for nodes -1, 0, 1, 2 (same expressions as srp_gather_kernel: same bits)
                const float fm1 = f - 1.f, fm2 = f - 2.f, fp1 = f + 1.f;
                const float c0 = -(1.f / 6.f) * f * fm1 * fm2;
                const float c1 = 0.5f * fp1 * fm1 * fm2;
                const float c2 = -0.5f * fp1 * f * fm2;
                const float c3 = (1.f / 6.f) * fp1 * f * fm1;
                const float* tp = s_cur + sm_off + (i0 - lo) - 1;
#pragma unroll
                for (int w = 0; w < kWc; ++w) {
                    if (w < wc) {
                        float v = c0 * tp[0];
                        v = fmaf(c1, tp[1], v);
                        v = fmaf(c2, tp[2], v);
                        v = fmaf(c3, tp[3], v);
                        acc[w][gi] += v;
                        tp += n;
                    }
                }
            }
            sm_off += wc * n;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[bufi]);              // this warp is done reading the stage
        if (cur.grp == n_groups - 1) {
#pragma unroll
            for (int w = 0; w < kWc; ++w)
                if (w < wc) {
#pragma unroll
                    for (int gi = 0; gi < GPT; ++gi) best[gi] = fmaxf(best[gi], acc[w][gi]);
                }
            if (cur.w0 + kWc >= p.Nw) {
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) {
                    const int slot = g_base + gi * kThreads + tid;
                    if (on[gi]) p.map[(size_t)b * p.G + p.perm[slot]] = best[gi];
                }
            }
        }
        cur = nxt;
        ++s;
    }
}

// ---- warp-specialised variant (round 2, second pass) -------------------------------------------------------------------
// ncu's source view of srp_gather_bulk_kernel (profiles/r02_notes.md section 10): 81 % of the stage entries found the
// stage's data not yet landed and 21 % of all warp samples sat in the `full` wait, although a stage's copies have a
// whole stage of gathers (~30 us) to complete.  The copies were not slow, their ISSUE was: warp 0 walked the stage's
// pairs serially (four dependent global loads per pair, a predicated single-lane copy per row) and then did a full
// share of the gathers, so it was the last warp of every stage and the other 24 waited for it.  A further 7.5 % waited
// for the per-pair range (rn / rlo) loads at the head of every pair.  Here
//   * an extra PRODUCER warp does nothing but stage: a lane per pair (ranges and offsets loaded in parallel, the stage
//     offsets by a warp scan), one cp.async.bulk per (pair, window) row, and it publishes {n, lo} of every staged pair
//     in shared memory next to the data (released by the same mbarrier arrive);
//   * the kThreads gather threads never touch the plan in global memory inside the pair loop.
// Same arithmetic in the same order as the other two kernels: bit-identical maps.
constexpr int kMaxStagePairs = kMaxMics * (kMaxMics - 1) / 2;

template <int GPT, int kThreads>
__global__ void __launch_bounds__(kThreads + 32, 1) srp_gather_ws_kernel(SrpGatherParams p, int ntiles, int n_items) {
    extern __shared__ __align__(128) float s_tab[];
    __shared__ uint64_t s_full[2], s_empty[2];
    __shared__ int2 s_desc[2][kMaxStagePairs];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;                 // gather warps; warp kWarps is the producer
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], kWarps);
        mbar_init(&s_empty[1], kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto groups_of = [&](int item) { const int t = item % ntiles; return p.tile_grp[t + 1] - p.tile_grp[t] - 1; };
    auto advance = [&](StageCursor c) {
        if (c.grp + 1 < groups_of(c.item)) {
            ++c.grp;
        } else if (c.w0 + kWc < p.Nw) {
            c.w0 += kWc;
            c.grp = 0;
        } else {
            c.item += gridDim.x;
            c.w0 = 0;
            c.grp = 0;
        }
        return c;
    };
    StageCursor cur{(int)blockIdx.x, 0, 0};
    if (cur.item >= n_items) return;

    if (warp == kWarps) {
        // ---------------- producer ----------------
        for (int s = 0; cur.item < n_items; ++s, cur = advance(cur)) {
            const int bufi = s & 1;
            // the buffer was last read by stage s - 2: every gather warp has arrived on its `empty` barrier
            if (s >= 2) mbar_wait(&s_empty[bufi], (uint32_t)(((s - 2) >> 1) & 1));
            const int b = cur.item / ntiles, t = cur.item % ntiles;
            const float* gcc_b = p.gcc + (size_t)b * p.tab_len * p.Nw;
            const int* rlo = p.rng_lo + (size_t)t * p.P;
            const int* rn = p.rng_n + (size_t)t * p.P;
            const int* gb = p.grp_flat + p.tile_grp[t];
            const int wc = min(kWc, p.Nw - cur.w0);
            const int p0 = gb[cur.grp], p1 = gb[cur.grp + 1];
            float* buf = s_tab + bufi * p.stage_floats;
            int carry = 0;
            for (int base = p0; base < p1; base += 32) {
                const int pp = base + lane;
                const bool valid = pp < p1;
                int n = 0, lo = 0, npd = 0, off = 0;
                if (valid) { n = rn[pp]; lo = rlo[pp]; npd = p.npad[pp]; off = p.off[pp]; }
                int incl = wc * n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int so = carry + incl - wc * n;
                if (valid) {
                    s_desc[bufi][pp - p0] = make_int2(n, lo);
                    const float* src = gcc_b + (size_t)p.Nw * off + (size_t)cur.w0 * npd + lo;
                    for (int w = 0; w < wc; ++w)
                        bulk_load(buf + so + w * n, src + (size_t)w * npd, (uint32_t)n * 4u, &s_full[bufi]);
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();   // every lane's descriptors are written before lane 0's releasing arrive
            if (lane == 0) mbar_expect_tx(&s_full[bufi], (uint32_t)carry * 4u);
        }
        return;
    }

    // ---------------- gather warps ----------------
    bool on[GPT];
    float best[GPT];
    float acc[kWc][GPT];
    for (int s = 0; cur.item < n_items; ++s) {
        const StageCursor nxt = advance(cur);
        const int bufi = s & 1;
        const int b = cur.item / ntiles, t = cur.item % ntiles;
        const int g_base = t * p.tile;
        const int* gb = p.grp_flat + p.tile_grp[t];
        const int n_groups = p.tile_grp[t + 1] - p.tile_grp[t] - 1;
        const int wc = min(kWc, p.Nw - cur.w0);
        if (cur.grp == 0) {
            if (cur.w0 == 0) {
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) {
                    const int sl = gi * kThreads + tid;
                    on[gi] = sl < p.tile && g_base + sl < p.G;
                    best[gi] = 0.f;
                }
            }
#pragma unroll
            for (int w = 0; w < kWc; ++w)
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) acc[w][gi] = 0.f;
        }
        const float* s_cur = s_tab + bufi * p.stage_floats;
        const int p0 = gb[cur.grp], pp_end = gb[cur.grp + 1];
        // lag positions of the first pair before the wait for the stage, those of pair pp + 1 while pair pp is gathered
        uint32_t qn[GPT];
#pragma unroll
        for (int gi = 0; gi < GPT; ++gi)
            qn[gi] = on[gi] ? __ldg(p.pos + (size_t)p0 * p.Gpad + g_base + gi * kThreads + tid) : 0u;
        mbar_wait(&s_full[bufi], (uint32_t)((s >> 1) & 1));
        int sm_off = 0;
        for (int pp = p0; pp < pp_end; ++pp) {
            const int2 d = s_desc[bufi][pp - p0];
            const int n = d.x, lo = d.y;
            uint32_t qc[GPT];
#pragma unroll
            for (int gi = 0; gi < GPT; ++gi) {
                qc[gi] = qn[gi];
                if (pp + 1 < pp_end && on[gi]) qn[gi] = __ldg(p.pos + (size_t)(pp + 1) * p.Gpad + g_base + gi * kThreads + tid);
            }
#pragma unroll
            for (int gi = 0; gi < GPT; ++gi) {
                if (!on[gi]) continue;
                const uint32_t q = qc[gi];
                const int i0 = (int)(q >> kFracBits);
                const float f = (float)(q & ((1u << kFracBits) - 1)) * (1.0f / (float)(1 << kFracBits));
                // 4-tap Lagrange weights for nodes -1, 0, 1, 2 (same expressions as srp_gather_kernel: same bits)
                const float fm1 = f - 1.f, fm2 = f - 2.f, fp1 = f + 1.f;
                const float c0 = -(1.f / 6.f) * f * fm1 * fm2;
                const float c1 = 0.5f * fp1 * fm1 * fm2;
                const float c2 = -0.5f * fp1 * f * fm2;
                const float c3 = (1.f / 6.f) * fp1 * f * fm1;
                const float* tp = s_cur + sm_off + (i0 - lo) - 1;
#pragma unroll
                for (int w = 0; w < kWc; ++w) {
                    if (w < wc) {
                        float v = c0 * tp[0];
                        v = fmaf(c1, tp[1], v);
                        v = fmaf(c2, tp[2], v);
                        v = fmaf(c3, tp[3], v);
                        acc[w][gi] += v;
                        tp += n;
                    }
                }
            }
            sm_off += wc * n;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[bufi]);              // this warp is done reading the stage
        if (cur.grp == n_groups - 1) {
#pragma unroll
            for (int w = 0; w < kWc; ++w)
                if (w < wc) {
#pragma unroll
                    for (int gi = 0; gi < GPT; ++gi) best[gi] = fmaxf(best[gi], acc[w][gi]);
                }
            if (cur.w0 + kWc >= p.Nw) {
#pragma unroll
                for (int gi = 0; gi < GPT; ++gi) {
                    const int slot = g_base + gi * kThreads + tid;
                    if (on[gi]) p.map[(size_t)b * p.G + p.perm[slot]] = best[gi];
                }
            }
        }
        cur = nxt;
    }
}

template <int GPT, int kThreads>
int launch_ws_t(const SrpGatherParams& p, cudaStream_t s) {
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(srp_gather_ws_kernel<GPT, kThreads>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSmemBudget));
    }
    const int ntiles = (p.G + p.tile - 1) / p.tile;
    const long long items = (long long)ntiles * p.B;
    const int grid = (int)(items < kNumSms ? items : kNumSms);
    srp_gather_ws_kernel<GPT, kThreads><<<grid, kThreads + 32, 2 * (size_t)p.stage_floats * sizeof(float), s>>>(p, ntiles, (int)items);
    ASW_LAUNCH_CHECK("srp_gather_ws_kernel");
    return ASW_OK;
}

template <int GPT, int kThreads>
int launch_bulk_t(const SrpGatherParams& p, cudaStream_t s) {
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(srp_gather_bulk_kernel<GPT, kThreads>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSmemBudget));
    }
    const int ntiles = (p.G + p.tile - 1) / p.tile;
    const long long items = (long long)ntiles * p.B;
    const int grid = (int)(items < kNumSms ? items : kNumSms);
    srp_gather_bulk_kernel<GPT, kThreads><<<grid, kThreads, 2 * (size_t)p.stage_floats * sizeof(float), s>>>(p, ntiles, (int)items);
    ASW_LAUNCH_CHECK("srp_gather_bulk_kernel");
    return ASW_OK;
}

template <int GPT, int kThreads>
int launch_t(const SrpGatherParams& p, cudaStream_t s) {
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        ASW_CUDA_CHECK(cudaFuncSetAttribute(srp_gather_kernel<GPT, kThreads>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSmemBudget));
    }
    dim3 grid((p.G + p.tile - 1) / p.tile, p.B);
    srp_gather_kernel<GPT, kThreads><<<grid, kThreads, 2 * (size_t)p.stage_floats * sizeof(float), s>>>(p);
    ASW_LAUNCH_CHECK("srp_gather_kernel");
    return ASW_OK;
}

// Hypercubes per CTA.  Every CTA stages the mixture's whole GCC table group by group (copies are asynchronous, but
// the per-group barriers and the shared-memory writes are a fixed cost per CTA, ~s_eq hypercubes' worth of gathers),
// and one CTA fits per SM, so the kernel runs in ceil(CTAs / 148) rounds of (s_eq + tile): pick the tile that
// minimises it.  Calibration (C2, B = 32, G = 21181): tile 1184 -> 576 CTAs, 4 rounds of 45 us; tile 1632 -> 416 CTAs,
// 3 rounds of 50 us  =>  fixed cost ~32 us = ~2850 hypercubes; tiles above 2048 use the 3-per-thread variant.
constexpr int kThreads3 = 800;                  // 3 hypercubes per thread need ~80 registers: 25 warps per CTA
constexpr int kMaxTile = 3 * kThreads3;
constexpr int kWsThreads = kMaxThreads - 32;   // gather threads of the warp-specialised kernel (+ one producer warp)

int choose_tile(int G, int B, int P, int tab_len) {
    const int kMinTile = 512;
    // a k-d tile touches roughly a quarter of every pair's table (C2: 0.20 - 0.27 for tiles of 1184 - 2368)
    const double s_eq = 0.6 * (double)tab_len / (double)(P > 0 ? P : 1);
    int best_tile = 2 * kMaxThreads;
    double best_cost = 1e300;
    for (int nt = (G + kMaxTile - 1) / kMaxTile; nt <= (G + kMinTile - 1) / kMinTile; ++nt) {
        int tile = ((G + nt - 1) / nt + 31) / 32 * 32;
        if (tile > kMaxTile) continue;
        if (tile < kMinTile) tile = kMinTile;
        const long long ctas = (long long)B * ((G + tile - 1) / tile);
        const long long rounds = (ctas + kNumSms - 1) / kNumSms;
        const double cost = (double)rounds * (s_eq + tile);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_tile = tile;
        }
    }
    return best_tile;
}

}  // namespace

int srp_gather_windows_per_chunk() { return kWc; }
int srp_gather_smem_budget() { return kSmemBudget; }

int srp_gather_choose_tile(int G, int B, int P, int tab_len) { return choose_tile(G, B, P, tab_len); }

int launch_srp_gather(const SrpGatherParams& p, cudaStream_t s) {
    if ((size_t)p.stage_floats * sizeof(float) > (size_t)kSmemBudget) {
        set_error("srp_gather: stage of %zu bytes exceeds the shared-memory budget", (size_t)p.stage_floats * sizeof(float));
        return ASW_ERR_RANGE;
    }
    // ASW_GATHER=legacy: the round-1 kernel (per-lane cp.async ring, one CTA per tile), kept for A/B measurements
    static const bool legacy = [] { const char* e = getenv("ASW_GATHER"); return e && strcmp(e, "legacy") == 0; }();
    if (legacy) {
        if (p.tile > 2 * kMaxThreads) return launch_t<3, kThreads3>(p, s);
        if (p.tile > kMaxThreads) return launch_t<2, kMaxThreads>(p, s);
        return launch_t<1, kMaxThreads>(p, s);
    }
    // ASW_GATHER=bulk: the first persistent kernel (warp 0 stages AND gathers), kept for A/B measurements
    static const bool bulk = [] { const char* e = getenv("ASW_GATHER"); return e && strcmp(e, "bulk") == 0; }();
    if (bulk) {
        if (p.tile > 2 * kMaxThreads) return launch_bulk_t<3, kThreads3>(p, s);
        if (p.tile > kMaxThreads) return launch_bulk_t<2, kMaxThreads>(p, s);
        return launch_bulk_t<1, kMaxThreads>(p, s);
    }
    if (p.P > kMaxStagePairs) {
        set_error("srp_gather: %d pairs exceed the stage descriptor table (%d)", p.P, kMaxStagePairs);
        return ASW_ERR_RANGE;
    }
    // kWsThreads gather threads + the producer warp = 1024 threads
    if (p.tile > 2 * kWsThreads) return launch_ws_t<3, kThreads3>(p, s);
    if (p.tile > kWsThreads) return launch_ws_t<2, kWsThreads>(p, s);
    return launch_ws_t<1, kWsThreads>(p, s);
}

}  // namespace asw
