// api.cu -- C ABI of libasw.so (declared in include/asw.h): handle management, host-side table
// construction and the launch sequence of the scoring path.
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace asw {

int carve_all() {
    static const int v = [] { const char* e = getenv("ASW_CARVE"); return e ? atoi(e) : 0; }();
    return v;
}

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace asw

using namespace asw;

struct asw_srp {
    int device = 0, M = 0, P = 0, G = 0, Gpad = 0;
    int nfft = 0, hop = 0, bin0 = 0, bin1 = 0, F = 0, U = 0;
    float tol = 0.f;
    int frame_mode = ASW_FRAMES_FLOOR;
    int stft_path = ASW_STFT_AUTO;
    // per-pair lag-table layout (host copies) and device mirrors
    std::vector<int> lag_lo, n_entries, npad, off;
    int tab_len = 0;
    int *d_lag_lo = nullptr, *d_n_entries = nullptr, *d_npad = nullptr, *d_off = nullptr;
    uint32_t* d_pos = nullptr;   // [P][Gpad] Q12.20, slot order
    int* d_perm = nullptr;       // [Gpad] slot -> hypercube (-1 padding)
    float2* d_tw1024 = nullptr;  // [1024]
    float2* d_twpost = nullptr;  // [F]
    float* d_fir = nullptr;      // [(U-1)][12] upsampling weights of gcc.cu
    // staging groups of the gather kernel, one set per window-chunk size
    // staging plans of the gather kernel, one per (hypercubes per CTA, windows per chunk) combination seen so far
    struct GatherPlan {
        int tile = 0, wc = 0, stage_floats = 0;
        int *d_rng_lo = nullptr, *d_rng_n = nullptr, *d_tile_grp = nullptr, *d_grp_flat = nullptr;
    };
    std::vector<GatherPlan> plans;
    std::vector<uint32_t> pos_host;   // [P][Gpad] copy of d_pos (plans are built from it)
    // workspace (grown on demand)
    float2* d_cc_part = nullptr;
    size_t cc_part_cap = 0;
    float2* d_px = nullptr;       // [B][Nw][Nf][M][F] spectra of the split STFT path
    size_t px_cap = 0;
    float2* d_cc = nullptr;
    size_t cc_cap = 0;
    float* d_gcc = nullptr;
    size_t gcc_cap = 0;
    // shape of the last score call (for the stage taps)
    int last_B = 0, last_Nw = 0;
    bool gcc_internal = false;   // the last call left its GCC tables in d_gcc (asw_srp_score), not in a caller buffer
};

namespace {

// Order of the hypercubes inside the gather kernel: a k-d split of the lag vectors (always along the pair whose lags
// spread most, cut at a multiple of 32) so that the 32 hypercubes of a warp are neighbours in every pair's lag table.
// Their 4-tap gathers then fall into a window of a few dozen table entries -- distinct banks or a broadcast --
// instead of 32 unrelated addresses (the cluster order of the reference walks 1.6 m of grid per warp).
void kd_order(const double* lag, int P, int* idx, int n) {
    if (n <= 32) return;
    int best_p = 0;
    double best_spread = -1.0;
    for (int p = 0; p < P; ++p) {
        double mn = lag[(size_t)idx[0] * P + p], mx = mn;
        for (int i = 1; i < n; ++i) {
            const double v = lag[(size_t)idx[i] * P + p];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mx - mn > best_spread) {
            best_spread = mx - mn;
            best_p = p;
        }
    }
    int mid = ((n / 2 + 31) / 32) * 32;
    if (mid >= n) mid = n - 32 > 0 ? ((n - 1) / 32) * 32 : n / 2;
    std::nth_element(idx, idx + mid, idx + n, [&](int a, int b) {
        const double va = lag[(size_t)a * P + best_p], vb = lag[(size_t)b * P + best_p];
        return va < vb || (va == vb && a < b);
    });
    kd_order(lag, P, idx, mid);
    kd_order(lag, P, idx + mid, n - mid);
}

// Staging plan for CTAs of `tile` hypercubes (slot order) and `wc` windows per chunk: for every tile and pair the
// range of table entries the tile's 4-tap gathers touch (a k-d tile covers about a quarter of a pair's lags), and the
// partition of the pairs into groups that fit one shared-memory stage.
int build_plan(asw_srp* h, int tile, int wc, const asw_srp::GatherPlan** out) {
    for (const auto& pl : h->plans)
        if (pl.tile == tile && pl.wc == wc) {
            *out = &pl;
            return ASW_OK;
        }
    const int budget = srp_gather_smem_budget();
    const int P = h->P, G = h->G, ntiles = (G + tile - 1) / tile;
    std::vector<int> lo((size_t)ntiles * P), nn((size_t)ntiles * P), tile_grp(ntiles + 1), flat;
    int max_bytes = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int s0 = t * tile, s1 = std::min(G, s0 + tile);
        tile_grp[t] = (int)flat.size();
        flat.push_back(0);
        int cur = 0;
        for (int p = 0; p < P; ++p) {
            const uint32_t* row = h->pos_host.data() + (size_t)p * h->Gpad;
            int mn = INT_MAX, mx = INT_MIN;
            for (int sl = s0; sl < s1; ++sl) {
                const int i0 = (int)(row[sl] >> kFracBits);
                mn = std::min(mn, i0);
                mx = std::max(mx, i0);
            }
            const int l = (mn - 1) & ~3;                         // taps i0 - 1 .. i0 + 2
            const int n = ((mx + 2 - l + 1) + 3) & ~3;
            lo[(size_t)t * P + p] = l;
            nn[(size_t)t * P + p] = n;
            const int bytes = wc * n * (int)sizeof(float);
            if (l < 0 || l + n > h->npad[p] || bytes > budget) {
                set_error("gather plan: pair %d needs %d entries x %d windows per stage (limit %d bytes); lower the "
                          "oversampling", p, n, wc, budget);
                return ASW_ERR_RANGE;
            }
            if (cur + bytes > budget) {
                flat.push_back(p);
                cur = 0;
            }
            cur += bytes;
            max_bytes = std::max(max_bytes, cur);
        }
        flat.push_back(P);
    }
    tile_grp[ntiles] = (int)flat.size();
    asw_srp::GatherPlan pl;
    pl.tile = tile;
    pl.wc = wc;
    pl.stage_floats = max_bytes / (int)sizeof(float);
    auto up = [](int** d, const std::vector<int>& v) {
        cudaError_t e = cudaMalloc(d, sizeof(int) * v.size());
        if (e == cudaSuccess) e = cudaMemcpy(*d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice);
        return e;
    };
    cudaError_t e = up(&pl.d_rng_lo, lo);
    if (e == cudaSuccess) e = up(&pl.d_rng_n, nn);
    if (e == cudaSuccess) e = up(&pl.d_tile_grp, tile_grp);
    if (e == cudaSuccess) e = up(&pl.d_grp_flat, flat);
    if (e != cudaSuccess) {
        set_error("gather plan: %s", cudaGetErrorString(e));
        cudaFree(pl.d_rng_lo); cudaFree(pl.d_rng_n); cudaFree(pl.d_tile_grp); cudaFree(pl.d_grp_flat);
        return ASW_ERR_CUDA;
    }
    h->plans.reserve(16);                                        // pointers into the vector stay valid
    if (h->plans.size() >= 16) {                                 // a caller cycling through many batch shapes: start over
        for (auto& q : h->plans) { cudaFree(q.d_rng_lo); cudaFree(q.d_rng_n); cudaFree(q.d_tile_grp); cudaFree(q.d_grp_flat); }
        h->plans.clear();
    }
    h->plans.push_back(pl);
    *out = &h->plans.back();
    return ASW_OK;
}

template <typename T>
int ensure(T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return ASW_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need * sizeof(T));
    if (e != cudaSuccess) {
        set_error("device allocation of %zu bytes failed: %s", need * sizeof(T), cudaGetErrorString(e));
        return ASW_ERR_ALLOC;
    }
    *cap = need;
    return ASW_OK;
}

}  // namespace

extern "C" {

int asw_version(void) { return 100; }
const char* asw_last_error(void) { return g_err; }
long long asw_launch_count(void) { return g_launches.load(); }

int asw_srp_num_windows(int T, int win_len) {
    if (win_len < 2) return 0;
    const int step = win_len / 2;
    int n = 0;
    for (int j = 0; j < T / step - 1; ++j) {
        if (j * step + win_len > T) break;
        ++n;
    }
    return n;
}

int asw_srp_num_frames(int win_len, int nfft, int hop) {
    if (win_len < nfft || hop <= 0) return 0;
    return (win_len - nfft) / hop + 1;
}

int asw_srp_num_frames_mode(int win_len, int nfft, int hop, int frame_mode) {
    if (frame_mode == ASW_FRAMES_FLOOR) return asw_srp_num_frames(win_len, nfft, hop);
    if (win_len < 1 || hop <= 0) return 0;
    if (win_len < nfft) return 1;                                  // one frame, zero padded
    return (win_len - nfft + hop - 1) / hop + 1;
}

int asw_srp_set_stft_path(asw_srp_t* h, int path) {
    if (!h || path < ASW_STFT_AUTO || path > ASW_STFT_GENERIC) {
        set_error("asw_srp_set_stft_path: null handle or unknown path %d", path);
        return ASW_ERR_ARG;
    }
    h->stft_path = path;
    return ASW_OK;
}

int asw_srp_set_frame_mode(asw_srp_t* h, int frame_mode) {
    if (!h || (frame_mode != ASW_FRAMES_FLOOR && frame_mode != ASW_FRAMES_PAD_TAIL)) {
        set_error("asw_srp_set_frame_mode: null handle or unknown mode %d", frame_mode);
        return ASW_ERR_ARG;
    }
    h->frame_mode = frame_mode;
    return ASW_OK;
}

int asw_srp_create(asw_srp_t** out, int device, int M, int G, const double* lag, int nfft, int hop, int bin0, int bin1,
                   float tol, int oversample) {
    if (!out || !lag) {
        set_error("asw_srp_create: null argument");
        return ASW_ERR_ARG;
    }
    *out = nullptr;
    if (M < 2 || M > kMaxMics || G < 1) {
        set_error("asw_srp_create: M=%d (2..%d) / G=%d unsupported", M, kMaxMics, G);
        return ASW_ERR_ARG;
    }
    if (nfft != kNfft || hop != kHop) {
        set_error("asw_srp_create: only nfft=%d, hop=%d is implemented (got %d, %d)", kNfft, kHop, nfft, hop);
        return ASW_ERR_ARG;
    }
    if (bin0 < 0 || bin1 <= bin0 || bin1 > kNfft / 2 || bin1 - bin0 > 256) {
        set_error("asw_srp_create: bin range [%d, %d) unsupported", bin0, bin1);
        return ASW_ERR_ARG;
    }
    const int U = oversample == 0 ? 4 : oversample;
    if (U != 1 && U != 2 && U != 4 && U != 8) {
        set_error("asw_srp_create: oversample must be 1, 2, 4 or 8");
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(device);      // the caller's current device is restored on return
    if (!guard.ok) {
        set_error("cannot make CUDA device %d current (no CUDA device, or a bad index)", device);
        return ASW_ERR_CUDA;
    }

    asw_srp* h = new asw_srp();
    h->device = device;
    h->M = M;
    h->P = M * (M - 1) / 2;
    h->G = G;
    h->Gpad = ((G + 2047) / 2048) * 2048;
    h->nfft = nfft;
    h->hop = hop;
    h->bin0 = bin0;
    h->bin1 = bin1;
    h->F = bin1 - bin0;
    h->U = U;
    h->tol = tol;
    const int P = h->P;

    // per-pair lag range -> table layout
    h->lag_lo.resize(P);
    h->n_entries.resize(P);
    h->npad.resize(P);
    h->off.resize(P);
    int off = 0;
    for (int p = 0; p < P; ++p) {
        double mn = lag[p], mx = lag[p];
        for (int g = 1; g < G; ++g) {
            const double v = lag[(size_t)g * P + p];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (!(mn == mn) || !(mx == mx) || fabs(mn) > 1e6 || fabs(mx) > 1e6) {
            set_error("asw_srp_create: non-finite or absurd lag for pair %d", p);
            delete h;
            return ASW_ERR_ARG;
        }
        const int lo = (int)floor(mn) - 2, hi = (int)ceil(mx) + 2;
        const int n = (hi - lo) * U + 1;
        if (n > kMaxEntries - 4) {
            set_error("asw_srp_create: pair %d spans %d lag-table entries (limit %d); lower the oversampling", p, n,
                      kMaxEntries - 4);
            delete h;
            return ASW_ERR_RANGE;
        }
        h->lag_lo[p] = lo;
        h->n_entries[p] = n;
        h->npad[p] = (n + 3) & ~3;
        h->off[p] = off;
        off += h->npad[p];
    }
    h->tab_len = off;

    // slot -> hypercube permutation of the gather kernel (see kd_order); the map is still written in hypercube order
    std::vector<int> perm(h->Gpad, -1);
    for (int g = 0; g < G; ++g) perm[g] = g;
    kd_order(lag, P, perm.data(), G);

    // fixed-point positions, transposed to [P][Gpad] in slot order; padding slots point at a valid interior entry
    std::vector<uint32_t> pos((size_t)P * h->Gpad, 2u << kFracBits);
    const double one = (double)(1u << kFracBits);
    for (int slot = 0; slot < G; ++slot) {
        const int g = perm[slot];
        for (int p = 0; p < P; ++p) {
            const double x = (lag[(size_t)g * P + p] - (double)h->lag_lo[p]) * (double)U;
            double q = floor(x * one + 0.5);
            const double qmax = (double)(h->n_entries[p] - 3) * one;  // keep i0 + 2 inside the table
            if (q < one) q = one;
            if (q > qmax) q = qmax;
            pos[(size_t)p * h->Gpad + slot] = (uint32_t)q;
        }
    }

    std::vector<float2> tw(kNc), twp(h->F);
    for (int t = 0; t < kNc; ++t) {
        const double a = -2.0 * M_PI * (double)t / (double)kNc;
        tw[t] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int f = 0; f < h->F; ++f) {
        const double a = -2.0 * M_PI * (double)(bin0 + f) / (double)kNfft;
        twp[f] = make_float2((float)cos(a), (float)sin(a));
    }

    // 12-tap Lagrange upsampling weights for fractions fr/U, nodes -5..+6 (gcc.cu stage B)
    std::vector<float> fir((size_t)(U > 1 ? U - 1 : 1) * 12, 0.f);
    for (int fr = 1; fr < U; ++fr) {
        const double x = (double)fr / (double)U;
        for (int a = 0; a < 12; ++a) {
            double num = 1.0, den = 1.0;
            for (int c = 0; c < 12; ++c) {
                if (c == a) continue;
                num *= x - (double)(c - 5);
                den *= (double)(a - c);
            }
            fir[(size_t)(fr - 1) * 12 + a] = (float)(num / den);
        }
    }

    int rc = ASW_OK;
    do {
#define TRY(expr)                                                                          \
    {                                                                                      \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                     \
            rc = ASW_ERR_CUDA;                                                             \
            break;                                                                         \
        }                                                                                  \
    }
        TRY(cudaMalloc(&h->d_lag_lo, sizeof(int) * P));
        TRY(cudaMalloc(&h->d_n_entries, sizeof(int) * P));
        TRY(cudaMalloc(&h->d_npad, sizeof(int) * P));
        TRY(cudaMalloc(&h->d_off, sizeof(int) * P));
        TRY(cudaMalloc(&h->d_pos, sizeof(uint32_t) * pos.size()));
        TRY(cudaMalloc(&h->d_perm, sizeof(int) * perm.size()));
        TRY(cudaMemcpy(h->d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice));
        TRY(cudaMalloc(&h->d_tw1024, sizeof(float2) * kNc));
        TRY(cudaMalloc(&h->d_twpost, sizeof(float2) * h->F));
        TRY(cudaMalloc(&h->d_fir, sizeof(float) * fir.size()));
        TRY(cudaMemcpy(h->d_fir, fir.data(), sizeof(float) * fir.size(), cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_lag_lo, h->lag_lo.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_n_entries, h->n_entries.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_npad, h->npad.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_off, h->off.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_pos, pos.data(), sizeof(uint32_t) * pos.size(), cudaMemcpyHostToDevice));
        h->pos_host = pos;
        TRY(cudaMemcpy(h->d_tw1024, tw.data(), sizeof(float2) * kNc, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_twpost, twp.data(), sizeof(float2) * h->F, cudaMemcpyHostToDevice));
#undef TRY
    } while (0);
    if (rc != ASW_OK) {
        asw_srp_destroy(h);
        return rc;
    }
    *out = h;
    return ASW_OK;
}

int asw_srp_destroy(asw_srp_t* h) {
    if (!h) return ASW_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->d_lag_lo);
    cudaFree(h->d_n_entries);
    cudaFree(h->d_npad);
    cudaFree(h->d_off);
    cudaFree(h->d_pos);
    cudaFree(h->d_perm);
    cudaFree(h->d_px);
    cudaFree(h->d_tw1024);
    cudaFree(h->d_twpost);
    cudaFree(h->d_fir);
    for (auto& q : h->plans) { cudaFree(q.d_rng_lo); cudaFree(q.d_rng_n); cudaFree(q.d_tile_grp); cudaFree(q.d_grp_flat); }
    cudaFree(h->d_cc_part);
    cudaFree(h->d_cc);
    cudaFree(h->d_gcc);
    delete h;
    return ASW_OK;
}

// Stage 1 + 2 of the scoring path: STFT + PHAT + cross-spectra, then the pair GCC lag tables into `gcc_out`
// ([B][Nw * tab_len], pair-major).  Nw >= 1 and Nf >= 1 are the caller's responsibility.
static int run_gcc(asw_srp* h, const float* mix_dev, int B, int T, int win_len, int Nw, int Nf, float* gcc_out,
                   cudaStream_t s) {
    StftCcParams sp{};
    sp.mix = mix_dev;
    sp.tw1024 = h->d_tw1024;
    sp.twpost = h->d_twpost;
    sp.B = B;
    sp.M = h->M;
    sp.T = T;
    sp.Nw = Nw;
    sp.step = win_len / 2;
    sp.win_len = win_len;
    sp.Nf = Nf;
    sp.bin0 = h->bin0;
    sp.F = h->F;
    sp.P = h->P;
    sp.tol = h->tol;
    const bool fast = stft_cc_warp_supported(sp);
    // Frame groups: a CTA owns FG consecutive frames of one (mixture, window) and writes one partial
    // cross-spectrum.  FG is a constant, NOT a function of the batch size: the order in which frame
    // products are summed must not depend on how mixtures are batched or sharded across GPUs, so that
    // B mixtures on one GPU and B/n mixtures on each of n GPUs give bit-identical maps.
    // 17 frames per CTA (4 groups per 67-frame window): measured 8 / 12 / 17 / 23 / 34 -> 27.92 / 27.69 / 27.54 /
    // 27.47 / 27.34 ms per 768-mixture step (fewer partial spectra to write and to sum in gcc.cu) against 733 / 778 /
    // 742 / 800 / 808 us for a single mixture through Apply_SRP_PHAT (fewer CTAs); ASW_FG overrides for experiments.
    static const int FG = [] { const char* e = getenv("ASW_FG"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 17; }();
    const int NG = (Nf + FG - 1) / FG;
    sp.NG = NG;
    sp.FG = FG;

    int rc;
    if ((rc = ensure(&h->d_cc_part, &h->cc_part_cap, (size_t)B * Nw * NG * h->F * h->P)) != ASW_OK) return rc;
    if ((rc = ensure(&h->d_cc, &h->cc_cap, (size_t)B * Nw * h->F * h->P)) != ASW_OK) return rc;

    sp.cc_part = h->d_cc_part;
    // fused register kernel for M <= 8; beyond that (or on request) spectra go through global memory to a pair kernel;
    // the radix-4 shared-memory kernel remains as the fallback for bin ranges the warp FFT does not cover
    const bool split = stft_split_supported(sp) &&
                       (h->stft_path == ASW_STFT_SPLIT || (h->stft_path == ASW_STFT_AUTO && !fast));
    if (split) {
        if ((rc = ensure(&h->d_px, &h->px_cap, (size_t)B * Nw * Nf * h->M * h->F)) != ASW_OK) return rc;
        sp.px_out = h->d_px;
        rc = launch_stft_split(sp, s);
    } else if (fast && h->stft_path != ASW_STFT_GENERIC) {
        rc = launch_stft_cc_warp(sp, s);
    } else {
        rc = launch_stft_cc(sp, s);
    }
    if (rc != ASW_OK) return rc;

    GccParams gp{};
    gp.cc_part = h->d_cc_part;
    gp.gcc = gcc_out;
    gp.cc_out = h->d_cc;
    gp.lag_lo = h->d_lag_lo;
    gp.n_entries = h->d_n_entries;
    gp.npad = h->d_npad;
    gp.off = h->d_off;
    gp.fir = h->d_fir;
    gp.tw1024 = h->d_tw1024;
    gp.twpost = h->d_twpost;
    gp.max_int_lags = 0;
    for (int p = 0; p < h->P; ++p) {
        const int ni = (h->n_entries[p] - 1) / h->U + 1 + 12;
        if (ni > gp.max_int_lags) gp.max_int_lags = ni;
    }
    gp.B = B;
    gp.Nw = Nw;
    gp.NG = NG;
    gp.F = h->F;
    gp.P = h->P;
    gp.bin0 = h->bin0;
    gp.U = h->U;
    gp.tab_len = h->tab_len;
    gp.inv_nf = 1.0f / (float)Nf;
    gp.scale = (float)(1.0 / ((double)h->F * (double)h->P));
    return launch_gcc(gp, s);
}

// Stage 3: steered response of every hypercube of the handle from GCC tables `gcc` ([B][Nw * tab_len]).
static int run_gather(asw_srp* h, const float* gcc, int B, int Nw, float* map_dev, cudaStream_t s) {
    const int wc = Nw < srp_gather_windows_per_chunk() ? Nw : srp_gather_windows_per_chunk();
    const int tile = srp_gather_choose_tile(h->G, B, h->P, h->tab_len);
    const asw_srp::GatherPlan* pl = nullptr;
    int rc = build_plan(h, tile, wc, &pl);
    if (rc != ASW_OK) return rc;
    SrpGatherParams rp{};
    rp.gcc = gcc;
    rp.pos = h->d_pos;
    rp.perm = h->d_perm;
    rp.npad = h->d_npad;
    rp.off = h->d_off;
    rp.rng_lo = pl->d_rng_lo;
    rp.rng_n = pl->d_rng_n;
    rp.tile_grp = pl->d_tile_grp;
    rp.grp_flat = pl->d_grp_flat;
    rp.map = map_dev;
    rp.B = B;
    rp.G = h->G;
    rp.Gpad = h->Gpad;
    rp.P = h->P;
    rp.Nw = Nw;
    rp.tab_len = h->tab_len;
    rp.tile = tile;
    rp.stage_floats = pl->stage_floats;
    return launch_srp_gather(rp, s);
}

static int check_batch(const char* who, const asw_srp* h, const void* a, const void* b, int B) {
    if (!h || !a || !b || B < 1) {
        set_error("%s: null handle/buffer or B < 1", who);
        return ASW_ERR_ARG;
    }
    if (B > 65535) {
        set_error("%s: B=%d exceeds the grid limit 65535; split the batch", who, B);
        return ASW_ERR_ARG;
    }
    return ASW_OK;
}

int asw_srp_score(asw_srp_t* h, const float* mix_dev, int B, int T, int win_len, float* map_dev, void* stream) {
    int rc = check_batch("asw_srp_score", h, mix_dev, map_dev, B);
    if (rc != ASW_OK) return rc;
    DeviceGuard guard(h->device);
    if (!guard.ok) {
        set_error("asw_srp_score: cannot make device %d current", h->device);
        return ASW_ERR_CUDA;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int Nw = asw_srp_num_windows(T, win_len);
    const int Nf = asw_srp_num_frames_mode(win_len, h->nfft, h->hop, h->frame_mode);
    if (Nw < 1 || Nf < 1) {
        // no analysis window fits: the reference leaves the map at its zero initialisation (:253)
        ASW_CUDA_CHECK(cudaMemsetAsync(map_dev, 0, sizeof(float) * (size_t)B * h->G, s));
        h->last_B = B;
        h->last_Nw = 0;
        return ASW_OK;
    }
    if ((rc = ensure(&h->d_gcc, &h->gcc_cap, (size_t)B * Nw * h->tab_len)) != ASW_OK) return rc;
    if ((rc = run_gcc(h, mix_dev, B, T, win_len, Nw, Nf, h->d_gcc, s)) != ASW_OK) return rc;
    if ((rc = run_gather(h, h->d_gcc, B, Nw, map_dev, s)) != ASW_OK) return rc;
    h->last_B = B;
    h->last_Nw = Nw;
    h->gcc_internal = true;
    return ASW_OK;
}

int asw_srp_gcc(asw_srp_t* h, const float* mix_dev, int B, int T, int win_len, float* gcc_dev, void* stream) {
    int rc = check_batch("asw_srp_gcc", h, mix_dev, gcc_dev, B);
    if (rc != ASW_OK) return rc;
    DeviceGuard guard(h->device);
    const int Nw = asw_srp_num_windows(T, win_len);
    const int Nf = asw_srp_num_frames_mode(win_len, h->nfft, h->hop, h->frame_mode);
    if (Nw < 1 || Nf < 1) {
        set_error("asw_srp_gcc: no analysis window of %d samples fits T=%d", win_len, T);
        return ASW_ERR_ARG;
    }
    rc = run_gcc(h, mix_dev, B, T, win_len, Nw, Nf, gcc_dev, (cudaStream_t)stream);
    if (rc == ASW_OK) {          // asw_srp_read_cc taps this call's cross-spectra; the tables went to the caller
        h->last_B = B;
        h->last_Nw = Nw;
        h->gcc_internal = false;
    }
    return rc;
}

int asw_srp_gather(asw_srp_t* h, const float* gcc_dev, int B, int Nw, float* map_dev, void* stream) {
    int rc = check_batch("asw_srp_gather", h, gcc_dev, map_dev, B);
    if (rc != ASW_OK) return rc;
    if (Nw < 1) {
        set_error("asw_srp_gather: Nw < 1");
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(h->device);
    return run_gather(h, gcc_dev, B, Nw, map_dev, (cudaStream_t)stream);
}

int asw_srp_read_cc(asw_srp_t* h, float* cc_dev, void* stream) {
    if (!h || !cc_dev) {
        set_error("asw_srp_read_cc: null argument");
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(h->device);
    const size_t n = (size_t)h->last_B * h->last_Nw * h->F * h->P;
    if (n == 0) return ASW_OK;
    ASW_CUDA_CHECK(cudaMemcpyAsync(cc_dev, h->d_cc, n * sizeof(float2), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ASW_OK;
}

int asw_srp_gcc_layout(asw_srp_t* h, int* lag_lo, int* n_entries, int* offset, int* table_len, int* oversample) {
    if (!h) {
        set_error("asw_srp_gcc_layout: null handle");
        return ASW_ERR_ARG;
    }
    for (int p = 0; p < h->P; ++p) {
        if (lag_lo) lag_lo[p] = h->lag_lo[p];
        if (n_entries) n_entries[p] = h->n_entries[p];
        if (offset) offset[p] = h->off[p];
    }
    if (table_len) *table_len = h->tab_len;
    if (oversample) *oversample = h->U;
    return ASW_OK;
}

int asw_srp_read_gcc(asw_srp_t* h, float* gcc_dev, void* stream) {
    if (!h || !gcc_dev) {
        set_error("asw_srp_read_gcc: null argument");
        return ASW_ERR_ARG;
    }
    if (!h->gcc_internal && h->last_Nw > 0) {
        set_error("asw_srp_read_gcc: the last call (asw_srp_gcc) wrote its tables to the caller's buffer");
        return ASW_ERR_ARG;
    }
    DeviceGuard guard(h->device);
    const size_t n = (size_t)h->last_B * h->last_Nw * h->tab_len;
    if (n == 0) return ASW_OK;
    ASW_CUDA_CHECK(cudaMemcpyAsync(gcc_dev, h->d_gcc, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ASW_OK;
}

int asw_map_topk(const float* map_dev, int B, int G, int K, int idx_offset, float* val_dev, int32_t* idx_dev,
                 void* stream) {
    if (!map_dev || !val_dev || !idx_dev || B < 1 || G < 0) {
        set_error("asw_map_topk: null buffer or bad shape");
        return ASW_ERR_ARG;
    }
    return launch_topk(map_dev, B, G, K, idx_offset, val_dev, idx_dev, (cudaStream_t)stream);
}

int asw_shift_stack(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev, int N, int B, int M,
                    int T, float* out_dev, void* stream) {
    if (!mix_dev || !shifts_dev || !out_dev || N < 0 || B < 1 || M < 1 || M > kMaxMics || T < 1) {
        set_error("asw_shift_stack: null buffer or bad shape (N=%d B=%d M=%d T=%d)", N, B, M, T);
        return ASW_ERR_ARG;
    }
    return launch_shift_stack(mix_dev, shifts_dev, mix_index_dev, N, B, M, T, out_dev, (cudaStream_t)stream);
}

int asw_pcm16_to_f32(const int16_t* pcm_dev, float* out_dev, long long n, void* stream) {
    if (!pcm_dev || !out_dev || n < 0) {
        set_error("asw_pcm16_to_f32: null buffer or negative length");
        return ASW_ERR_ARG;
    }
    return launch_pcm16_to_f32(reinterpret_cast<const short*>(pcm_dev), out_dev, (size_t)n, (cudaStream_t)stream);
}

int asw_patch_powers(float* x_dev, int N, int T, int window, int demean, float* mean_dev, float* power_dev,
                     float* maxavg_dev, int32_t* argmax_dev, void* stream) {
    if (N < 0 || T < 1 || window < 1 || (N > 0 && (!x_dev || !mean_dev || !power_dev || !maxavg_dev))) {
        set_error("asw_patch_powers: null buffer or bad shape (N=%d T=%d window=%d)", N, T, window);
        return ASW_ERR_ARG;
    }
    return launch_patch_powers(x_dev, N, T, window, demean, mean_dev, power_dev, maxavg_dev, argmax_dev,
                               (cudaStream_t)stream);
}

int asw_shift_stack_counted(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                            const int32_t* n_valid_dev, int n_base, int N, int B, int M, int T, float* out_dev,
                            void* stream) {
    if (!mix_dev || !shifts_dev || !mix_index_dev || !n_valid_dev || !out_dev || N < 0 || n_base < 0 || B < 1 || M < 1 ||
        M > kMaxMics || T < 1) {
        set_error("asw_shift_stack_counted: null buffer or bad shape (N=%d B=%d M=%d T=%d)", N, B, M, T);
        return ASW_ERR_ARG;
    }
    return launch_shift_stack_counted(mix_dev, shifts_dev, mix_index_dev, n_valid_dev, n_base, N, B, M, T, out_dev,
                                      (cudaStream_t)stream);
}

int asw_shift_stack_norm(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev, int N, int B,
                         int M, int T, float* out_dev, float* means_dev, float* stds_dev, double* work_dev,
                         void* stream) {
    if (!mix_dev || !shifts_dev || !out_dev || !means_dev || !stds_dev || !work_dev || N < 0 || B < 1 || M < 1 ||
        M > kMaxMics || T < 2) {
        set_error("asw_shift_stack_norm: null buffer or bad shape (N=%d B=%d M=%d T=%d)", N, B, M, T);
        return ASW_ERR_ARG;
    }
    return launch_shift_stack_norm(mix_dev, shifts_dev, mix_index_dev, N, B, M, T, out_dev, means_dev, stds_dev,
                                   work_dev, nullptr, 0, 0, nullptr, 0, (cudaStream_t)stream);
}

int asw_shift_stack_norm_tab(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev, int N, int B,
                             int M, int T, const double* tables_dev, int table_len, int max_lag, float* out_dev,
                             float* means_dev, float* stds_dev, double* work_dev, const int32_t* n_valid_dev, int n_base,
                             void* stream) {
    if (!mix_dev || !shifts_dev || !out_dev || !means_dev || !stds_dev || !work_dev || N < 0 || B < 1 || M < 1 ||
        M > kMaxMics || T < 2 || n_base < 0) {
        set_error("asw_shift_stack_norm_tab: null buffer or bad shape (N=%d B=%d M=%d T=%d)", N, B, M, T);
        return ASW_ERR_ARG;
    }
    if (tables_dev && (M < 2 || max_lag < 1 || table_len != 2 * M + (M * (M - 1) / 2) * (2 * max_lag + 1))) {
        set_error("asw_shift_stack_norm_tab: table_len=%d does not belong to M=%d, max_lag=%d", table_len, M, max_lag);
        return ASW_ERR_ARG;
    }
    return launch_shift_stack_norm(mix_dev, shifts_dev, mix_index_dev, N, B, M, T, out_dev, means_dev, stds_dev,
                                   work_dev, tables_dev, table_len, max_lag, n_valid_dev, n_base, (cudaStream_t)stream);
}

int asw_shift_stack_norm_grouped(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev, int N, int B,
                                 int M, int T, float* out_dev, float* means_dev, float* stds_dev, double* work_dev,
                                 int32_t* ranges_dev, const int32_t* n_valid_dev, int n_base, void* stream) {
    if (!mix_dev || !shifts_dev || !mix_index_dev || !out_dev || !means_dev || !stds_dev || !work_dev || !ranges_dev ||
        N < 0 || B < 1 || M < 1 || M > kMaxMics || T < 2 || n_base < 0) {
        set_error("asw_shift_stack_norm_grouped: null buffer or bad shape (N=%d B=%d M=%d T=%d)", N, B, M, T);
        return ASW_ERR_ARG;
    }
    return launch_shift_stack_norm_grouped(mix_dev, shifts_dev, mix_index_dev, N, B, M, T, out_dev, means_dev, stds_dev,
                                           work_dev, ranges_dev, n_valid_dev, n_base, (cudaStream_t)stream);
}

}  // extern "C"
