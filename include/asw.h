/*
 * asw.h -- C ABI of libasw.so: B200-native (sm_100a) SRP-PHAT scoring of TDoA
 * hypercubes + per-hypercube circular shift-and-stack.
 *
 * The reference (uw-x/AcousticSwarms-Speech) is 100 % Python and has no FFI of
 * its own; these entry points are what a binding for its hot path would call.
 * Each one cites the reference interface it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes stubs a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success or a negative asw_status code; the
 *     message of the last failure on the calling thread is asw_last_error().
 *   - all *_dev pointers are device pointers on the handle's device, contiguous,
 *     owned by the caller (e.g. torch tensors' data_ptr()).  The handle owns only
 *     its tables and workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are stream-ordered and do not synchronise unless stated.
 *   - no CPU fallback, no dispatch: without a CUDA device every call fails with
 *     ASW_ERR_CUDA.
 *   - a handle is not thread-safe (the reference is single-threaded per
 *     Mic_Array); use one handle per device / per host thread.
 *   - a handle is SINGLE-STREAM and not re-entrant: every handle (asw_srp_t,
 *     asw_peaks_t, asw_select_t, asw_corr_t) owns one workspace that each call
 *     reuses whatever stream it is given.  Two calls on the same handle must
 *     be ordered by the caller (same stream, or an event between the streams);
 *     calls on different handles are independent.
 *   - shift tables: a row whose mixture index is outside [0, B) is skipped by
 *     every shift-stack entry point (never read out of bounds).
 */
#ifndef ASW_H_
#define ASW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum asw_status {
    ASW_OK = 0,
    ASW_ERR_ARG = -1,      /* bad argument (null pointer, unsupported size)   */
    ASW_ERR_CUDA = -2,     /* CUDA runtime error (message has the details)    */
    ASW_ERR_ALLOC = -3,    /* device allocation failed                         */
    ASW_ERR_RANGE = -4     /* geometry exceeds a kernel limit (see message)    */
} asw_status;

typedef struct asw_srp asw_srp_t;

/* Library / build info. asw_version() = major*10000 + minor*100 + patch. */
int asw_version(void);
const char* asw_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process
 * (bench.py reports it as gpu_launches). */
long long asw_launch_count(void);

/* ---------------------------------------------------------------------------
 * SRP-PHAT scoring handle.
 *
 * Replaces the per-geometry state of SRP_PHAT.__init__ that the scoring loop
 * uses: the (G, F, P) float64 steering table `mode_mat_flat_real/imag`
 * (sep/Traditional_SP/SRP_Prunning.py:221-243, generate_mod_vector :368-381).
 * Instead of the table the handle keeps, per hypercube g and mic pair p
 * (i<j, row-major), the fractional pair lag in samples
 *     lag[g*P + p] = fs * (dist(g, mic_i) - dist(g, mic_j)) / C
 * with dist as in :375 (mic height ignored).  The host computes it in fp64
 * from the reference's own `grids` / `mic_pos`.
 *
 *   M          microphones (2..32); P = M(M-1)/2
 *   G          hypercubes (Grid_clusters)
 *   nfft, hop  STFT frame / hop; only nfft = 2048 is implemented (constants.py:27),
 *              hop = nfft/4 (SRP_Prunning.py:406)
 *   bin0,bin1  scored bins [bin0, bin1) (constants.py:24-26: 2, 200)
 *   tol        PHAT magnitude floor (SRP_Prunning.py:384: 1e-8)
 *   oversample lag-table oversampling U in {1,2,4,8}; 0 selects the default (4)
 */
int asw_srp_create(asw_srp_t** out, int device, int M, int G, const double* lag_samples,
                   int nfft, int hop, int bin0, int bin1, float tol, int oversample);
int asw_srp_destroy(asw_srp_t* h);

/* How an analysis window is cut into STFT frames.  The reference calls
 * pyroomacoustics.transform.stft.analysis(x, nfft, nfft // 4) (SRP_Prunning.py:406; pyroomacoustics==0.5.0,
 * requirements.txt:10), which is not vendored with it (assumption A1 of DESIGN.md):
 *   ASW_FRAMES_FLOOR     (default) floor((win - nfft) / hop) + 1 frames, samples that do not fill a frame are dropped;
 *   ASW_FRAMES_PAD_TAIL  ceil((win - nfft) / hop) + 1 frames, the last one completed with zeros -- the
 *                        "append zeros" handling of a ragged tail by a non-streaming STFT.
 * Both are implemented, tested against the oracle, and differ only when (win - nfft) % hop != 0 (36000- and
 * 24000-sample windows: 67 vs 68 and 43 vs 44 frames).  A maintainer who can run the pinned package re-pins A1
 * with this one call. */
enum { ASW_FRAMES_FLOOR = 0, ASW_FRAMES_PAD_TAIL = 1 };

/* Which kernels compute the STFT + PHAT + cross-spectra stage (:404-426); results agree to rounding (FUSED and
 * SPLIT bit for bit).  AUTO: FUSED for M <= 8, SPLIT for 9..32 mics, GENERIC where neither applies.
 *   ASW_STFT_SPLIT    warp FFT + PHAT to a [B][Nw][Nf][M][F] spectrum buffer, then a pair-product kernel (any M <= 32)
 *   ASW_STFT_GENERIC  radix-4 shared-memory FFT with the pair products in the same kernel (any M, any bin range) */
enum { ASW_STFT_AUTO = 0, ASW_STFT_SPLIT = 1, ASW_STFT_GENERIC = 2 };
int asw_srp_set_stft_path(asw_srp_t* h, int path);
int asw_srp_set_frame_mode(asw_srp_t* h, int frame_mode);
int asw_srp_num_frames_mode(int win_len, int nfft, int hop, int frame_mode);

/* SRP_PHAT.SRP_Map_WINDOW_new / SRP_Map_WINDOW_torch
 * (sep/Traditional_SP/SRP_Prunning.py:384-433) for a batch of B mixtures that
 * share the geometry: analysis-window framing (:393-403), rectangular-window
 * STFT (:404-409), per-channel PHAT (:414-416), pair cross-spectra averaged over
 * frames (:418-426), steered response at every hypercube (:428-429), max over
 * windows starting from zero (:253, :430).
 *   mix_dev  [B][M][T] float32           map_dev  [B][G] float32 (output)
 *   win_len  analysis-window length (sep/Mic_Array.py:160-163: 36000 | 24000)
 * Windows: step = win_len/2, j < T/step - 1, j*step + win_len <= T. */
int asw_srp_score(asw_srp_t* h, const float* mix_dev, int B, int T, int win_len,
                  float* map_dev, void* stream);

/* The two halves of asw_srp_score, for sharding the hypercubes of one geometry over several GPUs without recomputing
 * the transform stage on every rank: each rank runs asw_srp_gcc on ITS mixtures, the ranks all-gather the GCC lag
 * tables (NCCL), and each rank runs asw_srp_gather on ITS hypercubes for all mixtures.  All ranks' handles must be
 * built with the same per-pair lag range (dist.py appends the two rows of global per-pair minima / maxima to every
 * rank's lag slice), so that the table layout -- asw_srp_gcc_layout -- is the same everywhere.
 *   gcc_dev [B][Nw * table_len] float32, pair-major (pair p at Nw * off[p], then [Nw][npad[p]]) */
int asw_srp_gcc(asw_srp_t* h, const float* mix_dev, int B, int T, int win_len, float* gcc_dev, void* stream);
int asw_srp_gather(asw_srp_t* h, const float* gcc_dev, int B, int Nw, float* map_dev, void* stream);

/* Number of analysis windows / STFT frames the call above uses (host-side helper;
 * mirrors SRP_Prunning.py:393-403 and the pyroomacoustics frame rule). */
int asw_srp_num_windows(int T, int win_len);
int asw_srp_num_frames(int win_len, int nfft, int hop);

/* Stage taps for parity tests: results of the most recent asw_srp_score call.
 *   cc_dev   [B][Nw][F][P][2] float32 -- CC_flat (:426), real/imag interleaved
 *   gcc_dev  [B][P-segments...]       -- band-limited GCC lag tables (layout via
 *            asw_srp_gcc_layout), already scaled by 1/(F*P) */
int asw_srp_read_cc(asw_srp_t* h, float* cc_dev, void* stream);
int asw_srp_gcc_layout(asw_srp_t* h, int* lag_lo /*P*/, int* n_entries /*P*/, int* offset /*P*/,
                       int* table_len, int* oversample);
int asw_srp_read_gcc(asw_srp_t* h, float* gcc_dev, void* stream);

/* MAX_POWER (:432) and the K best hypercubes per mixture, descending by value,
 * ties broken by lower index.  Used for the multi-GPU merge and as the
 * pruning pre-filter; K in {1..1024}.
 *   val_dev [B][K] float32, idx_dev [B][K] int32 (index into [0,G) + idx_offset);
 *   entries beyond G are (-inf, -1). */
int asw_map_topk(const float* map_dev, int B, int G, int K, int idx_offset,
                 float* val_dev, int32_t* idx_dev, void* stream);

/* ---------------------------------------------------------------------------
 * PCM16 ingest: out[i] = pcm[i] / 32768 (float32, exact) -- the conversion soundfile/librosa apply when the
 * reference loads its PCM_16 wav files (sep/helpers/utils.py read_audio_file, sep/eval/get_items.py:10-44).
 * Lets a caller ship 16-bit audio over PCIe and expand it on the device; the float32 mixture it produces
 * is what every other entry point consumes.  Buffers must be 16-byte aligned. */
int asw_pcm16_to_f32(const int16_t* pcm_dev, float* out_dev, long long n, void* stream);

/* ---------------------------------------------------------------------------
 * Per-patch statistics of the separator's outputs, the numpy loops of binary_search_baseline
 * (sep/helpers/local_utils_3d.py:342-349) and Spotform_Small_Patch_Parallel (sep/Mic_Array.py:288-296):
 *     mean[n]   = mean(x[n])                                      (float32)
 *     x[n]     -= mean[n]            in place when demean != 0
 *     power[n]  = sum((x[n] - mean[n])^2)
 *     maxavg[n] = max_avg_power(x[n] - mean[n], window)[0]        (local_utils_3d.py:13-17: the largest RMS over
 *                 any `window`-sample box, zeros to the right of the signal), argmax[n] = start of that box
 * x_dev [N][T] float32; sums are carried in double (numpy / scipy results agree to float32 rounding).
 * argmax_dev may be NULL. */
int asw_patch_powers(float* x_dev, int N, int T, int window, int demean, float* mean_dev, float* power_dev,
                     float* maxavg_dev, int32_t* argmax_dev, void* stream);

/* ---------------------------------------------------------------------------
 * Shift-and-stack: the loop of DataParallelSpotModel.shift_and_sep
 * (sep/training/JointModel/network.py:75-83) with roll_by_gather (:12-25):
 *     out[n][c][t] = mix[mix_index[n]][c][(t + shifts[n][c]) mod T]
 * where shifts[n][0] = 0 and shifts[n][c] = round_half_even(float32(
 * patch.sample_offset[c-1])) (:81-82); any int32 value is accepted (reduced
 * mod T with Python semantics).
 *   mix_dev    [B][M][T] float32
 *   shifts_dev [N][M] int32
 *   mix_index_dev [N] int32 or NULL (all patches read mixture 0); a row whose index is outside [0, B) is
 *                 skipped (its output slot is left untouched), never read out of bounds
 *   out_dev    [N][M][T] float32 -- the network's input layout (batch, mic, time) */
int asw_shift_stack(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                    int N, int B, int M, int T, float* out_dev, void* stream);

/* Same, fused with normalize_input
 * (sep/training/SpeakerLocalization/network.py:28-40): 16-bit re-quantisation,
 * mean / unbiased std of the mic-average, (x - mean) / std.  The statistics are reduced in a fixed order
 * (results are reproducible bit for bit); the division is carried out as a multiplication by 1 / std
 * (within 1.5 ulp of torch's division, bar: 1e-4 relative).
 *   means_dev, stds_dev [N] float32 (outputs, needed by unnormalize_input :42-47)
 *   work_dev  [N][2] float64 scratch owned by the caller */
int asw_shift_stack_norm(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                         int N, int B, int M, int T, float* out_dev, float* means_dev, float* stds_dev,
                         double* work_dev, void* stream);

/* ---------------------------------------------------------------------------
 * Per-mixture correlation tables for the fused normalize_input
 * (sep/training/SpeakerLocalization/network.py:28-40 applied to the shifted stack of
 * sep/training/JointModel/network.py:75-85).  The mean and the unbiased std of a patch's mic-average
 *     ref[t] = 1/M sum_c q_c[(t + r_c) mod T],   q = round(x 2^15) / 2^15
 * are functions of per-mixture sums:  sum ref = 1/M sum_c S_c  and
 *     sum ref^2 = 1/M^2 (sum_c E_c + 2 sum_{c<c'} R_cc'(r_c' - r_c)),  R_cc'(l) = sum_t q_c[t] q_c'[(t+l) mod T].
 * asw_corr_tables computes, per mixture, S_c [M], E_c [M] and R_cc'(l) [P][2 max_lag + 1] (pairs i<j row-major,
 * l = -max_lag..max_lag) as float64 (S, E exact; R by overlap-save FFT correlation in fp32, ~3e-7 sqrt(E_c E_c')).
 * asw_shift_stack_norm_tab then needs M + M + P look-ups per patch instead of a pass over its M T samples; a patch
 * whose pair lag exceeds max_lag, or whose variance is small against the table's round-off, silently takes the
 * exact pass of asw_shift_stack_norm, so the result does not depend on max_lag.  Means / stds agree with the exact
 * pass to ~1e-6 relative (bar: 1e-4); results are reproducible bit for bit.
 *   asw_corr_create: max_lag in 1..512 samples (48 kHz: 3.6 m of aperture); M in 2..32
 *   asw_corr_tables: mix_dev [B][M][T] float32, T >= 4096; tables_dev [B][asw_corr_table_len] float64 out.
 * The handle owns the spectrum workspace (14 MB per mixture at 7 mics / 3 s, at most 512 MB: larger batches are
 * processed in chunks); it is single-stream and not re-entrant. */
typedef struct asw_corr asw_corr_t;
int asw_corr_create(asw_corr_t** out, int device, int M, int max_lag);
int asw_corr_destroy(asw_corr_t* h);
int asw_corr_table_len(const asw_corr_t* h);
int asw_corr_tables(asw_corr_t* h, const float* mix_dev, int B, int T, double* tables_dev, void* stream);
/* tables_dev may be NULL (exact pass for every patch).  With n_valid_dev != NULL the call covers rows
 * [n_base, n_base + N) of a device-built table (asw_build_shift_table / asw_build_fine_table) and skips the rows at
 * or beyond *n_valid_dev, like asw_shift_stack_counted; out / means / stds / work are indexed from 0. */
int asw_shift_stack_norm_tab(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                             int N, int B, int M, int T, const double* tables_dev, int table_len, int max_lag,
                             float* out_dev, float* means_dev, float* stds_dev, double* work_dev,
                             const int32_t* n_valid_dev, int n_base, void* stream);

/* The same fused shift-stack + normalize_input (sep/training/JointModel/network.py:75-85 +
 * sep/training/SpeakerLocalization/network.py:28-40) for a table whose rows are GROUPED BY MIXTURE (mix_index_dev
 * non-decreasing over the valid rows: what asw_build_shift_table and asw_build_fine_table produce).  One pass over each
 * mixture's audio serves all of its patches from shared-memory tiles, so a few dozen patches per mixture (the coarse
 * stage) cost neither a per-patch pass nor the per-mixture correlation tables.  Both moments are exact integer sums of
 * k = rint(x 2^15): the statistics are those of the unrounded mic average (1e-8 relative from asw_shift_stack_norm,
 * which rounds the average to float32 first) and do not depend on the order of accumulation.
 *   ranges_dev [B + 1] int32 scratch owned by the caller; the other arguments as asw_shift_stack_norm_tab.
 * Samples must satisfy |x| < 4 (32-bit partial sums; audio read from PCM files is in [-1, 1)).
 * M > 8 takes the per-patch pass of asw_shift_stack_norm. */
int asw_shift_stack_norm_grouped(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                                 int N, int B, int M, int T, float* out_dev, float* means_dev, float* stds_dev,
                                 double* work_dev, int32_t* ranges_dev, const int32_t* n_valid_dev, int n_base,
                                 void* stream);

/* ---------------------------------------------------------------------------
 * Peak picking on the device: fill_powermap_torch + MAX_POWER + find_valid_peak_new
 * (sep/Traditional_SP/SRP_Prunning.py:347-357, :432, :500-544) for a batch of maps.
 *   power_index [Lx][Ly][Lz] int32  cluster id of every 5 cm voxel, -1 where the voxel belongs to no
 *                                   cluster (keep-out box); the reference's POWER_INDEX with its
 *                                   zero-initialised non-members marked
 *   dis_matrix  [Lx][Ly] float64    SRP_Prunning.py:140-144
 *   threshold3  {ratio, floor, ceiling} (sep/Mic_Array.py:120), ratio2 = 4 (:500)
 * asw_peaks_find:
 *   map_dev       [B][G] float32
 *   peaks_dev     [B][max_peaks] int32 out: cluster ids in the reference's order (first occurrence in
 *                 C order over the interior voxels), padded with -1
 *   count_dev     [B] int32 out: number of peak clusters (may exceed max_peaks / 2048: list truncated)
 *   max_power_dev [B] float32 out: MAX_POWER
 * Comparisons are evaluated in double on the float32 map values, exactly as the host code does. */
typedef struct asw_peaks asw_peaks_t;
int asw_peaks_create(asw_peaks_t** out, int device, int Lx, int Ly, int Lz, int G, const int32_t* power_index,
                     const double* dis_matrix, const double* threshold3, double ratio2);
int asw_peaks_destroy(asw_peaks_t* h);
int asw_peaks_find(asw_peaks_t* h, const float* map_dev, int B, int32_t* peaks_dev, int max_peaks,
                   int32_t* count_dev, float* max_power_dev, void* stream);

/* ---------------------------------------------------------------------------
 * Greedy selection of the coarse hypercube patches on the device: SRP_PHAT.local_source_adaptive
 * (sep/Traditional_SP/SRP_Prunning.py:547-643, helpers :19-61) for a batch of mixtures, fed by
 * asw_peaks_find.  Decisions (trimming, covered peaks, "is any 1 cm voxel inside") are evaluated in
 * double on the same float64 volumes the reference builds in SRP_PHAT.__init__ (:148-170).
 *   cluster_offsets [G][D] int32    quantised TDoA vector of every cluster (Grid_cluster.sample_offset)
 *   off5_sorted     [n5][D] float64 Offset_5 voxel records ordered by bucket
 *                                   (floor(o0) - b0min) * NB1 + (floor(o1) - b1min)   (o1 bucket 0 if D == 1)
 *   off1_at5        [n5][D] float64 Offset_1 at the 1 cm voxel (5*iy, 5*ix, iz) coinciding with each of
 *                                   those 5 cm voxels, same order (NaN where it does not exist)
 *   vox5            [n5] int32      iy * Nx5 + ix of the ordered voxels
 *   bucket_start    [NB0*NB1 + 1] int32 first record of every bucket
 *   xx5, yy5        float64         the 5 cm grid coordinates (np.arange, :149-150)
 *   axis_range4     {x0, x1, y0, y1} (Axis_range, :146)
 *   off1            [Ny1][Nx1][Nz][D] float64  Offset_1 (:167-170)
 * asw_select_patches outputs, per mixture b (patch q < min(out_count[b], max_patches)):
 *   out_offsets [B][max_patches][D] int32  Patch.sample_offset
 *   out_width   [B][max_patches]    int32  Patch.width_list (identical in every dimension)
 *   out_peak    [B][max_patches]    int32  cluster id whose centre is Patch.peak_pos
 * Patch.area_points is not produced (the host builds it on demand for the patches that get subdivided). */
typedef struct asw_select asw_select_t;
int asw_select_create(asw_select_t** out, int device, int G, int D, int W, const int32_t* cluster_offsets,
                      const double* off5_sorted, const double* off1_at5, const int32_t* vox5,
                      const int32_t* bucket_start, int b0min, int b1min, int NB0, int NB1, int n5, int Nx5, int Ny5,
                      const double* xx5, const double* yy5, const double* axis_range4, const double* off1, int Ny1,
                      int Nx1, int Nz);
int asw_select_destroy(asw_select_t* h);
int asw_select_patches(asw_select_t* h, const float* map_dev, const int32_t* peaks_dev, int max_peaks,
                       const int32_t* count_dev, int B, int32_t* out_count_dev, int32_t* out_offsets_dev,
                       int32_t* out_width_dev, int32_t* out_peak_dev, int max_patches, void* stream);

/* Subdivision of kept coarse hypercubes into fine ones: search_area / binary_area_divide_width
 * (sep/helpers/local_utils_3d.py:212-335) with Patch.check_out (sep/Traditional_SP/Patch_3D.py:69-87), one
 * candidate per CTA.  Needs D <= 8 (M <= 9); larger arrays use the host code.
 *   centres_dev [n][D] int32, widths_dev [n] int32   the coarse patches (width identical in every dimension)
 *   upper_bound [D] float64 (host)                    upper_bound_pairwise (sep/Mic_Array.py:113-115)
 * Outputs per candidate c, leaf l < min(leaf_count[c], max_leaves), in the reference's order:
 *   leaf_off [n][max_leaves][D], leaf_w [n][max_leaves][D] int32   Patch.sample_offset / width_list
 *   leaf_npts [n][max_leaves] int32                                area_size()
 *   leaf_box [n][max_leaves][2][D] float64   closed TDoA box (lo, hi) selecting the leaf's member points among the
 *                                            candidate's area_points (so the host can rebuild them in order)
 *   leaf_centre [n][max_leaves][3] float64   Patch.center_pos() of the leaf = mean position of its member voxels
 *                                            (NaN for an empty leaf); may be NULL
 *   root_members [n][member_cap] int32       the candidate's own area_points (hyperbola_area_init, SRP_Prunning.py:41-61)
 *                                            as flat indices (iy * Nx1 + ix) * Nz + iz into the 1 cm volume, UNORDERED;
 *                                            sorted ascending they are the reference's order.  root_count [n] is their
 *                                            number (a count above member_cap means the list was truncated).  May be NULL.
 *   root_after [n][2][D] int32   the candidate's offsets / widths after check_out (the reference mutates it in place)
 *   status [n] int32             0 ok; 1/2/3 = member-list / node / leaf capacity exceeded (results incomplete) */
int asw_subdivide(asw_select_t* h, const int32_t* centres_dev, const int32_t* widths_dev, int n,
                  const double* upper_bound, int max_leaves, int32_t* leaf_count_dev, int32_t* leaf_off_dev,
                  int32_t* leaf_w_dev, int32_t* leaf_npts_dev, double* leaf_box_dev, double* leaf_centre_dev,
                  int32_t* root_after_dev, int32_t* status_dev, int32_t* root_members_dev, int member_cap,
                  int32_t* root_count_dev, void* stream);
/* Coordinates of the 1 cm volume (np.arange of SRP_Prunning.py:158-160: xx1 [Nx1], yy1 [Ny1], zz [Nz], host
 * arrays); needed once per handle before asw_subdivide is asked for leaf centres. */
int asw_select_set_grid1(asw_select_t* h, const double* xx1, const double* yy1, const double* zz);

/* Dense shift table of the fine stage, the patch-list assembly of Spotform_Small_Patch_Parallel
 * (sep/Mic_Array.py:244-262) for n candidates at once, from asw_subdivide's outputs: candidate i (skipped when
 * widths_dev[i] <= 0, the empty slots of a padded batch) contributes its min(leaf_count[i], max_leaves) leaves
 * followed by one centre patch at root_after[i][0], its offsets after check_out (:250-256).
 *   owner_dev      [n] int32 or NULL   mixture each candidate belongs to (-> mix_index_dev)
 *   shifts_dev     [capacity][D+1] int32 out (column 0 = 0), mix_index_dev [capacity] int32 out
 *   cand_index_dev [capacity] int32 out or NULL: candidate of every row
 *   cand_start_dev [n+1] int32 out: first row of every candidate (patches_indexes of :247, :262)
 *   n_total_dev    [1] int32 out = min(total rows, capacity) */
int asw_build_fine_table(const int32_t* leaf_count_dev, const int32_t* leaf_off_dev, const int32_t* root_after_dev,
                         const int32_t* widths_dev, const int32_t* owner_dev, int n, int max_leaves, int D,
                         int32_t* shifts_dev, int32_t* mix_index_dev, int32_t* cand_index_dev, int32_t* cand_start_dev,
                         int32_t* n_total_dev, int capacity, void* stream);

/* Dense shift table for asw_shift_stack from the per-mixture patch lists above:
 *   shifts_dev [capacity][D+1] int32 (column 0 = 0), mix_index_dev [capacity] int32,
 *   n_total_dev [1] int32 = min(total patches, capacity).  B <= 1024. */
int asw_build_shift_table(const int32_t* count_dev, const int32_t* offsets_dev, int B, int max_patches, int D,
                          int32_t* shifts_dev, int32_t* mix_index_dev, int32_t* n_total_dev, int capacity,
                          void* stream);

/* asw_shift_stack for patches [n_base, n_base + N) of a device-built table whose length lives on the
 * device: patches at or beyond *n_valid_dev are skipped (no host synchronisation needed between the
 * selection and the stacking).  out_dev receives N slots. */
int asw_shift_stack_counted(const float* mix_dev, const int32_t* shifts_dev, const int32_t* mix_index_dev,
                            const int32_t* n_valid_dev, int n_base, int N, int B, int M, int T, float* out_dev,
                            void* stream);

/* ---------------------------------------------------------------------------
 * Host-side hypercube table build: SRP_PHAT.Map_3D_TDoA + search_cluster
 * (sep/Traditional_SP/SRP_Prunning.py:277-344).  Pure CPU code (no device needed).
 *   offsets [Lx][Ly][Lz][D] int64  quantised TDoA vector of every voxel (:327-331)
 *   valid   [Lx][Ly][Lz]    uint8  check_valid (:266-275)
 *   label   [Lx][Ly][Lz]    int32  out: cluster id of the voxel, -1 if invalid
 *   order   [n_valid]       int32  out: flat voxel indices, cluster-major, members in the
 *                                  reference's breadth-first order
 *   cluster_start [n_valid + 1] int32 out: start of each cluster inside `order`
 * Clusters are numbered in first-seen C-order, exactly like the reference. */
int asw_geometry_cluster(const int64_t* offsets, const uint8_t* valid, int Lx, int Ly, int Lz, int D,
                         int32_t* label, int32_t* order, int32_t* cluster_start, int32_t* n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* ASW_H_ */
