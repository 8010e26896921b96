/*
 * cmc.h - C ABI of libcmc_b200.so: B200 (sm_100a) kernels for the cortico-muscular
 * coherence hot path of paulruesing/multimodal-biosignal-analysis.
 *
 * The reference has no FFI: its boundary is the Python call surface of
 *   src/pipeline/signal_features.py, data_surrogation.py and cbpa.py.
 * Each entry point below cites the reference code whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative CMC_E* code; the message of
 *     the last failure on the calling thread is returned by cmc_last_error();
 *   - no exception crosses the boundary, nothing is allocated on behalf of the
 *     caller: every buffer is caller-owned DEVICE memory (unless marked host), sized
 *     as documented; scratch space is passed in explicitly and sized by the
 *     matching *_workspace_bytes() function;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous with respect to the host and re-entrant per stream;
 *   - complex values are interleaved float pairs (re, im) == numpy complex64.
 */
#ifndef CMC_B200_H
#define CMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CMC_API __attribute__((visibility("default")))
#else
#define CMC_API
#endif

#define CMC_ABI_VERSION 1

#define CMC_OK 0
#define CMC_EINVAL (-1)      /* bad argument (shape, alignment, unsupported size) */
#define CMC_ECUDA (-2)       /* CUDA runtime / driver error */
#define CMC_EWORKSPACE (-3)  /* workspace too small */
#define CMC_EUNSUPPORTED (-4)/* valid request outside the built kernel set */

#define CMC_DETREND_NONE 0       /* multitaper MSC: signal_features.py:743-748 */
#define CMC_DETREND_CONSTANT 1   /* segment mean removed before windowing (scipy Welch) */
#define CMC_DETREND_POST_TAPER 2 /* mean of the tapered segment removed (periodogram, :419) */

#define CMC_SURR_SHIFT 0
#define CMC_SURR_PHASE 1
#define CMC_PHASE_TABLE_BITS 12  /* phase surrogates draw one of 4096 unit-circle phases exp(2 pi i a / 4096) */
#define CMC_FIX_SHIFT 30         /* cluster masses: int64 sums of rint(t * 2^30) */
#define CMC_T_CLAMP 65536.0

CMC_API int cmc_abi_version(void);
CMC_API const char* cmc_last_error(void);
/* number of kernels launched by this library in the calling process (bench bookkeeping) */
CMC_API int64_t cmc_launch_count(void);

/* ------------------------------------------------------------------------------------
 * K1  fused detrend + taper + batched real FFT over segments.
 * Replaces the per-window / per-taper `np.fft.rfft(window * taper)` of
 * signal_features.py:743-748 (multitaper MSC), :412-420 (multitaper PSD) and the
 * Welch segment transforms inside scipy.signal.coherence (preprocessing.py:1228).
 *
 *   x          [n_samples][ld]        float32, time-first, channel c at x[t*ld + c]
 *   seg_starts [n_seg]                int64 (device) first sample of every segment
 *   windows    [n_win][N]             float32 taper rows (DPSS, hann, ...)
 *   spec       [n_seg][n_win][F][spec_ld] complex64, F = bin_hi - bin_lo + 1,
 *              channel c at spec[...][c]; only columns [0, n_ch) are written
 *   N          powers of two in [128, 8192] run the shared-memory FFT kernels; any other length in
 *              [2, 2^20] falls back to a direct O(N F) DFT kernel (validation-size calls)
 * ---------------------------------------------------------------------------------- */
CMC_API int cmc_fft_segments(const float* x, int64_t n_samples, int n_ch, int64_t ld,
                     const int64_t* seg_starts, int n_seg,
                     const float* windows, int n_win, int N, int detrend,
                     int bin_lo, int bin_hi,
                     float* spec, int64_t spec_ld, void* stream);

/* Creates the per-device constant tables cmc_fft_segments uses for segment length N (done on first use otherwise: a
 * cudaMalloc and a synchronous copy, which a stream that is being captured into a CUDA graph cannot take - the first
 * call for a new N inside a capture fails with CMC_EINVAL and says so). */
CMC_API int cmc_fft_prepare(int N);

/* The same for TWO recordings of equal length that share segments, windows and bins (EEG and EMG of one
 * subject-condition): one launch of the pipelined kernel when both arrays qualify for its TMA path (N = 512, 1024,
 * 2048; channel pitches multiples of 4 floats, 16-byte aligned bases), otherwise two cmc_fft_segments calls - the
 * results are identical either way.  spec1 / spec2 use the row pitch spec_ld (e.g. two channel ranges of one array). */
CMC_API int cmc_fft_segments_pair(const float* x1, int n_ch1, int64_t ld1, float* spec1,
                          const float* x2, int n_ch2, int64_t ld2, float* spec2,
                          int64_t n_samples, const int64_t* seg_starts, int n_seg,
                          const float* windows, int n_win, int N, int detrend,
                          int bin_lo, int bin_hi, int64_t spec_ld, void* stream);

/* ------------------------------------------------------------------------------------
 * K1t  band-limited hann-windowed Welch spectra on the tensor cores.
 * The reference's Welch path is scipy.signal.coherence with its defaults (preprocessing.py:1228-1230): periodic
 * hann window, noverlap = nperseg / 2, detrend = 'constant'; its callers keep a narrow band (1 - 100 Hz).  For that
 * case the spectra are computed WITHOUT an FFT: BF16 x 3 tcgen05 GEMMs per half block of N / 2 samples (folded about
 * its centre: sums against a cos table, differences against a sin table, K = N / 4) give the rectangular-window
 * half-block sums P_h[b]; the epilogue applies the hann window as the three-tap filter 1/2 R[b] - 1/4 (R[b-1] +
 * R[b+1]) and adds the two halves of every segment (R_s[b] = P_h[b] + (-1)^b P_{h+1}[b]).  Every sample is
 * transformed once although segments overlap by half, and only the requested bins are computed.  Output identical
 * in meaning to cmc_fft_segments(.., windows = periodic hann, n_win = 1, ..) to ~3e-5 of the spectrum's rms (BF16
 * hi + lo operands, FP32 accumulation).
 *
 * A plan holds what depends on the segment table, N and the band only: the half-block list (segments that overlap
 * their predecessor by exactly N / 2 share a half block; any other segment simply costs two) and the table.
 *   seg_starts_host [n_seg]  int64 on the HOST
 *   N % 128 == 0, 256 <= N <= 16384; at most 102 bins, bin_hi + 2 <= N / 2; otherwise CMC_EUNSUPPORTED
 *     (the caller then uses cmc_fft_segments)
 * cmc_welch_hann_plan_create allocates a few MB of device memory for the plan and synchronises the device once;
 * destroy frees it.  cmc_welch_hann_spectra is asynchronous on `stream` and may be captured into a CUDA graph.
 *   x1 [n_samples][ld1], x2 [n_samples][ld2] (x2 may be NULL)   float32 recordings, 16-byte aligned rows
 *   spec1 / spec2 [n_seg][F][spec_ld] complex64, channel c of recording i at spec_i[..][c]
 * ---------------------------------------------------------------------------------- */
CMC_API int cmc_welch_hann_plan_create(const int64_t* seg_starts_host, int n_seg, int N, int bin_lo, int bin_hi,
                               void** plan_out);
CMC_API int cmc_welch_hann_plan_destroy(void* plan);
CMC_API int cmc_welch_hann_plan_info(const void* plan, int* n_half_blocks, int* n_segments, int* n_bins);
CMC_API int cmc_welch_hann_spectra(const void* plan, const float* x1, int n_ch1, int64_t ld1, float* spec1,
                           const float* x2, int n_ch2, int64_t ld2, float* spec2, int64_t n_samples,
                           int detrend, int64_t spec_ld, void* stream);

/* Power spectra from segment spectra: out[w][f][c] = base_scale * dbl(f) * mean_k |spec[w][k][f][c]|^2, with
 * dbl(f) = 2 for bins strictly inside (0, N/2) when one_sided != 0 (scipy density convention), optionally
 * followed by log10(|.| + 1e-10).  Replaces signal.periodogram + mean over tapers of multitaper_psd
 * (signal_features.py:417-437) and the segment average of signal.welch (signal_features.py:2116) with
 * W = 1, K = segments.   spec [W][K][F][ld_in] complex64, out [W][F][ld_out] float32. */
CMC_API int cmc_psd_from_spectra(const float* spec, int W, int K, int F, int n_ch, int64_t ld_in,
                         float base_scale, int one_sided, int bin_lo, int N, int log_scale,
                         float* out, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------
 * K2w  per-window multitaper magnitude-squared coherence with optional jackknife CI
 * and independence-threshold mask.  Replaces signal_features.py:750-796 and
 * jackknife_coherence_and_ci (:484-578) for all (EEG, EMG) pairs of every window.
 *
 *   X [W][K][F][ldx], Y [W][K][F][ldy]  complex64 spectra from cmc_fft_segments
 *   window_mask [W] uint8 or NULL       0 = skip (outputs left untouched = zeros)
 *   jackknife   0: coh = clip(|Sxy|^2 / (Sxx Syy), 0, 1)
 *               1: coh = leave-one-taper-out mean, ci_lo / ci_hi = Student-t CI in
 *                  Fisher-z space with critical value t_crit = t.ppf(1 - alpha/2, K-1)
 *   it_threshold  >= 0: significant = coh > it_threshold;  < 0: `significant` unused;  NaN (degenerate
 *                 Beta(K-2, K-2) for K <= 2 tapers): mask written as all zeros, like `coh > nan` in numpy
 *   coh, ci_lo, ci_hi [W][F][Ne][Nm] float32;  significant [W][F][Ne][Nm] uint8
 *   Any alignment works; with Nm even, 8-byte aligned float outputs and a 2-byte aligned mask (what every
 *   allocator returns) two adjacent pairs leave as one store.  K <= 15, K * (32 Ne + 12 Nm) <= 200 KB.
 *   The fused variant below returns bit-for-bit the values this one would (same instruction sequence).
 * ---------------------------------------------------------------------------------- */
CMC_API int cmc_msc_windows(const float* X, const float* Y, int W, int K, int F, int Ne, int Nm,
                    int64_t ldx, int64_t ldy, const uint8_t* window_mask,
                    int jackknife, float t_crit, float it_threshold,
                    float* coh, float* ci_lo, float* ci_hi, uint8_t* significant,
                    void* stream);

/* Same arithmetic fused with max_cmc_spectrograms_over_channels (:1132-1171): for every
 * (window, frequency, EEG channel) the EMG channel with the largest value is selected
 * (first index on ties, like np.argmax) and value / CI are gathered at that index, so the
 * (W, F, Ne, Nm) tensor is never materialised.  zero_nonsignificant != 0 reproduces
 * compute_task_wise_aggregated_cmc(enforce_independence_threshold=True) (:979-983).
 *   out_* [W][F][Ne] float32, out_arg [W][F][Ne] int32 (may be NULL) */
CMC_API int cmc_msc_windows_maxemg(const float* X, const float* Y, int W, int K, int F, int Ne, int Nm,
                           int64_t ldx, int64_t ldy, const uint8_t* window_mask,
                           int jackknife, float t_crit, float it_threshold,
                           int zero_nonsignificant,
                           float* out_coh, float* out_lo, float* out_hi, int32_t* out_arg,
                           void* stream);

/* ------------------------------------------------------------------------------------
 * K2  pooled cross-spectral density on the tensor cores (tcgen05, TF32 x3 split) fused
 * with the auto-spectra normalisation into magnitude-squared coherence.  Replaces the
 * averaged CSD / PSD / coherence arithmetic of signal_features.py:750-770 when the
 * average runs over L = segments (Welch, scipy.signal.coherence) or windows x tapers.
 *
 *   X [L][F][ldx], Y [L][F][ldy] complex64 spectra
 *   coh [F][Ne][Nm] float32 = clip(|sum_l conj(X) Y|^2 / (Sxx Syy), 0, 1)
 *   sxx [F][Ne], syy [F][Nm] float32 = sum_l |.|^2           (may be NULL)
 *   sxy [F][Ne][Nm] complex64 un-normalised cross spectrum    (may be NULL)
 *   ws  scratch of cmc_csd_workspace_bytes(); on return it holds the TF32 operand planes and
 *       auto-spectra that cmc_surrogate_null() consumes (and extends with shifted views).
 * ---------------------------------------------------------------------------------- */
CMC_API int64_t cmc_csd_workspace_bytes(int L, int F, int Ne, int Nm);
CMC_API int cmc_csd_msc(const float* X, const float* Y, int L, int F, int Ne, int Nm,
                int64_t ldx, int64_t ldy,
                float* coh, float* sxx, float* syy, float* sxy,
                void* ws, int64_t ws_bytes, void* stream);

/* Coherence only, straight from the spectra (same arguments and results as cmc_csd_msc): the spectra are
 * staged by TMA as MN-major operands (M = (channel, re/im), K = segment), split into TF32 hi/lo and reduced
 * to auto-spectra in shared memory - one read of X and Y, no pack pass, no operand planes.  ws only receives
 * the auto-spectra: cmc_csd_workspace_bytes_min(F, Ne, Nm) bytes suffice when ldx and ldy are even and X, Y
 * are 16-byte aligned; other layouts fall back to cmc_csd_msc and need its full workspace.
 * cmc_csd_operands fills the operand planes (and auto-spectra) of a full-size workspace for
 * cmc_surrogate_null without computing the coherence again. */
CMC_API int64_t cmc_csd_workspace_bytes_min(int F, int Ne, int Nm);
CMC_API int cmc_csd_coherence(const float* X, const float* Y, int L, int F, int Ne, int Nm,
                      int64_t ldx, int64_t ldy,
                      float* coh, float* sxx, float* syy, float* sxy,
                      void* ws, int64_t ws_bytes, void* stream);
CMC_API int cmc_csd_operands(const float* X, const float* Y, int L, int F, int Ne, int Nm,
                     int64_t ldx, int64_t ldy, void* ws, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K3  surrogate null on the cached (whitened) spectra left in `ws` by cmc_csd_msc.
 * New capability: the reference's stand-in is the analytic Beta(K-2,K-2) threshold of
 * signal_features.py:470-481; the construction is defined in oracle/surrogate.py.
 *
 *   mode CMC_SURR_SHIFT: surrogate s rotates the EMG segment index by
 *        shifts[s] * group (shifts int32 [s_end - s_begin], host-chosen, in [1, L/group))
 *   mode CMC_SURR_PHASE: one phase per surrogate, segment and frequency, shared by all EMG
 *        channels: table index = word (f & 3) of Philox4x32-10(key = seed, counter =
 *        (s, l, f >> 2, s >> 32)) >> 20; s runs over GLOBAL indices [s_begin, s_end) so
 *        results do not depend on how surrogates are sharded.  FP16 tensor-core GEMM (FP32
 *        accumulation; operands prescaled into the FP16 range; error-compensated hi/lo
 *        operand split for L <= 85 so that every surrogate coherence stays within 1e-4 of the
 *        float64 definition) with the phase panel resident in shared memory for K <= 512,
 *        streamed with the B tiles for longer segment axes (2 L <= 12288).
 *   Both modes: three-term TF32 (shift) / FP16 (phase) products, FP32 accumulation in TMEM.
 *   coh_obs  [F][Ne][Nm]  observed coherence to compare against
 *   exceed   [F][Ne][Nm]  uint32, += #{s : C_s >= coh_obs}   (caller zero-initialises)
 *   max_stat [s_end - s_begin] float32 max over (f, i, j) of C_s
 *   ws2 scratch of cmc_surrogate_workspace_bytes()
 * ---------------------------------------------------------------------------------- */
/* host copy of the FP16-rounded phase table (fp16(cos), fp16(sin)) of 2 pi a / 4096 the phase GEMM multiplies
 * with, as float pairs [4096][2] (diagnostics of the operand rounding; the definition uses the exact phases) */
CMC_API int cmc_phase_table(float* out_host);
CMC_API int64_t cmc_surrogate_workspace_bytes(int L, int F, int Ne, int Nm, int mode, int64_t n_surr);
CMC_API int cmc_surrogate_null(void* ws, int L, int F, int Ne, int Nm, int mode, int group,
                       const int32_t* shifts, uint64_t seed, int64_t s_begin, int64_t s_end,
                       const float* coh_obs, uint32_t* exceed, float* max_stat,
                       void* ws2, int64_t ws2_bytes, void* stream);
/* Same, restricted to the frequency bins [f_begin, f_end) of the F-bin problem (all pointers still
 * address the full arrays): exceed is only touched inside the range and max_stat is the maximum over
 * the range, so ranks that split the FREQUENCY axis combine with a sum and a max.  Every kernel of the
 * null (operand generation included) then shrinks with the range, which is what scales a single null
 * over GPUs; phases are indexed by the global bin, so the result does not depend on the split. */
CMC_API int cmc_surrogate_null_range(void* ws, int L, int F, int Ne, int Nm, int mode, int group,
                             const int32_t* shifts, uint64_t seed, int64_t s_begin, int64_t s_end,
                             int f_begin, int f_end,
                             const float* coh_obs, uint32_t* exceed, float* max_stat,
                             void* ws2, int64_t ws2_bytes, void* stream);

/* Per-pair null histograms of the phase surrogates [s_begin, s_end) (north star: "null histograms"; they give the
 * per-pair significance thresholds of BASELINE config 3 - the reference's stand-in is apply_threshold_filtering,
 * signal_features.py:581-604).  A surrogate coherence C of pair (i, j) at frequency f lands in bin
 *     floor((C - bin_lo[f][i][j]) * bin_scale[f][i][j])
 * of the pair's histogram; values under the window (negative bin) only raise below[f][i][j], values over it are
 * dropped.  bin_lo = NULL reads as 0 and bin_scale = NULL as n_bins, i.e. n_bins uniform bins over [0, 1]; per-pair
 * (lo, scale) arrays place / zoom the window (a quantile search resolves the upper tail only, which keeps most
 * surrogates out of the histogram - and out of the shared-memory atomics).  Same surrogates as cmc_surrogate_null
 * for the same (seed, s, l, f); counts are ADDED to hist / below, so chunks of the surrogate range accumulate.
 *   hist  [F][Ne][Nm][n_bins] uint32 (caller zero-initialises), 2 <= n_bins <= 128
 *   below [F][Ne][Nm] uint32 or NULL (caller zero-initialises)
 *   ws / ws2 as for cmc_surrogate_null (mode CMC_SURR_PHASE only; CMC_SURR_SHIFT returns CMC_EUNSUPPORTED).
 *   reuse_operands != 0: ws2 still holds the phase panel and cross-product rows that an earlier
 *   cmc_surrogate_null[_range / _hist] call generated for the SAME ws, seed, surrogate range and frequency range
 *   (e.g. the exceedance pass, or the previous zoom pass): their generation is skipped. */
CMC_API int cmc_surrogate_null_hist(void* ws, int L, int F, int Ne, int Nm, int mode, uint64_t seed,
                            int64_t s_begin, int64_t s_end, int f_begin, int f_end, int n_bins,
                            const float* bin_lo, const float* bin_scale, uint32_t* hist, uint32_t* below,
                            void* ws2, int64_t ws2_bytes, int reuse_operands, void* stream);

/* Rank selection in the histograms above: for every row of hist [n_rows][n_bins] the first bin whose running count,
 * started at below[row], exceeds k (the bin of the value of 0-based rank k among the counted and the `below` values);
 * below[row] is replaced by the count below that bin.  bin_out = -1: the rank lies under the window (below > k),
 * bin_out = n_bins: over it (the counts never reach k).  Drives the zoom passes of the per-pair quantiles. */
CMC_API int cmc_hist_select(const uint32_t* hist, int64_t n_rows, int n_bins, int k, int32_t* below, int32_t* bin_out,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * K4  cluster-based permutation test: sign-flip t-map -> threshold -> connected-component
 * labelling over a CSR adjacency -> cluster mass -> signed max statistic.  Replaces the
 * body of mne.stats.permutation_cluster_1samp_test / spatio_temporal_cluster_1samp_test
 * as called at cbpa.py:1027-1042 (algorithm restated in oracle/cbpa.py).
 *
 *   X      [n_subj][n_tests] float64, test index = t * n_ch + ch (C order)
 *   signs  [n_perm][n_subj]  int8 in {-1,+1}, host-chosen
 *   indptr [n_tests + 1], indices [nnz] int32 CSR adjacency (symmetric; diagonal ignored)
 *   tail   0: clusters of t > thr and of t < -thr;  1: t > thr;  -1: t < thr
 *   h0_fixed [p_end - p_begin] int64: signed cluster mass of largest magnitude (0 if no
 *            cluster) as sum of rint(clamp(t) * 2^CMC_FIX_SHIFT) - exact and order-free
 * cmc_cbpa_observed additionally returns the t-map, the label map (0 = no cluster,
 * k = k-th cluster in MNE order: t > thr clusters first, each group by smallest flat
 * index) and per-cluster masses (caller sizes mass_fixed / mass_f64 to n_tests).
 *   ws     scratch of cmc_cbpa_workspace_bytes() for BOTH calls, 16-byte aligned: labelling
 *          scratch plus a re-tiled copy of X (calls sharing a stream may share one workspace).
 *          cmc_cbpa_permute accepts X == NULL when ws still holds the tiled copy written by an
 *          earlier cmc_cbpa_observed / cmc_cbpa_permute call for the same data: the re-tile pass
 *          is then skipped.  Permutations beyond a CTA's first are claimed from a device counter
 *          kept in ws, so ragged permutation costs balance themselves.
 * ---------------------------------------------------------------------------------- */
CMC_API int64_t cmc_cbpa_workspace_bytes(int n_subj, int n_tests);
CMC_API int cmc_cbpa_permute(const double* X, int n_subj, int n_tests,
                     const int8_t* signs, int64_t p_begin, int64_t p_end,
                     double thr, int tail, const int32_t* indptr, const int32_t* indices,
                     int64_t* h0_fixed, void* ws, int64_t ws_bytes, void* stream);
CMC_API int cmc_cbpa_observed(const double* X, int n_subj, int n_tests, double thr, int tail,
                      const int32_t* indptr, const int32_t* indices,
                      double* t_obs, int32_t* labels, int64_t* mass_fixed, double* mass_f64,
                      int32_t* n_clusters, void* ws, int64_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMC_B200_H */
