/*
 * vhr_b200.h -- C ABI of the B200-native rPPG signal path (libvhr_b200.so).
 *
 * The reference (AngaBlue/video-heart-rate) is pure Python and has no FFI; its only
 * "operator API" is Python-level.  This header is the boundary a
 * maintainer binds with ctypes: every entry point below names the
 * reference function(s) whose arithmetic it replaces (file:line in /root/reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - d_* pointers are DEVICE pointers owned by the caller (e.g. torch tensors),
 *     contiguous, row-major; h_* pointers are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     asynchronous with respect to the host and ordered on `stream`, except the
 *     *_host entry points, which synchronise before returning.
 *   - Return value: 0 = VHR_OK, negative = error; vhr_last_error(ctx) gives the text.
 *     Nothing throws across the boundary.
 *   - One context per (device, host thread), driven by ONE stream at a time.  A context owns
 *     cached tables (twiddles, band mask, composite pyrUp weights) and a scratch arena (ROI
 *     partial sums, polygon row masks) that grow on demand outside the hot loop; a call on
 *     another stream than the context's previous call first waits for that call (event), so
 *     changing streams between calls is safe, two streams at once are not supported.
 *   - Frames are uint8, interleaved 3-channel, shape (T,H,W,3).  The library is
 *     channel-order agnostic (RGB per BASELINE.json; the reference's cv2 arrays are BGR:
 *     index 1 is green either way).
 *   - Rectangles are int32 [x1,y1,x2,y2) half-open, already normalised to
 *     0<=x1<=x2<=W, 0<=y1<=y2<=H (the Python host applies NumPy slice semantics,
 *     rppg_VIDEO.py:106).  An empty rectangle yields NaN means, like np.mean([]).
 */
#ifndef VHR_B200_H
#define VHR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VHR_ABI_VERSION 2
#define VHR_MAX_LEVELS 6
#define VHR_MAX_ROIS 8
#define VHR_MAX_POLY_VERTS 64

enum {
    VHR_OK = 0,
    VHR_ERR_INVALID = -1,   /* bad argument (shape, level count, NULL pointer ...) */
    VHR_ERR_CUDA = -2,      /* a CUDA runtime call failed; text in vhr_last_error   */
    VHR_ERR_NOMEM = -3,
    VHR_ERR_UNSUPPORTED = -4
};

typedef struct vhr_ctx vhr_ctx;

/* ---- context ----------------------------------------------------------------------- */
int vhr_abi_version(void);
int vhr_create(vhr_ctx** out, int device);
int vhr_destroy(vhr_ctx* ctx);
const char* vhr_last_error(const vhr_ctx* ctx);   /* ctx may be NULL (global last error) */
/* Number of kernels this context has launched since creation (bench.py gpu_launches). */
int64_t vhr_launch_count(const vhr_ctx* ctx);

/* ---- synthetic clips (measurement / parity infrastructure) -------
 * Bit-identical to oracle/synth.py: a pure integer function of (seed, clip, t, y, x, c).
 * d_pulse_q8: int32 (T,3) host-computed pulse table; face = half-open [x0,x1)x[y0,y1). */
typedef struct {
    uint32_t seed, clip;
    int32_t T, H, W;
    int32_t t0;                 /* first frame index to generate (frames t0 .. t0+T-1)  */
    int32_t face[4];            /* x0,y0,x1,y1 */
    int32_t base_q8[2][3];      /* [background, skin] x channel, Q8 */
    int32_t noise_gain;         /* Q8 per unit of the 4-byte sum */
} vhr_synth_params;
int vhr_synth_clip(vhr_ctx* ctx, const vhr_synth_params* p, const int32_t* d_pulse_q8,
                   uint8_t* d_frames, void* stream);

/* ---- EVM: Gaussian pyramid (no reference code; spec = cv2.pyrDown) --- */
/* dims[l] for l=0..levels: w_{l+1}=(w_l+1)/2.  Host helper, no GPU work. */
int vhr_pyr_dims(int W, int H, int levels, int32_t* w_out, int32_t* h_out);
/* Fused pyrDown cascade: uint8 (T,H,W,3) -> float32 (T,h_L,w_L,3), `levels` reductions in
 * one pass (levels 1..L-1 never touch HBM).  1 <= levels <= VHR_MAX_LEVELS. */
int vhr_pyrdown_cascade(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W,
                        int levels, float* d_level, void* stream);
/* Diagnostics, host only: the tile plan the tensor-core kernel (csrc/pyrdown_umma.cu: 4 levels, W % 80 == 0) uses for a
 * frame shape, so that tests can check its baked border weights on the CPU.  tiles: up to 12 x 8 int32 (first level-4
 * row, level-4 rows, first level-3 row, rows, first level-2 row, rows, first input row, k-steps); codes: 12 x 18 (offset
 * >> 4 of each k-step's weight slice in the blob, -1 past the end); wsp: 3 x 13 horizontal weights of level-2 pixels
 * 0, 1, w2-1; meta: tiles, strips, border slices, blob bytes; blob (may be NULL): weight band + border slices in the
 * UMMA K-major core-matrix layout.  VHR_ERR_UNSUPPORTED when the shape is not eligible. */
int vhr_pyrdown_umma_plan(int H, int W, int32_t* tiles, int32_t* codes, int32_t* wsp, int32_t* meta,
                          uint8_t* blob, int blob_cap);
/* Diagnostics: the 4-level cascade through the tensor-core kernel, which also copies the raw TMEM accumulators of one
 * (item, strip) to d_acc (128 x 240 uint32; item = frame * tiles + tile): exact integers (the vertical 13-tap sums of
 * the tile's level-2 rows over the strip's 240 byte columns), so a test holds the MMA stage to the plan bit for bit. */
int vhr_pyrdown_umma_accumulators(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, float* d_level,
                                  int item, int strip, uint32_t* d_acc, void* stream);

/* ---- EVM: temporal ideal bandpass (no reference code; spec = irfft(mask*rfft)) ----------
 * d_in/d_out float32 (T,P), time-major; keeps rfft bins with f_lo <= k*fps/T <= f_hi
 * (inclusive, the reference's mask style rppg_VIDEO.py:140,196), DC always dropped;
 * output is multiplied by `gain` (the EVM alpha).  In-place (d_out == d_in) allowed. */
int vhr_temporal_bandpass(vhr_ctx* ctx, const float* d_in, float* d_out, int T, int64_t P,
                          double fps, double f_lo, double f_hi, float gain, void* stream);
/* Number of kept bins and first/last kept bin (host helper; -1/-1 if none). */
int vhr_band_bins(int T, double fps, double f_lo, double f_hi, int* k_first, int* k_last);

/* ---- EVM: amplify-collapse-add-back with fused rectangle ROI means ----------------------
 * out(t,y,x,c) = float(frame) + pyrUp^levels(d_level)(t,y,x,c)      (spec = cv2.pyrUp)
 * d_out_f32 (T,H,W,3) float32 and/or d_out_u8 (T,H,W,3) uint8 (clip + round-half-up);
 * either may be NULL.  If K > 0 (<= VHR_MAX_ROIS): d_rects int32 (T,K,4), d_roi_mean float64 (T,K,3) gets
 * the per-channel mean of `out` over each rectangle (replaces get_avg over the cheek
 * slice, rppg_VIDEO.py:60-66,106-110, applied to the magnified frame).  Deterministic:
 * per-tile partial sums, then a fixed-order float64 reduction. */
int vhr_collapse_addback_roi(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames,
                             int T, int H, int W, int levels,
                             float* d_out_f32, uint8_t* d_out_u8,
                             const int32_t* d_rects, int K, double* d_roi_mean,
                             void* stream);
/* The same pass with landmark-POLYGON ROIs (forehead / cheek outlines; BASELINE.json north_star's
 * ROI stage; the reference itself only has the rectangles above, rppg_VIDEO.py:102-103):
 * d_poly int32 (T,K,Vmax,2) vertices (x,y), d_nvert int32 (T,K), 1 <= K <= VHR_MAX_ROIS.  Each
 * polygon is rasterised once per frame by the frozen exact-integer rule of vhr_poly_mask (bit-exact
 * masks) into row bit-masks, and the collapse accumulates the masked sums of the magnified frame in
 * the same pass; d_roi_mean (T,K,3) = masked mean (NaN for an empty mask), d_count int64 (T,K)
 * (optional) = pixels in each mask.
 * In both calls, when d_out_f32 == d_out_u8 == NULL only the image parts under a ROI are
 * evaluated (ROI-only mode: the measurement plugins' case). */
int vhr_collapse_addback_poly(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames,
                              int T, int H, int W, int levels,
                              float* d_out_f32, uint8_t* d_out_u8,
                              const int32_t* d_poly, const int32_t* d_nvert, int K, int Vmax,
                              double* d_roi_mean, int64_t* d_count, void* stream);

/* ---- ROI on raw uint8 frames -------------------------------------------------------------
 * Rectangle means, bit-exact with np.mean(roi[:,:,c]) in float64 (exact integer sums):
 * analysis/measurement/green_avg.py:34, rppg_VIDEO.py:66.  d_mean float64 (T,K,3).
 * d_paint (optional, may be NULL): int32 (T,NP,4) rectangles [x1,y1,x2,y2] whose
 * thickness-2 cv.rectangle outlines are drawn into the frame BEFORE the slice is averaged,
 * in order, with colours paint_rgb[NP][3] -- reproduces the overdraw quirk of
 * rppg_VIDEO.py:54,100-106 without modifying the frame. */
int vhr_roi_mean_rect_u8(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W,
                         const int32_t* d_rects, int K,
                         const int32_t* d_paint, int NP, const uint8_t* paint_rgb /*host, NP*3*/,
                         double* d_mean, void* stream);
/* Polygon ROIs (no reference code; frozen exact-integer rule).
 * d_poly int32 (T,K,Vmax,2) vertex (x,y); d_nvert int32 (T,K).  Means of uint8 frames
 * (exact) or float32 frames over the rasterised mask; d_count int64 (T,K) optional. */
int vhr_roi_mean_poly_u8(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W,
                         const int32_t* d_poly, const int32_t* d_nvert, int K, int Vmax,
                         double* d_mean, int64_t* d_count, void* stream);
int vhr_roi_mean_poly_f32(vhr_ctx* ctx, const float* d_frames, int T, int H, int W,
                          const int32_t* d_poly, const int32_t* d_nvert, int K, int Vmax,
                          double* d_mean, int64_t* d_count, void* stream);
/* Rasterise one polygon set into a uint8 mask (T,K,H,W) -- parity checks of the rule. */
int vhr_poly_mask(vhr_ctx* ctx, int T, int H, int W, const int32_t* d_poly,
                  const int32_t* d_nvert, int K, int Vmax, uint8_t* d_mask, void* stream);

/* ---- BPM estimation ----------------------------------------------------------------------
 * Batched over windows of one or more traces.  d_trace float64 (n_trace, C) time-major: sample i of
 * column c at d_trace[i * ld + c * cs] (contiguous: ld = C, cs = 1; the green means of the K ROIs
 * of a (T,K,3) trace: pointer to element [0,0,1], C = K, ld = 3K, cs = 3).  With C > 1 the chosen
 * bin is the per-column peak of the column with the largest peak (estimate_bpm.py:59-64).
 * Window w covers samples [start[w], start[w]+len[w]); max_len >= every len[w] (it sizes the
 * shared memory; windows longer than max_len yield NaN).  Results: d_bpm float64 (n_win)
 * (NaN where the reference returns None), d_bin int32 (n_win) = chosen FFT/rfft bin.
 *
 * detrend modes: 0 none; 1 float64 mean (rppg_VIDEO.py:399); 2 cast to float32, subtract
 * the float32 pairwise mean, as green_avg.py:42-43 does before estimate_bpm;
 * 3 float32 z-score (x - mean) / std (green_avg_psd_plot.py:174-175).
 */
enum { VHR_DETREND_NONE = 0, VHR_DETREND_F64 = 1, VHR_DETREND_F32 = 2, VHR_DETREND_ZSCORE_F32 = 3 };
/* analysis/utils/estimate_bpm.py:12-65 (mode 0: |X| over freqs>0, N>=8 required) and
 * rppg_VIDEO.py:129-147 (mode 1: mask on signed fftfreq, no length floor). */
enum { VHR_FFT_ANALYSIS = 0, VHR_FFT_VIDEO = 1 };
int vhr_bpm_fft(vhr_ctx* ctx, const double* d_trace, int n_trace, int C, int ld, int cs,
                const int32_t* d_start, const int32_t* d_len, int n_win, int max_len,
                double fs, double f_lo, double f_hi, int detrend, int mode,
                double* d_bpm, int32_t* d_bin, void* stream);

/* Zero-phase IIR/FIR bandpass + Welch peak: rppg_VIDEO.py:241-289 (bandpass_butterworth /
 * bandpass_cheby2 via sosfiltfilt; bandpass_fir via filtfilt) followed by
 * estimate_bpm_welch (:172-203).  The filter is given as coefficients (designed on the
 * host exactly where the reference designs them, sp.butter/cheby2/firwin):
 *   kind 0: none          (Welch on the detrended window; rppg_LIVESTREAM.py:347)
 *   kind 1: SOS           coef = float64 (n_sec,6), sosfiltfilt, odd padding
 *   kind 2: FIR           coef = float64 (n_taps), filtfilt(b,[1.0]), odd padding
 * Windows shorter than or equal to the pad length yield NaN / bin -1 (the reference
 * raises ValueError there).  d_filtered (optional) float64 (n_win, max_len) receives the
 * filtered windows.  welch_seconds = 9 in the reference. */
enum { VHR_FILT_NONE = 0, VHR_FILT_SOS = 1, VHR_FILT_FIR = 2 };
int vhr_bpm_welch(vhr_ctx* ctx, const double* d_trace, int n_trace,
                  const int32_t* d_start, const int32_t* d_len, int n_win,
                  double fs, double f_lo, double f_hi, int detrend,
                  int filt_kind, const double* h_coef, int n_coef,
                  double welch_seconds,
                  double* d_bpm, int32_t* d_bin, double* d_filtered, int max_len,
                  void* stream);

/* Causal SOS filter with carried state: rppg_LIVESTREAM.py:226-251 (live_sos_push).
 * Filters n samples of d_x through h_sos (n_sec,6); d_state float64 (n_sec,2) is read and
 * updated (zero it for live_sos_init / live_sos_reset). */
int vhr_sos_causal(vhr_ctx* ctx, const double* d_x, int n, const double* h_sos, int n_sec,
                   double* d_state, double* d_y, void* stream);

/* ---- ICA measurement: analysis/measurement/ica.py:36-72 ----------------------------
 * Batched FastICA (3 components, parallel algorithm, logcosh, unit-variance whitening; scikit-learn's
 * algorithm, which the reference calls at ica.py:36-44,65) over windows of a mean-BGR trace:
 * d_trace float64 (n_trace,3); window w = rows [start[w], start[w]+len[w]).  Per window: float32 cast and
 * per-channel std normalisation (ddof = 1, ica.py:56-61), centring, whitening, fixed point from the 3x3
 * start matrix h_w_init (row-major; the reference's is RandomState(0).normal(size=(3,3))), at most max_iter
 * iterations, tolerance tol.  d_sources float64 (n_win, max_len, 3) gets the unit-variance sources (NaN
 * padding behind len[w]); d_n_iter int32 (n_win) the iterations used, NEGATED when the fixed point did not
 * reach tol (the reference skips such windows, ica.py:64-69).  Feed d_sources to vhr_bpm_fft with C = 3
 * (ica.py:72).  Float64 arithmetic: tolerance contract, not bit parity (see csrc/ica.cu). */
int vhr_ica_fastica(vhr_ctx* ctx, const double* d_trace, int n_trace, const int32_t* d_start,
                    const int32_t* d_len, int n_win, int max_len, const double* h_w_init,
                    int max_iter, double tol, double* d_sources, int32_t* d_n_iter, void* stream);

/* ---- analysis-harness degradations and metric ---------------------------
 * Additive noise: clip(float(frame) + noise, 0, 255) truncated to uint8 -- analysis/degradation/
 * colour_noise.py:11-24.  The reference draws np.random.normal; here the draw is a counter-based
 * 12-term Irwin-Hall sum (twelve hash bytes of (seed, clip, t0 + t, byte index); mean 0,
 * std = noise_gain_q16 * 255.998 / 65536 LSB, tails to 5.98 std) in pure integer arithmetic, so
 * the CPU oracle regenerates it bit for bit.  In place allowed. */
int vhr_degrade_noise_u8(vhr_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int T, int H, int W,
                         int noise_gain_q16, uint32_t seed, uint32_t clip, int t0, void* stream);
/* Bit-depth quantisation: scale = 256 // 2**bits; (x // scale) * scale -- analysis/degradation/
 * colour_quantisation.py:12-25 (bits > 8 gives all zeros, like NumPy's uint8 // 0). */
int vhr_degrade_quantise_u8(vhr_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, long long n, int bits,
                            void* stream);
/* Step-hold truth alignment + MAE: analysis/utils/video_io.py:80-106 (searchsorted side='right'
 * minus 1, clipped) and analysis/metrics/mae.py:32-36.  d_meas float64 (m,2) [t_sec, bpm];
 * d_aligned float64 (m) truth HR per measurement; d_mae float64 (1). */
int vhr_align_mae(vhr_ctx* ctx, const double* d_truth_t, const double* d_truth_hr, int n_truth,
                  const double* d_meas, int m, double* d_aligned, double* d_mae, void* stream);

/* ---- host-buffer convenience (the reference-facing call: NumPy arrays in and out) -------
 * Whole EVM + ROI path on one clip held in HOST memory: H2D of the frames, the three EVM
 * kernels, fused ROI means, D2H of the (T,K,3) float64 ROI trace.  h_out_f32 may be NULL: the
 * magnified frames are then never materialised (ROI-only collapse).  h_frames should be
 * page-locked (pageable memory works but its copies do not overlap the kernels).  The device
 * arena of these calls stays cached in the context until vhr_trim / vhr_destroy. */
int vhr_evm_roi_host(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels,
                     double fps, double f_lo, double f_hi, float alpha,
                     const int32_t* h_rects, int K, double* h_roi_mean, float* h_out_f32);
/* Polygon form: h_poly int32 (T,K,Vmax,2), h_nvert int32 (T,K); h_count int64 (T,K) optional. */
int vhr_evm_poly_host(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels,
                      double fps, double f_lo, double f_hi, float alpha,
                      const int32_t* h_poly, const int32_t* h_nvert, int K, int Vmax,
                      double* h_roi_mean, int64_t* h_count, float* h_out_f32);
/* Release the context's cached device buffers (host-path arena, scratch arena). */
int vhr_trim(vhr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VHR_B200_H */
