// BPM estimation, batched over windows: detrend -> (optional zero-phase filter) -> spectrum
// -> first-maximum in-band bin -> BPM.  All arithmetic in float64 (the data is tiny); the
// frequency grid and the band comparison repeat NumPy's own float operations so the
// inclusive band edges, the chosen bin and the returned BPM are identical to the reference:
//
//   vhr_bpm_fft    analysis/utils/estimate_bpm.py:12-65  (and rppg_VIDEO.py:129-147)
//                  with the float32 detrend of analysis/measurement/green_avg.py:42-43
//   vhr_bpm_welch  rppg_VIDEO.py:172-203 estimate_bpm_welch (scipy.signal.welch: hann,
//                  nperseg = int(min(n, 9 fps)), 50 % overlap, per-segment constant detrend,
//                  density scaling, mean over segments) after the zero-phase filters of
//                  rppg_VIDEO.py:241-289 (sosfiltfilt / filtfilt, odd extension, lfilter_zi
//                  initial state)
//   vhr_sos_causal rppg_LIVESTREAM.py:226-251 live_sos_push (sosfilt with carried state)
#include "common.cuh"
#include "pairwise.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace {

constexpr int BT = 256;
constexpr int MAXSEC = 16;
constexpr int MAXTAPS = 128;
constexpr int MAXSPLIT = 8;        // CTAs (one thread-block cluster) sharing the bins of one window
constexpr int MAXCOL = 8;          // columns of a (T,C) signal
constexpr int MAXCOEF = (MAXSEC * 6 > MAXTAPS) ? MAXSEC * 6 : MAXTAPS;

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7FF8000000000000ll); }

// frequency of bin k exactly as numpy.fft.fftfreq / rfftfreq compute it:
//   val = 1.0 / (n * d), d = 1 / fs ; f = k * val
__device__ __forceinline__ double np_freq(int k, int n, double fs) {
    const double d = __ddiv_rn(1.0, fs);
    const double val = __ddiv_rn(1.0, __dmul_rn((double)n, d));
    return __dmul_rn((double)k, val);
}

// detrend a window in place (x: n float64 samples in shared memory); called by every thread of the
// block.  The element-wise steps are spread over the threads; each sum runs on thread 0 in
// NumPy's pairwise order (the order decides the last bit of the mean, hence near-tie bins).
__device__ void detrend_window(double* x, float* xf, int n, int mode) {
    __shared__ double s_m;
    __shared__ float s_mf;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (mode == VHR_DETREND_F64) {
        // np.mean(float64): pairwise sum / n  (rppg_VIDEO.py:399)
        if (tid == 0) s_m = __ddiv_rn(pairwise_sum_f64(x, n), (double)n);
        __syncthreads();
        const double m = s_m;
        for (int i = tid; i < n; i += nt) x[i] = __dsub_rn(x[i], m);
    } else if (mode == VHR_DETREND_F32) {
        // sig = float32(deque); sig - np.mean(sig): float32 pairwise mean (green_avg.py:42-43)
        for (int i = tid; i < n; i += nt) xf[i] = (float)x[i];
        __syncthreads();
        if (tid == 0) s_mf = __fdiv_rn(pairwise_sum_f32(xf, n), (float)n);
        __syncthreads();
        const float m = s_mf;
        for (int i = tid; i < n; i += nt) x[i] = (double)__fsub_rn(xf[i], m);
    } else if (mode == VHR_DETREND_ZSCORE_F32) {
        // (sig - np.mean(sig)) / np.std(sig) on a float32 vector (green_avg_psd_plot.py:174-175):
        // np.std = sqrt(pairwise_sum((x - mean)^2) / n), every step in float32
        for (int i = tid; i < n; i += nt) xf[i] = (float)x[i];
        __syncthreads();
        if (tid == 0) s_mf = __fdiv_rn(pairwise_sum_f32(xf, n), (float)n);
        __syncthreads();
        const float m = s_mf;
        for (int i = tid; i < n; i += nt) { const float d = __fsub_rn(xf[i], m); x[i] = (double)d; xf[i] = __fmul_rn(d, d); }
        __syncthreads();
        if (tid == 0) s_mf = __fsqrt_rn(__fdiv_rn(pairwise_sum_f32(xf, n), (float)n));
        __syncthreads();
        const float sd = s_mf;
        for (int i = tid; i < n; i += nt) x[i] = (double)__fdiv_rn((float)x[i], sd);
    }
    __syncthreads();
}

// |X_k|^2 of x[0..n) (float64), twiddles from a table tw[m] = (cos, sin)(2 pi m / n):
// the 32 lanes of a warp stride over the samples (butterfly reduction: every lane ends with the
// same value)
__device__ __forceinline__ double dft_power_warp(const double* x, int n, int k, const double2* tw) {
    const int lane = threadIdx.x & 31;
    double re = 0., im = 0.;
    const int dk = (int)(((long long)k * 32) % n);
    int m = (int)(((long long)k * lane) % n);
    for (int i = lane; i < n; i += 32) {
        const double2 w = tw[m];
        re = fma(x[i], w.x, re);
        im = fma(-x[i], w.y, im);
        m += dk;
        if (m >= n) m -= n;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, d);
        im += __shfl_xor_sync(0xffffffffu, im, d);
    }
    return re * re + im * im;
}

struct ArgMax {
    double v;
    int k;
};
// first-maximum (np.argmax) reduction over the block; entries with k < 0 are empty
__device__ ArgMax block_argmax(ArgMax a, ArgMax* sh) {
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, a.v, d);
        const int ok = __shfl_xor_sync(0xffffffffu, a.k, d);
        if (ok >= 0 && (a.k < 0 || ov > a.v || (ov == a.v && ok < a.k))) { a.v = ov; a.k = ok; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = a;
    __syncthreads();
    ArgMax r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        const ArgMax o = sh[w];
        if (o.k >= 0 && (r.k < 0 || o.v > r.v || (o.v == r.v && o.k < r.k))) r = o;
    }
    return r;   // same in every thread
}

struct FftArgs {
    const double* trace;
    int n_trace, C, ld, cs;
    const int32_t* start;
    const int32_t* len;
    double fs, f_lo, f_hi;
    int detrend, mode;
    double* bpm;
    int32_t* bin;
    int max_len;
};

// One thread-block CLUSTER per window: every CTA of the cluster loads and detrends the window (cheap, and it keeps the
// pairwise-sum order), takes a contiguous share of the candidate bins, and CTA 0 merges the shares through
// distributed shared memory in bin order (so ties still resolve to the lowest index, as np.argmax does).  A single
// long window (the whole-clip estimate of the measurement plugins: 1 800 samples, 199 bins) no longer runs on one SM.
__global__ void __launch_bounds__(BT) bpm_fft_kernel(const FftArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int nsplit = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ ArgMax shm[BT / 32];
    __shared__ ArgMax share[MAXCOL];                                 // this CTA's best bin per column (read by CTA 0)
    double2* tw = reinterpret_cast<double2*>(sm);                    // [max_len]  (16-byte aligned first)
    double* x = reinterpret_cast<double*>(tw + a.max_len);           // [max_len]
    float* xf = reinterpret_cast<float*>(x + a.max_len);             // [max_len]
    const int w = blockIdx.x / nsplit;
    const int s = a.start[w], n = a.len[w];
    const bool bad = (n < 1) || n > a.max_len || s < 0 || s + n > a.n_trace || (a.mode == VHR_FFT_ANALYSIS && n < 8);
    if (bad) {                                                       // the whole cluster takes this exit
        if (threadIdx.x == 0 && rank == 0) { a.bpm[w] = qnan(); a.bin[w] = -1; }
        return;
    }
    for (int m = threadIdx.x; m < n; m += BT) {
        double sv, cv;
        sincospi(2.0 * (double)m / (double)n, &sv, &cv);
        tw[m] = make_double2(cv, sv);
    }
    // positive-frequency bins 1..(n-1)/2 (freqs > 0 in fftfreq order), inclusive band.  The VIDEO
    // estimator masks the SIGNED fftfreq grid (rppg_VIDEO.py:137-140): with f_lo <= 0 the DC bin and
    // the negative-frequency bins (n-1)/2+1 .. n-1 are candidates too.
    const bool all_bins = a.mode == VHR_FFT_VIDEO && !(a.f_lo > 0.0);
    const int kpos = (n - 1) / 2;
    const int kmin = all_bins ? 0 : 1;
    const int kmax = all_bins ? n - 1 : kpos;
    // this CTA's share of the candidates: the in-band bins are one run k_lo..k_hi of the positive half (plus, in
    // all_bins mode, anything else): split the index range evenly, ascending with the rank
    int k_lo = kmin, k_hi = kmax;
    if (!all_bins) {
        // f(k) is monotonic in k: bisect with the very comparisons the bin loop makes
        int lo = kmin, hi = kmax + 1;                   // first k with f(k) >= f_lo
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (np_freq(mid, n, a.fs) >= a.f_lo) hi = mid; else lo = mid + 1; }
        k_lo = lo;
        lo = kmin - 1; hi = kmax;                       // last k with f(k) <= f_hi
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (np_freq(mid, n, a.fs) <= a.f_hi) lo = mid; else hi = mid - 1; }
        k_hi = lo;
    }
    const int per = (k_hi - k_lo + 1 + nsplit - 1) / nsplit;
    const int my_lo = k_lo + rank * per, my_hi = min(k_hi, my_lo + per - 1);
    for (int c = 0; c < a.C; ++c) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += BT) x[i] = a.trace[(size_t)(s + i) * a.ld + (size_t)c * a.cs];
        __syncthreads();
        detrend_window(x, xf, n, a.detrend);
        ArgMax mine;
        mine.v = 0.;
        mine.k = -1;
        // one warp per bin (long windows: a single window must not serialise 1800 samples per thread)
        for (int k = my_lo + (int)(threadIdx.x >> 5); k <= my_hi; k += BT / 32) {
            const double f = np_freq(k <= kpos ? k : k - n, n, a.fs);
            if (f >= a.f_lo && f <= a.f_hi) {           // warp-uniform
                // a negative-frequency bin has the magnitude of its mirror n - k (real input; NumPy's fft of a
                // real array mirrors it exactly), so ties between the two resolve to the lower index as in np.argmax
                const double p = dft_power_warp(x, n, k <= kpos ? k : n - k, tw);
                if (mine.k < 0 || p > mine.v) { mine.v = p; mine.k = k; }
            }
        }
        const ArgMax col = block_argmax(mine, shm);
        if (threadIdx.x == 0 && c < MAXCOL) share[c] = col;
    }
    cluster.sync();                                      // every CTA's shares are written
    if (rank == 0 && threadIdx.x == 0) {
        ArgMax best;
        best.v = 0.;
        best.k = -1;
        for (int c = 0; c < a.C; ++c) {
            ArgMax col;
            col.v = 0.;
            col.k = -1;
            for (int r = 0; r < nsplit; ++r) {           // ascending bins: a later share wins only if strictly larger
                const ArgMax o = *cluster.map_shared_rank(&share[c], r);
                if (o.k >= 0 && (col.k < 0 || o.v > col.v)) col = o;
            }
            // best channel: first maximum of the per-channel peak magnitudes
            if (col.k >= 0 && (best.k < 0 || col.v > best.v)) best = col;
        }
        if (best.k < 0) { a.bpm[w] = qnan(); a.bin[w] = -1; }
        else { a.bpm[w] = __dmul_rn(np_freq(best.k <= kpos ? best.k : best.k - n, n, a.fs), 60.0); a.bin[w] = best.k; }
    }
    cluster.sync();                                      // keep the shares alive until CTA 0 has read them
}

// ---- Welch path ----------------------------------------------------------------------------
struct WelchArgs {
    const double* trace;
    int n_trace;
    const int32_t* start;
    const int32_t* len;
    double fs, f_lo, f_hi;
    int detrend;
    int filt_kind, n_coef;        // SOS: n_coef = n_sections ; FIR: n_coef = taps
    double welch_seconds;
    double* bpm;
    int32_t* bin;
    double* filtered;
    int max_len, max_ext;
    double coef[MAXCOEF];          // by value: no shared constant bank between contexts
};

// scipy.signal.sosfilt_zi / lfilter_zi for one biquad (a0 == 1): zi = [c1 + c2, c2] with
// c = b - y_inf a, y_inf = sum(b)/sum(a); scale accumulates the DC gain of earlier sections.
__device__ void sos_zi(const double* sos, int nsec, double (*zi)[2]) {
    double scale = 1.0;
    for (int s = 0; s < nsec; ++s) {
        const double* b = sos + 6 * s;
        const double* a_ = b + 3;
        const double sb = __dadd_rn(__dadd_rn(b[0], b[1]), b[2]);
        const double sa = __dadd_rn(__dadd_rn(a_[0], a_[1]), a_[2]);
        const double yinf = __ddiv_rn(sb, sa);
        const double c1 = __dsub_rn(b[1], __dmul_rn(yinf, a_[1]));
        const double c2 = __dsub_rn(b[2], __dmul_rn(yinf, a_[2]));
        zi[s][0] = __dmul_rn(scale, __dadd_rn(c2, c1));
        zi[s][1] = __dmul_rn(scale, c2);
        scale = __dmul_rn(scale, yinf);
    }
}

// scipy's _sosfilt inner loop (direct form II transposed), in place over y[0..n)
__device__ void sosfilt_run(const double* sos, int nsec, double (*z)[2], double* y, int n, int stride) {
    for (int i = 0; i < n; ++i) {
        double xc = y[(ptrdiff_t)i * stride];
        for (int s = 0; s < nsec; ++s) {
            const double* c = sos + 6 * s;
            const double xn = xc;
            xc = __dadd_rn(__dmul_rn(c[0], xn), z[s][0]);
            z[s][0] = __dadd_rn(__dsub_rn(__dmul_rn(c[1], xn), __dmul_rn(c[4], xc)), z[s][1]);
            z[s][1] = __dsub_rn(__dmul_rn(c[2], xn), __dmul_rn(c[5], xc));
        }
        y[(ptrdiff_t)i * stride] = xc;
    }
}

__global__ void __launch_bounds__(BT) bpm_welch_kernel(const __grid_constant__ WelchArgs a) {
    const double* c_coef = a.coef;
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ ArgMax shm[BT / 32];
    double2* tw = reinterpret_cast<double2*>(sm);                     // [max_len]   (16-byte aligned first)
    double* x = reinterpret_cast<double*>(tw + a.max_len);            // [max_len]   window / filtered
    double* ext = x + a.max_len;                                      // [max_ext]   padded signal
    double* ext2 = ext + a.max_ext;                                   // [max_ext]   FIR scratch
    double* win = ext2 + a.max_ext;                                   // [max_len]   hann
    float* xf = reinterpret_cast<float*>(win + a.max_len);            // [max_len]
    const int w = blockIdx.x;
    const int s = a.start[w], n = a.len[w];
    int edge = 0;
    if (a.filt_kind == VHR_FILT_SOS) {
        // ntaps = 2 n_sections + 1 - min(#(b2 == 0), #(a2 == 0)); edge = 3 ntaps  (sosfiltfilt)
        int zb = 0, za = 0;
        for (int q = 0; q < a.n_coef; ++q) { zb += c_coef[6 * q + 2] == 0.0; za += c_coef[6 * q + 5] == 0.0; }
        edge = 3 * (2 * a.n_coef + 1 - min(zb, za));
    } else if (a.filt_kind == VHR_FILT_FIR) {
        edge = 3 * a.n_coef;                                            // max(len(a), len(b)) * 3
    }
    const bool bad = n < 1 || n > a.max_len || s < 0 || s + n > a.n_trace || (a.filt_kind != VHR_FILT_NONE && n <= edge);
    if (bad) {   // the reference raises ValueError here (rppg_VIDEO.py:404 at 5 FPS)
        if (threadIdx.x == 0) { a.bpm[w] = qnan(); a.bin[w] = -1; }
        if (a.filtered)
            for (int i = threadIdx.x; i < a.max_len; i += BT) a.filtered[(size_t)w * a.max_len + i] = qnan();
        return;
    }
    for (int i = threadIdx.x; i < n; i += BT) x[i] = a.trace[s + i];
    __syncthreads();
    detrend_window(x, xf, n, a.detrend);

    const int ne = n + 2 * edge;
    if (a.filt_kind != VHR_FILT_NONE) {
        // odd extension (scipy _arraytools.odd_ext)
        for (int i = threadIdx.x; i < ne; i += BT) {
            double v;
            if (i < edge) v = __dsub_rn(__dmul_rn(2.0, x[0]), x[edge - i]);
            else if (i < edge + n) v = x[i - edge];
            else v = __dsub_rn(__dmul_rn(2.0, x[n - 1]), x[n - 2 - (i - edge - n)]);
            ext[i] = v;
        }
        __syncthreads();
    }
    if (a.filt_kind == VHR_FILT_SOS) {
        if (threadIdx.x == 0) {
            double zi[MAXSEC][2], z[MAXSEC][2];
            sos_zi(c_coef, a.n_coef, zi);
            const double x0 = ext[0];
            for (int q = 0; q < a.n_coef; ++q) { z[q][0] = __dmul_rn(zi[q][0], x0); z[q][1] = __dmul_rn(zi[q][1], x0); }
            sosfilt_run(c_coef, a.n_coef, z, ext, ne, 1);
            const double y0 = ext[ne - 1];
            for (int q = 0; q < a.n_coef; ++q) { z[q][0] = __dmul_rn(zi[q][0], y0); z[q][1] = __dmul_rn(zi[q][1], y0); }
            sosfilt_run(c_coef, a.n_coef, z, ext + (ne - 1), ne, -1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += BT) x[i] = ext[edge + i];
        __syncthreads();
    } else if (a.filt_kind == VHR_FILT_FIR) {
        // lfilter(b,[1]) with zi = lfilter_zi(b,[1]) * x0 is a convolution over the signal
        // left-padded with the constant x0; forward into ext2, backward into ext.
        const int nt = a.n_coef;
        for (int i = threadIdx.x; i < ne; i += BT) {
            double acc = 0.;
            const double x0 = ext[0];
            for (int k = 0; k < nt; ++k) acc = fma(c_coef[k], (i - k >= 0) ? ext[i - k] : x0, acc);
            ext2[i] = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ne; i += BT) {
            // reversed sequence r[j] = ext2[ne-1-j]; output j maps back to index ne-1-j
            const int j = ne - 1 - i;
            double acc = 0.;
            const double y0 = ext2[ne - 1];
            for (int k = 0; k < nt; ++k) acc = fma(c_coef[k], (j - k >= 0) ? ext2[ne - 1 - (j - k)] : y0, acc);
            ext[i] = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += BT) x[i] = ext[edge + i];
        __syncthreads();
    }
    if (a.filtered)
        for (int i = threadIdx.x; i < a.max_len; i += BT)
            a.filtered[(size_t)w * a.max_len + i] = (i < n) ? x[i] : qnan();

    // estimate_bpm_welch: float32 cast, minus nanmean (float32), then welch in float64 here
    // (the float32 copy goes to xf; x is rewritten only after every thread has stored its part of
    // `filtered` above and has read the shared mean)
    __shared__ float welch_mean;
    for (int i = threadIdx.x; i < n; i += BT) xf[i] = (float)x[i];
    __syncthreads();
    if (threadIdx.x == 0) welch_mean = __fdiv_rn(pairwise_sum_f32(xf, n), (float)n);      // NumPy's pairwise order
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += BT) x[i] = (double)__fsub_rn(xf[i], welch_mean);
    int nperseg = (int)fmin((double)n, __dmul_rn(a.fs, a.welch_seconds));     // int(min(len(x), fps*9))
    if (nperseg < 1) nperseg = 1;
    const int noverlap = nperseg / 2;
    const int step = nperseg - noverlap;
    const int nseg = (n - nperseg) / step + 1;
    for (int m = threadIdx.x; m < nperseg; m += BT) {
        double sv, cv;
        sincospi(2.0 * (double)m / (double)nperseg, &sv, &cv);
        tw[m] = make_double2(cv, sv);
        win[m] = 0.5 - 0.5 * cv;                                               // periodic hann
    }
    __syncthreads();
    // per-segment mean (detrend='constant') -- reuse ext as [nseg] means
    for (int g = threadIdx.x; g < nseg; g += BT) {
        double sum = 0.;
        for (int i = 0; i < nperseg; ++i) sum += x[g * step + i];
        ext[g] = sum / (double)nperseg;
    }
    __syncthreads();
    const int kmax = nperseg / 2;         // rfft bins 0..nperseg/2
    ArgMax mine;
    mine.v = 0.;
    mine.k = -1;
    for (int k = threadIdx.x; k <= kmax; k += BT) {
        const double f = np_freq(k, nperseg, a.fs);
        if (!(f >= a.f_lo && f <= a.f_hi)) continue;
        double psd = 0.;
        for (int g = 0; g < nseg; ++g) {
            double re = 0., im = 0.;
            int m = 0;
            const double mu = ext[g];
            const double* xs = x + g * step;
            for (int i = 0; i < nperseg; ++i) {
                const double v = (xs[i] - mu) * win[i];
                const double2 tq = tw[m];
                re = fma(v, tq.x, re);
                im = fma(-v, tq.y, im);
                m += k;
                if (m >= nperseg) m -= nperseg;
            }
            psd += re * re + im * im;
        }
        // common positive factors (scale, 1/nseg) do not move the argmax; the one-sided
        // doubling does: every bin but DC and (even nperseg) Nyquist is doubled
        const bool single = (k == 0) || ((nperseg % 2 == 0) && k == kmax);
        if (!single) psd *= 2.0;
        if (mine.k < 0 || psd > mine.v) { mine.v = psd; mine.k = k; }
    }
    const ArgMax best = block_argmax(mine, shm);
    if (threadIdx.x == 0) {
        if (best.k < 0) { a.bpm[w] = qnan(); a.bin[w] = -1; }
        else { a.bpm[w] = __dmul_rn(np_freq(best.k, nperseg, a.fs), 60.0); a.bin[w] = best.k; }
    }
}

struct SosCoef {
    double c[MAXSEC * 6];
};
__global__ void sos_causal_kernel(const double* __restrict__ x, int n, int nsec, double* __restrict__ state,
                                  double* __restrict__ y, const __grid_constant__ SosCoef cf) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* c_coef = cf.c;
    double z[MAXSEC][2];
    for (int s = 0; s < nsec; ++s) { z[s][0] = state[2 * s]; z[s][1] = state[2 * s + 1]; }
    for (int i = 0; i < n; ++i) {
        double xc = x[i];
        for (int s = 0; s < nsec; ++s) {
            const double* c = c_coef + 6 * s;
            const double xn = xc;
            xc = __dadd_rn(__dmul_rn(c[0], xn), z[s][0]);
            z[s][0] = __dadd_rn(__dsub_rn(__dmul_rn(c[1], xn), __dmul_rn(c[4], xc)), z[s][1]);
            z[s][1] = __dsub_rn(__dmul_rn(c[2], xn), __dmul_rn(c[5], xc));
        }
        y[i] = xc;
    }
    for (int s = 0; s < nsec; ++s) { state[2 * s] = z[s][0]; state[2 * s + 1] = z[s][1]; }
}

}  // namespace

// The windows live on the device; their maximum length sizes the shared memory.  The caller
// (Python host) knows it, so it is passed explicitly instead of a device round trip.
static int check_windows(vhr_ctx* ctx, int n_trace, int n_win) {
    VHR_REQUIRE(ctx, n_trace >= 1 && n_win >= 1, "empty trace or window list");
    return VHR_OK;
}

extern "C" int vhr_bpm_fft(vhr_ctx* ctx, const double* d_trace, int n_trace, int C, int ld, int cs, const int32_t* d_start,
                           const int32_t* d_len, int n_win, int max_len, double fs, double f_lo, double f_hi,
                           int detrend, int mode, double* d_bpm, int32_t* d_bin, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_trace && d_start && d_len && d_bpm && d_bin, "null pointer");
    VHR_REQUIRE(ctx, C >= 1 && C <= MAXCOL && cs >= 1 && ld >= 1 && fs > 0, "bad arguments (1..8 columns)");
    VHR_REQUIRE(ctx, detrend >= 0 && detrend <= 3 && (mode == 0 || mode == 1), "bad detrend/mode");
    int rc = check_windows(ctx, n_trace, n_win);
    if (rc != VHR_OK) return rc;
    FftArgs a;
    a.trace = d_trace; a.n_trace = n_trace; a.C = C; a.ld = ld; a.cs = cs; a.start = d_start; a.len = d_len;
    a.fs = fs; a.f_lo = f_lo; a.f_hi = f_hi; a.detrend = detrend; a.mode = mode; a.bpm = d_bpm; a.bin = d_bin;
    VHR_REQUIRE(ctx, max_len >= 1 && max_len <= n_trace, "max_len must be 1..n_trace");
    a.max_len = max_len;
    const size_t smem = (size_t)a.max_len * (8 + 16 + 4);
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "bpm_fft: windows of %d samples need %zu bytes of shared memory", max_len, smem);
        return VHR_ERR_UNSUPPORTED;
    }
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(bpm_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // few windows: spread each over a cluster; many windows already fill the GPU (and the split repeats the detrend)
    int nsplit = 1;
    while (nsplit < MAXSPLIT && (long long)n_win * nsplit * 2 <= 2LL * ctx->num_sms) nsplit *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_win * nsplit, 1, 1);
    cfg.blockDim = dim3(BT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nsplit;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VHR_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, bpm_fft_kernel, a));
    return vhr_after_launch(ctx, "bpm_fft_kernel");
}

extern "C" int vhr_bpm_welch(vhr_ctx* ctx, const double* d_trace, int n_trace, const int32_t* d_start,
                             const int32_t* d_len, int n_win, double fs, double f_lo, double f_hi, int detrend,
                             int filt_kind, const double* h_coef, int n_coef, double welch_seconds, double* d_bpm,
                             int32_t* d_bin, double* d_filtered, int max_len, void* stream_) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_trace && d_start && d_len && d_bpm && d_bin, "null pointer");
    VHR_REQUIRE(ctx, fs > 0 && welch_seconds > 0, "bad arguments");
    VHR_REQUIRE(ctx, detrend >= 0 && detrend <= 3, "bad detrend");
    VHR_REQUIRE(ctx, max_len >= 1 && max_len <= n_trace, "max_len must be 1..n_trace");
    int rc = check_windows(ctx, n_trace, n_win);
    if (rc != VHR_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    int edge = 0;
    if (filt_kind == VHR_FILT_SOS) {
        VHR_REQUIRE(ctx, h_coef && n_coef >= 1 && n_coef <= MAXSEC, "SOS: 1..16 sections");
        edge = 3 * (2 * n_coef + 1);
    } else if (filt_kind == VHR_FILT_FIR) {
        VHR_REQUIRE(ctx, h_coef && n_coef >= 1 && n_coef <= MAXTAPS, "FIR: 1..128 taps");
        edge = 3 * n_coef;
    } else {
        VHR_REQUIRE(ctx, filt_kind == VHR_FILT_NONE, "bad filter kind");
    }
    WelchArgs a;
    a.trace = d_trace; a.n_trace = n_trace; a.start = d_start; a.len = d_len;
    a.fs = fs; a.f_lo = f_lo; a.f_hi = f_hi; a.detrend = detrend;
    a.filt_kind = filt_kind; a.n_coef = n_coef; a.welch_seconds = welch_seconds;
    a.bpm = d_bpm; a.bin = d_bin; a.filtered = d_filtered;
    a.max_len = max_len;
    a.max_ext = max_len + 2 * edge;
    memset(a.coef, 0, sizeof(a.coef));
    if (filt_kind == VHR_FILT_SOS) memcpy(a.coef, h_coef, sizeof(double) * 6 * n_coef);
    if (filt_kind == VHR_FILT_FIR) memcpy(a.coef, h_coef, sizeof(double) * n_coef);
    const size_t smem = (size_t)a.max_len * (8 + 16 + 8 + 4) + (size_t)a.max_ext * 16 + 16;
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "bpm_welch: windows of %d samples need %zu bytes of shared memory", max_len, smem);
        return VHR_ERR_UNSUPPORTED;
    }
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(bpm_welch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bpm_welch_kernel<<<n_win, BT, smem, stream>>>(a);
    return vhr_after_launch(ctx, "bpm_welch_kernel");
}

extern "C" int vhr_sos_causal(vhr_ctx* ctx, const double* d_x, int n, const double* h_sos, int n_sec,
                              double* d_state, double* d_y, void* stream_) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_x && h_sos && d_state && d_y, "null pointer");
    VHR_REQUIRE(ctx, n >= 1 && n_sec >= 1 && n_sec <= MAXSEC, "bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    SosCoef cf;
    memset(&cf, 0, sizeof(cf));
    memcpy(cf.c, h_sos, sizeof(double) * 6 * n_sec);
    sos_causal_kernel<<<1, 32, 0, stream>>>(d_x, n, n_sec, d_state, d_y, cf);
    return vhr_after_launch(ctx, "sos_causal_kernel");
}
