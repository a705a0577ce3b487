// SURVEY.md section 8(f) "next" rows: the analysis harness's degradations and metric, on device.
//
//   vhr_degrade_noise_u8     analysis/degradation/colour_noise.py:11-24 (add_gaussian_noise):
//                            clip(float32(frame) + noise, 0, 255).astype(uint8)  [astype truncates]
//                            The reference draws np.random.normal; here the draw is a 12-term
//                            Irwin-Hall sum (twelve hash bytes of three counter-based words; mean 0,
//                            std sigma, tails to +-5.98 sigma, excess kurtosis -0.1) -- pure integer
//                            arithmetic, so the CPU oracle reproduces it bit for bit.
//   vhr_degrade_quantise_u8  analysis/degradation/colour_quantisation.py:12-25 (quantise_colour):
//                            scale = 256 // 2**bits ; (frame // scale) * scale  (scale == 0, i.e.
//                            bits > 8, gives 0 like NumPy's uint8 // 0)
//   vhr_align_mae            analysis/utils/video_io.py:80-106 (interpolate_hr_to_frames: step-hold,
//                            searchsorted(side='right') - 1, clipped) + analysis/metrics/mae.py:32-36
//                            (mean |pred - truth|, NumPy pairwise float64 mean)
#include "common.cuh"
#include "pairwise.cuh"

namespace {

__device__ __forceinline__ uint32_t dmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) noise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                    long long total, unsigned frame_bytes, uint32_t seed, uint32_t clip,
                                                    int t0, int gain) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < total; base += stride) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long g = base + j;
            if (g >= total) break;
            const unsigned t = (unsigned)(g / frame_bytes);
            const uint32_t idx = (uint32_t)(g - (long long)t * frame_bytes);
            const uint32_t key = dmix32(seed * 0x9E3779B1u + clip * 0x85EBCA77u + (uint32_t)(t0 + (int)t) * 0xC2B2AE3Du + 0x3C6EF372u);
            int s = -1530;                                        // twelve bytes: mean 1530, std sqrt(65535) = 255.998
#pragma unroll
            for (uint32_t w = 0; w < 3; ++w) {
                const uint32_t r = dmix32((key + w * 0x9E3779B9u) ^ (idx * 0x27D4EB2Fu));
                s += (int)__dp4a(r, 0x01010101u, 0u);
            }
            int v = ((int)in[g] * 65536 + s * gain) >> 16;        // floor(x + noise); clip, then astype(uint8) truncates
            v = min(max(v, 0), 255);
            word |= (uint32_t)v << (8 * j);
        }
        if (base + 4 <= total && ((reinterpret_cast<uintptr_t>(out) & 3) == 0)) *reinterpret_cast<uint32_t*>(out + base) = word;
        else for (int j = 0; j < 4 && base + j < total; ++j) out[base + j] = (uint8_t)(word >> (8 * j));
    }
}

__global__ void __launch_bounds__(256) quantise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                       long long total, int scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int x = in[g];
        out[g] = scale > 0 ? (uint8_t)((x / scale) * scale) : (uint8_t)0;
    }
}

__global__ void align_mae_kernel(const double* __restrict__ tt, const double* __restrict__ th, int n,
                                 const double* __restrict__ meas, int m, double* __restrict__ aligned,
                                 double* __restrict__ absdiff, double* __restrict__ mae) {
    // phase 1: step-hold alignment (grid-stride); phase 2 (last block via a second launch): mean
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const double t = meas[2 * i];
        int lo = 0, hi = n;                 // searchsorted(side='right'): first index with tt[idx] > t
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tt[mid] <= t) lo = mid + 1; else hi = mid;
        }
        int idx = lo - 1;
        idx = min(max(idx, 0), n - 1);
        const double hr = th[idx];
        aligned[i] = hr;
        absdiff[i] = fabs(meas[2 * i + 1] - hr);
    }
    (void)mae;
}

__global__ void mean_kernel(const double* __restrict__ x, int m, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = m > 0 ? __ddiv_rn(pairwise_sum_f64(x, m), (double)m) : __longlong_as_double(0x7FF8000000000000ll);
}

}  // namespace

extern "C" int vhr_degrade_noise_u8(vhr_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int T, int H, int W,
                                    int noise_gain_q16, uint32_t seed, uint32_t clip, int t0, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_in && d_out, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, (long long)H * W * 3 < 0xFFFFFFFFll, "frame too large");
    VHR_REQUIRE(ctx, noise_gain_q16 >= 0 && noise_gain_q16 <= 1000000, "noise gain out of range (sigma <= 3900 LSB)");
    const long long total = (long long)T * H * W * 3;
    long long blocks = (total / 4 + 255) / 256;
    if (blocks > ctx->num_sms * 32) blocks = ctx->num_sms * 32;
    if (blocks < 1) blocks = 1;
    noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, total, (unsigned)(H * W * 3), seed, clip, t0, noise_gain_q16);
    return vhr_after_launch(ctx, "noise_kernel");
}

extern "C" int vhr_degrade_quantise_u8(vhr_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, long long n, int bits, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_in && d_out, "null pointer");
    VHR_REQUIRE(ctx, n >= 1 && bits >= 0 && bits <= 30, "bad arguments");
    const int scale = 256 / (1 << bits);                 // 256 // 2**bits ; 0 for bits > 8
    long long blocks = (n + 255) / 256;
    if (blocks > ctx->num_sms * 32) blocks = ctx->num_sms * 32;
    quantise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, scale);
    return vhr_after_launch(ctx, "quantise_kernel");
}

extern "C" int vhr_align_mae(vhr_ctx* ctx, const double* d_truth_t, const double* d_truth_hr, int n_truth,
                             const double* d_meas, int m, double* d_aligned, double* d_mae, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_truth_t && d_truth_hr && d_meas && d_aligned && d_mae, "null pointer");
    VHR_REQUIRE(ctx, n_truth >= 1 && m >= 1, "empty truth or measurement");
    void* scratch = nullptr;
    int rc = vhr_scratch(ctx, sizeof(double) * (size_t)m, &scratch);
    if (rc != VHR_OK) return rc;
    align_mae_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_truth_t, d_truth_hr, n_truth, d_meas, m, d_aligned,
                                                                         reinterpret_cast<double*>(scratch), d_mae);
    rc = vhr_after_launch(ctx, "align_mae_kernel");
    if (rc != VHR_OK) return rc;
    mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<double*>(scratch), m, d_mae);
    return vhr_after_launch(ctx, "mean_kernel");
}
