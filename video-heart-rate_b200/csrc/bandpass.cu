// Temporal ideal bandpass along the time axis of each pyramid-level pixel.
//
// No reference code exists for this stage; spec = oracle/evm.py:
// y = irfft(mask * rfft(x), n=T) with mask = (f_lo <= rfftfreq(T,1/fps) <= f_hi), DC dropped
// (inclusive edges like rppg_VIDEO.py:140,196).
//
// Form used here: band-limited DFT pair.  Only the B kept bins are ever needed
// (B = 199 of 901 at T = 1800), so each CTA takes PX pixel series (T x PX floats in shared
// memory), computes X_k = sum_t x_t e^{-2 pi i k t / T} for the kept k only, then
// y_t = (g / T) sum_k w_k Re(X_k e^{+2 pi i k t / T}), w_k = 2 (1 for the Nyquist bin).
// One HBM read and one HBM write of the (T,P) tensor; twiddles come from a T-entry table
// (float2 cos/sin computed in double on the host), indexed by (k t) mod T kept incrementally,
// so there is no argument-reduction error.  Works for ANY T (no radix restriction).
// This stage moves 24 p of the 18 WH + 48 p algorithmic bytes per frame (0.5 % at 1080p).
#include "common.cuh"
#include <math.h>
#include <vector>

namespace {

constexpr int PX = 8;          // pixel series per CTA (one 32-byte sector per time step)
constexpr int NTHREADS = 256;
static_assert(NTHREADS % PX == 0, "tile loader assumes a fixed column per thread");

struct BpArgs {
    const float* in;
    float* out;
    const float2* tw;     // T entries: (cos, sin)(2 pi m / T)
    int T;
    long long P;
    int k0, nb;           // kept bins k0 .. k0+nb-1
    float scale;          // gain / T
    int nyq;              // T/2 if T even else -1
};

__global__ void __launch_bounds__(NTHREADS) bandpass_dft_kernel(const BpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);                    // [T][PX]
    float* Xr = xs + (size_t)a.T * PX;                                 // [nb][PX]  (16-byte aligned: float4 loads)
    float* Xi = Xr + (size_t)a.nb * PX;                                // [nb][PX]
    float2* tw = reinterpret_cast<float2*>(Xi + (size_t)a.nb * PX);    // [T]
    const int T = a.T;
    const long long p0 = (long long)blockIdx.x * PX;
    const int npx = (int)min((long long)PX, a.P - p0);

    for (int m = threadIdx.x; m < T; m += NTHREADS) tw[m] = a.tw[m];
    // coalesced tile load: consecutive threads read consecutive floats of a time step
    // The series' first sample is subtracted on the way in: a constant has no energy in any
    // kept bin (DC is always dropped), and removing the ~150-LSB pedestal keeps the float32
    // accumulation error of the forward sums at the level of the pulse, not of the pedestal.
    {
        const int j = threadIdx.x % PX;                 // NTHREADS % PX == 0: fixed per thread
        const float x0 = (j < npx) ? __ldg(a.in + p0 + j) : 0.0f;
        for (int idx = threadIdx.x; idx < T * PX; idx += NTHREADS) {
            int t = idx / PX;
            xs[idx] = (j < npx) ? __ldg(a.in + (size_t)t * a.P + p0 + j) - x0 : 0.0f;
        }
    }
    __syncthreads();

    // forward: one thread per kept bin, all PX series at once (x_t broadcast from smem)
    for (int b = threadIdx.x; b < a.nb; b += NTHREADS) {
        const int k = a.k0 + b;
        float re[PX], im[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) { re[j] = 0.f; im[j] = 0.f; }
        int m = 0;
        for (int t = 0; t < T; ++t) {
            const float2 w = tw[m];
            const float4 xa = *reinterpret_cast<const float4*>(xs + (size_t)t * PX);
            const float4 xb = *reinterpret_cast<const float4*>(xs + (size_t)t * PX + 4);
            const float xv[PX] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                re[j] = fmaf(xv[j], w.x, re[j]);
                im[j] = fmaf(-xv[j], w.y, im[j]);
            }
            m += k;
            if (m >= T) m -= T;
        }
        const float wk = (k == a.nyq) ? 1.0f : 2.0f;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            Xr[b * PX + j] = re[j] * wk;
            Xi[b * PX + j] = im[j] * wk;
        }
    }
    __syncthreads();

    // inverse: one thread per time step
    for (int t = threadIdx.x; t < T; t += NTHREADS) {
        float acc[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[j] = 0.f;
        int m = (int)(((long long)a.k0 * t) % T);
        for (int b = 0; b < a.nb; ++b) {
            const float2 w = tw[m];
            const float4 ra = *reinterpret_cast<const float4*>(Xr + b * PX);
            const float4 rb = *reinterpret_cast<const float4*>(Xr + b * PX + 4);
            const float4 ia = *reinterpret_cast<const float4*>(Xi + b * PX);
            const float4 ib = *reinterpret_cast<const float4*>(Xi + b * PX + 4);
            const float rv[PX] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const float iv[PX] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
#pragma unroll
            for (int j = 0; j < PX; ++j) acc[j] = fmaf(rv[j], w.x, fmaf(-iv[j], w.y, acc[j]));
            m += t;
            if (m >= T) m -= T;
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) xs[(size_t)t * PX + j] = acc[j] * a.scale;   // own row only
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * PX; idx += NTHREADS) {
        int t = idx / PX, j = idx - t * PX;
        if (j < npx) a.out[(size_t)t * a.P + p0 + j] = xs[idx];
    }
}

__global__ void zero_fill_kernel(float* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = 0.f;
}

}  // namespace

static int ensure_twiddles(vhr_ctx* ctx, int T, cudaStream_t stream) {
    if (ctx->tw_T == T && ctx->tw) return VHR_OK;
    if (ctx->tw) {
        VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
        VHR_CHECK_CUDA(ctx, cudaFree(ctx->tw));
        ctx->tw = nullptr;
        ctx->tw_T = 0;
    }
    std::vector<float2> h((size_t)T);
    for (int m = 0; m < T; ++m) {
        // exact octant symmetry is not needed: double sincos of 2 pi m / T, rounded once
        double ang = 2.0 * M_PI * (double)m / (double)T;
        h[m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    VHR_CHECK_CUDA(ctx, cudaMalloc(&ctx->tw, sizeof(float2) * (size_t)T));
    VHR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->tw, h.data(), sizeof(float2) * (size_t)T, cudaMemcpyHostToDevice, stream));
    VHR_CHECK_CUDA(ctx, cudaStreamSynchronize(stream));   // h goes out of scope
    ctx->tw_T = T;
    return VHR_OK;
}

extern "C" int vhr_temporal_bandpass(vhr_ctx* ctx, const float* d_in, float* d_out, int T, int64_t P, double fps,
                                     double f_lo, double f_hi, float gain, void* stream_) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_in && d_out, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && P >= 1 && fps > 0, "bad shape");
    cudaStream_t stream = (cudaStream_t)stream_;
    int k0 = -1, k1 = -1;
    int nb = vhr_band_bins(T, fps, f_lo, f_hi, &k0, &k1);
    if (nb <= 0) {   // empty band: the filter output is identically zero
        zero_fill_kernel<<<ctx->num_sms * 4, 256, 0, stream>>>(d_out, (long long)T * P);
        return vhr_after_launch(ctx, "zero_fill_kernel");
    }
    int rc = ensure_twiddles(ctx, T, stream);
    if (rc != VHR_OK) return rc;
    BpArgs a;
    a.in = d_in; a.out = d_out; a.tw = ctx->tw; a.T = T; a.P = P;
    a.k0 = k0; a.nb = nb;
    a.scale = (float)((double)gain / (double)T);
    a.nyq = (T % 2 == 0) ? T / 2 : -1;
    size_t smem = (size_t)T * PX * 4 + (size_t)T * 8 + (size_t)nb * PX * 8;
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "bandpass: T=%d needs %zu bytes of shared memory (> %d)", T, smem, ctx->smem_optin);
        return VHR_ERR_UNSUPPORTED;
    }
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(bandpass_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (P + PX - 1) / PX;
    bandpass_dft_kernel<<<(unsigned)blocks, NTHREADS, smem, stream>>>(a);
    return vhr_after_launch(ctx, "bandpass_dft_kernel");
}
