// Temporal ideal bandpass along the time axis of each pyramid-level pixel.
//
// No reference code exists for this stage; spec = oracle/evm.py:
// y = irfft(mask * rfft(x), n=T) with mask = (f_lo <= rfftfreq(T,1/fps) <= f_hi), DC dropped
// (inclusive edges like rppg_VIDEO.py:140,196).
//
// Two forms, same result:
//  (1) T = 2^a 3^b 5^c (1800, 300, 150 ...): in-place mixed-radix FFT in shared memory, two real
//      series packed into one complex transform: decimation-in-frequency forward (natural ->
//      digit-reversed), multiply by a host-built mask laid out in digit-reversed order (kept
//      bins k and T-k, gain/T folded in), decimation-in-time inverse (digit-reversed -> natural).
//      No reordering pass, one buffer, radix-5/3/4/2 butterflies; ~20x fewer instructions than
//      form (2) at T = 1800.
//  (2) any other T: band-limited DFT pair (below).
//
// Form (2): band-limited DFT pair.  Only the B kept bins are ever needed
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
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------- form (1)
constexpr int FG = 4;            // complex series per CTA (= 8 pixel series, one 32-byte sector per time step)
constexpr int FT = 256;
constexpr int MAXPASS = 16;

struct FftArgs {
    const float* in;
    float* out;
    const float2* tw;        // T entries (cos, sin)(2 pi m / T)
    const float* mask;       // T entries, digit-reversed order: gain/T on kept bins, else 0
    int T;
    long long P;
    int npass;
    int radix[MAXPASS];
    int sub[MAXPASS];        // sub-transform length n at this pass (n_0 = T)
    unsigned inv_m[MAXPASS]; // magic reciprocal of m = n / r
    int pair_ok;             // P even: float2 global accesses allowed
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// exp(+2 pi i m / N) for the inner twiddles of the composite radices (m is a compile-time constant
// after unrolling, so the switch folds to two literals)
template <int N>
__device__ __forceinline__ float2 unit_root(int m) {
    if constexpr (N == 8) {
        switch (m) {
        case 0: return make_float2(1.000000000e+00f, 0.000000000e+00f);
        case 1: return make_float2(7.071067812e-01f, 7.071067812e-01f);
        case 2: return make_float2(0.0f, 1.000000000e+00f);
        case 3: return make_float2(-7.071067812e-01f, 7.071067812e-01f);
        case 4: return make_float2(-1.000000000e+00f, 0.0f);
        case 5: return make_float2(-7.071067812e-01f, -7.071067812e-01f);
        case 6: return make_float2(0.0f, -1.000000000e+00f);
        case 7: return make_float2(7.071067812e-01f, -7.071067812e-01f);
        }
    }
    if constexpr (N == 9) {
        switch (m) {
        case 0: return make_float2(1.000000000e+00f, 0.000000000e+00f);
        case 1: return make_float2(7.660444431e-01f, 6.427876097e-01f);
        case 2: return make_float2(1.736481777e-01f, 9.848077530e-01f);
        case 3: return make_float2(-5.000000000e-01f, 8.660254038e-01f);
        case 4: return make_float2(-9.396926208e-01f, 3.420201433e-01f);
        case 5: return make_float2(-9.396926208e-01f, -3.420201433e-01f);
        case 6: return make_float2(-5.000000000e-01f, -8.660254038e-01f);
        case 7: return make_float2(1.736481777e-01f, -9.848077530e-01f);
        case 8: return make_float2(7.660444431e-01f, -6.427876097e-01f);
        }
    }
    static_assert(N == 8 || N == 9, "unit_root");
    return make_float2(1.0f, 0.0f);
}

template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2 (&v)[R]);

// DFT of size R1*R2 in natural order, in registers: n = R2 n1 + n2, k = k1 + R1 k2;
// R1-point DFTs over n1, inner twiddles W_N^(n2 k1), R2-point DFTs over n2
template <int R1, int R2, bool INV>
__device__ __forceinline__ void dft_comp(float2 (&v)[R1 * R2]) {
    float2 A[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        float2 u[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) u[n1] = v[R2 * n1 + n2];
        dft_small<R1, INV>(u);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            if (n2 * k1 != 0) {
                const float2 w = unit_root<R1 * R2>((n2 * k1) % (R1 * R2));
                A[n2][k1] = cmul(u[k1], make_float2(w.x, INV ? w.y : -w.y));
            } else {
                A[n2][k1] = u[k1];
            }
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        float2 t[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) t[n2] = A[n2][k1];
        dft_small<R2, INV>(t);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
}

template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2 (&v)[R]) {
    if constexpr (R == 8) {
        dft_comp<4, 2, INV>(v);
    } else if constexpr (R == 9) {
        dft_comp<3, 3, INV>(v);
    } else if constexpr (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    } else if constexpr (R == 3) {
        const float2 t1 = cadd(v[1], v[2]);
        const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 d = csub(v[1], v[2]);
        const float2 t3 = mul_mi<INV>(make_float2(0.8660254037844386f * d.x, 0.8660254037844386f * d.y));
        v[0] = cadd(v[0], t1); v[1] = cadd(t2, t3); v[2] = csub(t2, t3);
    } else if constexpr (R == 4) {
        const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]);
        const float2 t3 = mul_mi<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2); v[2] = csub(t0, t2); v[1] = cadd(t1, t3); v[3] = csub(t1, t3);
    } else {
        static_assert(R == 5, "radix");
        const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f, s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;
        const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const float2 n1 = mul_mi<INV>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        const float2 n2 = mul_mi<INV>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
        v[1] = cadd(m1, n1); v[4] = csub(m1, n1); v[2] = cadd(m2, n2); v[3] = csub(m2, n2);
    }
}

// one radix-R pass over all FG series held in z[FG][T]; DIF (forward) twiddles after the
// butterfly, DIT (inverse) conjugate twiddles before it
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(float2* __restrict__ z, const float2* __restrict__ tws, int T, int n,
                                         unsigned inv_m, const float* __restrict__ mask) {
    const int m = n / R, step = T / n, nb = T / R;
    for (int b = threadIdx.x; b < nb; b += FT) {
        const int block = inv_m ? (int)__umulhi((unsigned)b, inv_m) : b;     // inv_m == 0 encodes m == 1
        const int j = b - block * m;
        const int base = block * n + j;
        float2 w[R];
#pragma unroll
        for (int q = 1; q < R; ++q) {
            const float2 t = tws[j * q * step];
            w[q] = make_float2(t.x, INV ? t.y : -t.y);
        }
#pragma unroll
        for (int s = 0; s < FG; ++s) {
            float2* zs = z + (size_t)s * T + base;
            float2 v[R];
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = zs[q * m];
            if (INV) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            dft_small<R, INV>(v);
            if (!INV) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            if (mask) {          // last forward pass: the band mask (gain / T on kept bins) rides on the store
#pragma unroll
                for (int q = 0; q < R; ++q) { const float mk = __ldg(mask + base + q * m); v[q] = make_float2(v[q].x * mk, v[q].y * mk); }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) zs[q * m] = v[q];
        }
    }
}

// the same pass for the composite radices (8, 9): the outer twiddles of a butterfly are loaded once
// and reused for the FG series (the kernel is bound by shared-memory wavefronts, and the R - 1
// twiddle loads per butterfly and series nearly doubled the read traffic of these passes)
template <int R, bool INV>
__device__ __forceinline__ void fft_pass_big(float2* __restrict__ z, const float2* __restrict__ tws, int T, int n,
                                             unsigned inv_m, const float* __restrict__ mask) {
    const int m = n / R, step = T / n, nb = T / R;
    for (int b = threadIdx.x; b < nb; b += FT) {
        const int block = inv_m ? (int)__umulhi((unsigned)b, inv_m) : b;     // inv_m == 0 encodes m == 1
        const int j = b - block * m;
        const int base = block * n + j;
        const int js = j * step;
        float2 w[R];
        if (js != 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) { const float2 t = tws[js * q]; w[q] = make_float2(t.x, INV ? t.y : -t.y); }
        }
        float mk[R];
        if (mask) {
#pragma unroll
            for (int q = 0; q < R; ++q) mk[q] = __ldg(mask + base + q * m);
        }
#pragma unroll 1
        for (int s = 0; s < FG; ++s) {
            float2* zs = z + (size_t)s * T + base;
            float2 v[R];
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = zs[q * m];
            if (INV && js != 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            dft_small<R, INV>(v);
            if (!INV && js != 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            if (mask) {
#pragma unroll
                for (int q = 0; q < R; ++q) v[q] = make_float2(v[q].x * mk[q], v[q].y * mk[q]);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) zs[q * m] = v[q];
        }
    }
}

template <bool INV>
__device__ __forceinline__ void fft_dispatch(float2* z, const float2* tws, int T, int r, int n, unsigned inv_m,
                                             const float* mask) {
    switch (r) {
        case 9: fft_pass_big<9, INV>(z, tws, T, n, inv_m, mask); break;
        case 8: fft_pass_big<8, INV>(z, tws, T, n, inv_m, mask); break;
        case 5: fft_pass<5, INV>(z, tws, T, n, inv_m, mask); break;
        case 3: fft_pass<3, INV>(z, tws, T, n, inv_m, mask); break;
        case 4: fft_pass<4, INV>(z, tws, T, n, inv_m, mask); break;
        default: fft_pass<2, INV>(z, tws, T, n, inv_m, mask); break;
    }
}

__global__ void __launch_bounds__(FT) bandpass_fft_kernel(const __grid_constant__ FftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* z = reinterpret_cast<float2*>(smem_raw);                  // [FG][T]
    float2* tws = z + (size_t)FG * a.T;                               // [T]
    const int T = a.T;
    const long long p0 = (long long)blockIdx.x * (2 * FG);
    for (int m = threadIdx.x; m < T; m += FT) tws[m] = a.tw[m];
    // tile load: (px 2s, px 2s+1) of time step t -> z[s][t]; the series' first sample is
    // subtracted (a constant only feeds the dropped DC bin; keeps float32 error at signal scale)
    {
        const int s = threadIdx.x % FG;                 // FT % FG == 0: fixed series per thread
        const long long pa = p0 + 2 * s;
        const bool va = pa < a.P, vb = pa + 1 < a.P;
        float2 x0 = make_float2(0.f, 0.f);
        if (va) x0.x = __ldg(a.in + pa);
        if (vb) x0.y = __ldg(a.in + pa + 1);
        // LB time steps per thread in flight at once: the tile load is pure latency otherwise
        // (ncu: 36 % of the kernel's stall samples sat on the first use of each loaded value)
        constexpr int LB = 8;
        for (int idx0 = threadIdx.x; idx0 < T * FG; idx0 += FT * LB) {
            float2 v[LB];
#pragma unroll
            for (int u = 0; u < LB; ++u) {
                const int idx = idx0 + u * FT;
                v[u] = make_float2(0.f, 0.f);
                if (idx < T * FG) {
                    const float* src = a.in + (size_t)(idx / FG) * a.P + pa;
                    if (a.pair_ok && vb) v[u] = __ldg(reinterpret_cast<const float2*>(src));
                    else { if (va) v[u].x = __ldg(src); if (vb) v[u].y = __ldg(src + 1); }
                }
            }
#pragma unroll
            for (int u = 0; u < LB; ++u) {
                const int idx = idx0 + u * FT;
                if (idx < T * FG) z[(size_t)s * T + idx / FG] = make_float2(v[u].x - x0.x, v[u].y - x0.y);
            }
        }
    }
    __syncthreads();
    for (int ps = 0; ps < a.npass; ++ps) {
        fft_dispatch<false>(z, tws, T, a.radix[ps], a.sub[ps], a.inv_m[ps], ps == a.npass - 1 ? a.mask : nullptr);
        __syncthreads();
    }
    for (int ps = a.npass - 1; ps >= 0; --ps) {
        fft_dispatch<true>(z, tws, T, a.radix[ps], a.sub[ps], a.inv_m[ps], nullptr);
        __syncthreads();
    }
    {
        const int s = threadIdx.x % FG;
        const long long pa = p0 + 2 * s;
        const bool va = pa < a.P, vb = pa + 1 < a.P;
        for (int idx = threadIdx.x; idx < T * FG; idx += FT) {
            const int t = idx / FG;
            const float2 v = z[(size_t)s * T + t];
            float* dst = a.out + (size_t)t * a.P + pa;
            if (a.pair_ok && vb) *reinterpret_cast<float2*>(dst) = v;
            else { if (va) dst[0] = v.x; if (vb) dst[1] = v.y; }
        }
    }
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

static int temporal_bandpass_impl(vhr_ctx* ctx, const float* d_in, float* d_out, int T, int64_t P, double fps,
                                  double f_lo, double f_hi, float gain, cudaStream_t stream);

extern "C" int vhr_temporal_bandpass(vhr_ctx* ctx, const float* d_in, float* d_out, int T, int64_t P, double fps,
                                     double f_lo, double f_hi, float gain, void* stream_) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_in && d_out, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && P >= 1 && fps > 0, "bad shape");
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = vhr_enter(ctx, stream);           // the twiddle table and the band mask are context-owned
    if (rc != VHR_OK) return rc;
    return vhr_leave(ctx, stream, temporal_bandpass_impl(ctx, d_in, d_out, T, P, fps, f_lo, f_hi, gain, stream));
}

static int temporal_bandpass_impl(vhr_ctx* ctx, const float* d_in, float* d_out, int T, int64_t P, double fps,
                                  double f_lo, double f_hi, float gain, cudaStream_t stream) {
    int k0 = -1, k1 = -1;
    int nb = vhr_band_bins(T, fps, f_lo, f_hi, &k0, &k1);
    if (nb <= 0) {   // empty band: the filter output is identically zero
        zero_fill_kernel<<<ctx->num_sms * 4, 256, 0, stream>>>(d_out, (long long)T * P);
        return vhr_after_launch(ctx, "zero_fill_kernel");
    }
    int rc = ensure_twiddles(ctx, T, stream);
    if (rc != VHR_OK) return rc;
    {   // form (1): mixed-radix FFT when T factors into 2, 3, 5 and the tile fits in shared memory
        const char* force = getenv("VHR_BANDPASS_DFT");          // test hook: exercise form (2)
        std::vector<int> radices;
        int n = T;
        // composite radices first (9 = 3x3, 8 = 4x2 in registers): fewer shared-memory passes.  (25 = 5x5 was
        // tried: 128 registers, 2 CTAs/SM and T/25 butterflies per series made the kernel 40 % slower.)
        for (int f : {9, 8, 5, 3}) while (n % f == 0) { radices.push_back(f); n /= f; }
        while (n % 4 == 0) { radices.push_back(4); n /= 4; }
        while (n % 2 == 0) { radices.push_back(2); n /= 2; }
        {   // pass order, measured at T = 1800 (profiles/README.md): small radices first, 8 then 9 last (a last
            // pass of radix 8 reads 64-byte runs per thread: 16-way bank conflicts, 0.40 ms vs 0.29 ms)
            std::vector<int> ordered;
            for (int f : {5, 3, 4, 2, 8, 9})
                for (int r : radices) if (r == f) ordered.push_back(r);
            radices = ordered;
        }
        if (const char* ord = getenv("VHR_BANDPASS_RADICES")) {   // tuning / test hook: "5,5,9,8" (must multiply to T)
            std::vector<int> r2;
            long long prod = 1;
            for (const char* c = ord; *c;) {
                const int v = atoi(c);
                if (v == 2 || v == 3 || v == 4 || v == 5 || v == 8 || v == 9) { r2.push_back(v); prod *= v; }
                while (*c && *c != ',') ++c;
                if (*c == ',') ++c;
            }
            if (prod == T && n == 1) radices = r2;
        }
        const size_t smem_fft = (size_t)T * 8 * (FG + 1);
        if (n == 1 && T >= 2 && (int)radices.size() <= MAXPASS && (long long)smem_fft <= ctx->smem_optin && T <= 65535 &&
            !(force && force[0] == '1')) {
            FftArgs fa;
            memset(&fa, 0, sizeof(fa));
            fa.in = d_in; fa.out = d_out; fa.tw = ctx->tw; fa.T = T; fa.P = P;
            fa.npass = (int)radices.size();
            int sub = T;
            for (int i = 0; i < fa.npass; ++i) {
                fa.radix[i] = radices[i];
                fa.sub[i] = sub;
                const unsigned mm = (unsigned)(sub / radices[i]);
                fa.inv_m[i] = mm > 1 ? 0xFFFFFFFFu / mm + 1u : 0u;       // exact for b < 65536
                sub /= radices[i];
            }
            fa.pair_ok = (P % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 7) == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 7) == 0);
            // mask in digit-reversed (DIF output) order; cached per (T, band, gain, pass order)
            unsigned long long rkey = 1;
            for (int r : radices) rkey = rkey * 16ull + (unsigned long long)r;
            if (!(ctx->mask && ctx->mask_T == T && ctx->mask_k0 == k0 && ctx->mask_k1 == k1 && ctx->mask_gain == gain &&
                  ctx->mask_radix_key == rkey)) {
                std::vector<float> hm((size_t)T);
                const float g = (float)((double)gain / (double)T);
                for (int pos = 0; pos < T; ++pos) {
                    int rem = pos, sb = T, k = 0, mul = 1;
                    for (int r : radices) { sb /= r; const int q = rem / sb; rem -= q * sb; k += mul * q; mul *= r; }
                    const int kk = k <= T - k ? k : T - k;          // bins k and T-k share |f|
                    hm[pos] = (kk >= k0 && kk <= k1) ? g : 0.0f;
                }
                if (ctx->mask && ctx->mask_T != T) {
                    VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
                    VHR_CHECK_CUDA(ctx, cudaFree(ctx->mask));
                    ctx->mask = nullptr;
                }
                if (!ctx->mask) VHR_CHECK_CUDA(ctx, cudaMalloc(&ctx->mask, sizeof(float) * (size_t)T));
                VHR_CHECK_CUDA(ctx, cudaStreamSynchronize(stream));      // previous users of the cached mask
                VHR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->mask, hm.data(), sizeof(float) * (size_t)T, cudaMemcpyHostToDevice, stream));
                VHR_CHECK_CUDA(ctx, cudaStreamSynchronize(stream));
                ctx->mask_T = T; ctx->mask_k0 = k0; ctx->mask_k1 = k1; ctx->mask_gain = gain; ctx->mask_radix_key = rkey;
            }
            fa.mask = ctx->mask;
            VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(bandpass_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft));
            const long long blocks = (P + 2 * FG - 1) / (2 * FG);
            bandpass_fft_kernel<<<(unsigned)blocks, FT, smem_fft, stream>>>(fa);
            return vhr_after_launch(ctx, "bandpass_fft_kernel");
        }
    }
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
