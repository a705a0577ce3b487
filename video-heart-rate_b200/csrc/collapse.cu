// EVM reconstruction: pyrUp^L collapse of the filtered, amplified level + add-back to the
// original uint8 frame, with the rectangle-ROI mean reduction fused into the same pass.
//
// No reference code exists for the collapse; spec = cv2.pyrUp on
// float32 (oracle/evm.py:pyrup): per axis even outputs (s[i-1] + 6 s[i] + s[i+1]) / 8, odd
// outputs (s[i] + s[i+1]) / 2, border low side reflect-101, high side replicate, sizes
// walking back the pyrDown chain.  The fused ROI mean replaces get_avg over the cheek slice
// (rppg_VIDEO.py:60-66,106-110) evaluated on the magnified frame.
//
// Design (DESIGN.md "collapse"): this kernel moves 15 of the 18 bytes per pixel of the
// whole EVM path (3 B/px read, 12 B/px written), so it is laid out around the output store:
//   * CTA = one 320 x 32 pixel tile of one frame; thread = 4 consecutive floats of a row
//     (one 16-byte store, one 4-byte load of the original pixels), marching down the rows.
//   * The pyrUp halo is one sample per level, so the whole chain for a tile (level L region
//     of ~23x5 samples up to a level-1 region of ~163x19) is rebuilt in shared memory per
//     tile; levels 1..L-1 never touch HBM.
//   * The last expansion is evaluated on the fly: horizontally (3 taps from the level-1
//     tile in smem, once per level-1 row) and vertically (3-row register window), fused
//     with the uint8 -> float conversion, the add-back and the store.
//   * ROI sums: per-thread float accumulators -> warp shuffle -> per-tile partial (double)
//     -> fixed-order finalize kernel.  No atomics: results are run-to-run identical.
#include "common.cuh"

namespace {

constexpr int TW = 320;                 // tile width in pixels (960 floats = 4 per thread)
constexpr int TH = 32;                  // tile height (even)
constexpr int NT = TW * 3 / 4;          // 240 threads own columns
constexpr int NTL = 256;                // launched threads (whole warps for the shuffles)
constexpr int KMAXF = 4;                // fused ROI rectangles per call

struct ColArgs {
    const float* lvl;
    const uint8_t* frames;
    float* out_f32;
    uint8_t* out_u8;
    int T, H, W, L;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    int tiles_x, tiles_y;
    const int32_t* rects;     // (T,K,4)
    int K;
    double* partial;          // (T, tiles_y*tiles_x, K, 3)
    int buf_odd_off, buf_even_off;   // float offsets into dynamic smem
    int vec_ok;               // W % 4 == 0 and bases aligned
};

// one axis of pyrUp: destination index d -> up to 3 (source index, weight) taps
struct Tap3 {
    int i0, i1, i2;
    float w0, w1, w2;
};
__device__ __forceinline__ Tap3 up_taps(int d, int n) {
    Tap3 t;
    const int m = d >> 1;
    const int mp = (m + 1 >= n) ? n - 1 : m + 1;                 // replicate high side
    if ((d & 1) == 0) {
        const int mm = (m - 1 < 0) ? (n > 1 ? 1 : 0) : m - 1;    // reflect-101 low side
        t.i0 = mm; t.i1 = m; t.i2 = mp;
        t.w0 = 0.125f; t.w1 = 0.75f; t.w2 = 0.125f;
    } else {
        t.i0 = m; t.i1 = mp; t.i2 = mp;
        t.w0 = 0.5f; t.w1 = 0.5f; t.w2 = 0.0f;
    }
    return t;
}

template <int KMAX, bool F32OUT, bool U8OUT>
__global__ void __launch_bounds__(NTL) collapse_kernel(const ColArgs a) {
    extern __shared__ __align__(16) float smf[];
    __shared__ float red[NTL / 32][KMAXF][3];

    const int tiles = a.tiles_x * a.tiles_y;
    const int t = blockIdx.x / tiles;
    const int tile = blockIdx.x - t * tiles;
    const int by = tile / a.tiles_x, bx = tile - by * a.tiles_x;
    const int x0 = bx * TW, x1 = min(a.W, x0 + TW);
    const int y0 = by * TH, y1 = min(a.H, y0 + TH);
    const int L = a.L;

    // regions [ra,rb] x [ca,cb] needed at each level (inclusive)
    int ra[VHR_MAX_LEVELS + 1], rb[VHR_MAX_LEVELS + 1], ca[VHR_MAX_LEVELS + 1], cb[VHR_MAX_LEVELS + 1];
    ra[0] = y0; rb[0] = y1 - 1; ca[0] = x0; cb[0] = x1 - 1;
    for (int l = 1; l <= L; ++l) {
        ra[l] = max(0, (ra[l - 1] >> 1) - 1);
        rb[l] = min(a.h[l] - 1, (rb[l - 1] >> 1) + 1);
        ca[l] = max(0, (ca[l - 1] >> 1) - 1);
        cb[l] = min(a.w[l] - 1, (cb[l - 1] >> 1) + 1);
    }

    // ---- level L region from HBM -------------------------------------------------------
    {
        float* dst = smf + ((L & 1) ? a.buf_odd_off : a.buf_even_off);
        const int rw = (cb[L] - ca[L] + 1) * 3, rh = rb[L] - ra[L] + 1;
        const float* src = a.lvl + ((size_t)t * a.h[L] * a.w[L]) * 3;
        for (int idx = threadIdx.x; idx < rw * rh; idx += NTL) {
            int r = idx / rw, j = idx - r * rw;
            dst[idx] = __ldg(src + ((size_t)(ra[L] + r) * a.w[L] + ca[L]) * 3 + j);
        }
    }
    // ---- levels L-1 .. 1 in shared memory ----------------------------------------------
    for (int l = L - 1; l >= 1; --l) {
        __syncthreads();
        const float* src = smf + (((l + 1) & 1) ? a.buf_odd_off : a.buf_even_off);
        float* dst = smf + ((l & 1) ? a.buf_odd_off : a.buf_even_off);
        const int sw = (cb[l + 1] - ca[l + 1] + 1) * 3;
        const int dwp = cb[l] - ca[l] + 1, dh = rb[l] - ra[l] + 1;
        const int nsrc_h = a.h[l + 1], nsrc_w = a.w[l + 1];
        for (int idx = threadIdx.x; idx < dwp * dh; idx += NTL) {
            const int r = idx / dwp, xq = idx - r * dwp;
            const Tap3 tv = up_taps(ra[l] + r, nsrc_h);
            const Tap3 th = up_taps(ca[l] + xq, nsrc_w);
            const float* r0 = src + (tv.i0 - ra[l + 1]) * sw;
            const float* r1 = src + (tv.i1 - ra[l + 1]) * sw;
            const float* r2 = src + (tv.i2 - ra[l + 1]) * sw;
            const int c0 = (th.i0 - ca[l + 1]) * 3, c1 = (th.i1 - ca[l + 1]) * 3, c2 = (th.i2 - ca[l + 1]) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v0 = tv.w0 * r0[c0 + c] + tv.w1 * r1[c0 + c] + tv.w2 * r2[c0 + c];
                float v1 = tv.w0 * r0[c1 + c] + tv.w1 * r1[c1 + c] + tv.w2 * r2[c1 + c];
                float v2 = tv.w0 * r0[c2 + c] + tv.w1 * r1[c2 + c] + tv.w2 * r2[c2 + c];
                dst[idx * 3 + c] = th.w0 * v0 + th.w1 * v1 + th.w2 * v2;
            }
        }
    }
    __syncthreads();

    // ---- last expansion fused with add-back, store and ROI sums ---------------------------
    // L >= 1: source = level-1 region (buf_odd); n1w/n1h = level-1 size
    const float* l1 = smf + a.buf_odd_off;
    const int l1w = (cb[1] - ca[1] + 1) * 3;
    const int n1w = a.w[1], n1h = a.h[1];
    const int fbase = 4 * threadIdx.x;                 // first flat float of this thread in the tile row
    int off[4][3];
    float wg[4][3];
    int gx[4], ch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int f = fbase + j;
        const int px = f / 3;
        ch[j] = f - px * 3;
        gx[j] = x0 + px;
        const int gxc = min(gx[j], x1 - 1);          // threads past the tile edge stay in range
        const Tap3 th = up_taps(gxc, n1w);
        off[j][0] = (th.i0 - ca[1]) * 3 + ch[j];
        off[j][1] = (th.i1 - ca[1]) * 3 + ch[j];
        off[j][2] = (th.i2 - ca[1]) * 3 + ch[j];
        wg[j][0] = th.w0; wg[j][1] = th.w1; wg[j][2] = th.w2;
    }
    const bool any_valid = gx[0] < x1;
    const bool all_valid = gx[3] < x1;

    auto hrow = [&](int r, float (&v)[4]) {
        const float* rp = l1 + (r - ra[1]) * l1w;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = wg[j][0] * rp[off[j][0]] + wg[j][1] * rp[off[j][1]] + wg[j][2] * rp[off[j][2]];
    };

    // ROI bookkeeping (block-uniform)
    int rx1[KMAXF], ry1[KMAXF], rx2[KMAXF], ry2[KMAXF];
    bool hit[KMAXF];
    float acc[KMAXF][4];
    if (KMAX > 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            hit[k] = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
            if (k < a.K) {
                const int32_t* rc = a.rects + ((size_t)t * a.K + k) * 4;
                rx1[k] = rc[0]; ry1[k] = rc[1]; rx2[k] = rc[2]; ry2[k] = rc[3];
                hit[k] = rx1[k] < x1 && rx2[k] > x0 && ry1[k] < y1 && ry2[k] > y0 && rx2[k] > rx1[k] && ry2[k] > ry1[k];
            }
        }
    }

    float hm1[4], h0[4], hp1[4];
    {
        const int m0 = y0 >> 1;
        const Tap3 tv = up_taps(y0, n1h);           // y0 is even: taps (m0-1 | 1, m0, m0+1 | n-1)
        hrow(tv.i0, hm1);
        hrow(tv.i1, h0);
        hrow(tv.i2, hp1);
        (void)m0;
    }
    const size_t row_elems = (size_t)a.W * 3;
    size_t gofs = ((size_t)t * a.H + y0) * row_elems + (size_t)x0 * 3 + fbase;
    for (int y = y0; y < y1; ++y, gofs += row_elems) {
        if ((y & 1) == 0 && y != y0) {
            // advance the window: m -> m+1
            const int m = y >> 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) { hm1[j] = h0[j]; h0[j] = hp1[j]; }
            hrow((m + 1 >= n1h) ? n1h - 1 : m + 1, hp1);
        }
        if (!any_valid) continue;
        // original pixels -> float (magic-number conversion: 0x4B0000xx = 2^23 + xx)
        uint32_t pw;
        if (a.vec_ok && all_valid) {
            pw = __ldg(reinterpret_cast<const uint32_t*>(a.frames + gofs));
        } else {
            pw = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gx[j] < x1) pw |= (uint32_t)__ldg(a.frames + gofs + j) << (8 * j);
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float f = __uint_as_float(__byte_perm(pw, 0x4B000000u, 0x7440 + j)) - 8388608.0f;
            o[j] = (y & 1) ? fmaf(h0[j] + hp1[j], 0.5f, f)
                           : fmaf(h0[j], 0.75f, fmaf(hm1[j] + hp1[j], 0.125f, f));
        }
        if (F32OUT) {
            if (a.vec_ok && all_valid) {
                *reinterpret_cast<float4*>(a.out_f32 + gofs) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gx[j] < x1) a.out_f32[gofs + j] = o[j];
            }
        }
        if (U8OUT) {
            uint32_t q = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v = fminf(fmaxf(o[j], 0.0f), 255.0f);
                q |= (uint32_t)(int)(v + 0.5f) << (8 * j);
            }
            if (a.vec_ok && all_valid) {
                *reinterpret_cast<uint32_t*>(a.out_u8 + gofs) = q;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gx[j] < x1) a.out_u8[gofs + j] = (uint8_t)(q >> (8 * j));
            }
        }
        if (KMAX > 0) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (hit[k] && y >= ry1[k] && y < ry2[k]) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (gx[j] >= rx1[k] && gx[j] < rx2[k] && gx[j] < x1) acc[k][j] += o[j];
                }
            }
        }
    }

    if (KMAX > 0) {
        bool any_hit = false;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) any_hit |= hit[k];
        if (any_hit) {       // block-uniform
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[0] += (ch[j] == 0) ? acc[k][j] : 0.f;
                    s[1] += (ch[j] == 1) ? acc[k][j] : 0.f;
                    s[2] += (ch[j] == 2) ? acc[k][j] : 0.f;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float v = s[c];
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                    if (lane == 0) red[wid][k][c] = v;
                }
            }
            __syncthreads();
            if (threadIdx.x < a.K * 3) {
                const int k = threadIdx.x / 3, c = threadIdx.x - 3 * k;
                double s = 0.0;
                for (int w = 0; w < NTL / 32; ++w) s += (double)red[w][k][c];
                a.partial[(((size_t)t * tiles + tile) * a.K + k) * 3 + c] = s;
            }
        }
    }
}

// fixed-order reduction of the per-tile partials over the tiles a rectangle touches
__global__ void roi_finalize_kernel(const double* __restrict__ partial, const int32_t* __restrict__ rects,
                                    int T, int K, int tiles_x, int tiles_y, double* __restrict__ mean) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * K * 3) return;
    const int c = idx % 3, k = (idx / 3) % K, t = idx / (3 * K);
    const int32_t* rc = rects + ((size_t)t * K + k) * 4;
    const int x1 = rc[0], y1 = rc[1], x2 = rc[2], y2 = rc[3];
    if (x2 <= x1 || y2 <= y1) {
        mean[idx] = __longlong_as_double(0x7FF8000000000000ll);   // NaN, like np.mean of an empty slice
        return;
    }
    const int tiles = tiles_x * tiles_y;
    double s = 0.0;
    for (int ty = y1 / TH; ty <= (y2 - 1) / TH && ty < tiles_y; ++ty)
        for (int tx = x1 / TW; tx <= (x2 - 1) / TW && tx < tiles_x; ++tx)
            s += partial[(((size_t)t * tiles + ty * tiles_x + tx) * K + k) * 3 + c];
    mean[idx] = s / ((double)(x2 - x1) * (double)(y2 - y1));
}

template <int KMAX>
int launch_collapse(vhr_ctx* ctx, const ColArgs& a, size_t smem, cudaStream_t stream) {
    const unsigned grid = (unsigned)((size_t)a.T * a.tiles_x * a.tiles_y);
#define VHR_COLLAPSE_LAUNCH(F, U)                                                                             \
    do {                                                                                                      \
        auto kern = collapse_kernel<KMAX, F, U>;                                                              \
        VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, NTL, smem, stream>>>(a);                                                                  \
    } while (0)
    if (a.out_f32 && a.out_u8) VHR_COLLAPSE_LAUNCH(true, true);
    else if (a.out_f32) VHR_COLLAPSE_LAUNCH(true, false);
    else if (a.out_u8) VHR_COLLAPSE_LAUNCH(false, true);
    else VHR_COLLAPSE_LAUNCH(false, false);
#undef VHR_COLLAPSE_LAUNCH
    return vhr_after_launch(ctx, "collapse_kernel");
}

}  // namespace

extern "C" int vhr_collapse_addback_roi(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames, int T, int H,
                                        int W, int levels, float* d_out_f32, uint8_t* d_out_u8,
                                        const int32_t* d_rects, int K, double* d_roi_mean, void* stream_) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_level && d_frames, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, levels >= 1 && levels <= VHR_MAX_LEVELS, "levels must be 1..6");
    VHR_REQUIRE(ctx, K >= 0 && K <= KMAXF, "fused ROI count must be 0..4");
    VHR_REQUIRE(ctx, K == 0 || (d_rects && d_roi_mean), "ROI pointers missing");
    VHR_REQUIRE(ctx, d_out_f32 || d_out_u8 || K > 0, "nothing to compute");
    cudaStream_t stream = (cudaStream_t)stream_;
    ColArgs a;
    memset(&a, 0, sizeof(a));
    a.lvl = d_level; a.frames = d_frames; a.out_f32 = d_out_f32; a.out_u8 = d_out_u8;
    a.T = T; a.H = H; a.W = W; a.L = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    a.tiles_x = (W + TW - 1) / TW;
    a.tiles_y = (H + TH - 1) / TH;
    a.rects = d_rects; a.K = K;
    // shared-memory budget: odd levels share one buffer, even levels the other
    size_t odd = 0, even = 0;
    {
        int rh = TH, rw = TW;
        for (int l = 1; l <= levels; ++l) {
            rh = (rh - 1) / 2 + 4;      // rows (rb>>1)+1 - ((ra>>1)-1) + 1 of an rh-row region
            rw = (rw - 1) / 2 + 4;
            size_t n = (size_t)rh * rw * 3;
            if (l & 1) odd = odd > n ? odd : n; else even = even > n ? even : n;
        }
    }
    a.buf_odd_off = 0;
    a.buf_even_off = (int)((odd + 3) & ~(size_t)3);
    const size_t smem = ((size_t)a.buf_even_off + even) * sizeof(float);
    a.vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 3) == 0) &&
               (!d_out_f32 || (reinterpret_cast<uintptr_t>(d_out_f32) & 15) == 0) &&
               (!d_out_u8 || (reinterpret_cast<uintptr_t>(d_out_u8) & 3) == 0);
    if (K > 0) {
        void* p = nullptr;
        int rc = vhr_scratch(ctx, sizeof(double) * (size_t)T * a.tiles_x * a.tiles_y * K * 3, &p);
        if (rc != VHR_OK) return rc;
        a.partial = reinterpret_cast<double*>(p);
    }
    int rc = (K > 0) ? launch_collapse<KMAXF>(ctx, a, smem, stream) : launch_collapse<0>(ctx, a, smem, stream);
    if (rc != VHR_OK) return rc;
    if (K > 0) {
        const int n = T * K * 3;
        roi_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a.partial, d_rects, T, K, a.tiles_x, a.tiles_y, d_roi_mean);
        rc = vhr_after_launch(ctx, "roi_finalize_kernel");
    }
    return rc;
}
