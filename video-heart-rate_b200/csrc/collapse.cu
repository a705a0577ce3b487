// EVM reconstruction: pyrUp^L collapse of the filtered, amplified level + add-back to the
// original uint8 frame, with the rectangle-ROI mean reduction fused into the same pass.
//
// No reference code exists for the collapse; spec = cv2.pyrUp on
// float32 (oracle/evm.py:pyrup): per axis even outputs (s[i-1] + 6 s[i] + s[i+1]) / 8, odd
// outputs (s[i] + s[i+1]) / 2, border low side reflect-101, high side replicate, sizes
// walking back the pyrDown chain.  The fused ROI mean replaces get_avg over the cheek slice
// (rppg_VIDEO.py:60-66,106-110) evaluated on the magnified frame.
//
// Design (DESIGN.md "collapse"): this kernel moves 15 of the 18 bytes per pixel of the whole
// EVM path (3 B/px read, 12 B/px written) and was instruction-bound in its first form, so it
// is organised to minimise instructions per output value:
//   * CTA = one TW x 32 pixel tile of one frame.  The pyrUp halo is one sample per level, so
//     the whole chain for a tile (level-L region of a few samples up to a level-1 region of
//     (TW/2+6) x 20) is rebuilt in shared memory; levels 1..L-1 never touch HBM.  Levels are
//     stored PLANAR (channel planes) and each expansion step maps one thread to one source
//     cell producing its 2x2 destination block (27 loads -> 12 values).
//   * Last expansion: one work item = 4 pixels x 2 rows (24 values): 18 LDS.64 from the
//     level-1 planes, vertical then horizontal interpolation in registers, uint8 -> float by
//     byte-permute + one add, add-back fused into the last FMA, three 16-byte stores per row.
//     No sliding window: items are independent, registers stay low, occupancy high.
//   * Frame-border fix-ups (reflect/replicate) and the ROI accumulation are compiled as
//     separate loop bodies selected by block-uniform flags, so interior tiles pay for neither.
//   * ROI sums: per-thread float accumulators -> warp shuffle -> per-tile partial (double)
//     -> fixed-order finalize kernel.  No atomics: results are run-to-run identical.
#include "common.cuh"

namespace {

constexpr int TH = 32;                  // tile height (rows); even
constexpr int MAXT = 320;               // max threads per CTA
constexpr int KMAXF = 4;                // fused ROI rectangles per call

struct LevelGeom {            // storage geometry of one level's region in shared memory
    int ro, co;               // image row / col stored at index 0
    int rs, ps;               // row stride, plane stride (floats)
    int off;                  // float offset of plane 0 in dynamic smem
};

struct ColArgs {
    const float* lvl;
    const uint8_t* frames;
    float* out_f32;
    uint8_t* out_u8;
    int T, H, W, L;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    int TW;                   // tile width in pixels (multiple of 128)
    int tiles_x, tiles_y;
    const int32_t* rects;     // (T,K,4)
    int K;
    double* partial;          // (T, tiles_y*tiles_x, K, 3)
    int buf_odd_off, buf_even_off;   // float offsets into dynamic smem
    int vec_ok;               // W % 4 == 0 and bases aligned
};

// border-mapped neighbours of source index p on an axis of length n (cv2.pyrUp rule)
__device__ __forceinline__ int nb_lo(int p, int n) { return (p - 1 < 0) ? (n > 1 ? 1 : 0) : p - 1; }
__device__ __forceinline__ int nb_hi(int p, int n) { return (p + 1 >= n) ? n - 1 : p + 1; }

// one expansion step in shared memory: every source cell (p,i) of level l+1 produces the 2x2
// destination block (2p..2p+1, 2i..2i+1) of level l, all three channel planes.
__device__ __forceinline__ void expand_level(const float* __restrict__ smf, const LevelGeom& s, int sn_h, int sn_w,
                                             float* __restrict__ smw, const LevelGeom& d, int p_lo, int p_hi,
                                             int i_lo, int i_hi, int tid, int nthreads) {
    const int ni = i_hi - i_lo + 1;
    const int ncell = (p_hi - p_lo + 1) * ni;
    for (int cell = tid; cell < ncell; cell += nthreads) {
        const int pr = cell / ni;
        const int p = p_lo + pr, i = i_lo + (cell - pr * ni);
        const int r0 = (nb_lo(p, sn_h) - s.ro) * s.rs, r1 = (min(p, sn_h - 1) - s.ro) * s.rs, r2 = (nb_hi(p, sn_h) - s.ro) * s.rs;
        const int c0 = nb_lo(i, sn_w) - s.co, c1 = min(i, sn_w - 1) - s.co, c2 = nb_hi(i, sn_w) - s.co;
        const int dbase = (2 * p - d.ro) * d.rs + (2 * i - d.co);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* sp = smf + s.off + ch * s.ps;
            float he[3], ho[3];
            const int rr[3] = {r0, r1, r2};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float x0 = sp[rr[q] + c0], x1 = sp[rr[q] + c1], x2 = sp[rr[q] + c2];
                he[q] = fmaf(x0 + x2, 0.125f, x1 * 0.75f);
                ho[q] = (x1 + x2) * 0.5f;
            }
            float* dp = smw + d.off + ch * d.ps + dbase;
            dp[0] = fmaf(he[0] + he[2], 0.125f, he[1] * 0.75f);
            dp[1] = fmaf(ho[0] + ho[2], 0.125f, ho[1] * 0.75f);
            dp[d.rs] = (he[1] + he[2]) * 0.5f;
            dp[d.rs + 1] = (ho[1] + ho[2]) * 0.5f;
        }
    }
}

template <int KMAX, bool F32OUT, bool U8OUT>
struct Tile {
    const ColArgs& a;
    int t, x0, x1, y0, y1;
    // level-1 geometry
    const float* l1;
    int l1_rs, l1_ps, l1_ro;
    int n1w, n1h;
    // ROI (block-uniform)
    int rx1[KMAXF > 0 ? KMAXF : 1], ry1[KMAXF > 0 ? KMAXF : 1], rx2[KMAXF > 0 ? KMAXF : 1], ry2[KMAXF > 0 ? KMAXF : 1];
    bool hit[KMAXF > 0 ? KMAXF : 1];
    float acc[KMAXF > 0 ? KMAXF : 1][3];

    __device__ Tile(const ColArgs& a_) : a(a_) {}

    // one work item: pixels X..X+3, rows 2m and 2m+1
    template <bool EDGE, bool ROI, bool VEC>
    __device__ __forceinline__ void item(int X, int m, int colofs) {
        const int rm = (nb_lo(m, n1h) - l1_ro) * l1_rs, rc = (m - l1_ro) * l1_rs, rp = (nb_hi(m, n1h) - l1_ro) * l1_rs;
        float ve[3][4], vo[3][4];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* pl = l1 + ch * l1_ps + colofs;
            const float2 a0 = *reinterpret_cast<const float2*>(pl + rm), a1 = *reinterpret_cast<const float2*>(pl + rm + 2);
            const float2 b0 = *reinterpret_cast<const float2*>(pl + rc), b1 = *reinterpret_cast<const float2*>(pl + rc + 2);
            const float2 c0 = *reinterpret_cast<const float2*>(pl + rp), c1 = *reinterpret_cast<const float2*>(pl + rp + 2);
            const float av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y}, cv[4] = {c0.x, c0.y, c1.x, c1.y};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ve[ch][j] = fmaf(av[j] + cv[j], 0.125f, bv[j] * 0.75f);
                vo[ch][j] = (bv[j] + cv[j]) * 0.5f;
            }
            if (EDGE) {
                const int J = X >> 1;
                if (J + 1 >= n1w) { ve[ch][2] = ve[ch][1]; vo[ch][2] = vo[ch][1]; }           // replicate high side
                if (J + 2 >= n1w) { ve[ch][3] = ve[ch][2]; vo[ch][3] = vo[ch][2]; }
                if (J - 1 < 0) { ve[ch][0] = ve[ch][2]; vo[ch][0] = vo[ch][2]; }              // reflect-101: col -1 -> col 1
            }
        }
        const size_t row_elems = (size_t)a.W * 3;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int y = 2 * m + half;
            if (y < y0 || y >= y1) continue;          // tile rows start even; only the frame's last odd row can be cut
            const size_t g = ((size_t)t * a.H + y) * row_elems + (size_t)X * 3;
            uint32_t w0, w1, w2;
            if (VEC) {
                const uint32_t* fp = reinterpret_cast<const uint32_t*>(a.frames + g);
                w0 = __ldg(fp); w1 = __ldg(fp + 1); w2 = __ldg(fp + 2);
            } else {
                w0 = w1 = w2 = 0;
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    if (X + k / 3 < x1) {
                        const uint32_t bv = __ldg(a.frames + g + k);
                        if (k < 4) w0 |= bv << (8 * k); else if (k < 8) w1 |= bv << (8 * (k - 4)); else w2 |= bv << (8 * (k - 8));
                    }
                }
            }
            const uint32_t ww[3] = {w0, w1, w2};
            float o[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int px = k / 3, ch = k - 3 * px;
                // uint8 -> float: 0x4B0000xx = 2^23 + xx
                const float f = __uint_as_float(__byte_perm(ww[k >> 2], 0x4B000000u, 0x7440 + (k & 3))) - 8388608.0f;
                const float (&v)[4] = half ? vo[ch] : ve[ch];
                float up;
                if (px == 0) up = fmaf(v[0] + v[2], 0.125f, fmaf(v[1], 0.75f, f));
                else if (px == 1) up = fmaf(v[1] + v[2], 0.5f, f);
                else if (px == 2) up = fmaf(v[1] + v[3], 0.125f, fmaf(v[2], 0.75f, f));
                else up = fmaf(v[2] + v[3], 0.5f, f);
                o[k] = up;
            }
            if (F32OUT) {
                if (VEC) {
                    float4* op = reinterpret_cast<float4*>(a.out_f32 + g);
                    op[0] = make_float4(o[0], o[1], o[2], o[3]);
                    op[1] = make_float4(o[4], o[5], o[6], o[7]);
                    op[2] = make_float4(o[8], o[9], o[10], o[11]);
                } else {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (X + k / 3 < x1) a.out_f32[g + k] = o[k];
                }
            }
            if (U8OUT) {
                uint32_t q[3] = {0, 0, 0};
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const float v = fminf(fmaxf(o[k], 0.0f), 255.0f);
                    q[k >> 2] |= (uint32_t)(int)(v + 0.5f) << (8 * (k & 3));
                }
                if (VEC) {
                    uint32_t* op = reinterpret_cast<uint32_t*>(a.out_u8 + g);
                    op[0] = q[0]; op[1] = q[1]; op[2] = q[2];
                } else {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (X + k / 3 < x1) a.out_u8[g + k] = (uint8_t)(q[k >> 2] >> (8 * (k & 3)));
                }
            }
            if (ROI && KMAX > 0) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (hit[k] && y >= ry1[k] && y < ry2[k]) {
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            if (X + px >= rx1[k] && X + px < rx2[k] && X + px < x1) {
                                acc[k][0] += o[3 * px]; acc[k][1] += o[3 * px + 1]; acc[k][2] += o[3 * px + 2];
                            }
                        }
                    }
                }
            }
        }
    }

    template <bool EDGE, bool ROI, bool VEC>
    __device__ __forceinline__ void run(int colofs_base) {
        const int cx = threadIdx.x, X = x0 + 4 * cx;
        if (X >= x1) return;
        const int colofs = colofs_base + 2 * cx;
        for (int m = (y0 >> 1) + threadIdx.y; 2 * m < y1; m += blockDim.y) item<EDGE, ROI, VEC>(X, m, colofs);
    }
};

template <int KMAX, bool F32OUT, bool U8OUT>
__global__ void __launch_bounds__(MAXT, 3) collapse_kernel(const ColArgs a) {
    extern __shared__ __align__(16) float smf[];
    __shared__ float red[MAXT / 32][KMAXF][3];

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int tiles = a.tiles_x * a.tiles_y;
    const int t = blockIdx.x / tiles;
    const int tile = blockIdx.x - t * tiles;
    const int by = tile / a.tiles_x, bx = tile - by * a.tiles_x;
    const int x0 = bx * a.TW, x1 = min(a.W, x0 + a.TW);
    const int y0 = by * TH, y1 = min(a.H, y0 + TH);
    const int L = a.L;

    // regions [ra,rb] x [ca,cb] needed at each level (inclusive) and their smem storage
    int ra[VHR_MAX_LEVELS + 1], rb[VHR_MAX_LEVELS + 1], ca[VHR_MAX_LEVELS + 1], cb[VHR_MAX_LEVELS + 1];
    LevelGeom g[VHR_MAX_LEVELS + 1];
    ra[0] = y0; rb[0] = y1 - 1; ca[0] = x0; cb[0] = x0 + a.TW - 1;   // full tile width: partial tiles still index within it
    for (int l = 1; l <= L; ++l) {
        ra[l] = max(0, (ra[l - 1] >> 1) - 1);
        rb[l] = min(a.h[l] - 1, (rb[l - 1] >> 1) + 1);
        ca[l] = max(0, (ca[l - 1] >> 1) - 1);
        cb[l] = min(a.w[l] - 1, (cb[l - 1] >> 1) + 1);
        if (cb[l] < ca[l]) cb[l] = ca[l];            // partial tile far right of a tiny level
        LevelGeom& q = g[l];
        q.ro = ra[l] & ~1;                            // 2x2 blocks start on even rows / cols
        q.co = (ca[l] & ~1) - 1;                      // odd origin: the pairs (J-1,J) read by the last stage are 8-byte aligned
        const int nrows = (rb[l] | 1) - q.ro + 1;
        q.rs = (((cb[l] | 1) + 3 - q.co + 1) + 1) & ~1;       // +3: slack columns read (and discarded) at the frame edge
        q.ps = nrows * q.rs;
        q.off = (l & 1) ? a.buf_odd_off : a.buf_even_off;
    }

    // ---- level L region from HBM (interleaved) into planar smem ---------------------------
    {
        const LevelGeom& q = g[L];
        const int rw = (cb[L] - ca[L] + 1) * 3, rh = rb[L] - ra[L] + 1;
        const float* src = a.lvl + ((size_t)t * a.h[L] * a.w[L]) * 3;
        for (int idx = tid; idx < rw * rh; idx += nthreads) {
            const int r = idx / rw, j = idx - r * rw;
            const int c = j / 3, ch = j - 3 * c;
            smf[q.off + ch * q.ps + (ra[L] + r - q.ro) * q.rs + (ca[L] + c - q.co)] =
                __ldg(src + ((size_t)(ra[L] + r) * a.w[L] + ca[L]) * 3 + j);
        }
    }
    // ---- levels L-1 .. 1 in shared memory ----------------------------------------------------
    for (int l = L - 1; l >= 1; --l) {
        __syncthreads();
        expand_level(smf, g[l + 1], a.h[l + 1], a.w[l + 1], smf, g[l], g[l].ro >> 1, rb[l] >> 1, (g[l].co + 1) >> 1,
                     cb[l] >> 1, tid, nthreads);
    }
    __syncthreads();

    // ---- last expansion fused with add-back, store and ROI sums -------------------------------
    Tile<KMAX, F32OUT, U8OUT> tl(a);
    tl.t = t; tl.x0 = x0; tl.x1 = x1; tl.y0 = y0; tl.y1 = y1;
    tl.l1 = smf + g[1].off; tl.l1_rs = g[1].rs; tl.l1_ps = g[1].ps; tl.l1_ro = g[1].ro;
    tl.n1w = a.w[1]; tl.n1h = a.h[1];
    bool any_hit = false;
    if (KMAX > 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            tl.hit[k] = false;
            tl.acc[k][0] = tl.acc[k][1] = tl.acc[k][2] = 0.f;
            if (k < a.K) {
                const int32_t* rc = a.rects + ((size_t)t * a.K + k) * 4;
                tl.rx1[k] = rc[0]; tl.ry1[k] = rc[1]; tl.rx2[k] = rc[2]; tl.ry2[k] = rc[3];
                tl.hit[k] = rc[0] < x1 && rc[2] > x0 && rc[1] < y1 && rc[3] > y0 && rc[2] > rc[0] && rc[3] > rc[1];
                any_hit |= tl.hit[k];
            }
        }
    }
    // storage index of level-1 column (x0/2 - 1): the first thread's pair (J-1, J)
    const int colofs_base = ((x0 >> 1) - 1) - g[1].co;
    const bool edge = (x0 == 0) || ((x0 + a.TW) >> 1) + 1 >= a.w[1];
    const bool vec = a.vec_ok && (x0 + a.TW <= a.W);
    if (vec) {
        if (any_hit) { if (edge) tl.template run<true, true, true>(colofs_base); else tl.template run<false, true, true>(colofs_base); }
        else { if (edge) tl.template run<true, false, true>(colofs_base); else tl.template run<false, false, true>(colofs_base); }
    } else {
        if (any_hit) tl.template run<true, true, false>(colofs_base); else tl.template run<true, false, false>(colofs_base);
    }

    if (KMAX > 0 && any_hit) {       // block-uniform
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = tl.acc[k][c];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == 0) red[wid][k][c] = v;
            }
        }
        __syncthreads();
        if (tid < a.K * 3) {
            const int k = tid / 3, c = tid - 3 * k;
            double s = 0.0;
            for (int w = 0; w < (nthreads + 31) / 32; ++w) s += (double)red[w][k][c];
            a.partial[(((size_t)t * tiles + tile) * a.K + k) * 3 + c] = s;
        }
    }
}

// fixed-order reduction of the per-tile partials over the tiles a rectangle touches
__global__ void roi_finalize_kernel(const double* __restrict__ partial, const int32_t* __restrict__ rects,
                                    int T, int K, int TW, int tiles_x, int tiles_y, double* __restrict__ mean) {
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
int launch_collapse(vhr_ctx* ctx, const ColArgs& a, dim3 block, size_t smem, cudaStream_t stream) {
    const unsigned grid = (unsigned)((size_t)a.T * a.tiles_x * a.tiles_y);
#define VHR_COLLAPSE_LAUNCH(F, U)                                                                             \
    do {                                                                                                      \
        auto kern = collapse_kernel<KMAX, F, U>;                                                              \
        VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
        kern<<<grid, block, smem, stream>>>(a);                                                               \
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
    // tile width: the largest of 640/512/384/256/128 that divides W, else 256 (last tile partial)
    int TW = 256;
    const int cand[5] = {640, 512, 384, 256, 128};
    for (int i = 0; i < 5; ++i)
        if (W % cand[i] == 0) { TW = cand[i]; break; }
    if (W < 256) TW = ((W + 127) / 128) * 128;
    a.TW = TW;
    a.tiles_x = (W + TW - 1) / TW;
    a.tiles_y = (H + TH - 1) / TH;
    const int cg = TW / 4;                              // thread columns (multiple of 32)
    int rg = MAXT / cg;                                 // row-pair groups
    if (rg < 1) rg = 1;
    if (rg > TH / 2) rg = TH / 2;
    dim3 block(cg, rg);
    a.rects = d_rects; a.K = K;
    // shared-memory budget: odd levels share one buffer, even levels the other (planar x3)
    size_t odd = 0, even = 0;
    {
        int rh = TH, rw = TW;
        for (int l = 1; l <= levels; ++l) {
            rh = rh / 2 + 5;
            rw = rw / 2 + 10;
            const size_t n = (size_t)3 * rh * rw;
            if (l & 1) odd = odd > n ? odd : n; else even = even > n ? even : n;
        }
    }
    a.buf_odd_off = 0;
    a.buf_even_off = (int)((odd + 3) & ~(size_t)3);
    const size_t smem = ((size_t)a.buf_even_off + even) * sizeof(float);
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "collapse: tile needs %zu bytes of shared memory (> %d)", smem, ctx->smem_optin);
        return VHR_ERR_UNSUPPORTED;
    }
    a.vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 3) == 0) &&
               (!d_out_f32 || (reinterpret_cast<uintptr_t>(d_out_f32) & 15) == 0) &&
               (!d_out_u8 || (reinterpret_cast<uintptr_t>(d_out_u8) & 3) == 0);
    if (K > 0) {
        void* p = nullptr;
        int rc = vhr_scratch(ctx, sizeof(double) * (size_t)T * a.tiles_x * a.tiles_y * K * 3, &p);
        if (rc != VHR_OK) return rc;
        a.partial = reinterpret_cast<double*>(p);
    }
    int rc = (K > 0) ? launch_collapse<KMAXF>(ctx, a, block, smem, stream) : launch_collapse<0>(ctx, a, block, smem, stream);
    if (rc != VHR_OK) return rc;
    if (K > 0) {
        const int n = T * K * 3;
        roi_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a.partial, d_rects, T, K, TW, a.tiles_x, a.tiles_y, d_roi_mean);
        rc = vhr_after_launch(ctx, "roi_finalize_kernel");
    }
    return rc;
}
