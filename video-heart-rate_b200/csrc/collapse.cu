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
// EVM path (3 B/px read, 12 B/px written).  Its first two forms were instruction-bound, so
// everything here is organised to minimise instructions per output value and to make every
// global store a full, contiguous 512-byte warp transaction:
//   * CTA = one TW x TH pixel tile of one frame.  The pyrUp halo is one sample per level, so
//     the chain for a tile (a few level-L samples up to a (TW/2+2) x (TH/2+2) level-1 region)
//     is rebuilt in shared memory; levels 1..L-1 never touch HBM.
//   * Regions are stored PLANAR (one plane per channel) with a one-sample APRON that holds
//     the border-mapped neighbours (reflect-101 low / replicate high), written explicitly
//     after each level is built.  Every stencil after that is a plain 3-tap with no index
//     mapping, and each expansion is two separable, float4-vectorised passes
//     (horizontal: 4 loads -> 4 values; vertical: 3 float4 loads -> 2 x 4 values).
//   * Last expansion: one work item = 4 pixels x 2 rows (24 values): 18 LDS.64 from the
//     level-1 planes, vertical then horizontal interpolation in registers, uint8 -> float by
//     byte-permute + one add, add-back fused into the last FMA.  Items are independent (no
//     sliding window), so registers stay low and occupancy high.
//   * Stores: a thread's 12 floats are 48 bytes, so direct float4 stores would write
//     half-sectors at a 48-byte lane stride.  Each warp instead transposes a row segment
//     (1536 B) through a private shared-memory strip and issues three fully coalesced STG.128.
//   * ROI sums: per-thread float accumulators -> warp shuffle -> per-tile partial (double)
//     -> fixed-order finalize kernel.  No atomics: results are run-to-run identical.
#include "common.cuh"
#include <stdlib.h>
#include <vector>

namespace {

constexpr int MAXT = 256;               // max threads per CTA
constexpr int KMAXF = 4;                // fused ROI rectangles per call

struct Geo {                  // storage of one level's region: planar, apron of >= 1 sample
    int c0, r0;               // image col / row stored at index 0
    int rs, ps;               // row stride, plane stride (floats; rs % 4 == 0)
    int off;                  // float offset of plane 0 in dynamic smem
};

// Level-L region + storage geometry per tile position (independent of the frame), host-built
// and served from constant memory: the gather that heads every tile's critical path needs it
// before anything else, and a constant-cache hit costs no DRAM round trip.
struct TileL {
    int ra, rb, ca, cb;
    int c0, r0, rs, ps;
    unsigned inv_nc3;         // magic reciprocal of 3 * (cb - ca + 3)
    int pad[3];
};
constexpr int MAX_CONST_TILES = 1024;
__constant__ TileL c_tileL[MAX_CONST_TILES];

struct ColArgs {
    int use_const_tiles;      // tiles_x * tiles_y <= MAX_CONST_TILES and c_tileL is valid for this launch
    const float* lvl;
    const uint8_t* frames;
    float* out_f32;
    uint8_t* out_u8;
    int T, H, W, L;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    int TW, TH;               // tile size in pixels (TW multiple of 128, TH even)
    int tiles_x, tiles_y;
    const int32_t* rects;     // (T,K,4)
    int K;
    double* partial;          // (T, tiles_y*tiles_x, K, 3)
    int bufX_off, bufY_off, bufH_off, stage_off;   // float offsets into dynamic smem
    int vec_ok;               // W % 4 == 0 and bases aligned
};

__device__ __forceinline__ int border_map(int x, int n) { return x < 0 ? (n > 1 ? 1 : 0) : (x >= n ? n - 1 : x); }

// Horizontal expansion (pass A): rows p_lo..p_hi of source level (stored in S, aprons valid)
// -> Hh rows (same row indexing as S) at destination column resolution, column groups
// g_lo..g_hi, group g = destination columns 4g-1 .. 4g+2 (stored 16-byte aligned).
__device__ __forceinline__ void expand_h(float* __restrict__ sm, const Geo S, const Geo Hh, int p_lo, int p_hi,
                                         int g_lo, int g_hi, int tid, int nthreads) {
    const int ng = g_hi - g_lo + 1;
    const int n = (p_hi - p_lo + 1) * ng;
    const unsigned inv = ng > 1 ? 0xFFFFFFFFu / (unsigned)ng + 1u : 0u;   // it / ng == umulhi(it, inv) for it < 65536; 0 encodes ng == 1
    for (int it = tid; it < n; it += nthreads) {
        const int pr = inv ? (int)__umulhi((unsigned)it, inv) : it;
        const int g = g_lo + (it - pr * ng), p = p_lo + pr;
        const int so = (p - S.r0) * S.rs + (2 * g - 1 - S.c0);
        const int ho = (p - Hh.r0) * Hh.rs + (4 * g - 1 - Hh.c0);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* sp = sm + S.off + ch * S.ps + so;
            const float x0 = sp[0], x1 = sp[1], x2 = sp[2], x3 = sp[3];
            float4 o;
            o.x = (x0 + x1) * 0.5f;                          // col 4g-1 (odd,  i = 2g-1)
            o.y = fmaf(x0 + x2, 0.125f, x1 * 0.75f);         // col 4g   (even, i = 2g)
            o.z = (x1 + x2) * 0.5f;                          // col 4g+1
            o.w = fmaf(x1 + x3, 0.125f, x2 * 0.75f);         // col 4g+2 (even, i = 2g+1)
            *reinterpret_cast<float4*>(sm + Hh.off + ch * Hh.ps + ho) = o;
        }
    }
}

// Vertical expansion (pass B): Hh rows p-1, p, p+1 -> destination rows 2p, 2p+1.
__device__ __forceinline__ void expand_v(float* __restrict__ sm, const Geo Hh, const Geo D, int p_lo, int p_hi,
                                         int g_lo, int g_hi, int tid, int nthreads) {
    const int ng = g_hi - g_lo + 1;
    const int n = (p_hi - p_lo + 1) * ng;
    const unsigned inv = ng > 1 ? 0xFFFFFFFFu / (unsigned)ng + 1u : 0u;
    for (int it = tid; it < n; it += nthreads) {
        const int pr = inv ? (int)__umulhi((unsigned)it, inv) : it;
        const int g = g_lo + (it - pr * ng), p = p_lo + pr;
        const int ho = (p - Hh.r0) * Hh.rs + (4 * g - 1 - Hh.c0);
        const int dofs = (2 * p - D.r0) * D.rs + (4 * g - 1 - D.c0);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* hp = sm + Hh.off + ch * Hh.ps + ho;
            const float4 a = *reinterpret_cast<const float4*>(hp - Hh.rs);
            const float4 b = *reinterpret_cast<const float4*>(hp);
            const float4 c = *reinterpret_cast<const float4*>(hp + Hh.rs);
            float4 e, o;
            e.x = fmaf(a.x + c.x, 0.125f, b.x * 0.75f); o.x = (b.x + c.x) * 0.5f;
            e.y = fmaf(a.y + c.y, 0.125f, b.y * 0.75f); o.y = (b.y + c.y) * 0.5f;
            e.z = fmaf(a.z + c.z, 0.125f, b.z * 0.75f); o.z = (b.z + c.z) * 0.5f;
            e.w = fmaf(a.w + c.w, 0.125f, b.w * 0.75f); o.w = (b.w + c.w) * 0.5f;
            float* dp = sm + D.off + ch * D.ps + dofs;
            *reinterpret_cast<float4*>(dp) = e;
            *reinterpret_cast<float4*>(dp + D.rs) = o;
        }
    }
}

// Write the apron of a freshly built region where the tile touches the image border:
// column -1 := column 1 (reflect-101), column n := column n-1 (replicate); then the same for
// rows (whole stored rows, aprons included).  Block-uniform conditions.
__device__ __forceinline__ void fix_aprons(float* __restrict__ sm, const Geo D, int ra, int rb, int ca, int cb,
                                           int nh, int nw, int tid, int nthreads) {
    const bool left = (ca == 0), right = (cb == nw - 1), top = (ra == 0), bot = (rb == nh - 1);
    if (left || right) {
        const int nr = rb - ra + 1;
        for (int it = tid; it < nr * 3; it += nthreads) {
            const int ch = it / nr, r = ra + (it - ch * nr);
            float* row = sm + D.off + ch * D.ps + (r - D.r0) * D.rs - D.c0;
            if (left) row[-1] = row[nw > 1 ? 1 : 0];
            if (right) row[nw] = row[nw - 1];
        }
    }
    if (top || bot) {
        __syncthreads();
        const int c_lo = ca - 1, ncol = cb - ca + 3;
        for (int it = tid; it < ncol * 3; it += nthreads) {
            const int ch = it / ncol, c = c_lo + (it - ch * ncol);
            float* col = sm + D.off + ch * D.ps + (c - D.c0) - D.r0 * D.rs;
            if (top) col[-1 * D.rs] = col[(nh > 1 ? 1 : 0) * D.rs];
            if (bot) col[nh * D.rs] = col[(nh - 1) * D.rs];
        }
    }
}

template <int KMAX, bool F32OUT, bool U8OUT>
struct Tile {
    const ColArgs& a;
    int t, x0, x1, y0, y1;
    const float* l1;          // level-1 plane 0, pre-offset so that l1[r * rs + c] = value(row r, col c)
    int l1_rs, l1_ps;
    float* stage;             // this warp's 1536-byte transpose strip
    int rx1[KMAXF > 0 ? KMAXF : 1], ry1[KMAXF > 0 ? KMAXF : 1], rx2[KMAXF > 0 ? KMAXF : 1], ry2[KMAXF > 0 ? KMAXF : 1];
    bool hit[KMAXF > 0 ? KMAXF : 1];
    float acc[KMAXF > 0 ? KMAXF : 1][3];

    __device__ Tile(const ColArgs& a_) : a(a_) {}

    // one work item: pixels X..X+3, rows 2m and 2m+1.  `active` is false for lanes right of
    // the tile's valid width (they still take part in the warp-wide store transpose).
    // issue the uint8 loads of rows 2m, 2m+1 for pixels X..X+3 (VEC path); nothing reads them here
    __device__ __forceinline__ void load_pixels(int X, int m, bool active, uint32_t (&wq)[2][3]) const {
        if (!active || 2 * m >= y1) return;
        const size_t row_elems = (size_t)a.W * 3;
        const size_t g0 = ((size_t)t * a.H + 2 * m) * row_elems + (size_t)X * 3;
        const uint32_t* fp = reinterpret_cast<const uint32_t*>(a.frames + g0);
        wq[0][0] = __ldg(fp); wq[0][1] = __ldg(fp + 1); wq[0][2] = __ldg(fp + 2);
        if (2 * m + 1 < y1) {
            const uint32_t* fq = reinterpret_cast<const uint32_t*>(a.frames + g0 + row_elems);
            wq[1][0] = __ldg(fq); wq[1][1] = __ldg(fq + 1); wq[1][2] = __ldg(fq + 2);
        }
    }

    template <bool ROI, bool VEC>
    __device__ __forceinline__ void item(int X, int m, bool active, int warp_X0, const uint32_t (&wq)[2][3]) {
        float ve[3][4], vo[3][4];
        const size_t row_elems = (size_t)a.W * 3;
        const size_t g0 = ((size_t)t * a.H + 2 * m) * row_elems + (size_t)X * 3;
        if (active) {
            const int J = X >> 1;
            const float* pl0 = l1 + (m - 1) * l1_rs + (J - 1);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float* pl = pl0 + ch * l1_ps;
                const float2 a0 = *reinterpret_cast<const float2*>(pl), a1 = *reinterpret_cast<const float2*>(pl + 2);
                const float2 b0 = *reinterpret_cast<const float2*>(pl + l1_rs), b1 = *reinterpret_cast<const float2*>(pl + l1_rs + 2);
                const float2 c0 = *reinterpret_cast<const float2*>(pl + 2 * l1_rs), c1 = *reinterpret_cast<const float2*>(pl + 2 * l1_rs + 2);
                const float av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y}, cv[4] = {c0.x, c0.y, c1.x, c1.y};
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // UNSCALED (x8 / x2): the scaling is folded into the horizontal weights
                    ve[ch][j] = fmaf(bv[j], 6.0f, av[j] + cv[j]);
                    vo[ch][j] = bv[j] + cv[j];
                }
            }
        }
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int y = 2 * m + half;
            if (y >= y1) continue;                    // warp-uniform (only the frame's last odd row can be cut)
            const size_t g = g0 + (half ? row_elems : 0);
            const size_t grow = g - (size_t)X * 3;
            float o[12];
            if (active) {
                uint32_t w0, w1, w2;
                if (VEC) {
                    w0 = wq[half][0]; w1 = wq[half][1]; w2 = wq[half][2];
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
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const int px = k / 3, ch = k - 3 * px;
                    // uint8 -> float: 0x4B0000xx = 2^23 + xx
                    const float f = __uint_as_float(__byte_perm(ww[k >> 2], 0x4B000000u, 0x7440 + (k & 3))) - 8388608.0f;
                    const float (&v)[4] = half ? vo[ch] : ve[ch];
                    // even row: v is x8 -> (v0+v2)/64 + 6 v1/64, (v1+v2)/16 ; odd row: v is x2 -> /16, 6/16, /4
                    const float we = half ? 0.0625f : 0.015625f, wc = half ? 0.375f : 0.09375f, wo = half ? 0.25f : 0.0625f;
                    float up;
                    if (px == 0) up = fmaf(v[0] + v[2], we, fmaf(v[1], wc, f));
                    else if (px == 1) up = fmaf(v[1] + v[2], wo, f);
                    else if (px == 2) up = fmaf(v[1] + v[3], we, fmaf(v[2], wc, f));
                    else up = fmaf(v[2] + v[3], wo, f);
                    o[k] = up;
                }
            }
            if (F32OUT) {
                if (VEC) {
                    // The warp's row segment (32 lanes x 48 B = 1536 contiguous bytes) goes through a
                    // private shared-memory strip (conflict-free 16-byte stores at a 12-word lane
                    // stride) and leaves as ONE bulk async copy issued by lane 0 (TMA engine,
                    // cp.async.bulk shared -> global): fully coalesced HBM writes with no LDS / STG
                    // wavefronts on the warp's LSU pipe.  Two strips alternate (one per row).
                    float* strip = stage + half * 384;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // strip's previous copy has read it
                    __syncwarp();
                    if (active) {
                        float4* sp = reinterpret_cast<float4*>(strip + 12 * lane);
                        sp[0] = make_float4(o[0], o[1], o[2], o[3]);
                        sp[1] = make_float4(o[4], o[5], o[6], o[7]);
                        sp[2] = make_float4(o[8], o[9], o[10], o[11]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
                    __syncwarp();
                    if (lane == 0) {
                        const int nbytes = min(32, (x1 - warp_X0) >> 2) * 48;
                        float* wout = a.out_f32 + grow + (size_t)warp_X0 * 3;
                        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(strip);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     :: "l"(wout), "r"(saddr), "r"(nbytes) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (active) {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (X + k / 3 < x1) a.out_f32[g + k] = o[k];
                }
            }
            if (U8OUT && active) {
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
            if (ROI && KMAX > 0 && active) {
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

    // `wfirst`: pixels of this thread's first item, loaded before the tile prologue so their
    // DRAM latency overlaps the region build; every later item's loads are issued one item ahead.
    template <bool ROI, bool VEC>
    __device__ __forceinline__ void run(uint32_t (&wfirst)[2][3]) {
        const int X = x0 + 4 * threadIdx.x;
        const int warp_X0 = x0 + 4 * (threadIdx.x & ~31);
        if (warp_X0 >= x1) return;                   // whole warp right of the image
        const bool active = X < x1;
        uint32_t wcur[2][3], wnext[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) { wcur[r][k] = wfirst[r][k]; wnext[r][k] = 0; }
        for (int m = (y0 >> 1) + threadIdx.y; 2 * m < y1; m += blockDim.y) {
            if (VEC) load_pixels(X, m + blockDim.y, active, wnext);
            item<ROI, VEC>(X, m, active, warp_X0, wcur);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) wcur[r][k] = wnext[r][k];
        }
        if (VEC && F32OUT && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
    const int y0 = by * a.TH, y1 = min(a.H, y0 + a.TH);
    const int L = a.L;

    // the first work item's pixels: in flight during the whole region build
    Tile<KMAX, F32OUT, U8OUT> tl(a);
    tl.t = t; tl.x0 = x0; tl.x1 = x1; tl.y0 = y0; tl.y1 = y1;
    uint32_t wfirst[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (a.vec_ok) tl.load_pixels(x0 + 4 * (int)threadIdx.x, (y0 >> 1) + (int)threadIdx.y, x0 + 4 * (int)threadIdx.x < x1, wfirst);

    // regions [ra,rb] x [ca,cb] (inclusive, inside the image) needed at each level and their
    // storage geometry: computed once per tile by threads 1..L into shared memory (dynamically
    // indexed per-thread arrays would live in local memory and be re-read in every loop).
    // Every thread also derives the level-L geometry itself so the gather from HBM below is
    // issued immediately, not behind the geometry barrier.
    __shared__ int s_reg[VHR_MAX_LEVELS + 1][4];       // ra, rb, ca, cb
    __shared__ Geo s_geo[VHR_MAX_LEVELS + 1];
    Geo qL;
    int raL, rbL, caL, cbL;
    unsigned invL;
    if (a.use_const_tiles) {
        const TileL q = c_tileL[tile];
        raL = q.ra; rbL = q.rb; caL = q.ca; cbL = q.cb;
        qL.c0 = q.c0; qL.r0 = q.r0; qL.rs = q.rs; qL.ps = q.ps; qL.off = 0;
        invL = q.inv_nc3;
    } else {
        int ra = y0, rb = y1 - 1, ca = x0, cb = x1 - 1;
        for (int l = 1; l <= L; ++l) {
            ra = max(0, (ra >> 1) - 1);
            rb = min(a.h[l] - 1, (rb >> 1) + 1);
            ca = max(0, (ca >> 1) - 1);
            cb = min(a.w[l] - 1, (cb >> 1) + 1);
        }
        raL = ra; rbL = rb; caL = ca; cbL = cb;
        qL.c0 = 4 * (ca >> 2) - 5;
        qL.r0 = ra - 2;
        qL.rs = ((cb + 8 - qL.c0) + 3) & ~3;
        qL.ps = (rb - ra + 5) * qL.rs;
        qL.off = 0;
        invL = 0xFFFFFFFFu / (unsigned)((cb - ca + 3) * 3) + 1u;
    }
    if (tid >= 1 && tid <= L) {
        int ra = y0, rb = y1 - 1, ca = x0, cb = x1 - 1;
        for (int l = 1; l <= tid; ++l) {
            ra = max(0, (ra >> 1) - 1);
            rb = min(a.h[l] - 1, (rb >> 1) + 1);
            ca = max(0, (ca >> 1) - 1);
            cb = min(a.w[l] - 1, (cb >> 1) + 1);
        }
        Geo q;
        q.c0 = 4 * (ca >> 2) - 5;                     // column 4g-1 lands on an index that is a multiple of 4
        q.r0 = ra - 2;
        q.rs = ((cb + 8 - q.c0) + 3) & ~3;
        q.ps = (rb - ra + 5) * q.rs;
        q.off = 0;                                    // every level lives at offset 0 (lifetimes do not overlap)
        s_geo[tid] = q;
        s_reg[tid][0] = ra; s_reg[tid][1] = rb; s_reg[tid][2] = ca; s_reg[tid][3] = cb;
    }

    // ---- level L region (+ apron, border-mapped at load time) from HBM into planar smem ----
    {
        const Geo q = qL;
        const int nr = rbL - raL + 3, nc3 = (cbL - caL + 3) * 3;
        const unsigned inv = invL;
        const float* src = a.lvl + ((size_t)t * a.h[L] * a.w[L]) * 3;
        for (int idx = tid; idx < nr * nc3; idx += nthreads) {
            const int r = (int)__umulhi((unsigned)idx, inv), j = idx - r * nc3;
            const int c = (int)__umulhi((unsigned)j, 0x55555556u), ch = j - 3 * c;
            const int ir = raL - 1 + r, ic = caL - 1 + c;
            smf[q.off + ch * q.ps + (ir - q.r0) * q.rs + (ic - q.c0)] =
                __ldg(src + ((size_t)border_map(ir, a.h[L]) * a.w[L] + border_map(ic, a.w[L])) * 3 + ch);
        }
    }
    // ---- levels L-1 .. 1: separable expansion, aprons rewritten after each level ----------
    for (int l = L - 1; l >= 1; --l) {
        __syncthreads();                              // level l+1 complete (and, first time round, the geometry written)
        const Geo S = s_geo[l + 1], D = s_geo[l];
        const int ral = s_reg[l][0], rbl = s_reg[l][1], cal = s_reg[l][2], cbl = s_reg[l][3];
        Geo hh;                                       // H rows: source row indexing, destination column indexing
        hh.c0 = D.c0; hh.rs = D.rs;
        hh.r0 = S.r0;
        hh.ps = (s_reg[l + 1][1] - s_reg[l + 1][0] + 5) * hh.rs;
        hh.off = a.bufH_off;
        const int g_lo = (cal + 1) >> 2, g_hi = (cbl + 1) >> 2;
        const int p_lo = ral >> 1, p_hi = rbl >> 1;
        expand_h(smf, S, hh, p_lo - 1, p_hi + 1, g_lo, g_hi, tid, nthreads);
        __syncthreads();
        expand_v(smf, hh, D, p_lo, p_hi, g_lo, g_hi, tid, nthreads);
        __syncthreads();
        fix_aprons(smf, D, ral, rbl, cal, cbl, a.h[l], a.w[l], tid, nthreads);
    }
    __syncthreads();
    const Geo g1 = s_geo[1];

    // ---- last expansion fused with add-back, store and ROI sums -------------------------------
    tl.l1 = smf + g1.off - g1.r0 * g1.rs - g1.c0;
    tl.l1_rs = g1.rs; tl.l1_ps = g1.ps;
    tl.stage = smf + a.stage_off + (tid >> 5) * 768;
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
    if (a.vec_ok) {
        if (any_hit) tl.template run<true, true>(wfirst); else tl.template run<false, true>(wfirst);
    } else {
        if (any_hit) tl.template run<true, false>(wfirst); else tl.template run<false, false>(wfirst);
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
                                    int T, int K, int TW, int TH, int tiles_x, int tiles_y, double* __restrict__ mean) {
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
    {
        const char* impl = getenv("VHR_COLLAPSE_IMPL");
        if (!impl || strcmp(impl, "tiled") != 0)
            return vhr_collapse_sep(ctx, d_level, d_frames, T, H, W, levels, d_out_f32, d_out_u8, d_rects, K, d_roi_mean, stream);
    }
    ColArgs a;
    memset(&a, 0, sizeof(a));
    a.lvl = d_level; a.frames = d_frames; a.out_f32 = d_out_f32; a.out_u8 = d_out_u8;
    a.T = T; a.H = H; a.W = W; a.L = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    // tile: 128 x 64 by default (best of the measured sweep, profiles/); tunable for experiments
    int TW = 128, TH = 64;
    if (const char* e = getenv("VHR_COLLAPSE_TW")) TW = atoi(e);
    if (const char* e = getenv("VHR_COLLAPSE_TH")) TH = atoi(e);
    if (TW < 128 || TW % 128 != 0 || TW > 1024 || TH < 2 || TH % 2 != 0 || TH > 128) {
        vhr_set_error(ctx, "collapse: bad tile %dx%d", TW, TH);
        return VHR_ERR_INVALID;
    }
    if (W < TW) TW = ((W + 127) / 128) * 128;
    a.TW = TW; a.TH = TH;
    a.tiles_x = (W + TW - 1) / TW;
    a.tiles_y = (H + TH - 1) / TH;
    const int cg = TW / 4;                              // thread columns (multiple of 32)
    int rg = MAXT / cg;                                 // row-pair groups
    if (rg < 1) rg = 1;
    if (rg > TH / 2) rg = TH / 2;
    dim3 block(cg, rg);
    VHR_REQUIRE(ctx, cg * rg <= MAXT, "tile too wide for one CTA");
    a.rects = d_rects; a.K = K;
    // shared memory: X = odd levels (level 1 is the largest), Y = even levels, H = pass-A rows
    size_t szX = 0, szY = 0, szH = 0;
    {
        int rh = TH, rw = TW;
        for (int l = 1; l <= levels; ++l) {
            const int rh_src_rows = rh / 2 / 2 + 4 + 5;            // rows of level l+1 region (+ slack), for H
            rh = rh / 2 + 3;                                      // rows rb-ra+1 of level l
            rw = rw / 2 + 3;
            const size_t rs = (size_t)((rw + 16 + 3) & ~3);
            const size_t n = 3 * (size_t)(rh + 4) * rs;
            if (l & 1) szX = szX > n ? szX : n; else szY = szY > n ? szY : n;
            const size_t nh = 3 * (size_t)rh_src_rows * rs;
            if (l < levels) szH = szH > nh ? szH : nh;
        }
    }
    // Lifetimes let the buffers overlap: a level is dead once pass A has turned it into H rows,
    // and the next level is only written by pass B (a barrier later), so every level lives at
    // offset 0; the store-transpose strips are only used after the last H is dead.
    const size_t szR = szX > szY ? szX : szY;
    const size_t szS = (size_t)((cg * rg + 31) / 32) * 768;
    a.bufX_off = 0;
    a.bufY_off = 0;
    a.bufH_off = (int)((szR + 3) & ~(size_t)3);
    a.stage_off = a.bufH_off;
    const size_t smem = ((size_t)a.bufH_off + (szH > szS ? szH : szS)) * sizeof(float);
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "collapse: tile needs %zu bytes of shared memory (> %d)", smem, ctx->smem_optin);
        return VHR_ERR_UNSUPPORTED;
    }
    // level-L geometry per tile position -> constant memory (rebuilt only when the shape changes)
    {
        static long long s_key = -1;                 // c_tileL is per module: one key for all contexts
        const long long key = ((long long)W << 40) ^ ((long long)H << 20) ^ ((long long)levels << 16) ^ ((long long)TW << 4) ^ (long long)TH;
        const int ntile = a.tiles_x * a.tiles_y;
        a.use_const_tiles = ntile <= MAX_CONST_TILES;
        if (a.use_const_tiles && s_key != key) {
            std::vector<TileL> tab((size_t)ntile);
            for (int by = 0; by < a.tiles_y; ++by)
                for (int bx = 0; bx < a.tiles_x; ++bx) {
                    const int x0 = bx * TW, x1 = (W < x0 + TW) ? W : x0 + TW, y0 = by * TH, y1 = (H < y0 + TH) ? H : y0 + TH;
                    int ra = y0, rb = y1 - 1, ca = x0, cb = x1 - 1;
                    for (int l = 1; l <= levels; ++l) {
                        ra = (ra >> 1) - 1 < 0 ? 0 : (ra >> 1) - 1;
                        rb = (rb >> 1) + 1 > a.h[l] - 1 ? a.h[l] - 1 : (rb >> 1) + 1;
                        ca = (ca >> 1) - 1 < 0 ? 0 : (ca >> 1) - 1;
                        cb = (cb >> 1) + 1 > a.w[l] - 1 ? a.w[l] - 1 : (cb >> 1) + 1;
                    }
                    TileL& q = tab[(size_t)by * a.tiles_x + bx];
                    memset(&q, 0, sizeof(q));
                    q.ra = ra; q.rb = rb; q.ca = ca; q.cb = cb;
                    q.c0 = 4 * (ca >> 2) - 5;
                    q.r0 = ra - 2;
                    q.rs = ((cb + 8 - q.c0) + 3) & ~3;
                    q.ps = (rb - ra + 5) * q.rs;
                    q.inv_nc3 = 0xFFFFFFFFu / (unsigned)((cb - ca + 3) * 3) + 1u;
                }
            VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());     // launches still reading the previous table
            VHR_CHECK_CUDA(ctx, cudaMemcpyToSymbol(c_tileL, tab.data(), tab.size() * sizeof(TileL)));
            s_key = key;
        }
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
    int rc = (K == 0) ? launch_collapse<0>(ctx, a, block, smem, stream)
           : (K == 1) ? launch_collapse<1>(ctx, a, block, smem, stream)
                      : launch_collapse<KMAXF>(ctx, a, block, smem, stream);
    if (rc != VHR_OK) return rc;
    if (K > 0) {
        const int n = T * K * 3;
        roi_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a.partial, d_rects, T, K, TW, TH, a.tiles_x, a.tiles_y, d_roi_mean);
        rc = vhr_after_launch(ctx, "roi_finalize_kernel");
    }
    return rc;
}
