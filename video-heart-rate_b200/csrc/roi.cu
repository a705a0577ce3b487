// ROI reductions on raw frames: rectangle means (reference-pinned) and polygon means / masks
// (frozen exact-integer rule; the reference has no polygon ROI).
//
// Rectangles replace np.mean(roi[:, :, c]) of rppg_VIDEO.py:60-66 / green_avg.py:34: the sum
// of uint8 values is an exact integer, and float64(sum) / float64(count) is one IEEE
// division -- bit-identical to NumPy's float64 pairwise mean.  The optional paint list
// reproduces the reference's overdraw quirk (rppg_VIDEO.py:54,100-106: the bbox, forehead
// and cheek outlines are drawn INTO the frame before the cheek slice is averaged) using the
// thickness-2 cv.rectangle raster model pinned in oracle/roi.py:outline_mask.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int RT = 256;
constexpr int MAXPAINT = 8;

struct PaintArgs {
    int np;
    uint8_t rgb[MAXPAINT][3];
};

__device__ __forceinline__ bool on_outline(int x, int y, int x1, int y1, int x2, int y2) {
    const int xa = min(x1, x2), xb = max(x1, x2), ya = min(y1, y2), yb = max(y1, y2);
    const bool horiz = (abs(y - ya) <= 1 || abs(y - yb) <= 1) && x >= xa && x <= xb;
    const bool vert = (abs(x - xa) <= 1 || abs(x - xb) <= 1) && y >= ya && y <= yb;
    return horiz || vert;
}

template <typename Tsum>
__device__ __forceinline__ Tsum block_sum(Tsum v, Tsum* sh) {
    // fixed-order tree: warp shuffle then warp 0 over the warp totals
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    Tsum r = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
    return r;   // valid in thread 0
}

__global__ void __launch_bounds__(RT) rect_mean_u8_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                           const int32_t* __restrict__ rects, int K,
                                                           const int32_t* __restrict__ paint, PaintArgs pa,
                                                           double* __restrict__ mean) {
    __shared__ unsigned long long sh[RT / 32];
    const int t = blockIdx.x / K, k = blockIdx.x - t * K;
    const int32_t* rc = rects + ((size_t)t * K + k) * 4;
    const int x1 = rc[0], y1 = rc[1], x2 = rc[2], y2 = rc[3];
    const int rw = x2 - x1, rh = y2 - y1;
    double* out = mean + ((size_t)t * K + k) * 3;
    if (rw <= 0 || rh <= 0 || x1 < 0 || y1 < 0 || x2 > W || y2 > H) {
        if (threadIdx.x < 3) out[threadIdx.x] = __longlong_as_double(0x7FF8000000000000ll);
        return;
    }
    int prc[MAXPAINT][4];
    for (int p = 0; p < pa.np; ++p)
        for (int q = 0; q < 4; ++q) prc[p][q] = paint[((size_t)t * pa.np + p) * 4 + q];
    const uint8_t* fr = frames + (size_t)t * H * W * 3;
    unsigned long long s0 = 0, s1 = 0, s2 = 0;
    const int npx = rw * rh;
    for (int idx = threadIdx.x; idx < npx; idx += RT) {
        const int r = idx / rw, c = idx - r * rw;
        const int x = x1 + c, y = y1 + r;
        const uint8_t* p = fr + ((size_t)y * W + x) * 3;
        unsigned v0 = p[0], v1 = p[1], v2 = p[2];
        for (int q = 0; q < pa.np; ++q) {
            if (on_outline(x, y, prc[q][0], prc[q][1], prc[q][2], prc[q][3])) {
                v0 = pa.rgb[q][0]; v1 = pa.rgb[q][1]; v2 = pa.rgb[q][2];
            }
        }
        s0 += v0; s1 += v1; s2 += v2;
    }
    const unsigned long long t0 = block_sum(s0, sh);
    const unsigned long long t1 = block_sum(s1, sh);
    const unsigned long long t2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        const double n = (double)npx;
        out[0] = __ddiv_rn((double)t0, n);
        out[1] = __ddiv_rn((double)t1, n);
        out[2] = __ddiv_rn((double)t2, n);
    }
}

// The same means without a paint list (the analysis harness' clean ROI, green_avg.py:34): one warp per
// ROI row, aligned 4-byte loads, three IDP4A per word (one per channel, the byte -> channel phase of a
// word is (byte offset) mod 3), head and tail bytes masked off.  Exact integer sums, so the means are
// bit-identical to the per-pixel kernel above.
__global__ void __launch_bounds__(RT) rect_mean_u8_rows_kernel(const uint8_t* __restrict__ frames, size_t total_bytes,
                                                                int H, int W, const int32_t* __restrict__ rects, int K,
                                                                double* __restrict__ mean) {
    __shared__ unsigned long long sh[RT / 32];
    const int t = blockIdx.x / K, k = blockIdx.x - t * K;
    const int32_t* rc = rects + ((size_t)t * K + k) * 4;
    const int x1 = rc[0], y1 = rc[1], x2 = rc[2], y2 = rc[3];
    const int rw = x2 - x1, rh = y2 - y1;
    double* out = mean + ((size_t)t * K + k) * 3;
    if (rw <= 0 || rh <= 0 || x1 < 0 || y1 < 0 || x2 > W || y2 > H) {
        if (threadIdx.x < 3) out[threadIdx.x] = __longlong_as_double(0x7FF8000000000000ll);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;       // (frames: 4-byte aligned, checked by the caller)
    unsigned s0 = 0, s1 = 0, s2 = 0;                                  // <= 255 * 4 * words per lane: far below 2^32
    unsigned long long a0 = 0, a1 = 0, a2 = 0;
    for (int r = warp; r < rh; r += RT / 32) {
        const size_t b0 = ((size_t)(y1 + r) * W + x1) * 3 + (size_t)t * H * W * 3;   // byte offset from `frames`
        const size_t w0 = b0 >> 2;                                     // the row's first aligned word
        const int head = (int)(b0 - (w0 << 2));                        // bytes of that word before the row (0..3)
        const int nbytes = rw * 3;
        const int nwords = (head + nbytes + 3) >> 2;
        const uint32_t* __restrict__ wp = reinterpret_cast<const uint32_t*>(frames) + w0;
        const bool clip_tail = ((w0 + nwords) << 2) > total_bytes;     // the clip's last, partial word is in this row
        constexpr int UB = 8;                                          // words per lane in flight (the loop is pure load latency otherwise)
        for (int wb = lane; wb < nwords; wb += 32 * UB) {
            uint32_t v[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int idx = wb + 32 * u;
                v[u] = 0;
                if (idx < nwords) {
                    if (!(clip_tail && idx == nwords - 1)) {
                        v[u] = __ldg(wp + idx);
                    } else {
                        const size_t o = (w0 + idx) << 2;
                        for (int j = 0; o + j < total_bytes; ++j) v[u] |= (uint32_t)frames[o + j] << (8 * j);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int d = 4 * (wb + 32 * u) - head;                 // offset of the word's byte 0 from the row's first byte (>= -3)
                uint32_t x = v[u];
                if (d < 0) x &= 0xFFFFFFFFu << (8 * -d);                // head bytes of the row's first word
                if (d + 4 > nbytes) x &= (d < nbytes) ? 0xFFFFFFFFu >> (8 * (d + 4 - nbytes)) : 0u;   // tail bytes of its last word
                const unsigned ph = (unsigned)(d + 3) % 3u;             // channel of byte 0
                // weight words: byte j counts for channel c iff (ph + j) % 3 == c
                const uint32_t m0 = ph == 0 ? 0x01000001u : (ph == 1 ? 0x00010000u : 0x00000100u);
                const uint32_t m1 = ph == 0 ? 0x00000100u : (ph == 1 ? 0x01000001u : 0x00010000u);
                const uint32_t m2 = ph == 0 ? 0x00010000u : (ph == 1 ? 0x00000100u : 0x01000001u);
                s0 = __dp4a(x, m0, s0);
                s1 = __dp4a(x, m1, s1);
                s2 = __dp4a(x, m2, s2);
            }
        }
        a0 += s0; a1 += s1; a2 += s2;
        s0 = s1 = s2 = 0;
    }
    const unsigned long long t0 = block_sum(a0, sh);
    const unsigned long long t1 = block_sum(a1, sh);
    const unsigned long long t2 = block_sum(a2, sh);
    if (threadIdx.x == 0) {
        const double n = (double)rw * (double)rh;
        out[0] = __ddiv_rn((double)t0, n);
        out[1] = __ddiv_rn((double)t1, n);
        out[2] = __ddiv_rn((double)t2, n);
    }
}

// ---- polygons ----------------------------------------------------------------------------
// inside(x,y) = on any edge (closed segment) OR even-odd parity with half-open spans
// (y0 <= y) != (y1 <= y) and the pixel strictly left of the crossing -- oracle/roi.py:poly_mask.
__device__ __forceinline__ bool poly_inside(int x, int y, const int2* __restrict__ v, int n) {
    bool on = false, par = false;
    for (int i = 0; i < n; ++i) {
        const int2 a = v[i], b = v[i + 1 == n ? 0 : i + 1];
        const long long dx = (long long)b.x - a.x, dy = (long long)b.y - a.y;
        const long long cr = dx * ((long long)y - a.y) - dy * ((long long)x - a.x);
        on |= (cr == 0) && x >= min(a.x, b.x) && x <= max(a.x, b.x) && y >= min(a.y, b.y) && y <= max(a.y, b.y);
        if (dy != 0) {
            const bool strad = (a.y <= y) != (b.y <= y);
            // tt = (x - x0) dy - dx (y - y0) = -cr
            const bool left = (dy > 0) ? (cr > 0) : (cr < 0);
            par ^= (strad && left);
        }
    }
    return on || par;
}

template <typename Tpix>
__global__ void __launch_bounds__(RT) poly_mean_kernel(const Tpix* __restrict__ frames, int H, int W,
                                                        const int32_t* __restrict__ poly, const int32_t* __restrict__ nvert,
                                                        int K, int Vmax, double* __restrict__ mean, long long* __restrict__ count) {
    __shared__ int2 verts[VHR_MAX_POLY_VERTS];
    __shared__ double shd[RT / 32];
    __shared__ unsigned long long shu[RT / 32];
    const int t = blockIdx.x / K, k = blockIdx.x - t * K;
    int n = nvert[(size_t)t * K + k];
    n = min(max(n, 0), min(Vmax, VHR_MAX_POLY_VERTS));
    const int32_t* pv = poly + (((size_t)t * K + k) * Vmax) * 2;
    for (int i = threadIdx.x; i < n; i += RT) verts[i] = make_int2(pv[2 * i], pv[2 * i + 1]);
    __syncthreads();
    int bx0 = W, by0 = H, bx1 = -1, by1 = -1;
    for (int i = 0; i < n; ++i) {
        bx0 = min(bx0, verts[i].x); bx1 = max(bx1, verts[i].x);
        by0 = min(by0, verts[i].y); by1 = max(by1, verts[i].y);
    }
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, W - 1); by1 = min(by1, H - 1);
    const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
    const Tpix* fr = frames + (size_t)t * H * W * 3;
    double s0 = 0, s1 = 0, s2 = 0;
    unsigned long long u0 = 0, u1 = 0, u2 = 0, cnt = 0;
    if (n > 0 && bw > 0 && bh > 0) {
        const int npx = bw * bh;
        for (int idx = threadIdx.x; idx < npx; idx += RT) {
            const int r = idx / bw, c = idx - r * bw;
            const int x = bx0 + c, y = by0 + r;
            if (poly_inside(x, y, verts, n)) {
                const Tpix* p = fr + ((size_t)y * W + x) * 3;
                if (sizeof(Tpix) == 1) { u0 += (unsigned)p[0]; u1 += (unsigned)p[1]; u2 += (unsigned)p[2]; }
                else { s0 += (double)p[0]; s1 += (double)p[1]; s2 += (double)p[2]; }
                ++cnt;
            }
        }
    }
    const unsigned long long ctot = block_sum(cnt, shu);
    double t0, t1, t2;
    if (sizeof(Tpix) == 1) {
        t0 = (double)block_sum(u0, shu); t1 = (double)block_sum(u1, shu); t2 = (double)block_sum(u2, shu);
    } else {
        t0 = block_sum(s0, shd); t1 = block_sum(s1, shd); t2 = block_sum(s2, shd);
    }
    if (threadIdx.x == 0) {
        double* out = mean + ((size_t)t * K + k) * 3;
        if (ctot == 0) {
            out[0] = out[1] = out[2] = __longlong_as_double(0x7FF8000000000000ll);
        } else {
            const double nn = (double)ctot;
            out[0] = __ddiv_rn(t0, nn); out[1] = __ddiv_rn(t1, nn); out[2] = __ddiv_rn(t2, nn);
        }
        if (count) count[(size_t)t * K + k] = (long long)ctot;
    }
}

// ---- scanline form of the same rule ----------------------------------------------------------
// The per-pixel test above costs O(vertices) int64 cross products per pixel of the bounding box
// (33 ms per 1080p clip for a forehead + two cheek polygons).  Per row y the rule decomposes:
//   * parity: an edge that straddles y ((y0 <= y) != (y1 <= y)) counts for the pixels strictly left
//     of its crossing xc, i.e. for x < ceil(xc) (exact int64 rational).  The number of straddling
//     edges of a closed polygon is even, so parity(x) = (number of edges with ceil(xc) <= x) mod 2:
//     toggle one bit per edge at ceil(xc) and take the inclusive prefix-XOR along the row;
//   * on-edge pixels: a horizontal edge at y contributes a run, any other edge at most the one lattice
//     point it passes through at y (dx (y - y0) divisible by dy).
// One warp per row, one lane per edge; the row mask lives in shared memory as bits.
constexpr int SCAN_MAXW = 256;          // mask words per row: bounding boxes up to 8192 pixels wide

// I = long long in general; int when every coordinate (and y) lies within +-16383, so that dx (y - y0) < 2^31: a 64-bit
// integer division costs ~100 instructions, and the three per edge and row were the whole cost of the row-mask kernel.
template <typename I>
__device__ __forceinline__ void scan_row_edges(int y, const int2* __restrict__ v, int n, int bx0, int bw,
                                               uint32_t* __restrict__ tog, uint32_t* __restrict__ edg) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < n; e += 32) {
        const int2 a = v[e], b = v[e + 1 == n ? 0 : e + 1];
        const I dx = (I)b.x - a.x, dy = (I)b.y - a.y;
        if (dy == 0) {
            if (a.y == y) {                                         // horizontal edge (or a repeated vertex) on this row
                const int c0 = max(min(a.x, b.x), bx0) - bx0, c1 = min(max(a.x, b.x), bx0 + bw - 1) - bx0;
                for (int w = c0 >> 5; c0 <= c1 && w <= (c1 >> 5); ++w) {
                    const int lo = max(c0 - 32 * w, 0), hi = min(c1 - 32 * w, 31);
                    atomicOr(&edg[w], (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo));
                }
            }
            continue;
        }
        const bool on_span = y >= min(a.y, b.y) && y <= max(a.y, b.y);
        const bool straddles = (a.y <= y) != (b.y <= y);
        if (!on_span && !straddles) continue;                        // (a straddling edge is always on the span)
        I num = dx * ((I)y - a.y), den = dy;
        if (den < 0) { num = -num; den = -den; }
        const I quo = num / den, rem = num - quo * den;              // truncated division, one divide
        if (on_span && rem == 0) {                                   // lattice point of the segment on this row
            const long long x = (long long)a.x + (long long)quo;
            if (x >= bx0 && x < (long long)bx0 + bw) atomicOr(&edg[(int)(x - bx0) >> 5], 1u << ((int)(x - bx0) & 31));
        }
        if (straddles) {
            long long q = (long long)quo;                            // ceil(num / den), den > 0
            if (num > 0 && rem != 0) ++q;
            const long long xi = (long long)a.x + q;                 // pixels x < xi are left of the crossing
            if (xi < (long long)bx0 + bw) {
                const int c = xi <= bx0 ? 0 : (int)(xi - bx0);
                atomicXor(&tog[c >> 5], 1u << (c & 31));
            }
        }
    }
}

__device__ __forceinline__ void scan_row_mask(int y, const int2* __restrict__ v, int n, int bx0, int bw,
                                              uint32_t* __restrict__ tog, uint32_t* __restrict__ edg, bool small = false) {
    const int lane = threadIdx.x & 31;
    const int nw = (bw + 31) >> 5;
    for (int w = lane; w < nw; w += 32) { tog[w] = 0u; edg[w] = 0u; }
    __syncwarp();
    if (small) scan_row_edges<int>(y, v, n, bx0, bw, tog, edg);
    else scan_row_edges<long long>(y, v, n, bx0, bw, tog, edg);
    __syncwarp();
    // inclusive prefix-XOR along the row, 32 words per step
    uint32_t carry = 0u;                                             // parity of all toggles in the words before this step
    for (int w0 = 0; w0 < nw; w0 += 32) {
        const int w = w0 + lane;
        uint32_t m = w < nw ? tog[w] : 0u;
        uint32_t par = __popc(m) & 1u;                               // parity of this word
        uint32_t inc = par;                                          // inclusive scan of the word parities over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc ^= o;
        }
        m ^= m << 1; m ^= m << 2; m ^= m << 4; m ^= m << 8; m ^= m << 16;
        if ((inc ^ par ^ carry) & 1u) m = ~m;                        // parity of everything left of this word
        if (w < nw) {
            m |= edg[w];
            if (w == nw - 1 && (bw & 31)) m &= 0xFFFFFFFFu >> (32 - (bw & 31));
            tog[w] = m;
        }
        carry ^= __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
}

template <typename Tpix>
__global__ void __launch_bounds__(RT) poly_mean_scan_kernel(const Tpix* __restrict__ frames, int H, int W,
                                                             const int32_t* __restrict__ poly, const int32_t* __restrict__ nvert,
                                                             int K, int Vmax, double* __restrict__ mean, long long* __restrict__ count) {
    __shared__ int2 verts[VHR_MAX_POLY_VERTS];
    __shared__ uint32_t rowmask[RT / 32][2][SCAN_MAXW];
    __shared__ double shd[RT / 32];
    __shared__ unsigned long long shu[RT / 32];
    const int t = blockIdx.x / K, k = blockIdx.x - t * K;
    int n = nvert[(size_t)t * K + k];
    n = min(max(n, 0), min(Vmax, VHR_MAX_POLY_VERTS));
    const int32_t* pv = poly + (((size_t)t * K + k) * Vmax) * 2;
    for (int i = threadIdx.x; i < n; i += RT) verts[i] = make_int2(pv[2 * i], pv[2 * i + 1]);
    __syncthreads();
    int bx0 = W, by0 = H, bx1 = -1, by1 = -1;
    for (int i = 0; i < n; ++i) {
        bx0 = min(bx0, verts[i].x); bx1 = max(bx1, verts[i].x);
        by0 = min(by0, verts[i].y); by1 = max(by1, verts[i].y);
    }
    // every |coordinate| and every row <= 16383: the 32-bit form of the edge arithmetic is exact
    const bool small = n > 0 && bx0 >= -16383 && by0 >= -16383 && bx1 <= 16383 && by1 <= 16383 && H <= 16383;
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, W - 1); by1 = min(by1, H - 1);
    const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
    const Tpix* fr = frames + (size_t)t * H * W * 3;
    double s0 = 0, s1 = 0, s2 = 0;
    unsigned long long u0 = 0, u1 = 0, u2 = 0, cnt = 0;
    if (n > 0 && bw > 0 && bh > 0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        uint32_t* tog = rowmask[warp][0];
        uint32_t* edg = rowmask[warp][1];
        for (int y = by0 + warp; y <= by1; y += RT / 32) {
            scan_row_mask(y, verts, n, bx0, bw, tog, edg, small);
            const Tpix* row = fr + ((size_t)y * W + bx0) * 3;
            for (int c0 = 0; c0 < bw; c0 += 32) {
                const uint32_t m = tog[c0 >> 5];                     // warp-uniform word
                if (m == 0u) continue;
                if ((m >> lane) & 1u) {
                    const Tpix* p = row + (size_t)(c0 + lane) * 3;
                    if (sizeof(Tpix) == 1) { u0 += (unsigned)p[0]; u1 += (unsigned)p[1]; u2 += (unsigned)p[2]; }
                    else { s0 += (double)p[0]; s1 += (double)p[1]; s2 += (double)p[2]; }
                }
                if (lane == 0) cnt += __popc(m);
            }
            __syncwarp();                                            // the mask is rebuilt for the warp's next row
        }
    }
    const unsigned long long ctot = block_sum(cnt, shu);
    double t0, t1, t2;
    if (sizeof(Tpix) == 1) {
        t0 = (double)block_sum(u0, shu); t1 = (double)block_sum(u1, shu); t2 = (double)block_sum(u2, shu);
    } else {
        t0 = block_sum(s0, shd); t1 = block_sum(s1, shd); t2 = block_sum(s2, shd);
    }
    if (threadIdx.x == 0) {
        double* out = mean + ((size_t)t * K + k) * 3;
        if (ctot == 0) {
            out[0] = out[1] = out[2] = __longlong_as_double(0x7FF8000000000000ll);
        } else {
            const double nn = (double)ctot;
            out[0] = __ddiv_rn(t0, nn); out[1] = __ddiv_rn(t1, nn); out[2] = __ddiv_rn(t2, nn);
        }
        if (count) count[(size_t)t * K + k] = (long long)ctot;
    }
}

// Row bit-masks of every polygon for the fused ROI of the collapse (collapse_sep.cu): the same
// scan_row_mask rows, laid out on frame-aligned words (bit x & 31 of word x >> 5) so that a collapse
// lane finds the 4 bits of its 4 pixels in one word.  Only the bounding box is written.
__global__ void __launch_bounds__(RT) poly_rowmask_kernel(int H, int W, const int32_t* __restrict__ poly,
                                                           const int32_t* __restrict__ nvert, int K, int Vmax,
                                                           uint32_t* __restrict__ mask, int MW, int32_t* __restrict__ box,
                                                           long long* __restrict__ count) {
    __shared__ int2 verts[VHR_MAX_POLY_VERTS];
    __shared__ uint32_t rowmask[RT / 32][2][SCAN_MAXW];
    __shared__ unsigned long long shu[RT / 32];
    const int tk = blockIdx.x;
    int n = nvert[tk];
    n = min(max(n, 0), min(Vmax, VHR_MAX_POLY_VERTS));
    const int32_t* pv = poly + ((size_t)tk * Vmax) * 2;
    for (int i = threadIdx.x; i < n; i += RT) verts[i] = make_int2(pv[2 * i], pv[2 * i + 1]);
    __syncthreads();
    int bx0 = W, by0 = H, bx1 = -1, by1 = -1;
    for (int i = 0; i < n; ++i) {
        bx0 = min(bx0, verts[i].x); bx1 = max(bx1, verts[i].x);
        by0 = min(by0, verts[i].y); by1 = max(by1, verts[i].y);
    }
    // every |coordinate| and every row <= 16383: the 32-bit form of the edge arithmetic is exact
    const bool small = n > 0 && bx0 >= -16383 && by0 >= -16383 && bx1 <= 16383 && by1 <= 16383 && H <= 16383;
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, W - 1); by1 = min(by1, H - 1);
    const bool some = n > 0 && bx1 >= bx0 && by1 >= by0;
    unsigned long long cnt = 0;
    if (some) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int ax0 = bx0 & ~31;                                   // word-aligned origin of the row masks
        const int bw = bx1 - ax0 + 1;
        const int nw = (bw + 31) >> 5;
        uint32_t* tog = rowmask[warp][0];
        uint32_t* edg = rowmask[warp][1];
        uint32_t* dst = mask + (size_t)tk * H * MW + (ax0 >> 5);
        for (int y = by0 + warp; y <= by1; y += RT / 32) {
            scan_row_mask(y, verts, n, ax0, bw, tog, edg, small);
            for (int w = lane; w < nw; w += 32) {
                const uint32_t m = tog[w];
                dst[(size_t)y * MW + w] = m;
                cnt += __popc(m);
            }
            __syncwarp();
        }
    }
    const unsigned long long ctot = block_sum(cnt, shu);
    if (threadIdx.x == 0) {
        int32_t* b = box + (size_t)tk * 4;
        if (some && ctot > 0) { b[0] = bx0; b[1] = by0; b[2] = bx1 + 1; b[3] = by1 + 1; }
        else { b[0] = b[1] = b[2] = b[3] = 0; }
        count[tk] = (long long)ctot;
    }
}

__global__ void __launch_bounds__(RT) poly_mask_kernel(int H, int W, const int32_t* __restrict__ poly,
                                                        const int32_t* __restrict__ nvert, int K, int Vmax,
                                                        uint8_t* __restrict__ mask) {
    __shared__ int2 verts[VHR_MAX_POLY_VERTS];
    const int tk = blockIdx.y;
    int n = nvert[tk];
    n = min(max(n, 0), min(Vmax, VHR_MAX_POLY_VERTS));
    const int32_t* pv = poly + ((size_t)tk * Vmax) * 2;
    for (int i = threadIdx.x; i < n; i += RT) verts[i] = make_int2(pv[2 * i], pv[2 * i + 1]);
    __syncthreads();
    const int npx = H * W;
    for (int idx = blockIdx.x * RT + threadIdx.x; idx < npx; idx += gridDim.x * RT) {
        const int y = idx / W, x = idx - y * W;
        mask[(size_t)tk * npx + idx] = (n > 0 && poly_inside(x, y, verts, n)) ? 1 : 0;
    }
}

}  // namespace

int vhr_poly_rowmask(vhr_ctx* ctx, int T, int H, int W, const int32_t* d_poly, const int32_t* d_nvert, int K, int Vmax,
                     uint32_t* d_mask, int MW, int32_t* d_box, long long* d_count, cudaStream_t stream) {
    VHR_REQUIRE(ctx, W <= 32 * SCAN_MAXW && MW >= (W + 31) / 32, "frame too wide for the row masks");
    poly_rowmask_kernel<<<(unsigned)((size_t)T * K), RT, 0, stream>>>(H, W, d_poly, d_nvert, K, Vmax, d_mask, MW, d_box, d_count);
    return vhr_after_launch(ctx, "poly_rowmask_kernel");
}

extern "C" int vhr_roi_mean_rect_u8(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W,
                                    const int32_t* d_rects, int K, const int32_t* d_paint, int NP,
                                    const uint8_t* paint_rgb, double* d_mean, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_frames && d_rects && d_mean, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1 && K >= 1, "bad shape");
    VHR_REQUIRE(ctx, NP >= 0 && NP <= MAXPAINT, "too many paint rectangles (max 8)");
    VHR_REQUIRE(ctx, NP == 0 || (d_paint && paint_rgb), "paint pointers missing");
    PaintArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.np = NP;
    for (int p = 0; p < NP; ++p)
        for (int c = 0; c < 3; ++c) pa.rgb[p][c] = paint_rgb[p * 3 + c];
    const char* px = getenv("VHR_RECT_PIXEL");                 // test hook: the per-pixel kernel
    if (NP == 0 && (reinterpret_cast<uintptr_t>(d_frames) & 3) == 0 && !(px && px[0] == '1')) {
        rect_mean_u8_rows_kernel<<<(unsigned)((size_t)T * K), RT, 0, (cudaStream_t)stream>>>(d_frames, (size_t)T * H * W * 3, H, W,
                                                                                              d_rects, K, d_mean);
        return vhr_after_launch(ctx, "rect_mean_u8_rows_kernel");
    }
    rect_mean_u8_kernel<<<(unsigned)((size_t)T * K), RT, 0, (cudaStream_t)stream>>>(d_frames, H, W, d_rects, K, d_paint, pa, d_mean);
    return vhr_after_launch(ctx, "rect_mean_u8_kernel");
}

template <typename Tpix>
static int poly_mean_impl(vhr_ctx* ctx, const Tpix* d_frames, int T, int H, int W, const int32_t* d_poly,
                          const int32_t* d_nvert, int K, int Vmax, double* d_mean, int64_t* d_count, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_frames && d_poly && d_nvert && d_mean, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1 && K >= 1, "bad shape");
    VHR_REQUIRE(ctx, Vmax >= 1 && Vmax <= VHR_MAX_POLY_VERTS, "Vmax must be 1..64");
    const char* px = getenv("VHR_POLY_PIXEL");                 // test hook: the per-pixel form of the rule
    if (W <= 32 * SCAN_MAXW && !(px && px[0] == '1')) {
        poly_mean_scan_kernel<Tpix><<<(unsigned)((size_t)T * K), RT, 0, (cudaStream_t)stream>>>(
            d_frames, H, W, d_poly, d_nvert, K, Vmax, d_mean, reinterpret_cast<long long*>(d_count));
        return vhr_after_launch(ctx, "poly_mean_scan_kernel");
    }
    poly_mean_kernel<Tpix><<<(unsigned)((size_t)T * K), RT, 0, (cudaStream_t)stream>>>(
        d_frames, H, W, d_poly, d_nvert, K, Vmax, d_mean, reinterpret_cast<long long*>(d_count));
    return vhr_after_launch(ctx, "poly_mean_kernel");
}

extern "C" int vhr_roi_mean_poly_u8(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, const int32_t* d_poly,
                                    const int32_t* d_nvert, int K, int Vmax, double* d_mean, int64_t* d_count, void* stream) {
    return poly_mean_impl<uint8_t>(ctx, d_frames, T, H, W, d_poly, d_nvert, K, Vmax, d_mean, d_count, stream);
}

extern "C" int vhr_roi_mean_poly_f32(vhr_ctx* ctx, const float* d_frames, int T, int H, int W, const int32_t* d_poly,
                                     const int32_t* d_nvert, int K, int Vmax, double* d_mean, int64_t* d_count, void* stream) {
    return poly_mean_impl<float>(ctx, d_frames, T, H, W, d_poly, d_nvert, K, Vmax, d_mean, d_count, stream);
}

extern "C" int vhr_poly_mask(vhr_ctx* ctx, int T, int H, int W, const int32_t* d_poly, const int32_t* d_nvert, int K,
                             int Vmax, uint8_t* d_mask, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_poly && d_nvert && d_mask, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1 && K >= 1, "bad shape");
    VHR_REQUIRE(ctx, Vmax >= 1 && Vmax <= VHR_MAX_POLY_VERTS, "Vmax must be 1..64");
    VHR_REQUIRE(ctx, (long long)T * K <= 65535, "T*K too large for one mask call");
    const int npx = H * W;
    dim3 grid((unsigned)min((npx + RT - 1) / RT, 1024), (unsigned)(T * K));
    poly_mask_kernel<<<grid, RT, 0, (cudaStream_t)stream>>>(H, W, d_poly, d_nvert, K, Vmax, d_mask);
    return vhr_after_launch(ctx, "poly_mask_kernel");
}
