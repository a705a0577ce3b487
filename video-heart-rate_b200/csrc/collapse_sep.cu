// EVM reconstruction, separable-composite form: pyrUp^L collapse of the filtered, amplified
// level + add-back to the original uint8 frame, rectangle-ROI sums fused into the same pass.
//
// No reference code exists for the collapse; spec = cv2.pyrUp on
// float32 applied L times (oracle/evm.py:pyrup / collapse).  Every pyrUp is linear and
// separable, borders included (low side reflect-101, high side replicate, odd sizes walking
// back the pyrDown chain), so L of them compose into ONE separable operator
//        up = Uy . level_L . Ux^T
// whose rows have at most FOUR non-zero weights (exact multiples of 8^-L, so exact in f32),
// with a first-tap index that never decreases and never steps by more than one.  The host
// composes Uy / Ux once per shape (exactly, in double) and the kernel evaluates
//        out(y, x, c) = f32(frame(y, x, c)) + sum_a wy[y][a] * Hx[yb[y] + a](x, c)
//        Hx[r](x, c)  = sum_b wx[x][b] * level_L(r, xb[x] + b, c)
// which costs 4 FMA per output value instead of rebuilding the level chain in shared memory
// (the tiled kernel this replaces spent ~25 thread-instructions per value and was issue-bound).
//
// Work decomposition: one WARP = one item = a 128-pixel column
// segment x a band of rows of one frame; a lane owns 4 pixels (12 values) of every row.
//   * the lane keeps a 4-row window of Hx for its 12 values in registers; when yb[y] steps,
//     the window shifts and one new Hx row is built (48 FMA per 2^L output rows) from
//     level_L values that were prefetched towards L1 one step earlier;
//   * the vertical weights wy[y] are warp-uniform;
//   * pixels: 3 x LDG.32 per lane per row, issued PF rows ahead; uint8 -> float by
//     byte-permute into 2^23 + b, one FADD;
//   * stores: the warp's 1536-byte row segment is staged in a private shared-memory strip
//     and leaves as one cp.async.bulk shared -> global (TMA engine), NSTRIP strips in
//     flight per warp; no barrier wider than a warp anywhere in the kernel;
//   * ROI sums: per-lane float accumulators -> warp shuffle -> one double partial per item
//     -> fixed-order finalize kernel.  No atomics: results are run-to-run identical.
//     Rectangles are tested by coordinates; polygons (forehead / cheek outlines) are rasterised
//     once per frame by poly_rowmask_kernel (roi.cu: the frozen exact-integer rule, one bit per
//     pixel, frame-aligned 32-pixel words) and an item reads the 4 bits of its lane's pixels;
//   * ROI-only calls (no frame output requested, the measurement plugins' case) retire every
//     item that touches no ROI before it requests a single pixel.
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>
#include <math.h>
#include <vector>

namespace {

constexpr int WARPS = 8;                // warps (independent items) per CTA
constexpr int NSTRIP = 4;               // bulk-store strips in flight per warp
constexpr int PF = 3;                   // pixel rows requested ahead of use (register path)
constexpr int DEPTH = 8;                // pixel rows in flight per warp (bulk-copy ring; power of two)
constexpr int WARP_SMEM = NSTRIP * 1536 + DEPTH * 384 + DEPTH * 8 + 64;   // bytes per warp, multiple of 128
constexpr int KMAXF = 4;                // fused ROI rectangles per call

struct SepArgs {
    const float* lvl;
    const uint8_t* frames;
    float* out_f32;
    uint8_t* out_u8;
    int T, H, W;
    int hL, wL;
    int BH;                   // rows per band
    int nbands, nsegs;
    long long n_items;
    const float4* xw;         // (nsegs*128) horizontal weights per pixel
    const int* xb;            // (nsegs*32)  first level-L column per aligned 4-pixel group
    const float4* yw;         // (H) vertical weights per row
    const int* yb;            // (H) first level-L row per output row
    const int32_t* rects;     // ROI k of frame t at rects[(t * Kstride + k) * 4]: rectangle, or a polygon's clamped bounding box
    int K, Kstride;
    const uint32_t* mask;     // RM == 2: (T, K, H, MW) row bit-masks of the polygons (bit x & 31 of word x >> 5)
    int MW;
    double* partial;          // (T, nbands, nsegs, K, 3)
};

// ---- PTX helpers (sm_100a) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
#ifdef VHR_WATCHDOG
__device__ __noinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (long long spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > 2000000) {
            if ((threadIdx.x & 31) == 0) printf("WATCHDOG(collapse) block %d warp %d parity %u\n", blockIdx.x, threadIdx.x >> 5, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}"
        :: "r"(bar), "r"(parity) : "memory");
}
#endif
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// LOAD: 0 = scalar pixel loads (any W), 1 = 3 x LDG.32 per lane (W % 4 == 0),
//       2 = one cp.async.bulk per warp row into a shared-memory ring (W % 16 == 0).
// RM: 0 = no ROI, 1 = rectangles, 2 = polygon row masks.
template <int KMAX, int RM, bool F32OUT, bool U8OUT, int LOAD>
__global__ void __launch_bounds__(WARPS * 32, 2) collapse_sep_kernel(const SepArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool VEC = LOAD >= 1;
    constexpr bool TMA = LOAD == 2;
    const int lane = threadIdx.x & 31;
    // warp-uniform values are broadcast from lane 0 so that the compiler keeps everything
    // derived from them (item decode, row addresses, bulk-copy operands) on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long item = (long long)blockIdx.x * WARPS + warp;
    if (item >= a.n_items) return;
    const int seg = (int)(item % a.nsegs);
    const long long q = item / a.nsegs;
    const int band = (int)(q % a.nbands);
    const int t = (int)(q / a.nbands);
    const int X0 = seg * 128;
    const int X = X0 + 4 * lane;
    const int y0 = band * a.BH, y1 = min(a.H, y0 + a.BH);
    const bool active = X < a.W;
    const int Xc = active ? X : X0;                    // address-safe column for idle lanes

    // ROIs that intersect this item (rectangle, or bounding box of a polygon)
    int rx1[KMAXF > 0 ? KMAXF : 1], ry1[KMAXF > 0 ? KMAXF : 1], rx2[KMAXF > 0 ? KMAXF : 1], ry2[KMAXF > 0 ? KMAXF : 1];
    bool hit[KMAXF > 0 ? KMAXF : 1];
    uint32_t moff[KMAXF > 0 ? KMAXF : 1];        // RM == 2: word offset of this lane's mask column in row 0 of ROI k (host checks < 2^32)
    float acc[KMAXF > 0 ? KMAXF : 1][3];
    bool any_hit = false;
    if (KMAX > 0) {
        const int xe = min(a.W, X0 + 128);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            hit[k] = false;
            acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
            if (k < a.K) {
                const int32_t* rc = a.rects + ((size_t)t * a.Kstride + k) * 4;
                rx1[k] = rc[0]; ry1[k] = rc[1]; rx2[k] = rc[2]; ry2[k] = rc[3];
                hit[k] = rc[0] < xe && rc[2] > X0 && rc[1] < y1 && rc[3] > y0 && rc[2] > rc[0] && rc[3] > rc[1];
                any_hit |= hit[k];
                if (RM == 2) moff[k] = (uint32_t)(((size_t)t * a.Kstride + k) * a.H * a.MW + (X >> 5));
            }
        }
    }
    // ROI-only call: an item outside every ROI has nothing to produce (warp-uniform exit, before any pixel is requested)
    if (!F32OUT && !U8OUT && !any_hit) return;

    // per-warp shared memory: NSTRIP output strips, then the pixel ring, then its barriers
    unsigned char* wsm = smem_raw + (size_t)warp * WARP_SMEM;
    float* mystrips = reinterpret_cast<float*>(wsm);
    unsigned char* ring = wsm + NSTRIP * 1536;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t bar_u32 = ring_u32 + DEPTH * 384;

    // first level-L column of this lane's 4 pixels (they share one 4-tap window)
    int coff[4];
    {
        const int xb = __ldg(a.xb + (X >> 2));
#pragma unroll
        for (int b = 0; b < 4; ++b) coff[b] = min(xb + b, a.wL - 1) * 3;
    }
    const float* Lf = a.lvl + (size_t)t * a.hL * a.wL * 3;
    const int lrow = a.wL * 3;

    // one row of Hx for this lane's 12 values: level-L row r (clamped) times the horizontal
    // weights.  The weights are re-read from the table (L1-resident) instead of being held in
    // 16 registers across the row loop; this runs once per 2^L output rows.
    auto hx_row = [&](int r, float (&o)[12]) {
        const float* p = Lf + (size_t)min(r, a.hL - 1) * lrow;
        float v[12];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) v[b * 3 + ch] = __ldg(p + coff[b] + ch);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 w = __ldg(a.xw + X + i);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float s = w.x * v[ch];
                s = fmaf(w.y, v[3 + ch], s);
                s = fmaf(w.z, v[6 + ch], s);
                s = fmaf(w.w, v[9 + ch], s);
                o[i * 3 + ch] = s;
            }
        }
    };
    auto prefetch_lrow = [&](int r) {
        const float* p = Lf + (size_t)min(r, a.hL - 1) * lrow;
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p + coff[0]));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p + coff[3] + 2));
    };

    // pixel pipeline
    const size_t row_bytes = (size_t)a.W * 3;
    const uint8_t* fwarp = a.frames + ((size_t)t * a.H + y0) * row_bytes + (size_t)X0 * 3;   // warp-uniform
    const uint32_t seg_bytes = (uint32_t)min(128, a.W - X0) * 3;                            // TMA: multiple of 16
    const int lane_off = (Xc - X0) * 3;
    uint32_t pix[PF][3];
    auto load_pix = [&](int y, uint32_t (&w)[3]) {                                            // LOAD 0 / 1
        if (y >= y1) return;
        const uint8_t* p = fwarp + (size_t)(y - y0) * row_bytes + lane_off;
        if (VEC) {
            const uint32_t* fp = reinterpret_cast<const uint32_t*>(p);
            w[0] = __ldg(fp); w[1] = __ldg(fp + 1); w[2] = __ldg(fp + 2);
        } else {
            w[0] = w[1] = w[2] = 0;
#pragma unroll
            for (int k = 0; k < 12; ++k)
                if (X + k / 3 < a.W) w[k >> 2] |= (uint32_t)__ldg(p + k) << (8 * (k & 3));
        }
    };
    if (TMA) {
        if (elect_one()) {
#pragma unroll
            for (int i = 0; i < DEPTH; ++i) mbar_init(bar_u32 + 8 * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int i = 0; i < DEPTH; ++i)
                if (y0 + i < y1) {
                    mbar_expect_tx(bar_u32 + 8 * i, seg_bytes);
                    bulk_g2s(ring_u32 + 384 * i, fwarp + (size_t)i * row_bytes, seg_bytes, bar_u32 + 8 * i);
                }
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            pix[i][0] = pix[i][1] = pix[i][2] = 0;
            load_pix(y0 + i, pix[i]);
        }
    }

    // Hx window for the first row of the band
    const int* ybp = a.yb + y0;
    const float4* ywp = a.yw + y0;
    int cur = __ldg(ybp);
    float Wn[4][12];
    hx_row(cur, Wn[0]); hx_row(cur + 1, Wn[1]); hx_row(cur + 2, Wn[2]); hx_row(cur + 3, Wn[3]);
    prefetch_lrow(cur + 4);

    const uint32_t nbytes = (uint32_t)min(32, (a.W - X0) >> 2) * 48;          // VEC: W % 4 == 0
    float* owarp = F32OUT ? a.out_f32 + (((size_t)t * a.H + y0) * a.W + X0) * 3 : nullptr;   // warp-uniform, this row
    uint8_t* ulane = U8OUT ? a.out_u8 + ((size_t)t * a.H + y0) * row_bytes + (size_t)Xc * 3 : nullptr;
    const uint8_t* fnext = fwarp + (size_t)DEPTH * row_bytes;                 // TMA: next row to request
    const int nrows = y1 - y0;

    // RM == 2: the mask word of this lane's 4 pixels is fetched one row ahead (a dependent load per row and ROI stalled
    // the items inside the bounding boxes: the polygon step ran 0.9 ms behind the rectangle one).  0 = no pixel of
    // this lane inside ROI k on that row (only the rows / words of the bounding box were ever written).
    uint32_t mnext[KMAXF > 0 ? KMAXF : 1];
    auto mask_word = [&](int k, int y) -> uint32_t {
        if (hit[k] && y >= ry1[k] && y < ry2[k] && y < y1 && X + 4 > rx1[k] && X < rx2[k])
            return __ldg(a.mask + (moff[k] + (uint32_t)y * (uint32_t)a.MW));
        return 0u;
    };
    if (KMAX > 0 && RM == 2 && any_hit && active) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) mnext[k] = mask_word(k, y0);
    }

    // The row loop exists twice: items that touch no ROI (88 % of them for a forehead + two cheeks) run an instance
    // that carries none of the ROI state -- with the ROI registers live the loop body was 33 instructions longer for
    // EVERY row (register shuffling at the 128-register cap), 1.0 ms per clip (profiles/README.md, round 2).
    auto row_loop = [&](auto roi_c) {
        constexpr bool ROI = decltype(roi_c)::value;
        for (int i = 0; i < nrows; ++i) {
            const int y = y0 + i;
            const int yb = __ldg(ybp + i);
            const float4 wy = __ldg(ywp + i);
            if (yb != cur) {                                 // warp-uniform; the host guarantees yb == cur + 1
                cur = yb;
    #pragma unroll
                for (int k = 0; k < 12; ++k) { Wn[0][k] = Wn[1][k]; Wn[1][k] = Wn[2][k]; Wn[2][k] = Wn[3][k]; }
                hx_row(cur + 3, Wn[3]);
                prefetch_lrow(cur + 4);
            }
            uint32_t w[3];
            const int slot = i & (DEPTH - 1);
            if (TMA) {
                mbar_wait(bar_u32 + 8 * slot, (uint32_t)(i / DEPTH) & 1u);
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(ring + 384 * slot + lane_off);
                w[0] = rp[0]; w[1] = rp[1]; w[2] = rp[2];
            } else {
                w[0] = pix[0][0]; w[1] = pix[0][1]; w[2] = pix[0][2];
    #pragma unroll
                for (int j = 0; j + 1 < PF; ++j) { pix[j][0] = pix[j + 1][0]; pix[j][1] = pix[j + 1][1]; pix[j][2] = pix[j + 1][2]; }
                load_pix(y + PF, pix[PF - 1]);
            }

            float o[12];
    #pragma unroll
            for (int k = 0; k < 12; ++k) {
                // uint8 -> float: 0x4B0000xx = 2^23 + xx
                float f = __uint_as_float(__byte_perm(w[k >> 2], 0x4B000000u, 0x7440 + (k & 3))) - 8388608.0f;
                f = fmaf(wy.x, Wn[0][k], f);
                f = fmaf(wy.y, Wn[1][k], f);
                f = fmaf(wy.z, Wn[2][k], f);
                f = fmaf(wy.w, Wn[3][k], f);
                o[k] = f;
            }

            if (F32OUT && VEC) {
                // the strip's previous bulk store must have finished reading it
                float* strip = mystrips + (i & (NSTRIP - 1)) * 384;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NSTRIP - 1) : "memory");
                __syncwarp();                                // strip free; every lane has read its ring slot
                if (active) {
                    float4* sp = reinterpret_cast<float4*>(strip + 12 * lane);
                    sp[0] = make_float4(o[0], o[1], o[2], o[3]);
                    sp[1] = make_float4(o[4], o[5], o[6], o[7]);
                    sp[2] = make_float4(o[8], o[9], o[10], o[11]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
                __syncwarp();
                if (elect_one()) {
                    bulk_s2g(owarp + (size_t)i * a.W * 3, smem_u32(strip), nbytes);
                    if (TMA && i + DEPTH < nrows) {
                        mbar_expect_tx(bar_u32 + 8 * slot, seg_bytes);
                        bulk_g2s(ring_u32 + 384 * slot, fnext + (size_t)i * row_bytes, seg_bytes, bar_u32 + 8 * slot);
                    }
                }
            } else {
                if (F32OUT && active) {
                    float* dst = owarp + (size_t)i * a.W * 3 + 12 * lane;
    #pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (X + k / 3 < a.W) dst[k] = o[k];
                }
                if (TMA) {
                    __syncwarp();                            // every lane has read its ring slot
                    if (i + DEPTH < nrows && elect_one()) {
                        mbar_expect_tx(bar_u32 + 8 * slot, seg_bytes);
                        bulk_g2s(ring_u32 + 384 * slot, fnext + (size_t)i * row_bytes, seg_bytes, bar_u32 + 8 * slot);
                    }
                }
            }
            if (U8OUT && active) {
                uint32_t qv[3] = {0, 0, 0};
    #pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const float v = fminf(fmaxf(o[k], 0.0f), 255.0f);
                    qv[k >> 2] |= (uint32_t)(int)(v + 0.5f) << (8 * (k & 3));
                }
                uint8_t* dst = ulane + (size_t)i * row_bytes;
                if (VEC) {
                    uint32_t* op = reinterpret_cast<uint32_t*>(dst);
                    op[0] = qv[0]; op[1] = qv[1]; op[2] = qv[2];
                } else {
    #pragma unroll
                    for (int k = 0; k < 12; ++k)
                        if (X + k / 3 < a.W) dst[k] = (uint8_t)(qv[k >> 2] >> (8 * (k & 3)));
                }
            }
            if (ROI && RM == 2 && any_hit && active) {
    #pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    // 4 mask bits of this lane's pixels (X % 4 == 0: they never straddle a word)
                    const uint32_t bits = (mnext[k] >> (X & 31)) & 0xFu;
                    mnext[k] = mask_word(k, y + 1);
                    if (bits) {
    #pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            if ((bits >> px) & 1u) {
                                acc[k][0] += o[3 * px]; acc[k][1] += o[3 * px + 1]; acc[k][2] += o[3 * px + 2];
                            }
                        }
                    }
                }
            }
            if (ROI && RM != 2 && any_hit && active) {
    #pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (hit[k] && y >= ry1[k] && y < ry2[k]) {
                        {
    #pragma unroll
                            for (int px = 0; px < 4; ++px) {
                                if (X + px >= rx1[k] && X + px < rx2[k] && X + px < a.W) {
                                    acc[k][0] += o[3 * px]; acc[k][1] += o[3 * px + 1]; acc[k][2] += o[3 * px + 2];
                                }
                            }
                        }
                    }
                }
            }
        }
    };
    if constexpr (KMAX > 1) {
        if (!any_hit) {                                  // nothing of the ROI state is live on this path
            row_loop(std::false_type{});
            if (VEC && F32OUT && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            return;
        }
        row_loop(std::true_type{});
    } else {
        row_loop(std::integral_constant<bool, (KMAX > 0)>{});      // one ROI: its state is cheap, one instance (8.90 vs 9.03 ms)
    }
    if (VEC && F32OUT && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");

    if (KMAX > 0 && any_hit) {                           // warp-uniform
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = acc[k][c];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == 0 && k < a.K)
                    a.partial[((((size_t)t * a.nbands + band) * a.nsegs + seg) * a.K + k) * 3 + c] = (double)v;
            }
        }
    }
}

// fixed-order reduction of the per-item partials over the items a ROI's rectangle / bounding box touches.
// Rectangles outside the frame give NaN (like rect_mean_u8_kernel: the caller is expected to have applied
// NumPy slice semantics, host.slice_rects); polygons divide by their rasterised pixel count.
__global__ void roi_finalize_sep_kernel(const double* __restrict__ partial, const int32_t* __restrict__ rects, int Kstride,
                                        const long long* __restrict__ count, int T, int K, int H, int W, int BH, int nbands,
                                        int nsegs, double* __restrict__ mean) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * K * 3) return;
    const int c = idx % 3, k = (idx / 3) % K, t = idx / (3 * K);
    const int32_t* rc = rects + ((size_t)t * Kstride + k) * 4;
    const int x1 = rc[0], y1 = rc[1], x2 = rc[2], y2 = rc[3];
    double* out = mean + ((size_t)t * Kstride + k) * 3 + c;
    const long long n = count ? count[(size_t)t * Kstride + k] : (long long)(x2 - x1) * (long long)(y2 - y1);
    if (x2 <= x1 || y2 <= y1 || x1 < 0 || y1 < 0 || x2 > W || y2 > H || n <= 0) {
        *out = __longlong_as_double(0x7FF8000000000000ll);   // NaN, like np.mean of an empty slice
        return;
    }
    double s = 0.0;
    for (int b = y1 / BH; b <= (y2 - 1) / BH && b < nbands; ++b)
        for (int sg = x1 / 128; sg <= (x2 - 1) / 128 && sg < nsegs; ++sg)
            s += partial[((((size_t)t * nbands + b) * nsegs + sg) * K + k) * 3 + c];
    *out = s / (double)n;
}

// ---- host: compose `levels` pyrUp steps along one axis into a banded matrix ---------------
// rows[o] = weights over the level-L samples that output sample o depends on (dense, double:
// every weight is an exact multiple of 8^-levels).  Follows oracle/evm.py:_pyrup_axis.
void compose_axis(int n0, int levels, std::vector<std::vector<double>>& rows, int& nL) {
    std::vector<int> n(levels + 1);
    n[0] = n0;
    for (int l = 1; l <= levels; ++l) n[l] = (n[l - 1] + 1) / 2;
    nL = n[levels];
    std::vector<std::vector<double>> cur((size_t)nL, std::vector<double>((size_t)nL, 0.0));
    for (int i = 0; i < nL; ++i) cur[i][i] = 1.0;
    for (int l = levels; l >= 1; --l) {
        const int ns = n[l], nd = n[l - 1];
        std::vector<std::vector<double>> nxt((size_t)nd, std::vector<double>((size_t)nL, 0.0));
        for (int o = 0; o < nd; ++o) {
            const int i = o >> 1;
            const int im1 = i - 1 < 0 ? (ns > 1 ? 1 : 0) : i - 1;      // reflect-101 low side
            const int ip1 = i + 1 >= ns ? ns - 1 : i + 1;              // replicate high side
            for (int j = 0; j < nL; ++j)
                nxt[o][j] = (o & 1) ? (cur[i][j] + cur[ip1][j]) * 0.5
                                    : (cur[im1][j] + 6.0 * cur[i][j] + cur[ip1][j]) * 0.125;
        }
        cur.swap(nxt);
    }
    rows.swap(cur);
}

// base/weight tables with `group` consecutive outputs sharing one base; n_pad >= n0 entries
// (padding: zero weights).  Returns false if some group needs more than 4 taps or the base
// sequence is not monotone with steps <= 1 (never the case for a pyrUp chain; checked anyway).
bool make_tables(int n0, int levels, int group, int n_pad, std::vector<float>& wt, std::vector<int>& base, int& nL) {
    std::vector<std::vector<double>> rows;
    compose_axis(n0, levels, rows, nL);
    wt.assign((size_t)n_pad * 4, 0.f);
    base.assign((size_t)(n_pad + group - 1) / group, 0);
    int prev = 0;
    for (int g0 = 0; g0 < n_pad; g0 += group) {
        int lo = nL, hi = -1;
        for (int o = g0; o < g0 + group && o < n0; ++o)
            for (int j = 0; j < nL; ++j)
                if (rows[o][j] != 0.0) { lo = j < lo ? j : lo; hi = j > hi ? j : hi; }
        int b;
        if (hi < 0) b = prev;                      // padding group
        else {
            if (hi - lo > 3) return false;
            b = lo;
            const int bmax = nL - 4 > 0 ? nL - 4 : 0;
            if (b > bmax) b = bmax;
            if (group == 1 && g0 > 0 && (b < prev || b - prev > 1)) return false;
        }
        base[g0 / group] = b;
        prev = b;
        for (int o = g0; o < g0 + group && o < n0; ++o)
            for (int k = 0; k < 4; ++k)
                if (b + k < nL) wt[(size_t)o * 4 + k] = (float)rows[o][b + k];
    }
    return true;
}

template <int KMAX, int RM>
int launch_sep(vhr_ctx* ctx, const SepArgs& a, int load, cudaStream_t stream) {
    const unsigned grid = (unsigned)((a.n_items + WARPS - 1) / WARPS);
    const size_t smem = (size_t)WARPS * WARP_SMEM;
#define VHR_SEP_LAUNCH(F, U, V)                                                                               \
    do {                                                                                                      \
        auto kern = collapse_sep_kernel<KMAX, RM, F, U, V>;                                                   \
        VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, WARPS * 32, smem, stream>>>(a);                                                          \
    } while (0)
#define VHR_SEP_LAUNCH_V(F, U)                                                                                \
    do {                                                                                                      \
        if (load == 2) VHR_SEP_LAUNCH(F, U, 2); else if (load == 1) VHR_SEP_LAUNCH(F, U, 1); else VHR_SEP_LAUNCH(F, U, 0); \
    } while (0)
    if (a.out_f32 && a.out_u8) VHR_SEP_LAUNCH_V(true, true);
    else if (a.out_f32) VHR_SEP_LAUNCH_V(true, false);
    else if (a.out_u8) VHR_SEP_LAUNCH_V(false, true);
    else VHR_SEP_LAUNCH_V(false, false);
#undef VHR_SEP_LAUNCH_V
#undef VHR_SEP_LAUNCH
    return vhr_after_launch(ctx, "collapse_sep_kernel");
}

struct RoiSpec {                 // one fused-ROI request
    const int32_t* rects;        // (T,K,4) rectangles, or NULL when polygons are given
    const int32_t* poly;         // (T,K,Vmax,2)
    const int32_t* nvert;        // (T,K)
    int K, Vmax;
    double* mean;                // (T,K,3)
    int64_t* count;              // (T,K) polygons only, optional
};

// Frames [0, T) of the pointers given; ROIs [k0, k0 + Kg) of `roi` (Kg <= KMAXF); frame outputs may be NULL.
int collapse_group(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames, int T, int H, int W, int levels,
                   float* d_out_f32, uint8_t* d_out_u8, const RoiSpec& roi, int k0, int Kg, cudaStream_t stream) {
    SepArgs a;
    memset(&a, 0, sizeof(a));
    a.lvl = d_level; a.frames = d_frames; a.out_f32 = d_out_f32; a.out_u8 = d_out_u8;
    a.T = T; a.H = H; a.W = W;
    const PyrDims d = vhr_make_dims(W, H, levels);
    a.hL = d.h[levels]; a.wL = d.w[levels];
    int BH = 0;
    if (const char* e = getenv("VHR_COLLAPSE_BH")) BH = atoi(e);
    if (BH < 1) BH = (H + 7) / 8;                       // 8 bands per frame by default
    if (BH < 32) BH = H < 32 ? H : 32;
    a.BH = BH;
    a.nbands = (H + BH - 1) / BH;
    a.nsegs = (W + 127) / 128;
    a.n_items = (long long)T * a.nbands * a.nsegs;
    VHR_REQUIRE(ctx, (a.n_items + WARPS - 1) / WARPS < 0x7fffffffLL, "too many work items");

    // composite weight tables (rebuilt only when the shape changes)
    const int Wp = a.nsegs * 128;
    const long long key = ((long long)W << 36) ^ ((long long)H << 12) ^ (long long)levels;
    const size_t tab_bytes = (size_t)Wp * 16 + (size_t)H * 16 + (size_t)(Wp / 4) * 4 + (size_t)H * 4;
    if (ctx->sep_key != key || !ctx->sep_tab) {
        std::vector<float> xw, yw;
        std::vector<int> xb, yb;
        int nLx = 0, nLy = 0;
        if (!make_tables(W, levels, 4, Wp, xw, xb, nLx) || !make_tables(H, levels, 1, H, yw, yb, nLy)) {
            vhr_set_error(ctx, "collapse: composite pyrUp weights do not fit a 4-tap window for %dx%d, %d levels", W, H, levels);
            return VHR_ERR_UNSUPPORTED;
        }
        VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());    // launches still reading the previous tables
        if (tab_bytes > ctx->sep_tab_bytes) {
            if (ctx->sep_tab) VHR_CHECK_CUDA(ctx, cudaFree(ctx->sep_tab));
            ctx->sep_tab = nullptr; ctx->sep_tab_bytes = 0; ctx->sep_key = -1;
            VHR_CHECK_CUDA(ctx, cudaMalloc(&ctx->sep_tab, tab_bytes));
            ctx->sep_tab_bytes = tab_bytes;
        }
        char* p = reinterpret_cast<char*>(ctx->sep_tab);
        VHR_CHECK_CUDA(ctx, cudaMemcpy(p, xw.data(), (size_t)Wp * 16, cudaMemcpyHostToDevice));
        VHR_CHECK_CUDA(ctx, cudaMemcpy(p + (size_t)Wp * 16, yw.data(), (size_t)H * 16, cudaMemcpyHostToDevice));
        VHR_CHECK_CUDA(ctx, cudaMemcpy(p + (size_t)Wp * 16 + (size_t)H * 16, xb.data(), (size_t)(Wp / 4) * 4, cudaMemcpyHostToDevice));
        VHR_CHECK_CUDA(ctx, cudaMemcpy(p + (size_t)Wp * 16 + (size_t)H * 16 + (size_t)(Wp / 4) * 4, yb.data(), (size_t)H * 4, cudaMemcpyHostToDevice));
        ctx->sep_key = key;
    }
    {
        char* p = reinterpret_cast<char*>(ctx->sep_tab);
        a.xw = reinterpret_cast<const float4*>(p);
        a.yw = reinterpret_cast<const float4*>(p + (size_t)Wp * 16);
        a.xb = reinterpret_cast<const int*>(p + (size_t)Wp * 16 + (size_t)H * 16);
        a.yb = reinterpret_cast<const int*>(p + (size_t)Wp * 16 + (size_t)H * 16 + (size_t)(Wp / 4) * 4);
    }
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 3) == 0) &&
                     (!d_out_f32 || (reinterpret_cast<uintptr_t>(d_out_f32) & 15) == 0) &&
                     (!d_out_u8 || (reinterpret_cast<uintptr_t>(d_out_u8) & 3) == 0);
    int load = vec ? 1 : 0;
    if (vec && W % 16 == 0 && (reinterpret_cast<uintptr_t>(d_frames) & 15) == 0) load = 2;
    if (const char* e = getenv("VHR_COLLAPSE_LOAD")) { const int v = atoi(e); if (v >= 0 && v < load) load = v; }

    const bool poly = Kg > 0 && roi.poly != nullptr;
    const int32_t* boxes = nullptr;       // rectangles or polygon bounding boxes of ROIs [k0, k0 + Kg), stride roi.K
    const long long* counts = nullptr;
    a.K = Kg; a.Kstride = roi.K;
    if (Kg > 0) {
        // scratch: [partials | polygon boxes (T,K,4) | polygon counts (T,K) | row masks (T,K,H,MW)]
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const int Kp = roi.K < KMAXF ? roi.K : KMAXF;            // every group of a call uses the same layout
        const size_t part_bytes = al(sizeof(double) * (size_t)T * a.nbands * a.nsegs * Kp * 3);
        const int MW = (W + 31) / 32;
        const size_t box_bytes = poly ? al(sizeof(int32_t) * 4 * (size_t)T * roi.K) : 0;
        const size_t cnt_bytes = poly ? al(sizeof(long long) * (size_t)T * roi.K) : 0;
        const size_t mask_bytes = poly ? al(sizeof(uint32_t) * (size_t)T * roi.K * H * MW) : 0;
        VHR_REQUIRE(ctx, mask_bytes / 4 < 0xFFFFFFFFull, "polygon row masks exceed 2^32 words: split the clip along T");
        void* p = nullptr;
        int rc = vhr_scratch(ctx, part_bytes + box_bytes + cnt_bytes + mask_bytes, &p);
        if (rc != VHR_OK) return rc;
        char* base = reinterpret_cast<char*>(p);
        a.partial = reinterpret_cast<double*>(base);
        if (poly) {
            int32_t* bx = reinterpret_cast<int32_t*>(base + part_bytes);
            long long* cn = reinterpret_cast<long long*>(base + part_bytes + box_bytes);
            uint32_t* mk = reinterpret_cast<uint32_t*>(base + part_bytes + box_bytes + cnt_bytes);
            // rasterise every polygon of the call once (k0 == 0 call; later groups of the same call reuse the scratch)
            if (k0 == 0) {
                rc = vhr_poly_rowmask(ctx, T, H, W, roi.poly, roi.nvert, roi.K, roi.Vmax, mk, MW, bx, cn, stream);
                if (rc != VHR_OK) return rc;
            }
            boxes = bx + 4 * k0; counts = cn + k0;
            a.mask = mk + (size_t)k0 * H * MW; a.MW = MW;
        } else {
            boxes = roi.rects + 4 * k0;
        }
        a.rects = boxes;
    }
    int rc;
    if (Kg == 0) rc = launch_sep<0, 0>(ctx, a, load, stream);
    else if (!poly) rc = (Kg == 1) ? launch_sep<1, 1>(ctx, a, load, stream) : launch_sep<KMAXF, 1>(ctx, a, load, stream);
    else rc = (Kg == 1) ? launch_sep<1, 2>(ctx, a, load, stream) : launch_sep<KMAXF, 2>(ctx, a, load, stream);
    if (rc != VHR_OK) return rc;
    if (Kg > 0) {
        const int n = T * Kg * 3;
        roi_finalize_sep_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a.partial, boxes, roi.K, counts, T, Kg, H, W, a.BH, a.nbands,
                                                                    a.nsegs, roi.mean + 3 * k0);
        rc = vhr_after_launch(ctx, "roi_finalize_sep_kernel");
        if (rc == VHR_OK && poly && roi.count && k0 + Kg >= roi.K) {
            VHR_CHECK_CUDA(ctx, cudaMemcpyAsync(roi.count, counts - k0, sizeof(long long) * (size_t)T * roi.K, cudaMemcpyDeviceToDevice, stream));
        }
    }
    return rc;
}

int collapse_impl(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames, int T, int H, int W, int levels,
                  float* d_out_f32, uint8_t* d_out_u8, const RoiSpec& roi, cudaStream_t stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_level && d_frames, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, levels >= 1 && levels <= VHR_MAX_LEVELS, "levels must be 1..6");
    VHR_REQUIRE(ctx, roi.K >= 0 && roi.K <= VHR_MAX_ROIS, "fused ROI count must be 0..8");
    VHR_REQUIRE(ctx, roi.K == 0 || ((roi.rects || (roi.poly && roi.nvert)) && roi.mean), "ROI pointers missing");
    VHR_REQUIRE(ctx, !roi.poly || (roi.Vmax >= 1 && roi.Vmax <= VHR_MAX_POLY_VERTS), "Vmax must be 1..64");
    VHR_REQUIRE(ctx, !roi.poly || W <= 8192, "polygon ROIs: frames up to 8192 pixels wide");
    VHR_REQUIRE(ctx, d_out_f32 || d_out_u8 || roi.K > 0, "nothing to compute");
    int rc = vhr_enter(ctx, stream);       // composite tables + scratch arena are context-owned
    if (rc != VHR_OK) return rc;
    // ROIs in groups of KMAXF: the first group rides on the frame-output pass, the others are ROI-only passes
    // (items outside their ROIs retire at once).  All groups share one scratch layout, so a later group must not
    // start before the previous finalize has read its partials: same stream, in order.
    int k0 = 0;
    do {
        const int Kg = roi.K - k0 < KMAXF ? roi.K - k0 : KMAXF;
        rc = collapse_group(ctx, d_level, d_frames, T, H, W, levels, k0 == 0 ? d_out_f32 : nullptr, k0 == 0 ? d_out_u8 : nullptr,
                            roi, k0, Kg, stream);
        k0 += KMAXF;
    } while (rc == VHR_OK && k0 < roi.K);
    return vhr_leave(ctx, stream, rc);
}

}  // namespace

extern "C" int vhr_collapse_addback_roi(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames, int T, int H, int W,
                                        int levels, float* d_out_f32, uint8_t* d_out_u8, const int32_t* d_rects, int K,
                                        double* d_roi_mean, void* stream) {
    RoiSpec roi;
    memset(&roi, 0, sizeof(roi));
    roi.rects = d_rects; roi.K = K; roi.mean = d_roi_mean;
    if (ctx && K > 0 && !d_rects) { vhr_set_error(ctx, "vhr_collapse_addback_roi: ROI pointers missing"); return VHR_ERR_INVALID; }
    return collapse_impl(ctx, d_level, d_frames, T, H, W, levels, d_out_f32, d_out_u8, roi, (cudaStream_t)stream);
}

extern "C" int vhr_collapse_addback_poly(vhr_ctx* ctx, const float* d_level, const uint8_t* d_frames, int T, int H, int W,
                                         int levels, float* d_out_f32, uint8_t* d_out_u8, const int32_t* d_poly,
                                         const int32_t* d_nvert, int K, int Vmax, double* d_roi_mean, int64_t* d_count,
                                         void* stream) {
    RoiSpec roi;
    memset(&roi, 0, sizeof(roi));
    roi.poly = d_poly; roi.nvert = d_nvert; roi.K = K; roi.Vmax = Vmax; roi.mean = d_roi_mean; roi.count = d_count;
    if (ctx && (K < 1 || !d_poly || !d_nvert)) { vhr_set_error(ctx, "vhr_collapse_addback_poly: polygon pointers missing"); return VHR_ERR_INVALID; }
    return collapse_impl(ctx, d_level, d_frames, T, H, W, levels, d_out_f32, d_out_u8, roi, (cudaStream_t)stream);
}
