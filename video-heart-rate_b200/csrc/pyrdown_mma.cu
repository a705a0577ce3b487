// Streaming pyrDown cascade, integer-tensor-core form, ONE WARP = ONE PRIVATE PIPELINE (W % 16 == 0 and
// W % 2^levels == 0, levels <= 5, 16-byte aligned frames; other shapes take pyrdown_stream.cu / pyrdown.cu).
// Same spec as pyrdown.cu (cv2.pyrDown float32 semantics; levels 1-2 exact integers * 2^-8l).
//
// What the ncu captures of the previous kernels said (profiles/README.md, round 2):
//   * pyrdown_stream.cu is issue-bound in its main loop (~220 instructions per level-1 row and warp, two thirds
//     of them byte unpacking, IDP2A dot products and shuffles of the HORIZONTAL 5-tap passes), and
//   * behind that, BOTH it and the first tensor-core cut were bound by the serial upper-level chain: levels >= 3
//     of a level-2 row were run by one warp, each row waiting for the previous one (~4 300 - 7 000 cycles per
//     row and CTA whatever the main loop cost; a third of the stall samples sat in mbarrier spins).
// Hence this design:
//   * horizontal passes of levels 1 and 2 on the integer tensor cores: the 5-tap, stride-2, 3-channel-interleaved
//     filter is a banded matrix, evaluated on the raw uint8 row bytes by mma.sync.m16n8k32 (u8 x u8 -> s32, SASS
//     IMMA.16832).  One MMA column = one block of 16 consecutive output values (interleaved channel bytes), whose
//     45-byte input window sits inside K = 64 = two k-steps; the 8 columns of an MMA are blocks 3 apart, so that
//     all of them have the same channel phase and share one constant weight fragment (three phases, 24 registers
//     of constants built once per thread).  A lane feeds the MMA with plain 8-byte shared-memory loads of the row
//     (the K order of the fragment is permuted to make them contiguous): no unpacking, no shuffles, exact sums;
//   * vertical passes on the accumulators: level 1 packed two 16-bit values per register (<= 65280) in the
//     incremental form out = A + 4 n1 + n2, A' = C + 4 n1 + 6 n2, C' = n2; a finished level-1 row is split into a
//     low-byte and a high-byte plane in a private shared-memory strip and goes through the SAME banded MMA twice
//     (level 1 is 3-channel interleaved too; lo + 256 hi recombine exactly in s32); level-2 vertical pass
//     incremental with two s32 partial sums per value;
//   * NO cross-warp communication at all.  A warp owns a strip of the FINAL level (12 px of level 4 at 4 levels)
//     and computes, privately, everything that strip depends on: 60 px of level 2 (for 48 owned), 128 px of level 1,
//     a 832-byte window of every input row that it fetches itself (TMA tensor box into its own ring).  Levels >= 3
//     run inline in the same warp, one lane per pixel, from small per-warp strips / 5-row H rings.  The price is
//     25 % redundant work at the strip borders; what it buys: no barrier wider than a warp, no spin loop except
//     the wait for the warp's own input rows, no serial chain.  A CTA is only a container of such warps;
//   * frame borders (reflect-101 left / right) are patched into the shared-memory rows of each level (6 + 3
//     values per row, edge strips only) instead of being special-cased in the arithmetic.
// Levels 1-2 stay bit-exact (integer arithmetic throughout); levels >= 3 are float32 as before.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int HR = 5;        // rows in a private H ring (exactly the vertical footprint)
constexpr int WSLOT = 832;   // bytes of one input row in a warp's ring (24 blocks x 32 bytes + the 64-byte window of the last one)
constexpr int GBYTES = 2 * WSLOT;   // input rows travel in groups of two = one TMA box, a multiple of 128 bytes
constexpr int PLANE = 512;   // bytes of one byte plane (low / high) of a warp's level-1 row: 384 values + over-read room
constexpr int N2 = 60;       // level-2 pixels a warp computes (12 MMA blocks of 16 values; the last 12 values are not valid)
constexpr int MAXL = 5;      // deepest pyramid this kernel takes (the strip a warp can own shrinks as 60 -> 24 -> 12 -> 4 px)

// pixels of level l (>= 2) a warp computes: 60, 28, 12, 4
__host__ __device__ constexpr int ncomp(int l) { return l == 2 ? 60 : l == 3 ? 28 : l == 4 ? 12 : 4; }

struct MmaArgs {
    const uint8_t* frames;
    float* out;
    int T, H, W, levels;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    long long total_rows;
    int ng;                   // two-row groups in each warp's input ring
    int rowbytes;             // 3 W
    int strips, wpc, ncg;     // strips per row, warps (strips) per CTA, CTAs per row share (column groups)
    int own;                  // pixels of the final level a strip owns (level-1 pixels when levels == 1)
    int step[MAXL + 1];       // first computed pixel of level l (2..L) in strip s: step[l] * s + boff[l]
    int boff[MAXL + 1];
    int warp_smem;            // bytes of shared memory per warp: [ng mbarriers][input ring][planes][strips / H rings]
    int in_off, plane_off;    // offsets inside a warp's block
    int strip_off[MAXL + 1];  // level l (2..L-1): finished row, interleaved floats (3 per pixel)
    int hring_off[MAXL + 1];  // level l (3..L): HR rows of horizontal sums
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
#ifdef VHR_WATCHDOG
__device__ __noinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (long long spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > 2000000) {
            if ((threadIdx.x & 31) == 0)
                printf("WATCHDOG(mma) block %d warp %d bar_off %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar & 0xffffu, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
#endif
__device__ __forceinline__ void tensor_g2s(uint32_t dst, const CUtensorMap* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// D = A(16x32, u8, row) * B(32x8, u8, col) [+ D]; s32 accumulators (SASS: IMMA.16832.U8.U8).  The first k-step of a
// chain takes a zero C operand (RZ) instead of zero-filled accumulator registers.
__device__ __forceinline__ void imma0(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0));
}
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One register of the constant weight fragment (operand A of the MMA) of channel phase `phase`, k-step `ks`.
// Row m of the fragment produces output value jl = 2 m (m < 8) or 2 (m - 8) + 1 of a 16-value block, so that a
// lane's two accumulator rows (g, g + 8) are two ADJACENT values; logical k-slot 4 q + i (+16) of the fragment
// holds window byte 8 q + i (+4): a lane's b0 / b1 operand registers are then the two halves of ONE 8-byte load.
// Value jl (channel c = (phase + jl) % 3) of a block whose first value has global index S reads input bytes
// 2 (S + jl) - c + 3 (d - 2), d = 0..4, with weights 1 4 6 4 1; the window starts at byte 2 S - 8.
__device__ __forceinline__ uint32_t weight_reg(int phase, int ks, int r, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int row = g + 8 * (r & 1);
    const int jl = row < 8 ? 2 * row : 2 * (row - 8) + 1;
    const int c = (phase + jl) % 3;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int P = 8 * q + i + 4 * (r >> 1) + 32 * ks;      // window byte of this k-slot
        const int d3 = P - 8 - 2 * jl + c + 30;                // 3 (d - 2) + 30
        uint32_t wgt = 0;
        if (d3 % 3 == 0) {
            const int d = d3 / 3 - 10 + 2;
            if (d == 0 || d == 4) wgt = 1;
            else if (d == 1 || d == 3) wgt = 4;
            else if (d == 2) wgt = 6;
        }
        v |= wgt << (8 * i);
    }
    return v;
}

// reflect-101 patch of one row held in shared memory (bytes or floats, 3 interleaved channels): idx0 = index of
// pixel 0 / channel 0, idxe = index of pixel w (the first one right of the row); pixels -2, -1 <- 2, 1 and w <- w - 2.
// Lanes 0..5 / 8..10 do the copies; the caller orders them with __syncwarp.
template <typename Tv>
__device__ __forceinline__ void patch_row(Tv* row, int idx0, bool left, int idxe, bool right, int lane) {
    if (left && lane < 6) row[idx0 + lane - 6] = row[idx0 + (lane < 3 ? lane + 6 : lane)];
    if (right && lane >= 8 && lane < 11) row[idxe + lane - 8] = row[idxe - 6 + lane - 8];
}

template <int L>
struct Pipe {
    const MmaArgs& a;
    const CUtensorMap* tmap;
    unsigned char* wsm;       // this warp's shared-memory block
    const int lane, strip;
    uint32_t AF[3][2][4];     // weight fragments: [channel phase][k-step][register]
    const unsigned char* rd;  // this lane's first operand bytes in row 0 of the warp's input ring
    unsigned char* ring0;     // row 0 of the warp's input ring
    unsigned char* plane;     // the warp's level-1 planes (low bytes, then high bytes at + PLANE)
    unsigned char* pw;        // this lane's store position in the low plane
    const unsigned char* pl;  // this lane's operand bytes in the low plane
    // frame borders inside this warp's rows: index of pixel 0 (left border present if >= 6) / of pixel w (right, -1 if outside)
    int in0, ine, l10, l1e;
    bool inL, l1L;
    int B[MAXL + 1];          // first computed pixel of level l (2..L)
    int own_lo, own_hi;       // final-level pixels this warp stores: [own_lo, own_hi)
    uint32_t wbar;            // shared address of the warp's group-0 "rows landed" barrier
    // the warp's input ring: consumer side (all lanes) and producer side (lane 0); unit = group of two rows
    int c_g, c_phase, g_cons;
    int p_g, g_issued, g_total, vg0;
    int g_int0, g_int1, box_y0;                   // groups [g_int0, g_int1) lie inside the frame: one TMA box at row box_y0 + 2 G
    uint32_t ring_u32;                            // shared address of the warp's input ring
    int src_off, cp_bytes, dst_off, box_x;        // the warp's byte range of an input row
    const uint8_t* frame;
    float* out_frame;
    int nextr[MAXL + 1], lastr[MAXL + 1];         // next / last row of each level in the current segment
    int hslot[MAXL + 1];                          // newest row of the H ring of levels >= 3
    uint32_t VA[6], VC[6];                        // level-1 vertical pass (packed pairs): partial sum of the next row, last input row

    __device__ Pipe(const MmaArgs& a_, const CUtensorMap* tm, unsigned char* s, int strip_)
        : a(a_), tmap(tm), wsm(s), lane((int)(threadIdx.x & 31)), strip(strip_) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int ph = 0; ph < 3; ++ph)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int r = 0; r < 4; ++r) AF[ph][ks][r] = weight_reg(ph, ks, r, lane);
#pragma unroll
        for (int l = 2; l <= MAXL; ++l) B[l] = a.step[l] * strip + a.boff[l];
        // level-2 value index of the first computed value: O2 = 3 B2 (levels == 1: a virtual level 2 with B2 = 60 strip);
        // level-1 block J (0..23) = values [2 O2 - 8 + 16 J, + 16): window = input bytes [4 O2 - 24 + 32 J, + 64)
        //   = ring bytes [8 + 32 J, + 64) with ring byte b <-> input byte 4 O2 - 32 + b;  MMA group G takes J = 3 n + G.
        const int O2 = 3 * B[2];
        ring0 = wsm + a.in_off;
        rd = ring0 + 8 + 96 * g + 8 * q;
        plane = wsm + a.plane_off;
        pw = plane + 96 * q + 2 * g;                         // value pair (2 g, 2 g + 1) of block 3 (2 q + e) + G: + 48 e + 16 G
        pl = plane + 96 * (g & 3) + 8 * q;                   // level-2 block J2 = 3 (g & 3) + G: window = plane bytes [32 J2, + 64)
        wbar = smem_u32(wsm);
        c_g = 0; c_phase = 0; g_cons = 0; p_g = 0; g_issued = 0; g_total = 0; vg0 = 0;
        g_int0 = 0; g_int1 = 0; box_y0 = 0;
        ring_u32 = smem_u32(ring0);
        const int lo = 4 * O2 - 32;                          // input byte of ring byte 0 (a multiple of 16: B2 % 4 == 0)
        box_x = lo / 4;                                      // (uint32 elements; outside the row = zero-filled by the TMA unit)
        src_off = max(lo, 0);
        dst_off = src_off - lo;
        cp_bytes = min(a.rowbytes, lo + WSLOT) - src_off;
        in0 = -lo;          inL = in0 >= 6 && in0 + 9 <= WSLOT;
        ine = a.rowbytes - lo;
        if (ine < 6 || ine + 3 > WSLOT) ine = -1;
        const int L1base = 2 * O2 - 8;
        l10 = -L1base;      l1L = l10 >= 6 && l10 + 9 <= PLANE;
        l1e = 3 * a.w[1] - L1base;
        if (l1e < 6 || l1e + 3 > PLANE) l1e = -1;
        own_lo = a.own * strip;
        own_hi = min(own_lo + a.own, a.w[L]);
#pragma unroll
        for (int l = 0; l <= MAXL; ++l) { nextr[l] = 0; lastr[l] = -1; hslot[l] = 0; }
    }

    // ---- the warp's input ring ------------------------------------------------------------------
    // Group G of a segment = virtual input rows vg0 + 2G, vg0 + 2G + 1 (reflect-101 at the frame's
    // top / bottom).  Inside the frame the two rows are one TMA box (cp.async.bulk.tensor.2d on a
    // (T*H) x (3W/4) uint32 view of the clip; columns outside the row are zero-filled); at the
    // frame's edges they are two plain bulk copies.
    __device__ __forceinline__ void issue_group(int G) {
        const uint32_t bar = wbar + 8 * p_g;
        const uint32_t dst = ring_u32 + p_g * GBYTES;
        if (G >= g_int0 && G < g_int1) {
            mbar_expect_tx(bar, (uint32_t)GBYTES);
            tensor_g2s(dst, tmap, box_x, box_y0 + 2 * G, bar);
        } else {
            const int v = vg0 + 2 * G;
            mbar_expect_tx(bar, 2u * (uint32_t)cp_bytes);
#pragma unroll
            for (int r = 0; r < 2; ++r)
                bulk_g2s(dst + r * WSLOT + dst_off, frame + (size_t)vhr_reflect101(v + r, a.H) * a.rowbytes + src_off,
                         (uint32_t)cp_bytes, bar);
        }
        p_g = (p_g + 1 == a.ng) ? 0 : p_g + 1;
    }
    // Every group the warp has consumed so far has been read by all its lanes: lane 0 requests the
    // next groups into those slots.  (Edge strips patched border bytes into those slots through the
    // generic proxy: order those writes before the bulk copies that will overwrite them.)
    __device__ __forceinline__ void refill() {
        if (inL || ine >= 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const int lim = min(g_total, g_cons + a.ng);
            while (g_issued < lim) issue_group(g_issued++);
        }
    }
    // wait for the next group of the segment; returns the offset of its first row in the warp's ring
    __device__ __forceinline__ int next_group() {
        mbar_wait(wbar + 8 * c_g, (uint32_t)c_phase);
        const int off = c_g * GBYTES;
        if (inL || ine >= 0) {                       // warp-uniform: frame borders of the two input rows
            patch_row(ring0 + off, in0, inL, ine, ine >= 0, lane);
            patch_row(ring0 + off + WSLOT, in0, inL, ine, ine >= 0, lane);
            __syncwarp();
        }
        ++g_cons;
        if (++c_g == a.ng) { c_g = 0; c_phase ^= 1; }
        return off;
    }

    // ---- the banded horizontal pass ---------------------------------------------------------------
    // 24 blocks x 16 values of one row from the lane's operand bytes at `p` (+ 32 per block of the lane's column):
    // 4 x LDS.64 + 6 x IMMA; result = the lane's 12 values as 6 packed pairs (lo 16 bits: value 2 g, hi: 2 g + 1),
    // pk[2 G + e] = pair of block 3 (2 q + e) + G.  PH0 = channel phase of group 0.
    template <int PH0>
    __device__ __forceinline__ void hrow_packed(const unsigned char* p, uint32_t (&pk)[6]) {
        uint2 b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = *reinterpret_cast<const uint2*>(p + 32 * k);
#pragma unroll
        for (int G = 0; G < 3; ++G) {
            int d[4];
            imma0(d, AF[(PH0 + G) % 3][0], b[G].x, b[G].y);
            imma(d, AF[(PH0 + G) % 3][1], b[G + 1].x, b[G + 1].y);
            pk[2 * G] = (uint32_t)d[0] + ((uint32_t)d[2] << 16);
            pk[2 * G + 1] = (uint32_t)d[1] + ((uint32_t)d[3] << 16);
        }
    }

    // ---- level 1: vertical pass on the packed horizontal sums ----------------------------------------
    //     out_r = A + 4 n1 + n2,   A' = C + 4 n1 + 6 n2,   C' = n2
    // with n1, n2 the horizontal sums of the new input rows 2r+1, 2r+2, C = row 2r and A = row(2r-2) + 4 row(2r-1)
    // + 6 row(2r).  All sums are exact integers (level 1 = sum / 256 <= 65280: 16 bits per packed half).
    __device__ __forceinline__ void prime() {       // first three input rows of a segment
        int off = next_group();                     // (the first row of a segment's group 0 is a filler)
        uint32_t x0[6], x1[6];
        hrow_packed<1>(rd + off + WSLOT, x0);
        off = next_group();
        hrow_packed<1>(rd + off, x1);
        hrow_packed<1>(rd + off + WSLOT, VC);
#pragma unroll
        for (int k = 0; k < 6; ++k) VA[k] = x0[k] + (x1[k] << 2) + VC[k] * 6u;
        refill();
    }
    __device__ __forceinline__ void l1_row(uint32_t (&v)[6]) {
        const int off = next_group();
        uint32_t n1[6], n2[6];
        hrow_packed<1>(rd + off, n1);
        hrow_packed<1>(rd + off + WSLOT, n2);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t t = n1[k] << 2;
            v[k] = VA[k] + t + n2[k];
            VA[k] = VC[k] + t + n2[k] * 6u;
            VC[k] = n2[k];
        }
    }
    // ---- level 2, horizontal: the finished level-1 row through the same banded MMA, as two byte planes ----------
    __device__ __forceinline__ void l2_hrow(const uint32_t (&v)[6], int (&x)[12]) {
        __syncwarp();                                // every lane has loaded its operands of the previous row
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            unsigned char* d = pw + 48 * (k & 1) + 16 * (k >> 1);
            *reinterpret_cast<uint16_t*>(d) = (uint16_t)__byte_perm(v[k], 0u, 0x4420);            // low bytes of the pair
            *reinterpret_cast<uint16_t*>(d + PLANE) = (uint16_t)__byte_perm(v[k], 0u, 0x4431);    // high bytes
        }
        __syncwarp();
        if (l1L || l1e >= 0) {                       // frame borders of level 1 (both planes)
            patch_row(plane, l10, l1L, l1e, l1e >= 0, lane);
            patch_row(plane + PLANE, l10, l1L, l1e, l1e >= 0, lane);
            __syncwarp();
        }
        uint2 bl[4], bh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bl[k] = *reinterpret_cast<const uint2*>(pl + 32 * k);
            bh[k] = *reinterpret_cast<const uint2*>(pl + PLANE + 32 * k);
        }
#pragma unroll
        for (int G = 0; G < 3; ++G) {
            int dl[4], dh[4];
            imma0(dl, AF[G][0], bl[G].x, bl[G].y);
            imma(dl, AF[G][1], bl[G + 1].x, bl[G + 1].y);
            imma0(dh, AF[G][0], bh[G].x, bh[G].y);
            imma(dh, AF[G][1], bh[G + 1].x, bh[G + 1].y);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[4 * G + i] = dl[i] + (dh[i] << 8);     // < 2^20
        }
    }
    __device__ __forceinline__ void l12_row(int (&x)[12]) {
        uint32_t v[6];
        l1_row(v);
        l2_hrow(v, x);
    }

    __device__ __forceinline__ void begin_segment(int t, int r0, int r1) {
        frame = a.frames + (size_t)t * a.H * a.rowbytes;
        out_frame = a.out + (size_t)t * a.h[L] * a.w[L] * 3;
        int f = r0, e = r1 - 1;
        nextr[L] = f; lastr[L] = e;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            f = max(0, 2 * f - 2);
            e = min(a.h[l] - 1, 2 * e + 2);
            nextr[l] = f; lastr[l] = e;
        }
        // input rows of the segment: virtual rows 2*first-2 .. 2*last+2 in groups of two, after one filler row
        vg0 = 2 * nextr[1] - 3;
        g_total = (lastr[1] - nextr[1] + 1) + 2;
        g_issued = 0;
        g_cons = 0;
        g_int0 = vg0 < 0 ? (1 - vg0) / 2 : 0;                 // first G with vg0 + 2G >= 0
        g_int1 = (a.H - vg0) / 2;                            // first G with vg0 + 2G + 1 >= H  (vg0 odd)
        box_y0 = t * a.H + vg0;
    }

    // ---- levels >= 3, inline in the same warp: one lane per pixel (3 channels) ---------------------------------------
    // Row r of level l-1 sits in the warp's strip of that level (interleaved floats, pixel B[l-1] + i at float 3 i).
    // Horizontal pass of the warp's ncomp(l) pixels into the H ring of level l (HR rows), then every level-l row whose
    // five H rows are present is finished (reflect-101 at the frame's top / bottom), written to the strip of level l
    // (or, at the last level, to global memory: the owned pixels only) and handed to level l+1.
    template <int l>
    __device__ __forceinline__ void upper(int r) {
        constexpr int NL = ncomp(l);
        const int hp = a.h[l - 1], wl = a.w[l];
        int hs = hslot[l] + 1;
        if (hs == HR) hs = 0;
        float* const hring = reinterpret_cast<float*>(wsm + a.hring_off[l]);
        const float* const src = reinterpret_cast<const float*>(wsm + a.strip_off[l - 1]);
        // pixel B[l] + lane of level l reads pixels 2 (B[l] + lane) - 2 + d of level l-1 = strip pixel 2 lane + d + off
        const int off = 2 * B[l] - 2 - B[l - 1];
        if (lane < NL) {
            const float* p = src + 3 * (2 * lane + off);
            float o[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                o[ch] = p[6 + ch] * 6.0f + (p[3 + ch] + p[9 + ch]) * 4.0f + p[ch] + p[12 + ch];
            float* hd = hring + (hs * NL + lane) * 3;
            hd[0] = o[0]; hd[1] = o[1]; hd[2] = o[2];
        }
        int nx = nextr[l];
        const int lst = lastr[l];
        __syncwarp();
        while (nx <= lst && min(2 * nx + 2, hp - 1) <= r) {
            const int q = nx;
            int so[5];                              // H-ring rows of the five source rows: row r' sits (r - r') slots behind the newest
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                int sl = hs - (r - vhr_reflect101(2 * q - 2 + d, hp));
                if (sl < 0) sl += HR;
                so[d] = sl * NL * 3;
            }
            if (lane < NL) {
                const float* hc = hring + 3 * lane;
                float v[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    v[ch] = (hc[so[2] + ch] * 6.0f + (hc[so[1] + ch] + hc[so[3] + ch]) * 4.0f + hc[so[0] + ch] + hc[so[4] + ch]) * (1.0f / 256.0f);
                if constexpr (l == L) {
                    const int px = B[l] + lane;
                    if (px >= own_lo && px < own_hi) {
                        float* o = out_frame + ((size_t)q * wl + px) * 3;
                        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
                    }
                } else {
                    float* sd = reinterpret_cast<float*>(wsm + a.strip_off[l]) + 3 * lane;
                    sd[0] = v[0]; sd[1] = v[1]; sd[2] = v[2];
                }
            }
            nx = q + 1;
            if constexpr (l < L) {
                __syncwarp();
                const int i0 = -3 * B[l], ie = 3 * (wl - B[l]);
                const bool lf = i0 >= 6 && i0 + 9 <= 3 * NL, rt = ie >= 6 && ie + 3 <= 3 * NL;
                if (lf || rt) {                      // frame borders of level l
                    patch_row(reinterpret_cast<float*>(wsm + a.strip_off[l]), i0, lf, ie, rt, lane);
                    __syncwarp();
                }
                upper<l + 1>(q);
            }
        }
        __syncwarp();
        hslot[l] = hs;
        nextr[l] = nx;
    }

    // A finished level-2 row: this lane holds 12 sums (exact, < 2^24) of which the lanes with q < 2 are real (the
    // MMA columns 4..7 of level 2 are duplicates): value s[4 G + e + 2 h] = level-2 value 96 q + 48 e + 16 G + 2 g + h
    // of the warp's 192-value window (the first 180 = 60 pixels are valid).
    __device__ __forceinline__ void emit2(int q, const int (&s)[12]) {
        float* dst;
        int jlo, jhi;                                // values of the window to store
        if constexpr (L == 2) {
            dst = out_frame + ((size_t)q * a.w[2] + B[2]) * 3;
            jlo = 3 * (own_lo - B[2]); jhi = 3 * (own_hi - B[2]);
        } else {
            dst = reinterpret_cast<float*>(wsm + a.strip_off[2]);
            jlo = 0; jhi = 3 * N2;
        }
        const int j0 = 96 * (lane & 3) + 2 * (lane >> 2);
        if ((lane & 3) < 2) {
#pragma unroll
            for (int G = 0; G < 3; ++G)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = j0 + 48 * e + 16 * G;
                    if (j >= jlo && j < jhi)
                        *reinterpret_cast<float2*>(dst + j) = make_float2((float)s[4 * G + e] * (1.0f / 65536.0f),
                                                                          (float)s[4 * G + e + 2] * (1.0f / 65536.0f));
                }
        }
        nextr[2] = q + 1;
        if constexpr (L >= 3) {
            __syncwarp();
            const int i0 = -3 * B[2], ie = 3 * (a.w[2] - B[2]);
            const bool lf = i0 >= 6 && i0 + 9 <= 3 * N2, rt = ie >= 6 && ie + 3 <= 3 * N2;
            if (lf || rt) {                          // frame borders of level 2
                patch_row(dst, i0, lf, ie, rt, lane);
                __syncwarp();
            }
            upper<3>(q);
        }
    }

    // ---- one segment ----------------------------------------------------------------------------
    __device__ __forceinline__ void run_segment() {
        refill();
        prime();
        if constexpr (L == 1) {
            int n = 0;
            // the lane's pairs: level-1 value 2 O2 - 8 + 96 q + 48 e + 16 G + 2 g (+1); stored if its pixel is owned
            const int v0 = 6 * B[2] - 8 + 96 * (lane & 3) + 2 * (lane >> 2);
            for (int r = nextr[1]; r <= lastr[1]; ++r) {
                uint32_t v[6];
                l1_row(v);
                float* o = out_frame + (size_t)r * a.w[1] * 3;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int j = v0 + 48 * (k & 1) + 16 * (k >> 1);
                    if (j >= 3 * own_lo && j < 3 * own_hi)
                        *reinterpret_cast<float2*>(o + j) = make_float2((float)(v[k] & 0xFFFFu) * (1.0f / 256.0f),
                                                                        (float)(v[k] >> 16) * (1.0f / 256.0f));
                }
                if (++n == 2) { n = 0; refill(); }
            }
        } else {
            const int h1 = a.h[1];
            int q = nextr[2];
            const int ql = lastr[2];
            // level-2 vertical pass, incremental: A = r(2q-2) + 4 r(2q-1) + 6 r(2q), C = r(2q)  (r = the level-2 horizontal
            // sums of a level-1 row); rows 0 .. 2 prime it at the top of a frame.  Bottom border: the rows that reflect
            // are 4 r(2q-1) + r(2q-2) = A - 6 C.
            int A[12], C[12];
            {
                int x0[12], x1[12];
                l12_row(x0);
                refill();
                l12_row(x1);
                refill();
                l12_row(C);
                refill();
                if (q == 0) {                        // rows -2,-1 reflect to 2,1
                    int s[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) s[k] = 6 * x0[k] + 8 * x1[k] + 2 * C[k];
                    emit2(0, s);
                    q = 1;
                }
#pragma unroll
                for (int k = 0; k < 12; ++k) A[k] = x0[k] + 4 * x1[k] + 6 * C[k];
            }
#pragma unroll 1
            for (; q <= ql; ++q) {
                const bool has1 = 2 * q + 1 <= h1 - 1, has2 = 2 * q + 2 <= h1 - 1;
                int s[12];
                if (has2) {                          // (has2 implies has1)
                    // n1 is consumed before the second row is built: at most two 12-value sets are live at a time
                    {
                        int n1[12];
                        l12_row(n1);
#pragma unroll
                        for (int k = 0; k < 12; ++k) {
                            s[k] = A[k] + 4 * n1[k];
                            A[k] = 4 * n1[k] + C[k];
                        }
                    }
                    l12_row(C);
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        s[k] += C[k];
                        A[k] += 6 * C[k];
                    }
                } else if (has1) {                   // row 2q+2 = h1 reflects to 2q
                    int n1[12];
                    l12_row(n1);
#pragma unroll
                    for (int k = 0; k < 12; ++k) s[k] = A[k] + 4 * n1[k] + C[k];
                } else {                             // rows 2q+1, 2q+2 reflect to 2q-1, 2q-2
#pragma unroll
                    for (int k = 0; k < 12; ++k) s[k] = 2 * A[k] - 6 * C[k];
                }
                refill();
                emit2(q, s);
            }
        }
    }
};

template <int L>
__global__ void __launch_bounds__(256, 2) pyrdown_mma_kernel(const MmaArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int cg = (int)(blockIdx.x % a.ncg), share = (int)(blockIdx.x / a.ncg), nshares = (int)(gridDim.x / a.ncg);
    const int strip = cg * a.wpc + warp;
    if (strip >= a.strips || share >= nshares) return;           // (warps never synchronise with one another)
    const long long lo = a.total_rows * share / nshares;
    const long long hi = a.total_rows * (share + 1) / nshares;
    if (lo >= hi) return;
    unsigned char* wsm = smem + (size_t)warp * a.warp_smem;
    if (lane == 0) {
        for (int b = 0; b < a.ng; ++b) mbar_init(smem_u32(wsm) + 8 * b, 1);                  // rows landed, per group
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the level-1 planes are read 64 bytes per block window: clear them once so that never-written tail bytes are defined
    for (int i = lane; i < 2 * PLANE / 4; i += 32) reinterpret_cast<uint32_t*>(wsm + a.plane_off)[i] = 0u;
    __syncwarp();
    Pipe<L> st(a, &tmap, wsm, strip);
    const int hL = a.h[L];
    long long pos = lo;
    while (pos < hi) {
        const int t = (int)(pos / hL);
        const int r0 = (int)(pos - (long long)t * hL);
        const long long frame_end = (long long)(t + 1) * hL;
        const int r1 = (int)((hi < frame_end ? hi : frame_end) - (long long)t * hL);
        st.begin_segment(t, r0, r1);
        st.run_segment();
        pos += r1 - r0;
    }
}

template <int L>
int launch_mma(vhr_ctx* ctx, const MmaArgs& a, const CUtensorMap& tmap, int threads, int smem_bytes, cudaStream_t stream) {
    auto kern = pyrdown_mma_kernel<L>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    VHR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes));
    if (per_sm < 1) return VHR_ERR_UNSUPPORTED;
    long long resident = (long long)per_sm * ctx->num_sms;
    long long nshares = resident / a.ncg;
    if (nshares < 1) nshares = 1;
    if (nshares > a.total_rows) nshares = a.total_rows;
    kern<<<(unsigned)(nshares * a.ncg), threads, smem_bytes, stream>>>(a, tmap);
    return vhr_after_launch(ctx, "pyrdown_mma_kernel");
}

}  // namespace

// Returns VHR_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible.
int vhr_pyrdown_mma(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                    cudaStream_t stream) {
    if (levels > MAXL || W % 16 != 0 || W % (1 << levels) != 0 || (reinterpret_cast<uintptr_t>(d_frames) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(d_level) & 15) != 0)
        return VHR_ERR_UNSUPPORTED;
    MmaArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = d_frames; a.out = d_level; a.T = T; a.H = H; a.W = W; a.levels = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    if (levels == 1 ? H < 2 : a.h[1] < 3) return VHR_ERR_UNSUPPORTED;   // the vertical passes are primed with three rows
    if (H < 8) return VHR_ERR_UNSUPPORTED;                               // reflected filler rows reach 5 rows outside the frame
    if (levels >= 2 && a.w[levels - 1] < 4) return VHR_ERR_UNSUPPORTED;  // border patches assume single reflections (px -2 <- 2, px w <- w - 2)
    a.total_rows = (long long)T * a.h[levels];
    a.rowbytes = 3 * W;
    // strip geometry: a warp computes 60 px of level 2 from pixel B2 = step2 * strip + boff2 (B2 % 4 == 0 keeps the input
    // window 16-byte aligned), 28 / 12 / 4 px of levels 3 / 4 / 5 from B_{l+1} = ceil((B_l + 2) / 2), and owns
    // `own` pixels of the last level (level-1 / level-2 pixels for levels 1 / 2: nothing is recomputed there)
    static const int OWN[MAXL + 1] = {0, 120, 60, 24, 12, 4};
    static const int STEP2[MAXL + 1] = {0, 60, 60, 48, 48, 32};
    static const int BOFF2[MAXL + 1] = {0, 0, 0, -4, -8, -16};
    a.own = OWN[levels];
    a.step[2] = STEP2[levels];
    a.boff[2] = BOFF2[levels];
    for (int l = 3; l <= MAXL; ++l) {
        a.step[l] = a.step[l - 1] / 2;
        a.boff[l] = (a.boff[l - 1] + 3 + 1000) / 2 - 500;            // ceil((boff + 2) / 2), also for negative boff (the step is even)
    }
    for (int l = 3; l <= levels; ++l) {                              // the owned pixels must be inside the computed ones
        const int lo_l = a.boff[l], hi_l = a.boff[l] + ncomp(l);
        if (l == levels && (lo_l > 0 || hi_l < a.own)) return VHR_ERR_UNSUPPORTED;
        (void)lo_l; (void)hi_l;
    }
    a.strips = (a.w[levels] + a.own - 1) / a.own;
    // a CTA is a container of independent warps: 4..8 strips, as few idle warps in the last column group as possible
    int best_wpc = 8, best_waste = 1 << 30;
    for (int wpc = 8; wpc >= 4; --wpc) {
        const int ncg = (a.strips + wpc - 1) / wpc;
        const int waste = ncg * wpc - a.strips;
        if (waste < best_waste) { best_waste = waste; best_wpc = wpc; }
    }
    if (a.strips < 4) best_wpc = a.strips;
    a.wpc = best_wpc;
    a.ncg = (a.strips + a.wpc - 1) / a.wpc;
    const int threads = a.wpc * 32;
    // per-warp shared memory
    auto al16 = [](int v) { return (v + 15) & ~15; };
    auto al128 = [](int v) { return (v + 127) & ~127; };
    int fixed = al16(2 * PLANE);
    for (int l = 2; l < levels; ++l) fixed = al16(fixed + 3 * ncomp(l) * 4 + 16);
    for (int l = 3; l <= levels; ++l) fixed = al16(fixed + HR * 3 * ncomp(l) * 4);
    int ng = 6;                                                       // ring depth: as deep as 16 warps per SM (128 registers each) allow
    const int ctas = 16 / a.wpc;                                      // CTAs per SM by registers
    const int budget = 233472 / ctas - 1024 - 256;                    // 228 KB per SM, 1 KB reserved per CTA
    while (ng > 3 && a.wpc * al128(128 + ng * GBYTES + fixed) > budget) --ng;
    a.ng = ng;
    a.in_off = 128;                                                   // (mbarriers in the first 128 bytes)
    a.plane_off = a.in_off + ng * GBYTES;
    int off = al16(a.plane_off + 2 * PLANE);
    for (int l = 2; l < levels; ++l) { a.strip_off[l] = off; off = al16(off + 3 * ncomp(l) * 4 + 16); }
    for (int l = 3; l <= levels; ++l) { a.hring_off[l] = off; off = al16(off + HR * 3 * ncomp(l) * 4); }
    a.warp_smem = al128(off);
    const int smem_bytes = a.wpc * a.warp_smem;
    if (smem_bytes > ctx->smem_optin) return VHR_ERR_UNSUPPORTED;
    // the clip as a 2-D uint32 tensor: (T*H) rows x (3W/4) elements; box = one warp's two-row group
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return VHR_ERR_UNSUPPORTED;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)(a.rowbytes / 4), (cuuint64_t)T * (cuuint64_t)H};
    const cuuint64_t gstr[1] = {(cuuint64_t)a.rowbytes};
    const cuuint32_t box[2] = {WSLOT / 4, 2};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(d_frames), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return VHR_ERR_UNSUPPORTED;
    switch (levels) {
        case 1: return launch_mma<1>(ctx, a, tmap, threads, smem_bytes, stream);
        case 2: return launch_mma<2>(ctx, a, tmap, threads, smem_bytes, stream);
        case 3: return launch_mma<3>(ctx, a, tmap, threads, smem_bytes, stream);
        case 4: return launch_mma<4>(ctx, a, tmap, threads, smem_bytes, stream);
        case 5: return launch_mma<5>(ctx, a, tmap, threads, smem_bytes, stream);
    }
    return VHR_ERR_INVALID;
}
