// Streaming pyrDown cascade, integer-tensor-core form (W % 16 == 0; W % 64 == 0 for >= 3 levels; 16-byte
// aligned frames).  Same spec as pyrdown.cu (cv2.pyrDown float32 semantics; levels 1-2 exact integers * 2^-8l),
// same streaming skeleton as pyrdown_stream.cu (persistent grid over (frame, final-row) shares, one private
// TMA input ring per warp, levels >= 3 by one warp per level-2 row in turn) -- what changes is where the
// instructions of levels 1 and 2 go.  The ncu captures of the previous kernel (profiles/r1_ncu_p_*) showed it
// issue-bound at 39 % of DRAM peak: ~220 instructions per level-1 row and warp, two thirds of them byte
// unpacking (PRMT), dot products (IDP2A) and shuffles of the HORIZONTAL 5-tap passes.  Here
//
//   * the horizontal 5-tap, stride-2, 3-channel-interleaved filter is a banded matrix, evaluated on the raw
//     uint8 row bytes by mma.sync.m16n8k32 (u8 x u8 -> s32, SASS IMMA.16832): one MMA column = one block
//     of 16 consecutive output values (interleaved channel bytes), whose 45-byte input window sits inside
//     K = 64 = two k-steps; the 8 columns of an MMA are blocks 3 apart, so that all of them have the same
//     channel phase and share one constant weight fragment (three phases, 24 registers of constants built
//     once per thread).  A lane feeds the MMA with plain 8-byte shared-memory loads of the row (the K order
//     of the fragment is permuted to make them contiguous) -- no unpacking, no shuffles, exact s32 sums;
//   * the vertical pass runs on the accumulators, packed two 16-bit values per register (<= 65280), in the
//     incremental form out = A + 4 n1 + n2, A' = C + 4 n1 + 6 n2, C' = n2 of the previous kernel;
//   * a finished level-1 row (16-bit values) is split into a low-byte and a high-byte plane in a private
//     shared-memory strip and goes through the SAME banded MMA twice (the weights are the same: level 1 is
//     3-channel interleaved too); lo + 256 hi recombine exactly in s32; level-2 vertical pass incremental
//     with three s32 partial sums per value (A, B = 4 r(2q-1) + r(2q-2) for the bottom border, C);
//   * a warp owns 60 level-2 pixels and RECOMPUTES the 8 level-1 values either side that its level-2 window
//     needs (384 level-1 values computed for 360 owned), so warps still never exchange level-1 data;
//   * frame borders (reflect-101 on the left / right) are patched into the shared-memory rows (6 + 3 bytes
//     per row, edge warps only) instead of being special-cased in the arithmetic.
//
// Level 2 rows go to the shared level-2 ring (interleaved RGB floats now) and levels >= 3 proceed exactly as
// in pyrdown_stream.cu.  Bit-exactness of levels 1-2 is unchanged (integer arithmetic throughout).
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int HR = 5;        // rows in a private H ring (exactly the vertical footprint)
constexpr int LANES = 30;    // 8-pixel input column groups owned by a warp (= 240 input px = 120 px of level 1 = 60 px of level 2)
constexpr int WSLOT = 832;   // bytes of one input row in a warp's ring: ring byte b <-> input byte 720 w - 32 + b
constexpr int GBYTES = 2 * WSLOT;   // input rows travel in groups of two = one TMA box, a multiple of 128 bytes
constexpr int RS = 4;        // level-2 ring slots = rows a warp may run ahead of the upper levels
constexpr int DLY = 2;       // the upper levels of row n start when their warp has published row n + DLY
constexpr int PLANE = 512;   // bytes of one byte plane (low / high) of a warp's level-1 row: 384 values + over-read room
constexpr int OWN2 = 6 * LANES;   // level-2 values (interleaved channel bytes) owned by a warp: 60 px x 3

struct MmaArgs {
    const uint8_t* frames;
    float* out;
    int T, H, W, levels;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    long long total_rows;
    int nt;                                // column groups = W / 8
    int ng;                                // two-row groups in each warp's input ring
    int rowbytes;                          // 3 W
    int in_off;                            // byte offset of warp 0's input ring (warp w: + w * ng * GBYTES), 128-B aligned
    int plane_off;                         // byte offset of warp 0's level-1 planes (warp w: + w * 2 * PLANE)
    int rbar_off;                          // byte offset of the RS row barriers, followed by the RS duty barriers
    int ring_off[VHR_MAX_LEVELS + 1];      // level 2: RS rows, interleaved (px + 2) * 3 + ch; levels 3..L-1: planar, double-buffered
    int ring_stride[VHR_MAX_LEVELS + 1];   // floats per row (level 2) / per channel plane row (levels >= 3)
    int hring_off[VHR_MAX_LEVELS + 1];     // levels 3..L: H rings (HR rows x 3 planes x w[l])
    int duty_off;                          // DutyState
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
// D += A(16x32, u8, row) * B(32x8, u8, col); s32 accumulators (SASS: IMMA.16832.U8.U8)
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// vector loads / stores of C consecutive floats (C = 4, 2, 1), naturally aligned
template <int C>
__device__ __forceinline__ void ldv(const float* p, float* x) {
    if constexpr (C == 4) { const float4 v = *reinterpret_cast<const float4*>(p); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    else if constexpr (C == 2) { const float2 v = *reinterpret_cast<const float2*>(p); x[0] = v.x; x[1] = v.y; }
    else x[0] = *p;
}
template <int C>
__device__ __forceinline__ void stv(float* p, const float* x) {
    if constexpr (C == 4) *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    else if constexpr (C == 2) *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]);
    else *p = x[0];
}

// Block-shared state of the levels >= 3 (they are processed by one warp at a time, in turn).
struct DutyState {
    int nextr[VHR_MAX_LEVELS + 1];
    int lastr[VHR_MAX_LEVELS + 1];
    int hslot[VHR_MAX_LEVELS + 1];
};

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

template <int L>
struct Stream {
    const MmaArgs& a;
    const CUtensorMap* tmap;
    unsigned char* smem;
    const int warp, lane;
    uint32_t AF[3][2][4];     // weight fragments: [channel phase][k-step][register]
    const unsigned char* rd;  // this lane's first operand bytes in row 0 of the warp's input ring
    unsigned char* ring0;     // row 0 of the warp's input ring
    unsigned char* plane;     // the warp's level-1 planes (low bytes, then high bytes at + PLANE)
    unsigned char* pw;        // this lane's store position in the low plane
    const unsigned char* pl;  // this lane's operand bytes in the low plane
    int e_in, e_l1;           // ring / plane byte index of the first byte right of the row end (pixel w: reflected from w - 2), or -1
    int own2;                 // level-2 values this warp publishes (<= OWN2; 0 if none)
    bool last2;               // this warp holds the right end of the level-2 row
    uint32_t bar0;            // shared address of mbarrier 0
    uint32_t wbar;            // shared address of the warp's group-0 "rows landed" barrier
    // the warp's input ring: consumer side (all lanes) and producer side (lane 0); unit = group of two rows
    int c_g, c_phase, g_cons;
    int p_g, g_issued, g_total, vg0;
    int g_int0, g_int1, box_y0;                   // groups [g_int0, g_int1) lie inside the frame: one TMA box at row box_y0 + 2 G
    uint32_t ring_u32;                            // shared address of the warp's input ring
    int src_off, cp_bytes, dst_off, box_x;        // the warp's byte range of an input row
    // Rows of level 2 are numbered across segments (dn = rows published so far, the same in every
    // warp).  Row barrier n % RS, phase n / RS: every warp has written its part of row n.
    // Duty barrier n % RS, phase n / RS: the upper levels have consumed row n.
    int dn, seg_n0, seg_q0;
    int duty_m, turn_w, turn_c;        // next row handed to run_duty, and the warp whose turn it is
    const uint8_t* frame;
    float* out_frame;
    int nextr[3], lastr[3];            // levels 1, 2 (levels >= 3: DutyState)
    int seg_next[VHR_MAX_LEVELS + 1], seg_last[VHR_MAX_LEVELS + 1];
    uint32_t VA[6], VC[6];             // level-1 vertical pass (packed pairs): partial sum of the next row, last input row

    __device__ Stream(const MmaArgs& a_, const CUtensorMap* tm, unsigned char* s)
        : a(a_), tmap(tm), smem(s), warp((int)(threadIdx.x >> 5)), lane((int)(threadIdx.x & 31)) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int ph = 0; ph < 3; ++ph)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int r = 0; r < 4; ++r) AF[ph][ks][r] = weight_reg(ph, ks, r, lane);
        ring0 = smem + a.in_off + warp * a.ng * GBYTES;
        // level-1 block J (0..23) of the warp = values [360 w - 8 + 16 J, + 16): window = ring bytes [8 + 32 J, + 64);
        // MMA group G (0..2) takes the blocks J = 3 n + G in its columns n = g.
        rd = ring0 + 8 + 96 * g + 8 * q;
        plane = smem + a.plane_off + warp * 2 * PLANE;
        pw = plane + 96 * q + 2 * g;                         // value pair (2 g, 2 g + 1) of block 3 (2 q + e) + G: + 48 e + 16 G
        pl = plane + 96 * (g & 3) + 8 * q;                   // level-2 block J2 = 3 (g & 3) + G: window = plane bytes [32 J2, + 64)
        bar0 = smem_u32(smem);
        wbar = bar0 + 8 * warp * a.ng;
        c_g = 0; c_phase = 0; g_cons = 0; p_g = 0; g_issued = 0; g_total = 0; vg0 = 0;
        g_int0 = 0; g_int1 = 0; box_y0 = 0;
        ring_u32 = smem_u32(ring0);
        // ring-row byte b of the warp <-> byte 24 * LANES * warp - 32 + b of the input row
        const int lo = 24 * LANES * warp - 32;
        box_x = lo / 4;                                   // (uint32 elements; negative = zero-filled by the TMA unit)
        src_off = max(lo, 0);
        dst_off = src_off - lo;
        cp_bytes = min(a.rowbytes, lo + WSLOT) - src_off;
        e_in = a.rowbytes - lo;
        if (e_in < 6 || e_in + 3 > WSLOT) e_in = -1;
        e_l1 = 3 * a.w[1] - (12 * LANES * warp - 8);
        if (e_l1 < 6 || e_l1 + 3 > PLANE) e_l1 = -1;
        if constexpr (L >= 2) {
            own2 = min(OWN2, 3 * a.w[2] - OWN2 * warp);
            last2 = own2 > 0 && OWN2 * (warp + 1) >= 3 * a.w[2];
        } else {
            own2 = min(2 * OWN2, 3 * a.w[1] - 2 * OWN2 * warp);      // level-1 values written by this warp
            last2 = false;
        }
        if (own2 < 0) own2 = 0;
        dn = 0; seg_n0 = 0; seg_q0 = 0;
        duty_m = 0; turn_w = 0; turn_c = 0;
    }
    __device__ __forceinline__ DutyState* duty_state() const { return reinterpret_cast<DutyState*>(smem + a.duty_off); }
    __device__ __forceinline__ void wait_row(int n) { mbar_wait(bar0 + a.rbar_off + 8 * (n & (RS - 1)), (uint32_t)((n / RS) & 1)); }
    __device__ __forceinline__ void wait_duty(int n) { mbar_wait(bar0 + a.rbar_off + 8 * (RS + (n & (RS - 1))), (uint32_t)((n / RS) & 1)); }

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
    // next groups into those slots.  (Edge warps patched border bytes into those slots through the
    // generic proxy: order those writes before the bulk copies that will overwrite them.)
    __device__ __forceinline__ void refill() {
        if (warp == 0 || e_in >= 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const int lim = min(g_total, g_cons + a.ng);
            while (g_issued < lim) issue_group(g_issued++);
        }
    }
    // wait for the next group of the segment; returns the offset of its first row in the warp's ring.
    // Frame borders: pixels -2, -1 <- 2, 1 (ring bytes 26..31 <- 38..40, 35..37) and pixel W <- W - 2.
    __device__ __forceinline__ int next_group() {
        mbar_wait(wbar + 8 * c_g, (uint32_t)c_phase);
        const int off = c_g * GBYTES;
        if (warp == 0 || e_in >= 0) {                // warp-uniform
            unsigned char* p = ring0 + off;
            if (warp == 0 && lane < 12) {
                const int r = lane >= 6, i = lane - 6 * r;
                p[r * WSLOT + 26 + i] = p[r * WSLOT + (i < 3 ? 38 : 32) + i];
            }
            if (e_in >= 0 && lane >= 16 && lane < 22) {
                const int r = lane >= 19, i = lane - 16 - 3 * r;
                p[r * WSLOT + e_in + i] = p[r * WSLOT + e_in - 6 + i];
            }
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
            int d[4] = {0, 0, 0, 0};
            imma(d, AF[(PH0 + G) % 3][0], b[G].x, b[G].y);
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
        if (warp == 0 || e_l1 >= 0) {                // frame borders of level 1: values -6..-1 <- 6..8, 3..5; pixel w1 <- w1 - 2
            if (warp == 0 && lane < 12) {
                const int h = lane >= 6, i = lane - 6 * h;
                plane[h * PLANE + 2 + i] = plane[h * PLANE + (i < 3 ? 14 : 8) + i];
            }
            if (e_l1 >= 0 && lane >= 16 && lane < 22) {
                const int h = lane >= 19, i = lane - 16 - 3 * h;
                plane[h * PLANE + e_l1 + i] = plane[h * PLANE + e_l1 - 6 + i];
            }
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
            int dl[4] = {0, 0, 0, 0}, dh[4] = {0, 0, 0, 0};
            imma(dl, AF[G][0], bl[G].x, bl[G].y);
            imma(dl, AF[G][1], bl[G + 1].x, bl[G + 1].y);
            imma(dh, AF[G][0], bh[G].x, bh[G].y);
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
        seg_next[L] = f; seg_last[L] = e;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            f = max(0, 2 * f - 2);
            e = min(a.h[l] - 1, 2 * e + 2);
            seg_next[l] = f; seg_last[l] = e;
        }
        nextr[1] = seg_next[1]; lastr[1] = seg_last[1];
        if constexpr (L >= 2) { nextr[2] = seg_next[2]; lastr[2] = seg_last[2]; }
        // input rows of the segment: virtual rows 2*first-2 .. 2*last+2 in groups of two, after one filler row
        vg0 = 2 * nextr[1] - 3;
        g_total = (lastr[1] - nextr[1] + 1) + 2;
        g_issued = 0;
        g_cons = 0;
        g_int0 = vg0 < 0 ? (1 - vg0) / 2 : 0;                 // first G with vg0 + 2G >= 0
        g_int1 = (a.H - vg0) / 2;                            // first G with vg0 + 2G + 1 >= H  (vg0 odd)
        box_y0 = t * a.H + vg0;
        seg_n0 = dn;
        seg_q0 = (L >= 2) ? nextr[L >= 2 ? 2 : 1] : 0;
    }

    // ---- levels >= 3, run by ONE warp per level-2 row (the warps take turns) --------------------
    // Row r of level l-1 is complete in ring l-1.  Level 2 (the source of l = 3): slot r % RS, interleaved, pixel p
    // channel c at float 3 (p + 2) + c (aprons px -2, -1 in front, px w behind).  Levels >= 3: slot r & 1, per
    // channel plane, aprons at float 2, 3 = px -2, -1, px p at float 4 + p, aprons px w, w + 1 behind.  A lane owns
    // N = 64 >> l adjacent pixels of level l: horizontal pass into the H ring of level l (HR rows), then every
    // level-l row whose five H rows are present is finished, written to ring l (or to global memory at the last
    // level) and handed to level l+1 by the same warp: no block barrier anywhere above level 2.
    template <int l>
    __device__ __forceinline__ void duty_row(int r, int src_slot) {
        constexpr int N = 64 >> l;                 // 8, 4, 2, 1 pixels per lane
        constexpr int C = N >= 4 ? 4 : N;          // pixels per vertical-pass chunk
        DutyState* ds = duty_state();
        const int wl = a.w[l], hp = a.h[l - 1];
        int hs = ds->hslot[l] + 1;
        if (hs == HR) hs = 0;
        float* const hring = reinterpret_cast<float*>(smem + a.hring_off[l]);
        if constexpr (l == 3) {
            // source = level-2 ring row, interleaved; a lane's 8 pixels in two halves of 4 (12 source pixels = 9 x LDS.128 each)
            const float* const src = reinterpret_cast<const float*>(smem + a.ring_off[2]) + src_slot * a.ring_stride[2];
            for (int px0 = lane * N; px0 < wl; px0 += 32 * N) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float x[36];                    // px 2 (px0 + 4 hf) - 2 .. + 11, interleaved (the last px is not used)
                    const float* p = src + 6 * (px0 + 4 * hf);
#pragma unroll
                    for (int k = 0; k < 9; ++k) ldv<4>(p + 4 * k, x + 4 * k);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float o[4];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            o[m] = x[3 * (2 * m + 2) + ch] * 6.0f + (x[3 * (2 * m + 1) + ch] + x[3 * (2 * m + 3) + ch]) * 4.0f +
                                   x[3 * (2 * m) + ch] + x[3 * (2 * m + 4) + ch];
                        stv<4>(hring + (hs * 3 + ch) * wl + px0 + 4 * hf, o);
                    }
                }
            }
        } else {
            const float* const src = reinterpret_cast<const float*>(smem + a.ring_off[l - 1]) + src_slot * 3 * a.ring_stride[l - 1];
            for (int px0 = lane * N; px0 < wl; px0 += 32 * N) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float* p = src + ch * a.ring_stride[l - 1] + 2 * px0 + 2;     // px 2 px0 - 2
                    float x[2 * N + 4];
                    ldv<2>(p, x);
                    if constexpr (N >= 2) {
#pragma unroll
                        for (int k = 0; k < N / 2; ++k) ldv<4>(p + 2 + 4 * k, x + 2 + 4 * k);
                    } else {
                        ldv<2>(p + 2, x + 2);
                    }
                    x[2 * N + 2] = p[2 * N + 2];
                    float o[N];
#pragma unroll
                    for (int m = 0; m < N; ++m)
                        o[m] = x[2 * m + 2] * 6.0f + (x[2 * m + 1] + x[2 * m + 3]) * 4.0f + x[2 * m] + x[2 * m + 4];
                    float* hd = hring + (hs * 3 + ch) * wl + px0;
#pragma unroll
                    for (int k = 0; k < N; k += C) stv<C>(hd + k, o + k);
                }
            }
        }
        int nx = ds->nextr[l];
        const int lst = ds->lastr[l];
        __syncwarp();
        while (nx <= lst && min(2 * nx + 2, hp - 1) <= r) {
            const int q = nx;
            int so[5];                              // H-ring rows of the five source rows: row r' sits (r - r') slots behind the newest
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                int sl = hs - (r - vhr_reflect101(2 * q - 2 + d, hp));
                if (sl < 0) sl += HR;
                so[d] = sl * 3 * wl;
            }
            float* const dring = (l < L) ? reinterpret_cast<float*>(smem + a.ring_off[l < L ? l : 3]) + (q & 1) * 3 * a.ring_stride[l < L ? l : 3] : nullptr;
            for (int px0 = lane * C; px0 < wl; px0 += 32 * C) {
                float v[3][C];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float* hc = hring + ch * wl + px0;
                    float t0[C], t1[C], t2[C], t3[C], t4[C];
                    ldv<C>(hc + so[0], t0); ldv<C>(hc + so[1], t1); ldv<C>(hc + so[2], t2);
                    ldv<C>(hc + so[3], t3); ldv<C>(hc + so[4], t4);
#pragma unroll
                    for (int m = 0; m < C; ++m)
                        v[ch][m] = (t2[m] * 6.0f + (t1[m] + t3[m]) * 4.0f + t0[m] + t4[m]) * (1.0f / 256.0f);
                }
                if constexpr (l == L) {
                    float t[3 * C];
#pragma unroll
                    for (int m = 0; m < C; ++m) { t[3 * m] = v[0][m]; t[3 * m + 1] = v[1][m]; t[3 * m + 2] = v[2][m]; }
                    float* o = out_frame + ((size_t)q * wl + px0) * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) stv<C>(o + C * k, t + C * k);
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) stv<C>(dring + ch * a.ring_stride[l] + 4 + px0, v[ch]);
                }
            }
            nx = q + 1;
            if constexpr (l < L) {
                __syncwarp();
                if (lane < 3) {                     // reflect-101 aprons of the new row, one channel per lane
                    float* pa = dring + lane * a.ring_stride[l];
                    pa[2] = pa[4 + vhr_reflect101(-2, wl)];
                    pa[3] = pa[4 + vhr_reflect101(-1, wl)];
                    pa[4 + wl] = pa[4 + vhr_reflect101(wl, wl)];
                    pa[5 + wl] = pa[4 + vhr_reflect101(wl + 1, wl)];
                }
                __syncwarp();
                duty_row<l + 1>(q, q & 1);
            }
        }
        __syncwarp();
        if (lane == 0) { ds->hslot[l] = hs; ds->nextr[l] = nx; }
    }

    // Upper levels of the next level-2 row (number duty_m), by the warp whose turn it is.  Turns
    // rotate with a skew (warp (n + n / nw) % nw) so that the heavier rows, which also finish rows
    // of the upper levels, do not always fall on the same warps.  The row must be complete and the
    // previous row's upper-level work done (shared H rings, DutyState).
    // Every parity wait below lags its barrier by less than one phase: the next phase of row
    // barrier n % RS needs this warp's signal for row n + RS, which follows duty(n) in program
    // order via the slot wait in publish(); the next phase of duty barrier (n-1) % RS needs duty(n).
    __device__ __forceinline__ void run_duty() {
        if constexpr (L >= 3) {
            const int n = duty_m;
            if (turn_w == warp) {
                wait_row(n);
                if (n >= 1) wait_duty(n - 1);
                duty_row<3>(seg_q0 + (n - seg_n0), n & (RS - 1));
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar0 + a.rbar_off + 8 * (RS + (n & (RS - 1)))) : "memory");
            }
            const int nw = blockDim.x >> 5;
            duty_m = n + 1;
            ++turn_w;
            if (++turn_c == nw) { turn_c = 0; ++turn_w; }
            if (turn_w >= nw) turn_w -= nw;
            if (turn_w >= nw) turn_w -= nw;
        }
    }
    // A finished level-2 row: this lane holds 12 sums (exact, < 2^24) of which the lanes with q < 2 are real (the
    // MMA columns 4..7 of level 2 are duplicates): value x[4 G + e + 2 h] = level-2 value 96 q + 48 e + 16 G + 2 g + h
    // of the warp's 192-value window.
    __device__ __forceinline__ void publish(int q, const int (&s)[12]) {
        const int n = dn;
        if constexpr (L >= 3) {
            // the ring slot's previous row (n - RS) has been consumed; this also keeps the signals of
            // row n out of the row barrier's previous phase
            if (n >= RS) wait_duty(n - RS);
        }
        {
            float* dst;
            if constexpr (L == 2) dst = out_frame + (size_t)q * a.w[2] * 3 + OWN2 * warp;
            else dst = reinterpret_cast<float*>(smem + a.ring_off[2]) + (n & (RS - 1)) * a.ring_stride[2] + 6 + OWN2 * warp;
            const int j0 = 96 * (lane & 3) + 2 * (lane >> 2);
            if ((lane & 3) < 2) {
#pragma unroll
                for (int G = 0; G < 3; ++G)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = j0 + 48 * e + 16 * G;
                        if (j < own2)
                            *reinterpret_cast<float2*>(dst + j) = make_float2((float)s[4 * G + e] * (1.0f / 65536.0f),
                                                                              (float)s[4 * G + e + 2] * (1.0f / 65536.0f));
                    }
            }
            if constexpr (L >= 3) {
                if (warp == 0 || last2) {            // aprons: px -2 <- px 2, px -1 <- px 1, px w2 <- px w2 - 2
                    __syncwarp();
                    float* row = reinterpret_cast<float*>(smem + a.ring_off[2]) + (n & (RS - 1)) * a.ring_stride[2];
                    if (warp == 0 && lane < 6) row[lane] = row[(lane < 3 ? 12 : 6) + lane];
                    if (last2 && lane >= 8 && lane < 11) row[6 + 3 * a.w[2] + lane - 8] = row[3 * a.w[2] + lane - 8];
                }
            }
        }
        if constexpr (L >= 3) {
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar0 + a.rbar_off + 8 * (n & (RS - 1))) : "memory");
        }
        dn = n + 1;
        nextr[2] = q + 1;
        if (n - DLY >= seg_n0) run_duty();           // an older row: complete by now, and nobody waits for its upper levels yet
    }

    // ---- one segment ----------------------------------------------------------------------------
    __device__ __forceinline__ void run_segment() {
        __syncthreads();                // every warp is done with the previous segment (upper levels included)
        if constexpr (L >= 3) {
            if (threadIdx.x == 0) {
                DutyState* ds = duty_state();
#pragma unroll
                for (int l = 3; l <= L; ++l) { ds->nextr[l] = seg_next[l]; ds->lastr[l] = seg_last[l]; }
            }
            __syncthreads();
        }
        refill();
        prime();
        if constexpr (L == 1) {
            int n = 0;
            const int j0 = 96 * (lane & 3) + 2 * (lane >> 2) - 8;            // first value of the lane's pairs, relative to the owned range
            for (int r = nextr[1]; r <= lastr[1]; ++r) {
                uint32_t v[6];
                l1_row(v);
                float* o = out_frame + (size_t)r * a.w[1] * 3 + 2 * OWN2 * warp;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int j = j0 + 48 * (k & 1) + 16 * (k >> 1);
                    if (j >= 0 && j < own2)
                        *reinterpret_cast<float2*>(o + j) = make_float2((float)(v[k] & 0xFFFFu) * (1.0f / 256.0f),
                                                                        (float)(v[k] >> 16) * (1.0f / 256.0f));
                }
                if (++n == 2) { n = 0; refill(); }
            }
        } else {
            const int h1 = a.h[1];
            int q = nextr[2];
            const int ql = lastr[2];
            // level-2 vertical pass, incremental: A = r(2q-2) + 4 r(2q-1) + 6 r(2q), B = 4 r(2q-1) + r(2q-2), C = r(2q)
            // (r = the level-2 horizontal sums of a level-1 row); rows 0 .. 2 prime it at the top of a frame
            int A[12], B[12], C[12];
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
                    publish(0, s);
                    q = 1;
                }
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    B[k] = 4 * x1[k] + x0[k];
                    A[k] = B[k] + 6 * C[k];
                }
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
                            B[k] = 4 * n1[k] + C[k];
                        }
                    }
                    l12_row(C);
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        s[k] += C[k];
                        A[k] = B[k] + 6 * C[k];
                    }
                } else if (has1) {                   // row 2q+2 = h1 reflects to 2q
                    int n1[12];
                    l12_row(n1);
#pragma unroll
                    for (int k = 0; k < 12; ++k) s[k] = A[k] + 4 * n1[k] + C[k];
                } else {                             // rows 2q+1, 2q+2 reflect to 2q-1, 2q-2
#pragma unroll
                    for (int k = 0; k < 12; ++k) s[k] = A[k] + B[k];
                }
                refill();
                publish(q, s);
            }
            while (duty_m < dn) run_duty();          // the last rows of the segment
        }
    }
};

template <int L, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? 2 : 1) pyrdown_mma_kernel(const MmaArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    const long long lo = a.total_rows * blockIdx.x / gridDim.x;
    const long long hi = a.total_rows * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int b = 0; b < nw * a.ng; ++b) mbar_init(smem_u32(smem) + 8 * b, 1);             // rows landed, per warp and group
        for (int b = 0; b < RS; ++b) {
            mbar_init(smem_u32(smem) + a.rbar_off + 8 * b, nw);                                // row barriers: one arrival per warp
            mbar_init(smem_u32(smem) + a.rbar_off + 8 * (RS + b), 1);                          // duty barriers: one arrival per row
        }
        DutyState* ds = reinterpret_cast<DutyState*>(smem + a.duty_off);
        for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { ds->hslot[l] = 0; ds->nextr[l] = 0; ds->lastr[l] = -1; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the level-1 planes are read 64 bytes per block window: clear them once so that never-written tail bytes are defined
    {
        const int nw = blockDim.x >> 5;
        uint32_t* pz = reinterpret_cast<uint32_t*>(smem + a.plane_off);
        for (int i = threadIdx.x; i < nw * 2 * PLANE / 4; i += blockDim.x) pz[i] = 0u;
    }
    Stream<L> st(a, &tmap, smem);
    const int hL = a.h[L];
    long long pos = lo;
    while (pos < hi) {
        const int t = (int)(pos / hL);
        const int r0 = (int)(pos - (long long)t * hL);
        const long long frame_end = (long long)(t + 1) * hL;
        const int r1 = (int)((hi < frame_end ? hi : frame_end) - (long long)t * hL);
        st.begin_segment(t, r0, r1);
        st.run_segment();               // starts with a block barrier: the previous segment's shared rows are dead
        pos += r1 - r0;
    }
}

template <int L, int MAXT>
int launch_mma(vhr_ctx* ctx, const MmaArgs& a, const CUtensorMap& tmap, int threads, int smem_bytes, cudaStream_t stream) {
    auto kern = pyrdown_mma_kernel<L, MAXT>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    VHR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes));
    if (per_sm < 1) return VHR_ERR_UNSUPPORTED;
    long long grid = (long long)per_sm * ctx->num_sms;
    if (grid > a.total_rows) grid = a.total_rows;
    kern<<<(int)grid, threads, smem_bytes, stream>>>(a, tmap);
    return vhr_after_launch(ctx, "pyrdown_mma_kernel");
}

template <int MAXT>
int dispatch_mma(vhr_ctx* ctx, const MmaArgs& a, const CUtensorMap& tmap, int threads, int smem_bytes, cudaStream_t stream) {
    switch (a.levels) {
        case 1: return launch_mma<1, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 2: return launch_mma<2, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 3: return launch_mma<3, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 4: return launch_mma<4, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 5: return launch_mma<5, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 6: return launch_mma<6, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
    }
    return VHR_ERR_INVALID;
}

}  // namespace

// Returns VHR_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible.
int vhr_pyrdown_mma(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                    cudaStream_t stream) {
    if (W % 16 != 0 || W > 8 * LANES * 16 || (reinterpret_cast<uintptr_t>(d_frames) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(d_level) & 15) != 0)
        return VHR_ERR_UNSUPPORTED;
    if (levels >= 3 && W % 64 != 0) return VHR_ERR_UNSUPPORTED;     // a lane of the upper-level warp owns 64 >> l pixels
    MmaArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = d_frames; a.out = d_level; a.T = T; a.H = H; a.W = W; a.levels = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    if (levels == 1 ? H < 2 : a.h[1] < 3) return VHR_ERR_UNSUPPORTED;   // the vertical passes are primed with three rows
    a.total_rows = (long long)T * a.h[levels];
    a.nt = W / 8;
    a.rowbytes = 3 * W;
    const int warps = (a.nt + LANES - 1) / LANES;
    const int threads = warps * 32;
    auto al16 = [](int v) { return (v + 15) & ~15; };
    int fixed = al16(warps * 2 * PLANE);                            // everything but the input rings
    if (levels >= 3) {
        a.ring_stride[2] = (3 * (a.w[2] + 4) + 3) & ~3;             // interleaved: px p channel c at float 3 (p + 2) + c
        fixed = al16(fixed + RS * a.ring_stride[2] * 4);
    }
    for (int l = 3; l < levels; ++l) {
        a.ring_stride[l] = (a.w[l] + 8 + 3) & ~3;                   // px p at float 4 + p; aprons at 2, 3 and w + 4, w + 5
        fixed = al16(fixed + 2 * 3 * a.ring_stride[l] * 4);
    }
    for (int l = 3; l <= levels; ++l) fixed = al16(fixed + HR * 3 * a.w[l] * 4);
    // per-warp input rings: as deep as two CTAs per SM allow (2 groups are consumed between two refills)
    const int budget = (threads <= 256 ? ctx->smem_optin / 2 - 2048 : ctx->smem_optin - 1024);
    int ng = 6;
    auto al128 = [](int v) { return (v + 127) & ~127; };
    auto head = [&](int n) { return al128(al16(8 * (warps * n + 2 * RS)) + al16((int)sizeof(DutyState))); };
    while (ng > 3 && head(ng) + warps * ng * GBYTES + fixed > budget) --ng;
    if (head(ng) + warps * ng * GBYTES + fixed > ctx->smem_optin) return VHR_ERR_UNSUPPORTED;
    a.ng = ng;
    a.rbar_off = 8 * warps * ng;
    a.duty_off = al16(8 * (warps * ng + 2 * RS));
    int off = head(ng);
    a.in_off = off;
    off = al16(off + warps * ng * GBYTES);
    a.plane_off = off;
    off = al16(off + warps * 2 * PLANE);
    if (levels >= 3) {
        a.ring_off[2] = off;
        off = al16(off + RS * a.ring_stride[2] * 4);
    }
    for (int l = 3; l < levels; ++l) {
        a.ring_off[l] = off;
        off = al16(off + 2 * 3 * a.ring_stride[l] * 4);
    }
    for (int l = 3; l <= levels; ++l) {
        a.hring_off[l] = off;
        off = al16(off + HR * 3 * a.w[l] * 4);
    }
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
    if (threads <= 256) return dispatch_mma<256>(ctx, a, tmap, threads, off, stream);
    return dispatch_mma<512>(ctx, a, tmap, threads, off, stream);
}
