// Streaming CUDA-core form of the fused pyrDown cascade (W % 16 == 0; W % 64 == 0 for >= 3 levels; 16-byte aligned
// frames).  Since round 2 it serves the shapes and level counts the tensor-core kernel (pyrdown_umma.cu: 4 levels,
// W % 80 == 0) does not take; the generic kernel in pyrdown.cu covers the rest, same arithmetic.
//
// Same spec as pyrdown.cu (cv2.pyrDown float32 semantics; levels 1-2 exact integers * 2^-8l).  Current form (v6;
// the stage-by-stage history with the ncu captures is in profiles/README.md, the design in DESIGN.md section 4.1b):
//   * A thread owns a fixed column group: 8 input pixels -> 4 px of level 1 -> 2 px of level 2; a warp owns 30 groups,
//     lanes 0 and 31 are halo lanes that recompute the neighbour warp's edge group.  Levels 1 and 2 live in registers.
//     Level 1 runs its VERTICAL pass first, on the raw bytes split into packed 16-bit lanes with two incremental partial
//     rows (A + 4 n1 + n2), so only one row per level-1 row goes through the horizontal pass (IDP2A on same-channel
//     pixel pairs, neighbours by warp shuffle); the level-2 horizontal pass takes the finished level-1 row by shuffle.
//   * Input per WARP: each warp has its own shared-memory ring of two-row groups, one cp.async.bulk.tensor.2d
//     (UTMALDG.2D) per group on a per-group mbarrier; reflected rows at the frame's top / bottom by cp.async.bulk.
//     No block barrier in the steady state.
//   * Levels >= 3 (1/16 of the data) by ONE warp per level-2 row, in turn: every warp publishes its part of the
//     finished level-2 row in a 4-slot ring (row mbarrier), the warp on duty runs the whole upper cascade for that row
//     (duty mbarrier frees the slot).
//   * Persistent grid over the flattened (frame, final-row) space, equal contiguous shares.
// Bound by instruction issue (3.0e9 warp-instructions per 1080p clip, 73 % issue-active, 39 % of DRAM peak): 3.57 ms.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int HR = 5;        // rows in a private H ring (exactly the vertical footprint)
constexpr int LANES = 30;    // column groups per warp (lanes 1..30); lanes 0 / 31 are halo lanes
constexpr int WSLOT = 832;   // bytes of one input row in a warp's ring (>= 32 column groups x 24 B + 8 B either side)
constexpr int GBYTES = 2 * WSLOT;   // input rows travel in groups of two = one TMA box, a multiple of 128 bytes
constexpr int RS = 4;        // level-2 ring slots = rows a warp may run ahead of the upper levels
constexpr int DLY = 2;       // the upper levels of row n start when their warp has published row n + DLY

struct StreamArgs {
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
    int rbar_off;                          // byte offset of the RS row barriers, followed by the RS duty barriers
    int ring_off[VHR_MAX_LEVELS + 1];      // levels 2..L-1: newest rows, float planar, double-buffered
    int ring_stride[VHR_MAX_LEVELS + 1];   // floats per channel plane row
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
                printf("WATCHDOG block %d warp %d bar_off %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar & 0xffffu, parity);
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

template <int L>
struct Stream {
    const StreamArgs& a;
    const CUtensorMap* tmap;
    unsigned char* smem;
    const int i;              // column group of this lane (-1 / >= nt on idle halo lanes)
    const bool own;           // lane owns outputs of column group i
    const bool first_col, last_col;
    const unsigned char* rd;  // this lane's 36 bytes in row 0 of the warp's input ring
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
    uint32_t VA[12], VC[12];           // level-1 vertical pass: partial sum of the next row, last input row (split layout)

    __device__ Stream(const StreamArgs& a_, const CUtensorMap* tm, unsigned char* s, int col)
        : a(a_), tmap(tm), smem(s), i(col), own((threadIdx.x & 31) >= 1 && (threadIdx.x & 31) <= LANES && col < a_.nt),
          first_col(col == 0), last_col(col == a_.nt - 1) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        rd = smem + a.in_off + warp * a.ng * GBYTES + 24 * lane;
        bar0 = smem_u32(smem);
        wbar = bar0 + 8 * warp * a.ng;
        c_g = 0; c_phase = 0; g_cons = 0; p_g = 0; g_issued = 0; g_total = 0; vg0 = 0;
        g_int0 = 0; g_int1 = 0; box_y0 = 0;
        ring_u32 = smem_u32(smem + a.in_off + warp * a.ng * GBYTES);
        // ring-row byte b of the warp <-> byte 24 * LANES * warp - 32 + b of the input row
        const int lo = 24 * LANES * warp - 32;
        box_x = lo / 4;                                   // (uint32 elements; negative = zero-filled by the TMA unit)
        src_off = max(lo, 0);
        dst_off = src_off - lo;
        cp_bytes = min(a.rowbytes, lo + WSLOT) - src_off;
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
    // next groups into those slots.
    __device__ __forceinline__ void refill() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            const int lim = min(g_total, g_cons + a.ng);
            while (g_issued < lim) issue_group(g_issued++);
        }
    }
    // ---- level 1, vertical pass first ---------------------------------------------------------
    // A lane reads only its own 24 bytes of a row (3 x LDS.64).  Each word is split into two
    // registers of two 16-bit lanes (bytes 0,2 and bytes 1,3): the vertical 5-tap then runs on packed
    // values (<= 4080) for two bytes per instruction, and only ONE row per level-1 row goes through
    // the horizontal pass (3 IDP2A per value on same-channel pixel pairs; the neighbours' pairs come
    // by shuffle).  Per level-1 row r the vertical pass is incremental:
    //     out_r = A + 4 n1 + n2,   A' = C + 4 n1 + 6 n2,   C' = n2
    // with n1, n2 the new rows 2r+1, 2r+2, C = row 2r and A = row(2r-2) + 4 row(2r-1) + 6 row(2r).
    // All sums are exact integers (level 1 = sum / 256), identical to the horizontal-first order.
    __device__ __forceinline__ void load_unpack(const unsigned char* p, uint32_t (&u)[12]) {
        const uint2 q0 = *reinterpret_cast<const uint2*>(p + 8), q1 = *reinterpret_cast<const uint2*>(p + 16);
        const uint2 q2 = *reinterpret_cast<const uint2*>(p + 24);
        const uint32_t w[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            u[2 * k] = __byte_perm(w[k], 0u, 0x4240);            // bytes 0, 2
            u[2 * k + 1] = __byte_perm(w[k], 0u, 0x4341);        // bytes 1, 3
        }
    }
    // (register, half) of element e = 3 px + ch of the lane's 24 values in the split layout
    __host__ __device__ static constexpr int el_reg(int e) { return 2 * (e >> 2) + (e & 1); }
    __host__ __device__ static constexpr int el_half(int e) { return (e & 3) >> 1; }
    // same-channel pair (px 2k, px 2k+1) of channel c from the split layout: one PRMT
    template <int K, int CH>
    __device__ __forceinline__ uint32_t pair_of(const uint32_t (&v)[12]) const {
        constexpr int e1 = 6 * K + CH, e2 = e1 + 3;
        constexpr uint32_t sel = (uint32_t)(2 * el_half(e1)) | ((uint32_t)(2 * el_half(e1) + 1) << 4) |
                                 ((uint32_t)(4 + 2 * el_half(e2)) << 8) | ((uint32_t)(4 + 2 * el_half(e2) + 1) << 12);
        return __byte_perm(v[el_reg(e1)], v[el_reg(e2)], sel);
    }
    template <int CH>
    __device__ __forceinline__ void hpass_channel(const uint32_t (&vv)[12], uint32_t (&out)[6]) {
        const uint32_t p0 = pair_of<0, CH>(vv), p1 = pair_of<1, CH>(vv), p2 = pair_of<2, CH>(vv), p3 = pair_of<3, CH>(vv);
        uint32_t pl = __shfl_up_sync(0xffffffffu, p3, 1);          // px -2, -1
        uint32_t pr = __shfl_down_sync(0xffffffffu, p0, 1);        // px 8
        if (first_col) pl = __byte_perm(p1, p0, 0x7610);           // px -2,-1 <- px 2,1
        if (last_col) pr = p3;                                     // px 8 <- px 6
        uint32_t o0 = __dp2a_lo(pl, 0x0401u, 0u);
        o0 = __dp2a_lo(p0, 0x0406u, o0);
        o0 = __dp2a_lo(p1, 0x0001u, o0);
        uint32_t o1 = __dp2a_lo(p0, 0x0401u, 0u);
        o1 = __dp2a_lo(p1, 0x0406u, o1);
        o1 = __dp2a_lo(p2, 0x0001u, o1);
        uint32_t o2 = __dp2a_lo(p1, 0x0401u, 0u);
        o2 = __dp2a_lo(p2, 0x0406u, o2);
        o2 = __dp2a_lo(p3, 0x0001u, o2);
        uint32_t o3 = __dp2a_lo(p2, 0x0401u, 0u);
        o3 = __dp2a_lo(p3, 0x0406u, o3);
        o3 = __dp2a_lo(pr, 0x0001u, o3);
        out[2 * CH] = __byte_perm(o0, o1, 0x5410);                 // level-1 px 4i, 4i+1 (values <= 65280)
        out[2 * CH + 1] = __byte_perm(o2, o3, 0x5410);             // level-1 px 4i+2, 4i+3
    }
    // wait for the next group of the segment; returns the address of its first row for this lane
    __device__ __forceinline__ const unsigned char* next_group() {
        mbar_wait(wbar + 8 * c_g, (uint32_t)c_phase);
        const unsigned char* p = rd + c_g * GBYTES;
        ++g_cons;
        if (++c_g == a.ng) { c_g = 0; c_phase ^= 1; }
        return p;
    }

    __device__ __forceinline__ void prime() {       // first three input rows of a segment
        const unsigned char* p = next_group();      // (the first row of a segment's group 0 is a filler)
        uint32_t x0[12], x1[12];
        load_unpack(p + WSLOT, x0);
        p = next_group();
        load_unpack(p, x1);
        load_unpack(p + WSLOT, VC);
#pragma unroll
        for (int k = 0; k < 12; ++k) VA[k] = x0[k] + (x1[k] << 2) + VC[k] * 6u;
        refill();
    }
    __device__ __forceinline__ void l1_row(uint32_t (&v)[6]) {
        const unsigned char* p = next_group();
        uint32_t vv[12];
#pragma unroll
        for (int h = 0; h < 3; ++h) {      // one LDS.64 of each new row at a time: 8 of the 24 values
            const uint2 qa = *reinterpret_cast<const uint2*>(p + 8 + 8 * h);
            const uint2 qb = *reinterpret_cast<const uint2*>(p + WSLOT + 8 + 8 * h);
            const uint32_t wa[2] = {qa.x, qa.y}, wb[2] = {qb.x, qb.y};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const int k = 4 * h + 2 * j + par;
                    const uint32_t n1 = __byte_perm(wa[j], 0u, par ? 0x4341 : 0x4240);
                    const uint32_t n2 = __byte_perm(wb[j], 0u, par ? 0x4341 : 0x4240);
                    const uint32_t t = n1 << 2;
                    vv[k] = VA[k] + t + n2;                 // <= 4080 per 16-bit lane
                    VA[k] = VC[k] + t + n2 * 6u;
                    VC[k] = n2;
                }
            }
        }
        hpass_channel<0>(vv, v);
        hpass_channel<1>(vv, v);
        hpass_channel<2>(vv, v);
    }
    // level-2 horizontal pass of a finished level-1 row: neighbours' pixels by shuffle
    __device__ __forceinline__ void l2_hrow(const uint32_t (&v)[6], uint32_t (&o)[6]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t pv = __shfl_up_sync(0xffffffffu, v[2 * c + 1], 1);     // level-1 px 4i-2, 4i-1
            uint32_t nv = __shfl_down_sync(0xffffffffu, v[2 * c], 1);       // level-1 px 4i+4
            if (first_col) pv = __byte_perm(v[2 * c + 1], v[2 * c], 0x7610);   // px -2,-1 <- px 2,1
            if (last_col) nv = v[2 * c + 1];                                   // px w1 <- px w1-2
            uint32_t o0 = __dp2a_lo(pv, 0x0401u, 0u);
            o0 = __dp2a_lo(v[2 * c], 0x0406u, o0);
            o0 = __dp2a_lo(v[2 * c + 1], 0x0001u, o0);
            uint32_t o1 = __dp2a_lo(v[2 * c], 0x0401u, 0u);
            o1 = __dp2a_lo(v[2 * c + 1], 0x0406u, o1);
            o1 = __dp2a_lo(nv, 0x0001u, o1);
            o[2 * c] = o0; o[2 * c + 1] = o1;
        }
    }
    __device__ __forceinline__ void l12_row(uint32_t (&o)[6]) {
        uint32_t v[6];
        l1_row(v);
        l2_hrow(v, o);
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
    // Row r of level l-1 is complete in ring l-1 (slot r & 1; per channel plane: aprons at float
    // 2,3 = px -2,-1, px p at float 4+p, aprons px w, w+1 behind).  A lane owns N = 64 >> l
    // adjacent pixels of level l: horizontal pass into the H ring of level l (HR rows), then
    // every level-l row whose five H rows are present is finished, written to ring l (or to
    // global memory at the last level) and handed to level l+1 by the same warp: no block
    // barrier anywhere above level 2.
    template <int l>
    __device__ __forceinline__ void duty_row(int r, int src_slot) {
        constexpr int N = 64 >> l;                 // 8, 4, 2, 1 pixels per lane
        constexpr int C = N >= 4 ? 4 : N;          // pixels per vertical-pass chunk
        DutyState* ds = duty_state();
        const int lane = threadIdx.x & 31;
        const int wl = a.w[l], hp = a.h[l - 1];
        int hs = ds->hslot[l] + 1;
        if (hs == HR) hs = 0;
        float* const hring = reinterpret_cast<float*>(smem + a.hring_off[l]);
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
            float* const dring = (l < L) ? reinterpret_cast<float*>(smem + a.ring_off[l < L ? l : 2]) + (q & 1) * 3 * a.ring_stride[l < L ? l : 2] : nullptr;
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
                    float* pl = dring + lane * a.ring_stride[l];
                    pl[2] = pl[4 + vhr_reflect101(-2, wl)];
                    pl[3] = pl[4 + vhr_reflect101(-1, wl)];
                    pl[4 + wl] = pl[4 + vhr_reflect101(wl, wl)];
                    pl[5 + wl] = pl[4 + vhr_reflect101(wl + 1, wl)];
                }
                __syncwarp();
                duty_row<l + 1>(q, q & 1);
            }
        }
        __syncwarp();
        if (lane == 0) { ds->hslot[l] = hs; ds->nextr[l] = nx; }
    }

    // ---- level 2: vertical pass + hand-over --------------------------------------------------
    __device__ __forceinline__ void l2_vpass(const uint32_t (&xa)[6], const uint32_t (&xb)[6], const uint32_t (&xc)[6],
                                             const uint32_t (&xd)[6], const uint32_t (&xe)[6], float (&f)[6]) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t s = (xa[k] + xe[k]) + 4u * (xb[k] + xd[k]) + 6u * xc[k];     // < 2^24: exact
            f[k] = (float)s * (1.0f / 65536.0f);
        }
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
            if (turn_w == (int)(threadIdx.x >> 5)) {
                wait_row(n);
                if (n >= 1) wait_duty(n - 1);
                duty_row<3>(seg_q0 + (n - seg_n0), n & (RS - 1));
                __syncwarp();
                if ((threadIdx.x & 31) == 0)
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
    __device__ __forceinline__ void publish(int q, const float (&f)[6]) {
        const int n = dn;
        if constexpr (L >= 3) {
            // the ring slot's previous row (n - RS) has been consumed; this also keeps the signals of
            // row n out of the row barrier's previous phase
            if (n >= RS) wait_duty(n - RS);
        }
        if (own) {
            if constexpr (L == 2) {
                float2* o = reinterpret_cast<float2*>(out_frame + ((size_t)q * a.w[2] + 2 * i) * 3);
                o[0] = make_float2(f[0], f[2]);      // px 2i: c0 c1
                o[1] = make_float2(f[4], f[1]);      //        c2 | px 2i+1: c0
                o[2] = make_float2(f[3], f[5]);      //        c1 c2
            } else {
                float* const dst = reinterpret_cast<float*>(smem + a.ring_off[2]) + (n & (RS - 1)) * 3 * a.ring_stride[2];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    *reinterpret_cast<float2*>(dst + ch * a.ring_stride[2] + 4 + 2 * i) = make_float2(f[2 * ch], f[2 * ch + 1]);
                if (i <= 1 || last_col) {            // aprons: px -2 <- px 2, px -1 <- px 1, px w2 <- px w2-2
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float* pl = dst + ch * a.ring_stride[2];
                        if (i == 1) pl[2] = f[2 * ch];
                        if (i == 0) pl[3] = f[2 * ch + 1];
                        if (last_col) pl[4 + a.w[2]] = f[2 * ch];
                    }
                }
            }
        }
        if constexpr (L >= 3) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0)
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
            for (int r = nextr[1]; r <= lastr[1]; ++r) {
                uint32_t v[6];
                l1_row(v);
                if (own) {
                    float* o = out_frame + ((size_t)r * a.w[1] + 4 * i) * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        o[ch] = (float)(v[2 * ch] & 0xFFFFu) * (1.0f / 256.0f);
                        o[3 + ch] = (float)(v[2 * ch] >> 16) * (1.0f / 256.0f);
                        o[6 + ch] = (float)(v[2 * ch + 1] & 0xFFFFu) * (1.0f / 256.0f);
                        o[9 + ch] = (float)(v[2 * ch + 1] >> 16) * (1.0f / 256.0f);
                    }
                }
                if (++n == 2) { n = 0; refill(); }
            }
        } else {
            const int h1 = a.h[1];
            int q = nextr[2];
            const int ql = lastr[2];
            uint32_t x0[6], x1[6], x2[6];
            // level-2 H rows of level-1 rows 2q-2 .. 2q (rows 0 .. 2 at the top of a frame)
#pragma unroll 1
            for (int p = 0; p < 3; ++p) {
                uint32_t o[6];
                l12_row(o);
#pragma unroll
                for (int k = 0; k < 6; ++k) { x0[k] = x1[k]; x1[k] = x2[k]; x2[k] = o[k]; }
                refill();
            }
            if (q == 0) {                            // rows -2,-1 reflect to 2,1
                float f[6];
                l2_vpass(x2, x1, x0, x1, x2, f);
                publish(0, f);
                q = 1;
            }
#pragma unroll 1
            for (; q <= ql; ++q) {
                uint32_t x3[6], x4[6];
                const bool has1 = 2 * q + 1 <= h1 - 1, has2 = 2 * q + 2 <= h1 - 1;
                if (has1) {
                    l12_row(x3);
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) x3[k] = x1[k];          // row h1 reflects to h1-2 = 2q-1
                }
                if (has2) {
                    l12_row(x4);
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) x4[k] = has1 ? x2[k] : x0[k];   // row 2q+2 reflects to 2q / 2q-2
                }
                float f[6];
                l2_vpass(x0, x1, x2, x3, x4, f);
#pragma unroll
                for (int k = 0; k < 6; ++k) { x0[k] = x2[k]; x1[k] = x3[k]; x2[k] = x4[k]; }
                refill();
                publish(q, f);
            }
            while (duty_m < dn) run_duty();          // the last rows of the segment
        }
    }
};

template <int L, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? 2 : 1) pyrdown_stream_kernel(const StreamArgs a, const __grid_constant__ CUtensorMap tmap) {
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
    const int col = (int)(threadIdx.x >> 5) * LANES + (int)(threadIdx.x & 31) - 1;
    Stream<L> st(a, &tmap, smem, col);
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
int launch_stream(vhr_ctx* ctx, const StreamArgs& a, const CUtensorMap& tmap, int threads, int smem_bytes, cudaStream_t stream) {
    auto kern = pyrdown_stream_kernel<L, MAXT>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    VHR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes));
    if (per_sm < 1) return VHR_ERR_UNSUPPORTED;
    long long grid = (long long)per_sm * ctx->num_sms;
    if (grid > a.total_rows) grid = a.total_rows;
    kern<<<(int)grid, threads, smem_bytes, stream>>>(a, tmap);
    return vhr_after_launch(ctx, "pyrdown_stream_kernel");
}

template <int MAXT>
int dispatch_stream(vhr_ctx* ctx, const StreamArgs& a, const CUtensorMap& tmap, int threads, int smem_bytes, cudaStream_t stream) {
    switch (a.levels) {
        case 1: return launch_stream<1, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 2: return launch_stream<2, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 3: return launch_stream<3, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 4: return launch_stream<4, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 5: return launch_stream<5, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
        case 6: return launch_stream<6, MAXT>(ctx, a, tmap, threads, smem_bytes, stream);
    }
    return VHR_ERR_INVALID;
}

}  // namespace

// Returns VHR_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible.
int vhr_pyrdown_stream(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                       cudaStream_t stream) {
    if (W % 16 != 0 || W > 8 * LANES * 16 || (reinterpret_cast<uintptr_t>(d_frames) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(d_level) & 15) != 0)
        return VHR_ERR_UNSUPPORTED;
    if (levels >= 3 && W % 64 != 0) return VHR_ERR_UNSUPPORTED;     // a lane of the upper-level warp owns 64 >> l pixels
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = d_frames; a.out = d_level; a.T = T; a.H = H; a.W = W; a.levels = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    if (levels == 1 ? H < 2 : a.h[1] < 3) return VHR_ERR_UNSUPPORTED;   // the register windows assume >= 3 level-1 rows
    a.total_rows = (long long)T * a.h[levels];
    a.nt = W / 8;
    a.rowbytes = 3 * W;
    const int warps = (a.nt + LANES - 1) / LANES;
    const int threads = warps * 32;
    auto al16 = [](int v) { return (v + 15) & ~15; };
    int fixed = 0;                                                  // everything but the input rings
    for (int l = 2; l < levels; ++l) {
        a.ring_stride[l] = (a.w[l] + 8 + 3) & ~3;                   // px p at float 4 + p; aprons at 2, 3 and w + 4, w + 5
        fixed = al16(fixed + (l == 2 ? RS : 2) * 3 * a.ring_stride[l] * 4);
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
    for (int l = 2; l < levels; ++l) {
        a.ring_off[l] = off;
        off = al16(off + (l == 2 ? RS : 2) * 3 * a.ring_stride[l] * 4);
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
    if (threads <= 256) return dispatch_stream<256>(ctx, a, tmap, threads, off, stream);
    return dispatch_stream<512>(ctx, a, tmap, threads, off, stream);
}
