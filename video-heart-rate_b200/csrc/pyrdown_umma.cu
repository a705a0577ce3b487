// Fused 4-level pyrDown cascade on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
// uint8 (T,H,W,3) -> float32 (T,h4,w4,3); W % 80 == 0, 16-byte aligned frames, levels == 4.  Every other shape takes
// pyrdown_stream.cu / pyrdown.cu.  Same spec as pyrdown.cu (cv2.pyrDown float32 semantics; no
// reference code exists for this stage; oracle/evm.py:pyrdown_cascade).
//
// Why: the streaming kernel (pyrdown_stream.cu) is instruction-issue bound at 49 % of the HBM roofline (~220
// instructions per level-1 row and warp, profiles/r1_ncu_p_*), and the mma.sync attempts (profiles/README.md, round 2)
// only moved the load to the LSU, because every register fragment goes shared memory -> registers.  tcgen05.mma reads
// both operands from shared memory itself, so the full-resolution data never passes through a register:
//
//   * VERTICAL, two levels at once, on the tensor core.  pyrDown o pyrDown along a column is one 13-tap stride-4 filter
//     (1 4 10 20 31 40 44 40 31 20 10 4 1, sum 256; exact in integers).  For a tile of <= 128 level-2 rows it is the
//     GEMM  D[128 level-2 rows x 240 byte columns] += A[128 x 32] . B[32 image rows x 240 bytes]  per 32 input rows:
//     B is the raw image, one TMA box per 128 bytes of width (SWIZZLE_128B = the canonical MN-major UMMA layout, so no
//     transposition anywhere), A is the banded weight slice.  The band is shift-invariant (32 input rows = 8 output
//     rows), so ONE 8 KB zero-padded band in shared memory serves every k-step through the descriptor's start address;
//     the few slices that touch a frame border (reflect-101 folded into the weights by the host, cv2 semantics at both
//     levels) are kept as explicit 4 KB slices.  A k-step only has weights for ~11 consecutive level-2 rows, so it is issued
//     as an M = 64 MMA on the half of the tile that holds them (both halves at the seam); the remaining 11x zero
//     padding costs nothing: the tensor pipe is far from busy at the HBM-bound rate.
//   * HORIZONTAL in registers, on data that is already 4x smaller.  An epilogue thread owns one level-2 row: it reads its
//     240 accumulators (tcgen05.ld) and applies the same 13-tap filter with compile-time weights; 30 accumulator columns
//     and three level-2 pixels are carried from strip to strip, so strips do not overlap (each strip emits the 20 level-2
//     pixels ending one pixel before its right edge; one all-zero flush strip ends the row).  It goes on to the
//     horizontal pass of level 3 (10 pixels per strip) and hands those to
//   * two level-3 warps (vertical pass of level 3 from shared memory, horizontal pass of level 4 in registers) and one
//     level-4 warp (vertical pass, store).  Rows are cut into tiles of <= 29 level-4 rows whose level-3 / level-2 / input
//     rows are recomputed at the seams (<= 16 % more shared-memory fill, no extra HBM traffic: the seams hit L2), which
//     makes every tile self-contained: no inter-CTA hand-over, no scratch in HBM.
//
// Warp roles (288 threads, 1 CTA per SM, persistent over (frame, tile) items): 0 TMA producer, 1 MMA issuer + TMEM
// owner, 2-5 level-2 warps (TMEM lane quarter = warp % 4), 6-7 level-3 warps, 8 level-4 warp.  Bounded buffers with
// full / empty mbarriers between every pair of stages; TMEM holds two 240-column accumulators.
// Exactness: level 2 is an exact integer (< 2^24); levels 3-4 accumulate in float32 (tests: <= 1e-4 of full scale,
// measured ~2e-7).
// Two things the ncu captures taught (profiles/README.md, round 2): (1) the single-thread role loops must be cheap -- they
// run warp-uniform so that descriptors and barrier addresses stay in uniform registers and one elected lane issues; with
// ~50 instructions per k-step on one thread the whole kernel waited for its issuers.  (2) The epilogue must be small code:
// five unrolled chunk copies per strip (19 KB) starved the level-2 warps on instruction fetch; one instance, five trips.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <vector>

namespace {

#ifndef UMMA_NSTAGE
#define UMMA_NSTAGE 9
#endif
constexpr int NSTAGE = UMMA_NSTAGE;     // image stages of 64 rows x 256 bytes = two k-steps (6 / 8 / 9 stages: 1.892 / 1.860 / 1.854 ms)
constexpr int STAGE_BYTES = 16384;
constexpr int PX = 20;                  // level-2 pixels per strip
constexpr int NB = 12 * PX;             // fresh bytes per strip = MMA N
constexpr int MAX_N4 = 29;              // level-4 rows per tile (4 n + 9 level-2 rows <= 128)
constexpr int MAX_TILES = 12;
constexpr int MAX_KS = 18;              // k-steps per tile, even (band offsets 0..16; a 17th / 18th step only ever meets unused rows)
#ifndef UMMA_MAX_SPECIAL
#define UMMA_MAX_SPECIAL 6
#endif
constexpr int MAX_SPECIAL = UMMA_MAX_SPECIAL;
constexpr int BAND_BYTES = 8192;
constexpr int SLICE_BYTES = 4096;
constexpr int L3H_PITCH = 144;          // floats per column: even rows at 0.., odd rows at 80.. (conflict-free both ways)
constexpr int L3H_ODD = 80;
constexpr int L3H_COLS = 30;
constexpr int L4H_PITCH = 80;
constexpr int L4H_ODD = 48;
constexpr int L4H_COLS = 15;
constexpr int THREADS = 288;

constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + NSTAGE * STAGE_BYTES;
constexpr int OFF_L3H = OFF_A + BAND_BYTES + MAX_SPECIAL * SLICE_BYTES;
constexpr int OFF_L4H = OFF_L3H + 2 * L3H_COLS * L3H_PITCH * 4;
constexpr int OFF_BAR = OFF_L4H + 2 * L4H_COLS * L4H_PITCH * 4;
constexpr int NBAR = 2 * NSTAGE + 12;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;     // + slack to align the dynamic base to 1024

struct UmmaTile {
    int a, n4;        // first level-4 row, rows
    int g0, n3;       // first level-3 row held by the tile, rows (<= 64)
    int r0, nr;       // first level-2 row, rows (<= 128)
    int i0, nks;      // first input row fetched (may be negative), k-steps of 32 rows (even: a stage holds two)
};

struct UmmaArgs {
    float* out;
    int T, H, W;
    int h2, h3, h4, w4;
    int nstrips, ntiles;
    long long items;                      // T * ntiles
    const uint8_t* blob;                  // band + special slices (global), blob_bytes
    int blob_bytes;
    UmmaTile tile[MAX_TILES];
    alignas(4) unsigned short code[MAX_TILES][MAX_KS];   // per k-step (read in pairs = one stage): bits 0-11 byte offset >> 4 of the
                                                         // 128-row A slice inside the blob, bits 12-13 which 64-row halves carry
                                                         // weights, bits 14-15 which of them this k-step writes first in a strip
    int wsp[3][13];                       // horizontal weights of level-2 pixels 0, 1 and w2 - 1 (window coordinates)
    uint32_t* dbg;                        // optional: raw accumulators of (dbg_item, dbg_strip), 128 x 240
    int dbg_item, dbg_strip;
    int mode;                             // measurement hook (VHR_UMMA_MODE): bit 0 = level-2 warps skip their arithmetic, bit 1 = no MMA is issued
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
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
        if (spin > 4000000) {
            if ((threadIdx.x & 31) == 0)
                printf("WATCHDOG umma block %d warp %d bar_off %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar & 0xffffu, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
#endif
__device__ __forceinline__ void tma_box3d(uint32_t dst, const CUtensorMap* tmap, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                   "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld48(uint32_t taddr, uint32_t* d) {
    tmem_ld16(taddr, d);
    tmem_ld16(taddr + 16, d + 16);
    tmem_ld16(taddr + 32, d + 32);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout: start >> 4 at [0,14), leading byte offset
// >> 4 at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48), layout type at [61,64)).
//   A: K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); the two 16-byte K chunks of a
//      k-step 128 bytes apart (leading), 8-row groups 256 bytes apart (stride).
//   B: MN-major, SWIZZLE_128B: 128 bytes of N contiguous per K row (what a TMA box row is), 8-row groups 1024 bytes
//      apart (stride), the second 128 bytes of N = the second box, 8192 bytes on (leading); the second k-step of a
//      stage starts 4096 bytes into each box.
constexpr uint64_t A_DESC_HI = (uint64_t)((256u >> 4) | (1u << 14)) << 32 | (uint64_t)(128u >> 4) << 16;
constexpr uint64_t B_DESC_HI = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32 | (uint64_t)(8192u >> 4) << 16;
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 (2 at [4,6)), A = B = unsigned 8 bit (0 at [7,10) and
// [10,13)), A K-major (0 at 15), B MN-major (1 at 16), N >> 3 at [17,23), M >> 4 at [24,29).
// M = 64: a k-step (32 input rows) only has weights for ~11 consecutive level-2 rows, i.e. for one 64-row half of the
// tile (two at the seam), so only that half is multiplied: half the tensor work (and power) of the M = 128 form.  An
// M = 64 accumulator occupies 16 lanes of each TMEM lane quarter: row i of half h sits in lane 32 (i / 16) + 16 h + i % 16.
constexpr uint32_t IDESC = (2u << 4) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((64u >> 4) << 24);

__device__ __forceinline__ int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i < 0 ? -i : i;
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// ---- level-2 warps: one chunk of 48 accumulator columns = 4 level-2 pixels = 2 level-3 pixels of the row ------------
// One code instance for the five chunks of a strip (the unrolled form, 19 KB per strip, starved on instruction fetch:
// 66 % of the level-2 warps' samples were no_instruction, profiles/r2_ncu_v_pyrdown_umma_v2.txt).
// v = 30 carried columns + the 48 of this chunk; level-2 slot e of the chunk sits at v[12 e + 3 j + ch], j = 0..12.
__device__ __forceinline__ void l2_values(const uint32_t (&v)[78], float (&l2)[4][3]) {
#pragma unroll
    for (int el = 0; el < 4; ++el) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const int b = 12 * el + ch;
            const uint32_t acc = (v[b] + v[b + 36]) + 4u * (v[b + 3] + v[b + 33]) + 10u * (v[b + 6] + v[b + 30]) +
                                 20u * (v[b + 9] + v[b + 27]) + 31u * (v[b + 12] + v[b + 24]) + 40u * (v[b + 15] + v[b + 21]) +
                                 44u * v[b + 18];
            l2[el][ch] = __uint2float_rn(acc);
        }
    }
}
// frame borders (first chunk of the first / of the flush strip only): level-2 pixels 0, 1 (slots 1, 2 of the first strip)
// and w2 - 1 (slot 0 of the flush strip) carry reflect-101 folded into their weights; the virtual pixels -1, -2 and w2
// mirror their neighbours
__device__ __forceinline__ void l2_borders(const UmmaArgs& a, const uint32_t (&v)[78], float (&l2)[4][3], float (&p3)[3][3], const bool first) {
    if (first) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 13; ++j) acc += (uint32_t)a.wsp[k][j] * v[12 * (k + 1) + 3 * j + ch];
                l2[k + 1][ch] = __uint2float_rn(acc);
            }
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { l2[0][ch] = l2[2][ch]; p3[2][ch] = l2[3][ch]; }
    } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 13; ++j) acc += (uint32_t)a.wsp[2][j] * v[3 * j + ch];
            l2[0][ch] = __uint2float_rn(acc);
            l2[1][ch] = p3[2][ch];
        }
    }
}

template <bool DBG>
__global__ void __launch_bounds__(THREADS, 1) pyrdown_umma_kernel(const __grid_constant__ UmmaArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;     // warp-uniform for the compiler
    const uint32_t bar0 = sbase + OFF_BAR;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * NSTAGE + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 + b); };
    auto bar_l3full = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 4 + b); };
    auto bar_l3empty = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 6 + b); };
    auto bar_l4full = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 8 + b); };
    auto bar_l4empty = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 10 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

    // weight band + border slices: global -> shared (generic proxy), then made visible to the tensor core (async proxy)
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.blob);
        uint4* dst = reinterpret_cast<uint4*>(smem + OFF_A);
        for (int i = threadIdx.x; i < a.blob_bytes / 16; i += THREADS) dst[i] = __ldg(src + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 128);
            mbar_init(bar_l3full(b), 128); mbar_init(bar_l3empty(b), 64);
            mbar_init(bar_l4full(b), 64); mbar_init(bar_l4empty(b), 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(sbase + OFF_TMEM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int S = a.nstrips;

    if (warp == 0) {
        // ---- TMA producer: the whole warp runs the loop with warp-uniform values (operands stay in uniform registers),
        // one elected lane issues.  One stage = 64 input rows x 256 bytes = two boxes.
        if (lane == 0) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap) : "memory");
        uint32_t st = 0, ph = 0;
        for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
            const int f = (int)(item / a.ntiles);
            const int t = (int)(item - (long long)f * a.ntiles);
            const int i0 = a.tile[t].i0, nkp = a.tile[t].nks >> 1;
            for (int s = 0; s < S; ++s) {
                const int x = NB * s;
                for (int kp = 0; kp < nkp; ++kp) {
                    mbar_wait(bar_empty(st), ph ^ 1);
                    if (elect_one()) {
                        const uint32_t dst = sbase + OFF_B + st * STAGE_BYTES;
                        mbar_expect_tx(bar_full(st), STAGE_BYTES);
                        tma_box3d(dst, &tmap, x, i0 + 64 * kp, f, bar_full(st));
                        tma_box3d(dst + 8192, &tmap, x + 128, i0 + 64 * kp, f, bar_full(st));
                    }
                    __syncwarp();
                    if (++st == NSTAGE) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (same style): two MMAs (k-steps) per stage ------------------------------------------------------
        uint32_t st = 0, ph = 0, sc = 0;
        const uint32_t a_base = (sbase + OFF_A) >> 4, b_base = (sbase + OFF_B) >> 4;
        for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
            const int f = (int)(item / a.ntiles);
            const int t = (int)(item - (long long)f * a.ntiles);
            const int nkp = a.tile[t].nks >> 1;
            const uint32_t* codes = reinterpret_cast<const uint32_t*>(&a.code[t][0]);
            for (int s = 0; s < S; ++s, ++sc) {
                const uint32_t buf = sc & 1;
                mbar_wait(bar_tempty(buf), ((sc >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + 256u * buf;
                for (int kp = 0; kp < nkp; ++kp) {
                    const uint32_t cc = codes[kp];
                    mbar_wait(bar_full(st), ph);
                    tc_fence_after();
                    if (elect_one() && !(a.mode & 2)) {
                        const uint32_t b0 = b_base + st * (STAGE_BYTES >> 4);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {                    // the stage's two k-steps
                            const uint32_t c = j ? cc >> 16 : cc & 0xFFFFu;
                            const uint64_t bdesc = B_DESC_HI | (uint64_t)((b0 + 256u * j) & 0x3FFFu);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {                // the tile's two 64-row halves
                                if ((c >> (12 + h)) & 1u)
                                    tc_mma_i8(d_tmem + ((uint32_t)(16 * h) << 16),
                                              A_DESC_HI | (uint64_t)((a_base + (c & 0xFFFu) + 128u * h) & 0x3FFFu), bdesc, IDESC,
                                              ((c >> (14 + h)) & 1u) ^ 1u);
                            }
                        }
                    }
                    if (elect_one()) tc_commit(bar_empty(st));
                    __syncwarp();
                    if (++st == NSTAGE) { st = 0; ph ^= 1; }
                }
                if (elect_one()) tc_commit(bar_tfull(buf));
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ---- level-2 warps: accumulators -> level 2 (13-tap) -> horizontal pass of level 3 ------------------------------
        const int q = warp & 3;
        const int m = 64 * (lane >> 4) + 16 * q + (lane & 15);           // TMEM lane 32 q + lane holds this level-2 row (M = 64 halves)
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const uint32_t ridx = 4u * ((m & 1) * L3H_ODD + (m >> 1));
        uint32_t sc = 0;
        for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
            const int f = (int)(item / a.ntiles);
            const int t = (int)(item - (long long)f * a.ntiles);
            const bool active = 16 * q < a.tile[t].nr;
            uint32_t v[78];                      // [0, 30): columns carried from the previous chunk, [30, 78): this chunk
            float p3[3][3];                      // the three level-2 pixels before this chunk
#pragma unroll
            for (int i = 0; i < 78; ++i) v[i] = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) { p3[i][0] = 0.f; p3[i][1] = 0.f; p3[i][2] = 0.f; }
            for (int s = 0; s < S; ++s, ++sc) {
                const int buf = sc & 1;
                const uint32_t ph = (sc >> 1) & 1;
                uint32_t dst = sbase + OFF_L3H + buf * (L3H_COLS * L3H_PITCH * 4) + ridx;
                mbar_wait(bar_l3empty(buf), ph ^ 1);
                mbar_wait(bar_tfull(buf), ph);
                tc_fence_after();
                if (active) {
                    uint32_t ta = tlane + 256u * buf;
                    const bool border = s == 0 || s == S - 1;
                    tmem_ld48(ta, v + 30);
#pragma unroll 1
                    for (int c = 0; c < 5; ++c) {
                        tmem_wait_ld();
                        if (DBG) {
                            if (a.dbg && item == a.dbg_item && s == a.dbg_strip) {
#pragma unroll
                                for (int i = 0; i < 48; ++i) a.dbg[m * NB + 48 * c + i] = v[30 + i];
                            }
                        }
                        float l2[4][3];
                        if (!(a.mode & 1)) l2_values(v, l2);
                        if (c == 0 && border) l2_borders(a, v, l2, p3, s == 0);
#pragma unroll
                        for (int i = 0; i < 30; ++i) v[i] = v[48 + i];
                        ta += 48;
                        if (c < 4) {
                            tmem_ld48(ta, v + 30);
                        } else {
                            tc_fence_before();
                            mbar_arrive(bar_tempty(buf));
                        }
                        // horizontal pass of level 3: pixels 2c and 2c+1 of the strip from level-2 slots 4c-3 .. 4c+3
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float e0 = p3[0][ch], e1 = p3[1][ch], e2 = p3[2][ch], e3 = l2[0][ch], e4 = l2[1][ch], e5 = l2[2][ch], e6 = l2[3][ch];
                            sts_f32(dst + ch * (L3H_PITCH * 4), (e0 + e4) + 4.0f * (e1 + e3) + 6.0f * e2);
                            sts_f32(dst + (3 + ch) * (L3H_PITCH * 4), (e2 + e6) + 4.0f * (e3 + e5) + 6.0f * e4);
                            p3[0][ch] = e4; p3[1][ch] = e5; p3[2][ch] = e6;
                        }
                        dst += 6 * L3H_PITCH * 4;
                    }
                } else {
                    tc_fence_before();
                    mbar_arrive(bar_tempty(buf));
                }
                mbar_arrive(bar_l3full(buf));
            }
        }
    } else if (warp < 8) {
        // ---- level-3 warps: vertical pass of level 3, horizontal pass of level 4 ------------------------------------------
        const int i = (warp - 6) * 32 + lane;
        const uint32_t widx = 4u * ((i & 1) * L4H_ODD + (i >> 1));
        uint32_t sc = 0;
        for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
            const int f = (int)(item / a.ntiles);
            const UmmaTile& tl = a.tile[(int)(item - (long long)f * a.ntiles)];
            const int g = tl.g0 + min(i, tl.n3 - 1);
            int idx[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int r = refl101(2 * g - 2 + d, a.h2) - tl.r0;
                idx[d] = 4 * ((r & 1) * L3H_ODD + (r >> 1));
            }
            float c3[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { c3[k][0] = 0.f; c3[k][1] = 0.f; c3[k][2] = 0.f; }
            for (int s = 0; s < S; ++s, ++sc) {
                const int buf = sc & 1;
                const uint32_t ph = (sc >> 1) & 1;
                const uint32_t src = sbase + OFF_L3H + buf * (L3H_COLS * L3H_PITCH * 4);
                mbar_wait(bar_l3full(buf), ph);
                float l3[13][3];                    // slots -3 .. 9
#pragma unroll
                for (int e = 0; e < 10; ++e) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const uint32_t p = src + (e * 3 + ch) * (L3H_PITCH * 4);
                        l3[e + 3][ch] = (lds_f32(p + idx[0]) + lds_f32(p + idx[4])) + 4.0f * (lds_f32(p + idx[1]) + lds_f32(p + idx[3])) +
                                        6.0f * lds_f32(p + idx[2]);
                    }
                }
                mbar_arrive(bar_l3empty(buf));
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (s == 0) { l3[3][ch] = l3[5][ch]; c3[2][ch] = l3[6][ch]; }
                    if (s == S - 1) l3[4][ch] = c3[2][ch];
                    l3[0][ch] = c3[0][ch]; l3[1][ch] = c3[1][ch]; l3[2][ch] = c3[2][ch];
                }
                const uint32_t dst = sbase + OFF_L4H + buf * (L4H_COLS * L4H_PITCH * 4) + widx;
                mbar_wait(bar_l4empty(buf), ph ^ 1);
#pragma unroll
                for (int e = 0; e < 5; ++e) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        sts_f32(dst + (e * 3 + ch) * (L4H_PITCH * 4),
                                (l3[2 * e][ch] + l3[2 * e + 4][ch]) + 4.0f * (l3[2 * e + 1][ch] + l3[2 * e + 3][ch]) + 6.0f * l3[2 * e + 2][ch]);
                }
                mbar_arrive(bar_l4full(buf));
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) { c3[0][ch] = l3[10][ch]; c3[1][ch] = l3[11][ch]; c3[2][ch] = l3[12][ch]; }
            }
        }
    } else {
        // ---- level-4 warp: vertical pass of level 4, store ---------------------------------------------------------------
        uint32_t sc = 0;
        for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
            const int f = (int)(item / a.ntiles);
            const UmmaTile& tl = a.tile[(int)(item - (long long)f * a.ntiles)];
            const bool valid = lane < tl.n4;
            const int r4 = tl.a + min(lane, tl.n4 - 1);
            int idx[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int r = refl101(2 * r4 - 2 + d, a.h3) - tl.g0;
                idx[d] = 4 * ((r & 1) * L4H_ODD + (r >> 1));
            }
            float* orow = a.out + ((size_t)f * a.h4 + r4) * (size_t)a.w4 * 3;
            for (int s = 0; s < S; ++s, ++sc) {
                const int buf = sc & 1;
                const uint32_t ph = (sc >> 1) & 1;
                const uint32_t src = sbase + OFF_L4H + buf * (L4H_COLS * L4H_PITCH * 4);
                mbar_wait(bar_l4full(buf), ph);
                float o[15];
#pragma unroll
                for (int c = 0; c < 15; ++c) {
                    const uint32_t p = src + c * (L4H_PITCH * 4);
                    o[c] = ((lds_f32(p + idx[0]) + lds_f32(p + idx[4])) + 4.0f * (lds_f32(p + idx[1]) + lds_f32(p + idx[3])) +
                            6.0f * lds_f32(p + idx[2])) * 2.3283064365386963e-10f;   // 2^-32
                }
                mbar_arrive(bar_l4empty(buf));
                if (valid) {
                    float* o0 = orow + (5 * s - 1) * 3;
                    const int e_lo = s == 0 ? 1 : 0, e_hi = s == S - 1 ? 1 : 5;
#pragma unroll
                    for (int e = 0; e < 5; ++e) {
                        if (e >= e_lo && e < e_hi) { o0[3 * e] = o[3 * e]; o0[3 * e + 1] = o[3 * e + 1]; o0[3 * e + 2] = o[3 * e + 2]; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

// ---- host: tile plan, baked weight slices ---------------------------------------------------------------------------
const int W5[5] = {1, 4, 6, 4, 1};
const int W13[13] = {1, 4, 10, 20, 31, 40, 44, 40, 31, 20, 10, 4, 1};

int hrefl(int i, int n) { return vhr_reflect101(i, n); }

// weights of level-2 sample r over the level-0 samples (reflect-101 at both levels): dense over [lo, lo + 16)
struct Comp { int lo; int w[24]; };
Comp composite_row(int r, int n1, int n0) {
    Comp c;
    c.lo = 4 * r - 8;
    for (int k = 0; k < 24; ++k) c.w[k] = 0;
    for (int d2 = 0; d2 < 5; ++d2) {
        const int l1 = hrefl(2 * r - 2 + d2, n1);
        for (int d1 = 0; d1 < 5; ++d1) {
            const int l0 = hrefl(2 * l1 - 2 + d1, n0);
            const int k = l0 - c.lo;
            if (k >= 0 && k < 24) c.w[k] += W5[d2] * W5[d1];
            else c.lo = -1000000;      // cannot happen for n0 >= 8 (checked by the caller through the weight sum)
        }
    }
    return c;
}

struct UmmaPlan {
    int H = 0, W = 0;
    int h[5], w[5];
    int ntiles = 0, nstrips = 0, nspecial = 0;
    UmmaTile tile[MAX_TILES];
    unsigned short code[MAX_TILES][MAX_KS];
    unsigned char half[MAX_TILES][MAX_KS];      // bit h: the k-step has weights for rows 64 h .. 64 h + 63
    int wsp[3][13];
    std::vector<uint8_t> blob;
};

inline int canon(int m, int k) { return (m >> 3) * 256 + (k >> 4) * 128 + (m & 7) * 16 + (k & 15); }

// tiles of n4 level-4 rows (the last one takes the rest); VHR_OK, or VHR_ERR_UNSUPPORTED when the shape is not eligible
int make_plan_n4(int H, int W, int n4, UmmaPlan& p) {
    if (W % 80 != 0 || W < 160 || H < 64 || W > 16384 || H > 16384) return VHR_ERR_UNSUPPORTED;
    p.H = H; p.W = W;
    p.h[0] = H; p.w[0] = W;
    for (int l = 1; l <= 4; ++l) { p.h[l] = (p.h[l - 1] + 1) / 2; p.w[l] = (p.w[l - 1] + 1) / 2; }
    if (p.h[3] < 3 || p.w[3] < 3) return VHR_ERR_UNSUPPORTED;
    p.ntiles = (p.h[4] + n4 - 1) / n4;
    if (p.ntiles > MAX_TILES || n4 > MAX_N4) return VHR_ERR_UNSUPPORTED;
    p.nstrips = p.w[2] / PX + 1;
    p.blob.assign(BAND_BYTES, 0);
    for (int q = -128; q < 128; ++q)
        for (int k = 0; k < 32; ++k) {
            const int j = k - 4 * q - 2;
            if (j >= 0 && j <= 12) p.blob[canon(q + 128, k)] = (uint8_t)W13[j];
        }
    p.nspecial = 0;
    for (int t = 0; t < p.ntiles; ++t) {
        UmmaTile& tl = p.tile[t];
        tl.a = t * n4;
        tl.n4 = std::min(n4, p.h[4] - tl.a);
        if (tl.n4 < 1) return VHR_ERR_UNSUPPORTED;
        int g0 = 1 << 30, g1 = -1, r0 = 1 << 30, r1 = -1;
        for (int r = tl.a; r < tl.a + tl.n4; ++r)
            for (int d = 0; d < 5; ++d) { const int g = hrefl(2 * r - 2 + d, p.h[3]); g0 = std::min(g0, g); g1 = std::max(g1, g); }
        for (int g = g0; g <= g1; ++g)
            for (int d = 0; d < 5; ++d) { const int r = hrefl(2 * g - 2 + d, p.h[2]); r0 = std::min(r0, r); r1 = std::max(r1, r); }
        tl.g0 = g0; tl.n3 = g1 - g0 + 1; tl.r0 = r0; tl.nr = r1 - r0 + 1;
        if (tl.nr > 128 || tl.n3 > 64 || tl.n4 > 32) return VHR_ERR_UNSUPPORTED;
        tl.i0 = 4 * r0 - 8;
        int hi = 0;
        std::vector<Comp> rows(tl.nr);
        for (int m = 0; m < tl.nr; ++m) {
            rows[m] = composite_row(r0 + m, p.h[1], H);
            int sum = 0;
            for (int k = 0; k < 24; ++k) {
                if (rows[m].w[k]) { hi = std::max(hi, rows[m].lo + k); sum += rows[m].w[k]; }
                if (rows[m].w[k] > 255) return VHR_ERR_UNSUPPORTED;
            }
            if (rows[m].lo < -100000 || sum != 256) return VHR_ERR_UNSUPPORTED;
        }
        tl.nks = 2 * ((hi - tl.i0 + 1 + 63) / 64);           // whole stages of 64 rows; a trailing k-step carries zero weights
        if (tl.nks > MAX_KS || tl.nks < 2) return VHR_ERR_UNSUPPORTED;
        for (int ks = 0; ks < tl.nks; ++ks) {
            bool generic = true;
            uint8_t sl[SLICE_BYTES];
            memset(sl, 0, sizeof(sl));
            for (int m = 0; m < tl.nr; ++m)
                for (int k = 0; k < 32; ++k) {
                    const int row = tl.i0 + 32 * ks + k;              // input row
                    const int kk = row - rows[m].lo;
                    const int wv = (row >= 0 && row < H && kk >= 0 && kk < 24) ? rows[m].w[kk] : 0;
                    const int j = 32 * ks + k - 4 * m - 2;
                    const int gv = (j >= 0 && j <= 12) ? W13[j] : 0;
                    if (wv != gv) generic = false;
                    sl[canon(m, k)] = (uint8_t)wv;
                }
            // which 64-row halves of the tile this k-step has weights for (rows past the tile do not count)
            unsigned halves = 0;
            for (int m = 0; m < tl.nr; ++m)
                for (int k = 0; k < 32; ++k)
                    if (sl[canon(m, k)]) halves |= 1u << (m >> 6);
            p.half[t][ks] = (unsigned char)halves;
            if (generic && ks <= 16) {
                p.code[t][ks] = (unsigned short)(((16 - ks) * 256) >> 4);
            } else {
                int found = -1;
                for (int q = 0; q < p.nspecial && found < 0; ++q)
                    if (memcmp(p.blob.data() + BAND_BYTES + q * SLICE_BYTES, sl, SLICE_BYTES) == 0) found = q;
                if (found < 0) {
                    if (p.nspecial == MAX_SPECIAL) return VHR_ERR_UNSUPPORTED;
                    found = p.nspecial++;
                    p.blob.insert(p.blob.end(), sl, sl + SLICE_BYTES);
                }
                p.code[t][ks] = (unsigned short)((BAND_BYTES + found * SLICE_BYTES) >> 4);
            }
        }
    }
    // horizontal: every level-2 pixel but 0, 1 and w2 - 1 must be the plain 13-tap
    const int special_px[3] = {0, 1, p.w[2] - 1};
    for (int x = 0; x < p.w[2]; ++x) {
        Comp c = composite_row(x, p.w[1], W);
        int wv[13], sum = 0;
        bool plain = true;
        for (int j = 0; j < 13; ++j) {
            const int px = 4 * x - 6 + j;
            const int k = px - c.lo;
            wv[j] = (px >= 0 && px < W && k >= 0 && k < 24) ? c.w[k] : 0;
            sum += wv[j];
            if (wv[j] != W13[j]) plain = false;
        }
        if (c.lo < -100000 || sum != 256) return VHR_ERR_UNSUPPORTED;
        int which = -1;
        for (int k = 0; k < 3; ++k) if (x == special_px[k]) which = k;
        if (which >= 0) for (int j = 0; j < 13; ++j) p.wsp[which][j] = wv[j];
        else if (!plain) return VHR_ERR_UNSUPPORTED;
    }
    return VHR_OK;
}

// The split with the fewest 64-row stages per frame (ties: the smaller tile height).
int make_plan(int H, int W, UmmaPlan& best) {
    if (H < 32) return VHR_ERR_UNSUPPORTED;
    const int h4 = (((((H + 1) / 2 + 1) / 2 + 1) / 2) + 1) / 2;
    const int ntiles = (h4 + MAX_N4 - 1) / MAX_N4;
    int best_cost = 1 << 30, rc_any = VHR_ERR_UNSUPPORTED;
    for (int n4 = (h4 + ntiles - 1) / ntiles; n4 <= MAX_N4; ++n4) {
        UmmaPlan p;
        const int rc = make_plan_n4(H, W, n4, p);
        if (rc != VHR_OK) continue;
        int cost = 0;
        for (int t = 0; t < p.ntiles; ++t) cost += p.tile[t].nks;
        if (cost < best_cost) { best_cost = cost; best = p; rc_any = VHR_OK; }
    }
    return rc_any;
}

}  // namespace

// Diagnostics (tests): the tile plan of a frame shape.  tiles: ntiles x 8 int32 (a, n4, g0, n3, r0, nr, i0, nks); codes:
// ntiles x 18 (A-slice offset >> 4 inside the blob); wsp: 3 x 13; meta: ntiles, nstrips, nspecial, blob bytes.
// blob may be NULL.  Host only.
extern "C" int vhr_pyrdown_umma_plan(int H, int W, int32_t* tiles, int32_t* codes, int32_t* wsp, int32_t* meta, uint8_t* blob,
                                     int blob_cap) {
    UmmaPlan p;
    const int rc = make_plan(H, W, p);
    if (rc != VHR_OK) return rc;
    for (int t = 0; t < p.ntiles; ++t) {
        const UmmaTile& tl = p.tile[t];
        const int v[8] = {tl.a, tl.n4, tl.g0, tl.n3, tl.r0, tl.nr, tl.i0, tl.nks};
        for (int k = 0; k < 8; ++k) tiles[8 * t + k] = v[k];
        for (int ks = 0; ks < MAX_KS; ++ks) codes[MAX_KS * t + ks] = ks < tl.nks ? p.code[t][ks] : -1;
    }
    for (int k = 0; k < 39; ++k) wsp[k] = p.wsp[k / 13][k % 13];
    meta[0] = p.ntiles; meta[1] = p.nstrips; meta[2] = p.nspecial; meta[3] = (int)p.blob.size();
    if (blob) {
        if ((int)p.blob.size() > blob_cap) return VHR_ERR_INVALID;
        memcpy(blob, p.blob.data(), p.blob.size());
    }
    return VHR_OK;
}

static int umma_launch(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level, cudaStream_t stream,
                       uint32_t* d_acc, int acc_item, int acc_strip) {
    if (levels != 4 || (reinterpret_cast<uintptr_t>(d_frames) & 15) != 0) return VHR_ERR_UNSUPPORTED;
    if (SMEM_BYTES > ctx->smem_optin) return VHR_ERR_UNSUPPORTED;
    // plan cached per thread and shape, blob cached per context and shape
    static thread_local UmmaPlan plan;
    if (plan.H != H || plan.W != W) {
        UmmaPlan p;
        const int rc = make_plan(H, W, p);
        if (rc != VHR_OK) return rc;
        plan = p;
    }
    const long long key = ((long long)H << 32) | (unsigned)W;
    const uint8_t* d_blob = nullptr;
    for (int i = 0; i < ctx->umma_n && !d_blob; ++i)
        if (ctx->umma_key[i] == key) d_blob = static_cast<const uint8_t*>(ctx->umma_blob[i]);
    if (!d_blob) {
        int slot = ctx->umma_n;
        if (slot == vhr_ctx::UMMA_SLOTS) {                 // table full: nothing may still be reading the entry we replace
            VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
            slot = ctx->umma_next;
            ctx->umma_next = (ctx->umma_next + 1) % vhr_ctx::UMMA_SLOTS;
        } else {
            VHR_CHECK_CUDA(ctx, cudaMalloc(&ctx->umma_blob[slot], BAND_BYTES + MAX_SPECIAL * SLICE_BYTES));
            ctx->umma_n = slot + 1;
        }
        // synchronous copy into memory no kernel reads yet; kernels launched after this call see it on any stream
        VHR_CHECK_CUDA(ctx, cudaMemcpy(ctx->umma_blob[slot], plan.blob.data(), plan.blob.size(), cudaMemcpyHostToDevice));
        ctx->umma_key[slot] = key;
        d_blob = static_cast<const uint8_t*>(ctx->umma_blob[slot]);
    }
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
    const cuuint64_t gdim[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)T};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * (cuuint64_t)H};
    const cuuint32_t box[3] = {128, 64, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_frames), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return VHR_ERR_UNSUPPORTED;
    UmmaArgs a;
    memset(&a, 0, sizeof(a));
    a.out = d_level;
    a.T = T; a.H = H; a.W = W;
    a.h2 = plan.h[2]; a.h3 = plan.h[3]; a.h4 = plan.h[4]; a.w4 = plan.w[4];
    a.nstrips = plan.nstrips; a.ntiles = plan.ntiles;
    a.items = (long long)T * plan.ntiles;
    a.blob = d_blob;
    a.blob_bytes = (int)plan.blob.size();
    for (int t = 0; t < plan.ntiles; ++t) {
        a.tile[t] = plan.tile[t];
        unsigned seen = 0;
        for (int ks = 0; ks < MAX_KS; ++ks) {
            const unsigned halves = ks < plan.tile[t].nks ? plan.half[t][ks] : 0u;
            const unsigned first = halves & ~seen;          // the MMA that must overwrite the half instead of accumulating
            seen |= halves;
            a.code[t][ks] = (unsigned short)((plan.code[t][ks] & 0xFFFu) | (halves << 12) | (first << 14));
        }
    }
    memcpy(a.wsp, plan.wsp, sizeof(a.wsp));
    a.dbg = d_acc; a.dbg_item = acc_item; a.dbg_strip = acc_strip;
    const char* mode = getenv("VHR_UMMA_MODE");
    a.mode = mode ? atoi(mode) : 0;
    auto kern = a.dbg ? pyrdown_umma_kernel<true> : pyrdown_umma_kernel<false>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    long long grid = ctx->num_sms;
    if (grid > a.items) grid = a.items;
    kern<<<(int)grid, THREADS, SMEM_BYTES, stream>>>(a, tmap);
    return vhr_after_launch(ctx, "pyrdown_umma_kernel");
}

int vhr_pyrdown_umma(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level, cudaStream_t stream) {
    return umma_launch(ctx, d_frames, T, H, W, levels, d_level, stream, nullptr, 0, 0);
}

// Diagnostics (tests): the cascade through the tensor-core kernel, which also copies the raw TMEM accumulators of one
// (item, strip) -- item = frame * tiles + tile -- to d_acc (128 x 240 uint32: the vertical 13-tap sums of the tile's
// level-2 rows over the strip's 240 byte columns).  They are exact integers, so a test can hold the MMA stage (tensor map,
// swizzle, descriptors, baked border weights) to the plan bit for bit.  VHR_ERR_UNSUPPORTED when the shape is not eligible.
extern "C" int vhr_pyrdown_umma_accumulators(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, float* d_level, int item,
                                             int strip, uint32_t* d_acc, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_frames && d_level && d_acc, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1 && item >= 0 && strip >= 0, "bad arguments");
    return umma_launch(ctx, d_frames, T, H, W, 4, d_level, (cudaStream_t)stream, d_acc, item, strip);
}
