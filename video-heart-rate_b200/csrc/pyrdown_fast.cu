// Fast path of the fused pyrDown cascade (W % 16 == 0, 8-byte aligned frames); the generic
// kernel in pyrdown.cu covers every other shape with the same arithmetic.
//
// Same spec as pyrdown.cu (cv2.pyrDown float32 semantics; levels 1-2 exact).  What differs is
// how the work is laid out, driven by the first ncu profile (profiles/README.md: the generic
// kernel spent 55 % of its instructions in the shared-memory steps of levels >= 2 and ~40 %
// of level 1 in address / reflect arithmetic):
//   * HORIZONTAL FIRST at every level.  A thread owns fixed output columns of every level
//     (4 px of level 1, 2 px of level 2, 1 px of levels >= 3) and filters each incoming row
//     horizontally as soon as it exists: level 1 with 4 IDP4A per value on the raw bytes,
//     level 2 with 3 IDP2A per value on planar uint16 pairs, levels >= 3 with 5 float taps.
//   * The horizontally filtered rows of a thread's own columns live in a THREAD-PRIVATE
//     5-row ring in shared memory (level 1: packed registers), so the vertical 5-tap needs no
//     barrier; the only shared data is the newest finished row of each level (double
//     buffered, planar per channel), one barrier per produced row.
//   * Persistent grid over the flattened (frame, final-row) space as before; rows stream
//     top to bottom, each input byte is read from HBM once.
#include "common.cuh"

namespace {

constexpr int HR = 5;     // rows in a private H ring (exactly the vertical footprint)

struct FastArgs {
    const uint8_t* frames;
    float* out;
    int T, H, W, levels;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    long long total_rows;
    int nt[VHR_MAX_LEVELS + 1];        // threads owning columns at level l
    int ring_off[VHR_MAX_LEVELS + 1];  // byte offsets: shared row rings (levels 1..L-1)
    int ring_stride[VHR_MAX_LEVELS + 1];   // elements per channel plane row
    int hring_off[VHR_MAX_LEVELS + 1]; // byte offsets: private H rings (levels 2..L)
};

__host__ __device__ constexpr int fdiv4(int b) { return b >= 0 ? b / 4 : -((-b + 3) / 4); }
__host__ __device__ constexpr uint32_t tap_word_f(int j0, int wi) {
    uint32_t r = 0;
    const int wt[5] = {1, 4, 6, 4, 1};
    for (int d = 0; d < 5; ++d) {
        int b = j0 + 3 * (d - 2);
        int w = fdiv4(b);
        if (w == wi) r |= (uint32_t)wt[d] << (8 * (b - 4 * w));
    }
    return r;
}
template <int O, int WI>
struct TapF {
    __device__ static __forceinline__ uint32_t run(const uint32_t (&wd)[9], uint32_t acc) {
        constexpr uint32_t k = tap_word_f(6 * (O / 3) + (O % 3), WI);
        if (k != 0) acc = __dp4a(wd[WI + 2], k, acc);
        return TapF<O, WI + 1>::run(wd, acc);
    }
};
template <int O>
struct TapF<O, 7> {
    __device__ static __forceinline__ uint32_t run(const uint32_t (&)[9], uint32_t acc) { return acc; }
};
// 12 outputs (pixel m = 0..3, channel c) -> 6 packed registers: hp[2c] = (m0 | m1 << 16),
// hp[2c+1] = (m2 | m3 << 16): ready for planar (per-channel) 8-byte stores
template <int C>
struct HPassF {
    __device__ static __forceinline__ void run(const uint32_t (&wd)[9], uint32_t (&hp)[6]) {
        const uint32_t o0 = TapF<0 + C, -2>::run(wd, 0u), o1 = TapF<3 + C, -2>::run(wd, 0u);
        const uint32_t o2 = TapF<6 + C, -2>::run(wd, 0u), o3 = TapF<9 + C, -2>::run(wd, 0u);
        hp[2 * C] = __byte_perm(o0, o1, 0x5410);
        hp[2 * C + 1] = __byte_perm(o2, o3, 0x5410);
        HPassF<C + 1>::run(wd, hp);
    }
};
template <>
struct HPassF<3> {
    __device__ static __forceinline__ void run(const uint32_t (&)[9], uint32_t (&)[6]) {}
};

// Issue the loads of 9 words = bytes [24i-8, 24i+28) of the row at byte offset rowofs.  Nothing
// here reads a loaded value: the frame-edge words are patched by fix_row_edges() right before
// the row is consumed, so the loads stay in flight across a whole iteration (a predicated-off
// instruction that names a pending register still waits on the scoreboard).
__device__ __forceinline__ void load_row_fast(const uint8_t* __restrict__ tp, unsigned rowofs, int i, int nt1,
                                              uint32_t (&wd)[9]) {
    const uint2* p = reinterpret_cast<const uint2*>(tp + rowofs);     // tp = frame + 24 i
    const uint2 o0 = __ldg(p), o1 = __ldg(p + 1), o2 = __ldg(p + 2);
    wd[2] = o0.x; wd[3] = o0.y; wd[4] = o1.x; wd[5] = o1.y; wd[6] = o2.x; wd[7] = o2.y;
    const uint2 l = __ldg(p - (i > 0 ? 1 : 0));                        // thread 0: harmless in-bounds reload, patched later
    wd[0] = l.x; wd[1] = l.y;
    wd[8] = __ldg(reinterpret_cast<const uint32_t*>(tp + rowofs + (i < nt1 - 1 ? 24 : 20)));
}
__device__ __forceinline__ void fix_row_edges(int i, int nt1, uint32_t (&wd)[9]) {
    if (i == 0) {
        wd[0] = __byte_perm(wd[3], 0, 0x3244);                 // pixels -2,-1 reflect to 2,1
        const uint32_t tt = __byte_perm(wd[2], wd[3], 0x5430);
        wd[1] = __byte_perm(tt, wd[4], 0x3214);
    }
    if (i == nt1 - 1) wd[8] = __byte_perm(wd[6], wd[7], 0x0432);   // pixel W reflects to W-2
}

template <int L>
struct FastStream {
    const FastArgs& a;
    unsigned char* smem;
    const uint8_t* tp;          // frame + 24 * threadIdx.x
    float* out_frame;
    int nextr[VHR_MAX_LEVELS + 1];   // next row to produce per level
    int lastr[VHR_MAX_LEVELS + 1];   // last row this segment needs per level
    uint32_t win[5][6];
    uint32_t pre[2][9];
    // per-thread byte offsets into shared memory, computed once (the row / channel parts of
    // every address are block-uniform and added from uniform registers)
    int tapofs[VHR_MAX_LEVELS + 1][5];   // levels >= 3: byte offsets of the 5 horizontal taps in a ring row of level l-1
    int hps[VHR_MAX_LEVELS + 1];         // plane stride (bytes) of the private H ring of level l
    int rps[VHR_MAX_LEVELS + 1];         // plane stride (bytes) of the shared row ring of level l

    __device__ FastStream(const FastArgs& a_, unsigned char* s) : a(a_), smem(s) {
        const int i = threadIdx.x;
#pragma unroll
        for (int l = 1; l <= L; ++l) {
            hps[l] = a.w[l] * 4;
            rps[l] = a.ring_stride[l] * (l == 1 ? 2 : 4);
            if (l >= 3) {
                const int wp = a.w[l - 1];
#pragma unroll
                for (int d = 0; d < 5; ++d) tapofs[l][d] = 4 * vhr_reflect101(2 * min(i, a.w[l] - 1) - 2 + d, wp);
            }
        }
    }

    __device__ __forceinline__ void begin_segment(int t, int r0, int r1) {
        const uint8_t* frame = a.frames + (size_t)t * a.H * a.W * 3;
        tp = frame + 24 * threadIdx.x;
        out_frame = a.out + (size_t)t * a.h[L] * a.w[L] * 3;
        int f = r0, e = r1 - 1;
        nextr[L] = f; lastr[L] = e;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            f = max(0, 2 * f - 2);
            e = min(a.h[l] - 1, 2 * e + 2);
            nextr[l] = f; lastr[l] = e;
        }
    }

    __device__ __forceinline__ unsigned rowofs(int r) const { return (unsigned)vhr_reflect101(r, a.H) * (unsigned)(a.W * 3); }

    // ---- level >= 2: a finished row `r` of level l-1 sits in its shared ring --------------
    template <int l>
    __device__ __forceinline__ void on_row(int r) {
        const int i = threadIdx.x;
        const bool own = i < a.nt[l];
        // block-uniform pieces of the addresses
        unsigned char* const hbase = smem + a.hring_off[l] + (r % HR) * 3 * hps[l];        // H-ring slot of source row r
        const unsigned char* const rsrc = smem + a.ring_off[l - 1] + (r & 1) * 3 * rps[l - 1];
        // horizontal pass of this thread's columns
        if (own) {
            if constexpr (l == 2) {
                // 2 px x 3 ch from planar uint16 pairs (apron of 4 entries on the left)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const unsigned char* p = rsrc + ch * rps[1] + 8 * i + 4;                 // px 4i-2
                    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(p);
                    const uint2 w12 = *reinterpret_cast<const uint2*>(p + 4);               // px 4i .. 4i+3
                    const uint32_t w3 = *reinterpret_cast<const uint32_t*>(p + 12);
                    uint32_t o0 = __dp2a_lo(w0, 0x0401u, 0u);
                    o0 = __dp2a_lo(w12.x, 0x0406u, o0);
                    o0 = __dp2a_lo(w12.y, 0x0001u, o0);
                    uint32_t o1 = __dp2a_lo(w12.x, 0x0401u, 0u);
                    o1 = __dp2a_lo(w12.y, 0x0406u, o1);
                    o1 = __dp2a_lo(w3, 0x0001u, o1);
                    *reinterpret_cast<uint2*>(hbase + ch * hps[2] + 8 * i) = make_uint2(o0, o1);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const unsigned char* p = rsrc + ch * rps[l - 1];
                    const float t0 = *reinterpret_cast<const float*>(p + tapofs[l][0]);
                    const float t1 = *reinterpret_cast<const float*>(p + tapofs[l][1]);
                    const float t2 = *reinterpret_cast<const float*>(p + tapofs[l][2]);
                    const float t3 = *reinterpret_cast<const float*>(p + tapofs[l][3]);
                    const float t4 = *reinterpret_cast<const float*>(p + tapofs[l][4]);
                    *reinterpret_cast<float*>(hbase + ch * hps[l] + 4 * i) = t2 * 6.0f + (t1 + t3) * 4.0f + t0 + t4;
                }
            }
        }
        // vertical pass: every level-l row whose five source rows are now present
        const int hp = a.h[l - 1];
        while (nextr[l] <= lastr[l] && min(2 * nextr[l] + 2, hp - 1) <= r) {
            const int q = nextr[l];
            // slot byte offsets of the five source rows (block-uniform; reflected at the frame's
            // top / bottom only)
            int so[5];
            const int ss = 3 * hps[l];
            if (2 * q - 2 >= 0 && 2 * q + 2 <= hp - 1) {        // interior: five consecutive rows, newest = 2q+2
                int sl = (2 * q - 2) % HR;
#pragma unroll
                for (int d = 0; d < 5; ++d) { so[d] = sl * ss; sl = (sl == HR - 1) ? 0 : sl + 1; }
            } else {
#pragma unroll
                for (int d = 0; d < 5; ++d) so[d] = (vhr_reflect101(2 * q - 2 + d, hp) % HR) * ss;
            }
            const unsigned char* const hb = smem + a.hring_off[l];
            if (own) {
                if constexpr (l == 2) {
                    unsigned char* const dst = smem + a.ring_off[2] + (q & 1) * 3 * rps[2];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const unsigned char* hc = hb + ch * hps[2] + 8 * i;
                        const uint2 va = *reinterpret_cast<const uint2*>(hc + so[0]);
                        const uint2 vb = *reinterpret_cast<const uint2*>(hc + so[1]);
                        const uint2 vc = *reinterpret_cast<const uint2*>(hc + so[2]);
                        const uint2 vd = *reinterpret_cast<const uint2*>(hc + so[3]);
                        const uint2 ve = *reinterpret_cast<const uint2*>(hc + so[4]);
                        const uint32_t s0 = (va.x + ve.x) + 4u * (vb.x + vd.x) + 6u * vc.x;     // < 2^24: exact
                        const uint32_t s1 = (va.y + ve.y) + 4u * (vb.y + vd.y) + 6u * vc.y;
                        const float f0 = (float)s0 * (1.0f / 65536.0f), f1 = (float)s1 * (1.0f / 65536.0f);
                        if constexpr (L == 2) {
                            float* o = out_frame + ((size_t)q * a.w[2] + 2 * i) * 3 + ch;
                            o[0] = f0; o[3] = f1;
                        } else {
                            *reinterpret_cast<float2*>(dst + ch * rps[2] + 8 * i) = make_float2(f0, f1);
                        }
                    }
                } else {
                    unsigned char* const dst = smem + a.ring_off[l < L ? l : 1] + (q & 1) * 3 * rps[l];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const unsigned char* hc = hb + ch * hps[l] + 4 * i;
                        const float s = *reinterpret_cast<const float*>(hc + so[2]) * 6.0f +
                                        (*reinterpret_cast<const float*>(hc + so[1]) + *reinterpret_cast<const float*>(hc + so[3])) * 4.0f +
                                        *reinterpret_cast<const float*>(hc + so[0]) + *reinterpret_cast<const float*>(hc + so[4]);
                        const float v = s * (1.0f / 256.0f);
                        if constexpr (l == L) out_frame[((size_t)q * a.w[l] + i) * 3 + ch] = v;
                        else *reinterpret_cast<float*>(dst + ch * rps[l] + 4 * i) = v;
                    }
                }
            }
            nextr[l] = q + 1;
            if constexpr (l < L) {
                __syncthreads();                    // row q of level l visible to the neighbours
                on_row<l + 1>(q);
            }
        }
    }

    // ---- level 1 -----------------------------------------------------------------------------
    __device__ __forceinline__ void run_segment() {
        const int i = threadIdx.x;
        const bool own = i < a.nt[1];
        const int first = nextr[1], last = lastr[1];
        if (own) {
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                uint32_t wd[9];
                load_row_fast(tp, rowofs(2 * first - 2 + d), i, a.nt[1], wd);
                fix_row_edges(i, a.nt[1], wd);
                HPassF<0>::run(wd, win[d]);
            }
        }
        for (int r = first; r <= last; ++r) {
            uint32_t v[6];
            if (own) {
                if (r != first) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { win[0][k] = win[2][k]; win[1][k] = win[3][k]; win[2][k] = win[4][k]; }
                    fix_row_edges(i, a.nt[1], pre[0]);
                    fix_row_edges(i, a.nt[1], pre[1]);
                    HPassF<0>::run(pre[0], win[3]);
                    HPassF<0>::run(pre[1], win[4]);
                }
                if (r < last) {     // rows entering the window next: in flight across the upper levels' work
                    load_row_fast(tp, rowofs(2 * r + 3), i, a.nt[1], pre[0]);
                    load_row_fast(tp, rowofs(2 * r + 4), i, a.nt[1], pre[1]);
                }
                // vertical pass, packed 16-bit lanes (max 65280: no carry between halves)
#pragma unroll
                for (int k = 0; k < 6; ++k) v[k] = (win[0][k] + win[4][k]) + ((win[1][k] + win[3][k]) << 2) + win[2][k] * 6u;
                if constexpr (L == 1) {
                    float* o = out_frame + ((size_t)r * a.w[1] + 4 * i) * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        o[ch] = (float)(v[2 * ch] & 0xFFFFu) * (1.0f / 256.0f);
                        o[3 + ch] = (float)(v[2 * ch] >> 16) * (1.0f / 256.0f);
                        o[6 + ch] = (float)(v[2 * ch + 1] & 0xFFFFu) * (1.0f / 256.0f);
                        o[9 + ch] = (float)(v[2 * ch + 1] >> 16) * (1.0f / 256.0f);
                    }
                } else {
                    unsigned char* const dst = smem + a.ring_off[1] + (r & 1) * 3 * rps[1] + 8 * i;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        unsigned char* p = dst + ch * rps[1];
                        *reinterpret_cast<uint2*>(p + 8) = make_uint2(v[2 * ch], v[2 * ch + 1]);        // px 4i .. 4i+3 (apron of 4)
                        // reflect-101 aprons: px -2,-1 <- px 2,1 ; px w1, w1+1 <- px w1-2, w1-3
                        if (i == 0) *reinterpret_cast<uint32_t*>(p + 4) = __byte_perm(v[2 * ch + 1], v[2 * ch], 0x7610);
                        if (i == a.nt[1] - 1) *reinterpret_cast<uint32_t*>(p + 16) = __byte_perm(v[2 * ch + 1], v[2 * ch], 0x7610);
                    }
                }
            }
            if constexpr (L >= 2) {
                __syncthreads();
                on_row<2>(r);
            }
        }
    }
};

template <int L>
__global__ void __launch_bounds__(256, 2) pyrdown_fast_kernel(const FastArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const long long lo = a.total_rows * blockIdx.x / gridDim.x;
    const long long hi = a.total_rows * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;
    FastStream<L> st(a, smem);
    const int hL = a.h[L];
    long long pos = lo;
    while (pos < hi) {
        const int t = (int)(pos / hL);
        const int r0 = (int)(pos - (long long)t * hL);
        const long long frame_end = (long long)(t + 1) * hL;
        const int r1 = (int)((hi < frame_end ? hi : frame_end) - (long long)t * hL);
        __syncthreads();    // previous segment's shared rows no longer read
        st.begin_segment(t, r0, r1);
        st.run_segment();
        pos += r1 - r0;
    }
}

template <int L>
int launch_fast(vhr_ctx* ctx, const FastArgs& a, int threads, int smem_bytes, cudaStream_t stream) {
    auto kern = pyrdown_fast_kernel<L>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    VHR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes));
    if (per_sm < 1) return VHR_ERR_UNSUPPORTED;
    long long grid = (long long)per_sm * ctx->num_sms;
    if (grid > a.total_rows) grid = a.total_rows;
    kern<<<(int)grid, threads, smem_bytes, stream>>>(a);
    return vhr_after_launch(ctx, "pyrdown_fast_kernel");
}

}  // namespace

// Returns VHR_ERR_UNSUPPORTED (without setting an error) when the shape is not eligible.
int vhr_pyrdown_fast(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                     cudaStream_t stream) {
    if (W % 16 != 0 || W > 2048 || (reinterpret_cast<uintptr_t>(d_frames) & 7) != 0) return VHR_ERR_UNSUPPORTED;
    FastArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = d_frames; a.out = d_level; a.T = T; a.H = H; a.W = W; a.levels = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    a.total_rows = (long long)T * a.h[levels];
    a.nt[1] = a.w[1] / 4;
    for (int l = 2; l <= levels; ++l) a.nt[l] = (l == 2) ? a.w[2] / 2 : a.w[l];
    int threads = ((a.nt[1] + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    if (threads > 256) return VHR_ERR_UNSUPPORTED;
    for (int l = 2; l <= levels; ++l)
        if (a.nt[l] > threads) return VHR_ERR_UNSUPPORTED;
    int off = 0;
    auto al16 = [](int v) { return (v + 15) & ~15; };
    if (levels >= 2) {
        a.ring_off[1] = off;
        a.ring_stride[1] = (a.w[1] + 8 + 7) & ~7;                 // uint16 entries per plane row (apron 4 + 4)
        off = al16(off + 2 * 3 * a.ring_stride[1] * 2);
        for (int l = 2; l < levels; ++l) {
            a.ring_off[l] = off;
            a.ring_stride[l] = (a.w[l] + 3) & ~3;
            off = al16(off + 2 * 3 * a.ring_stride[l] * 4);
        }
        for (int l = 2; l <= levels; ++l) {
            a.hring_off[l] = off;
            off = al16(off + HR * 3 * a.w[l] * 4);
        }
    }
    if (off > ctx->smem_optin) return VHR_ERR_UNSUPPORTED;
    switch (levels) {
        case 1: return launch_fast<1>(ctx, a, threads, off, stream);
        case 2: return launch_fast<2>(ctx, a, threads, off, stream);
        case 3: return launch_fast<3>(ctx, a, threads, off, stream);
        case 4: return launch_fast<4>(ctx, a, threads, off, stream);
        case 5: return launch_fast<5>(ctx, a, threads, off, stream);
        case 6: return launch_fast<6>(ctx, a, threads, off, stream);
    }
    return VHR_ERR_INVALID;
}
