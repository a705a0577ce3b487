// Fused Gaussian-pyramid pyrDown cascade: uint8 (T,H,W,3) -> float32 (T,hL,wL,3).
//
// No reference code exists for this stage; the arithmetic spec is
// cv2.pyrDown on float32 -- separable [1 4 6 4 1]/16, BORDER_REFLECT_101, (n+1)/2 outputs
// sampled at even coordinates (oracle/evm.py:pyrdown).
//
// Design (HBM-bound byte work; DESIGN.md "pyrdown"):
//   * Persistent grid.  The flattened (frame, final-level row) space is cut into equal
//     contiguous shares, one per CTA, so every CTA streams whole frames top to bottom and
//     pays the 2*(2^L - 1)-row vertical halo at most twice.  Rows are full width: no
//     horizontal halo at all.
//   * Level 1 lives in registers.  A thread owns 8 input pixels (24 bytes) of a row and
//     their 4 level-1 outputs.  The horizontal 5-tap pass is 4 IDP4A per output on the
//     raw bytes (the taps of one output fall into 4 aligned words; the weight vectors are
//     compile-time constants), results are packed two per register (<= 4080 fits 16 bit),
//     and the vertical pass is packed 16-bit integer arithmetic on a 5-row sliding window
//     (<= 65280, exact).  Each input byte is loaded from HBM exactly once.
//   * Level-1 numerators (uint16, exact) and the upper levels (float32) go through small
//     shared-memory rings of 6 rows; a level-l row is produced as soon as its five
//     level-(l-1) rows exist.  Only the final level is written to HBM.
//   * Exactness: levels 1 and 2 are exact integers scaled by 2^-8 / 2^-16, hence bit-exact
//     with cv2 / the float64 oracle; levels >= 3 round in float32 (tests: <= 1e-4 rel).
#include "common.cuh"
#include <stdlib.h>

// tensor-core kernel (pyrdown_umma.cu) and streaming kernel (pyrdown_stream.cu); VHR_ERR_UNSUPPORTED when the shape is not eligible
int vhr_pyrdown_umma(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                     cudaStream_t stream);
int vhr_pyrdown_stream(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels, float* d_level,
                       cudaStream_t stream);

namespace {

constexpr int RING = 6;

struct PyrArgs {
    const uint8_t* frames;
    float* out;
    int T, H, W, levels;
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
    long long total_rows;          // T * h[levels]
    int nt1;                       // threads active in level 1 = ceil(w[1]/4)
    int ring_off[VHR_MAX_LEVELS + 1];   // byte offset of ring l in dynamic smem (l = 1..levels-1)
    int ring_stride[VHR_MAX_LEVELS + 1];// elements per ring row
    int tmp_off;
};

// Weight word for IDP4A: taps of the output whose centre byte (relative to the thread's
// first own byte) is j0, restricted to aligned word `wi` (word 0 = own bytes 0..3).
__host__ __device__ constexpr int floordiv4(int b) { return b >= 0 ? b / 4 : -((-b + 3) / 4); }
__host__ __device__ constexpr uint32_t tap_word(int j0, int wi) {
    uint32_t r = 0;
    const int wt[5] = {1, 4, 6, 4, 1};
    for (int d = 0; d < 5; ++d) {
        int b = j0 + 3 * (d - 2);
        int w = floordiv4(b);
        if (w == wi) r |= (uint32_t)wt[d] << (8 * (b - 4 * w));
    }
    return r;
}

template <int O, int WI>
struct TapAcc {
    // accumulate words WI..6 for output O (12 outputs: m = O/3, c = O%3, j0 = 6m + c)
    __device__ static __forceinline__ uint32_t run(const uint32_t (&wd)[9], uint32_t acc) {
        constexpr uint32_t k = tap_word(6 * (O / 3) + (O % 3), WI);
        if (k != 0) acc = __dp4a(wd[WI + 2], k, acc);
        return TapAcc<O, WI + 1>::run(wd, acc);
    }
};
template <int O>
struct TapAcc<O, 7> {
    __device__ static __forceinline__ uint32_t run(const uint32_t (&)[9], uint32_t acc) { return acc; }
};

template <int O>
struct HPass {
    // 12 outputs -> 6 packed registers (output 2j in the low half, 2j+1 in the high half)
    __device__ static __forceinline__ void run(const uint32_t (&wd)[9], uint32_t (&hp)[6]) {
        uint32_t lo = TapAcc<O, -2>::run(wd, 0u);
        uint32_t hi = TapAcc<O + 1, -2>::run(wd, 0u);
        hp[O / 2] = __byte_perm(lo, hi, 0x5410);
        HPass<O + 2>::run(wd, hp);
    }
};
template <>
struct HPass<12> {
    __device__ static __forceinline__ void run(const uint32_t (&)[9], uint32_t (&)[6]) {}
};

// ---- row loaders: 9 words = bytes [24i-8, 24i+28) of input row `row` -------------------
// ALIGNED: W % 8 == 0 and the frame base 8-byte aligned: every row starts 8-byte aligned.
template <bool ALIGNED>
__device__ __forceinline__ void load_row(const PyrArgs& a, const uint8_t* __restrict__ frame, int row, int i,
                                         uint32_t (&wd)[9]) {
    const uint8_t* rp = frame + (size_t)row * a.W * 3;
    if (ALIGNED) {
        const uint2* p = reinterpret_cast<const uint2*>(rp + 24 * i);
        uint2 o0 = __ldg(p), o1 = __ldg(p + 1), o2 = __ldg(p + 2);
        wd[2] = o0.x; wd[3] = o0.y; wd[4] = o1.x; wd[5] = o1.y; wd[6] = o2.x; wd[7] = o2.y;
        if (i > 0) {
            uint2 l = __ldg(p - 1);
            wd[0] = l.x; wd[1] = l.y;
        } else {
            // pixels -2,-1 reflect to 2,1: bytes -6..-4 <- 6..8, bytes -3..-1 <- 3..5
            wd[0] = __byte_perm(wd[3], 0, 0x3244);                          // (0,0,b6,b7)
            uint32_t t = __byte_perm(wd[2], wd[3], 0x5430);                 // (b0,b3,b4,b5)
            wd[1] = __byte_perm(t, wd[4], 0x3214);                          // (b8,b3,b4,b5)
        }
        if (i < a.nt1 - 1) {
            wd[8] = __ldg(reinterpret_cast<const uint32_t*>(rp + 24 * i + 24));
        } else {
            // pixel W reflects to W-2 = own pixel 6: bytes 24..26 <- 18..20
            wd[8] = __byte_perm(wd[6], wd[7], 0x0432);                      // (b18,b19,b20,x)
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) wd[k] = 0;
#pragma unroll
        for (int q = -2; q <= 8; ++q) {
            int px = vhr_reflect101(8 * i + q, a.W);
            // threads past the right edge compute padding that is never read
            px = min(max(px, 0), a.W - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int rel = 3 * q + c + 8;                  // 0..35 (byte -8 -> 0)
                uint32_t v = __ldg(rp + 3 * px + c);
                wd[rel >> 2] |= v << (8 * (rel & 3));
            }
        }
    }
}

template <int L, bool ALIGNED>
struct Stream {
    const PyrArgs& a;
    unsigned char* smem;
    const uint8_t* frame;       // current frame base
    float* out_frame;           // current frame's final-level output base
    int next[VHR_MAX_LEVELS + 1];   // next row to produce per level
    // level-1 sliding window (packed H rows for input rows 2r-2..2r+2)
    uint32_t win[5][6];
    int l1_ready_for;           // window currently positioned so that rows 2r-2..2r+2 are for r == l1_ready_for, or -1
    uint32_t pre[2][9];         // prefetched words of the two rows that enter the window next
    int pre_for;                // r for which `pre` holds rows 2r+1, 2r+2, or -1

    __device__ Stream(const PyrArgs& a_, unsigned char* s) : a(a_), smem(s) {}

    __device__ __forceinline__ uint16_t* ring1(int row) const {
        return reinterpret_cast<uint16_t*>(smem + a.ring_off[1]) + (size_t)(row % RING) * a.ring_stride[1];
    }
    __device__ __forceinline__ float* ringf(int l, int row) const {
        return reinterpret_cast<float*>(smem + a.ring_off[l]) + (size_t)(row % RING) * a.ring_stride[l];
    }

    __device__ __forceinline__ void begin_segment(int t, int r0) {
        frame = a.frames + (size_t)t * a.H * a.W * 3;
        out_frame = a.out + (size_t)t * a.h[L] * a.w[L] * 3;
        // first row each level must produce for final row r0
        int f = r0;
        next[L] = r0;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            f = max(0, 2 * f - 2);
            next[l] = f;
        }
        l1_ready_for = -100;      // sentinels must not collide with r - 1 for r = 0
        pre_for = -100;
    }

    // ---- level 1 -------------------------------------------------------------------------
    __device__ __forceinline__ void hrow(int in_row, uint32_t (&hp)[6]) {
        uint32_t wd[9];
        load_row<ALIGNED>(a, frame, vhr_reflect101(in_row, a.H), threadIdx.x, wd);
        HPass<0>::run(wd, hp);
    }

    __device__ __forceinline__ void produce_l1(int r) {
        const int i = threadIdx.x;
        if (i < a.nt1) {
            if (l1_ready_for != r) {
                if (l1_ready_for == r - 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { win[0][k] = win[2][k]; win[1][k] = win[3][k]; win[2][k] = win[4][k]; }
                    if (pre_for == r) {
                        HPass<0>::run(pre[0], win[3]);
                        HPass<0>::run(pre[1], win[4]);
                    } else {
                        hrow(2 * r + 1, win[3]);
                        hrow(2 * r + 2, win[4]);
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < 5; ++d) hrow(2 * r - 2 + d, win[d]);
                }
                l1_ready_for = r;
            }
            // prefetch the two rows that enter the window for r+1 (loads stay in flight
            // across the shared-memory stages of the upper levels)
            if (r + 1 < a.h[1]) {
                load_row<ALIGNED>(a, frame, vhr_reflect101(2 * r + 3, a.H), i, pre[0]);
                load_row<ALIGNED>(a, frame, vhr_reflect101(2 * r + 4, a.H), i, pre[1]);
                pre_for = r + 1;
            } else {
                pre_for = -100;
            }
            // vertical pass, packed 16-bit lanes (max 65280: no carry between halves)
            uint32_t v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = (win[0][k] + win[4][k]) + ((win[1][k] + win[3][k]) << 2) + win[2][k] * 6u;
            if (L == 1) {
                float* o = out_frame + ((size_t)r * a.w[1]) * 3 + 12 * i;
                const int lim = a.w[1] * 3 - 12 * i;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if (2 * k < lim) o[2 * k] = (float)(v[k] & 0xFFFFu) * (1.0f / 256.0f);
                    if (2 * k + 1 < lim) o[2 * k + 1] = (float)(v[k] >> 16) * (1.0f / 256.0f);
                }
            } else {
                uint2* o = reinterpret_cast<uint2*>(ring1(r) + 12 * i);
                o[0] = make_uint2(v[0], v[1]);
                o[1] = make_uint2(v[2], v[3]);
                o[2] = make_uint2(v[4], v[5]);
            }
        }
    }

    // ---- levels >= 2 from the ring of level l-1 -------------------------------------------
    template <int l>
    __device__ __forceinline__ void produce_upper(int r) {
        const int wp = a.w[l - 1], hp_ = a.h[l - 1], wl = a.w[l];
        const int nprev = wp * 3;
        int rr[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) rr[d] = vhr_reflect101(2 * r - 2 + d, hp_);
        float* tmp = reinterpret_cast<float*>(smem + a.tmp_off);
        __syncthreads();   // ring l-1 rows complete; tmp free
        if constexpr (l == 2) {
            const uint16_t* p0 = ring1(rr[0]); const uint16_t* p1 = ring1(rr[1]); const uint16_t* p2 = ring1(rr[2]);
            const uint16_t* p3 = ring1(rr[3]); const uint16_t* p4 = ring1(rr[4]);
            for (int j = threadIdx.x; j < nprev; j += blockDim.x) {
                int s = ((int)p0[j] + (int)p4[j]) + 4 * ((int)p1[j] + (int)p3[j]) + 6 * (int)p2[j];
                tmp[j] = (float)s;      // < 2^20, exact
            }
        } else {
            const float* p0 = ringf(l - 1, rr[0]); const float* p1 = ringf(l - 1, rr[1]); const float* p2 = ringf(l - 1, rr[2]);
            const float* p3 = ringf(l - 1, rr[3]); const float* p4 = ringf(l - 1, rr[4]);
            for (int j = threadIdx.x; j < nprev; j += blockDim.x)
                tmp[j] = p2[j] * 6.0f + (p1[j] + p3[j]) * 4.0f + p0[j] + p4[j];
        }
        __syncthreads();
        const float scale = (l == 2) ? (1.0f / 65536.0f) : (1.0f / 256.0f);
        float* dst = (l == L) ? (out_frame + (size_t)r * wl * 3) : ringf(l, r);
        for (int o = threadIdx.x; o < wl * 3; o += blockDim.x) {
            int x = o / 3, c = o - 3 * x;
            int x0 = vhr_reflect101(2 * x - 2, wp), x1 = vhr_reflect101(2 * x - 1, wp), x2 = 2 * x;
            int x3 = vhr_reflect101(2 * x + 1, wp), x4 = vhr_reflect101(2 * x + 2, wp);
            float s = tmp[3 * x2 + c] * 6.0f + (tmp[3 * x1 + c] + tmp[3 * x3 + c]) * 4.0f + tmp[3 * x0 + c] + tmp[3 * x4 + c];
            dst[o] = s * scale;
        }
    }

    // ---- demand-driven schedule -----------------------------------------------------------
    template <int l>
    __device__ __forceinline__ void ensure(int upto) {
        if (upto > a.h[l] - 1) upto = a.h[l] - 1;
        while (next[l] <= upto) {
            const int r = next[l];
            if constexpr (l == 1) {
                produce_l1(r);
            } else {
                ensure<l - 1>(2 * r + 2);
                produce_upper<l>(r);
            }
            next[l] = r + 1;
        }
    }
};

template <int L, bool ALIGNED, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 ? 2 : 1)) pyrdown_kernel(const PyrArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const long long lo = a.total_rows * blockIdx.x / gridDim.x;
    const long long hi = a.total_rows * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;
    Stream<L, ALIGNED> st(a, smem);
    const int hL = a.h[L];
    long long pos = lo;
    while (pos < hi) {
        const int t = (int)(pos / hL);
        const int r0 = (int)(pos - (long long)t * hL);
        const long long frame_end = (long long)(t + 1) * hL;
        const int r1 = (int)((hi < frame_end ? hi : frame_end) - (long long)t * hL);
        __syncthreads();    // previous segment's smem reads done
        st.begin_segment(t, r0);
        st.template ensure<L>(r1 - 1);
        pos += r1 - r0;
    }
}

template <int L, bool ALIGNED, int MAXT>
int launch_t(vhr_ctx* ctx, const PyrArgs& a, int threads, int smem_bytes, cudaStream_t stream) {
    auto kern = pyrdown_kernel<L, ALIGNED, MAXT>;
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    VHR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes));
    if (per_sm < 1) {
        vhr_set_error(ctx, "pyrdown: kernel does not fit (threads %d, smem %d)", threads, smem_bytes);
        return VHR_ERR_UNSUPPORTED;
    }
    long long grid = (long long)per_sm * ctx->num_sms;
    if (grid > a.total_rows) grid = a.total_rows;
    kern<<<(int)grid, threads, smem_bytes, stream>>>(a);
    return vhr_after_launch(ctx, "pyrdown_kernel");
}

template <int L, bool ALIGNED>
int launch(vhr_ctx* ctx, const PyrArgs& a, int threads, int smem_bytes, cudaStream_t stream) {
    return threads <= 256 ? launch_t<L, ALIGNED, 256>(ctx, a, threads, smem_bytes, stream)
                          : launch_t<L, ALIGNED, 1024>(ctx, a, threads, smem_bytes, stream);
}

template <bool ALIGNED>
int dispatch(vhr_ctx* ctx, const PyrArgs& a, int threads, int smem_bytes, cudaStream_t s) {
    switch (a.levels) {
        case 1: return launch<1, ALIGNED>(ctx, a, threads, smem_bytes, s);
        case 2: return launch<2, ALIGNED>(ctx, a, threads, smem_bytes, s);
        case 3: return launch<3, ALIGNED>(ctx, a, threads, smem_bytes, s);
        case 4: return launch<4, ALIGNED>(ctx, a, threads, smem_bytes, s);
        case 5: return launch<5, ALIGNED>(ctx, a, threads, smem_bytes, s);
        case 6: return launch<6, ALIGNED>(ctx, a, threads, smem_bytes, s);
    }
    return VHR_ERR_INVALID;
}

}  // namespace

extern "C" int vhr_pyrdown_cascade(vhr_ctx* ctx, const uint8_t* d_frames, int T, int H, int W, int levels,
                                   float* d_level, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_frames && d_level, "null pointer");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, levels >= 1 && levels <= VHR_MAX_LEVELS, "levels must be 1..6");
    VHR_REQUIRE(ctx, W <= 8192, "W > 8192 unsupported");
    {
        // Dispatch: the tcgen05 kernel (pyrdown_umma.cu: 4 levels, W % 80 == 0), then the streaming kernel
        // (pyrdown_stream.cu: W % 16 == 0, W % 64 == 0 for >= 3 levels), then the generic kernel below.  Test /
        // measurement hooks: VHR_PYRDOWN_IMPL=stream skips the tcgen05 kernel, VHR_PYRDOWN_GENERIC=1 forces the generic one.
        const char* force = getenv("VHR_PYRDOWN_GENERIC");
        const char* impl = getenv("VHR_PYRDOWN_IMPL");
        if (!(force && force[0] == '1')) {
            int rc = VHR_ERR_UNSUPPORTED;
            if (!(impl && impl[0] == 's')) rc = vhr_pyrdown_umma(ctx, d_frames, T, H, W, levels, d_level, (cudaStream_t)stream);
            if (rc == VHR_ERR_UNSUPPORTED) rc = vhr_pyrdown_stream(ctx, d_frames, T, H, W, levels, d_level, (cudaStream_t)stream);
            if (rc != VHR_ERR_UNSUPPORTED) return rc;
        }
    }
    PyrArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = d_frames;
    a.out = d_level;
    a.T = T; a.H = H; a.W = W; a.levels = levels;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= VHR_MAX_LEVELS; ++l) { a.w[l] = d.w[l]; a.h[l] = d.h[l]; }
    a.total_rows = (long long)T * a.h[levels];
    a.nt1 = (a.w[1] + 3) / 4;
    int threads = ((a.nt1 + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    int off = 0;
    if (levels >= 2) {
        a.ring_off[1] = off;
        a.ring_stride[1] = a.nt1 * 12;
        off += RING * a.ring_stride[1] * 2;
        off = (off + 15) & ~15;
        for (int l = 2; l < levels; ++l) {
            a.ring_off[l] = off;
            a.ring_stride[l] = (a.w[l] * 3 + 3) & ~3;
            off += RING * a.ring_stride[l] * 4;
        }
        a.tmp_off = off;
        off += ((a.w[1] * 3 + 3) & ~3) * 4;
    }
    const bool aligned = (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 7) == 0);
    if (off > ctx->smem_optin) {
        vhr_set_error(ctx, "pyrdown: W=%d needs %d bytes of shared memory (> %d)", W, off, ctx->smem_optin);
        return VHR_ERR_UNSUPPORTED;
    }
    return aligned ? dispatch<true>(ctx, a, threads, off, (cudaStream_t)stream)
                   : dispatch<false>(ctx, a, threads, off, (cudaStream_t)stream);
}
