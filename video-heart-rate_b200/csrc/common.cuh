// Shared declarations for libvhr_b200.so (sm_100a).  See include/vhr_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/vhr_b200.h"

struct vhr_ctx {
    int device = 0;
    int num_sms = 148;
    int smem_optin = 0;
    char err[512] = {0};
    std::atomic<int64_t> launches{0};
    // scratch arena (device), grown on demand outside the hot loop
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // cached twiddle table for the temporal bandpass
    float2* tw = nullptr;
    int tw_T = 0;
    // cached digit-reversed band mask of the FFT bandpass
    float* mask = nullptr;
    int mask_T = 0, mask_k0 = -1, mask_k1 = -1;
    float mask_gain = 0.f;
    unsigned long long mask_radix_key = 0;    // pass order of the FFT the mask's digit-reversed layout belongs to
    // stream hand-over (vhr_enter / vhr_leave): the cached tables and the scratch arena are shared by every
    // call of a context, so a call on another stream than the previous one first waits for that one
    cudaStream_t last_stream = nullptr;
    cudaEvent_t last_ev = nullptr;
    bool have_last = false;
    // context-owned buffers of the *_host convenience path
    void* hostpath = nullptr;
    size_t hostpath_bytes = 0;
    cudaStream_t hp_copy = nullptr, hp_comp = nullptr;     // created on first use, destroyed with the context
    // composite pyrUp weight tables of the collapse (collapse_sep.cu), keyed by shape
    void* sep_tab = nullptr;
    size_t sep_tab_bytes = 0;
    long long sep_key = -1;
    // weight bands + border slices of the tensor-core pyrDown (pyrdown_umma.cu), one read-only blob per frame shape.
    // A blob is never rewritten or freed while the context lives (kernels on any stream may be reading it); when the
    // table is full the whole device is synchronised before the oldest entry is replaced.
    static constexpr int UMMA_SLOTS = 8;
    void* umma_blob[UMMA_SLOTS] = {nullptr};
    long long umma_key[UMMA_SLOTS] = {0};
    int umma_n = 0, umma_next = 0;
};

void vhr_set_error(vhr_ctx* ctx, const char* fmt, ...);
// Contract: a context is used by ONE stream at a time.  Entry points that read or rewrite context-owned
// device state (twiddles, band mask, composite pyrUp tables, scratch arena) bracket their work with
// vhr_enter / vhr_leave: when the stream differs from the previous call's, the new stream first waits
// for the previous call's work, so switching streams between calls is safe; two streams driving one
// context concurrently is not supported.
int vhr_enter(vhr_ctx* ctx, cudaStream_t stream);
int vhr_leave(vhr_ctx* ctx, cudaStream_t stream, int rc);
int vhr_scratch(vhr_ctx* ctx, size_t bytes, void** out);
// roi.cu: rasterise polygons (T,K,Vmax,2) into frame-aligned row bit-masks (T,K,H,MW) (only the rows and
// words of each polygon's clamped bounding box are written), their bounding boxes (T,K,4) [x1,y1,x2,y2)
// (all zero when empty) and pixel counts (T,K).  Used by the fused polygon ROI of the collapse.
int vhr_poly_rowmask(vhr_ctx* ctx, int T, int H, int W, const int32_t* d_poly, const int32_t* d_nvert, int K, int Vmax,
                     uint32_t* d_mask, int MW, int32_t* d_box, long long* d_count, cudaStream_t stream);

#define VHR_CHECK_CUDA(ctx, expr)                                                        \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            vhr_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                          cudaGetErrorString(_e));                                       \
            return VHR_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define VHR_REQUIRE(ctx, cond, msg)                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            vhr_set_error(ctx, "%s: %s", __func__, msg);                                 \
            return VHR_ERR_INVALID;                                                      \
        }                                                                                \
    } while (0)

#define VHR_LAUNCHED(ctx, n) (ctx)->launches.fetch_add((n), std::memory_order_relaxed)

static inline int vhr_after_launch(vhr_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        vhr_set_error(ctx, "launch %s -> %s", what, cudaGetErrorString(e));
        return VHR_ERR_CUDA;
    }
    VHR_LAUNCHED(ctx, 1);
    return VHR_OK;
}

__host__ __device__ static inline int vhr_reflect101(int i, int n) {
    // BORDER_REFLECT_101 for |excursion| < n (all uses here: excursion <= 2)
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    if (i < 0) i = -i;            // n == 2, i == 3 -> -1 -> 1
    return i;
}

struct PyrDims {
    int w[VHR_MAX_LEVELS + 1];
    int h[VHR_MAX_LEVELS + 1];
};
static inline PyrDims vhr_make_dims(int W, int H, int levels) {
    PyrDims d;
    d.w[0] = W;
    d.h[0] = H;
    for (int l = 1; l <= VHR_MAX_LEVELS; ++l) {
        d.w[l] = l <= levels ? (d.w[l - 1] + 1) / 2 : 0;
        d.h[l] = l <= levels ? (d.h[l - 1] + 1) / 2 : 0;
    }
    return d;
}
