// Host-buffer convenience entry points: the call a NumPy user makes (frames in host memory in, ROI
// traces in host memory out).  H2D of the clip in chunks on a copy stream, the pyrDown cascade chasing
// the copies chunk by chunk on a compute stream, then the temporal bandpass, the collapse + fused ROI
// means and the D2H of the (T,K,3) trace.
//   * When no magnified frames are requested (h_out_f32 == NULL) nothing of the (T,H,W,3) float32
//     output exists on the device either: the collapse runs ROI-only (collapse_sep.cu retires every
//     item outside the ROIs), so the tail behind the last H2D chunk is the bandpass plus the ROI rows.
//   * The device arena (frames + level + optional output + ROI descriptors) and the two streams are
//     context-owned and reused by later calls; vhr_trim() releases them.
//   * h_frames should be page-locked (cudaHostAlloc / torch pin_memory): cudaMemcpyAsync from pageable
//     memory is staged through the driver and does not overlap the pyrDown kernels.
#include "common.cuh"
#include <vector>

namespace {

struct HostRoi {
    const int32_t* rects;     // host (T,K,4) or NULL
    const int32_t* poly;      // host (T,K,Vmax,2) or NULL
    const int32_t* nvert;     // host (T,K)
    int K, Vmax;
    double* mean;             // host (T,K,3)
    int64_t* count;           // host (T,K), polygons only, optional
};

int ensure_streams(vhr_ctx* ctx) {
    if (!ctx->hp_copy) VHR_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->hp_copy, cudaStreamNonBlocking));
    if (!ctx->hp_comp) VHR_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->hp_comp, cudaStreamNonBlocking));
    return VHR_OK;
}

int evm_host_impl(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels, double fps, double f_lo,
                  double f_hi, float alpha, const HostRoi& roi, float* h_out_f32) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, h_frames != nullptr, "null frames");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, levels >= 1 && levels <= VHR_MAX_LEVELS, "levels must be 1..6");
    VHR_REQUIRE(ctx, roi.K >= 0 && roi.K <= VHR_MAX_ROIS, "K must be 0..8");
    VHR_REQUIRE(ctx, roi.K == 0 || ((roi.rects || (roi.poly && roi.nvert)) && roi.mean), "ROI pointers missing");
    VHR_REQUIRE(ctx, roi.K > 0 || h_out_f32, "nothing to compute");
    VHR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_streams(ctx);
    if (rc != VHR_OK) return rc;
    const bool poly = roi.K > 0 && roi.poly != nullptr;
    PyrDims d = vhr_make_dims(W, H, levels);
    const size_t frame_bytes = (size_t)H * W * 3;
    const size_t nframes = frame_bytes * T;
    const size_t P = (size_t)d.w[levels] * d.h[levels] * 3;
    const size_t Kn = roi.K ? roi.K : 1;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t roi_in_bytes = poly ? al((size_t)T * Kn * roi.Vmax * 8) + al((size_t)T * Kn * 4) : al((size_t)T * Kn * 16);
    const size_t need = al(nframes) + al(P * T * 4) + (h_out_f32 ? al(nframes * 4) : 0) + roi_in_bytes + al((size_t)T * Kn * 24) +
                        al((size_t)T * Kn * 8);
    if (need > ctx->hostpath_bytes) {
        if (ctx->hostpath) {
            VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
            VHR_CHECK_CUDA(ctx, cudaFree(ctx->hostpath));
            ctx->hostpath = nullptr;
            ctx->hostpath_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->hostpath, need);
        if (e != cudaSuccess) {
            vhr_set_error(ctx, "vhr_evm_*_host: cudaMalloc(%zu) -> %s", need, cudaGetErrorString(e));
            return VHR_ERR_NOMEM;
        }
        ctx->hostpath_bytes = need;
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(ctx->hostpath);
    uint8_t* b_frames = base;                                       base += al(nframes);
    float* b_level = reinterpret_cast<float*>(base);                base += al(P * T * 4);
    float* b_out = nullptr;
    if (h_out_f32) { b_out = reinterpret_cast<float*>(base);        base += al(nframes * 4); }
    int32_t* b_rects = nullptr; int32_t* b_poly = nullptr; int32_t* b_nvert = nullptr;
    if (poly) {
        b_poly = reinterpret_cast<int32_t*>(base);                  base += al((size_t)T * Kn * roi.Vmax * 8);
        b_nvert = reinterpret_cast<int32_t*>(base);                 base += al((size_t)T * Kn * 4);
    } else {
        b_rects = reinterpret_cast<int32_t*>(base);                 base += al((size_t)T * Kn * 16);
    }
    double* b_means = reinterpret_cast<double*>(base);              base += al((size_t)T * Kn * 24);
    int64_t* b_count = reinterpret_cast<int64_t*>(base);

    cudaStream_t s_copy = ctx->hp_copy, s_comp = ctx->hp_comp;
    std::vector<cudaEvent_t> evs;
    // the previous call of this context (any stream) must be done with the arena and the caches
    rc = vhr_enter(ctx, s_copy);
    if (rc == VHR_OK) rc = vhr_enter(ctx, s_comp);
    // ~256 MiB chunks: large enough for PCIe efficiency, small enough to overlap pyrDown
    int chunk = (int)((size_t)(256u << 20) / frame_bytes);
    if (chunk < 1) chunk = 1;
    auto h2d = [&](void* dst, const void* src, size_t bytes, const char* what) {
        if (rc != VHR_OK) return;
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_copy);
        if (e != cudaSuccess) { vhr_set_error(ctx, "H2D %s -> %s", what, cudaGetErrorString(e)); rc = VHR_ERR_CUDA; }
    };
    if (roi.K > 0) {
        if (poly) {
            h2d(b_poly, roi.poly, sizeof(int32_t) * 2 * (size_t)T * roi.K * roi.Vmax, "polygons");
            h2d(b_nvert, roi.nvert, sizeof(int32_t) * (size_t)T * roi.K, "vertex counts");
        } else {
            h2d(b_rects, roi.rects, sizeof(int32_t) * 4 * (size_t)T * roi.K, "rects");
        }
    }
    for (int t0 = 0; t0 < T && rc == VHR_OK; t0 += chunk) {
        const int tn = (T - t0 < chunk) ? T - t0 : chunk;
        cudaError_t e = cudaMemcpyAsync(b_frames + (size_t)t0 * frame_bytes, h_frames + (size_t)t0 * frame_bytes,
                                        (size_t)tn * frame_bytes, cudaMemcpyHostToDevice, s_copy);
        cudaEvent_t ev = nullptr;
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) { evs.push_back(ev); e = cudaEventRecord(ev, s_copy); }
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_comp, ev, 0);
        if (e != cudaSuccess) { vhr_set_error(ctx, "H2D chunk -> %s", cudaGetErrorString(e)); rc = VHR_ERR_CUDA; break; }
        rc = vhr_pyrdown_cascade(ctx, b_frames + (size_t)t0 * frame_bytes, tn, H, W, levels, b_level + (size_t)t0 * P, s_comp);
    }
    if (rc == VHR_OK) rc = vhr_temporal_bandpass(ctx, b_level, b_level, T, (int64_t)P, fps, f_lo, f_hi, alpha, s_comp);
    if (rc == VHR_OK) {
        if (poly)
            rc = vhr_collapse_addback_poly(ctx, b_level, b_frames, T, H, W, levels, b_out, nullptr, b_poly, b_nvert, roi.K, roi.Vmax,
                                           b_means, roi.count ? b_count : nullptr, s_comp);
        else
            rc = vhr_collapse_addback_roi(ctx, b_level, b_frames, T, H, W, levels, b_out, nullptr, roi.K ? b_rects : nullptr, roi.K,
                                          roi.K ? b_means : nullptr, s_comp);
    }
    auto d2h = [&](void* dst, const void* src, size_t bytes, const char* what) {
        if (rc != VHR_OK) return;
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s_comp);
        if (e != cudaSuccess) { vhr_set_error(ctx, "D2H %s -> %s", what, cudaGetErrorString(e)); rc = VHR_ERR_CUDA; }
    };
    if (roi.K > 0) d2h(roi.mean, b_means, sizeof(double) * 3 * (size_t)T * roi.K, "means");
    if (poly && roi.count) d2h(roi.count, b_count, sizeof(int64_t) * (size_t)T * roi.K, "counts");
    if (h_out_f32) d2h(h_out_f32, b_out, nframes * 4, "frames");
    // every path (errors included) drains both streams before the events go away and the host buffers are released
    cudaError_t e1 = cudaStreamSynchronize(s_copy);
    cudaError_t e2 = cudaStreamSynchronize(s_comp);
    if (rc == VHR_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        vhr_set_error(ctx, "vhr_evm_*_host: sync -> %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = VHR_ERR_CUDA;
    }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    return rc;
}

}  // namespace

extern "C" int vhr_evm_roi_host(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels, double fps,
                                double f_lo, double f_hi, float alpha, const int32_t* h_rects, int K,
                                double* h_roi_mean, float* h_out_f32) {
    HostRoi roi;
    memset(&roi, 0, sizeof(roi));
    roi.rects = h_rects; roi.K = K; roi.mean = h_roi_mean;
    if (ctx && K > 0 && !h_rects) { vhr_set_error(ctx, "vhr_evm_roi_host: ROI pointers missing"); return VHR_ERR_INVALID; }
    return evm_host_impl(ctx, h_frames, T, H, W, levels, fps, f_lo, f_hi, alpha, roi, h_out_f32);
}

extern "C" int vhr_evm_poly_host(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels, double fps,
                                 double f_lo, double f_hi, float alpha, const int32_t* h_poly, const int32_t* h_nvert,
                                 int K, int Vmax, double* h_roi_mean, int64_t* h_count, float* h_out_f32) {
    HostRoi roi;
    memset(&roi, 0, sizeof(roi));
    roi.poly = h_poly; roi.nvert = h_nvert; roi.K = K; roi.Vmax = Vmax; roi.mean = h_roi_mean; roi.count = h_count;
    if (ctx && (K < 1 || !h_poly || !h_nvert || Vmax < 1 || Vmax > VHR_MAX_POLY_VERTS)) {
        vhr_set_error(ctx, "vhr_evm_poly_host: polygon arguments missing or out of range");
        return VHR_ERR_INVALID;
    }
    return evm_host_impl(ctx, h_frames, T, H, W, levels, fps, f_lo, f_hi, alpha, roi, h_out_f32);
}

// Release the context's cached device buffers (host-path arena, scratch arena); tables stay.
extern "C" int vhr_trim(vhr_ctx* ctx) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->hostpath) { VHR_CHECK_CUDA(ctx, cudaFree(ctx->hostpath)); ctx->hostpath = nullptr; ctx->hostpath_bytes = 0; }
    if (ctx->scratch) { VHR_CHECK_CUDA(ctx, cudaFree(ctx->scratch)); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    return VHR_OK;
}
