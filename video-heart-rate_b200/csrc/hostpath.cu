// Host-buffer convenience entry point: the call a NumPy user makes (frames in host memory
// in, ROI traces in host memory out).  H2D of the clip in chunks on a copy stream, the
// pyrDown cascade chasing the copies chunk by chunk on a compute stream, then the temporal
// bandpass, the collapse + fused ROI means and the D2H of the (T,K,3) trace.
#include "common.cuh"
#include <vector>

namespace {
struct HostPathBufs {
    uint8_t* frames;
    float* level;
    float* out;
    int32_t* rects;
    double* means;
};
}  // namespace

extern "C" int vhr_evm_roi_host(vhr_ctx* ctx, const uint8_t* h_frames, int T, int H, int W, int levels, double fps,
                                double f_lo, double f_hi, float alpha, const int32_t* h_rects, int K,
                                double* h_roi_mean, float* h_out_f32) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, h_frames != nullptr, "null frames");
    VHR_REQUIRE(ctx, T >= 1 && H >= 1 && W >= 1, "bad shape");
    VHR_REQUIRE(ctx, levels >= 1 && levels <= VHR_MAX_LEVELS, "levels must be 1..6");
    VHR_REQUIRE(ctx, K >= 0 && K <= 4, "K must be 0..4");
    VHR_REQUIRE(ctx, K == 0 || (h_rects && h_roi_mean), "ROI pointers missing");
    VHR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    PyrDims d = vhr_make_dims(W, H, levels);
    const size_t frame_bytes = (size_t)H * W * 3;
    const size_t nframes = frame_bytes * T;
    const size_t P = (size_t)d.w[levels] * d.h[levels] * 3;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t need = al(nframes) + al(P * T * 4) + al(nframes * 4) + al((size_t)T * (K ? K : 1) * 16) + al((size_t)T * (K ? K : 1) * 24);
    if (need > ctx->hostpath_bytes) {
        if (ctx->hostpath) {
            VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
            VHR_CHECK_CUDA(ctx, cudaFree(ctx->hostpath));
            ctx->hostpath = nullptr;
            ctx->hostpath_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->hostpath, need);
        if (e != cudaSuccess) {
            vhr_set_error(ctx, "vhr_evm_roi_host: cudaMalloc(%zu) -> %s", need, cudaGetErrorString(e));
            return VHR_ERR_NOMEM;
        }
        ctx->hostpath_bytes = need;
    }
    HostPathBufs b;
    unsigned char* base = reinterpret_cast<unsigned char*>(ctx->hostpath);
    b.frames = base;                                   base += al(nframes);
    b.level = reinterpret_cast<float*>(base);          base += al(P * T * 4);
    b.out = reinterpret_cast<float*>(base);            base += al(nframes * 4);
    b.rects = reinterpret_cast<int32_t*>(base);        base += al((size_t)T * (K ? K : 1) * 16);
    b.means = reinterpret_cast<double*>(base);

    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    VHR_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    VHR_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    int rc = VHR_OK;
    std::vector<cudaEvent_t> evs;
    // ~256 MiB chunks: large enough for PCIe efficiency, small enough to overlap pyrDown
    int chunk = (int)((size_t)(256u << 20) / frame_bytes);
    if (chunk < 1) chunk = 1;
    if (K > 0) {
        cudaError_t e = cudaMemcpyAsync(b.rects, h_rects, sizeof(int32_t) * 4 * (size_t)T * K, cudaMemcpyHostToDevice, s_copy);
        if (e != cudaSuccess) { vhr_set_error(ctx, "H2D rects -> %s", cudaGetErrorString(e)); rc = VHR_ERR_CUDA; }
    }
    for (int t0 = 0; t0 < T && rc == VHR_OK; t0 += chunk) {
        const int tn = (T - t0 < chunk) ? T - t0 : chunk;
        cudaError_t e = cudaMemcpyAsync(b.frames + (size_t)t0 * frame_bytes, h_frames + (size_t)t0 * frame_bytes,
                                        (size_t)tn * frame_bytes, cudaMemcpyHostToDevice, s_copy);
        cudaEvent_t ev = nullptr;
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) { evs.push_back(ev); e = cudaEventRecord(ev, s_copy); }
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_comp, ev, 0);
        if (e != cudaSuccess) { vhr_set_error(ctx, "H2D chunk -> %s", cudaGetErrorString(e)); rc = VHR_ERR_CUDA; break; }
        rc = vhr_pyrdown_cascade(ctx, b.frames + (size_t)t0 * frame_bytes, tn, H, W, levels, b.level + (size_t)t0 * P, s_comp);
    }
    if (rc == VHR_OK) rc = vhr_temporal_bandpass(ctx, b.level, b.level, T, (int64_t)P, fps, f_lo, f_hi, alpha, s_comp);
    if (rc == VHR_OK)
        rc = vhr_collapse_addback_roi(ctx, b.level, b.frames, T, H, W, levels, b.out, nullptr, K ? b.rects : nullptr, K,
                                      K ? b.means : nullptr, s_comp);
    if (rc == VHR_OK && K > 0) {
        cudaError_t e = cudaMemcpyAsync(h_roi_mean, b.means, sizeof(double) * 3 * (size_t)T * K, cudaMemcpyDeviceToHost, s_comp);
        if (e != cudaSuccess) { vhr_set_error(ctx, "D2H means -> %s", cudaGetErrorString(e)); rc = VHR_ERR_CUDA; }
    }
    if (rc == VHR_OK && h_out_f32) {
        cudaError_t e = cudaMemcpyAsync(h_out_f32, b.out, nframes * 4, cudaMemcpyDeviceToHost, s_comp);
        if (e != cudaSuccess) { vhr_set_error(ctx, "D2H frames -> %s", cudaGetErrorString(e)); rc = VHR_ERR_CUDA; }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_copy);
    cudaError_t e2 = cudaStreamSynchronize(s_comp);
    if (rc == VHR_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        vhr_set_error(ctx, "vhr_evm_roi_host: sync -> %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = VHR_ERR_CUDA;
    }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    cudaStreamDestroy(s_copy);
    cudaStreamDestroy(s_comp);
    return rc;
}
