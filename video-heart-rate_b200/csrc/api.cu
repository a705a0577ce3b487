// Context management and small host helpers of the C ABI (include/vhr_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <math.h>
#include <new>

static char g_last_error[512] = {0};

void vhr_set_error(vhr_ctx* ctx, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    if (ctx) memcpy(ctx->err, g_last_error, sizeof(g_last_error));
}

int vhr_scratch(vhr_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            VHR_CHECK_CUDA(ctx, cudaDeviceSynchronize());
            VHR_CHECK_CUDA(ctx, cudaFree(ctx->scratch));
            ctx->scratch = nullptr;
            ctx->scratch_bytes = 0;
        }
        size_t want = bytes + (bytes >> 2) + 4096;
        cudaError_t e = cudaMalloc(&ctx->scratch, want);
        if (e != cudaSuccess) {
            vhr_set_error(ctx, "scratch cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
            return VHR_ERR_NOMEM;
        }
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return VHR_OK;
}

int vhr_enter(vhr_ctx* ctx, cudaStream_t stream) {
    if (ctx->have_last && ctx->last_stream != stream)
        VHR_CHECK_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->last_ev, 0));
    return VHR_OK;
}

int vhr_leave(vhr_ctx* ctx, cudaStream_t stream, int rc) {
    if (!ctx->last_ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->last_ev, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            vhr_set_error(ctx, "cudaEventCreate -> %s", cudaGetErrorString(e));
            return rc != VHR_OK ? rc : VHR_ERR_CUDA;
        }
    }
    cudaError_t e = cudaEventRecord(ctx->last_ev, stream);
    if (e != cudaSuccess) {
        vhr_set_error(ctx, "cudaEventRecord -> %s", cudaGetErrorString(e));
        return rc != VHR_OK ? rc : VHR_ERR_CUDA;
    }
    ctx->last_stream = stream;
    ctx->have_last = true;
    return rc;
}

extern "C" {

int vhr_abi_version(void) { return VHR_ABI_VERSION; }

int vhr_create(vhr_ctx** out, int device) {
    if (!out) return VHR_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        vhr_set_error(nullptr, "vhr_create: no CUDA device (%s)", cudaGetErrorString(e));
        return VHR_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        vhr_set_error(nullptr, "vhr_create: device %d out of range (0..%d)", device, ndev - 1);
        return VHR_ERR_INVALID;
    }
    vhr_ctx* ctx = new (std::nothrow) vhr_ctx();
    if (!ctx) return VHR_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        vhr_set_error(nullptr, "vhr_create: device %d -> %s", device, cudaGetErrorString(e));
        delete ctx;
        return VHR_ERR_CUDA;
    }
    if (prop.major < 10) {
        vhr_set_error(nullptr, "vhr_create: device %d is sm_%d%d; this library is built for sm_100a only",
                      device, prop.major, prop.minor);
        delete ctx;
        return VHR_ERR_UNSUPPORTED;
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = ctx;
    return VHR_OK;
}

int vhr_destroy(vhr_ctx* ctx) {
    if (!ctx) return VHR_OK;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->tw) cudaFree(ctx->tw);
    if (ctx->mask) cudaFree(ctx->mask);
    if (ctx->hostpath) cudaFree(ctx->hostpath);
    if (ctx->sep_tab) cudaFree(ctx->sep_tab);
    for (int i = 0; i < ctx->umma_n; ++i) if (ctx->umma_blob[i]) cudaFree(ctx->umma_blob[i]);
    if (ctx->last_ev) cudaEventDestroy(ctx->last_ev);
    if (ctx->hp_copy) cudaStreamDestroy(ctx->hp_copy);
    if (ctx->hp_comp) cudaStreamDestroy(ctx->hp_comp);
    delete ctx;
    return VHR_OK;
}

const char* vhr_last_error(const vhr_ctx* ctx) { return ctx ? ctx->err : g_last_error; }

int64_t vhr_launch_count(const vhr_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int vhr_pyr_dims(int W, int H, int levels, int32_t* w_out, int32_t* h_out) {
    if (W < 1 || H < 1 || levels < 0 || levels > VHR_MAX_LEVELS || !w_out || !h_out) return VHR_ERR_INVALID;
    PyrDims d = vhr_make_dims(W, H, levels);
    for (int l = 0; l <= levels; ++l) {
        w_out[l] = d.w[l];
        h_out[l] = d.h[l];
    }
    return VHR_OK;
}

// rfftfreq exactly as NumPy computes it: val = 1.0/(n*d), f_k = k*val, d = 1/fps
// (numpy/fft/_helper.py rfftfreq).  Inclusive edges, DC dropped.
int vhr_band_bins(int T, double fps, double f_lo, double f_hi, int* k_first, int* k_last) {
    if (T < 1 || !(fps > 0)) return VHR_ERR_INVALID;
    volatile double d = 1.0 / fps;
    volatile double nd = (double)T * d;
    volatile double val = 1.0 / nd;
    int first = -1, last = -1, count = 0;
    for (int k = 1; k <= T / 2; ++k) {
        volatile double f = (double)k * val;
        if (f >= f_lo && f <= f_hi) {
            if (first < 0) first = k;
            last = k;
            ++count;
        }
    }
    if (k_first) *k_first = first;
    if (k_last) *k_last = last;
    return count;
}

}  // extern "C"
