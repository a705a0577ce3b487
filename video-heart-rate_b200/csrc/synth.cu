// Synthetic clip generator, bit-identical to oracle/synth.py (pure integer function of
// (seed, clip, t, y, x, c)).  One thread writes 16 consecutive bytes.
#include "common.cuh"

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

struct SynthArgs {
    vhr_synth_params p;
};

__global__ void __launch_bounds__(256) synth_kernel(SynthArgs a, const int32_t* __restrict__ pulse,
                                                    uint8_t* __restrict__ out) {
    const vhr_synth_params& p = a.p;
    const int64_t frame_bytes = (int64_t)p.H * p.W * 3;
    const int64_t total = frame_bytes * p.T;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < total; base += stride) {
        uint32_t words[4] = {0, 0, 0, 0};
        int tl = (int)(base / frame_bytes);
        uint32_t idx = (uint32_t)(base - (int64_t)tl * frame_bytes);
        int cur_t = -1;
        uint32_t key = 0;
        int32_t pq[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (base + j >= total) break;
            if (idx >= (uint32_t)frame_bytes) {
                idx -= (uint32_t)frame_bytes;
                ++tl;
            }
            if (tl != cur_t) {
                cur_t = tl;
                uint32_t t = (uint32_t)(p.t0 + tl);
                key = mix32(p.seed * 0x9E3779B1u + p.clip * 0x85EBCA77u + t * 0xC2B2AE3Du + 0x165667B1u);
                pq[0] = pulse[(p.t0 + tl) * 3 + 0];
                pq[1] = pulse[(p.t0 + tl) * 3 + 1];
                pq[2] = pulse[(p.t0 + tl) * 3 + 2];
            }
            uint32_t pix = idx / 3u;
            int c = (int)(idx - pix * 3u);
            int y = (int)(pix / (uint32_t)p.W);
            int x = (int)(pix - (uint32_t)y * (uint32_t)p.W);
            int face = (x >= p.face[0]) & (x < p.face[2]) & (y >= p.face[1]) & (y < p.face[3]);
            uint32_t r = mix32(key ^ (idx * 0x27D4EB2Fu));
            int s = (int)((r & 255u) + ((r >> 8) & 255u) + ((r >> 16) & 255u) + (r >> 24)) - 510;
            int v = p.base_q8[face][c] + face * pq[c] + s * p.noise_gain;
            v = (v + 128) >> 8;
            v = min(max(v, 0), 255);
            words[j >> 2] |= (uint32_t)v << ((j & 3) * 8);
            ++idx;
        }
        if (base + 16 <= total && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            *reinterpret_cast<uint4*>(out + base) = make_uint4(words[0], words[1], words[2], words[3]);
        } else {
            for (int j = 0; j < 16 && base + j < total; ++j) out[base + j] = (uint8_t)(words[j >> 2] >> ((j & 3) * 8));
        }
    }
}

extern "C" int vhr_synth_clip(vhr_ctx* ctx, const vhr_synth_params* p, const int32_t* d_pulse_q8,
                              uint8_t* d_frames, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, p && d_pulse_q8 && d_frames, "null pointer");
    VHR_REQUIRE(ctx, p->T >= 1 && p->H >= 1 && p->W >= 1, "bad shape");
    VHR_REQUIRE(ctx, (int64_t)p->H * p->W * 3 < (int64_t)0xFFFFFFFFll, "frame too large");
    SynthArgs a;
    a.p = *p;
    int64_t total = (int64_t)p->T * p->H * p->W * 3;
    int64_t units = (total + 15) / 16;
    int blocks = (int)((units + 255) / 256);
    int maxb = ctx->num_sms * 16;
    if (blocks > maxb) blocks = maxb;
    synth_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, d_pulse_q8, d_frames);
    return vhr_after_launch(ctx, "synth_kernel");
}
