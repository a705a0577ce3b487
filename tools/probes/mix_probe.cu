// Ceiling probe for the collapse's traffic mix: read N uint8, write N float32 (1 byte in : 4 bytes out, the collapse
// moves 1 : 3.9), plain vectorised loads / stores, no arithmetic worth mentioning.  Prints achieved GB/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_probe mix_probe.cu && ./mix_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ in16, float4* __restrict__ out, size_t n16) {
    // a lane converts 4 bytes -> one float4; consecutive lanes take consecutive words: 128 B read, 512 B written per warp
    // and instruction; 4 independent words per lane and trip
    const uint32_t* in = reinterpret_cast<const uint32_t*>(in16);
    const size_t n4 = n16 * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (i + k * stride < n4) ? (MODE == 2 ? __ldcs(in + i + k * stride) : __ldg(in + i + k * stride)) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k * stride >= n4) break;
            float4 f;
            f.x = (float)(w[k] & 0xFF); f.y = (float)((w[k] >> 8) & 0xFF); f.z = (float)((w[k] >> 16) & 0xFF); f.w = (float)(w[k] >> 24);
            if (MODE >= 1) __stcs(out + i + k * stride, f); else out[i + k * stride] = f;
        }
    }
}

template <int MODE>
void run(const char* name, const uint4* in, float4* out, size_t n16, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<MODE><<<blocks, 256>>>(in, out, n16);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        mix_kernel<MODE><<<blocks, 256>>>(in, out, n16);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("{\"variant\": \"%s\", \"blocks\": %d, \"ms\": %.4f, \"GBs\": %.1f}\n", name, blocks, best, (double)n16 * 80.0 / best / 1e6);
}

int main() {
    const size_t n = (size_t)1800 * 1080 * 1920 * 3;      // one clip
    const size_t n16 = n / 16;
    uint4* in; float4* out;
    if (cudaMalloc(&in, n) != cudaSuccess || cudaMalloc(&out, n * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(in, 7, n);
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 64}) {
        run<0>("plain", in, out, n16, blocks);
        run<1>("st.cs", in, out, n16, blocks);
        run<2>("ld.cs+st.cs", in, out, n16, blocks);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
