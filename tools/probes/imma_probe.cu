// Micro-probe: issue rate of legacy integer tensor-core MMA (mma.sync.m16n8k32 u8 x s8 -> s32) on sm_100a,
// next to the integer dot product (IDP4A) the pyrDown kernel uses today.  Answers one question for
// can the level-1 horizontal 5-tap of the pyrDown cascade run as banded IMMA tiles without the
// tensor pipe becoming the limiter?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_u8s8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void imma_loop(int iters, int* out, long long* cycles) {
    uint32_t a[4] = {0x01010101u * (threadIdx.x & 3), 0x02020202u, 0x01020304u, 0x04030201u};
    uint32_t b[2] = {threadIdx.x * 0x01010101u, 0x03030303u};
    int d[ILP][4];
#pragma unroll
    for (int j = 0; j < ILP; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) mma_u8s8(d[j], a, b);
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void dp4a_loop(int iters, int* out, long long* cycles) {
    uint32_t a = threadIdx.x * 0x01010101u, b = 0x01040601u;
    int d[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) d[j] = j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) d[j] = __dp4a((int)(a + j), (int)b, d[j]);
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    int* out; long long* cyc;
    cudaMalloc(&out, sizeof(int) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    const int iters = 20000;
    printf("{\"probe\": \"imma_m16n8k32_u8s8\", \"iters\": %d, \"rows\": [\n", iters);
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int ilp : {1, 4}) {
            long long h[148];
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            if (ilp == 1) imma_loop<1><<<148, warps * 32>>>(iters, out, cyc); else imma_loop<4><<<148, warps * 32>>>(iters, out, cyc);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            if (ilp == 1) imma_loop<1><<<148, warps * 32>>>(iters, out, cyc); else imma_loop<4><<<148, warps * 32>>>(iters, out, cyc);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double mmas_per_sm = (double)iters * ilp * warps;
            printf("  {\"op\": \"imma\", \"warps_per_sm\": %d, \"ilp\": %d, \"cycles_per_mma_per_sm\": %.3f, \"ms\": %.3f},\n", warps, ilp,
                   (double)h[0] / mmas_per_sm, ms);
            if (ilp == 1) dp4a_loop<1><<<148, warps * 32>>>(iters, out, cyc); else dp4a_loop<4><<<148, warps * 32>>>(iters, out, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("  {\"op\": \"dp4a\", \"warps_per_sm\": %d, \"ilp\": %d, \"cycles_per_warp_instr_per_sm\": %.3f},\n", warps, ilp,
                   (double)h[0] / mmas_per_sm);
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("  {\"cuda\": \"%s\"}\n]}\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
