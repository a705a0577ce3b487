// NumPy-exact pairwise summation, shared by the BPM and metric kernels.
#pragma once
#include <cuda_runtime.h>

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src), which is what np.mean /
// np.nanmean of a contiguous float vector evaluates.  The recursion (split at n/2 rounded
// down to a multiple of 8 until blocks are <= 128 long) is run with an explicit stack: device
// recursion would need more than the default per-thread stack.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename F>
__device__ F pairwise_leaf(const F* a, int n) {
    if (n < 8) {
        F res = (F)0;
        for (int i = 0; i < n; ++i) res = add_rn(res, a[i]);
        return res;
    }
    F r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = add_rn(r[j], a[i + j]);
    }
    F res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])), add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
    for (; i < n; ++i) res = add_rn(res, a[i]);
    return res;
}

template <typename F>
__device__ F pairwise_sum(const F* a, int n) {
    if (n <= 128) return pairwise_leaf(a, n);
    int s_off[40], s_len[40];
    bool s_comb[40];
    F vals[40];
    int sp = 0, vp = 0;
    s_off[sp] = 0; s_len[sp] = n; s_comb[sp] = false; ++sp;
    while (sp > 0) {
        --sp;
        const int off = s_off[sp], len = s_len[sp];
        if (s_comb[sp]) {
            const F r = vals[--vp];
            const F l = vals[--vp];
            vals[vp++] = add_rn(l, r);
        } else if (len <= 128) {
            vals[vp++] = pairwise_leaf(a + off, len);
        } else {
            int n2 = len / 2;
            n2 -= n2 % 8;
            s_off[sp] = off; s_len[sp] = len; s_comb[sp] = true; ++sp;              // combine after both halves
            s_off[sp] = off + n2; s_len[sp] = len - n2; s_comb[sp] = false; ++sp;    // right half (evaluated second)
            s_off[sp] = off; s_len[sp] = n2; s_comb[sp] = false; ++sp;               // left half (evaluated first)
        }
    }
    return vals[0];
}
__device__ __forceinline__ float pairwise_sum_f32(const float* a, int n) { return pairwise_sum<float>(a, n); }
__device__ __forceinline__ double pairwise_sum_f64(const double* a, int n) { return pairwise_sum<double>(a, n); }

