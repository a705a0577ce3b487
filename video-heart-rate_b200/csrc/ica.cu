// Batched FastICA for the analysis harness' ICA measurement:
// /root/reference/analysis/measurement/ica.py:36-72 --
//     signal = float32(window of mean-BGR rows);  signal /= std(signal, axis=0, ddof=1)     (:56-61)
//     sources = FastICA(n_components=3, algorithm="parallel", fun="logcosh", max_iter=300, tol=1e-6,
//                       whiten="unit-variance", random_state=0).fit_transform(signal)        (:36-44, :65)
//     skip the window on ConvergenceWarning (:64-69);  bpm = estimate_bpm(sources, fs)       (:72)
// The algorithm is scikit-learn's (third-party, not under /root/reference; 1.9.0 in this image,
// sklearn/decomposition/_fastica.py): centre, whiten by the SVD of the centred data (singular vectors sign-fixed
// so that their first component is non-negative, singular values decreasing), scale by sqrt(n), symmetric
// decorrelation W <- (W W^T)^(-1/2) W of the seeded 3x3 start matrix, then the parallel fixed point
//     W1 = symdecor( E[tanh(W x) x^T] - diag(E[1 - tanh^2(W x)]) W ),   lim = max_j | |<W1_j, W_j>| - 1 |
// until lim < tol, sources = W K x scaled to unit variance.
//
// One WARP per window: the window's whitened samples live in shared memory, a lane owns every 32nd sample, the
// twelve sums of an iteration are warp-shuffle reductions and the 3x3 eigen-decompositions (Jacobi) run
// redundantly in every lane.  Everything after the float32 preprocessing is float64, whereas scikit-learn keeps
// float32 for float32 input: the fixed point is the same, the iteration count and the last bits are not, so the
// contract is a TOLERANCE one (tests/test_gpu_round2.py: identical spectral-peak bin on the windows where
// scikit-learn converges), not bit parity -- the reference's own set of emitted rows depends on float32 LAPACK
// rounding (whether lim crosses 1e-6 within 300 iterations) and cannot be reproduced by any independent code.
#include "common.cuh"
#include <math.h>

namespace {

struct IcaArgs {
    const double* trace;      // (n_trace, 3)
    int n_trace;
    const int32_t* start;
    const int32_t* len;
    int max_len, max_iter;
    double tol;
    double w_init[9];
    double* sources;          // (n_win, max_len, 3)
    int32_t* n_iter;          // (n_win): iterations used; negated when the fixed point did not converge
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi rotations: A = V diag(ev) V^T.
__device__ void eig3(const double (&Ain)[3][3], double (&ev)[3], double (&V)[3][3]) {
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { A[i][j] = Ain[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int p = r == 2 ? 1 : 0, q = r == 0 ? 1 : 2;          // (0,1) (0,2) (1,2)
            if (A[p][q] == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {                              // A <- A J
                const double akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq;
                A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {                              // A <- J^T A
                const double apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk;
                A[q][k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; ++k) {
                const double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq;
                V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    for (int i = 0; i < 3; ++i) ev[i] = A[i][i];
}

// W <- (W W^T)^(-1/2) W   (sklearn _sym_decorrelation; eigenvalues clipped at the smallest normal double)
__device__ void sym_decorrelation(double (&W)[3][3]) {
    double M[3][3], ev[3], U[3][3], R[3][3], O[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = W[i][0] * W[j][0] + W[i][1] * W[j][1] + W[i][2] * W[j][2];
    eig3(M, ev, U);
    for (int k = 0; k < 3; ++k) ev[k] = 1.0 / sqrt(fmax(ev[k], 2.2250738585072014e-308));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * ev[0] * U[j][0] + U[i][1] * ev[1] * U[j][1] + U[i][2] * ev[2] * U[j][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O[i][j] = R[i][0] * W[0][j] + R[i][1] * W[1][j] + R[i][2] * W[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W[i][j] = O[i][j];
}

__global__ void __launch_bounds__(32) ica_fastica_kernel(const IcaArgs a) {
    extern __shared__ __align__(16) unsigned char sm[];
    double* X1 = reinterpret_cast<double*>(sm);                   // whitened samples, [3][max_len]
    float* xc = reinterpret_cast<float*>(X1 + 3 * a.max_len);     // centred, std-normalised samples, [3][max_len]
    const int w = blockIdx.x, lane = threadIdx.x;
    const int s0 = a.start[w], n = a.len[w];
    double* out = a.sources + (size_t)w * a.max_len * 3;
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    if (n < 4 || n > a.max_len || s0 < 0 || s0 + n > a.n_trace) {
        for (int i = lane; i < a.max_len * 3; i += 32) out[i] = nan;
        if (lane == 0) a.n_iter[w] = 0;
        return;
    }
    const double dn = (double)n;
    // ---- ica.py:53-61: float32 cast, per-channel std (ddof = 1, 0 -> 1), divide -----------------------
    double mean[3], sd[3];
    for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) acc += (double)(float)a.trace[(size_t)(s0 + i) * 3 + c];
        mean[c] = warp_sum(acc) / dn;
        acc = 0.0;
        for (int i = lane; i < n; i += 32) { const double d = (double)(float)a.trace[(size_t)(s0 + i) * 3 + c] - mean[c]; acc += d * d; }
        float s = (float)sqrt(warp_sum(acc) / (dn - 1.0));
        if (s == 0.0f) s = 1.0f;
        sd[c] = (double)s;
    }
    // ---- FastICA: centre (the mean of x / sd) -----------------------------------------------------------
    for (int c = 0; c < 3; ++c) {
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) {
            const float v = __fdiv_rn((float)a.trace[(size_t)(s0 + i) * 3 + c], (float)sd[c]);
            xc[c * a.max_len + i] = v;
            acc += (double)v;
        }
        const float m = (float)(warp_sum(acc) / dn);
        for (int i = lane; i < n; i += 32) xc[c * a.max_len + i] = __fsub_rn(xc[c * a.max_len + i], m);
    }
    __syncwarp();
    // ---- whitening: eigen-decomposition of X X^T = left singular vectors / squared singular values --------
    double Cm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            double acc = 0.0;
            for (int k = lane; k < n; k += 32) acc += (double)xc[i * a.max_len + k] * (double)xc[j * a.max_len + k];
            Cm[i][j] = Cm[j][i] = warp_sum(acc);
        }
    double ev[3], U[3][3];
    eig3(Cm, ev, U);
    int ord[3] = {0, 1, 2};                                          // singular values in decreasing order
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (ev[ord[j]] < ev[ord[j + 1]]) { const int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    double K[3][3];                                                  // K = (u / d)^T, u *= sign(u[0])
    for (int j = 0; j < 3; ++j) {
        const int cidx = ord[j];
        const double d = sqrt(fmax(ev[cidx], 1e-300));
        const double sg = U[0][cidx] < 0.0 ? -1.0 : 1.0;
        for (int c = 0; c < 3; ++c) K[j][c] = sg * U[c][cidx] / d;
    }
    const double sq = sqrt(dn);
    for (int i = lane; i < n; i += 32) {
        const double x0 = xc[i], x1 = xc[a.max_len + i], x2 = xc[2 * a.max_len + i];
        for (int j = 0; j < 3; ++j) X1[j * a.max_len + i] = (K[j][0] * x0 + K[j][1] * x1 + K[j][2] * x2) * sq;
    }
    __syncwarp();
    // ---- parallel fixed point (_ica_par, logcosh, alpha = 1) ---------------------------------------------------
    double W[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W[i][j] = (double)(float)a.w_init[3 * i + j];     // w_init is cast to the data's float32
    sym_decorrelation(W);
    int it = 0;
    bool converged = false;
    for (; it < a.max_iter; ++it) {
        double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, gp[3] = {0, 0, 0};
        for (int i = lane; i < n; i += 32) {
            const double x0 = X1[i], x1 = X1[a.max_len + i], x2 = X1[2 * a.max_len + i];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double t = tanh(W[j][0] * x0 + W[j][1] * x1 + W[j][2] * x2);
                G[j][0] += t * x0; G[j][1] += t * x1; G[j][2] += t * x2;
                gp[j] += 1.0 - t * t;
            }
        }
        double W1[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double g = warp_sum(gp[j]) / dn;
#pragma unroll
            for (int k = 0; k < 3; ++k) W1[j][k] = warp_sum(G[j][k]) / dn - g * W[j][k];
        }
        sym_decorrelation(W1);
        double lim = 0.0;
        for (int j = 0; j < 3; ++j) {
            const double dot = W1[j][0] * W[j][0] + W1[j][1] * W[j][1] + W1[j][2] * W[j][2];
            lim = fmax(lim, fabs(fabs(dot) - 1.0));
        }
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) W[j][k] = W1[j][k];
        if (lim < a.tol) { converged = true; ++it; break; }
    }
    // ---- sources = W K x, scaled to unit variance (whiten="unit-variance") ---------------------------------------
    double WK[3][3];
    for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 3; ++c) WK[j][c] = W[j][0] * K[0][c] + W[j][1] * K[1][c] + W[j][2] * K[2][c];
    double sm1[3] = {0, 0, 0}, sm2[3] = {0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        const double x0 = xc[i], x1 = xc[a.max_len + i], x2 = xc[2 * a.max_len + i];
        for (int j = 0; j < 3; ++j) {
            const double sj = WK[j][0] * x0 + WK[j][1] * x1 + WK[j][2] * x2;
            X1[j * a.max_len + i] = sj;
            sm1[j] += sj;
            sm2[j] += sj * sj;
        }
    }
    double inv[3];
    for (int j = 0; j < 3; ++j) {
        const double m = warp_sum(sm1[j]) / dn;
        const double var = fmax(warp_sum(sm2[j]) / dn - m * m, 0.0);
        inv[j] = var > 0.0 ? 1.0 / sqrt(var) : 1.0;
    }
    __syncwarp();
    for (int i = lane; i < a.max_len; i += 32)
        for (int j = 0; j < 3; ++j) out[(size_t)i * 3 + j] = i < n ? X1[j * a.max_len + i] * inv[j] : nan;
    if (lane == 0) a.n_iter[w] = converged ? it : -it;
}

}  // namespace

extern "C" int vhr_ica_fastica(vhr_ctx* ctx, const double* d_trace, int n_trace, const int32_t* d_start, const int32_t* d_len,
                               int n_win, int max_len, const double* h_w_init, int max_iter, double tol, double* d_sources,
                               int32_t* d_n_iter, void* stream) {
    VHR_REQUIRE(ctx, ctx != nullptr, "null context");
    VHR_REQUIRE(ctx, d_trace && d_start && d_len && h_w_init && d_sources && d_n_iter, "null pointer");
    VHR_REQUIRE(ctx, n_trace >= 1 && n_win >= 1 && max_len >= 1 && max_len <= n_trace, "bad window arguments");
    VHR_REQUIRE(ctx, max_iter >= 1 && tol > 0, "bad iteration arguments");
    IcaArgs a;
    a.trace = d_trace; a.n_trace = n_trace; a.start = d_start; a.len = d_len; a.max_len = max_len; a.max_iter = max_iter;
    a.tol = tol; a.sources = d_sources; a.n_iter = d_n_iter;
    for (int i = 0; i < 9; ++i) a.w_init[i] = h_w_init[i];
    const size_t smem = (size_t)max_len * 3 * (sizeof(double) + sizeof(float));
    if ((long long)smem > ctx->smem_optin) {
        vhr_set_error(ctx, "ica: windows of %d samples need %zu bytes of shared memory", max_len, smem);
        return VHR_ERR_UNSUPPORTED;
    }
    VHR_CHECK_CUDA(ctx, cudaFuncSetAttribute(ica_fastica_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ica_fastica_kernel<<<n_win, 32, smem, (cudaStream_t)stream>>>(a);
    return vhr_after_launch(ctx, "ica_fastica_kernel");
}
