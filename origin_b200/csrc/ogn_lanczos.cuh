// Pieces shared by the two Lanczos users (step04 greedy PCA in ogn_pca.cu, step08 line estimation in ogn_lines.cu):
// a deterministic block-wide FP64 sum and the host-side eigen-solver of the small tridiagonal matrices.
#pragma once

#include <math.h>

#include <vector>

#include <cuda_runtime.h>

namespace ogn_lz {

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    if (warp == 0) {
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// largest eigenpair of the symmetric tridiagonal (alpha, beta) of order k: cyclic Jacobi on the dense form
inline void tridiag_top(const std::vector<double> &alpha, const std::vector<double> &beta, int k, double *theta, std::vector<double> *y) {
    std::vector<double> A((size_t)k * k, 0.0), V((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) {
        A[(size_t)i * k + i] = alpha[i];
        V[(size_t)i * k + i] = 1.0;
        if (i + 1 < k) A[(size_t)i * k + i + 1] = A[(size_t)(i + 1) * k + i] = beta[i];
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < k; ++p) {
            diag += A[(size_t)p * k + p] * A[(size_t)p * k + p];
            for (int q = p + 1; q < k; ++q) off += A[(size_t)p * k + q] * A[(size_t)p * k + q];
        }
        if (off <= 1e-34 * diag) break;
        for (int p = 0; p < k; ++p)
            for (int q = p + 1; q < k; ++q) {
                const double apq = A[(size_t)p * k + q];
                if (fabs(apq) < 1e-300) continue;
                const double tau = (A[(size_t)q * k + q] - A[(size_t)p * k + p]) / (2.0 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int i = 0; i < k; ++i) {
                    const double aip = A[(size_t)i * k + p], aiq = A[(size_t)i * k + q];
                    A[(size_t)i * k + p] = c * aip - s * aiq;
                    A[(size_t)i * k + q] = s * aip + c * aiq;
                }
                for (int i = 0; i < k; ++i) {
                    const double api = A[(size_t)p * k + i], aqi = A[(size_t)q * k + i];
                    A[(size_t)p * k + i] = c * api - s * aqi;
                    A[(size_t)q * k + i] = s * api + c * aqi;
                }
                for (int i = 0; i < k; ++i) {
                    const double vip = V[(size_t)i * k + p], viq = V[(size_t)i * k + q];
                    V[(size_t)i * k + p] = c * vip - s * viq;
                    V[(size_t)i * k + q] = s * vip + c * viq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < k; ++i)
        if (A[(size_t)i * k + i] > A[(size_t)best * k + best]) best = i;
    *theta = A[(size_t)best * k + best];
    y->assign(k, 0.0);
    for (int i = 0; i < k; ++i) (*y)[i] = V[(size_t)i * k + best];
}


}  // namespace ogn_lz
