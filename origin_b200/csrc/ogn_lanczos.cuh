// Pieces shared by the two Lanczos users (step04 greedy PCA in ogn_pca.cu, step08 line estimation in ogn_lines.cu):
// a deterministic block-wide FP64 sum and the host-side eigen-solver of the small tridiagonal matrices.
#pragma once

#include <math.h>

#include <cmath>

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

// Largest eigenpair of the symmetric tridiagonal (alpha, beta) of order k (beta[i] couples i and i + 1):
// the eigenvalue by bisection on the Sturm sequence down to the last bit, the eigenvector by inverse iteration
// with a pivoted tridiagonal LU.  O(k) per bisection step / solve: microseconds for the k <= 40 used here (the
// cyclic Jacobi this replaces took 2 ms at k = 40 and was the whole cost of the batched line estimation).
inline void tridiag_top(const std::vector<double> &alpha, const std::vector<double> &beta, int k, double *theta, std::vector<double> *y) {
    y->assign(k, 0.0);
    if (k == 1) {
        *theta = alpha[0];
        (*y)[0] = 1.0;
        return;
    }
    double lo = alpha[0], hi = alpha[0], tnorm = 0.0;
    for (int i = 0; i < k; ++i) {
        const double r = (i ? fabs(beta[i - 1]) : 0.0) + (i + 1 < k ? fabs(beta[i]) : 0.0);
        lo = fmin(lo, alpha[i] - r);
        hi = fmax(hi, alpha[i] + r);
        tnorm = fmax(tnorm, fabs(alpha[i]) + r);
    }
    const double tiny = tnorm > 0.0 ? tnorm * 4.9e-32 : 1e-300;   // ~ eps^2 |T|
    // all k eigenvalues lie below x  <=>  the Sturm sequence of T - x I has k negative terms
    auto all_below = [&](double x) {
        double q = alpha[0] - x;
        if (q == 0.0) q = -tiny;
        if (q >= 0.0) return false;
        for (int i = 1; i < k; ++i) {
            q = (alpha[i] - x) - beta[i - 1] * beta[i - 1] / q;
            if (q == 0.0) q = -tiny;
            if (q >= 0.0) return false;
        }
        return true;
    };
    hi += 2.2e-16 * tnorm + 1e-300;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        if (all_below(mid)) hi = mid; else lo = mid;
    }
    const double th = 0.5 * (lo + hi);
    *theta = th;
    // inverse iteration: (T - th I) z = y, three rounds from a vector with no zero component
    std::vector<double> dl(k), d(k), du(k), du2(k), b(k);
    for (int i = 0; i < k; ++i) b[i] = 1.0 / sqrt((double)k) * ((i & 1) ? 0.9 : 1.1);
    for (int round = 0; round < 3; ++round) {
        for (int i = 0; i < k; ++i) {
            d[i] = alpha[i] - th;
            dl[i] = du[i] = i + 1 < k ? beta[i] : 0.0;
            du2[i] = 0.0;
        }
        for (int i = 0; i + 1 < k; ++i) {
            if (fabs(d[i]) >= fabs(dl[i])) {
                if (d[i] == 0.0) d[i] = tiny;
                const double f = dl[i] / d[i];
                d[i + 1] -= f * du[i];
                b[i + 1] -= f * b[i];
            } else {   // swap rows i and i + 1
                const double f = d[i] / dl[i];
                d[i] = dl[i];
                const double t = d[i + 1];
                d[i + 1] = du[i] - f * t;
                if (i + 2 < k) {
                    du2[i] = du[i + 1];
                    du[i + 1] = -f * du[i + 1];
                }
                du[i] = t;
                const double tb = b[i];
                b[i] = b[i + 1];
                b[i + 1] = tb - f * b[i + 1];
            }
        }
        if (d[k - 1] == 0.0) d[k - 1] = tiny;
        b[k - 1] /= d[k - 1];
        b[k - 2] = (b[k - 2] - du[k - 2] * b[k - 1]) / d[k - 2];
        for (int i = k - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
        double big = 0.0;
        for (int i = 0; i < k; ++i) big = fmax(big, fabs(b[i]));
        if (!(big > 0.0) || !std::isfinite(big)) {   // breakdown: fall back to the start vector of this round
            for (int i = 0; i < k; ++i) b[i] = i == 0 ? 1.0 : 0.0;
            big = 1.0;
        }
        double nrm = 0.0;
        for (int i = 0; i < k; ++i) { b[i] /= big; nrm += b[i] * b[i]; }
        nrm = sqrt(nrm);
        for (int i = 0; i < k; ++i) b[i] /= nrm;
    }
    for (int i = 0; i < k; ++i) (*y)[i] = b[i];
}

}  // namespace ogn_lz
