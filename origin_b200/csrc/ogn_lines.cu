// step08: line estimation (estimation_line / GridAnalysis / method_PCA_wgt / LS_deconv_wgt, lib_origin.py:1482-1938).
//
// For every detection the reference cuts a (P + 2 size_grid)^2 x nz minicube out of the RAW cube, and for each of the
// (1 + 2 size_grid)^2 spatial offsets of its grid runs method_PCA_wgt on the P x P x nz window: two rank-1 SVDs
// (scipy svds, k = 1) of a [nz][P*P] matrix, two orthogonal projections and two weighted least-squares
// deconvolutions with the FSF.  The windows are independent, so they are processed here as ONE BATCH of problems:
// every kernel takes the problem index from the grid, the tall matrices [nz][P*P] of a problem are contiguous
// (their rows are what a block or a warp streams), and all arithmetic is FP64 like the reference.
//
//   build_kernel        Xs = raw / sqrt(var), W = 1 / sqrt(var) (0 outside the image: raw 0 / var +inf there, as
//                       estimation_line pads its minicube, :1893-1897), Xc = Xs - row mean                (:1574-1579)
//   gram_kernel / symv  the first left singular vector comes from the Gram matrix G = X^T X (P*P x P*P, the small side,
//                       as scipy's svds works on the smaller of X^H X / X X^H): G is formed once per SVD and the
//                       batched Lanczos iteration (full reorthogonalisation, restarted, tridiagonal solve on the
//                       host, as ogn_pca.cu) runs on it, 2 nz / (P*P) times cheaper per step than on X X^T;
//                       u = X v / |X v|.  OGN_LINES_NO_GRAM=1 iterates on X X^T instead (gemv_t_* / gemv_n)
//   project_deconv      residual = Xs - u (u^T X) and LS_deconv_wgt on it, fused per wavelength          (:1583-1587, :1482-1510)
//   clean_kernel        data_clean = (data - psf * line) / sqrt(var), centred                            (:1589-1597)
//   dct_denoise         U = D0 D0^T u with the first order_dct + 1 DCT atoms                              (:1600-1603)
//
// The host part of GridAnalysis (peak search, flux / mse criteria, :1700-1790) consumes the two [nz] vectors each
// problem returns; it lives in origin_b200/lib_origin.py.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "ogn_common.cuh"
#include "ogn_lanczos.cuh"

namespace {
using namespace ogn_lz;

constexpr int LT = 256;
constexpr int EL_M = 64;          // most Krylov vectors per restart cycle (layout of Q / scal / y)
constexpr int EL_M_DEFAULT = 40;  // ... used unless OGN_LINES_KRYLOV says otherwise
constexpr int EL_CYCLES = 200;
constexpr int EL_ZSEG = 128;      // wavelengths per partial sum of X^T q

template <typename T>
__global__ void build_kernel(const T *__restrict__ raw, const T *__restrict__ var, int nz, int ny, int nx, int P,
                             const int *__restrict__ centres, double *__restrict__ Xs, double *__restrict__ W,
                             double *__restrict__ Xc) {
    const int z = blockIdx.x, p = blockIdx.y, n = P * P, half = P / 2;
    const int cy = centres[2 * p], cx = centres[2 * p + 1];
    const size_t base = ((size_t)p * nz + z) * n;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int yy = cy - half + j / P, xx = cx - half + j % P;
        double w = 0.0, v = 0.0;
        if (yy >= 0 && yy < ny && xx >= 0 && xx < nx) {
            const size_t o = ((size_t)z * ny + yy) * nx + xx;
            w = 1.0 / sqrt((double)var[o]);               // var = +inf (masked / NaN voxels, origin.py:262-274): weight 0
            v = (double)raw[o] * w;
        }
        Xs[base + j] = v;
        W[base + j] = w;
        s += v;
    }
    const double mean = block_sum(s) / n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) Xc[base + j] = Xs[base + j] - mean;
}

// part[p][seg][j] = sum_{z in seg} v[p][z] M[p][z][j]
__global__ void gemv_t_partial_kernel(const double *__restrict__ M, int nz, int n, const double *__restrict__ v, int ldv,
                                      double *__restrict__ part, int nseg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y, p = blockIdx.z;
    if (j >= n) return;
    const int z0 = seg * EL_ZSEG, z1 = min(nz, z0 + EL_ZSEG);
    const double *Mp = M + (size_t)p * nz * n;
    const double *vp = v + (size_t)p * ldv;
    double a = 0.0;
    for (int z = z0; z < z1; ++z) a = fma(vp[z], Mp[(size_t)z * n + j], a);
    part[((size_t)p * nseg + seg) * n + j] = a;
}
__global__ void gemv_t_finish_kernel(const double *__restrict__ part, int nseg, int n, double *__restrict__ c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (j >= n) return;
    double a = 0.0;
    for (int g = 0; g < nseg; ++g) a += part[((size_t)p * nseg + g) * n + j];
    c[(size_t)p * n + j] = a;
}
// y[p][z] = sum_j M[p][z][j] c[p][j]
__global__ void gemv_n_kernel(const double *__restrict__ M, int nz, int n, const double *__restrict__ c, size_t ldc,
                              double *__restrict__ y, int ldy) {
    const int z = blockIdx.x, p = blockIdx.y;
    const double *row = M + ((size_t)p * nz + z) * n;
    const double *cp = c + (size_t)p * ldc;
    double a = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) a = fma(row[j], cp[j], a);
    a = block_sum(a);
    if (threadIdx.x == 0) y[(size_t)p * ldy + z] = a;
}

// G[p] = M[p]^T M[p] for the [nz][n] matrices M[p] (n x n, symmetric): one block per 64 x 64 tile of the upper
// triangle, 8 x 4 outputs per thread, wavelengths staged through shared memory GK at a time and summed in
// order (deterministic); off-diagonal tiles are mirrored on store.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(128) gram_kernel(const double *__restrict__ M, int nz, int n, double *__restrict__ G, int ntile) {
    __shared__ double sa[GK][GT], sb[GK][GT];
    int t = blockIdx.x, ta = 0;
    while (t >= ntile - ta) { t -= ntile - ta; ++ta; }
    const int tb = ta + t, p = blockIdx.y;
    const double *Mp = M + (size_t)p * nz * n;
    double *Gp = G + (size_t)p * n * n;
    const int a0 = ta * GT, b0 = tb * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // rows a0 + ty + 8 i, columns b0 + tx + 16 j
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int z0 = 0; z0 < nz; z0 += GK) {
        for (int i = threadIdx.x; i < GK * GT; i += 128) {
            const int zz = i / GT, c = i % GT, z = z0 + zz;
            sa[zz][c] = (z < nz && a0 + c < n) ? Mp[(size_t)z * n + a0 + c] : 0.0;
            sb[zz][c] = (z < nz && b0 + c < n) ? Mp[(size_t)z * n + b0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int zz = 0; zz < GK; ++zz) {
            double av[8], bv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = sa[zz][ty + 8 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sb[zz][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = a0 + ty + 8 * i, b = b0 + tx + 16 * j;
            if (a < n && b < n) {
                Gp[(size_t)a * n + b] = acc[i][j];
                if (ta != tb) Gp[(size_t)b * n + a] = acc[i][j];
            }
        }
}
// y[p][r] = sum_j G[p][r][j] q[p][j]: one warp per row, 8 rows per block
__global__ void symv_kernel(const double *__restrict__ G, int n, const double *__restrict__ q, size_t ldq, double *__restrict__ y,
                            int ldy) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, p = blockIdx.y;
    if (r >= n) return;
    const double *row = G + ((size_t)p * n + r) * n, *qp = q + (size_t)p * ldq;
    double a = 0.0;
    for (int j = lane; j < n; j += 32) a = fma(row[j], qp[j], a);
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) y[(size_t)p * ldy + r] = a;
}
__global__ void normalise_kernel(double *__restrict__ u, int nz) {
    double *up = u + (size_t)blockIdx.x * nz;
    double s = 0.0;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) s = fma(up[i], up[i], s);
    const double nrm = sqrt(block_sum(s));
    for (int i = threadIdx.x; i < nz; i += blockDim.x) up[i] = nrm > 0.0 ? up[i] / nrm : 0.0;
}

// ---- batched Lanczos vector kernels: one block per problem; Q[p][EL_M + 1][len], scal[p][2 EL_M] --------------
__global__ void seed_kernel(double *__restrict__ Q, int nz, size_t qstride) {
    double *q = Q + (size_t)blockIdx.x * qstride;
    double s = 0.0;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) {
        const double v = 1.0 + 0.5 * sin(0.7 * i + 0.3) + 0.25 * cos(2.3 * i);
        q[i] = v;
        s = fma(v, v, s);
    }
    const double nrm = sqrt(block_sum(s));
    for (int i = threadIdx.x; i < nz; i += blockDim.x) q[i] /= nrm;
}
// w -= alpha q_j + beta_{j-1} q_{j-1}, full reorthogonalisation against q_0..q_j (twice), beta_j = ||w||, q_{j+1} = w / beta_j
__global__ void lanczos_step_kernel(double *__restrict__ Q, double *__restrict__ w, double *__restrict__ scal, int j, int nz,
                                    size_t qstride) {
    __shared__ double h[EL_M + 1];
    const int p = blockIdx.x;
    double *Qp = Q + (size_t)p * qstride, *wp = w + (size_t)p * nz, *sp = scal + (size_t)p * 2 * EL_M;
    const double *qj = Qp + (size_t)j * nz;
    double s = 0.0;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) s = fma(qj[i], wp[i], s);
    const double alpha = block_sum(s);
    const double bprev = j ? sp[EL_M + j - 1] : 0.0;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) wp[i] -= alpha * qj[i] + (j ? bprev * qj[i - nz] : 0.0);
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {
        for (int k = 0; k <= j; ++k) {
            double t = 0.0;
            for (int i = threadIdx.x; i < nz; i += blockDim.x) t = fma(Qp[(size_t)k * nz + i], wp[i], t);
            t = block_sum(t);
            if (threadIdx.x == 0) h[k] = t;
            __syncthreads();
        }
        for (int i = threadIdx.x; i < nz; i += blockDim.x) {
            double v = wp[i];
            for (int k = 0; k <= j; ++k) v -= h[k] * Qp[(size_t)k * nz + i];
            wp[i] = v;
        }
        __syncthreads();
    }
    s = 0.0;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) s = fma(wp[i], wp[i], s);
    const double beta = sqrt(block_sum(s));
    double *qn = Qp + (size_t)(j + 1) * nz;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) qn[i] = beta > 0.0 ? wp[i] / beta : 0.0;
    if (threadIdx.x == 0) { sp[j] = alpha; sp[EL_M + j] = beta; }
}
// u[p] = normalised sum_i y[p][i] Q[p][i]; optionally restart: Q[p][0] = u[p]
__global__ void combine_kernel(double *__restrict__ Q, const double *__restrict__ y, const int *__restrict__ kdim, int nz,
                               size_t qstride, double *__restrict__ u, int restart) {
    const int p = blockIdx.x, k = kdim[p];
    double *Qp = Q + (size_t)p * qstride, *up = u + (size_t)p * nz;
    const double *yp = y + (size_t)p * EL_M;
    double s = 0.0;
    for (int t = threadIdx.x; t < nz; t += blockDim.x) {
        double v = 0.0;
        for (int i = 0; i < k; ++i) v = fma(yp[i], Qp[(size_t)i * nz + t], v);
        up[t] = v;
        s = fma(v, v, s);
    }
    const double nrm = sqrt(block_sum(s));
    for (int t = threadIdx.x; t < nz; t += blockDim.x) {
        const double v = nrm > 0.0 ? up[t] / nrm : 0.0;
        up[t] = v;
        if (restart) Qp[t] = v;
    }
}

// FSF sample j of wavelength z for one window.  Single field (coef == nullptr): psf[z][j].  Weighted mosaic: the
// reference combines the fields' FSFs with the weight maps cut to the window, sum_f wgt_f[j] psf_f[z][j]
// (GridAnalysis, lib_origin.py:1713-1717; products rounded, then added in field order like np.sum(axis=0));
// `coef` holds the window's [nf][n] factors, psf the fields' cubes [nf][nz][n].
__device__ __forceinline__ double fsf_sample(const double *__restrict__ psf, const double *__restrict__ coef, int nf, int nz,
                                             int n, int z, int j) {
    if (!coef) return psf[(size_t)z * n + j];
    double a = 0.0;
    for (int f = 0; f < nf; ++f) a = __dadd_rn(a, __dmul_rn(coef[(size_t)f * n + j], psf[((size_t)f * nz + z) * n + j]));
    return a;
}

// residual = Xs - u c^T ; line[z] = varest * sum_j psf W residual ; varest[z] = 1 / sum_j (psf W)^2   (LS_deconv_wgt)
__global__ void project_deconv_kernel(const double *__restrict__ Xs, const double *__restrict__ W, const double *__restrict__ psf,
                                      const double *__restrict__ coef, int nf, int nz, int n, const double *__restrict__ u,
                                      const double *__restrict__ c, double *__restrict__ line, double *__restrict__ linevar) {
    const int z = blockIdx.x, p = blockIdx.y;
    const size_t base = ((size_t)p * nz + z) * n;
    const double uz = u[(size_t)p * nz + z];
    const double *cp = c + (size_t)p * n;
    const double *cf = coef ? coef + (size_t)p * nf * n : nullptr;
    double num = 0.0, den = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double pw = fsf_sample(psf, cf, nf, nz, n, z, j) * W[base + j];
        num = fma(pw, Xs[base + j] - uz * cp[j], num);
        den = fma(pw, pw, den);
    }
    num = block_sum(num);
    den = block_sum(den);
    if (threadIdx.x == 0) {
        const double ve = 1.0 / den;                 // 1 / 0 = inf and 0 * inf = NaN, as numpy
        line[(size_t)p * nz + z] = num * ve;
        linevar[(size_t)p * nz + z] = ve;
    }
}

// Xc = (data - psf line) / sqrt(var), rows centred:  Xs - psf line W, minus the row mean
__global__ void clean_kernel(const double *__restrict__ Xs, const double *__restrict__ W, const double *__restrict__ psf,
                             const double *__restrict__ coef, int nf, int nz, int n, const double *__restrict__ line,
                             double *__restrict__ Xc) {
    const int z = blockIdx.x, p = blockIdx.y;
    const size_t base = ((size_t)p * nz + z) * n;
    const double lz = line[(size_t)p * nz + z];
    const double *cf = coef ? coef + (size_t)p * nf * n : nullptr;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        // conv_wgt multiplies by (|psf| > 0): a no-op for the product psf * line; a NaN line (no valid voxel at this
        // wavelength) times a zero weight is NaN in numpy as well
        const double v = Xs[base + j] - fsf_sample(psf, cf, nf, nz, n, z, j) * lz * W[base + j];
        Xc[base + j] = v;
        s += v;
    }
    const double mean = block_sum(s) / n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) Xc[base + j] -= mean;
}

// U = D0 (D0^T u): projection on the first M DCT atoms (orthogonal_projection(D0, U), lib_origin.py:1601-1603)
__global__ void dct_denoise_kernel(const double *__restrict__ d0, int M, int nz, double *__restrict__ u) {
    extern __shared__ double a[];
    double *up = u + (size_t)blockIdx.x * nz;
    for (int i = 0; i < M; ++i) {
        double s = 0.0;
        for (int z = threadIdx.x; z < nz; z += blockDim.x) s = fma(d0[(size_t)z * M + i], up[z], s);
        s = block_sum(s);
        if (threadIdx.x == 0) a[i] = s;
        __syncthreads();
    }
    for (int z = threadIdx.x; z < nz; z += blockDim.x) {
        double v = 0.0;
        for (int i = 0; i < M; ++i) v = fma(d0[(size_t)z * M + i], a[i], v);
        up[z] = v;
    }
}

struct LineWork {
    double *Xs, *W, *Xc, *part, *c, *u, *w, *Q, *scal, *y, *psf, *d0, *line1, *var1, *G, *v;
    const double *coef;   // weighted mosaics: [npos][nf][P*P] window factors of the fields' FSFs (else nullptr)
    int *kdim, *centres;
    int nseg, nf;
    size_t qstride;
};

int gemv_t(ogn_ctx *ctx, const LineWork &wk, const double *M, int nz, int n, int nb, const double *v, int ldv, double *c) {
    gemv_t_partial_kernel<<<dim3(ogn_div_up(n, 128), wk.nseg, nb), 128, 0, ctx->stream>>>(M, nz, n, v, ldv, wk.part, wk.nseg);
    OGN_LAUNCH_CHECK("gemv_t_partial_kernel");
    gemv_t_finish_kernel<<<dim3(ogn_div_up(n, 128), nb), 128, 0, ctx->stream>>>(wk.part, wk.nseg, n, c);
    OGN_LAUNCH_CHECK("gemv_t_finish_kernel");
    return OGN_OK;
}

// first left singular vectors of the nb matrices Xc[p] into wk.u[p]
int batched_top_vectors(ogn_ctx *ctx, const LineWork &wk, int nz, int n, int nb, int *matvecs) {
    const bool gram = wk.G != nullptr;
    const int len = gram ? n : nz;   // length of the Lanczos vectors
    static const int m_cfg = getenv("OGN_LINES_KRYLOV") ? std::max(2, std::min(EL_M, atoi(getenv("OGN_LINES_KRYLOV")))) : EL_M_DEFAULT;
    const int m = std::min(m_cfg, len);
    if (gram) {
        const int ntile = ogn_div_up(n, GT);
        gram_kernel<<<dim3(ntile * (ntile + 1) / 2, nb), 128, 0, ctx->stream>>>(wk.Xc, nz, n, wk.G, ntile);
        OGN_LAUNCH_CHECK("gram_kernel");
    }
    seed_kernel<<<nb, 1024, 0, ctx->stream>>>(wk.Q, len, wk.qstride);
    OGN_LAUNCH_CHECK("seed_kernel");
    std::vector<double> host((size_t)nb * 2 * EL_M), yall((size_t)nb * EL_M, 0.0), alpha(m), beta(m), y;
    std::vector<int> kdim(nb, m);
    std::vector<char> done(nb, 0);
    for (int cycle = 0; cycle < EL_CYCLES; ++cycle) {
        for (int j = 0; j < m; ++j) {
            if (gram) {
                symv_kernel<<<dim3(ogn_div_up(n, 8), nb), 256, 0, ctx->stream>>>(wk.G, n, wk.Q + (size_t)j * len, wk.qstride, wk.w, len);
                OGN_LAUNCH_CHECK("symv_kernel");                                                           // w = G q_j
            } else {
                OGN_TRY(gemv_t(ctx, wk, wk.Xc, nz, n, nb, wk.Q + (size_t)j * nz, (int)wk.qstride, wk.c));   // c = X^T q_j
                gemv_n_kernel<<<dim3(nz, nb), 128, 0, ctx->stream>>>(wk.Xc, nz, n, wk.c, n, wk.w, nz);    // w = X c
                OGN_LAUNCH_CHECK("gemv_n_kernel");
            }
            ++*matvecs;
            lanczos_step_kernel<<<nb, 1024, 0, ctx->stream>>>(wk.Q, wk.w, wk.scal, j, len, wk.qstride);
            OGN_LAUNCH_CHECK("lanczos_step_kernel");
        }
        OGN_CUDA(cudaMemcpyAsync(host.data(), wk.scal, host.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
        bool all = true;
        for (int p = 0; p < nb; ++p) {
            const double *hp = host.data() + (size_t)p * 2 * EL_M;
            for (int j = 0; j < m; ++j) { alpha[j] = hp[j]; beta[j] = hp[EL_M + j]; }
            int k = m;
            double scale = 0.0;
            for (int j = 0; j < m; ++j) scale = std::max(scale, fabs(alpha[j]));
            for (int j = 0; j < m - 1; ++j)
                if (!(beta[j] > 1e-14 * scale)) { k = j + 1; break; }
            double theta = 0.0;
            tridiag_top(alpha, beta, k, &theta, &y);
            kdim[p] = k;
            for (int i = 0; i < k; ++i) yall[(size_t)p * EL_M + i] = y[i];
            const double resid = k < m ? 0.0 : fabs(beta[k - 1] * y[k - 1]);
            done[p] = !(theta > 0.0) || resid <= 1e-13 * theta;
            all = all && done[p];
        }
        OGN_CUDA(cudaMemcpyAsync(wk.y, yall.data(), yall.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        OGN_CUDA(cudaMemcpyAsync(wk.kdim, kdim.data(), kdim.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        combine_kernel<<<nb, 1024, 0, ctx->stream>>>(wk.Q, wk.y, wk.kdim, len, wk.qstride, gram ? wk.v : wk.u, all ? 0 : 1);
        OGN_LAUNCH_CHECK("combine_kernel");
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));   // yall / kdim are host vectors reused by the next cycle
        if (all) break;
    }
    if (gram) {   // u = X v / |X v|
        gemv_n_kernel<<<dim3(nz, nb), 128, 0, ctx->stream>>>(wk.Xc, nz, n, wk.v, n, wk.u, nz);
        OGN_LAUNCH_CHECK("gemv_n_kernel");
        normalise_kernel<<<nb, 1024, 0, ctx->stream>>>(wk.u, nz);
        OGN_LAUNCH_CHECK("normalise_kernel");
    }
    return OGN_OK;
}

template <typename T>
int run_lines(ogn_ctx *ctx, const T *raw, const T *var, int nz, int ny, int nx, int P, const int *d_centres, int nb,
              int order_dct, const LineWork &wk, const double *coef, double *d_line, double *d_var, int *matvecs) {
    const int n = P * P;   // coef: the factors of THIS batch's first window (or nullptr)
    build_kernel<T><<<dim3(nz, nb), LT, 0, ctx->stream>>>(raw, var, nz, ny, nx, P, d_centres, wk.Xs, wk.W, wk.Xc);
    OGN_LAUNCH_CHECK("build_kernel");
    // first PCA: continuum model from the principal vector of the centred, standardised window (:1578-1584)
    OGN_TRY(batched_top_vectors(ctx, wk, nz, n, nb, matvecs));
    OGN_TRY(gemv_t(ctx, wk, wk.Xc, nz, n, nb, wk.u, nz, wk.c));
    project_deconv_kernel<<<dim3(nz, nb), LT, 0, ctx->stream>>>(wk.Xs, wk.W, wk.psf, coef, wk.nf, nz, n, wk.u, wk.c, wk.line1, wk.var1);
    OGN_LAUNCH_CHECK("project_deconv_kernel");
    // remove the first line estimate convolved with the FSF, second PCA (:1589-1598)
    clean_kernel<<<dim3(nz, nb), LT, 0, ctx->stream>>>(wk.Xs, wk.W, wk.psf, coef, wk.nf, nz, n, wk.line1, wk.Xc);
    OGN_LAUNCH_CHECK("clean_kernel");
    OGN_TRY(batched_top_vectors(ctx, wk, nz, n, nb, matvecs));
    if (order_dct >= 0) {
        const int M = order_dct + 1;
        dct_denoise_kernel<<<nb, 1024, M * sizeof(double), ctx->stream>>>(wk.d0, M, nz, wk.u);
        OGN_LAUNCH_CHECK("dct_denoise_kernel");
    }
    // continuum = U U^T data_st_pca (the UNcentred standardised window, :1606), final LS deconvolution (:1611)
    OGN_TRY(gemv_t(ctx, wk, wk.Xs, nz, n, nb, wk.u, nz, wk.c));
    project_deconv_kernel<<<dim3(nz, nb), LT, 0, ctx->stream>>>(wk.Xs, wk.W, wk.psf, coef, wk.nf, nz, n, wk.u, wk.c, d_line, d_var);
    OGN_LAUNCH_CHECK("project_deconv_kernel");
    return OGN_OK;
}

}  // namespace

// Shared body of the two entry points: nf = 1 and coef == nullptr for a single field.
static int line_estimates_impl(ogn_ctx *ctx, const char *who, const void *raw, const void *var, int dtype, int nz, int ny,
                               int nx, const double *psf, int nf, int P, const double *coef, const int *centres, int npos,
                               int order_dct, double *line, double *linevar, int *info) {
    if (!ctx) return OGN_ERR_ARG;
    if (!raw || !var || !psf || !centres || !line || !linevar || nz <= 1 || ny <= 0 || nx <= 0 || npos <= 0 || P < 1 ||
        !(P & 1) || nf < 1 || (nf > 1 && !coef))
        return ogn_fail(ctx, OGN_ERR_ARG, "%s: bad arguments", who);
    if (dtype != OGN_F32 && dtype != OGN_F64) return ogn_fail(ctx, OGN_ERR_ARG, "%s: unknown dtype", who);
    if (order_dct + 1 > nz || order_dct > 255) return ogn_fail(ctx, OGN_ERR_ARG, "%s: order_dct out of range", who);
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t es = dtype == OGN_F64 ? 8 : 4, vol = (size_t)nz * ny * nx;
    const int n = P * P;
    const void *d_raw = nullptr, *d_var = nullptr, *d_psf = nullptr, *d_cen = nullptr;
    OGN_TRY(ogn_input(ctx, "el_raw", raw, vol * es, &d_raw));
    OGN_TRY(ogn_input(ctx, "el_var", var, vol * es, &d_var));
    OGN_TRY(ogn_input(ctx, "el_psf", psf, (size_t)nf * nz * n * 8, &d_psf));
    OGN_TRY(ogn_input(ctx, "el_centres", centres, (size_t)npos * 2 * sizeof(int), &d_cen));
    const void *d_coef = nullptr;
    if (coef) OGN_TRY(ogn_input(ctx, "el_coef", coef, (size_t)npos * nf * n * 8, &d_coef));
    // OGN_LINES_NO_GRAM=1: Lanczos on X X^T (length-nz vectors, two passes over X per step) instead of on the Gram matrix
    static const bool no_gram = getenv("OGN_LINES_NO_GRAM") != nullptr;
    const int len = std::max(n, nz);
    // problems per batch: three [nz][P*P] FP64 matrices (+ the Gram matrix) each, within ~12 GB of scratch
    const size_t per = (size_t)nz * n * 8 * 3 + (no_gram ? 0 : (size_t)n * n * 8);
    const int nb_max = (int)std::max<size_t>(1, std::min<size_t>((size_t)npos, ((size_t)12 << 30) / per));
    LineWork wk;
    wk.nseg = ogn_div_up(nz, EL_ZSEG);
    wk.qstride = (size_t)(EL_M + 1) * len;
    wk.psf = const_cast<double *>(static_cast<const double *>(d_psf));
    wk.coef = static_cast<const double *>(d_coef);
    wk.nf = nf;
    OGN_TRY(ogn_scratch_t(ctx, "el_Xs", (size_t)nb_max * nz * n, &wk.Xs));
    OGN_TRY(ogn_scratch_t(ctx, "el_W", (size_t)nb_max * nz * n, &wk.W));
    OGN_TRY(ogn_scratch_t(ctx, "el_Xc", (size_t)nb_max * nz * n, &wk.Xc));
    OGN_TRY(ogn_scratch_t(ctx, "el_part", (size_t)nb_max * wk.nseg * n, &wk.part));
    OGN_TRY(ogn_scratch_t(ctx, "el_c", (size_t)nb_max * n, &wk.c));
    OGN_TRY(ogn_scratch_t(ctx, "el_u", (size_t)nb_max * nz, &wk.u));
    OGN_TRY(ogn_scratch_t(ctx, "el_w", (size_t)nb_max * len, &wk.w));
    OGN_TRY(ogn_scratch_t(ctx, "el_v", (size_t)nb_max * n, &wk.v));
    wk.G = nullptr;
    if (!no_gram) OGN_TRY(ogn_scratch_t(ctx, "el_G", (size_t)nb_max * n * n, &wk.G));
    ctx->variants["step08"] = no_gram ? "lanczos:xxt" : "lanczos:gram";
    OGN_TRY(ogn_scratch_t(ctx, "el_Q", (size_t)nb_max * wk.qstride, &wk.Q));
    OGN_TRY(ogn_scratch_t(ctx, "el_scal", (size_t)nb_max * 2 * EL_M, &wk.scal));
    OGN_TRY(ogn_scratch_t(ctx, "el_y", (size_t)nb_max * EL_M, &wk.y));
    OGN_TRY(ogn_scratch_t(ctx, "el_kdim", (size_t)nb_max, &wk.kdim));
    OGN_TRY(ogn_scratch_t(ctx, "el_line1", (size_t)nb_max * nz, &wk.line1));
    OGN_TRY(ogn_scratch_t(ctx, "el_var1", (size_t)nb_max * nz, &wk.var1));
    wk.d0 = nullptr;
    if (order_dct >= 0) {   // DCTMAT(nl, order_dct), lib_origin.py:127-146
        const int M = order_dct + 1;
        std::vector<double> h((size_t)nz * M);
        const double scale = sqrt(2.0 / nz), step = M_PI / nz;
        for (int z = 0; z < nz; ++z)
            for (int j = 0; j < M; ++j) h[(size_t)z * M + j] = scale * cos((z + 0.5) * step * j) * (j == 0 ? 1.0 / sqrt(2.0) : 1.0);
        OGN_TRY(ogn_scratch_t(ctx, "el_d0", h.size(), &wk.d0));
        OGN_CUDA(cudaMemcpyAsync(wk.d0, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    void *d_line = nullptr, *d_lvar = nullptr;
    OGN_TRY(ogn_output(ctx, "el_line_out", line, (size_t)npos * nz * 8, &d_line));
    OGN_TRY(ogn_output(ctx, "el_var_out", linevar, (size_t)npos * nz * 8, &d_lvar));
    int matvecs = 0;
    ogn_timer t_(ctx, "line_estimates");
    for (int p0 = 0; p0 < npos; p0 += nb_max) {
        const int nb = std::min(nb_max, npos - p0);
        const int *cen = static_cast<const int *>(d_cen) + 2 * p0;
        double *ol = static_cast<double *>(d_line) + (size_t)p0 * nz, *ov = static_cast<double *>(d_lvar) + (size_t)p0 * nz;
        const double *cf = wk.coef ? wk.coef + (size_t)p0 * nf * n : nullptr;
        if (dtype == OGN_F64)
            OGN_TRY(run_lines<double>(ctx, (const double *)d_raw, (const double *)d_var, nz, ny, nx, P, cen, nb, order_dct, wk, cf, ol, ov, &matvecs));
        else
            OGN_TRY(run_lines<float>(ctx, (const float *)d_raw, (const float *)d_var, nz, ny, nx, P, cen, nb, order_dct, wk, cf, ol, ov, &matvecs));
    }
    OGN_TRY(ogn_output_commit(ctx, line, d_line, (size_t)npos * nz * 8));
    OGN_TRY(ogn_output_commit(ctx, linevar, d_lvar, (size_t)npos * nz * 8));
    if (info) { info[0] = matvecs; info[1] = nb_max; }
    return ogn_finish_call(ctx);
}

// method_PCA_wgt (lib_origin.py:1535-1617) for a batch of P x P x nz windows of the raw cube, i.e. everything
// GridAnalysis (:1620-1790) computes per grid offset before its scalar criteria.
//   raw, var     [nz][ny][nx], float32 / float64 (`dtype`), host or device; var = +inf marks invalid voxels
//   psf          [nz][P][P] float64, single field (host or device)
//   centres      [npos][2] int32 (y, x): centre of each window; windows may stick out of the image
//   order_dct    order of the DCT that denoises the second eigenvector (< 0: PCA LS only, order_dct=None)
//   line, linevar [npos][nz] float64 out (host or device): estimated line and its theoretical variance
extern "C" int ogn_line_estimates(ogn_ctx *ctx, const void *raw, const void *var, int dtype, int nz, int ny, int nx,
                                  const double *psf, int P, const int *centres, int npos, int order_dct, double *line,
                                  double *linevar, int *info) {
    return line_estimates_impl(ctx, "ogn_line_estimates", raw, var, dtype, nz, ny, nx, psf, 1, P, nullptr, centres, npos,
                               order_dct, line, linevar, info);
}

// The same for weighted mosaics (wght is not None): the FSF of a window is the combination
// sum_f coef[p][f][j] psf[f][z][j] of the fields' FSFs, where coef holds the weight maps cut to the window
// (GridAnalysis, lib_origin.py:1713-1717; the host mirror builds it, including the way the reference's loop
// compounds the factors from one grid offset to the next).
//   psf   [nf][nz][P][P] float64;  coef  [npos][nf][P][P] float64 (host or device)
extern "C" int ogn_line_estimates_fields(ogn_ctx *ctx, const void *raw, const void *var, int dtype, int nz, int ny, int nx,
                                         const double *psf, int nf, int P, const double *coef, const int *centres, int npos,
                                         int order_dct, double *line, double *linevar, int *info) {
    if (ctx && !coef) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_line_estimates_fields: coef is required");
    return line_estimates_impl(ctx, "ogn_line_estimates_fields", raw, var, dtype, nz, ny, nx, psf, nf, P, coef, centres,
                               npos, order_dct, line, linevar, info);
}
