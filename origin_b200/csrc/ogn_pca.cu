// step04: greedy PCA (Compute_GreedyPCA, lib_origin.py:858-954; O2test :957-974; orthogonal_projection :76-88).
//
// The reference alternates, per area of the field, between a second-order test per spaxel (mean_z x^2), a
// background signature b (mean spectrum of the quietest 1/Noise_population of the spaxels below the threshold),
// the first left singular vector u of the nuisance spectra orthogonalised to b (scipy.sparse.linalg.svds, k = 1),
// and the rank-1 deflation faint -= u (u^T faint) of EVERY spectrum of the area, until no spaxel exceeds the
// threshold.  All of it is passes over tall matrices [nz][spaxels] with spaxels contiguous: HBM-bound
// GEMV-shaped kernels, here in FP64 like the reference (the decisions "test > threshold" steer the iteration).
//
//   colsumsq_*      test[s] = mean_z F[z][s]^2                        (O2test), z-segmented + fixed-order finish
//   mean_cols       b[z] = mean_j F[z][cols[j]]                       background signature
//   gather_cols     X[z][j] = F[z][px[j]]                             nuisance block
//   gemv_t_*        c[j] = sum_z v[z] M[z][j]                         M^T v, z-segmented + fixed-order finish
//   gemv_n          y[z] = sum_j M[z][j] c[j]
//   rank1_scale     X = (X - b c^T) / sum(b^2)                        lib_origin.py:908-909
//   deflate_*       F -= u c^T fused with the new colsumsq partials   :927 + :930 in one pass over the area
//   Lanczos (full reorthogonalisation, restarted) on A = X X^T for u; the top eigenpair of the m x m tridiagonal matrix
//   is found on the host by Sturm bisection + inverse iteration (ogn_lanczos.cuh).  svds(k=1) converges its ARPACK iteration to machine precision; the
//   restart loop here stops at a relative residual of 1e-13, so both give the same vector up to sign, and the
//   projector u u^T does not depend on the sign.
//
// The spaxel selection between two deflations (where / argsort on a vector of `spaxels` doubles) is host code,
// like the reference's, including its indexing quirk: positions inside the COMPRESSED vector test[test > 0] are
// used as column indices of the full block (lib_origin.py:895-903).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "ogn_common.cuh"
#include "ogn_lanczos.cuh"

namespace {
using namespace ogn_lz;

constexpr int PT = 256;
constexpr int LANCZOS_M = 48;          // most Krylov vectors per restart cycle (layout of Q / scal / y)
constexpr int LANCZOS_M_DEFAULT = 16;  // ... used unless OGN_PCA_KRYLOV says otherwise (12 / 16 / 24 / 32 / 48: 89 / 89 / 119 / 142 / 233 ms on the probe)
constexpr int LANCZOS_CYCLES = 200;

template <typename T>
__global__ void gather_area_kernel(const T *__restrict__ cube, int64_t ld, const int64_t *__restrict__ cols, int64_t n,
                                   double *__restrict__ F) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    if (j < n) F[(size_t)z * n + j] = (double)cube[(size_t)z * ld + (cols ? cols[j] : j)];
}
template <typename T>
__global__ void scatter_area_kernel(const double *__restrict__ F, int64_t n, const int64_t *__restrict__ cols, int64_t ld,
                                    T *__restrict__ cube) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    if (j < n) cube[(size_t)z * ld + (cols ? cols[j] : j)] = (T)F[(size_t)z * n + j];
}

__global__ void colsumsq_partial_kernel(const double *__restrict__ F, int nz, int64_t n, int zseg, double *__restrict__ part) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    double a = 0.0;
    for (int z = z0; z < z1; ++z) {
        const double v = F[(size_t)z * n + s];
        a = fma(v, v, a);
    }
    part[(size_t)blockIdx.y * n + s] = a;
}
// out[s] = scale * sum over the segments, in segment order
__global__ void seg_finish_kernel(const double *__restrict__ part, int nseg, int64_t n, double scale, double *__restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double a = 0.0;
    for (int g = 0; g < nseg; ++g) a += part[(size_t)g * n + s];
    out[s] = a * scale;
}

__global__ void mean_cols_kernel(const double *__restrict__ F, int64_t n, const int64_t *__restrict__ cols, int nb,
                                 double *__restrict__ b) {
    const int z = blockIdx.x;
    double a = 0.0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) a += F[(size_t)z * n + cols[j]];
    a = block_sum(a);
    if (threadIdx.x == 0) b[z] = nb > 0 ? a / nb : nan("");   // mean of no column: NaN, like np.mean
}

__global__ void gemv_t_partial_kernel(const double *__restrict__ M, int nz, int64_t n, const double *__restrict__ v, int zseg,
                                      double *__restrict__ part) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    double a = 0.0;
    for (int z = z0; z < z1; ++z) a = fma(v[z], M[(size_t)z * n + j], a);
    part[(size_t)blockIdx.y * n + j] = a;
}

__global__ void gemv_n_kernel(const double *__restrict__ M, int64_t n, const double *__restrict__ c, double *__restrict__ y) {
    const int z = blockIdx.x;
    double a = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) a = fma(M[(size_t)z * n + j], c[j], a);
    a = block_sum(a);
    if (threadIdx.x == 0) y[z] = a;
}

__global__ void rank1_scale_kernel(double *__restrict__ X, int64_t n, const double *__restrict__ b, const double *__restrict__ c,
                                   const double *__restrict__ bb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    if (j < n) X[(size_t)z * n + j] = (X[(size_t)z * n + j] - b[z] * c[j]) / bb[0];
}

// F[z][s] -= u[z] c[s], and the partial column sums of squares of the deflated block
__global__ void deflate_partial_kernel(double *__restrict__ F, int nz, int64_t n, const double *__restrict__ u,
                                       const double *__restrict__ c, int zseg, double *__restrict__ part) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    const double cs = c[s];
    double a = 0.0;
    for (int z = z0; z < z1; ++z) {
        const size_t o = (size_t)z * n + s;
        const double v = F[o] - u[z] * cs;
        F[o] = v;
        a = fma(v, v, a);
    }
    part[(size_t)blockIdx.y * n + s] = a;
}

// ---- small vectors (length nz): one block each ---------------------------------------------------------------
__global__ void dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int n, double *__restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(a[i], b[i], s);
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}
// w -= alpha q + beta qprev   (alpha = q.w is computed here and stored)
__global__ void lanczos_orth_kernel(double *__restrict__ w, const double *__restrict__ q, const double *__restrict__ qprev,
                                    const double *__restrict__ beta_prev, int n, double *__restrict__ alpha_out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(q[i], w[i], s);
    const double alpha = block_sum(s);
    const double bp = qprev ? beta_prev[0] : 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] -= alpha * q[i] + (qprev ? bp * qprev[i] : 0.0);
    if (threadIdx.x == 0) alpha_out[0] = alpha;
}
// full reorthogonalisation against Q[0..k): w -= Q (Q^T w); one block, the k coefficients in shared memory
__global__ void reorth_kernel(const double *__restrict__ Q, int k, int n, double *__restrict__ w) {
    __shared__ double h[LANCZOS_M + 1];
    for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int t = threadIdx.x; t < n; t += blockDim.x) s = fma(Q[(size_t)i * n + t], w[t], s);
        s = block_sum(s);
        if (threadIdx.x == 0) h[i] = s;
        __syncthreads();
    }
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        double v = w[t];
        for (int i = 0; i < k; ++i) v -= h[i] * Q[(size_t)i * n + t];
        w[t] = v;
    }
}
// beta = ||w||, qnext = w / beta (qnext may alias w)
__global__ void norm_scale_kernel(const double *__restrict__ w, int n, double *__restrict__ beta_out, double *__restrict__ qnext) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(w[i], w[i], s);
    const double beta = sqrt(block_sum(s));
    for (int i = threadIdx.x; i < n; i += blockDim.x) qnext[i] = beta > 0.0 ? w[i] / beta : 0.0;
    if (threadIdx.x == 0 && beta_out) beta_out[0] = beta;
}
// u = sum_i y[i] Q[i], normalised
__global__ void combine_kernel(const double *__restrict__ Q, const double *__restrict__ y, int k, int n, double *__restrict__ u) {
    double s = 0.0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        double v = 0.0;
        for (int i = 0; i < k; ++i) v = fma(y[i], Q[(size_t)i * n + t], v);
        u[t] = v;
        s = fma(v, v, s);
    }
    const double nrm = sqrt(block_sum(s));
    for (int t = threadIdx.x; t < n; t += blockDim.x) u[t] = nrm > 0.0 ? u[t] / nrm : 0.0;
}
__global__ void seed_kernel(double *__restrict__ q, int n) {   // a fixed, non-special start vector
    for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] = 1.0 + 0.5 * sin(0.7 * i + 0.3) + 0.25 * cos(2.3 * i);
}

struct PcaWork {
    double *F = nullptr, *X = nullptr, *part = nullptr, *test = nullptr, *b = nullptr, *c = nullptr, *u = nullptr, *w = nullptr;
    double *Q = nullptr, *scal = nullptr, *yd = nullptr;
    int64_t *idx = nullptr;
    int nseg = 1, zseg = 1;
};

int segments_for(int nz, int64_t n, int sm_count) {
    // enough (column block, segment) blocks to fill the device, at least 32 planes per segment
    const int64_t colblocks = (n + PT - 1) / PT;
    int nseg = (int)std::max<int64_t>(1, std::min<int64_t>(nz / 32 > 0 ? nz / 32 : 1, ((int64_t)sm_count * 8 + colblocks - 1) / colblocks));
    return std::max(1, std::min(nseg, 64));
}

// c = M^T v (M: [nz][n])
int gemv_t(ogn_ctx *ctx, const PcaWork &wk, const double *M, int nz, int64_t n, const double *v, double *c) {
    const int nseg = segments_for(nz, n, ctx->sm_count), zseg = ogn_div_up(nz, nseg);
    dim3 grid(ogn_div_up(n, PT), nseg);
    gemv_t_partial_kernel<<<grid, PT, 0, ctx->stream>>>(M, nz, n, v, zseg, wk.part);
    OGN_LAUNCH_CHECK("gemv_t_partial_kernel");
    seg_finish_kernel<<<ogn_div_up(n, PT), PT, 0, ctx->stream>>>(wk.part, nseg, n, 1.0, c);
    OGN_LAUNCH_CHECK("seg_finish_kernel");
    return OGN_OK;
}

// first left singular vector of X [nz][npx] into wk.u (unit norm)
int top_left_vector(ogn_ctx *ctx, const PcaWork &wk, int nz, int64_t npx, int *matvecs) {
    static const int m_cfg = getenv("OGN_PCA_KRYLOV") ? std::max(2, std::min(LANCZOS_M, atoi(getenv("OGN_PCA_KRYLOV")))) : LANCZOS_M_DEFAULT;
    const int m = std::min(m_cfg, nz);
    seed_kernel<<<1, 1024, 0, ctx->stream>>>(wk.w, nz);
    OGN_LAUNCH_CHECK("seed_kernel");
    norm_scale_kernel<<<1, 1024, 0, ctx->stream>>>(wk.w, nz, nullptr, wk.Q);
    OGN_LAUNCH_CHECK("norm_scale_kernel");
    std::vector<double> alpha(m), beta(m), y, host(2 * m);
    double *d_alpha = wk.scal, *d_beta = wk.scal + m;
    for (int cycle = 0; cycle < LANCZOS_CYCLES; ++cycle) {
        for (int j = 0; j < m; ++j) {
            const double *qj = wk.Q + (size_t)j * nz;
            OGN_TRY(gemv_t(ctx, wk, wk.X, nz, npx, qj, wk.c));                         // c = X^T q_j
            gemv_n_kernel<<<nz, PT, 0, ctx->stream>>>(wk.X, npx, wk.c, wk.w);          // w = X c
            OGN_LAUNCH_CHECK("gemv_n_kernel");
            ++*matvecs;
            lanczos_orth_kernel<<<1, 1024, 0, ctx->stream>>>(wk.w, qj, j ? qj - nz : nullptr, j ? d_beta + j - 1 : nullptr, nz,
                                                            d_alpha + j);
            OGN_LAUNCH_CHECK("lanczos_orth_kernel");
            for (int rep = 0; rep < 2; ++rep) {                                        // "twice is enough"
                reorth_kernel<<<1, 1024, 0, ctx->stream>>>(wk.Q, j + 1, nz, wk.w);
                OGN_LAUNCH_CHECK("reorth_kernel");
            }
            // q_{j+1} (the slot after the last one receives the residual direction; it is not used)
            norm_scale_kernel<<<1, 1024, 0, ctx->stream>>>(wk.w, nz, d_beta + j, wk.Q + (size_t)(j + 1) * nz);
            OGN_LAUNCH_CHECK("norm_scale_kernel");
        }
        OGN_CUDA(cudaMemcpyAsync(host.data(), wk.scal, 2 * m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < m; ++j) { alpha[j] = host[j]; beta[j] = host[m + j]; }
        // an invariant subspace was found when some beta vanishes: only the leading block is a valid tridiagonal
        int k = m;
        double scale = 0.0;
        for (int j = 0; j < m; ++j) scale = std::max(scale, fabs(alpha[j]));
        for (int j = 0; j < m - 1; ++j)
            if (!(beta[j] > 1e-14 * scale)) { k = j + 1; break; }
        double theta = 0.0;
        tridiag_top(alpha, beta, k, &theta, &y);
        OGN_CUDA(cudaMemcpyAsync(wk.yd, y.data(), k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        combine_kernel<<<1, 1024, 0, ctx->stream>>>(wk.Q, wk.yd, k, nz, wk.u);
        OGN_LAUNCH_CHECK("combine_kernel");
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));     // y is a host vector reused by the next cycle
        const double resid = k < m ? 0.0 : fabs(beta[k - 1] * y[k - 1]);
        if (!(theta > 0.0) || resid <= 1e-13 * theta) return OGN_OK;
        OGN_CUDA(cudaMemcpyAsync(wk.Q, wk.u, nz * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));   // restart from the Ritz vector
    }
    return OGN_OK;   // not converged to 1e-13 after all cycles: the best Ritz vector is used (as a capped ARPACK run would)
}

}  // namespace

// Compute_GreedyPCA (lib_origin.py:858-954) on the spaxels `cols` (NULL: all) of a [nz][ld] cube.
//   cube      float32 / float64, host or device; read only
//   test0     NULL (the O2 test of the block is computed, as Compute_PCA_threshold does, :836) or [n] doubles
//   faint     [nz][ld] cube of `out_dtype` (host or device): only the columns `cols` are written
//   map_o2    [n] doubles out: iterations each spaxel spent above the threshold (:887)
//   info      {nstop, iterations, matvecs}
extern "C" int ogn_greedy_pca(ogn_ctx *ctx, const void *cube, int dtype, int nz, int64_t ld, const int64_t *cols, int64_t n,
                              const double *test0, double thres, double noise_population, int itermax, void *faint,
                              int out_dtype, double *map_o2, int *info) {
    if (!ctx) return OGN_ERR_ARG;
    if (!cube || !faint || nz <= 0 || ld <= 0 || n <= 0 || n > ld || !(noise_population > 0) || itermax < 0)
        return ogn_fail(ctx, OGN_ERR_ARG, "ogn_greedy_pca: bad arguments");
    if ((dtype != OGN_F32 && dtype != OGN_F64) || (out_dtype != OGN_F32 && out_dtype != OGN_F64))
        return ogn_fail(ctx, OGN_ERR_ARG, "ogn_greedy_pca: unknown dtype");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t es = dtype == OGN_F64 ? 8 : 4, eo = out_dtype == OGN_F64 ? 8 : 4;
    const void *d_cube = nullptr;
    OGN_TRY(ogn_input(ctx, "pca_cube", cube, (size_t)nz * ld * es, &d_cube));
    const void *d_cols = nullptr;
    if (cols) OGN_TRY(ogn_input(ctx, "pca_cols", cols, (size_t)n * 8, &d_cols));
    PcaWork wk;
    wk.nseg = segments_for(nz, n, ctx->sm_count);
    wk.zseg = ogn_div_up(nz, wk.nseg);
    OGN_TRY(ogn_scratch_t(ctx, "pca_F", (size_t)nz * n, &wk.F));
    OGN_TRY(ogn_scratch_t(ctx, "pca_X", (size_t)nz * n, &wk.X));
    OGN_TRY(ogn_scratch_t(ctx, "pca_part", (size_t)64 * n, &wk.part));
    OGN_TRY(ogn_scratch_t(ctx, "pca_test", (size_t)n, &wk.test));
    OGN_TRY(ogn_scratch_t(ctx, "pca_b", (size_t)nz, &wk.b));
    OGN_TRY(ogn_scratch_t(ctx, "pca_c", (size_t)n, &wk.c));
    OGN_TRY(ogn_scratch_t(ctx, "pca_u", (size_t)nz, &wk.u));
    OGN_TRY(ogn_scratch_t(ctx, "pca_w", (size_t)nz, &wk.w));
    OGN_TRY(ogn_scratch_t(ctx, "pca_Q", (size_t)(LANCZOS_M + 1) * nz, &wk.Q));
    OGN_TRY(ogn_scratch_t(ctx, "pca_scal", (size_t)2 * LANCZOS_M + 8, &wk.scal));
    OGN_TRY(ogn_scratch_t(ctx, "pca_y", (size_t)LANCZOS_M + 1, &wk.yd));
    OGN_TRY(ogn_scratch_t(ctx, "pca_idx", (size_t)n, &wk.idx));

    ogn_timer t_(ctx, "pca_greedy");
    const dim3 ggrid(ogn_div_up(n, PT), nz);
    if (dtype == OGN_F64)
        gather_area_kernel<double><<<ggrid, PT, 0, ctx->stream>>>((const double *)d_cube, ld, (const int64_t *)d_cols, n, wk.F);
    else
        gather_area_kernel<float><<<ggrid, PT, 0, ctx->stream>>>((const float *)d_cube, ld, (const int64_t *)d_cols, n, wk.F);
    OGN_LAUNCH_CHECK("gather_area_kernel");

    std::vector<double> test(n), mapo2(n, 0.0);
    auto fetch_test = [&]() -> int {
        OGN_CUDA(cudaMemcpyAsync(test.data(), wk.test, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
        return OGN_OK;
    };
    if (test0) {
        std::copy(test0, test0 + n, test.begin());     // host vector, as the reference's testO2 list
    } else {
        const dim3 grid(ogn_div_up(n, PT), wk.nseg);
        colsumsq_partial_kernel<<<grid, PT, 0, ctx->stream>>>(wk.F, nz, n, wk.zseg, wk.part);
        OGN_LAUNCH_CHECK("colsumsq_partial_kernel");
        seg_finish_kernel<<<ogn_div_up(n, PT), PT, 0, ctx->stream>>>(wk.part, wk.nseg, n, 1.0 / nz, wk.test);
        OGN_LAUNCH_CHECK("seg_finish_kernel");
        OGN_TRY(fetch_test());
    }

    int nstop = 0, nbiter = 0, matvecs = 0;
    std::vector<int64_t> pypx, nind, pick;
    std::vector<double> tv;
    auto nuisance = [&]() {
        pypx.clear();
        for (int64_t s = 0; s < n; ++s)
            if (test[s] > thres) pypx.push_back(s);           // :873 / :943
    };
    nuisance();
    while (!pypx.empty()) {
        ++nbiter;
        for (int64_t s : pypx) mapo2[s] += 1.0;                // :887
        if (nbiter > itermax) { ++nstop; break; }             // :888-891
        // background: the quietest spaxels below the threshold (:894-903).  Positions in the compressed vector
        // test[test > 0] are used as column indices, exactly like the reference.
        tv.clear();
        for (int64_t s = 0; s < n; ++s)
            if (test[s] > 0) tv.push_back(test[s]);
        nind.clear();
        for (int64_t i = 0; i < (int64_t)tv.size(); ++i)
            if (tv[i] <= thres) nind.push_back(i);
        std::stable_sort(nind.begin(), nind.end(), [&](int64_t a, int64_t b) { return tv[a] < tv[b]; });
        const int64_t nb = std::min<int64_t>((int64_t)nind.size(), 1 + (int64_t)((double)nind.size() / noise_population));
        pick.assign(nind.begin(), nind.begin() + nb);
        if (nb > 0) OGN_CUDA(cudaMemcpyAsync(wk.idx, pick.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        mean_cols_kernel<<<nz, PT, 0, ctx->stream>>>(wk.F, n, wk.idx, (int)nb, wk.b);
        OGN_LAUNCH_CHECK("mean_cols_kernel");
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));          // `pick` is reused below
        // nuisance block, orthogonalised to the background signature (:906-909)
        const int64_t npx = (int64_t)pypx.size();
        OGN_CUDA(cudaMemcpyAsync(wk.idx, pypx.data(), (size_t)npx * 8, cudaMemcpyHostToDevice, ctx->stream));
        gather_area_kernel<double><<<dim3(ogn_div_up(npx, PT), nz), PT, 0, ctx->stream>>>(wk.F, n, wk.idx, npx, wk.X);
        OGN_LAUNCH_CHECK("gather_area_kernel");
        OGN_TRY(gemv_t(ctx, wk, wk.X, nz, npx, wk.b, wk.c));                                         // c = b^T X
        dot_kernel<<<1, 1024, 0, ctx->stream>>>(wk.b, wk.b, nz, wk.scal + 2 * LANCZOS_M);            // sum b^2
        OGN_LAUNCH_CHECK("dot_kernel");
        rank1_scale_kernel<<<dim3(ogn_div_up(npx, PT), nz), PT, 0, ctx->stream>>>(wk.X, npx, wk.b, wk.c, wk.scal + 2 * LANCZOS_M);
        OGN_LAUNCH_CHECK("rank1_scale_kernel");
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));          // `pypx` is rebuilt below
        if (npx == 1) break;                                   // :912-913
        OGN_TRY(top_left_vector(ctx, wk, nz, npx, &matvecs));  // U[:, 0] of svds(x_red, k=1), :924
        // faint -= U U^T faint (:927) and the new test (:930) in one pass
        OGN_TRY(gemv_t(ctx, wk, wk.F, nz, n, wk.u, wk.c));
        deflate_partial_kernel<<<dim3(ogn_div_up(n, PT), wk.nseg), PT, 0, ctx->stream>>>(wk.F, nz, n, wk.u, wk.c, wk.zseg, wk.part);
        OGN_LAUNCH_CHECK("deflate_partial_kernel");
        seg_finish_kernel<<<ogn_div_up(n, PT), PT, 0, ctx->stream>>>(wk.part, wk.nseg, n, 1.0 / nz, wk.test);
        OGN_LAUNCH_CHECK("seg_finish_kernel");
        OGN_TRY(fetch_test());
        nuisance();
    }

    void *d_faint = nullptr;
    const bool faint_dev = ogn_is_device_ptr(faint);
    if (faint_dev) d_faint = faint;
    else {
        // host output: stage the whole cube (the columns outside the area keep what the caller put there)
        const void *staged = nullptr;
        OGN_TRY(ogn_input(ctx, "pca_faint_out", faint, (size_t)nz * ld * eo, &staged));
        d_faint = const_cast<void *>(staged);
    }
    if (out_dtype == OGN_F64)
        scatter_area_kernel<double><<<ggrid, PT, 0, ctx->stream>>>(wk.F, n, (const int64_t *)d_cols, ld, (double *)d_faint);
    else
        scatter_area_kernel<float><<<ggrid, PT, 0, ctx->stream>>>(wk.F, n, (const int64_t *)d_cols, ld, (float *)d_faint);
    OGN_LAUNCH_CHECK("scatter_area_kernel");
    if (!faint_dev) OGN_TRY(ogn_output_commit(ctx, faint, d_faint, (size_t)nz * ld * eo));
    if (map_o2) std::copy(mapo2.begin(), mapo2.end(), map_o2);
    if (info) { info[0] = nstop; info[1] = nbiter; info[2] = matvecs; }
    return ogn_finish_call(ctx);
}
