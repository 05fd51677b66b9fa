// step01: DCT continuum fit and standardisation
// (dct_residual lib_origin.py:150-240, DCTMAT :127-146, Preprocessing.run steps.py:431-465).
//
// Everything that feeds the subtraction raw - cont is FP64: the continuum can be
// 10^3 times the noise, so an FP32 fit would leave residuals above the 1e-5
// parity bound.  B200 has a full-rate FP64 pipe, and the whole
// fit is ~100 DFMA per voxel.
//
//   K5a dct_fit_kernel    one thread per spaxel walks lambda once: weighted Gram
//                         matrix D^T W D (upper triangle in registers), D^T W s
//                         and D^T s, "any masked voxel" flag; then an in-register
//                         Cholesky solve -> order+1 coefficients per spaxel
//   K5b dct_synth_kernel  cont = D0 coef; optionally data = raw - cont (f32) and
//                         per-wavelength partial sums / counts of unmasked data
//   K5c standardise_kernel (data - mean)/sqrt(var), cont/sqrt(var) and the four
//                         per-spaxel reductions
#include <math.h>

#include <vector>

#include "ogn_common.cuh"

template <typename T>
__device__ __forceinline__ double ld_as_f64(const T *p, size_t i) { return (double)p[i]; }

// coef layout: [M][S] (spaxel fastest) so that warps read/write it coalesced.
template <int M, typename T>
__global__ void __launch_bounds__(128)
dct_fit_kernel(const T *__restrict__ raw, const T *__restrict__ var, const uint8_t *__restrict__ mask,
               const double *__restrict__ d0,  // [nz][M]
               int nz, size_t S, int approx, double *__restrict__ coef) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    constexpr int NG = M * (M + 1) / 2;
    double G[NG], bw[M], b0[M];
#pragma unroll
    for (int i = 0; i < NG; ++i) G[i] = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { bw[i] = 0.0; b0[i] = 0.0; }
    bool any_masked = false;
    for (int z = 0; z < nz; ++z) {
        const size_t o = (size_t)z * S + s;
        const double v = ld_as_f64(raw, o);
        double d[M];
#pragma unroll
        for (int i = 0; i < M; ++i) d[i] = __ldg(d0 + (size_t)z * M + i);
#pragma unroll
        for (int i = 0; i < M; ++i) b0[i] = fma(d[i], v, b0[i]);
        if (!approx) {
            const double w = 1.0 / ld_as_f64(var, o);
            any_masked |= mask[o] != 0;
            const double vw = v * w;
            int g = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const double dw = d[i] * w;
                bw[i] = fma(d[i], vw, bw[i]);
#pragma unroll
                for (int j = i; j < M; ++j) { G[g] = fma(dw, d[j], G[g]); ++g; }
            }
        }
    }
    double c[M];
    if (approx || any_masked) {
        // unweighted projection, D0 has orthonormal columns (lib_origin.py:192, :237)
#pragma unroll
        for (int i = 0; i < M; ++i) c[i] = b0[i];
    } else {
        // Cholesky G = L L^T on the packed upper triangle, then two triangular solves
        // (the reference inverts G, lib_origin.py:233-235; same solution to round-off)
        auto at = [](int i, int j) { return i * M - i * (i - 1) / 2 + (j - i); };  // i <= j
#pragma unroll
        for (int i = 0; i < M; ++i) {
#pragma unroll
            for (int j = i; j < M; ++j) {
                double sum = G[at(i, j)];
#pragma unroll
                for (int k = 0; k < i; ++k) sum -= G[at(k, i)] * G[at(k, j)];
                G[at(i, j)] = (j == i) ? sqrt(sum) : sum / G[at(i, i)];
            }
        }
#pragma unroll
        for (int i = 0; i < M; ++i) {  // L y = bw   (L[i][k] = U[k][i])
            double sum = bw[i];
#pragma unroll
            for (int k = 0; k < i; ++k) sum -= G[at(k, i)] * c[k];
            c[i] = sum / G[at(i, i)];
        }
#pragma unroll
        for (int i = M - 1; i >= 0; --i) {  // L^T x = y
            double sum = c[i];
#pragma unroll
            for (int k = i + 1; k < M; ++k) sum -= G[at(i, k)] * c[k];
            c[i] = sum / G[at(i, i)];
        }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) coef[(size_t)i * S + s] = c[i];
}

// Any order (arrays in local memory): the slow, general twin of dct_fit_kernel.
constexpr int DCT_MAXM = 33;
template <typename T>
__global__ void dct_fit_generic_kernel(const T *__restrict__ raw, const T *__restrict__ var,
                                       const uint8_t *__restrict__ mask, const double *__restrict__ d0, int M,
                                       int nz, size_t S, int approx, double *__restrict__ coef,
                                       double *__restrict__ gram_ws /* [S][M*M] */) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double *G = gram_ws + s * (size_t)M * M;
    double bw[DCT_MAXM], b0[DCT_MAXM], c[DCT_MAXM];
    for (int i = 0; i < M * M; ++i) G[i] = 0.0;
    for (int i = 0; i < M; ++i) { bw[i] = 0.0; b0[i] = 0.0; }
    bool any_masked = false;
    for (int z = 0; z < nz; ++z) {
        const size_t o = (size_t)z * S + s;
        const double v = ld_as_f64(raw, o);
        const double *d = d0 + (size_t)z * M;
        for (int i = 0; i < M; ++i) b0[i] = fma(d[i], v, b0[i]);
        if (!approx) {
            const double w = 1.0 / ld_as_f64(var, o);
            any_masked |= mask[o] != 0;
            for (int i = 0; i < M; ++i) {
                bw[i] = fma(d[i] * w, v, bw[i]);
                for (int j = i; j < M; ++j) G[i * M + j] = fma(d[i] * w, d[j], G[i * M + j]);
            }
        }
    }
    if (approx || any_masked) {
        for (int i = 0; i < M; ++i) c[i] = b0[i];
    } else {
        for (int i = 0; i < M; ++i)
            for (int j = i; j < M; ++j) {
                double sum = G[i * M + j];
                for (int k = 0; k < i; ++k) sum -= G[k * M + i] * G[k * M + j];
                G[i * M + j] = (j == i) ? sqrt(sum) : sum / G[i * M + i];
            }
        for (int i = 0; i < M; ++i) {
            double sum = bw[i];
            for (int k = 0; k < i; ++k) sum -= G[k * M + i] * c[k];
            c[i] = sum / G[i * M + i];
        }
        for (int i = M - 1; i >= 0; --i) {
            double sum = c[i];
            for (int k = i + 1; k < M; ++k) sum -= G[i * M + k] * c[k];
            c[i] = sum / G[i * M + i];
        }
    }
    for (int i = 0; i < M; ++i) coef[(size_t)i * S + s] = c[i];
}

// cont[z][s] = sum_i d0[z][i] coef[i][s].  One thread per spaxel, walking lambda.
//   cont_out (f64 or f32, may be NULL), data_out = raw - cont (f32, NaN->excluded by mask),
//   lambda_sum / lambda_cnt: per-wavelength sums over the unmasked voxels of this launch.
template <typename T, typename TO>
__global__ void __launch_bounds__(128)
dct_synth_kernel(const T *__restrict__ raw, const uint8_t *__restrict__ mask, const double *__restrict__ d0, int M,
                 int nz, size_t S, const double *__restrict__ coef, TO *__restrict__ cont_out,
                 double *__restrict__ cont64, float *__restrict__ data_out, double *__restrict__ lambda_sum,
                 double *__restrict__ lambda_cnt, int nx, int wy0, int wy1, int wx0, int wx1) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < S;
    // only spaxels inside the owned window contribute to the per-wavelength sums (multi-GPU tiles)
    const int sy = (int)(s / nx), sx = (int)(s - (size_t)sy * nx);
    const bool counted = live && sy >= wy0 && sy < wy1 && sx >= wx0 && sx < wx1;
    double c[DCT_MAXM];
    for (int i = 0; i < M; ++i) c[i] = live ? coef[(size_t)i * S + s] : 0.0;
    for (int z = 0; z < nz; ++z) {
        const double *d = d0 + (size_t)z * M;
        double cont = 0.0;
        for (int i = 0; i < M; ++i) cont = fma(d[i], c[i], cont);
        double v = 0.0, n = 0.0;
        if (live) {
            const size_t o = (size_t)z * S + s;
            if (cont_out) cont_out[o] = (TO)cont;
            if (cont64) cont64[o] = cont;
            if (data_out) {
                const double r = ld_as_f64(raw, o) - cont;
                data_out[o] = (float)r;
                if (counted && !mask[o]) { v = r; n = 1.0; }
            }
        }
        if (lambda_sum) {
            for (int o = 16; o; o >>= 1) {
                v += __shfl_xor_sync(0xffffffffu, v, o);
                n += __shfl_xor_sync(0xffffffffu, n, o);
            }
            if ((threadIdx.x & 31) == 0 && n > 0.0) {
                atomicAdd(lambda_sum + z, v);
                atomicAdd(lambda_cnt + z, n);
            }
        }
    }
}

// steps.py:439-450, :463-465, :472, :480 for one spaxel per thread.
template <typename T>
__global__ void __launch_bounds__(128)
standardise_kernel(const float *__restrict__ data, const double *__restrict__ cont, const T *__restrict__ var,
                   const uint8_t *__restrict__ mask, const double *__restrict__ mean, int nz, size_t S,
                   float *__restrict__ cube_std, float *__restrict__ cont_dct, double *__restrict__ ima_std,
                   double *__restrict__ ima_dct, double *__restrict__ cont_sumsq, double *__restrict__ o2map) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double a_std = 0, a_dct = 0, a_c2 = 0, a_o2 = 0;
    for (int z = 0; z < nz; ++z) {
        const size_t o = (size_t)z * S + s;
        const double sd = sqrt(ld_as_f64(var, o));
        double v = 0.0;
        if (!mask[o]) v = ((double)data[o] - mean[z]) / sd;
        const float cd = (float)(cont[o] / sd);
        if (cube_std) cube_std[o] = (float)v;
        if (cont_dct) cont_dct[o] = cd;
        a_std += v;
        a_o2 += v * v;
        a_dct += (double)cd;
        a_c2 += (double)cd * (double)cd;
    }
    if (ima_std) ima_std[s] = a_std / nz;
    if (o2map) o2map[s] = a_o2 / nz;
    if (ima_dct) ima_dct[s] = a_dct / nz;
    if (cont_sumsq) cont_sumsq[s] = a_c2;
}

// -------------------------------------------------------------------------------------------

static int upload_dctmat(ogn_ctx *ctx, int nz, int M, const double **d0_dev) {
    // DCTMAT, lib_origin.py:143-145
    std::vector<double> d0((size_t)nz * M);
    const double scale = sqrt(2.0 / nz), step = M_PI / nz;
    for (int z = 0; z < nz; ++z)
        for (int j = 0; j < M; ++j) {
            double v = scale * cos((z + 0.5) * step * j);
            if (j == 0) v *= 1.0 / sqrt(2.0);
            d0[(size_t)z * M + j] = v;
        }
    double *d = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dctmat", d0.size(), &d));
    OGN_CUDA(cudaMemcpyAsync(d, d0.data(), d0.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    *d0_dev = d;
    return OGN_OK;
}

template <typename T>
static int run_fit(ogn_ctx *ctx, const T *raw, const T *var, const uint8_t *mask, const double *d0, int M, int nz,
                   size_t S, int approx, double *coef) {
    const int blocks = ogn_div_up((int64_t)S, 128);
    if (M == 11) {
        dct_fit_kernel<11, T><<<blocks, 128, 0, ctx->stream>>>(raw, var, mask, d0, nz, S, approx, coef);
        OGN_LAUNCH_CHECK("dct_fit_kernel");
    } else {
        double *ws = nullptr;
        OGN_TRY(ogn_scratch_t(ctx, "dct_gram_ws", S * (size_t)M * M, &ws));
        dct_fit_generic_kernel<T><<<blocks, 128, 0, ctx->stream>>>(raw, var, mask, d0, M, nz, S, approx, coef, ws);
        OGN_LAUNCH_CHECK("dct_fit_generic_kernel");
    }
    return OGN_OK;
}

struct DctInputs {
    const void *raw = nullptr, *var = nullptr;
    const uint8_t *mask = nullptr;
};

static int stage_dct_inputs(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype, const uint8_t *mask,
                            size_t vol, int approx, DctInputs *in) {
    if (in_dtype != OGN_F32 && in_dtype != OGN_F64) return ogn_fail(ctx, OGN_ERR_ARG, "unknown dtype %d", in_dtype);
    if (!raw) return ogn_fail(ctx, OGN_ERR_ARG, "raw is NULL");
    const size_t es = in_dtype == OGN_F64 ? 8 : 4;
    OGN_TRY(ogn_input(ctx, "dct_raw", raw, vol * es, &in->raw));
    if (var) OGN_TRY(ogn_input(ctx, "dct_var", var, vol * es, &in->var));
    else if (!approx) return ogn_fail(ctx, OGN_ERR_ARG, "var is required unless approx");
    if (mask) {
        const void *d = nullptr;
        OGN_TRY(ogn_input(ctx, "dct_mask", mask, vol, &d));
        in->mask = (const uint8_t *)d;
    } else if (!approx) {
        return ogn_fail(ctx, OGN_ERR_ARG, "mask is required unless approx");
    }
    return OGN_OK;
}

static int check_dct_args(ogn_ctx *ctx, int nz, int ny, int nx, int order) {
    if (!ctx) return OGN_ERR_ARG;
    if (nz <= 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "cube shape (%d,%d,%d) is empty", nz, ny, nx);
    if (order < 0 || order + 1 > DCT_MAXM || order + 1 > nz)
        return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "dct order %d not in [0, min(%d, nz-1)]", order, DCT_MAXM - 1);
    return OGN_OK;
}

extern "C" int ogn_dct_residual(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype, const uint8_t *mask,
                                int nz, int ny, int nx, int order, int approx, void *cont, int out_dtype) {
    OGN_TRY(check_dct_args(ctx, nz, ny, nx, order));
    if (!cont) return ogn_fail(ctx, OGN_ERR_ARG, "cont is NULL");
    if (out_dtype != OGN_F32 && out_dtype != OGN_F64) return ogn_fail(ctx, OGN_ERR_ARG, "unknown dtype %d", out_dtype);
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t S = (size_t)ny * nx, vol = S * nz;
    const int M = order + 1;
    DctInputs in;
    OGN_TRY(stage_dct_inputs(ctx, raw, var, in_dtype, mask, vol, approx, &in));
    const double *d0 = nullptr;
    OGN_TRY(upload_dctmat(ctx, nz, M, &d0));
    double *coef = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dct_coef", S * M, &coef));
    void *d_cont = nullptr;
    OGN_TRY(ogn_output(ctx, "dct_cont_out", cont, vol * (out_dtype == OGN_F64 ? 8 : 4), &d_cont));
    const int blocks = ogn_div_up((int64_t)S, 128);
    if (in_dtype == OGN_F64) {
        OGN_TRY(run_fit<double>(ctx, (const double *)in.raw, (const double *)in.var, in.mask, d0, M, nz, S, approx, coef));
    } else {
        OGN_TRY(run_fit<float>(ctx, (const float *)in.raw, (const float *)in.var, in.mask, d0, M, nz, S, approx, coef));
    }
    if (out_dtype == OGN_F64)
        dct_synth_kernel<float, double><<<blocks, 128, 0, ctx->stream>>>(nullptr, nullptr, d0, M, nz, S, coef,
                                                                        (double *)d_cont, nullptr, nullptr, nullptr, nullptr,
                                                                        nx, 0, ny, 0, nx);
    else
        dct_synth_kernel<float, float><<<blocks, 128, 0, ctx->stream>>>(nullptr, nullptr, d0, M, nz, S, coef,
                                                                       (float *)d_cont, nullptr, nullptr, nullptr, nullptr,
                                                                       nx, 0, ny, 0, nx);
    OGN_LAUNCH_CHECK("dct_synth_kernel");
    OGN_TRY(ogn_output_commit(ctx, cont, d_cont, vol * (out_dtype == OGN_F64 ? 8 : 4)));
    return ogn_finish_call(ctx);
}

extern "C" int ogn_preprocess_begin(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype,
                                    const uint8_t *mask, int nz, int ny, int nx, int order, int approx,
                                    const int *owned, double *lambda_sum, double *lambda_cnt) {
    OGN_TRY(check_dct_args(ctx, nz, ny, nx, order));
    const int wy0 = owned ? owned[0] : 0, wy1 = owned ? owned[1] : ny;
    const int wx0 = owned ? owned[2] : 0, wx1 = owned ? owned[3] : nx;
    if (wy0 < 0 || wy1 > ny || wx0 < 0 || wx1 > nx || wy0 > wy1 || wx0 > wx1)
        return ogn_fail(ctx, OGN_ERR_ARG, "owned window [%d,%d)x[%d,%d) outside the %dx%d field", wy0, wy1, wx0, wx1, ny, nx);
    if (!var || !mask) return ogn_fail(ctx, OGN_ERR_ARG, "var and mask are required");
    if (!lambda_sum || !lambda_cnt) return ogn_fail(ctx, OGN_ERR_ARG, "lambda_sum / lambda_cnt are NULL");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t S = (size_t)ny * nx, vol = S * nz;
    const int M = order + 1;
    DctInputs in;
    OGN_TRY(stage_dct_inputs(ctx, raw, var, in_dtype, mask, vol, approx, &in));
    const double *d0 = nullptr;
    OGN_TRY(upload_dctmat(ctx, nz, M, &d0));
    double *coef = nullptr, *cont64 = nullptr;
    float *data = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dct_coef", S * M, &coef));
    OGN_TRY(ogn_scratch_t(ctx, "prep_cont64", vol, &cont64));
    OGN_TRY(ogn_scratch_t(ctx, "prep_data", vol, &data));
    void *d_sum = nullptr, *d_cnt = nullptr;
    OGN_TRY(ogn_output(ctx, "prep_lsum", lambda_sum, (size_t)nz * 8, &d_sum));
    OGN_TRY(ogn_output(ctx, "prep_lcnt", lambda_cnt, (size_t)nz * 8, &d_cnt));
    OGN_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)nz * 8, ctx->stream));
    OGN_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)nz * 8, ctx->stream));
    const int blocks = ogn_div_up((int64_t)S, 128);
    if (in_dtype == OGN_F64) {
        OGN_TRY(run_fit<double>(ctx, (const double *)in.raw, (const double *)in.var, in.mask, d0, M, nz, S, approx, coef));
        dct_synth_kernel<double, double><<<blocks, 128, 0, ctx->stream>>>((const double *)in.raw, in.mask, d0, M, nz, S,
                                                                         coef, nullptr, cont64, data, (double *)d_sum,
                                                                         (double *)d_cnt, nx, wy0, wy1, wx0, wx1);
    } else {
        OGN_TRY(run_fit<float>(ctx, (const float *)in.raw, (const float *)in.var, in.mask, d0, M, nz, S, approx, coef));
        dct_synth_kernel<float, double><<<blocks, 128, 0, ctx->stream>>>((const float *)in.raw, in.mask, d0, M, nz, S,
                                                                        coef, nullptr, cont64, data, (double *)d_sum,
                                                                        (double *)d_cnt, nx, wy0, wy1, wx0, wx1);
    }
    OGN_LAUNCH_CHECK("dct_synth_kernel");
    ctx->prep.active = true;
    ctx->prep.nz = nz; ctx->prep.ny = ny; ctx->prep.nx = nx; ctx->prep.in_dtype = in_dtype;
    ctx->prep.var = in.var; ctx->prep.mask = in.mask; ctx->prep.data = data; ctx->prep.cont = cont64;
    OGN_TRY(ogn_output_commit(ctx, lambda_sum, d_sum, (size_t)nz * 8));
    OGN_TRY(ogn_output_commit(ctx, lambda_cnt, d_cnt, (size_t)nz * 8));
    return ogn_finish_call(ctx);
}

extern "C" int ogn_preprocess_finish(ogn_ctx *ctx, const double *lambda_mean, float *cube_std, float *cont_dct,
                                     double *ima_std, double *ima_dct, double *cont_sumsq, double *o2map) {
    if (!ctx) return OGN_ERR_ARG;
    if (!ctx->prep.active) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_preprocess_finish without ogn_preprocess_begin");
    if (!lambda_mean) return ogn_fail(ctx, OGN_ERR_ARG, "lambda_mean is NULL");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const ogn_prep_state &st = ctx->prep;
    const int nz = st.nz;
    const size_t S = (size_t)st.ny * st.nx, vol = S * nz;
    const void *d_mean = nullptr;
    OGN_TRY(ogn_input(ctx, "prep_mean", lambda_mean, (size_t)nz * 8, &d_mean));
    void *d_std = nullptr, *d_cd = nullptr, *d_is = nullptr, *d_id = nullptr, *d_c2 = nullptr, *d_o2 = nullptr;
    if (cube_std) OGN_TRY(ogn_output(ctx, "prep_cube_std", cube_std, vol * 4, &d_std));
    if (cont_dct) OGN_TRY(ogn_output(ctx, "prep_cont_dct", cont_dct, vol * 4, &d_cd));
    if (ima_std) OGN_TRY(ogn_output(ctx, "prep_ima_std", ima_std, S * 8, &d_is));
    if (ima_dct) OGN_TRY(ogn_output(ctx, "prep_ima_dct", ima_dct, S * 8, &d_id));
    if (cont_sumsq) OGN_TRY(ogn_output(ctx, "prep_c2", cont_sumsq, S * 8, &d_c2));
    if (o2map) OGN_TRY(ogn_output(ctx, "prep_o2", o2map, S * 8, &d_o2));
    const int blocks = ogn_div_up((int64_t)S, 128);
    if (st.in_dtype == OGN_F64)
        standardise_kernel<double><<<blocks, 128, 0, ctx->stream>>>(st.data, st.cont, (const double *)st.var, st.mask,
                                                                   (const double *)d_mean, nz, S, (float *)d_std,
                                                                   (float *)d_cd, (double *)d_is, (double *)d_id,
                                                                   (double *)d_c2, (double *)d_o2);
    else
        standardise_kernel<float><<<blocks, 128, 0, ctx->stream>>>(st.data, st.cont, (const float *)st.var, st.mask,
                                                                  (const double *)d_mean, nz, S, (float *)d_std,
                                                                  (float *)d_cd, (double *)d_is, (double *)d_id,
                                                                  (double *)d_c2, (double *)d_o2);
    OGN_LAUNCH_CHECK("standardise_kernel");
    OGN_TRY(ogn_output_commit(ctx, cube_std, d_std, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, cont_dct, d_cd, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, ima_std, d_is, S * 8));
    OGN_TRY(ogn_output_commit(ctx, ima_dct, d_id, S * 8));
    OGN_TRY(ogn_output_commit(ctx, cont_sumsq, d_c2, S * 8));
    OGN_TRY(ogn_output_commit(ctx, o2map, d_o2, S * 8));
    ctx->prep.active = false;
    return ogn_finish_call(ctx);
}
