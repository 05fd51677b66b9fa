// Fused step-level entry points: the whole array part of one ORIGIN step in one call, so
// that intermediates stay on the device and every product crosses PCIe at most once.
#include "ogn_common.cuh"

// ComputeTGLR.run (steps.py:768-802): Correlation_GLR_test + masking + maxmap/minmap +
// compute_local_max, with correl / correl_min / profile never leaving the device in between.
extern "C" int ogn_step05(ogn_ctx *ctx, const void *cube, int cube_dtype, int nz, int ny, int nx, int nfields,
                          const double *const *fsf, int psize, const double *const *weights, const double *taps,
                          const int *tap_offsets, int nprof, const uint8_t *mask, int sz, int sy, int sx,
                          float *correl, float *correl_min, uint8_t *profile, float *maxmap, float *minmap,
                          float *dense_max, float *dense_min, int64_t *max_index, float *max_value,
                          int64_t *min_index, float *min_value, int64_t capacity, int64_t *counts) {
    if (!ctx) return OGN_ERR_ARG;
    if (nz <= 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "cube shape (%d,%d,%d) is empty", nz, ny, nx);
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t vol = (size_t)nz * ny * nx, img = (size_t)ny * nx;
    void *d_correl = nullptr, *d_cmin = nullptr, *d_prof = nullptr, *d_maxmap = nullptr, *d_minmap = nullptr;
    // correl and correl_min are needed on the device even when the caller does not want them back
    OGN_TRY(ogn_output(ctx, "s5_correl", correl, vol * 4, &d_correl));
    OGN_TRY(ogn_output(ctx, "s5_correl_min", correl_min, vol * 4, &d_cmin));
    if (profile) OGN_TRY(ogn_output(ctx, "s5_profile", profile, vol, &d_prof));
    if (maxmap) OGN_TRY(ogn_output(ctx, "s5_maxmap", maxmap, img * 4, &d_maxmap));
    if (minmap) OGN_TRY(ogn_output(ctx, "s5_minmap", minmap, img * 4, &d_minmap));
    const void *d_mask = nullptr;
    if (mask) OGN_TRY(ogn_input(ctx, "s5_mask", mask, vol, &d_mask));

    OGN_TRY(ogn_tglr(ctx, cube, cube_dtype, nz, ny, nx, nfields, fsf, psize, weights, taps, tap_offsets, nprof,
                     (const uint8_t *)d_mask, (float *)d_correl, (float *)d_cmin, (uint8_t *)d_prof,
                     (float *)d_maxmap, (float *)d_minmap));
    OGN_TRY(ogn_output_commit(ctx, correl, d_correl, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, correl_min, d_cmin, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, profile, d_prof, vol));
    OGN_TRY(ogn_output_commit(ctx, maxmap, d_maxmap, img * 4));
    OGN_TRY(ogn_output_commit(ctx, minmap, d_minmap, img * 4));
    int rc = ogn_local_extrema(ctx, (const float *)d_correl, (const float *)d_cmin, (const uint8_t *)d_mask, nz, ny,
                               nx, sz, sy, sx, dense_max, dense_min, max_index, max_value, min_index, min_value,
                               capacity, counts);
    if (rc != OGN_OK && rc != OGN_ERR_OVERFLOW) return rc;
    OGN_TRY(ogn_finish_call(ctx));
    return rc;
}
