// Fused step-level entry points: the whole array part of one ORIGIN step in one call, so
// that intermediates stay on the device and every product crosses PCIe at most once.
#include <stdlib.h>

#include <algorithm>

#include "ogn_common.cuh"

namespace {

struct Step05Args {
    const void *cube; int cube_dtype, nz, ny, nx, nfields;
    const double *const *fsf; int psize;
    const double *const *weights; const double *taps; const int *tap_offsets; int nprof;
    const uint8_t *mask; int sz, sy, sx;
    float *correl, *correl_min; uint8_t *profile; float *maxmap, *minmap, *dense_max, *dense_min;
    int64_t *max_index; float *max_value; int64_t *min_index; float *min_value; int64_t capacity; int64_t *counts;
    int mask_bits = 0;   // mask is bit-packed (numpy.packbits order, [nz][ny][nx] flattened), host memory
};

int get_event(ogn_ctx *ctx, size_t i, cudaEvent_t *ev) {
    while (ctx->events.size() <= i) {
        cudaEvent_t e;
        OGN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    *ev = ctx->events[i];
    return OGN_OK;
}

// out[i] = (bits[i >> 3] >> (7 - (i & 7))) & 1 for the voxels of `rows` image rows of every plane
// (numpy.packbits order); one thread per output byte quad
__global__ void unpack_mask_rows_kernel(const uint8_t *__restrict__ bits, uint8_t *__restrict__ out, int nz, size_t img,
                                        size_t off, size_t count) {
    const size_t per_plane = count / 4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_plane * nz) return;
    const size_t z = i / per_plane, q = i - z * per_plane;
    const size_t v = z * img + off + 4 * q;            // first voxel of the quad (a multiple of 4)
    const unsigned byte = bits[v >> 3];
    const unsigned sh = 4 - (unsigned)(v & 4);         // high nibble first
    const unsigned nib = (byte >> sh) & 15u;
    uchar4 o;
    o.x = (nib >> 3) & 1u; o.y = (nib >> 2) & 1u; o.z = (nib >> 1) & 1u; o.w = nib & 1u;
    *reinterpret_cast<uchar4 *>(out + v) = o;
}

// Host cube in, products to wherever the caller wants each of them (host: copied back slab by slab; device:
// written in place by the kernels, e.g. the lazily fetched correl_min / profile of the step mirror): the field
// is cut into slabs of image rows; the upload of slab s+1, the K1/K2 pass on slab s and the download of the
// products of slab s-1 overlap on three streams.
int step05_streamed(ogn_ctx *ctx, const Step05Args &a, const ogn_tglr_setup_t &st, ogn_window owned, ogn_place place,
                    ogn_window w) {
    const int nz = a.nz, ny = a.ny, nx = a.nx;
    const size_t vol = (size_t)nz * ny * nx, img = (size_t)ny * nx, plane_b = img * 4;
    // Rows per slab.  The call is PCIe-bound (measured 57 GB/s one way, 46 GB/s each way when both directions
    // are busy), so what matters is the pipeline ramp — the first upload and the last download — which
    // shrinks with the slab: 82 / 76 / 73.5 ms at 64 / 32 / 16 rows for 3681x320x320.  16 still covers the
    // FSF halo (the P/2 rows below a slab arrive with the next one); K1 runs half-empty 32-row warp patches
    // on such slabs, which stays hidden behind the copies.
    static const int SLAB = getenv("OGN_SLAB") ? std::max(16, atoi(getenv("OGN_SLAB"))) : 16;
    const int slab = std::max(SLAB, a.psize / 2 + 1);
    const int nslab = ogn_div_up(ny, slab);
    if (!ctx->h2d_stream) OGN_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) OGN_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    float *d_cube = nullptr, *d_correl = nullptr, *d_cmin = nullptr, *d_maxmap = nullptr, *d_minmap = nullptr;
    uint8_t *d_mask = nullptr, *d_prof = nullptr, *d_bits = nullptr;
    const bool correl_dev = a.correl && ogn_is_device_ptr(a.correl), cmin_dev = a.correl_min && ogn_is_device_ptr(a.correl_min);
    const bool prof_dev = a.profile && ogn_is_device_ptr(a.profile);
    OGN_TRY(ogn_scratch_t(ctx, "cube", vol, &d_cube));
    if (correl_dev) d_correl = a.correl; else OGN_TRY(ogn_scratch_t(ctx, "s5_correl", vol, &d_correl));
    if (cmin_dev) d_cmin = a.correl_min; else OGN_TRY(ogn_scratch_t(ctx, "s5_correl_min", vol, &d_cmin));
    if (prof_dev) d_prof = a.profile; else if (a.profile) OGN_TRY(ogn_scratch_t(ctx, "s5_profile", vol, &d_prof));
    if (a.mask) OGN_TRY(ogn_scratch_t(ctx, "s5_mask", vol, &d_mask));
    if (a.mask && a.mask_bits) OGN_TRY(ogn_scratch_t(ctx, "s5_mask_bits", (vol + 7) / 8 + 16, &d_bits));
    if (a.maxmap) OGN_TRY(ogn_scratch_t(ctx, "s5_maxmap", img, &d_maxmap));
    if (a.minmap) OGN_TRY(ogn_scratch_t(ctx, "s5_minmap", img, &d_minmap));
    OGN_TRY(ogn_tglr_init_maps(ctx, ctx->stream, d_maxmap, d_minmap, img));
    cudaEvent_t ev;
    // the side streams start after everything already queued on the main stream (setup, scratch reuse)
    OGN_TRY(get_event(ctx, 0, &ev));
    OGN_CUDA(cudaEventRecord(ev, ctx->stream));
    OGN_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, ev, 0));
    OGN_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ev, 0));
    const float *h_cube = static_cast<const float *>(a.cube);
    for (int s = 0; s < nslab; ++s) {
        const int y0 = s * slab, rows = std::min(slab, ny - y0);
        const size_t off = (size_t)y0 * nx;
        OGN_CUDA(cudaMemcpy2DAsync(d_cube + off, plane_b, h_cube + off, plane_b, (size_t)rows * nx * 4, nz,
                                   cudaMemcpyHostToDevice, ctx->h2d_stream));
        if (a.mask && a.mask_bits) {
            // nx % 8 == 0 (checked by the caller): a slab of rows is a whole number of packed bytes per plane
            OGN_CUDA(cudaMemcpy2DAsync(d_bits + off / 8, img / 8, a.mask + off / 8, img / 8, (size_t)rows * nx / 8, nz,
                                       cudaMemcpyHostToDevice, ctx->h2d_stream));
            const size_t count = (size_t)rows * nx;
            unpack_mask_rows_kernel<<<ogn_div_up((int64_t)(count / 4) * nz, 256), 256, 0, ctx->h2d_stream>>>(d_bits, d_mask, nz, img,
                                                                                                          off, count);
            OGN_LAUNCH_CHECK("unpack_mask_rows_kernel");
        } else if (a.mask) {
            OGN_CUDA(cudaMemcpy2DAsync(d_mask + off, img, a.mask + off, img, (size_t)rows * nx, nz,
                                       cudaMemcpyHostToDevice, ctx->h2d_stream));
        }
        OGN_TRY(get_event(ctx, 1 + s, &ev));
        OGN_CUDA(cudaEventRecord(ev, ctx->h2d_stream));
    }
    for (int s = 0; s < nslab; ++s) {
        const int y0 = s * slab, rows = std::min(slab, ny - y0);
        // slab s needs input rows up to y0 + rows + P/2: they arrive with slab s+1 (slab > P/2)
        OGN_TRY(get_event(ctx, 1 + std::min(s + 1, nslab - 1), &ev));
        OGN_CUDA(cudaStreamWaitEvent(ctx->stream, ev, 0));
        // the rows of this slab inside the computed window (a tile computes its owned window grown by the extremum
        // radius; a whole field everything)
        const int cy0 = std::max(y0, w.y0), cy1 = std::min(y0 + rows, w.y1);
        if (cy0 >= cy1) continue;
        OGN_TRY(ogn_tglr_window(ctx, ctx->stream, st, d_cube, d_mask, ogn_window{cy0, cy1, w.x0, w.x1}, d_correl,
                                d_cmin, d_prof, d_maxmap, d_minmap));
        OGN_TRY(get_event(ctx, 1 + nslab + s, &ev));
        OGN_CUDA(cudaEventRecord(ev, ctx->stream));
        OGN_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ev, 0));
        const size_t off = (size_t)cy0 * nx;
        const int crows = cy1 - cy0;
        if (a.correl && !correl_dev)
            OGN_CUDA(cudaMemcpy2DAsync(a.correl + off, plane_b, d_correl + off, plane_b, (size_t)crows * nx * 4, nz,
                                       cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (a.correl_min && !cmin_dev)
            OGN_CUDA(cudaMemcpy2DAsync(a.correl_min + off, plane_b, d_cmin + off, plane_b, (size_t)crows * nx * 4, nz,
                                       cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (a.profile && !prof_dev)
            OGN_CUDA(cudaMemcpy2DAsync(a.profile + off, img, d_prof + off, img, (size_t)crows * nx, nz,
                                       cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    // extremum pass on the complete device cubes while the last products are still travelling
    int rc = ogn_extrema_run(ctx, d_correl, d_cmin, d_mask, nz, ny, nx, owned, place, a.sz, a.sy, a.sx, a.dense_max,
                             a.dense_min, a.max_index, a.max_value, a.min_index, a.min_value, a.capacity, a.counts);
    if (rc != OGN_OK && rc != OGN_ERR_OVERFLOW) return rc;
    OGN_TRY(ogn_output_commit(ctx, a.maxmap, d_maxmap, img * 4));
    OGN_TRY(ogn_output_commit(ctx, a.minmap, d_minmap, img * 4));
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    OGN_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    ctx->host_output_pending = false;
    return rc;
}

int step05_run(ogn_ctx *ctx, const Step05Args &a, const int *tile) {
    if (!ctx) return OGN_ERR_ARG;
    const int nz = a.nz, ny = a.ny, nx = a.nx;
    if (nz <= 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "cube shape (%d,%d,%d) is empty", nz, ny, nx);
    OGN_CUDA(cudaSetDevice(ctx->device));
    ogn_place place{ny, nx, 0, 0};
    ogn_window owned{0, ny, 0, nx};
    if (tile) {
        place = ogn_place{tile[0], tile[1], tile[2], tile[3]};
        owned = ogn_window{tile[4], tile[5], tile[6], tile[7]};
        if (owned.y0 < 0 || owned.x0 < 0 || owned.y1 > ny || owned.x1 > nx || owned.y0 >= owned.y1 || owned.x0 >= owned.x1)
            return ogn_fail(ctx, OGN_ERR_ARG, "owned window [%d,%d)x[%d,%d) outside the %dx%d sub-cube", owned.y0,
                            owned.y1, owned.x0, owned.x1, ny, nx);
    }
    if (a.sz < 1 || a.sy < 1 || a.sx < 1 || !(a.sz & 1) || !(a.sy & 1) || !(a.sx & 1))
        return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "window (%d,%d,%d): only odd sizes are supported", a.sz, a.sy, a.sx);
    OGN_HT("step05 enter");
    ogn_tglr_guard guard(ctx);              // constant-memory taps: see ogn_common.cuh
    ogn_timer t_span(ctx, "step05_span");   // first to last device operation of the call (gaps included)
    ogn_tglr_setup_t st;
    OGN_TRY(ogn_tglr_setup(ctx, nz, ny, nx, &place, a.nfields, a.fsf, a.psize, a.weights, a.taps, a.tap_offsets,
                           a.nprof, true, &st));
    OGN_HT("setup done");
    float *peer_dst = nullptr;
    if (tile && ctx->local_gather && ctx->local_gather_is_peer) {
        // the gathered cube lives on another GPU: the owned window of correl is copied there as soon as the spectral
        // kernel has produced it (before the extremum pass and whatever the caller enqueues next), on the peer stream
        peer_dst = ctx->local_gather;
    } else if (tile && ctx->local_gather) {   // this rank owns the gathered cube: K2 stores its owned voxels there as well
        st.gather2.dst = ctx->local_gather;
        st.gather2.ny = place.gny; st.gather2.nx = place.gnx; st.gather2.dy = place.gy0; st.gather2.dx = place.gx0;
        st.gather2.y0 = owned.y0; st.gather2.y1 = owned.y1; st.gather2.x0 = owned.x0; st.gather2.x1 = owned.x1;
    }
    ctx->local_gather = nullptr;
    const size_t vol = (size_t)nz * ny * nx, img = (size_t)ny * nx;

    const bool host_in = !ogn_is_device_ptr(a.cube);
    static const bool no_stream = getenv("OGN_NO_STREAM") != nullptr;
    if (a.mask_bits && (nx % 8 != 0 || ogn_is_device_ptr(a.mask)))
        return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "a bit-packed mask must be host memory and nx a multiple of 8 (nx = %d)", nx);
    // host cube: slab-pipelined path (every product goes to its own destination, host or device)
    // TGLR on the owned window grown by the extremum radius (clipped to the sub-cube)
    // ... and to the left down to a multiple of 4 columns: K1's TMA box starts P/2 = 12 columns left of the
    // window and a TMA box must start on a 16-byte boundary (the few extra columns lie inside the sub-cube
    // and belong to a neighbour; computing them is harmless)
    const ogn_window w{std::max(0, owned.y0 - a.sy / 2), std::min(ny, owned.y1 + a.sy / 2),
                       std::max(0, owned.x0 - a.sx / 2) / 4 * 4, std::min(nx, owned.x1 + a.sx / 2)};
    const bool streamed = !no_stream && host_in && a.cube_dtype == OGN_F32 && !st.pervoxel && ny >= 64 && nx % 4 == 0 &&
                          (!a.mask || !ogn_is_device_ptr(a.mask)) && !st.gather2.dst;
    ctx->variants["step05"] = streamed ? "slab-pipelined" : "resident";
    if (streamed) return step05_streamed(ctx, a, st, owned, place, w);
    if (a.mask_bits) return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "a bit-packed mask is only taken on the streamed host path");

    void *d_correl = nullptr, *d_cmin = nullptr, *d_prof = nullptr, *d_maxmap = nullptr, *d_minmap = nullptr;
    // correl and correl_min are needed on the device even when the caller does not want them back
    OGN_TRY(ogn_output(ctx, "s5_correl", a.correl, vol * 4, &d_correl));
    OGN_TRY(ogn_output(ctx, "s5_correl_min", a.correl_min, vol * 4, &d_cmin));
    if (a.profile) OGN_TRY(ogn_output(ctx, "s5_profile", a.profile, vol, &d_prof));
    if (a.maxmap) OGN_TRY(ogn_output(ctx, "s5_maxmap", a.maxmap, img * 4, &d_maxmap));
    if (a.minmap) OGN_TRY(ogn_output(ctx, "s5_minmap", a.minmap, img * 4, &d_minmap));
    const void *d_mask = nullptr;
    if (a.mask) OGN_TRY(ogn_input(ctx, "s5_mask", a.mask, vol, &d_mask));
    const float *d_cube = nullptr;
    OGN_TRY(ogn_input_cube_f32(ctx, "cube", a.cube, a.cube_dtype, vol, &d_cube));
    OGN_TRY(ogn_tglr_init_maps(ctx, ctx->stream, (float *)d_maxmap, (float *)d_minmap, img));
    OGN_TRY(ogn_tglr_window(ctx, ctx->stream, st, d_cube, (const uint8_t *)d_mask, w, (float *)d_correl,
                            (float *)d_cmin, (uint8_t *)d_prof, (float *)d_maxmap, (float *)d_minmap));
    OGN_HT("tglr enqueued");
    if (peer_dst) OGN_TRY(ogn_scatter_tile(ctx, (const float *)d_correl, nz, ny, nx, tile, peer_dst));
    OGN_TRY(ogn_output_commit(ctx, a.correl, d_correl, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, a.correl_min, d_cmin, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, a.profile, d_prof, vol));
    OGN_TRY(ogn_output_commit(ctx, a.maxmap, d_maxmap, img * 4));
    OGN_TRY(ogn_output_commit(ctx, a.minmap, d_minmap, img * 4));
    int rc = ogn_extrema_run(ctx, (const float *)d_correl, (const float *)d_cmin, (const uint8_t *)d_mask, nz, ny, nx,
                             owned, place, a.sz, a.sy, a.sx, a.dense_max, a.dense_min, a.max_index, a.max_value,
                             a.min_index, a.min_value, a.capacity, a.counts);
    if (rc != OGN_OK && rc != OGN_ERR_OVERFLOW) return rc;
    OGN_TRY(ogn_finish_call(ctx));
    OGN_HT("step05 done");
    return rc;
}

}  // namespace

// ComputeTGLR.run (steps.py:768-802): Correlation_GLR_test + masking + maxmap/minmap +
// compute_local_max, with correl / correl_min / profile never leaving the device in between.
extern "C" int ogn_step05(ogn_ctx *ctx, const void *cube, int cube_dtype, int nz, int ny, int nx, int nfields,
                          const double *const *fsf, int psize, const double *const *weights, const double *taps,
                          const int *tap_offsets, int nprof, const uint8_t *mask, int sz, int sy, int sx,
                          float *correl, float *correl_min, uint8_t *profile, float *maxmap, float *minmap,
                          float *dense_max, float *dense_min, int64_t *max_index, float *max_value,
                          int64_t *min_index, float *min_value, int64_t capacity, int64_t *counts) {
    const Step05Args a{cube, cube_dtype, nz, ny, nx, nfields, fsf, psize, weights, taps, tap_offsets, nprof, mask,
                       sz, sy, sx, correl, correl_min, profile, maxmap, minmap, dense_max, dense_min, max_index,
                       max_value, min_index, min_value, capacity, counts};
    return step05_run(ctx, a, nullptr);
}

extern "C" int ogn_step05_tile(ogn_ctx *ctx, const void *cube, int cube_dtype, int nz, int ny, int nx,
                               const int *tile, int nfields, const double *const *fsf, int psize,
                               const double *const *weights, const double *taps, const int *tap_offsets, int nprof,
                               const uint8_t *mask, int sz, int sy, int sx, float *correl, float *correl_min,
                               uint8_t *profile, float *maxmap, float *minmap, int64_t *max_index, float *max_value,
                               int64_t *min_index, float *min_value, int64_t capacity, int64_t *counts) {
    if (!tile) return ogn_fail(ctx, OGN_ERR_ARG, "tile is NULL");
    const Step05Args a{cube, cube_dtype, nz, ny, nx, nfields, fsf, psize, weights, taps, tap_offsets, nprof, mask,
                       sz, sy, sx, correl, correl_min, profile, maxmap, minmap, nullptr, nullptr, max_index,
                       max_value, min_index, min_value, capacity, counts};
    return step05_run(ctx, a, tile);
}

// ogn_step05 with the mask given bit-packed (numpy.packbits of the flattened [nz][ny][nx] bool cube: 1/8 of the
// bytes on the PCIe link; unpacked on the device slab by slab).  Host cube only, nx a multiple of 8.
extern "C" int ogn_step05_bits(ogn_ctx *ctx, const void *cube, int cube_dtype, int nz, int ny, int nx, int nfields,
                               const double *const *fsf, int psize, const double *const *weights, const double *taps,
                               const int *tap_offsets, int nprof, const uint8_t *mask_bits, int sz, int sy, int sx,
                               float *correl, float *correl_min, uint8_t *profile, float *maxmap, float *minmap,
                               int64_t *max_index, float *max_value, int64_t *min_index, float *min_value,
                               int64_t capacity, int64_t *counts) {
    Step05Args a{cube, cube_dtype, nz, ny, nx, nfields, fsf, psize, weights, taps, tap_offsets, nprof, mask_bits,
                 sz, sy, sx, correl, correl_min, profile, maxmap, minmap, nullptr, nullptr, max_index,
                 max_value, min_index, min_value, capacity, counts};
    a.mask_bits = mask_bits ? 1 : 0;
    return step05_run(ctx, a, nullptr);
}
