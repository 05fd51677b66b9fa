/*
 * libogn — B200-native (sm_100a) implementation of ORIGIN's detection hot path.
 *
 * C ABI.  Plain pointers and sizes only; no torch / numpy types.  Every entry
 * point replaces one numerical function of the reference's L2 library
 * (musevlt/origin, muse_origin/lib_origin.py) or one inline block of its step
 * layer (muse_origin/steps.py); the line ranges are quoted next to each
 * declaration.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - Cubes are C-ordered [nz][ny][nx] (lambda, y, x), as numpy arrays in the
 *    reference.  Images are [ny][nx].
 *  - Every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE
 *    pointer on the context's device; the library asks the CUDA runtime which
 *    (cudaPointerGetAttributes) and stages host buffers itself.  NULL output
 *    pointers mean "do not return this product".
 *  - Buffers are caller-owned.  The context owns all scratch memory.
 *  - Return value: 0 (OGN_OK) or a negative ogn_status; ogn_last_error() gives
 *    the message.  There is no CPU fallback: without a usable CUDA device
 *    ogn_create fails.
 *  - One context per (process, device, stream).  Calls on one context must be
 *    serialised by the caller.  TGLR calls (ogn_tglr, ogn_step05*) of different
 *    contexts on the same device are serialised by the library: the profile
 *    taps of the call in flight live in __constant__ memory, which the device's
 *    contexts share, so the enqueue sections hold a process-wide lock and a
 *    context that follows another one makes its stream wait for the event the
 *    other recorded behind its last TGLR kernel (processes sharing one GPU are
 *    separate CUDA contexts with their own constants).  All work
 *    is enqueued on the context's stream and
 *    every entry point that returns data to host memory synchronises that
 *    stream before returning.  Calls whose outputs are all device pointers
 *    return without synchronising.
 */
#ifndef OGN_H
#define OGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGN_VERSION 100

#if defined(__GNUC__)
#define OGN_API __attribute__((visibility("default")))
#else
#define OGN_API
#endif

typedef struct ogn_ctx ogn_ctx;

typedef enum {
    OGN_OK = 0,
    OGN_ERR_CUDA = -1,        /* a CUDA runtime/driver call failed            */
    OGN_ERR_ARG = -2,         /* invalid argument                             */
    OGN_ERR_NOMEM = -3,       /* device or host allocation failed             */
    OGN_ERR_UNSUPPORTED = -4, /* valid request the kernels do not cover       */
    OGN_ERR_OVERFLOW = -5     /* an output list was too small (count returned) */
} ogn_status;

typedef enum { OGN_F32 = 0, OGN_F64 = 1 } ogn_dtype;

/* ---- context ---------------------------------------------------------- */

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t (or
 * NULL for the legacy default stream); the caller keeps ownership of it. */
OGN_API int ogn_create(int device, void *stream, ogn_ctx **out);
OGN_API void ogn_destroy(ogn_ctx *ctx);
/* Message of the last failing call on `ctx` (or of the last failing
 * ogn_create when ctx == NULL).  Valid until the next call. */
OGN_API const char *ogn_last_error(const ogn_ctx *ctx);
OGN_API int ogn_version(void);
/* Block until everything enqueued on the context's stream has finished. */
OGN_API int ogn_synchronize(ogn_ctx *ctx);
/* Number of kernels this context has launched so far (bench bookkeeping). */
OGN_API int64_t ogn_launch_count(const ogn_ctx *ctx);
/* Per-stage CUDA-event timing on the context's stream (off by default).
 * ogn_timing_report synchronises, writes "stage:ms;stage:ms;..." for every
 * stage timed since the previous report into buf, and clears the list. */
OGN_API int ogn_timing_enable(ogn_ctx *ctx, int on);
OGN_API int ogn_timing_report(ogn_ctx *ctx, char *buf, size_t size);
/* Diagnostics of the last ogn_tglr / ogn_fsf_stage / ogn_step05* call on `ctx`: *folded receives 1
 * when every FSF plane was mirror-symmetric in y (as float32) and the spatial stage therefore added
 * footprint rows dy and P-1-dy before multiplying (half the FFMAs of _convolve_fsf's direct form,
 * lib_origin.py:1027-1043), 0 when the general kernel ran.  Synchronises the stream. */
OGN_API int ogn_fsf_folded(ogn_ctx *ctx, int *folded);
/* Which code path each stage took on its last launch on `ctx`, as "stage=path;..." (stages: k1, k2, k3,
 * step01, step05).  The reference has one code path per function (lib_origin.py:150, 1027, 1070, 1220); this
 * library picks between kernels by shape, alignment and OGN_* diagnostic switches, and the tests pin the
 * defaults with this call.  No device work. */
OGN_API int ogn_variants(ogn_ctx *ctx, char *buf, size_t size);
/* Release the context's scratch memory (it is re-grown on demand). */
OGN_API int ogn_trim(ogn_ctx *ctx);
/* Pinned host memory for fast staging (optional; any host pointer works). */
OGN_API int ogn_host_alloc(size_t bytes, void **out);
OGN_API int ogn_host_free(void *ptr);

/* ---- step05: TGLR matched filter --------------------------------------- */

/* Correlation_GLR_test (lib_origin.py:1070-1217, helpers :1027-1066) fused
 * with the masking / maxmap / minmap glue of ComputeTGLR.run
 * (steps.py:781-793).
 *
 *  cube        [nz][ny][nx], dtype `cube_dtype` (computed in f32)
 *  nfields     1 for a single FSF (reference: weights is None), else the
 *              number of mosaic fields
 *  fsf         nfields pointers to [nz][psize][psize] float64 FSF cubes
 *              (orig.PSF, origin.py:590-646)
 *  weights     NULL (single field) or nfields pointers to [ny][nx] float64
 *              weight maps (orig.wfields)
 *  taps        the profiles AFTER the preparation of lib_origin.py:1155-1165
 *              (cut at pcut, L2-normalised, mean-subtracted), float64,
 *              concatenated; profile k is taps[tap_offsets[k]:tap_offsets[k+1]]
 *  mask        NULL, or [nz][ny][nx] bytes (non-zero = masked): correl and
 *              profile are zeroed under the mask (steps.py:781,788),
 *              correl_min is not
 *  correl, correl_min   [nz][ny][nx] float32 out
 *  profile     [nz][ny][nx] uint8 out (first-wins argmax, lib_origin.py:1210)
 *  maxmap      [ny][nx] float32 out = max_z correl (after masking, steps.py:792)
 *  minmap      [ny][nx] float32 out = min_z correl_min (steps.py:793)
 */
OGN_API int ogn_tglr(ogn_ctx *ctx,
             const void *cube, int cube_dtype, int nz, int ny, int nx,
             int nfields, const double *const *fsf, int psize,
             const double *const *weights,
             const double *taps, const int *tap_offsets, int nprof,
             const uint8_t *mask,
             float *correl, float *correl_min, uint8_t *profile,
             float *maxmap, float *minmap);

/* The two intermediate cubes of the spatial stage only (_convolve_fsf,
 * lib_origin.py:1027-1043, summed over fields :1143-1147); float32
 * [nz][ny][nx] each.  Used by the parity tests. */
OGN_API int ogn_fsf_stage(ogn_ctx *ctx,
                  const void *cube, int cube_dtype, int nz, int ny, int nx,
                  int nfields, const double *const *fsf, int psize,
                  const double *const *weights,
                  float *cube_fsf, float *norm_fsf);

/* ---- 3-D local extrema -------------------------------------------------- */

/* compute_local_max (lib_origin.py:1220-1256): local maxima of `a` and of
 * `-b` over a (sz,sy,sx) window (odd sizes, scipy 'reflect' edges), excluding
 * masked voxels.  `a` and `b` are float32 [nz][ny][nx] (they may alias: step01
 * calls it on (cube_std, cube_std), steps.py:453).
 *
 * Dense products (either may be NULL): dense_max = a * keep(a),
 * dense_min = (-b) * keep(-b), float32 [nz][ny][nx].
 *
 * Compact products: the kept voxels as lists sorted by (z, y, x) — i.e. by
 * C-order linear index, the order of np.where (steps.py:958).  max_index /
 * min_index receive the linear voxel index (int64), max_value / min_value the
 * value (for the minima list: -b, as in the reference).  `capacity` is the
 * size of each list; counts[0], counts[1] always receive the true counts and
 * OGN_ERR_OVERFLOW is returned when a list did not fit.  Lists may be NULL
 * (counts only).  When `counts` is a DEVICE array and all products are device
 * buffers the call does not synchronise and cannot report an overflow: the
 * lists are then truncated at `capacity` and the caller compares counts with
 * capacity once it reads them (same for ogn_step05 / ogn_step05_tile). */
OGN_API int ogn_local_extrema(ogn_ctx *ctx,
                      const float *a, const float *b, const uint8_t *mask,
                      int nz, int ny, int nx, int sz, int sy, int sx,
                      float *dense_max, float *dense_min,
                      int64_t *max_index, float *max_value,
                      int64_t *min_index, float *min_value,
                      int64_t capacity, int64_t *counts);

/* The whole array part of ComputeTGLR.run (steps.py:768-802) in one call:
 * ogn_tglr followed by ogn_local_extrema on the masked correl / correl_min,
 * with the intermediates staying on the device.  Arguments are those of the
 * two functions; correl and correl_min may be NULL when only the extremum
 * lists are wanted.  Returns OGN_ERR_OVERFLOW like ogn_local_extrema (all
 * other outputs are complete in that case). */
OGN_API int ogn_step05(ogn_ctx *ctx,
               const void *cube, int cube_dtype, int nz, int ny, int nx,
               int nfields, const double *const *fsf, int psize,
               const double *const *weights,
               const double *taps, const int *tap_offsets, int nprof,
               const uint8_t *mask, int sz, int sy, int sx,
               float *correl, float *correl_min, uint8_t *profile,
               float *maxmap, float *minmap,
               float *dense_max, float *dense_min,
               int64_t *max_index, float *max_value,
               int64_t *min_index, float *min_value,
               int64_t capacity, int64_t *counts);

/* ogn_step05 (ComputeTGLR.run, steps.py:768-802) for a HOST cube with the mask given bit-packed: mask_bits = numpy.packbits of the flattened
 * [nz][ny][nx] boolean mask (MSB first), 1/8 of the bytes on the PCIe link; it is unpacked on the device slab
 * by slab.  nx must be a multiple of 8.  Every product may be a host buffer (copied back slab by slab while
 * later slabs are computed) or a device buffer (written in place, e.g. products the caller fetches lazily). */
OGN_API int ogn_step05_bits(ogn_ctx *ctx,
                    const void *cube, int cube_dtype, int nz, int ny, int nx,
                    int nfields, const double *const *fsf, int psize,
                    const double *const *weights,
                    const double *taps, const int *tap_offsets, int nprof,
                    const uint8_t *mask_bits, int sz, int sy, int sx,
                    float *correl, float *correl_min, uint8_t *profile,
                    float *maxmap, float *minmap,
                    int64_t *max_index, float *max_value,
                    int64_t *min_index, float *min_value,
                    int64_t capacity, int64_t *counts);

/* ogn_step05 (ComputeTGLR.run, steps.py:768-802; the reference is single-process) on one spatial tile of a
 * larger field (multi-GPU runs).  `cube` is
 * the [nz][ny][nx] sub-cube cut from the field with its halo;
 *   tile = {gny, gnx, gy0, gx0, oy0, oy1, ox0, ox1}
 * says that spaxel (0,0) of the sub-cube is spaxel (gy0,gx0) of the gny x gnx
 * field and that the rank owns rows [oy0,oy1) and columns [ox0,ox1) of the
 * sub-cube.  Sub-cube edges that are not field edges must lie at least
 * psize/2 + 1 pixels outside the owned window.  Only the owned window grown by
 * the extremum radius is computed; the products are sub-cube shaped and valid
 * on that window; the extremum lists hold the owned voxels with linear indices
 * of the WHOLE field ([nz][gny][gnx]), so per-rank lists concatenate. */
OGN_API int ogn_step05_tile(ogn_ctx *ctx,
                    const void *cube, int cube_dtype, int nz, int ny, int nx,
                    const int *tile,
                    int nfields, const double *const *fsf, int psize,
                    const double *const *weights,
                    const double *taps, const int *tap_offsets, int nprof,
                    const uint8_t *mask, int sz, int sy, int sx,
                    float *correl, float *correl_min, uint8_t *profile,
                    float *maxmap, float *minmap,
                    int64_t *max_index, float *max_value,
                    int64_t *min_index, float *min_value,
                    int64_t capacity, int64_t *counts);

/* ---- multi-GPU: gather of owned tiles over NVLink peer memory ------------ */

/* north_star: "correl is gathered to rank 0" (the reference has no distributed code).  Rank 0
 * allocates the full [nz][gny][gnx] product with ogn_peer_alloc, which also returns the 64-byte
 * CUDA IPC handle of the buffer; the host layer ships the handle to the other ranks (any transport:
 * torch.distributed, MPI, a file), which map the buffer with ogn_peer_open.  ogn_scatter_tile then
 * copies the window a rank owns of its sub-cube product `src` ([nz][ny][nx] device memory; `tile`
 * as in ogn_step05_tile) into `dst` — the mapped peer buffer, or the local one on the owning rank —
 * with a kernel whose stores cross NVLink.  The copy is enqueued on an internal stream behind the
 * work already queued on the context's stream and the call returns at once, so the transfer overlaps
 * the next step; a later TGLR call that overwrites `src` waits for it on the device.  ogn_peer_join
 * makes the context's stream wait for all copies enqueued so far, ogn_peer_sync the host; the
 * ranks still need a barrier of their own before the owner reads `dst`. */
OGN_API int ogn_peer_alloc(ogn_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char *handle64);
OGN_API int ogn_peer_free(ogn_ctx *ctx, void *dev_ptr);
OGN_API int ogn_peer_open(ogn_ctx *ctx, const unsigned char *handle64, void **dev_ptr);
OGN_API int ogn_peer_close(ogn_ctx *ctx, void *dev_ptr);
OGN_API int ogn_scatter_tile(ogn_ctx *ctx, const float *src, int nz, int ny, int nx, const int *tile, float *dst);
/* Make the next ogn_step05_tile call deliver the window it owns of correl into `dst` by itself.  On the rank
 * that owns the gathered cube (`dst` from ogn_peer_alloc) the spectral kernel stores the owned voxels there as
 * well; on the other ranks (`dst` from ogn_peer_open) the copy of ogn_scatter_tile is enqueued right behind the
 * spectral kernel, before the extremum pass, so it overlaps the rest of the step and the next one.  One-shot. */
OGN_API int ogn_set_local_gather(ogn_ctx *ctx, float *dst);
/* Stagger: every following ogn_scatter_tile of this context waits `microseconds` on the device (on the copy's
 * stream, not the context's) before it starts.  The ranks share the destination GPU's NVLink ingress; taking
 * turns (delay = turn x time of one tile on the link alone) keeps a source's memory system free of stalled
 * remote stores except during its own turn.  0 switches it off. */
OGN_API int ogn_peer_set_delay(ogn_ctx *ctx, int microseconds);
OGN_API int ogn_peer_join(ogn_ctx *ctx);
OGN_API int ogn_peer_sync(ogn_ctx *ctx);

/* ---- step06: purity threshold counts ------------------------------------ */

/* Statistics Compute_threshold_purity derives its default threshold list from
 * (lib_origin.py:1424-1439), computed from the compact extremum lists:
 *   stats[0] = max of the maxima list, stats[1] = max of the background minima
 *   list (entries whose spaxel has segmask != 0 are dropped, :1428-1429),
 *   spaxel_max [ny][nx] float32 = per-spaxel max over lambda of the dense
 *   local-max cube (0 where a spaxel has no positive maximum), the argument of
 *   the median at :1438.
 * `segmask` is NULL or [ny][nx] bytes, non-zero = spaxel belongs to a source
 * (segmap != 0). */
OGN_API int ogn_purity_stats(ogn_ctx *ctx,
                     const int64_t *max_index, const float *max_value, int64_t nmax,
                     const int64_t *min_index, const float *min_value, int64_t nmin,
                     const uint8_t *segmask, int ny, int nx,
                     double *stats, float *spaxel_max);

/* The counting loop of Compute_threshold_purity (lib_origin.py:1443-1449):
 * n1[t] = #{maxima > thresholds[t]}, n0[t] = #{background minima >
 * thresholds[t]}, int64 each.  Multi-GPU callers sum n0/n1 across ranks. */
OGN_API int ogn_purity_counts(ogn_ctx *ctx,
                      const int64_t *max_index, const float *max_value, int64_t nmax,
                      const int64_t *min_index, const float *min_value, int64_t nmin,
                      const uint8_t *segmask, int ny, int nx,
                      const double *thresholds, int nthresh,
                      int64_t *n1, int64_t *n0);

/* Synchronisation-free variant of the same counting loop (lib_origin.py:1443-1449) for device-resident pipelines
 * (multi-GPU steps): every pointer is a device
 * pointer; the lists are the capacity-sized buffers of an ogn_step05 / ogn_step05_tile / ogn_local_extrema
 * call whose `counts` output was a DEVICE array (such a call does not synchronise either), and
 * list_counts is that array: the kernels read the true list lengths from it.  n1 / n0 can be handed to
 * an NCCL allreduce without the host ever seeing them.  A list that overflowed `capacity` cannot be
 * reported by a return code here: the kernel then adds -2^56 to every count of that list, so a NEGATIVE
 * count (also after a SUM over ranks) means "extremum list overflow: repeat the step with a larger
 * capacity". */
OGN_API int ogn_purity_counts_dev(ogn_ctx *ctx,
                          const int64_t *max_index, const float *max_value,
                          const int64_t *min_index, const float *min_value, int64_t capacity,
                          const int64_t *list_counts, const uint8_t *segmask, int ny, int nx,
                          const double *thresholds, int nthresh, int64_t *n1, int64_t *n0);

/* ---- step07: thresholding ------------------------------------------------ */

/* The np.where block of Detection.run (steps.py:956-974): entries of a
 * compact extremum list with value > threshold, in list order (= C order),
 * optionally with the argmax profile read from `profile`
 * ([nz][ny][nx] uint8, or NULL).  out_count receives the true count;
 * OGN_ERR_OVERFLOW if capacity was too small. */
OGN_API int ogn_threshold_extract(ogn_ctx *ctx,
                          const int64_t *index, const float *value, int64_t n,
                          double threshold, const uint8_t *profile,
                          int64_t *out_index, float *out_value, uint8_t *out_profile,
                          int64_t capacity, int64_t *out_count);

/* ---- step01: DCT continuum ------------------------------------------------ */

/* dct_residual (lib_origin.py:150-240, DCTMAT :127-146).  raw and var are
 * [nz][ny][nx] of `in_dtype` with the NaN convention of origin.py:262-274
 * (raw 0 / var +inf where invalid); mask is [nz][ny][nx] bytes.  The
 * continuum is computed in float64 and returned as `out_dtype`. */
OGN_API int ogn_dct_residual(ogn_ctx *ctx,
                     const void *raw, const void *var, int in_dtype,
                     const uint8_t *mask, int nz, int ny, int nx,
                     int order, int approx,
                     void *cont, int out_dtype);

/* Preprocessing.run, array part (steps.py:431-465, :472, :480), in two
 * phases so that the per-wavelength mean (np.nanmean over all spaxels,
 * steps.py:442) can be reduced across ranks in between.
 *
 * begin : continuum fit (the order+1 coefficients per spaxel stay on the device;
 *         the continuum itself is re-synthesised where needed, never stored),
 *         and the per-wavelength partial sums lambda_sum[nz] / lambda_cnt[nz]
 *         (float64, host or device) of the unmasked data = raw - cont of THIS
 *         cube (tile).  `owned` is NULL or {y0, y1, x0, x1}: only spaxels of that
 *         window of the given cube enter the sums (a tile's halo belongs to its
 *         neighbours).  Device inputs must stay valid until ogn_preprocess_finish.
 * finish: given the global per-wavelength mean, standardise and reduce:
 *         cube_std [nz][ny][nx] f32, cont_dct [nz][ny][nx] f32 (= cont/sqrt(var)),
 *         ima_std, ima_dct, cont_sumsq (= sum_z cont_dct^2), o2map
 *         (= mean_z cube_std^2) : [ny][nx] float64.  Any output may be NULL.
 */
OGN_API int ogn_preprocess_begin(ogn_ctx *ctx,
                         const void *raw, const void *var, int in_dtype,
                         const uint8_t *mask, int nz, int ny, int nx,
                         int order, int approx, const int *owned,
                         double *lambda_sum, double *lambda_cnt);
OGN_API int ogn_preprocess_finish(ogn_ctx *ctx, const double *lambda_mean,
                          float *cube_std, float *cont_dct,
                          double *ima_std, double *ima_dct,
                          double *cont_sumsq, double *o2map);

/* Both phases in one call, for a whole field on one device (no cross-rank reduction in between): the
 * per-wavelength mean (np.nanmean(data, axis=(1, 2)), steps.py:442; NaN for a plane without unmasked voxel)
 * is formed on the device and, optionally, returned in lambda_mean[nz].  Outputs as in
 * ogn_preprocess_finish; a call whose cubes are all device buffers does not synchronise. */
OGN_API int ogn_preprocess(ogn_ctx *ctx,
                   const void *raw, const void *var, int in_dtype,
                   const uint8_t *mask, int nz, int ny, int nx,
                   int order, int approx, double *lambda_mean,
                   float *cube_std, float *cont_dct,
                   double *ima_std, double *ima_dct,
                   double *cont_sumsq, double *o2map);

/* ---- step04: greedy PCA -------------------------------------------------------- */

/* Compute_GreedyPCA (lib_origin.py:858-954, with O2test :957-974 and orthogonal_projection :76-88) on the
 * spaxels `cols[0..n)` (column indices; NULL = all, then n == ld) of a [nz][ld] cube: until no spaxel's
 * second-order test mean_z x^2 exceeds `thres`, the first left singular vector of the nuisance spectra
 * (orthogonalised to the mean spectrum of the quietest 1 / noise_population of the background spaxels) is
 * projected out of every spectrum of the block.  FP64 on the device.
 *   cube    float32 / float64 (`dtype`), host or device, read only
 *   test0   NULL (the test of the block is computed: Compute_PCA_threshold, :836) or [n] float64, host
 *   faint   [nz][ld] cube of `out_dtype`, host or device: only the columns `cols` are written, so the caller
 *           starts from a copy of `cube` like the reference (Compute_GreedyPCA_area, :797) or passes cube itself
 *   map_o2  [n] float64, host: number of iterations each spaxel spent above the threshold (:887)
 *   info    {nstop (iteration limit hit, :888-891), iterations, operator applications of the SVD} */
OGN_API int ogn_greedy_pca(ogn_ctx *ctx, const void *cube, int dtype, int nz, int64_t ld,
                   const int64_t *cols, int64_t n, const double *test0, double thres,
                   double noise_population, int itermax, void *faint, int out_dtype,
                   double *map_o2, int *info);

/* ---- step08: line estimation ----------------------------------------------------- */

/* method_PCA_wgt (lib_origin.py:1535-1617, with LS_deconv_wgt :1482-1510 and conv_wgt :1513-1532) for a batch of
 * P x P x nz windows of the raw cube - what GridAnalysis (:1620-1790) computes at every offset of its grid for every
 * detection before its scalar criteria (peak search, flux / mse), which stay host code.  FP64 on the device.
 *   raw, var      [nz][ny][nx] float32 / float64 (`dtype`), host or device; var = +inf marks invalid voxels
 *                 (origin.py:262-274); windows may stick out of the image (raw 0 / var +inf there, :1893-1897)
 *   psf           [nz][P][P] float64, single field
 *   centres       [npos][2] int32: (y, x) of the centre of each window
 *   order_dct     order of the DCT that denoises the second eigenvector; < 0 = PCA LS only (order_dct=None)
 *   line, linevar [npos][nz] float64, host or device: estimated line and its theoretical variance
 *   info          NULL or {operator applications of the SVDs, problems per batch} */
OGN_API int ogn_line_estimates(ogn_ctx *ctx, const void *raw, const void *var, int dtype,
                       int nz, int ny, int nx, const double *psf, int P,
                       const int *centres, int npos, int order_dct,
                       double *line, double *linevar, int *info);

/* The same for weighted mosaics (estimation_line with wght, lib_origin.py:1899-1906): the FSF of a window is
 * the combination of the fields' FSFs with the weight maps cut to that window, which GridAnalysis forms at
 * every grid offset (:1713-1717),  psf_eff[p][z][j] = sum_f coef[p][f][j] * psf[f][z][j]  (products rounded and
 * added in field order, as np.sum(axis=0)).  The host mirror fills `coef`, including the way the reference's
 * loop re-uses the combined FSF of the previous offset.
 *   psf    [nf][nz][P][P] float64, host or device
 *   coef   [npos][nf][P][P] float64, host or device */
OGN_API int ogn_line_estimates_fields(ogn_ctx *ctx, const void *raw, const void *var, int dtype,
                       int nz, int ny, int nx, const double *psf, int nf, int P,
                       const double *coef, const int *centres, int npos, int order_dct,
                       double *line, double *linevar, int *info);

#ifdef __cplusplus
}
#endif
#endif /* OGN_H */
