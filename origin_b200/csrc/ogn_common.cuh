// Shared plumbing of libogn: context, scratch arena, host<->device staging.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ogn.h"

struct ogn_buf {
    void *p = nullptr;
    size_t cap = 0;
};

// State kept between ogn_preprocess_begin and ogn_preprocess_finish.
struct ogn_prep_state {
    bool active = false;
    int nz = 0, ny = 0, nx = 0, in_dtype = 0;
    const void *raw = nullptr;      // device
    const void *var = nullptr;      // device
    const uint8_t *mask = nullptr;  // device
    const double *coef = nullptr;   // device: DCT coefficients [M][S]; the continuum is re-synthesised, never stored
    const double *d0 = nullptr;     // device: DCTMAT [nz][M]
    double *d0p = nullptr;          // device: padded DCTMAT rows + mean column (streamed kernels), or nullptr
    int M = 0;
};

struct ogn_timing_entry {
    std::string name;
    cudaEvent_t start = nullptr, stop = nullptr;
};

struct ogn_ctx {
    bool timing = false;
    std::vector<ogn_timing_entry> timings;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;
    std::map<std::string, ogn_buf> bufs;  // named device scratch, grow-only
    std::map<std::string, ogn_buf> pins;  // named pinned host scratch, grow-only
    bool host_output_pending = false;     // a D2H copy to caller memory was enqueued
    ogn_prep_state prep;
    int dct_tab_nz = 0, dct_tab_M = 0;    // (nz, order + 1) the cached DCT tables were built for (ogn_dct.cu)
    // side streams of the streamed host path (created on first use)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> events;
    // peer gather (ogn_peer.cu)
    cudaStream_t peer_stream = nullptr;            // stream of the latest scatter (one of peer_streams)
    cudaStream_t peer_streams[2] = {nullptr, nullptr};   // consecutive scatters alternate: a staggered copy's delay must not queue behind the previous copy
    int peer_turn = 0;
    cudaEvent_t peer_ev_begin = nullptr, peer_ev_end = nullptr, peer_ev_end2 = nullptr;
    std::vector<void *> peer_owned, peer_mapped;
    std::map<const void *, cudaEvent_t> readers;   // source buffer -> event of the last scatter reading it
    // zero-copy staging (ogn_api.cu): small host -> device tables and device -> host scalars travel through
    // mapped pinned memory read / written by kernels, so the hot path never queues on a copy engine
    unsigned char *stage_h = nullptr, *stage_d = nullptr;
    size_t stage_cap = 0;
    cudaEvent_t stage_ev = nullptr;
    bool stage_pending = false;
    int64_t *res_h = nullptr, *res_d = nullptr;    // 32 mapped int64 result slots
    float *local_gather = nullptr;                 // ogn_set_local_gather: consumed by the next ogn_step05_tile
    int peer_delay_us = 0;                         // ogn_peer_set_delay
    cudaEvent_t tglr_done = nullptr;               // recorded behind the last TGLR kernel of a call (ogn_tglr_guard)
    std::map<std::string, std::string> variants;   // stage -> code path of its last launch (ogn_variants)
    bool local_gather_is_peer = false;             // ... it is another device's buffer (mapped with ogn_peer_open)
};

// Batches small host -> device copies into ONE kernel that reads a mapped pinned staging buffer over
// PCIe and writes every destination (no copy-engine operation, one launch instead of one memcpy each).
struct ogn_upload_item { void *dst; unsigned off, bytes; };
constexpr int OGN_UPLOAD_MAX = 16;
struct ogn_upload_table { ogn_upload_item item[OGN_UPLOAD_MAX]; int n; };
struct ogn_uploader {
    ogn_ctx *ctx;
    size_t used = 0;
    ogn_upload_table tab;
    explicit ogn_uploader(ogn_ctx *c) : ctx(c) { tab.n = 0; }
    int add(void *dst, const void *src, size_t bytes);   // bytes: multiple of 4
    int flush(cudaStream_t stream);
};
// p[0 .. bytes) = repeated 32-bit `word` (bytes: multiple of 4), by a kernel on `stream`
int ogn_fill_words(ogn_ctx *ctx, cudaStream_t stream, void *p, unsigned word, size_t bytes);
// mapped result slots (device pointer, host pointer); valid after the stream that wrote them is synchronised
int ogn_result_slots(ogn_ctx *ctx, int64_t **dev, int64_t **host);

int ogn_fail(ogn_ctx *ctx, int code, const char *fmt, ...);

// Host wall-clock trace points (OGN_HOST_TRACE=1): "label +ms" since the previous point, on stderr.
void ogn_host_trace(const char *label);
#define OGN_HT(label) ogn_host_trace(label)

// Scoped CUDA-event timer on the context's stream (active only after ogn_timing_enable).
struct ogn_timer {
    ogn_ctx *ctx;
    int slot = -1;
    ogn_timer(ogn_ctx *c, const char *name) : ctx(c) {
        if (!ctx->timing) return;
        ogn_timing_entry e;
        e.name = name;
        if (cudaEventCreate(&e.start) != cudaSuccess || cudaEventCreate(&e.stop) != cudaSuccess) return;
        cudaEventRecord(e.start, ctx->stream);
        ctx->timings.push_back(e);
        slot = (int)ctx->timings.size() - 1;
    }
    ~ogn_timer() {
        if (slot >= 0) cudaEventRecord(ctx->timings[slot].stop, ctx->stream);
    }
};

#define OGN_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ogn_fail(ctx, OGN_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

#define OGN_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != OGN_OK) return rc__; \
    } while (0)

#define OGN_LAUNCH_CHECK(name)                                                               \
    do {                                                                                     \
        ctx->launches++;                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return ogn_fail(ctx, OGN_ERR_CUDA, "launch of %s failed: %s", name,              \
                            cudaGetErrorString(e__));                                        \
    } while (0)

// Named device scratch buffer of at least `bytes` (contents undefined).
int ogn_scratch(ogn_ctx *ctx, const char *name, size_t bytes, void **out);
// Named pinned host scratch.
int ogn_pinned(ogn_ctx *ctx, const char *name, size_t bytes, void **out);
// true when `p` can be dereferenced by kernels on the context's device.
bool ogn_is_device_ptr(const void *p);
// Device view of an input: `p` itself when it is a device pointer, else a copy
// in the scratch buffer `name` (async on the context stream).
int ogn_input(ogn_ctx *ctx, const char *name, const void *p, size_t bytes, const void **dev);
// Device buffer for an output: `p` itself when it is a device pointer, else
// the scratch buffer `name`; ogn_output_commit copies it back when needed.
int ogn_output(ogn_ctx *ctx, const char *name, void *p, size_t bytes, void **dev);
int ogn_output_commit(ogn_ctx *ctx, void *p, const void *dev, size_t bytes);
// Synchronise the stream if any host output is pending.
int ogn_finish_call(ogn_ctx *ctx);

template <typename T>
static inline int ogn_scratch_t(ogn_ctx *ctx, const char *name, size_t count, T **out) {
    void *p = nullptr;
    int rc = ogn_scratch(ctx, name, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

static inline int ogn_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t ogn_round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Elementwise conversion kernels (ogn_api.cu)
int ogn_convert_f64_to_f32(ogn_ctx *ctx, const double *src, float *dst, size_t n);
int ogn_convert_f32_to_f64(ogn_ctx *ctx, const float *src, double *dst, size_t n);
// Device f32 view [n] of a cube given as host/device f32/f64.
int ogn_input_cube_f32(ogn_ctx *ctx, const char *name, const void *p, int dtype, size_t n,
                       const float **dev);

// ---- tiled / windowed execution (ogn_tglr.cu, ogn_extrema.cu, ogn_steps.cu) -------------------
struct ogn_window { int y0, y1, x0, x1; };        // half-open, coordinates of the sub-cube passed in
struct ogn_place { int gny, gnx, gy0, gx0; };     // where that sub-cube sits in the whole field

// Dictionary of the folded spectral kernel K2f (ogn_tglr_fold.cu), passed by value as a kernel
// parameter so that the taps are read through the uniform datapath from the parameter bank.
namespace k2f {
constexpr int ZB = 8;       // wavelengths per thread and pass
constexpr int GMAX = 10;    // profiles sharing one set of folded samples
constexpr int MAXG = 16;    // groups
constexpr int MAXT = 2048;  // tap table entries
constexpr int MAXB = 32;    // blocks of ZB tap distances per group (half-lengths up to 248)
struct FoldDict {
    unsigned char nab[MAXG][MAXB]; // [group][block]: profiles of the group active in the block (a suffix of the slots)
    int k[MAXG][GMAX];             // [group][slot]: profile index, -1 = empty slot
    int toff[MAXG], nblk[MAXG];    // tap table offset (floats) and number of blocks of the group
    int ngroups, G, hmax, nprof;   // hmax: largest half-length, padded to a multiple of ZB
};
}  // namespace k2f

// Second destination of K2's correl stores (multi-GPU): the rank that owns the gathered cube writes the
// voxels it owns straight into it, instead of copying them afterwards.
struct ogn_gather2 {
    float *dst = nullptr;          // [nz][ny][nx] on this device, or nullptr
    int ny = 0, nx = 0;            // field size
    int dy = 0, dx = 0;            // sub-cube origin in the field
    int y0 = 0, y1 = 0, x0 = 0, x1 = 0;   // owned window, sub-cube coordinates
};

struct ogn_tglr_setup_t {
    int nz = 0, ny = 0, nx = 0, P = 0, WP = 0, nfields = 0, nprof = 0;
    bool pervoxel = false;
    ogn_place place{0, 0, 0, 0};
    float *w32 = nullptr, *w32sq = nullptr, *rs = nullptr;
    int *asym = nullptr;   // device flag: != 0 when the FSF weights are not mirror-symmetric in y
    int nzp = 0, ncy = 0, ncx = 0;
    float *d_taps = nullptr, *d_taps_sq = nullptr;
    const void *d_desc = nullptr;
    int ntaps_total = 0, reach = 0, woff_min = 0;
    std::vector<const double *> w_dev;
    std::vector<ogn_window> w_box;   // per field: bounding box of the support of its weight map (sub-cube coordinates)
    ogn_gather2 gather2;
    std::shared_ptr<k2f::FoldDict> fold;   // set when the dictionary qualifies for K2f (taps already in constant memory)
};

// The profile taps of a TGLR call in flight live in __constant__ memory, which all contexts of a device share.
// Every entry point that runs the spectral stage holds this guard from before the setup (which rewrites the
// constants) until its last kernel is enqueued: the enqueue sections of different contexts cannot interleave on the
// host (process-wide mutex), and a context that follows another one on the same device first makes its stream
// wait for the event the other recorded behind its last TGLR kernel.
struct ogn_tglr_guard {
    ogn_ctx *ctx;
    explicit ogn_tglr_guard(ogn_ctx *c);
    ~ogn_tglr_guard();
    ogn_tglr_guard(const ogn_tglr_guard &) = delete;
    ogn_tglr_guard &operator=(const ogn_tglr_guard &) = delete;
};
void ogn_tglr_forget(ogn_ctx *ctx);   // ogn_destroy: the context leaves the per-device bookkeeping

int ogn_tglr_setup(ogn_ctx *ctx, int nz, int ny, int nx, const ogn_place *place, int nfields,
                   const double *const *fsf, int psize, const double *const *weights, const double *taps,
                   const int *tap_offsets, int nprof, bool need_spectral, ogn_tglr_setup_t *st);
int ogn_tglr_window(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, const float *dcube,
                    const uint8_t *dmask, ogn_window w, float *d_correl, float *d_cmin, uint8_t *d_prof,
                    float *d_maxmap, float *d_minmap);
bool ogn_k2f_prepare(const double *taps, const int *tap_offsets, int nprof, k2f::FoldDict *d, std::vector<float> *table);
int ogn_k2f_upload(ogn_ctx *ctx, ogn_uploader *up, const std::vector<float> &table);
int ogn_k2f_launch(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, ogn_window w, const float *cube_fsf,
                   int pitch, const uint8_t *mask, float *correl, float *correl_min, uint8_t *profile, float *maxmap,
                   float *minmap);
int ogn_wait_readers(ogn_ctx *ctx, cudaStream_t stream, const void *buf);
int ogn_tglr_init_maps(ogn_ctx *ctx, cudaStream_t stream, float *d_maxmap, float *d_minmap, size_t img);
int ogn_extrema_run(ogn_ctx *ctx, const float *a, const float *b, const uint8_t *mask, int nz, int ny, int nx,
                    ogn_window owned, ogn_place place, int sz, int sy, int sx, float *dense_max, float *dense_min,
                    int64_t *max_index, float *max_value, int64_t *min_index, float *min_value, int64_t capacity,
                    int64_t *counts);
