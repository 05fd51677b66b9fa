// libogn context, scratch arena and staging helpers.
#include <stdarg.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "ogn_common.cuh"
#include "ogn_tma.cuh"

static std::string g_create_error;

int ogn_fail(ogn_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

void ogn_host_trace(const char *label) {
    static const bool on = getenv("OGN_HOST_TRACE") != nullptr;
    if (!on) return;
    static timespec prev = {0, 0};
    timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    const double ms = (now.tv_sec - prev.tv_sec) * 1e3 + (now.tv_nsec - prev.tv_nsec) * 1e-6;
    fprintf(stderr, "[ogn %d] %-28s +%.3f ms\n", (int)getpid(), label, prev.tv_sec ? ms : 0.0);
    prev = now;
}

extern "C" int ogn_version(void) { return OGN_VERSION; }

extern "C" int ogn_create(int device, void *stream, ogn_ctx **out) {
    ogn_ctx *ctx = nullptr;
    if (!out) return ogn_fail(nullptr, OGN_ERR_ARG, "ogn_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ogn_fail(nullptr, OGN_ERR_CUDA,
                        "ogn_create: no CUDA device (%s); libogn has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev)
        return ogn_fail(nullptr, OGN_ERR_ARG, "ogn_create: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return ogn_fail(nullptr, OGN_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return ogn_fail(nullptr, OGN_ERR_UNSUPPORTED,
                        "ogn_create: device %d is sm_%d%d; libogn is built for sm_100a (B200) only",
                        device, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return ogn_fail(nullptr, OGN_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new ogn_ctx();
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return OGN_OK;
}

extern "C" int ogn_trim(ogn_ctx *ctx) {
    if (!ctx) return OGN_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->bufs)
        if (kv.second.p) cudaFree(kv.second.p);
    ctx->bufs.clear();
    for (auto &kv : ctx->pins)
        if (kv.second.p) cudaFreeHost(kv.second.p);
    ctx->pins.clear();
    ctx->prep = ogn_prep_state();
    ctx->dct_tab_nz = ctx->dct_tab_M = 0;
    for (auto e : ctx->events) cudaEventDestroy(e);
    ctx->events.clear();
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    ctx->h2d_stream = ctx->d2h_stream = nullptr;
    for (auto &ps : ctx->peer_streams)
        if (ps) {
            cudaStreamSynchronize(ps);
            cudaStreamDestroy(ps);
            ps = nullptr;
        }
    ctx->peer_stream = nullptr;
    for (auto &kv : ctx->readers) cudaEventDestroy(kv.second);
    ctx->readers.clear();
    if (ctx->peer_ev_begin) cudaEventDestroy(ctx->peer_ev_begin);
    if (ctx->peer_ev_end) cudaEventDestroy(ctx->peer_ev_end);
    if (ctx->peer_ev_end2) cudaEventDestroy(ctx->peer_ev_end2);
    ctx->peer_ev_begin = ctx->peer_ev_end = ctx->peer_ev_end2 = nullptr;
    return OGN_OK;
}

extern "C" void ogn_destroy(ogn_ctx *ctx) {
    if (!ctx) return;
    ogn_tglr_forget(ctx);
    ogn_trim(ctx);
    if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
    if (ctx->res_h) cudaFreeHost(ctx->res_h);
    if (ctx->stage_ev) cudaEventDestroy(ctx->stage_ev);
    for (void *p : ctx->peer_mapped) cudaIpcCloseMemHandle(p);
    for (void *p : ctx->peer_owned) cudaFree(p);
    delete ctx;
}

extern "C" const char *ogn_last_error(const ogn_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int ogn_synchronize(ogn_ctx *ctx) {
    if (!ctx) return OGN_ERR_ARG;
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->host_output_pending = false;
    return OGN_OK;
}

extern "C" int ogn_timing_enable(ogn_ctx *ctx, int on) {
    if (!ctx) return OGN_ERR_ARG;
    ctx->timing = on != 0;
    return OGN_OK;
}

// "name:ms;name:ms;..." for every timed stage since the last report; clears the list.
extern "C" int ogn_timing_report(ogn_ctx *ctx, char *buf, size_t size) {
    if (!ctx || !buf || size == 0) return OGN_ERR_ARG;
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto ps : ctx->peer_streams)
        if (ps) OGN_CUDA(cudaStreamSynchronize(ps));   // peer_scatter entries live there
    std::string out;
    static const bool with_offsets = getenv("OGN_TIMING_OFFSETS") != nullptr;   // "name@start_ms:duration_ms" (timelines)
    cudaEvent_t first = ctx->timings.empty() ? nullptr : ctx->timings.front().start;
    for (auto &e : ctx->timings) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.start, e.stop) == cudaSuccess) {
            char tmp[200];
            float off = 0.f;
            if (with_offsets && cudaEventElapsedTime(&off, first, e.start) == cudaSuccess)
                snprintf(tmp, sizeof(tmp), "%s@%.4f:%.6f;", e.name.c_str(), off, ms);
            else
                snprintf(tmp, sizeof(tmp), "%s:%.6f;", e.name.c_str(), ms);
            cudaGetLastError();
            out += tmp;
        } else {
            cudaGetLastError();
        }
    }
    for (auto &e : ctx->timings) {   // after the loop: the first start event is the origin of the offsets
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    ctx->timings.clear();
    snprintf(buf, size, "%s", out.c_str());
    return OGN_OK;
}

extern "C" int ogn_variants(ogn_ctx *ctx, char *buf, size_t size) {
    if (!ctx || !buf || size == 0) return OGN_ERR_ARG;
    std::string out;
    for (auto &kv : ctx->variants) out += kv.first + "=" + kv.second + ";";
    snprintf(buf, size, "%s", out.c_str());
    return OGN_OK;
}

extern "C" int ogn_fsf_folded(ogn_ctx *ctx, int *folded) {
    if (!ctx || !folded) return OGN_ERR_ARG;
    auto it = ctx->bufs.find("fsf_asym");
    if (it == ctx->bufs.end() || !it->second.p) return ogn_fail(ctx, OGN_ERR_ARG, "no TGLR call has run on this context");
    int asym = 1;
    OGN_CUDA(cudaMemcpyAsync(&asym, it->second.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    *folded = asym == 0;
    return OGN_OK;
}

extern "C" int64_t ogn_launch_count(const ogn_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ogn_host_alloc(size_t bytes, void **out) {
    if (!out) return OGN_ERR_ARG;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return ogn_fail(nullptr, OGN_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return OGN_OK;
}

extern "C" int ogn_host_free(void *ptr) {
    return cudaFreeHost(ptr) == cudaSuccess ? OGN_OK : OGN_ERR_CUDA;
}

// ---- zero-copy staging ------------------------------------------------------------------------
__global__ void upload_kernel(const unsigned char *__restrict__ src, const ogn_upload_table tab) {
    const ogn_upload_item it = tab.item[blockIdx.x];
    const unsigned *s = reinterpret_cast<const unsigned *>(src + it.off);
    unsigned *d = reinterpret_cast<unsigned *>(it.dst);
    for (unsigned i = threadIdx.x; i < it.bytes / 4; i += blockDim.x) d[i] = s[i];
}

__global__ void fill_words_kernel(unsigned *__restrict__ p, unsigned word, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = word;
}

int ogn_fill_words(ogn_ctx *ctx, cudaStream_t stream, void *p, unsigned word, size_t bytes) {
    if (bytes == 0) return OGN_OK;
    const size_t n = bytes / 4;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    fill_words_kernel<<<blocks, 256, 0, stream>>>(static_cast<unsigned *>(p), word, n);
    OGN_LAUNCH_CHECK("fill_words_kernel");
    return OGN_OK;
}

static int stage_reserve(ogn_ctx *ctx, size_t bytes) {
    if (ctx->stage_cap >= bytes) return OGN_OK;
    if (ctx->stage_h) {
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
        OGN_CUDA(cudaFreeHost(ctx->stage_h));
        ctx->stage_h = ctx->stage_d = nullptr;
        ctx->stage_cap = 0;
    }
    const size_t cap = std::max<size_t>(bytes, 256 * 1024);
    void *hp = nullptr, *dp = nullptr;
    cudaError_t e = cudaHostAlloc(&hp, cap, cudaHostAllocMapped);
    if (e != cudaSuccess) return ogn_fail(ctx, OGN_ERR_NOMEM, "cudaHostAlloc(mapped, %zu) failed: %s", cap, cudaGetErrorString(e));
    OGN_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
    ctx->stage_h = static_cast<unsigned char *>(hp);
    ctx->stage_d = static_cast<unsigned char *>(dp);
    ctx->stage_cap = cap;
    return OGN_OK;
}

int ogn_uploader::add(void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return OGN_OK;
    if (bytes % 4) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_uploader: item size must be a multiple of 4");
    if (used == 0 && ctx->stage_pending) {  // the previous batch may still be read by its kernel
        OGN_CUDA(cudaEventSynchronize(ctx->stage_ev));
        ctx->stage_pending = false;
    }
    size_t off = (used + 15) / 16 * 16;
    if (off + bytes > ctx->stage_cap) {
        if (tab.n) {   // send what is staged, wait until it has been read, start over at offset 0
            OGN_TRY(flush(ctx->stream));
            OGN_CUDA(cudaEventSynchronize(ctx->stage_ev));
            ctx->stage_pending = false;
            off = 0;
        }
        if (bytes > ctx->stage_cap) OGN_TRY(stage_reserve(ctx, bytes));
    }
    if (tab.n >= OGN_UPLOAD_MAX) {   // table full: same
        OGN_TRY(flush(ctx->stream));
        OGN_CUDA(cudaEventSynchronize(ctx->stage_ev));
        ctx->stage_pending = false;
        off = 0;
    }
    memcpy(ctx->stage_h + off, src, bytes);
    tab.item[tab.n++] = ogn_upload_item{dst, (unsigned)off, (unsigned)bytes};
    used = off + bytes;
    return OGN_OK;
}

int ogn_uploader::flush(cudaStream_t stream) {
    if (tab.n == 0) return OGN_OK;
    upload_kernel<<<tab.n, 256, 0, stream>>>(ctx->stage_d, tab);
    OGN_LAUNCH_CHECK("upload_kernel");
    if (!ctx->stage_ev) OGN_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev, cudaEventDisableTiming));
    OGN_CUDA(cudaEventRecord(ctx->stage_ev, stream));
    ctx->stage_pending = true;
    tab.n = 0;
    used = 0;
    return OGN_OK;
}

int ogn_result_slots(ogn_ctx *ctx, int64_t **dev, int64_t **host) {
    if (!ctx->res_h) {
        void *hp = nullptr, *dp = nullptr;
        cudaError_t e = cudaHostAlloc(&hp, 32 * sizeof(int64_t), cudaHostAllocMapped);
        if (e != cudaSuccess) return ogn_fail(ctx, OGN_ERR_NOMEM, "cudaHostAlloc(mapped) failed: %s", cudaGetErrorString(e));
        OGN_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
        ctx->res_h = static_cast<int64_t *>(hp);
        ctx->res_d = static_cast<int64_t *>(dp);
    }
    *dev = ctx->res_d;
    *host = ctx->res_h;
    return OGN_OK;
}

int ogn_scratch(ogn_ctx *ctx, const char *name, size_t bytes, void **out) {
    ogn_buf &b = ctx->bufs[name];
    if (bytes == 0) bytes = 16;
    if (b.cap < bytes) {
        if (b.p) {
            // pending work on the stream may still use the old buffer
            OGN_CUDA(cudaStreamSynchronize(ctx->stream));
            OGN_CUDA(cudaFree(b.p));
            b.p = nullptr;
            b.cap = 0;
        }
        size_t cap = (bytes + 255) / 256 * 256;
        cudaError_t e = cudaMalloc(&b.p, cap);
        if (e != cudaSuccess) {
            b.p = nullptr;
            return ogn_fail(ctx, OGN_ERR_NOMEM, "cudaMalloc(%zu bytes) for '%s' failed: %s", cap, name,
                            cudaGetErrorString(e));
        }
        b.cap = cap;
    }
    *out = b.p;
    return OGN_OK;
}

int ogn_pinned(ogn_ctx *ctx, const char *name, size_t bytes, void **out) {
    ogn_buf &b = ctx->pins[name];
    if (bytes == 0) bytes = 16;
    if (b.cap < bytes) {
        if (b.p) {
            OGN_CUDA(cudaStreamSynchronize(ctx->stream));
            OGN_CUDA(cudaFreeHost(b.p));
            b.p = nullptr;
            b.cap = 0;
        }
        cudaError_t e = cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            b.p = nullptr;
            return ogn_fail(ctx, OGN_ERR_NOMEM, "cudaHostAlloc(%zu bytes) for '%s' failed: %s", bytes, name,
                            cudaGetErrorString(e));
        }
        b.cap = bytes;
    }
    *out = b.p;
    return OGN_OK;
}

bool ogn_is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int ogn_input(ogn_ctx *ctx, const char *name, const void *p, size_t bytes, const void **dev) {
    if (!p) return ogn_fail(ctx, OGN_ERR_ARG, "input '%s' is NULL", name);
    if (ogn_is_device_ptr(p)) {
        *dev = p;
        return OGN_OK;
    }
    void *d = nullptr;
    OGN_TRY(ogn_scratch(ctx, name, bytes, &d));
    OGN_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    return OGN_OK;
}

int ogn_output(ogn_ctx *ctx, const char *name, void *p, size_t bytes, void **dev) {
    if (p && ogn_is_device_ptr(p)) {
        *dev = p;
        return OGN_OK;
    }
    return ogn_scratch(ctx, name, bytes, dev);
}

int ogn_output_commit(ogn_ctx *ctx, void *p, const void *dev, size_t bytes) {
    if (!p || p == dev) return OGN_OK;
    OGN_CUDA(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->host_output_pending = true;
    return OGN_OK;
}

int ogn_finish_call(ogn_ctx *ctx) {
    if (ctx->host_output_pending) {
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->host_output_pending = false;
    }
    return OGN_OK;
}

// ---------------------------------------------------------------------------

template <typename S, typename D>
__global__ void convert_kernel(const S *__restrict__ src, D *__restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = static_cast<D>(src[i]);
}

int ogn_convert_f64_to_f32(ogn_ctx *ctx, const double *src, float *dst, size_t n) {
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
    if (blocks == 0) return OGN_OK;
    convert_kernel<double, float><<<blocks, 256, 0, ctx->stream>>>(src, dst, n);
    OGN_LAUNCH_CHECK("convert_f64_f32");
    return OGN_OK;
}

int ogn_convert_f32_to_f64(ogn_ctx *ctx, const float *src, double *dst, size_t n) {
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
    if (blocks == 0) return OGN_OK;
    convert_kernel<float, double><<<blocks, 256, 0, ctx->stream>>>(src, dst, n);
    OGN_LAUNCH_CHECK("convert_f32_f64");
    return OGN_OK;
}

int ogn_input_cube_f32(ogn_ctx *ctx, const char *name, const void *p, int dtype, size_t n,
                       const float **dev) {
    if (dtype == OGN_F32) {
        const void *d = nullptr;
        OGN_TRY(ogn_input(ctx, name, p, n * sizeof(float), &d));
        *dev = static_cast<const float *>(d);
        return OGN_OK;
    }
    if (dtype != OGN_F64) return ogn_fail(ctx, OGN_ERR_ARG, "unknown dtype %d for '%s'", dtype, name);
    std::string stage = std::string(name) + ".f64";
    const void *d64 = nullptr;
    OGN_TRY(ogn_input(ctx, stage.c_str(), p, n * sizeof(double), &d64));
    float *d32 = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, name, n, &d32));
    OGN_TRY(ogn_convert_f64_to_f32(ctx, static_cast<const double *>(d64), d32, n));
    *dev = d32;
    return OGN_OK;
}

// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int get_encoder(ogn_ctx *ctx, PFN_encodeTiled *out) {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return ogn_fail(ctx, OGN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    *out = encode;
    return OGN_OK;
}

// 2-D tiled map over a [rows][cols] array of 1- or 4-byte elements (row stride in bytes, a multiple of 16);
// out-of-range elements are zero-filled.
int ogn_make_map_2d(ogn_ctx *ctx, CUtensorMap *map, const void *base, int elem_bytes, uint64_t rows, uint64_t cols,
                    uint64_t row_stride_bytes, int box_cols, int box_rows) {
    PFN_encodeTiled encode = nullptr;
    OGN_TRY(get_encoder(ctx, &encode));
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = encode(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ogn_fail(ctx, OGN_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed with CUresult %d", (int)r);
    return OGN_OK;
}

int ogn_make_tile_map(ogn_ctx *ctx, CUtensorMap *map, const float *base, int nz, int ny, int nx, int pitch,
                      int box_x, int box_y, int box_z, bool nan_fill) {
    PFN_encodeTiled encode = nullptr;
    OGN_TRY(get_encoder(ctx, &encode));
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)ny * pitch * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, (cuuint32_t)box_z};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ogn_fail(ctx, OGN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return OGN_OK;
}

