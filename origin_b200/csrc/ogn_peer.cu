// Multi-GPU: gather of the owned product tiles into one rank's full cube over NVLink peer memory.
//
// The reference has no distributed code; north_star asks for "correl gathered to rank 0".  Instead
// of packing tiles, sending them with NCCL and unpacking them on rank 0, every rank writes the
// voxels it owns straight into rank 0's [nz][gny][gnx] cube: rank 0 allocates the cube with
// ogn_peer_alloc and publishes its CUDA IPC handle, the other ranks map it with ogn_peer_open, and
// ogn_scatter_tile enqueues one strided 3-D peer-to-peer copy (copy engine by default, an SM copy
// kernel on request) whose writes travel over NVLink / NVSwitch; on rank 0 itself the same call
// writes local memory.  The copy runs on the context's peer stream behind the work already queued
// on the main stream, so the transfer of step i overlaps the kernels of step i+1; the library makes
// a later TGLR call that overwrites the source buffer wait for the copy that still reads it.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ogn_common.cuh"

namespace {

// dst[z][gy0 + y][gx0 + x] = src[z][oy0 + y][ox0 + x] for the owned window; VEC floats per thread
template <int VEC>
__global__ void __launch_bounds__(256)
scatter_tile_kernel(const float *__restrict__ src, int ny, int nx, int oy0, int ox0, int oh, int ow,
                    float *__restrict__ dst, int gny, int gnx, int gy0, int gx0, long long nrow_total) {
    const int per_row = ow / VEC;
    const long long total = nrow_total * per_row;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int UN = 4;   // independent loads in flight per thread before the first store
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += UN * stride) {
        float4 v[UN];
        float *d[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = i0 + u * stride;
            d[u] = nullptr;
            if (i < total) {
                const long long rowid = i / per_row;            // z * oh + y
                const int xq = (int)(i - rowid * per_row) * VEC;
                const int z = (int)(rowid / oh), y = (int)(rowid - (long long)z * oh);
                const float *s = src + ((size_t)z * ny + oy0 + y) * nx + ox0 + xq;
                d[u] = dst + ((size_t)z * gny + gy0 + y) * gnx + gx0 + xq;
                if (VEC == 4) v[u] = __ldg(reinterpret_cast<const float4 *>(s));
                else v[u].x = __ldg(s);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (d[u]) {
                if (VEC == 4) *reinterpret_cast<float4 *>(d[u]) = v[u];
                else *d[u] = v[u].x;
            }
    }
}

// device-side wait on the peer stream (one thread): staggers the ranks' copies into the shared destination
__global__ void delay_kernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(2000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

// Consecutive scatters alternate between two streams: with a staggered start (ogn_peer_set_delay) the wait of
// copy i+1 would otherwise queue behind copy i and every copy would start later than the one before.
int peer_stream(ogn_ctx *ctx, cudaStream_t *out) {
    cudaStream_t &ps = ctx->peer_streams[ctx->peer_turn & 1];
    ctx->peer_turn++;
    if (!ps) OGN_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    ctx->peer_stream = ps;
    *out = ps;
    return OGN_OK;
}
int sync_peer_streams(ogn_ctx *ctx) {
    for (auto ps : ctx->peer_streams)
        if (ps) OGN_CUDA(cudaStreamSynchronize(ps));
    return OGN_OK;
}

}  // namespace

extern "C" int ogn_peer_alloc(ogn_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char *handle) {
    if (!ctx || !dev_ptr || !handle || bytes == 0) return ctx ? ogn_fail(ctx, OGN_ERR_ARG, "ogn_peer_alloc: bad arguments") : OGN_ERR_ARG;
    OGN_CUDA(cudaSetDevice(ctx->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);  // a whole allocation of its own: the IPC handle maps exactly this buffer
    if (e != cudaSuccess) return ogn_fail(ctx, OGN_ERR_NOMEM, "cudaMalloc(%zu) for a peer buffer failed: %s", bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return ogn_fail(ctx, OGN_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle, &h, sizeof(h));
    ctx->peer_owned.push_back(p);
    *dev_ptr = p;
    return OGN_OK;
}

extern "C" int ogn_peer_free(ogn_ctx *ctx, void *dev_ptr) {
    if (!ctx) return OGN_ERR_ARG;
    auto it = std::find(ctx->peer_owned.begin(), ctx->peer_owned.end(), dev_ptr);
    if (it == ctx->peer_owned.end()) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_peer_free: not a buffer of ogn_peer_alloc");
    OGN_CUDA(cudaSetDevice(ctx->device));
    OGN_CUDA(cudaDeviceSynchronize());
    OGN_CUDA(cudaFree(dev_ptr));
    ctx->peer_owned.erase(it);
    return OGN_OK;
}

extern "C" int ogn_peer_open(ogn_ctx *ctx, const unsigned char *handle, void **dev_ptr) {
    if (!ctx || !handle || !dev_ptr) return ctx ? ogn_fail(ctx, OGN_ERR_ARG, "ogn_peer_open: bad arguments") : OGN_ERR_ARG;
    OGN_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ogn_fail(ctx, OGN_ERR_CUDA, "cudaIpcOpenMemHandle failed: %s (is peer access between the GPUs available?)",
                        cudaGetErrorString(e));
    }
    ctx->peer_mapped.push_back(p);
    *dev_ptr = p;
    return OGN_OK;
}

extern "C" int ogn_peer_close(ogn_ctx *ctx, void *dev_ptr) {
    if (!ctx) return OGN_ERR_ARG;
    auto it = std::find(ctx->peer_mapped.begin(), ctx->peer_mapped.end(), dev_ptr);
    if (it == ctx->peer_mapped.end()) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_peer_close: not a mapping of ogn_peer_open");
    OGN_CUDA(cudaSetDevice(ctx->device));
    OGN_TRY(sync_peer_streams(ctx));
    OGN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    ctx->peer_mapped.erase(it);
    return OGN_OK;
}

extern "C" int ogn_scatter_tile(ogn_ctx *ctx, const float *src, int nz, int ny, int nx, const int *tile, float *dst) {
    if (!ctx) return OGN_ERR_ARG;
    if (!src || !dst || !tile || nz <= 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_scatter_tile: bad arguments");
    const int gny = tile[0], gnx = tile[1], gy0 = tile[2], gx0 = tile[3];
    const int oy0 = tile[4], oy1 = tile[5], ox0 = tile[6], ox1 = tile[7];
    if (oy0 < 0 || ox0 < 0 || oy1 > ny || ox1 > nx || oy0 >= oy1 || ox0 >= ox1 || gy0 < 0 || gx0 < 0 ||
        gy0 + ny > gny || gx0 + nx > gnx)
        return ogn_fail(ctx, OGN_ERR_ARG, "ogn_scatter_tile: tile does not fit the %dx%d field", gny, gnx);
    if (!ogn_is_device_ptr(src)) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_scatter_tile: src must be device memory");
    OGN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t ps;
    OGN_TRY(peer_stream(ctx, &ps));
    // the copy starts behind the producer of src on the main stream ...
    if (!ctx->peer_ev_begin) OGN_CUDA(cudaEventCreateWithFlags(&ctx->peer_ev_begin, cudaEventDisableTiming));
    OGN_CUDA(cudaEventRecord(ctx->peer_ev_begin, ctx->stream));
    OGN_CUDA(cudaStreamWaitEvent(ps, ctx->peer_ev_begin, 0));
    if (ctx->peer_delay_us > 0) {
        delay_kernel<<<1, 1, 0, ps>>>((unsigned long long)ctx->peer_delay_us * 1000ull);
        OGN_LAUNCH_CHECK("delay_kernel");
    }
    const int oh = oy1 - oy0, ow = ox1 - ox0;
    const long long nrow = (long long)nz * oh;
    const bool vec = ow % 4 == 0 && nx % 4 == 0 && gnx % 4 == 0 && ox0 % 4 == 0 && (gx0 + ox0) % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const long long items = nrow * (vec ? ow / 4 : ow);
    // a modest grid: NVLink is saturated by a few SMs' worth of stores, the rest stay with the main stream
    static const int max_blocks = getenv("OGN_SCATTER_BLOCKS") ? atoi(getenv("OGN_SCATTER_BLOCKS")) : ctx->sm_count;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, max_blocks));
    // Default: the copy engine moves the tile (one strided 3-D copy; measured 741-782 GB/s into rank 0
    // from 7 peers, and no SM is taken from the main stream's kernels).  OGN_SCATTER_KERNEL=1 selects the
    // SM copy kernel instead (703 GB/s with 148 blocks, but it competes with K1 for registers).
    static const bool use_dma = getenv("OGN_SCATTER_KERNEL") == nullptr;
    // stage timing (ogn_timing_enable): the copy's own duration on the peer stream
    ogn_timing_entry t_copy;
    if (ctx->timing) {
        t_copy.name = "peer_scatter";
        cudaEventCreate(&t_copy.start);
        cudaEventCreate(&t_copy.stop);
        cudaEventRecord(t_copy.start, ps);
    }
    if (use_dma) {
        cudaMemcpy3DParms p3 = {};
        p3.srcPtr = make_cudaPitchedPtr(const_cast<float *>(src), (size_t)nx * 4, nx, ny);
        p3.srcPos = make_cudaPos((size_t)ox0 * 4, oy0, 0);
        p3.dstPtr = make_cudaPitchedPtr(dst, (size_t)gnx * 4, gnx, gny);
        p3.dstPos = make_cudaPos((size_t)(gx0 + ox0) * 4, gy0 + oy0, 0);
        p3.extent = make_cudaExtent((size_t)ow * 4, oh, nz);
        p3.kind = cudaMemcpyDefault;
        OGN_CUDA(cudaMemcpy3DAsync(&p3, ps));
    } else if (vec) {
        // a copy block holds no shared memory; without a stated preference the driver configures an idle SM it
        // lands on for maximum L1, and the TGLR kernels (38-48 KB of shared memory per block) then cannot place a
        // block there until the copy block has left - measured as a full serialisation of the next step behind the
        // copy.  Asking for the largest shared-memory carve-out keeps the SM usable for them.
        static bool carve_set = false;
        if (!carve_set) {
            cudaFuncSetAttribute(scatter_tile_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(scatter_tile_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve_set = true;
        }
        scatter_tile_kernel<4><<<blocks, 256, 0, ps>>>(src, ny, nx, oy0, ox0, oh, ow, dst, gny, gnx, gy0 + oy0, gx0 + ox0, nrow);
    } else
        scatter_tile_kernel<1><<<blocks, 256, 0, ps>>>(src, ny, nx, oy0, ox0, oh, ow, dst, gny, gnx, gy0 + oy0, gx0 + ox0, nrow);
    if (!use_dma) OGN_LAUNCH_CHECK("scatter_tile_kernel");
    if (ctx->timing) {
        cudaEventRecord(t_copy.stop, ps);
        ctx->timings.push_back(t_copy);
    }
    // ... and whoever overwrites src next on the main stream waits for it (ogn_wait_readers)
    cudaEvent_t done = nullptr;
    auto it = ctx->readers.find(src);
    if (it != ctx->readers.end()) done = it->second;
    else {
        // callers that hand a fresh buffer every step would grow the map without bound: entries whose copy
        // has completed protect nothing any more
        if (ctx->readers.size() >= 16)
            for (auto r = ctx->readers.begin(); r != ctx->readers.end();)
                if (cudaEventQuery(r->second) == cudaSuccess) {
                    cudaEventDestroy(r->second);
                    r = ctx->readers.erase(r);
                } else ++r;
        OGN_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        ctx->readers[src] = done;
    }
    OGN_CUDA(cudaEventRecord(done, ps));
    return OGN_OK;
}

// The next ogn_step05_tile call on this context also stores the window it owns of `correl` into `dst`
// (the [nz][gny][gnx] cube of ogn_peer_alloc on THIS device): the owner of the gathered cube needs no
// copy of its own tile.  One-shot; pass NULL to cancel.
extern "C" int ogn_set_local_gather(ogn_ctx *ctx, float *dst) {
    if (!ctx) return OGN_ERR_ARG;
    if (dst && !ogn_is_device_ptr(dst)) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_set_local_gather: dst must be device memory");
    ctx->local_gather = dst;
    ctx->local_gather_is_peer = false;
    // a buffer of ogn_peer_open lives on another GPU (cudaPointerGetAttributes reports IPC mappings as local)
    if (dst) ctx->local_gather_is_peer = std::find(ctx->peer_mapped.begin(), ctx->peer_mapped.end(), (void *)dst) != ctx->peer_mapped.end();
    return OGN_OK;
}

// Main stream waits for every scatter enqueued so far (no host synchronisation).
extern "C" int ogn_peer_join(ogn_ctx *ctx) {
    if (!ctx) return OGN_ERR_ARG;
    if (!ctx->peer_stream) return OGN_OK;
    OGN_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->peer_ev_end) OGN_CUDA(cudaEventCreateWithFlags(&ctx->peer_ev_end, cudaEventDisableTiming));
    if (!ctx->peer_ev_end2) OGN_CUDA(cudaEventCreateWithFlags(&ctx->peer_ev_end2, cudaEventDisableTiming));
    cudaEvent_t evs[2] = {ctx->peer_ev_end, ctx->peer_ev_end2};
    for (int i = 0; i < 2; ++i)
        if (ctx->peer_streams[i]) {
            OGN_CUDA(cudaEventRecord(evs[i], ctx->peer_streams[i]));
            OGN_CUDA(cudaStreamWaitEvent(ctx->stream, evs[i], 0));
        }
    return OGN_OK;
}

// Host waits for every scatter enqueued so far.  The caller still needs a barrier across ranks
// before the owner of the destination reads it.
extern "C" int ogn_peer_sync(ogn_ctx *ctx) {
    if (!ctx) return OGN_ERR_ARG;
    if (!ctx->peer_stream) return OGN_OK;
    OGN_CUDA(cudaSetDevice(ctx->device));
    return sync_peer_streams(ctx);
}

// A kernel on `stream` is about to overwrite `buf`: wait for the scatter that may still read it.
int ogn_wait_readers(ogn_ctx *ctx, cudaStream_t stream, const void *buf) {
    if (!buf || ctx->readers.empty()) return OGN_OK;
    auto it = ctx->readers.find(buf);
    if (it != ctx->readers.end()) OGN_CUDA(cudaStreamWaitEvent(stream, it->second, 0));
    return OGN_OK;
}

// Every following ogn_scatter_tile of this context waits `microseconds` on the device (peer stream) before its copy.
// All ranks push into ONE destination GPU, whose NVLink ingress they share: pushing at the same time, every source
// keeps its memory system full of stalled remote stores for the whole gather, and its own latency-bound kernels
// slow down several times (measured: a 0.02 ms kernel with atomics took 0.75 ms).  Staggered by rank - each source
// alone on the link for its turn - the gather lasts as long but a source is congested only during its own turn.
extern "C" int ogn_peer_set_delay(ogn_ctx *ctx, int microseconds) {
    if (!ctx || microseconds < 0) return ctx ? ogn_fail(ctx, OGN_ERR_ARG, "ogn_peer_set_delay: negative delay") : OGN_ERR_ARG;
    ctx->peer_delay_us = microseconds;
    return OGN_OK;
}
