// 3-D local extrema (compute_local_max, lib_origin.py:1220-1256), ordered compaction,
// step06 purity counts (lib_origin.py:1424-1449) and step07 thresholding (steps.py:956-974).
//
//   K3  local_extrema3_tma_kernel  3x3x3: TMA-staged (NaN out-of-bounds fill), float4 lanes, one warp
//                                  per image row, no block barrier; flag words = one bit per voxel
//       local_extrema3_kernel      3x3x3 scalar fallback (nx % 4 != 0), local_extrema_kernel any odd window
//   K3b flag_totals / flag_offsets / flag_scatter   order-preserving compaction of both flag arrays
//                                  into (linear index, value) lists = np.where order, 3 launches
//   K4  purity_stats / purity_counts kernels over the lists
//   K6  threshold_select       order-preserving selection value > threshold
#include <math.h>
#include <stdlib.h>

#include "ogn_common.cuh"
#include "ogn_tma.cuh"

// A maximum filter with scipy's 'reflect' boundary only ever sees in-range
// samples (the reflected positions lie inside the clipped window), so the
// out-of-range part of the window is simply skipped.
__global__ void local_extrema_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                     const uint8_t *__restrict__ mask, int nz, int ny, int nx, int rz, int ry,
                                     int rx, int oy0, int oy1, int ox0a, int ox0, int ox1,
                                     float *__restrict__ dense_max, float *__restrict__ dense_min,
                                     uint32_t *__restrict__ flag_max, uint32_t *__restrict__ flag_min, int nxw) {
    const int x = ox0a + blockIdx.x * 32 + threadIdx.x;
    const int y = oy0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (y >= oy1) return;  // whole warp (a warp is one row of 32 x)
    bool keep_a = false, keep_b = false;
    float va = 0.f, vb = 0.f;
    size_t idx = 0;
    if (x >= ox0 && x < ox1) {
        idx = ((size_t)z * ny + y) * nx + x;
        va = a[idx];
        vb = b[idx];
        float ma = -INFINITY, mb = INFINITY;
        const int z0 = max(0, z - rz), z1 = min(nz - 1, z + rz);
        const int y0 = max(0, y - ry), y1 = min(ny - 1, y + ry);
        const int x0 = max(0, x - rx), x1 = min(nx - 1, x + rx);
        for (int zz = z0; zz <= z1; ++zz)
            for (int yy = y0; yy <= y1; ++yy) {
                const size_t row = ((size_t)zz * ny + yy) * nx;
                for (int xx = x0; xx <= x1; ++xx) {
                    ma = fmaxf(ma, a[row + xx]);
                    mb = fminf(mb, b[row + xx]);
                }
            }
        const bool free_voxel = !(mask && mask[idx]);
        keep_a = free_voxel && va == ma;
        keep_b = free_voxel && vb == mb;
        if (dense_max) dense_max[idx] = keep_a ? va : 0.f;
        if (dense_min) dense_min[idx] = keep_b ? -vb : 0.f;
    }
    const uint32_t wa = __ballot_sync(0xffffffffu, keep_a);
    const uint32_t wb = __ballot_sync(0xffffffffu, keep_b);
    if (threadIdx.x == 0) {
        const size_t w = ((size_t)z * (oy1 - oy0) + (y - oy0)) * nxw + blockIdx.x;
        flag_max[w] = wa;
        flag_min[w] = wb;
    }
}

// 3 x 3 x 3 window (the default, steps.py:427,756): a block owns a 32 x 16 spatial tile and
// walks a run of wavelength planes with 18 warps, one per tile row including the two halo
// rows.  Per plane every thread loads its own voxel of both arrays (one coalesced 128-byte
// row per warp; lanes 0 and 31 also fetch the neighbouring tile's column), takes the 3-wide
// max / min along x with two shuffles, the 3-wide max / min along y through a
// double-buffered shared-memory tile (one barrier per plane), and keeps the plane extrema
// of z-1, z, z+1 in registers, so each voxel is read once per array from L2/HBM.  The loads
// of plane p+1 are issued before plane p is reduced (software pipeline).
constexpr int EX_TY = 16;
constexpr int EX_CZ = 64;

struct ExPlane {
    float va, ea, vb, eb;  // own voxel and (lanes 0 / 31) the neighbouring tile's voxel
    uint32_t m;
};

__global__ void __launch_bounds__(32 * (EX_TY + 2), 2)
local_extrema3_kernel(const float *__restrict__ a, const float *__restrict__ b, const uint8_t *__restrict__ mask,
                      int nz, int ny, int nx, int oy0, int oy1, int ox0a, int ox0, int ox1,
                      float *__restrict__ dense_max, float *__restrict__ dense_min,
                      uint32_t *__restrict__ flag_max, uint32_t *__restrict__ flag_min, int nxw) {
    __shared__ float sa[2][EX_TY + 2][32];
    __shared__ float sb[2][EX_TY + 2][32];
    const int lane = threadIdx.x, row = threadIdx.y;  // row 0 and EX_TY+1 are halo rows
    const int x0 = ox0a + blockIdx.x * 32, x = x0 + lane;
    const int y = oy0 + blockIdx.y * EX_TY + row - 1;
    const int zc0 = blockIdx.z * EX_CZ, zc1 = min(nz, zc0 + EX_CZ);
    const bool row_ok = y >= 0 && y < ny;
    const bool c_ok = row_ok && x < nx;
    const bool e_ok = row_ok && (lane == 0 ? x0 > 0 : (lane == 31 && x0 + 32 < nx));
    const bool out_row = row >= 1 && row <= EX_TY && y < oy1;
    const bool out_col = x >= ox0 && x < ox1;
    const bool is_l = lane == 0, is_r = lane == 31;
    const size_t plane = (size_t)ny * nx;
    const size_t oc = c_ok ? (size_t)y * nx + x : 0;
    const size_t oe = e_ok ? (size_t)y * nx + (is_l ? x0 - 1 : x0 + 32) : 0;

    // running pointers at the next plane to load; everything below is pointer bumps, no index math
    int pp = max(zc0 - 1, 0);
    const float *ac = a + (size_t)pp * plane + oc, *ae = a + (size_t)pp * plane + oe;
    const float *bc = b + (size_t)pp * plane + oc, *be = b + (size_t)pp * plane + oe;
    const uint8_t *mc = mask ? mask + (size_t)pp * plane + oc : nullptr;
    const bool has_mask = mask != nullptr;
    float *dmax_p = dense_max ? dense_max + (size_t)zc0 * plane + oc : nullptr;
    float *dmin_p = dense_min ? dense_min + (size_t)zc0 * plane + oc : nullptr;
    const size_t wstride = (size_t)(oy1 - oy0) * nxw;
    size_t widx = ((size_t)zc0 * (oy1 - oy0) + (y - oy0)) * nxw + blockIdx.x;

    auto load_next = [&](bool valid) {
        ExPlane pl;
        pl.va = pl.ea = -INFINITY;
        pl.vb = pl.eb = INFINITY;
        pl.m = 0;
        if (valid) {  // warp-uniform
            if (c_ok) { pl.va = __ldg(ac); pl.vb = __ldg(bc); if (has_mask) pl.m = *mc; }
            if (e_ok) { pl.ea = __ldg(ae); pl.eb = __ldg(be); }
        }
        ac += plane; ae += plane; bc += plane; be += plane; mc += plane;
        ++pp;
        return pl;
    };

    float pa_prev = -INFINITY, pa_cur = -INFINITY, ca_cur = 0.f;
    float pb_prev = INFINITY, pb_cur = INFINITY, cb_cur = 0.f;
    uint32_t m_cur = 0;
    ExPlane cur;
    if (zc0 > 0) {
        cur = load_next(true);          // plane zc0-1
    } else {
        cur.va = cur.ea = -INFINITY;    // no plane below the cube
        cur.vb = cur.eb = INFINITY;
        cur.m = 0;
    }
    int buf = 0;
#pragma unroll 2
    for (int p = zc0 - 1; p <= zc1; ++p) {
        // software pipeline: the loads of plane p+1 are in flight while plane p is reduced
        const ExPlane nxt = load_next(pp <= zc1 && pp < nz);
        {
            float l = __shfl_up_sync(0xffffffffu, cur.va, 1), r = __shfl_down_sync(0xffffffffu, cur.va, 1);
            l = is_l ? cur.ea : l;
            r = is_r ? cur.ea : r;
            sa[buf][row][lane] = fmaxf(cur.va, fmaxf(l, r));
            l = __shfl_up_sync(0xffffffffu, cur.vb, 1);
            r = __shfl_down_sync(0xffffffffu, cur.vb, 1);
            l = is_l ? cur.eb : l;
            r = is_r ? cur.eb : r;
            sb[buf][row][lane] = fminf(cur.vb, fminf(l, r));
        }
        __syncthreads();
        if (out_row) {  // warp-uniform
            const float m9a = fmaxf(sa[buf][row - 1][lane], fmaxf(sa[buf][row][lane], sa[buf][row + 1][lane]));
            const float m9b = fminf(sb[buf][row - 1][lane], fminf(sb[buf][row][lane], sb[buf][row + 1][lane]));
            if (p > zc0) {  // output plane q = p - 1 in [zc0, zc1)
                const bool keep_a = out_col && !m_cur && ca_cur == fmaxf(pa_prev, fmaxf(pa_cur, m9a));
                const bool keep_b = out_col && !m_cur && cb_cur == fminf(pb_prev, fminf(pb_cur, m9b));
                if (dmax_p) {
                    if (out_col) *dmax_p = keep_a ? ca_cur : 0.f;
                    dmax_p += plane;
                }
                if (dmin_p) {
                    if (out_col) *dmin_p = keep_b ? -cb_cur : 0.f;
                    dmin_p += plane;
                }
                const uint32_t wa = __ballot_sync(0xffffffffu, keep_a);
                const uint32_t wb = __ballot_sync(0xffffffffu, keep_b);
                if (is_l) {
                    flag_max[widx] = wa;
                    flag_min[widx] = wb;
                }
                widx += wstride;
            }
            pa_prev = pa_cur; pa_cur = m9a;
            pb_prev = pb_cur; pb_cur = m9b;
        }
        ca_cur = cur.va;
        cb_cur = cur.vb;
        m_cur = cur.m;
        cur = nxt;
        buf ^= 1;
    }
}

// TMA-staged, vectorised, barrier-free 3 x 3 x 3 pass (the default when nx % 4 == 0).
//  * The raw planes arrive by TMA: a producer warp issues, per plane, two 3-D bulk tensor loads (box
//    {136 x, TY+2 y, 1 z} of each array, starting 4 columns / 1 row outside the tile so the halo comes
//    with it) into a ring of E4_NST shared-memory stages guarded by full / empty mbarriers.  Elements
//    outside the cube are filled with NaN by the TMA unit: fmaxf / fminf ignore NaN operands, which is
//    exactly scipy's 'reflect' maximum filter (out-of-range neighbours never win) — no edge tests.
//  * Every consumer warp owns one image row of the tile and is independent of the others (no block
//    barrier, no exchange buffer): per plane it reads rows y-1, y, y+1 of both arrays (three LDS.128
//    per array: every lane owns FOUR consecutive x, a warp covers 128 voxels), takes the vertical max
//    in registers, the horizontal max with two shuffles (lanes 0 / 31 fetch the halo column), and keeps
//    the 3x3 plane extrema of z-1, z, z+1 in registers.  Warps drift up to E4_NST - 1 planes apart, so
//    the HBM latency of one is hidden by the arithmetic of the others.
// Flag words keep the layout of the scalar kernels (one 32-bit word per 32 consecutive x): the 4-bit
// nibbles of 8 neighbouring lanes are OR-ed with three shuffles.
constexpr int E4_TY = 14;    // output rows per block = consumer warps (+ 1 producer warp)
constexpr int E4_CZ = 64;    // planes per block
constexpr int E4_NST = 4;    // raw-plane stages
constexpr int E4_BX = 136;   // TMA box width: 128 + 4 columns either side (16-byte granularity)
constexpr int E4_ROWS = E4_TY + 2;
constexpr int E4_STAGE_FLOATS = 2 * E4_ROWS * E4_BX;   // a rows, then b rows
constexpr int E4_THREADS = 32 * (E4_TY + 1);
constexpr int E4_SMEM = E4_NST * E4_STAGE_FLOATS * 4 + 2 * E4_NST * 8;

__device__ __forceinline__ float max3(float x, float y, float z) { return fmaxf(fmaxf(x, y), z); }
__device__ __forceinline__ float min3(float x, float y, float z) { return fminf(fminf(x, y), z); }
__device__ __forceinline__ float4 max3v(float4 p, float4 q, float4 r) {
    return make_float4(max3(p.x, q.x, r.x), max3(p.y, q.y, r.y), max3(p.z, q.z, r.z), max3(p.w, q.w, r.w));
}
__device__ __forceinline__ float4 min3v(float4 p, float4 q, float4 r) {
    return make_float4(min3(p.x, q.x, r.x), min3(p.y, q.y, r.y), min3(p.z, q.z, r.z), min3(p.w, q.w, r.w));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma::smem_u32(bar)) : "memory");
}
// bit i of the result: p_i == q_i
__device__ __forceinline__ uint32_t eq4(float4 p, float4 q) {
    uint32_t k = 0;
    if (p.x == q.x) k |= 1u;
    if (p.y == q.y) k |= 2u;
    if (p.z == q.z) k |= 4u;
    if (p.w == q.w) k |= 8u;
    return k;
}

template <bool DENSE>
__global__ void __launch_bounds__(E4_THREADS, 2)
local_extrema3_tma_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                          const uint8_t *__restrict__ mask, int nz, int ny, int nx, int oy0, int oy1, int ox0a,
                          int ox0, int ox1, float *__restrict__ dense_max, float *__restrict__ dense_min,
                          uint32_t *__restrict__ flag_max, uint32_t *__restrict__ flag_min, int nxw) {
    using namespace tma;
    extern __shared__ __align__(128) unsigned char e4_smem[];
    float *raw = reinterpret_cast<float *>(e4_smem);                                      // [NST][2][ROWS][BX]
    uint64_t *full = reinterpret_cast<uint64_t *>(e4_smem + E4_NST * E4_STAGE_FLOATS * 4);  // [NST]
    uint64_t *empty = full + E4_NST;                                                      // [NST]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // warp E4_TY is the producer
    const int x0 = ox0a + blockIdx.x * 128, x = x0 + 4 * lane;
    const int ty0 = oy0 + blockIdx.y * E4_TY;                    // first output row of the block
    const int zc0 = blockIdx.z * E4_CZ, zc1 = min(nz, zc0 + E4_CZ);
    // planes zc0-1 .. zc1 are needed (NaN outside [0, nz)); the loop walks a multiple of E4_NST planes,
    // the extra ones are loaded and reduced but never output
    const int nplanes = (zc1 - zc0 + 2 + E4_NST - 1) / E4_NST * E4_NST;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < E4_NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], E4_TY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == E4_TY) {  // ---- producer ---------------------------------------------------------------
        if (lane == 0) {
            for (int it = 0; it < nplanes; ++it) {
                const int st = it % E4_NST;
                if (it >= E4_NST) mbar_wait(&empty[st], ((it / E4_NST) - 1) & 1);  // every consumer has read the stage
                float *dst = raw + st * E4_STAGE_FLOATS;
                mbar_expect_tx(&full[st], E4_STAGE_FLOATS * 4);
                tma_load_3d(dst, &a_map, &full[st], x0 - 4, ty0 - 1, zc0 - 1 + it);
                tma_load_3d(dst + E4_ROWS * E4_BX, &b_map, &full[st], x0 - 4, ty0 - 1, zc0 - 1 + it);
            }
        }
        return;
    }

    // ---- consumers: warp w owns output row ty0 + w (box row w + 1) ----------------------------------
    const int y = ty0 + warp;
    const bool is_l = lane == 0, is_r = lane == 31;
    const bool out_row = y < oy1;
    const bool m_ok = mask != nullptr && out_row && x < nx;
    uint32_t colbits = 0;  // voxels of this lane that are outputs
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (out_row && x + i >= ox0 && x + i < ox1) colbits |= 1u << i;
    const size_t plane = (size_t)ny * nx;
    const int wcol = blockIdx.x * 4 + (lane >> 3);
    const bool w_ok = out_row && (lane & 7) == 0 && wcol < nxw;
    const uint32_t yx = out_row ? (uint32_t)y * nx + x : 0u;   // offset inside a plane
    const uint32_t wyx = (out_row ? (uint32_t)(y - oy0) : 0u) * nxw + wcol;
    const uint32_t wstride = (uint32_t)(oy1 - oy0) * nxw;
    const int nib = 4 * (lane & 7);

    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const float4 pinf4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    float4 pa_prev = ninf4, pa_cur = ninf4, ca_cur = ninf4;
    float4 pb_prev = pinf4, pb_cur = pinf4, cb_cur = pinf4;
    // this lane's float4 in box row `warp` (the row above the output row) of array a, stage 0; the halo
    // column is read by every lane (column 132 by lane 31, column 3 by the others: two addresses, no
    // conflict) and used by lanes 0 / 31 only
    const float *base = raw + warp * E4_BX + 4 + 4 * lane;
    const float *ebase = raw + warp * E4_BX + (is_r ? 132 : 3);
    // iteration `it` reduces plane zc0-1+it and outputs plane q = zc0-2+it; the mask word of plane q+2 is
    // requested in the same iteration (two planes ahead)
    const uint8_t *mptr = mask + (size_t)zc0 * plane + yx;
    uint32_t m_use = 0, m_a = 0, m_b = 0, phase = 0;

#pragma unroll 1
    for (int it0 = 0; it0 < nplanes; it0 += E4_NST) {
#pragma unroll
        for (int st = 0; st < E4_NST; ++st) {
            const int q = zc0 - 2 + it0 + st;
            m_use = m_a;
            m_a = m_b;
            m_b = 0;
            if (m_ok && q + 2 < zc1) m_b = __ldg(reinterpret_cast<const uint32_t *>(mptr));
            mptr += plane;
            mbar_wait(&full[st], phase);
            const float *ra = base + st * E4_STAGE_FLOATS, *rb = ra + E4_ROWS * E4_BX;
            const float *qa = ebase + st * E4_STAGE_FLOATS, *qb = qa + E4_ROWS * E4_BX;
            // array a, then array b (keeps the live registers low): vertical extremum of the three rows,
            // horizontal extremum by shuffles, then the window extremum over the three planes
            float4 wa, wb, na, nb;  // window extrema, new centre rows
            {
                const float4 r0 = *reinterpret_cast<const float4 *>(ra);
                na = *reinterpret_cast<const float4 *>(ra + E4_BX);
                const float4 r2 = *reinterpret_cast<const float4 *>(ra + 2 * E4_BX);
                const float e = max3(qa[0], qa[E4_BX], qa[2 * E4_BX]);  // halo column
                const float4 v = max3v(r0, na, r2);
                float l = __shfl_up_sync(0xffffffffu, v.w, 1), r = __shfl_down_sync(0xffffffffu, v.x, 1);
                l = is_l ? e : l;
                r = is_r ? e : r;
                const float4 m9 = make_float4(max3(l, v.x, v.y), max3(v.x, v.y, v.z), max3(v.y, v.z, v.w), max3(v.z, v.w, r));
                wa = max3v(pa_prev, pa_cur, m9);
                pa_prev = pa_cur;
                pa_cur = m9;
            }
            {
                const float4 r0 = *reinterpret_cast<const float4 *>(rb);
                nb = *reinterpret_cast<const float4 *>(rb + E4_BX);
                const float4 r2 = *reinterpret_cast<const float4 *>(rb + 2 * E4_BX);
                const float e = min3(qb[0], qb[E4_BX], qb[2 * E4_BX]);
                const float4 v = min3v(r0, nb, r2);
                float l = __shfl_up_sync(0xffffffffu, v.w, 1), r = __shfl_down_sync(0xffffffffu, v.x, 1);
                if (lane == 0) mbar_arrive(&empty[st]);  // the shuffles above made every lane's samples land
                l = is_l ? e : l;
                r = is_r ? e : r;
                const float4 m9 = make_float4(min3(l, v.x, v.y), min3(v.x, v.y, v.z), min3(v.y, v.z, v.w), min3(v.z, v.w, r));
                wb = min3v(pb_prev, pb_cur, m9);
                pb_prev = pb_cur;
                pb_cur = m9;
            }
            uint32_t free4 = colbits;  // bit i set: voxel i is an output and not masked
            if (m_use & 0x000000ffu) free4 &= ~1u;
            if (m_use & 0x0000ff00u) free4 &= ~2u;
            if (m_use & 0x00ff0000u) free4 &= ~4u;
            if (m_use & 0xff000000u) free4 &= ~8u;
            const uint32_t ka = eq4(ca_cur, wa) & free4, kb = eq4(cb_cur, wb) & free4;
            const bool emit = q >= zc0 && q < zc1;  // block-uniform
            if (DENSE && emit && colbits) {
                const size_t oidx = (size_t)q * plane + yx;
                const float4 da = make_float4((ka & 1u) ? ca_cur.x : 0.f, (ka & 2u) ? ca_cur.y : 0.f,
                                              (ka & 4u) ? ca_cur.z : 0.f, (ka & 8u) ? ca_cur.w : 0.f);
                const float4 db = make_float4((kb & 1u) ? -cb_cur.x : 0.f, (kb & 2u) ? -cb_cur.y : 0.f,
                                              (kb & 4u) ? -cb_cur.z : 0.f, (kb & 8u) ? -cb_cur.w : 0.f);
                if (colbits == 0xfu) {
                    if (dense_max) *reinterpret_cast<float4 *>(dense_max + oidx) = da;
                    if (dense_min) *reinterpret_cast<float4 *>(dense_min + oidx) = db;
                } else {
                    const float dav[4] = {da.x, da.y, da.z, da.w}, dbv[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (colbits & (1u << i)) {
                            if (dense_max) dense_max[oidx + i] = dav[i];
                            if (dense_min) dense_min[oidx + i] = dbv[i];
                        }
                }
            }
            // nibbles of 8 neighbouring lanes -> one flag word (bit = x offset inside the word)
            uint32_t fa = ka << nib, fb = kb << nib;
            fa |= __shfl_xor_sync(0xffffffffu, fa, 1); fb |= __shfl_xor_sync(0xffffffffu, fb, 1);
            fa |= __shfl_xor_sync(0xffffffffu, fa, 2); fb |= __shfl_xor_sync(0xffffffffu, fb, 2);
            fa |= __shfl_xor_sync(0xffffffffu, fa, 4); fb |= __shfl_xor_sync(0xffffffffu, fb, 4);
            if (emit && w_ok) {
                const size_t widx = (size_t)q * wstride + wyx;
                flag_max[widx] = fa;
                flag_min[widx] = fb;
            }
            ca_cur = na;
            cb_cur = nb;
        }
        phase ^= 1u;
    }
}

// ---- exclusive scan of popcounts (3 phases, 1024 words per block) ------------------------
constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const int base = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + inc - v;
}

// phase 1: per-block totals of popc(flags) (or of a 0/1 predicate array)
template <bool POPC>
__global__ void scan_block_totals_kernel(const uint32_t *__restrict__ in, size_t n, int64_t *__restrict__ block_tot) {
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int v = i < n ? (POPC ? __popc(in[i]) : (int)in[i]) : 0;
    int total;
    block_exclusive_scan(v, &total);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

// phase 2: one block turns block totals into exclusive offsets; total -> *grand
__global__ void scan_block_offsets_kernel(int64_t *__restrict__ block_tot, int nblocks, int64_t *__restrict__ grand) {
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int v = i < nblocks ? (int)block_tot[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (i < nblocks) block_tot[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry;
}

// geometry of the flag words (phase 3 of the compaction turns set bits into voxel indices)
struct ExMap {
    int ny, nx;            // dims of the arrays the flags were computed on (the sub-cube)
    int oy0, ony, ox0a;    // flag rows cover y in [oy0, oy0+ony), flag word k starts at x = ox0a + 32 k
    int gny, gnx, gy0, gx0;  // placement of the sub-cube in the whole field (indices reported globally)
};

// ---- K4 -----------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// stats[0] = max(list), spaxel_max[spaxel] = max(spaxel_max, value) ; background-only when segmask given
__global__ void list_stats_kernel(const int64_t *__restrict__ index, const float *__restrict__ value, int64_t n,
                                  const uint8_t *__restrict__ segmask, int64_t img, float *__restrict__ stat,
                                  float *__restrict__ spaxel_max) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float m = -INFINITY;
    if (i < n) {
        const float v = value[i];
        const int64_t sp = index[i] % img;
        if (!(segmask && segmask[sp])) m = v;
        if (spaxel_max) atomic_max_f32(spaxel_max + sp, v);
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_f32(stat, m);
}

// counts[t] += #{entries (background only when segmask) with value > thresholds[t]}
__global__ void list_counts_kernel(const int64_t *__restrict__ index, const float *__restrict__ value, int64_t n,
                                   const int64_t *__restrict__ n_dev,   // optional: true length (device), n = bound
                                   const uint8_t *__restrict__ segmask, int64_t img,
                                   const double *__restrict__ thresholds, int nthresh,
                                   unsigned long long *__restrict__ counts) {
    extern __shared__ unsigned int sm_counts[];
    if (n_dev) {
        // the list overflowed its capacity in the asynchronous call that filled it: the counts would be short.
        // Nobody can raise from here, so block 0 poisons every count with -2^56 (it stays negative through an
        // allreduce SUM over any realistic number of ranks); callers test for negative counts.
        if (*n_dev > n && blockIdx.x == 0)
            for (int t = threadIdx.x; t < nthresh; t += blockDim.x) atomicAdd(&counts[t], 0ull - (1ull << 56));
        n = min(n, *n_dev);
    }
    if ((int64_t)blockIdx.x * blockDim.x >= n) return;   // whole block past the end (uniform)
    for (int t = threadIdx.x; t < nthresh; t += blockDim.x) sm_counts[t] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    double v = 0;
    if (live) {
        v = (double)value[i];
        if (segmask && segmask[index[i] % img]) live = false;
    }
    for (int t = 0; t < nthresh; ++t) {
        const bool hit = live && v > thresholds[t];
        const unsigned int b = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sm_counts[t], __popc(b));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nthresh; t += blockDim.x)
        if (sm_counts[t]) atomicAdd(&counts[t], (unsigned long long)sm_counts[t]);
}

// ---- K6 -----------------------------------------------------------------------------------
__global__ void threshold_flag_kernel(const float *__restrict__ value, int64_t n, double thr, uint32_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (double)value[i] > thr ? 1u : 0u;
}

__global__ void threshold_scatter_kernel(const uint32_t *__restrict__ flag, size_t n, const int64_t *__restrict__ block_off,
                                         const int64_t *__restrict__ index, const float *__restrict__ value,
                                         const uint8_t *__restrict__ profile, int64_t *__restrict__ out_index,
                                         float *__restrict__ out_value, uint8_t *__restrict__ out_profile,
                                         int64_t capacity) {
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t f = i < n ? flag[i] : 0u;
    const int ex = block_exclusive_scan((int)f, nullptr);
    if (!f) return;
    const int64_t pos = block_off[blockIdx.x] + ex;
    if (pos >= capacity) return;
    const int64_t idx = index[i];
    if (out_index) out_index[pos] = idx;
    if (out_value) out_value[pos] = value[i];
    if (out_profile) out_profile[pos] = profile ? profile[idx] : 0;
}

// -------------------------------------------------------------------------------------------

// ---- compaction of the two flag arrays (maxima, minima) in three launches ------------------
// A thread owns CF_ITEMS consecutive flag words (two 16-byte loads), a block CF_THREADS * CF_ITEMS;
// blockIdx.y selects the list.  Phase 1: set bits per block; phase 2: one block per list turns the
// block totals into exclusive offsets; phase 3: every thread re-reads its words and writes its
// entries at offset(block) + exclusive prefix inside the block, which keeps C order = np.where order.
constexpr int CF_THREADS = 256;
constexpr int CF_ITEMS = 8;
constexpr int CF_WORDS = CF_THREADS * CF_ITEMS;

__device__ __forceinline__ void cf_load(const uint32_t *__restrict__ flags, size_t nwords, size_t base, uint32_t (&w)[CF_ITEMS]) {
    if (base + CF_ITEMS <= nwords) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(flags + base));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(flags + base) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < CF_ITEMS; ++j) w[j] = base + j < nwords ? flags[base + j] : 0u;
    }
}

__global__ void __launch_bounds__(CF_THREADS)
flag_totals_kernel(const uint32_t *__restrict__ fa, const uint32_t *__restrict__ fb, size_t nwords,
                   int64_t *__restrict__ tot, int nblocks) {
    const uint32_t *flags = blockIdx.y ? fb : fa;
    uint32_t w[CF_ITEMS];
    cf_load(flags, nwords, (size_t)blockIdx.x * CF_WORDS + (size_t)threadIdx.x * CF_ITEMS, w);
    int c = 0;
#pragma unroll
    for (int j = 0; j < CF_ITEMS; ++j) c += __popc(w[j]);
    int total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tot[(size_t)blockIdx.y * nblocks + blockIdx.x] = total;
}

// grand totals go to device memory and (optionally) to mapped pinned host memory / a caller's device array
__global__ void flag_offsets_kernel(int64_t *__restrict__ tot, int nblocks, int64_t *__restrict__ grand,
                                    int64_t *__restrict__ grand_mapped, int64_t *__restrict__ grand_user) {
    int64_t *t = tot + (size_t)blockIdx.x * nblocks;
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? (int)t[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < nblocks) t[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        grand[blockIdx.x] = carry;
        if (grand_mapped) grand_mapped[blockIdx.x] = carry;
        if (grand_user) grand_user[blockIdx.x] = carry;
    }
}

__global__ void __launch_bounds__(CF_THREADS)
flag_scatter_kernel(const uint32_t *__restrict__ fa, const uint32_t *__restrict__ fb, size_t nwords,
                    const int64_t *__restrict__ tot, int nblocks, const float *__restrict__ src_a,
                    const float *__restrict__ src_b, ExMap m, int nxw, int64_t *__restrict__ idx_a,
                    float *__restrict__ val_a, int64_t *__restrict__ idx_b, float *__restrict__ val_b, int64_t capacity) {
    const bool second = blockIdx.y != 0;
    const uint32_t *flags = second ? fb : fa;
    const float *src = second ? src_b : src_a;
    const float sign = second ? -1.f : 1.f;
    int64_t *out_index = second ? idx_b : idx_a;
    float *out_value = second ? val_b : val_a;
    const size_t base = (size_t)blockIdx.x * CF_WORDS + (size_t)threadIdx.x * CF_ITEMS;
    uint32_t w[CF_ITEMS];
    cf_load(flags, nwords, base, w);
    int c = 0;
#pragma unroll
    for (int j = 0; j < CF_ITEMS; ++j) c += __popc(w[j]);
    const int ex = block_exclusive_scan(c, nullptr);
    if (!c) return;
    int64_t pos = tot[(size_t)blockIdx.y * nblocks + blockIdx.x] + ex;
#pragma unroll 1
    for (int j = 0; j < CF_ITEMS; ++j) {
        uint32_t v = w[j];
        if (!v) continue;
        const size_t i = base + j;
        const size_t rowid = i / nxw;  // z*ony + (y - oy0)
        const int xw = (int)(i - rowid * nxw);
        const int z = (int)(rowid / m.ony), y = (int)(rowid - (size_t)z * m.ony) + m.oy0;
        const size_t lbase = ((size_t)z * m.ny + y) * m.nx + m.ox0a + (size_t)xw * 32;
        const int64_t gbase = ((int64_t)z * m.gny + (y + m.gy0)) * m.gnx + m.gx0 + m.ox0a + (int64_t)xw * 32;
        while (v) {
            const int bit = __ffs(v) - 1;
            v &= v - 1;
            if (pos < capacity) {
                out_index[pos] = gbase + bit;
                out_value[pos] = sign * src[lbase + bit];
            }
            ++pos;
        }
    }
}

// counts -> d_counts[0..1]; lists (when given) in C order
static int compact_flags2(ogn_ctx *ctx, const uint32_t *flag_max, const uint32_t *flag_min, size_t nwords,
                          const float *a, const float *b, ExMap emap, int nxw, int64_t *max_index, float *max_value,
                          int64_t *min_index, float *min_value, int64_t capacity, int64_t *d_counts,
                          int64_t *mapped_counts, int64_t *user_counts) {
    const int nblocks = ogn_div_up((int64_t)nwords, CF_WORDS);
    int64_t *tot = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "scan_tot2", (size_t)2 * nblocks + 2, &tot));
    flag_totals_kernel<<<dim3(nblocks, 2), CF_THREADS, 0, ctx->stream>>>(flag_max, flag_min, nwords, tot, nblocks);
    OGN_LAUNCH_CHECK("flag_totals_kernel");
    flag_offsets_kernel<<<2, 1024, 0, ctx->stream>>>(tot, nblocks, d_counts, mapped_counts, user_counts);
    OGN_LAUNCH_CHECK("flag_offsets_kernel");
    if (max_index && max_value && min_index && min_value && capacity > 0) {
        flag_scatter_kernel<<<dim3(nblocks, 2), CF_THREADS, 0, ctx->stream>>>(flag_max, flag_min, nwords, tot, nblocks, a, b,
                                                                             emap, nxw, max_index, max_value, min_index,
                                                                             min_value, capacity);
        OGN_LAUNCH_CHECK("flag_scatter_kernel");
    }
    return OGN_OK;
}

// Local extrema of the voxels of `owned` (window of the [nz][ny][nx] sub-cube placed at `place` in
// the whole field).  Dense products are sub-cube shaped (only the window is written); list
// indices are linear indices of the whole field.
int ogn_extrema_run(ogn_ctx *ctx, const float *a, const float *b, const uint8_t *mask, int nz, int ny, int nx,
                    ogn_window owned, ogn_place place, int sz, int sy, int sx, float *dense_max, float *dense_min,
                    int64_t *max_index, float *max_value, int64_t *min_index, float *min_value, int64_t capacity,
                    int64_t *counts) {
    if (!ctx) return OGN_ERR_ARG;
    if (nz <= 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "cube shape (%d,%d,%d) is empty", nz, ny, nx);
    if (sz < 1 || sy < 1 || sx < 1 || !(sz & 1) || !(sy & 1) || !(sx & 1))
        return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "window (%d,%d,%d): only odd sizes are supported", sz, sy, sx);
    if (!a || !b) return ogn_fail(ctx, OGN_ERR_ARG, "a / b must not be NULL");
    if (nz > 65535) return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "nz = %d exceeds the launch grid", nz);
    if (owned.y0 < 0 || owned.x0 < 0 || owned.y1 > ny || owned.x1 > nx || owned.y0 >= owned.y1 || owned.x0 >= owned.x1)
        return ogn_fail(ctx, OGN_ERR_ARG, "owned window outside the sub-cube");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t vol = (size_t)nz * ny * nx;
    const int ony = owned.y1 - owned.y0;
    const int ox0a = owned.x0 / 32 * 32;
    const int nxw = ogn_div_up(owned.x1 - ox0a, 32);
    const size_t nwords = (size_t)nz * ony * nxw;

    const void *da = nullptr, *db = nullptr, *dm = nullptr;
    OGN_TRY(ogn_input(ctx, "ext_a", a, vol * 4, &da));
    if (b == a) db = da;
    else OGN_TRY(ogn_input(ctx, "ext_b", b, vol * 4, &db));
    if (mask) OGN_TRY(ogn_input(ctx, "ext_mask", mask, vol, &dm));

    void *d_dmax = nullptr, *d_dmin = nullptr;
    if (dense_max) OGN_TRY(ogn_output(ctx, "ext_dense_max", dense_max, vol * 4, &d_dmax));
    if (dense_min) OGN_TRY(ogn_output(ctx, "ext_dense_min", dense_min, vol * 4, &d_dmin));
    uint32_t *flag_max = nullptr, *flag_min = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "ext_flag_max", nwords, &flag_max));
    OGN_TRY(ogn_scratch_t(ctx, "ext_flag_min", nwords, &flag_min));

    ogn_timer *t_k3 = new ogn_timer(ctx, "k3_local_extrema");
    // OGN_K3_SCALAR=1 forces the scalar shared-memory kernel (also used when nx % 4 != 0)
    static const bool no_v4_kernel = getenv("OGN_K3_SCALAR") != nullptr;
    const bool vec_ok = nx % 4 == 0 && ((reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(db)) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(dm) & 3) == 0 &&
                        (!d_dmax || (reinterpret_cast<uintptr_t>(d_dmax) & 15) == 0) &&
                        (!d_dmin || (reinterpret_cast<uintptr_t>(d_dmin) & 15) == 0);
    if (sz == 3 && sy == 3 && sx == 3 && vec_ok && !no_v4_kernel) {
        ctx->variants["k3"] = "tma3x3x3";
        CUtensorMap a_map, b_map;
        OGN_TRY(ogn_make_tile_map(ctx, &a_map, (const float *)da, nz, ny, nx, nx, E4_BX, E4_ROWS, 1, true));
        OGN_TRY(ogn_make_tile_map(ctx, &b_map, (const float *)db, nz, ny, nx, nx, E4_BX, E4_ROWS, 1, true));
        auto kern = (d_dmax || d_dmin) ? local_extrema3_tma_kernel<true> : local_extrema3_tma_kernel<false>;
        OGN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, E4_SMEM));
        dim3 block(E4_THREADS);
        dim3 grid(ogn_div_up(nxw, 4), ogn_div_up(ony, E4_TY), ogn_div_up(nz, E4_CZ));
        kern<<<grid, block, E4_SMEM, ctx->stream>>>(a_map, b_map, (const uint8_t *)dm, nz, ny, nx,
                                                                        owned.y0, owned.y1, ox0a, owned.x0, owned.x1,
                                                                        (float *)d_dmax, (float *)d_dmin, flag_max,
                                                                        flag_min, nxw);
        delete t_k3;
        OGN_LAUNCH_CHECK("local_extrema3_tma_kernel");
    } else if (sz == 3 && sy == 3 && sx == 3) {
        ctx->variants["k3"] = "scalar3x3x3";
        dim3 block(32, EX_TY + 2);
        dim3 grid(nxw, ogn_div_up(ony, EX_TY), ogn_div_up(nz, EX_CZ));
        local_extrema3_kernel<<<grid, block, 0, ctx->stream>>>((const float *)da, (const float *)db,
                                                              (const uint8_t *)dm, nz, ny, nx, owned.y0, owned.y1, ox0a,
                                                              owned.x0, owned.x1, (float *)d_dmax, (float *)d_dmin,
                                                              flag_max, flag_min, nxw);
        delete t_k3;
        OGN_LAUNCH_CHECK("local_extrema3_kernel");
    } else {
        ctx->variants["k3"] = "generic";
        dim3 block(32, 8);
        dim3 grid(nxw, ogn_div_up(ony, 8), nz);
        local_extrema_kernel<<<grid, block, 0, ctx->stream>>>((const float *)da, (const float *)db,
                                                             (const uint8_t *)dm, nz, ny, nx, sz / 2, sy / 2, sx / 2,
                                                             owned.y0, owned.y1, ox0a, owned.x0, owned.x1,
                                                             (float *)d_dmax, (float *)d_dmin, flag_max, flag_min, nxw);
        delete t_k3;
        OGN_LAUNCH_CHECK("local_extrema_kernel");
    }
    ogn_timer t_compact(ctx, "k3_compaction");

    const bool want_lists = max_index && max_value && min_index && min_value && capacity > 0;
    void *d_maxi = nullptr, *d_maxv = nullptr, *d_mini = nullptr, *d_minv = nullptr;
    if (want_lists) {
        OGN_TRY(ogn_output(ctx, "ext_max_index", max_index, (size_t)capacity * 8, &d_maxi));
        OGN_TRY(ogn_output(ctx, "ext_max_value", max_value, (size_t)capacity * 4, &d_maxv));
        OGN_TRY(ogn_output(ctx, "ext_min_index", min_index, (size_t)capacity * 8, &d_mini));
        OGN_TRY(ogn_output(ctx, "ext_min_value", min_value, (size_t)capacity * 4, &d_minv));
    }
    int64_t *d_counts = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "ext_counts", (size_t)2, &d_counts));
    const ExMap emap{ny, nx, owned.y0, ony, ox0a, place.gny, place.gnx, place.gy0, place.gx0};
    // the two counts come back through mapped pinned memory written by the kernel (no copy-engine read-back,
    // which would queue behind a peer gather in flight)
    int64_t *res_d = nullptr, *res_h = nullptr;
    OGN_TRY(ogn_result_slots(ctx, &res_d, &res_h));
    int64_t *user_counts = counts && ogn_is_device_ptr(counts) ? counts : nullptr;
    OGN_TRY(compact_flags2(ctx, flag_max, flag_min, nwords, (const float *)da, (const float *)db, emap, nxw,
                           want_lists ? (int64_t *)d_maxi : nullptr, (float *)d_maxv, (int64_t *)d_mini, (float *)d_minv,
                           want_lists ? capacity : 0, d_counts, res_d, user_counts));
    OGN_HT("extrema enqueued");
    // Asynchronous mode: the caller gave a DEVICE counts array and every list / dense product is a device
    // buffer, so nothing has to come back to the host: no synchronisation, the call returns with the
    // kernels still in flight (lists are truncated at `capacity`; the caller compares counts and capacity
    // when it eventually reads them).
    const bool dev_lists = !want_lists || (ogn_is_device_ptr(max_index) && ogn_is_device_ptr(max_value) &&
                                           ogn_is_device_ptr(min_index) && ogn_is_device_ptr(min_value));
    if (user_counts && dev_lists && (!dense_max || d_dmax == dense_max) && (!dense_min || d_dmin == dense_min)) return OGN_OK;
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    OGN_HT("extrema counts back");
    const int64_t h_counts[2] = {res_h[0], res_h[1]};
    if (counts && !user_counts) { counts[0] = h_counts[0]; counts[1] = h_counts[1]; }
    OGN_TRY(ogn_output_commit(ctx, dense_max, d_dmax, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, dense_min, d_dmin, vol * 4));
    if (want_lists) {
        const size_t nmax = (size_t)std::min<int64_t>(h_counts[0], capacity);
        const size_t nmin = (size_t)std::min<int64_t>(h_counts[1], capacity);
        OGN_TRY(ogn_output_commit(ctx, max_index, d_maxi, nmax * 8));
        OGN_TRY(ogn_output_commit(ctx, max_value, d_maxv, nmax * 4));
        OGN_TRY(ogn_output_commit(ctx, min_index, d_mini, nmin * 8));
        OGN_TRY(ogn_output_commit(ctx, min_value, d_minv, nmin * 4));
    }
    OGN_TRY(ogn_finish_call(ctx));
    if (want_lists && (h_counts[0] > capacity || h_counts[1] > capacity))
        return ogn_fail(ctx, OGN_ERR_OVERFLOW, "extremum lists need %lld / %lld entries, capacity is %lld",
                        (long long)h_counts[0], (long long)h_counts[1], (long long)capacity);
    return OGN_OK;
}

extern "C" int ogn_local_extrema(ogn_ctx *ctx, const float *a, const float *b, const uint8_t *mask, int nz, int ny,
                                 int nx, int sz, int sy, int sx, float *dense_max, float *dense_min,
                                 int64_t *max_index, float *max_value, int64_t *min_index, float *min_value,
                                 int64_t capacity, int64_t *counts) {
    return ogn_extrema_run(ctx, a, b, mask, nz, ny, nx, ogn_window{0, ny, 0, nx}, ogn_place{ny, nx, 0, 0}, sz, sy, sx,
                           dense_max, dense_min, max_index, max_value, min_index, min_value, capacity, counts);
}

__global__ void fill_f32_kernel2(float *p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct ListsDev {
    const int64_t *maxi = nullptr, *mini = nullptr;
    const float *maxv = nullptr, *minv = nullptr;
    const uint8_t *seg = nullptr;
};

static int stage_lists(ogn_ctx *ctx, const int64_t *max_index, const float *max_value, int64_t nmax,
                       const int64_t *min_index, const float *min_value, int64_t nmin, const uint8_t *segmask,
                       size_t img, ListsDev *out) {
    const void *d = nullptr;
    if (nmax > 0) {
        OGN_TRY(ogn_input(ctx, "pur_maxi", max_index, (size_t)nmax * 8, &d)); out->maxi = (const int64_t *)d;
        OGN_TRY(ogn_input(ctx, "pur_maxv", max_value, (size_t)nmax * 4, &d)); out->maxv = (const float *)d;
    }
    if (nmin > 0) {
        OGN_TRY(ogn_input(ctx, "pur_mini", min_index, (size_t)nmin * 8, &d)); out->mini = (const int64_t *)d;
        OGN_TRY(ogn_input(ctx, "pur_minv", min_value, (size_t)nmin * 4, &d)); out->minv = (const float *)d;
    }
    if (segmask) {
        OGN_TRY(ogn_input(ctx, "pur_seg", segmask, img, &d)); out->seg = (const uint8_t *)d;
    }
    return OGN_OK;
}

extern "C" int ogn_purity_stats(ogn_ctx *ctx, const int64_t *max_index, const float *max_value, int64_t nmax,
                                const int64_t *min_index, const float *min_value, int64_t nmin,
                                const uint8_t *segmask, int ny, int nx, double *stats, float *spaxel_max) {
    if (!ctx) return OGN_ERR_ARG;
    if (nmax < 0 || nmin < 0 || ny <= 0 || nx <= 0) return ogn_fail(ctx, OGN_ERR_ARG, "invalid sizes");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t img = (size_t)ny * nx;
    ListsDev L;
    OGN_TRY(stage_lists(ctx, max_index, max_value, nmax, min_index, min_value, nmin, segmask, img, &L));
    float *d_stat = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "pur_stat", (size_t)2, &d_stat));
    fill_f32_kernel2<<<1, 32, 0, ctx->stream>>>(d_stat, 2, -INFINITY);
    OGN_LAUNCH_CHECK("fill_f32_kernel");
    void *d_sp = nullptr;
    if (spaxel_max) {
        OGN_TRY(ogn_output(ctx, "pur_spaxel_max", spaxel_max, img * 4, &d_sp));
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, d_sp, 0u, img * 4));
    }
    if (nmax > 0) {
        list_stats_kernel<<<ogn_div_up(nmax, 256), 256, 0, ctx->stream>>>(L.maxi, L.maxv, nmax, nullptr, (int64_t)img,
                                                                          d_stat, (float *)d_sp);
        OGN_LAUNCH_CHECK("list_stats_kernel");
    }
    if (nmin > 0) {
        list_stats_kernel<<<ogn_div_up(nmin, 256), 256, 0, ctx->stream>>>(L.mini, L.minv, nmin, L.seg, (int64_t)img,
                                                                          d_stat + 1, nullptr);
        OGN_LAUNCH_CHECK("list_stats_kernel");
    }
    float h_stat[2];
    OGN_CUDA(cudaMemcpyAsync(h_stat, d_stat, sizeof(h_stat), cudaMemcpyDeviceToHost, ctx->stream));
    OGN_TRY(ogn_output_commit(ctx, spaxel_max, d_sp, img * 4));
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->host_output_pending = false;
    if (stats) { stats[0] = h_stat[0]; stats[1] = h_stat[1]; }
    return OGN_OK;
}

extern "C" int ogn_purity_counts(ogn_ctx *ctx, const int64_t *max_index, const float *max_value, int64_t nmax,
                                 const int64_t *min_index, const float *min_value, int64_t nmin,
                                 const uint8_t *segmask, int ny, int nx, const double *thresholds, int nthresh,
                                 int64_t *n1, int64_t *n0) {
    if (!ctx) return OGN_ERR_ARG;
    if (nmax < 0 || nmin < 0 || ny <= 0 || nx <= 0 || nthresh < 1 || nthresh > 8192 || !thresholds || !n1 || !n0)
        return ogn_fail(ctx, OGN_ERR_ARG, "invalid arguments (nthresh must be in 1..8192)");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t img = (size_t)ny * nx;
    ListsDev L;
    OGN_TRY(stage_lists(ctx, max_index, max_value, nmax, min_index, min_value, nmin, segmask, img, &L));
    const void *d_thr = nullptr;
    OGN_TRY(ogn_input(ctx, "pur_thr", thresholds, (size_t)nthresh * 8, &d_thr));
    void *d_n1 = nullptr, *d_n0 = nullptr;
    OGN_TRY(ogn_output(ctx, "pur_n1", n1, (size_t)nthresh * 8, &d_n1));
    OGN_TRY(ogn_output(ctx, "pur_n0", n0, (size_t)nthresh * 8, &d_n0));
    OGN_TRY(ogn_fill_words(ctx, ctx->stream, d_n1, 0u, (size_t)nthresh * 8));
    OGN_TRY(ogn_fill_words(ctx, ctx->stream, d_n0, 0u, (size_t)nthresh * 8));
    const size_t sm = (size_t)nthresh * sizeof(unsigned int);
    if (nmax > 0) {
        list_counts_kernel<<<ogn_div_up(nmax, 256), 256, sm, ctx->stream>>>(L.maxi, L.maxv, nmax, nullptr, nullptr, (int64_t)img,
                                                                            (const double *)d_thr, nthresh,
                                                                            (unsigned long long *)d_n1);
        OGN_LAUNCH_CHECK("list_counts_kernel");
    }
    if (nmin > 0) {
        list_counts_kernel<<<ogn_div_up(nmin, 256), 256, sm, ctx->stream>>>(L.mini, L.minv, nmin, nullptr, L.seg, (int64_t)img,
                                                                            (const double *)d_thr, nthresh,
                                                                            (unsigned long long *)d_n0);
        OGN_LAUNCH_CHECK("list_counts_kernel");
    }
    OGN_TRY(ogn_output_commit(ctx, n1, d_n1, (size_t)nthresh * 8));
    OGN_TRY(ogn_output_commit(ctx, n0, d_n0, (size_t)nthresh * 8));
    return ogn_finish_call(ctx);
}

// Device-only, synchronisation-free variant: the lists are the capacity-sized device buffers an
// asynchronous ogn_step05* call filled, their true lengths are read on the device from list_counts
// ({#maxima, #minima}, the `counts` output of that call).
extern "C" int ogn_purity_counts_dev(ogn_ctx *ctx, const int64_t *max_index, const float *max_value,
                                     const int64_t *min_index, const float *min_value, int64_t capacity,
                                     const int64_t *list_counts, const uint8_t *segmask, int ny, int nx,
                                     const double *thresholds, int nthresh, int64_t *n1, int64_t *n0) {
    if (!ctx) return OGN_ERR_ARG;
    if (capacity < 0 || ny <= 0 || nx <= 0 || nthresh < 1 || nthresh > 8192 || !thresholds || !n1 || !n0 || !list_counts)
        return ogn_fail(ctx, OGN_ERR_ARG, "invalid arguments (nthresh must be in 1..8192)");
    const void *ptrs[] = {max_index, max_value, min_index, min_value, list_counts, thresholds, n1, n0};
    for (const void *p : ptrs)
        if (!p || !ogn_is_device_ptr(p)) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_purity_counts_dev takes device pointers only");
    if (segmask && !ogn_is_device_ptr(segmask)) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_purity_counts_dev takes device pointers only");
    OGN_CUDA(cudaSetDevice(ctx->device));
    if (n0 == n1 + nthresh) {   // one contiguous [2][nthresh] array: one fill
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, n1, 0u, (size_t)nthresh * 16));
    } else {
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, n1, 0u, (size_t)nthresh * 8));
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, n0, 0u, (size_t)nthresh * 8));
    }
    if (capacity == 0) return OGN_OK;
    const size_t sm = (size_t)nthresh * sizeof(unsigned int);
    const int64_t img = (int64_t)ny * nx;
    list_counts_kernel<<<ogn_div_up(capacity, 256), 256, sm, ctx->stream>>>(max_index, max_value, capacity, list_counts,
                                                                            nullptr, img, thresholds, nthresh,
                                                                            (unsigned long long *)n1);
    OGN_LAUNCH_CHECK("list_counts_kernel");
    list_counts_kernel<<<ogn_div_up(capacity, 256), 256, sm, ctx->stream>>>(min_index, min_value, capacity, list_counts + 1,
                                                                            segmask, img, thresholds, nthresh,
                                                                            (unsigned long long *)n0);
    OGN_LAUNCH_CHECK("list_counts_kernel");
    return OGN_OK;
}

extern "C" int ogn_threshold_extract(ogn_ctx *ctx, const int64_t *index, const float *value, int64_t n,
                                     double threshold, const uint8_t *profile, int64_t *out_index, float *out_value,
                                     uint8_t *out_profile, int64_t capacity, int64_t *out_count) {
    if (!ctx) return OGN_ERR_ARG;
    if (n < 0 || capacity < 0) return ogn_fail(ctx, OGN_ERR_ARG, "invalid sizes");
    OGN_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) {
        if (out_count) *out_count = 0;
        return OGN_OK;
    }
    if (profile && !ogn_is_device_ptr(profile))
        return ogn_fail(ctx, OGN_ERR_ARG, "profile cube must be a device pointer (gather the rows on the host otherwise)");
    const void *d_idx = nullptr, *d_val = nullptr;
    OGN_TRY(ogn_input(ctx, "thr_index", index, (size_t)n * 8, &d_idx));
    OGN_TRY(ogn_input(ctx, "thr_value", value, (size_t)n * 4, &d_val));
    uint32_t *flag = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "thr_flag", (size_t)n, &flag));
    threshold_flag_kernel<<<ogn_div_up(n, 256), 256, 0, ctx->stream>>>((const float *)d_val, n, threshold, flag);
    OGN_LAUNCH_CHECK("threshold_flag_kernel");
    const int nblocks = ogn_div_up(n, SCAN_BLOCK);
    int64_t *block_tot = nullptr, *d_count = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "thr_scan", (size_t)nblocks + 1, &block_tot));
    OGN_TRY(ogn_scratch_t(ctx, "thr_count", (size_t)1, &d_count));
    scan_block_totals_kernel<false><<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(flag, (size_t)n, block_tot);
    OGN_LAUNCH_CHECK("scan_block_totals_kernel");
    scan_block_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(block_tot, nblocks, d_count);
    OGN_LAUNCH_CHECK("scan_block_offsets_kernel");
    void *d_oi = nullptr, *d_ov = nullptr, *d_op = nullptr;
    if (capacity > 0) {
        if (out_index) OGN_TRY(ogn_output(ctx, "thr_out_index", out_index, (size_t)capacity * 8, &d_oi));
        if (out_value) OGN_TRY(ogn_output(ctx, "thr_out_value", out_value, (size_t)capacity * 4, &d_ov));
        if (out_profile) OGN_TRY(ogn_output(ctx, "thr_out_profile", out_profile, (size_t)capacity, &d_op));
        threshold_scatter_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(flag, (size_t)n, block_tot, (const int64_t *)d_idx,
                                                                          (const float *)d_val, profile, (int64_t *)d_oi,
                                                                          (float *)d_ov, (uint8_t *)d_op, capacity);
        OGN_LAUNCH_CHECK("threshold_scatter_kernel");
    }
    int64_t h_count = 0;
    OGN_CUDA(cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
    OGN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_count) *out_count = h_count;
    const size_t m = (size_t)std::min<int64_t>(h_count, capacity);
    OGN_TRY(ogn_output_commit(ctx, out_index, d_oi, m * 8));
    OGN_TRY(ogn_output_commit(ctx, out_value, d_ov, m * 4));
    OGN_TRY(ogn_output_commit(ctx, out_profile, d_op, m));
    OGN_TRY(ogn_finish_call(ctx));
    if (h_count > capacity && (out_index || out_value || out_profile))
        return ogn_fail(ctx, OGN_ERR_OVERFLOW, "threshold list needs %lld entries, capacity is %lld",
                        (long long)h_count, (long long)capacity);
    return OGN_OK;
}
