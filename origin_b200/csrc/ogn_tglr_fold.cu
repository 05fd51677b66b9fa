// K2f: spectral stage of the TGLR for dictionaries of symmetric, width-sorted profiles (both shipped
// dictionaries are: Gaussians of increasing FWHM, lib_origin.py:1155-1165 keeps them symmetric).
//
//   num_k[z] = sum_j d_k[j] c[z + h_k - j]                                   (lib_origin.py:1046-1060)
//            = d_k[h_k] c[z] + sum_{m=1..h_k} d_k[h_k + m] (c[z + m] + c[z - m])       (d_k symmetric)
//
// The folded samples s_m[z] = c[z + m] + c[z - m] do not depend on the profile, so a thread computes
// them once per tap distance m (one FADD per output) and feeds them to the accumulators of up to G
// profiles at once: sum_k (h_k + 1) FFMAs + max_k h_k FADDs per voxel instead of sum_k (2 h_k + 1)
// FFMAs — 362 + 45 against 704 for Dico_FWHM_2_12.
//
// Layout: as K2 — the 32 lanes of a warp are 32 consecutive x of one image row; a block of NW warps
// walks wavelength chunks; the column window of a chunk (+ the longest half-profile either side,
// zero-filled outside [0, nz) by the TMA unit = the zero padding of the reference's linear
// convolution) arrives by TMA into one of two shared-memory stages while the previous chunk is being
// reduced.  A warp owns `np` consecutive passes of ZB = 8 wavelengths per chunk.  Per pass and profile
// group a thread keeps two register rings of 8 window samples (forward c[z+m], backward c[z-m]); a tap
// distance costs two LDS (prefetched one step ahead), 8 FADDs and 8 FFMAs per active profile whose tap
// is a uniform-register operand read from the kernel parameter bank.  Profiles are grouped by width
// (groups of G = 10, or 3 for Dico_3FWHM) and sorted inside the group, so the set of profiles that
// still have a tap at distance m is a suffix of the group: one table lookup and one indexed branch
// per distance.  Normalisation (table lookup), max / first-wins argmax / min, mask, maxmap / minmap
// are fused exactly as in K2 (lib_origin.py:1197-1217, steps.py:781-793).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "ogn_common.cuh"
#include "ogn_tma.cuh"

namespace k2f {
using namespace tma;

__host__ __device__ static inline int cls_of(int y, int n, int P) {
    int half = P / 2;
    if (n < P) return y;
    if (y < half) return y;
    if (y >= n - half) return P - (n - y);
    return half;
}

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float *addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    const uint32_t d = smem_u32(dst);
    const int bytes = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// one tap distance for the top `na` slots of the group (slots are sorted by half-length, slot G-1 is
// the widest profile): a dense switch that falls through from the first active slot to the last,
// i.e. one indexed branch per distance instead of a test per profile
template <int S, int G, int GP>
__device__ __forceinline__ void fold_slot(const float (&t)[GP], const float (&s)[ZB], float (&acc)[G][ZB]) {
    if (S < G) {
#pragma unroll
        for (int i = 0; i < ZB; ++i) acc[S < G ? S : 0][i] = fmaf(t[S < G ? S : 0], s[i], acc[S < G ? S : 0][i]);
    }
}
template <int G, int GP>
__device__ __forceinline__ void fold_fma(int na, const float (&t)[GP], const float (&s)[ZB], float (&acc)[G][ZB]) {
    static_assert(G <= 10, "fold_fma handles up to 10 slots");
    switch (G - na) {  // first active slot
        case 0: fold_slot<0, G, GP>(t, s, acc); [[fallthrough]];
        case 1: fold_slot<1, G, GP>(t, s, acc); [[fallthrough]];
        case 2: fold_slot<2, G, GP>(t, s, acc); [[fallthrough]];
        case 3: fold_slot<3, G, GP>(t, s, acc); [[fallthrough]];
        case 4: fold_slot<4, G, GP>(t, s, acc); [[fallthrough]];
        case 5: fold_slot<5, G, GP>(t, s, acc); [[fallthrough]];
        case 6: fold_slot<6, G, GP>(t, s, acc); [[fallthrough]];
        case 7: fold_slot<7, G, GP>(t, s, acc); [[fallthrough]];
        case 8: fold_slot<8, G, GP>(t, s, acc); [[fallthrough]];
        case 9: fold_slot<9, G, GP>(t, s, acc); [[fallthrough]];
        default: break;
    }
}

// acc[g][i] = num_{slot g}[zb + i] for the profiles of group `grp`; wrow points at c[zb] of this lane
template <int G>
__device__ __forceinline__ void fold_group(const float *__restrict__ wrow, const FoldDict &d, int grp,
                                           float (&acc)[G][ZB]) {
    constexpr int GP = (G + 3) / 4 * 4;
    const int toff = d.toff[grp];
    float F[ZB], B[ZB];
#pragma unroll
    for (int i = 0; i < ZB; ++i) {
        F[i] = wrow[i * 32];
        B[i] = F[i];
    }
    float fn = wrow[ZB * 32], bn = wrow[-32];  // samples of distance 1, requested a step ahead
    {
        float t[GP];
#pragma unroll
        for (int q = 0; q < GP / 4; ++q) {
            const float4 v = d.t4[(toff >> 2) + q];
            t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < ZB; ++i) acc[g][i] = t[g] * F[i];  // centre tap (0 for empty slots)
    }
    const unsigned char *na_tab = d.na[grp];
#pragma unroll 1
    for (int jb = 1;; jb += ZB) {
#pragma unroll
        for (int jj = 0; jj < ZB; ++jj) {
            const int j = jb + jj;
            const int na = na_tab[j];  // profiles of the group that reach distance j (0 past the widest)
            if (na == 0) return;
            // ring slot of absolute sample m is m mod ZB; jb = 1 mod ZB makes every index static
            F[jj % ZB] = fn;                 // c[zb + ZB - 1 + j]
            B[(ZB - 1 - jj) % ZB] = bn;      // c[zb - j]
            fn = wrow[(ZB + j) * 32];        // distance j + 1 (one row of slack past the widest profile)
            bn = wrow[-(j + 1) * 32];
            float t[GP];
#pragma unroll
            for (int q = 0; q < GP / 4; ++q) {
                const float4 v = d.t4[((toff + j * GP) >> 2) + q];
                t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
            }
            float s[ZB];
#pragma unroll
            for (int i = 0; i < ZB; ++i) s[i] = F[(i + 1 + jj) % ZB] + B[(i + ZB - 1 - jj) % ZB];
            fold_fma<G, GP>(na, t, s, acc);
        }
    }
}

template <int G, int NW>
__global__ void __launch_bounds__(NW * 32, (G <= 3 ? 5 : 3))
folded_glr_kernel(const __grid_constant__ CUtensorMap num_map, const __grid_constant__ FoldDict dict,
                  int nz, int wny, int wnx,                    // window (= K1 output) dims
                  int oy_off, int ox_off, int ony, int onx,    // window origin inside the [nz][ony][onx] products
                  int cy_off, int cx_off, int gny, int gnx,    // window origin / size of the global field (edge classes)
                  int np, int box_rows, int nbox,
                  const float *__restrict__ rs, int nzp, int ncls, int ncx, int P, int stage_rs, int stage_mask,
                  const uint8_t *__restrict__ mask,
                  float *__restrict__ correl, float *__restrict__ correl_min, uint8_t *__restrict__ profile,
                  float *__restrict__ maxmap, float *__restrict__ minmap) {
    // shared memory: [2 stages of window rows][mbarriers][2 x NW warps of mask rows][2 x NW warps of rs rows]
    extern __shared__ __align__(128) float smem[];
    const int nprof = dict.nprof, hmax = dict.hmax;
    const int wz = np * ZB;       // planes per warp and chunk
    const int cz = NW * wz;       // planes per chunk
    const int win_rows = box_rows * nbox;
    const int stage_floats = win_rows * 32;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * stage_floats);
    uint8_t *mask_sm = reinterpret_cast<uint8_t *>(bars + 2);                                  // [2][NW][wz][32]
    float *rs_sm = reinterpret_cast<float *>(mask_sm + (stage_mask ? 2 * NW * wz * 32 : 0));   // [2][NW][nprof][wz]

    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int x0 = blockIdx.x * 32, x = x0 + lane, y = blockIdx.y;   // window coordinates
    const int oy = y + oy_off, ox = x + ox_off;                       // coordinates in the product cubes
    const int nchunk = (nz + cz - 1) / cz;
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int cls_base = cls_of(y + cy_off, gny, P) * ncx + cls_of(min(x + cx_off, gnx - 1), gnx, P);
    // the warp shares one denominator row when all its lanes are interior in x
    const bool rs_staged = stage_rs && gnx >= P && x0 + cx_off >= P / 2 && x0 + cx_off + 31 < gnx - P / 2;

    auto issue_window = [&](int chunk, int stage) {  // elected thread
        float *dst = smem + stage * stage_floats;
        mbar_expect_tx(&bars[stage], stage_bytes);
        const int zbase = chunk * cz - hmax - 1;
        for (int b = 0; b < nbox; ++b) tma_load_3d(dst + b * box_rows * 32, &num_map, &bars[stage], x0, y, zbase + b * box_rows);
    };
    // per-warp side inputs of one chunk: its wz mask rows and its nprof denominator rows (cp.async)
    auto issue_side = [&](int chunk, int stage) {
        const int z0 = chunk * cz + warp * wz;
        if (stage_mask && mask) {
            uint8_t *dst = mask_sm + ((stage * NW + warp) * wz) * 32;
            for (int c = lane; c < 2 * wz; c += 32) {
                const int row = c >> 1, half = c & 1;
                const bool ok = z0 + row < nz && x0 + ox_off + 16 * half < onx;
                const uint8_t *src = ok ? mask + ((size_t)(z0 + row) * ony + oy) * onx + x0 + ox_off + 16 * half : mask;
                cp_async16(dst + row * 32 + 16 * half, src, ok);
            }
        }
        if (rs_staged && z0 < nz) {
            float *dst = rs_sm + (size_t)(stage * NW + warp) * nprof * wz;
            const float *src = rs + (size_t)cls_base * nzp + z0;
            const int per = wz / 4;
            for (int c = lane; c < nprof * per; c += 32) {
                const int k = c / per, part = c - k * per;
                cp_async16(dst + k * wz + 4 * part, src + (size_t)k * ncls * nzp + 4 * part, true);
            }
        }
    };

    int chunk = blockIdx.z;
    if (chunk < nchunk) {
        if (threadIdx.x == 0) issue_window(chunk, 0);
        issue_side(chunk, 0);
    }

    uint32_t phase_bits = 0;
    for (int it = 0; chunk < nchunk; chunk += gridDim.z, ++it) {
        const int stage = it & 1;
        const int next = chunk + gridDim.z;
        cp_async_wait_all();  // this chunk's side inputs (issued one iteration ago) have landed
        __syncwarp();
        if (next < nchunk) {
            if (threadIdx.x == 0) issue_window(next, stage ^ 1);
            issue_side(next, stage ^ 1);
        }
        mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
        phase_bits ^= 1u << stage;

        const float *win = smem + stage * stage_floats;
        float cmax = -INFINITY, cmin = INFINITY;
#pragma unroll 1
        for (int pass = 0; pass < np; ++pass) {
            const int zl = warp * wz + pass * ZB;   // first plane of the pass inside the chunk
            const int zb = chunk * cz + zl;
            if (zb >= nz) break;
            const float *wrow = win + (zl + hmax + 1) * 32 + lane;  // c[zb] of this lane's column
            float mx[ZB], mn[ZB];
            int arg[ZB];
#pragma unroll
            for (int i = 0; i < ZB; ++i) { mx[i] = -INFINITY; mn[i] = INFINITY; arg[i] = 0; }
            const float *rs_col = rs + (size_t)cls_base * nzp + zb;
            const float *rs_warp = rs_sm + (size_t)(stage * NW + warp) * nprof * wz + pass * ZB;

#pragma unroll 1
            for (int grp = 0; grp < dict.ngroups; ++grp) {
                float acc[G][ZB];
                fold_group<G>(wrow, dict, grp, acc);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int k = dict.k[grp][g];
                    if (k < 0) continue;  // empty slot (uniform)
                    float r[ZB];
                    if (rs_staged) {
                        const float4 *rp = reinterpret_cast<const float4 *>(rs_warp + k * wz);  // broadcast LDS.128
#pragma unroll
                        for (int q = 0; q < ZB / 4; ++q) {
                            const float4 v = rp[q];
                            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
                        }
                    } else {
                        const float4 *rp = reinterpret_cast<const float4 *>(rs_col + (size_t)k * ncls * nzp);
#pragma unroll
                        for (int q = 0; q < ZB / 4; ++q) {
                            const float4 v = __ldg(rp + q);
                            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < ZB; ++i) {
                        const float t = acc[g][i] * r[i];
                        arg[i] = t > mx[i] ? k : arg[i];   // strict: the lowest k wins ties (lib_origin.py:1210)
                        mx[i] = fmaxf(mx[i], t);
                        mn[i] = fminf(mn[i], t);
                    }
                }
            }

            if (x < wnx) {
                uint32_t mbits = 0;
                if (mask) {
                    if (stage_mask) {
                        const uint8_t *mrow = mask_sm + ((stage * NW + warp) * wz + pass * ZB) * 32 + lane;
#pragma unroll
                        for (int i = 0; i < ZB; ++i) mbits |= (mrow[i * 32] ? 1u : 0u) << i;
                    } else {
                        uint8_t mv[ZB];
#pragma unroll
                        for (int i = 0; i < ZB; ++i)
                            mv[i] = (zb + i < nz) ? mask[((size_t)(zb + i) * ony + oy) * onx + ox] : (uint8_t)0;
#pragma unroll
                        for (int i = 0; i < ZB; ++i) mbits |= (mv[i] ? 1u : 0u) << i;
                    }
                }
#pragma unroll
                for (int i = 0; i < ZB; ++i) {
                    const int z = zb + i;
                    if (z < nz) {
                        const size_t o = ((size_t)z * ony + oy) * onx + ox;
                        const bool masked = (mbits >> i) & 1u;
                        const float c = masked ? 0.f : mx[i];
                        if (correl) correl[o] = c;
                        if (correl_min) correl_min[o] = mn[i];
                        if (profile) profile[o] = masked ? (uint8_t)0 : (uint8_t)arg[i];
                        cmax = fmaxf(cmax, c);
                        cmin = fminf(cmin, mn[i]);
                    }
                }
            }
        }
        if (x < wnx && cmax > -INFINITY) {
            if (maxmap) atomic_max_float(maxmap + (size_t)oy * onx + ox, cmax);
            if (minmap) atomic_min_float(minmap + (size_t)oy * onx + ox, cmin);
        }
        __syncthreads();  // the stage is free again before the next iteration refills it
    }
}

}  // namespace k2f

// ---- host side ---------------------------------------------------------------------------------

// Builds the folded dictionary when every profile has odd length, is symmetric (to 1e-12 of its
// largest tap; the two halves are averaged) and the profiles come in non-decreasing width, so that
// walking the groups in order visits k = 0, 1, ... like the reference's loop (lib_origin.py:1207).
bool ogn_k2f_prepare(const double *taps, const int *tap_offsets, int nprof, k2f::FoldDict *d) {
    using namespace k2f;
    // measured on B200 at 3681x320x320: 12.4 ms against 13.4 ms for K2 with the 20 profiles of
    // Dico_FWHM_2_12, but 5.4 ms against 2.6 ms with the 3 of Dico_3FWHM (too little work per tap
    // distance to pay for the per-distance bookkeeping): small dictionaries stay on K2
    static const bool disabled = getenv("OGN_K2_NOFOLD") != nullptr;
    static const bool forced = getenv("OGN_K2_FOLD") != nullptr;
    if (disabled || nprof < 1 || (nprof <= 3 && !forced)) return false;
    const int G = nprof <= 3 ? 3 : GMAX, GP = (G + 3) / 4 * 4;
    const int ngroups = (nprof + G - 1) / G;
    if (ngroups > MAXG) return false;
    int prev_h = -1;
    for (int k = 0; k < nprof; ++k) {
        const int L = tap_offsets[k + 1] - tap_offsets[k];
        if (L < 1 || L % 2 == 0) return false;
        const int h = (L - 1) / 2;
        if (h < prev_h || h > 254) return false;
        prev_h = h;
        const double *p = taps + tap_offsets[k];
        double amax = 0;
        for (int j = 0; j < L; ++j) amax = std::max(amax, fabs(p[j]));
        for (int j = 0; j < h; ++j)
            if (fabs(p[j] - p[L - 1 - j]) > 1e-12 * amax) return false;
    }
    memset(d, 0, sizeof(*d));
    d->nprof = nprof; d->ngroups = ngroups; d->G = G; d->hmax = prev_h;
    int off = 0;
    for (int grp = 0; grp < ngroups; ++grp) {
        const int k0 = grp * G, nact = std::min(G, nprof - k0);
        const int kl = k0 + nact - 1;
        const int H = (tap_offsets[kl + 1] - tap_offsets[kl] - 1) / 2;   // the widest of the group
        if (off + (H + 1) * GP > MAXT) return false;
        d->toff[grp] = off;
        d->H[grp] = H;
        for (int g = 0; g < G; ++g) {
            const int k = k0 + g - (G - nact);   // active profiles sit in the top slots, widest last
            d->k[grp][g] = (g >= G - nact) ? k : -1;
        }
        for (int j = 0; j <= H + 1 && j < 256; ++j) {
            int na = 0;
            for (int g = 0; g < G; ++g) {
                const int k = d->k[grp][g];
                if (k < 0) continue;
                const int L = tap_offsets[k + 1] - tap_offsets[k], h = (L - 1) / 2;
                if (j <= h) {
                    ++na;
                    const double *p = taps + tap_offsets[k];
                    if (j <= H) reinterpret_cast<float *>(d->t4)[off + j * GP + g] = (float)(0.5 * (p[h + j] + p[h - j]));
                }
            }
            d->na[grp][j] = (unsigned char)na;
        }
        off += (H + 1) * GP;
    }
    return true;
}

template <int G, int NW>
static int launch_folded(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, ogn_window w,
                         const float *cube_fsf, int pitch, const uint8_t *mask, float *correl, float *correl_min,
                         uint8_t *profile, float *maxmap, float *minmap) {
    using namespace k2f;
    auto kern = folded_glr_kernel<G, NW>;
    const FoldDict &d = *st.fold;
    const int wny = w.y1 - w.y0, wnx = w.x1 - w.x0;
    const size_t limit = 100 * 1024;
    // passes per warp and chunk: as many as fit (longer chunks re-read less of the window halo)
    static const int np_env = getenv("OGN_K2F_NP") ? atoi(getenv("OGN_K2F_NP")) : 0;
    int np = np_env > 0 ? np_env : (G <= 3 ? 4 : 2);
    int box_rows = 0, nbox = 0, stage_mask = 0, stage_rs = 0;
    size_t smem = 0;
    for (;; --np) {
        if (np < 1)
            return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "profile dictionary does not fit the shared memory of the folded kernel");
        const int wz = np * ZB, cz = NW * wz;
        const int need_rows = cz + 2 * (d.hmax + 1);
        nbox = ogn_div_up(need_rows, 256);
        box_rows = ogn_div_up(need_rows, nbox);
        smem = (size_t)2 * nbox * box_rows * 32 * sizeof(float) + 16;
        stage_mask = mask && st.nx % 16 == 0 && w.x0 % 16 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0;
        if (stage_mask) smem += (size_t)2 * NW * wz * 32;
        const size_t rs_bytes = (size_t)2 * NW * st.nprof * wz * sizeof(float);
        stage_rs = smem + rs_bytes <= limit;
        if (stage_rs) smem += rs_bytes;
        if (smem <= limit) break;
    }
    OGN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUtensorMap num_map;
    OGN_TRY(ogn_make_tile_map(ctx, &num_map, cube_fsf, st.nz, wny, wnx, pitch, 32, 1, box_rows));
    const int cz = NW * np * ZB;
    const int nchunk = ogn_div_up(st.nz, cz);
    const int cols = (pitch / 32) * wny;
    // enough blocks for ~8 waves of resident blocks, at most one block per chunk
    int zsplit = ogn_div_up((int64_t)ctx->sm_count * (G <= 3 ? 5 : 3) * 8, cols);
    zsplit = std::max(1, std::min(zsplit, nchunk));
    dim3 grid(pitch / 32, wny, zsplit);
    if (grid.y > 65535) return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "cube too large for the K2 launch grid");
    kern<<<grid, NW * 32, smem, stream>>>(num_map, d, st.nz, wny, wnx, w.y0, w.x0, st.ny, st.nx, w.y0 + st.place.gy0,
                                          w.x0 + st.place.gx0, st.place.gny, st.place.gnx, np, box_rows, nbox, st.rs,
                                          st.nzp, st.ncy * st.ncx, st.ncx, st.P, stage_rs, stage_mask, mask, correl,
                                          correl_min, profile, maxmap, minmap);
    OGN_LAUNCH_CHECK("folded_glr_kernel");
    return OGN_OK;
}

int ogn_k2f_launch(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, ogn_window w, const float *cube_fsf,
                   int pitch, const uint8_t *mask, float *correl, float *correl_min, uint8_t *profile, float *maxmap,
                   float *minmap) {
    if (st.fold->G <= 3)
        return launch_folded<3, 4>(ctx, stream, st, w, cube_fsf, pitch, mask, correl, correl_min, profile, maxmap, minmap);
    return launch_folded<k2f::GMAX, 4>(ctx, stream, st, w, cube_fsf, pitch, mask, correl, correl_min, profile, maxmap,
                                       minmap);
}
