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
// distance costs two LDS (prefetched one step ahead), 8 FADDs and 8 FFMAs per profile, whose tap is a
// uniform-register operand read from constant memory (LDCU.64).  Profiles are grouped by width (groups
// of G = 10, or 3 for tiny dictionaries) and every half-length is padded with zero taps to a multiple
// of ZB, so the code of one turn of the rings (8 distances, all slots of the group) is a single
// straight-line block that loops — it stays in the instruction cache, which measured faster than
// skipping the zero taps through per-width code variants (11.0 ms against 13.4 ms, Dico_FWHM_2_12).
// Normalisation (table lookup), max / first-wins argmax / min, mask, maxmap / minmap are fused exactly
// as in K2 (lib_origin.py:1197-1217, steps.py:781-793).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "ogn_common.cuh"
#include "ogn_tma.cuh"
#include "ogn_tglr_dev.cuh"

namespace k2f {
using namespace tma;

using namespace ogn_dev;

// Tap table in constant memory, read as float4 through the uniform datapath (LDCU.64 pairs): the taps
// enter the FFMAs as uniform-register operands and cost no vector registers.
//   c_ftaps[(toff + j * GP) / 4 + q] = taps of slots 4q .. 4q+3 at distance j (0 beyond a profile's h_k)
__constant__ float4 c_ftaps[MAXT / 4];

// Packed FP32 (sm_100a): fma.rn.f32x2 -> SASS FFMA2 Rd.F32x2, Ra.F32x2, URb.F32 (uniform tap broadcast to both
// halves), Rc.F32x2.  Two FMAs of one lane per instruction: measured 72.5 TFLOP/s against 60 for scalar FFMAs
// of the same shape (acc += uniform tap * sample; tools/ffma2_probe.cu), because the register file delivers
// a 64-bit pair per operand read, and half the issue slots.
// accumulators of a group: scalar floats, or pairs of consecutive wavelengths for the packed variant
template <int G, bool PK> struct Acc;
template <int G> struct Acc<G, false> {
    float v[G][ZB];
    __device__ __forceinline__ float get(int g, int i) const { return v[g][i]; }
};
template <int G> struct Acc<G, true> {
    f32x2 v[G][ZB / 2];
    __device__ __forceinline__ float get(int g, int i) const {
        float lo, hi;
        unpack2(v[g][i >> 1], lo, hi);
        return (i & 1) ? hi : lo;
    }
};

// Eight consecutive tap distances (one turn of the register rings) for the NA widest slots of the
// group.  Everything is static: ring indices, the slots touched, the tap offsets relative to t4.
//   wj points at row (jb - 1) of the forward window of this lane, i.e. wj[(ZB + jj) * 32] is c[zb+ZB-1+j]
template <int NA, int G, bool PK>
__device__ __forceinline__ void fold_block(const float *__restrict__ wf, const float *__restrict__ wb,
                                           const float4 *__restrict__ t4, float (&F)[ZB], float (&B)[ZB],
                                           float &fn, float &bn, Acc<G, PK> &acc) {
    constexpr int GP = (G + 3) / 4 * 4;
    constexpr int Q0 = (G - NA) / 4;   // first float4 of the row that holds an active slot
#pragma unroll
    for (int jj = 0; jj < ZB; ++jj) {
        // ring slot of absolute sample m is m mod ZB; the block starts at j = 1 mod ZB: static indices
        F[jj % ZB] = fn;                       // c[zb + ZB - 1 + j]
        B[(ZB - 1 - jj) % ZB] = bn;            // c[zb - j]
        fn = wf[(jj + 1) * 32];                // samples of distance j + 1, requested a step ahead
        bn = wb[-(jj + 1) * 32];
        float t[GP];
#pragma unroll
        for (int q = Q0; q < GP / 4; ++q) {
            const float4 v = t4[jj * (GP / 4) + q];
            t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
        float s[ZB];
#pragma unroll
        for (int i = 0; i < ZB; ++i) s[i] = F[(i + 1 + jj) % ZB] + B[(i + ZB - 1 - jj) % ZB];
        if constexpr (PK) {
            f32x2 s2[ZB / 2];
#pragma unroll
            for (int i = 0; i < ZB / 2; ++i) s2[i] = pack2(s[2 * i], s[2 * i + 1]);
#pragma unroll
            for (int g = G - NA; g < G; ++g) {
                const f32x2 tt = pack2(t[g], t[g]);   // uniform: becomes the UR.F32 broadcast operand
#pragma unroll
                for (int i = 0; i < ZB / 2; ++i) acc.v[g][i] = fma2(tt, s2[i], acc.v[g][i]);
            }
        } else {
#pragma unroll
            for (int g = G - NA; g < G; ++g)
#pragma unroll
                for (int i = 0; i < ZB; ++i) acc.v[g][i] = fmaf(t[g], s[i], acc.v[g][i]);
        }
    }
}

// acc[g][i] = num_{slot g}[zb + i] for the profiles of group `grp`; wrow points at c[zb] of this lane.
// The half-lengths are padded to multiples of ZB with zero taps (host side), so the number of active
// slots is constant over a block of ZB distances: one table lookup and one switch per block.
template <int G, bool PK>
__device__ __forceinline__ void fold_group(const float *__restrict__ wrow, const FoldDict &d, int grp,
                                           Acc<G, PK> &acc) {
    constexpr int GP = (G + 3) / 4 * 4;
    // G = 10: one code path (all slots; taps past a profile's half-length are zero).  Measured with the packed
    // kernel as well: per-NA variants {1,2,4,7,10} execute 14 % fewer FFMA2s and run 2 % SLOWER (9.82 vs 9.63 ms,
    // Dico_FWHM_2_12) - the kernel is bound by latency at 3 warps per scheduler, not by the FMA pipe.
    constexpr bool FULL_ONLY = G > 3;
    const float4 *t4 = c_ftaps + (d.toff[grp] >> 2);
    float F[ZB], B[ZB];
#pragma unroll
    for (int i = 0; i < ZB; ++i) {
        F[i] = wrow[i * 32];
        B[i] = F[i];
    }
    float fn = wrow[ZB * 32], bn = wrow[-32];  // samples of distance 1
    {
        float t[GP];
#pragma unroll
        for (int q = 0; q < GP / 4; ++q) {
            const float4 v = t4[q];
            t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
        if constexpr (PK) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const f32x2 tt = pack2(t[g], t[g]);
#pragma unroll
                for (int i = 0; i < ZB / 2; ++i) acc.v[g][i] = mul2(tt, pack2(F[2 * i], F[2 * i + 1]));
            }
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int i = 0; i < ZB; ++i) acc.v[g][i] = t[g] * F[i];  // centre tap (0 for empty slots)
        }
    }
    t4 += GP / 4;  // distance 1
    const float *wf = wrow + ZB * 32, *wb = wrow - 32;
    const int nblk = d.nblk[grp];
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        if (FULL_ONLY) {
            // one code path (all slots; taps past a profile's half-length are zero): the loop body stays
            // resident in the instruction cache, which matters more than the ~15 % extra FFMAs
            fold_block<G, G, PK>(wf, wb, t4, F, B, fn, bn, acc);
        } else {
            switch (d.nab[grp][blk]) {  // active slots in this block of distances (uniform)
                case 1: fold_block<1, G, PK>(wf, wb, t4, F, B, fn, bn, acc); break;
                case 2: fold_block<(G >= 2 ? 2 : G), G, PK>(wf, wb, t4, F, B, fn, bn, acc); break;
                default: fold_block<G, G, PK>(wf, wb, t4, F, B, fn, bn, acc); break;
            }
        }
        wf += ZB * 32;
        wb -= ZB * 32;
        t4 += ZB * (GP / 4);
    }
}

template <int G, int NW, bool G2, bool PK, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
folded_glr_kernel(const __grid_constant__ CUtensorMap num_map, const __grid_constant__ FoldDict dict,
                  int nz, int wny, int wnx,                    // window (= K1 output) dims
                  int oy_off, int ox_off, int ony, int onx,    // window origin inside the [nz][ony][onx] products
                  int cy_off, int cx_off, int gny, int gnx,    // window origin / size of the global field (edge classes)
                  int np, int box_rows, int nbox,
                  const float *__restrict__ rs, int nzp, int ncls, int ncx, int P, int stage_rs, int stage_mask,
                  const uint8_t *__restrict__ mask,
                  float *__restrict__ correl, float *__restrict__ correl_min, uint8_t *__restrict__ profile,
                  float *__restrict__ maxmap, float *__restrict__ minmap, const ogn_gather2 g2) {
    // shared memory: [2 stages of window rows][mbarriers][2 x NW warps of mask rows][2 x NW warps of rs rows]
    extern __shared__ __align__(128) float smem[];
    const int nprof = dict.nprof, hmax = dict.hmax;   // hmax: padded to a multiple of ZB
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
    // G2: this rank owns the gathered cube and stores the voxels it owns there as well (multi-GPU)
    const bool own2 = G2 && oy >= g2.y0 && oy < g2.y1 && ox >= g2.x0 && ox < g2.x1;
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
            if (!rs_staged) {
                // blocks on the field's left / right edge: every lane has its own edge class, so the 1/sqrt(den) rows
                // cannot be staged as one broadcast row per profile.  Their 32-byte sectors are requested into L1 now,
                // a whole pass of arithmetic before the epilogues read them (no register is held meanwhile).
                for (int k = 0; k < nprof; ++k)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rs_col + (size_t)k * ncls * nzp));
            }

#pragma unroll 1
            for (int grp = 0; grp < dict.ngroups; ++grp) {
                Acc<G, PK> acc;
                fold_group<G, PK>(wrow, dict, grp, acc);
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
                        const float t = acc.get(g, i) * r[i];
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
                        if (G2 && own2) g2.dst[((size_t)z * g2.ny + oy + g2.dy) * g2.nx + ox + g2.dx] = c;
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
bool ogn_k2f_prepare(const double *taps, const int *tap_offsets, int nprof, k2f::FoldDict *d, std::vector<float> *table) {
    using namespace k2f;
    // Measured on B200 at 3681x320x320: 11.0 ms against 13.4 ms for K2 with the 20 profiles of
    // Dico_FWHM_2_12, but 2.8 ms against 2.5 ms with the 3 of Dico_3FWHM (the fold saves 53 + 29 against 103
    // FP32 instructions per voxel there, less than the bookkeeping costs): small dictionaries stay on K2.
    // OGN_K2_NOFOLD=1 forces K2, OGN_K2_FOLD=1 forces K2f.
    static const bool disabled = getenv("OGN_K2_NOFOLD") != nullptr;
    static const bool forced = getenv("OGN_K2_FOLD") != nullptr;
    if (disabled || nprof < 1 || (nprof <= 3 && !forced)) return false;
    // groups of 5 (104 registers, 4 blocks per SM) measured 9.52 ms against 9.61 ms for groups of 10 (146 registers,
    // 3 blocks): occupancy is not what holds the kernel back; 128 registers forced on groups of 10 spill (12.2 ms)
    const int G = nprof <= 3 ? 3 : GMAX, GP = (G + 3) / 4 * 4;
    const int ngroups = (nprof + G - 1) / G;
    if (ngroups > MAXG) return false;
    int prev_h = -1;
    for (int k = 0; k < nprof; ++k) {
        const int L = tap_offsets[k + 1] - tap_offsets[k];
        if (L < 1 || L % 2 == 0) return false;
        const int h = (L - 1) / 2;
        if (h < prev_h || h > 8 * MAXB - 8) return false;
        prev_h = h;
        const double *p = taps + tap_offsets[k];
        double amax = 0;
        for (int j = 0; j < L; ++j) amax = std::max(amax, fabs(p[j]));
        for (int j = 0; j < h; ++j)
            if (fabs(p[j] - p[L - 1 - j]) > 1e-12 * amax) return false;
    }
    memset(d, 0, sizeof(*d));
    table->clear();
    d->nprof = nprof; d->ngroups = ngroups; d->G = G;
    auto hpad_of = [&](int k) {   // half-length padded to a multiple of ZB (zero taps), at least ZB
        const int h = (tap_offsets[k + 1] - tap_offsets[k] - 1) / 2;
        return std::max(ZB, (h + ZB - 1) / ZB * ZB);
    };
    d->hmax = hpad_of(nprof - 1);
    for (int grp = 0; grp < ngroups; ++grp) {
        const int k0 = grp * G, nact = std::min(G, nprof - k0);
        const int H = hpad_of(k0 + nact - 1);   // the widest of the group
        const int off = (int)table->size();
        if (off + (H + 1) * GP > MAXT) return false;
        d->toff[grp] = off;
        d->nblk[grp] = H / ZB;
        table->resize(off + (size_t)(H + 1) * GP, 0.f);
        for (int g = 0; g < G; ++g) d->k[grp][g] = (g >= G - nact) ? k0 + g - (G - nact) : -1;   // widest last
        for (int blk = 0; blk < H / ZB; ++blk) {
            int na = 0;
            for (int g = 0; g < G; ++g)
                if (d->k[grp][g] >= 0 && hpad_of(d->k[grp][g]) > blk * ZB) ++na;
            d->nab[grp][blk] = (unsigned char)na;
        }
        for (int g = 0; g < G; ++g) {
            const int k = d->k[grp][g];
            if (k < 0) continue;
            const int L = tap_offsets[k + 1] - tap_offsets[k], h = (L - 1) / 2;
            const double *p = taps + tap_offsets[k];
            for (int j = 0; j <= h; ++j) (*table)[off + j * GP + g] = (float)(0.5 * (p[h + j] + p[h - j]));
        }
    }
    return true;
}

// host table -> constant memory, through the setup's zero-copy upload kernel
int ogn_k2f_upload(ogn_ctx *ctx, ogn_uploader *up, const std::vector<float> &table) {
    void *sym = nullptr;
    OGN_CUDA(cudaGetSymbolAddress(&sym, k2f::c_ftaps));
    return up->add(sym, table.data(), table.size() * sizeof(float));
}

template <int G, int NW, bool PK, int MINB, bool G2 = false>
static int launch_folded(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, ogn_window w,
                         const float *cube_fsf, int pitch, const uint8_t *mask, float *correl, float *correl_min,
                         uint8_t *profile, float *maxmap, float *minmap) {
    using namespace k2f;
    if (!G2 && st.gather2.dst)
        return launch_folded<G, NW, PK, MINB, true>(ctx, stream, st, w, cube_fsf, pitch, mask, correl, correl_min, profile, maxmap, minmap);
    auto kern = folded_glr_kernel<G, NW, G2, PK, MINB>;
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
    int zsplit = ogn_div_up((int64_t)ctx->sm_count * MINB * 8, cols);
    zsplit = std::max(1, std::min(zsplit, nchunk));
    dim3 grid(pitch / 32, wny, zsplit);
    if (grid.y > 65535) return ogn_fail(ctx, OGN_ERR_UNSUPPORTED, "cube too large for the K2 launch grid");
    kern<<<grid, NW * 32, smem, stream>>>(num_map, d, st.nz, wny, wnx, w.y0, w.x0, st.ny, st.nx, w.y0 + st.place.gy0,
                                          w.x0 + st.place.gx0, st.place.gny, st.place.gnx, np, box_rows, nbox, st.rs,
                                          st.nzp, st.ncy * st.ncx, st.ncx, st.P, stage_rs, stage_mask, mask, correl,
                                          correl_min, profile, maxmap, minmap, st.gather2);
    OGN_LAUNCH_CHECK("folded_glr_kernel");
    return OGN_OK;
}

int ogn_k2f_launch(ogn_ctx *ctx, cudaStream_t stream, const ogn_tglr_setup_t &st, ogn_window w, const float *cube_fsf,
                   int pitch, const uint8_t *mask, float *correl, float *correl_min, uint8_t *profile, float *maxmap,
                   float *minmap) {
    // OGN_K2F_SCALAR=1: scalar FFMAs instead of the packed FFMA2 form (diagnostic; both are parity-tested)
    static const bool scalar = getenv("OGN_K2F_SCALAR") != nullptr;
#define OGN_K2F(G_, PK_, MB_) launch_folded<G_, 4, PK_, MB_>(ctx, stream, st, w, cube_fsf, pitch, mask, correl, correl_min, \
                                                           profile, maxmap, minmap)
    ctx->variants["k2"] = std::string("folded:g") + (st.fold->G <= 3 ? "3" : std::to_string(k2f::GMAX)) + (scalar ? ":ffma" : ":ffma2");
    if (st.fold->G <= 3) return scalar ? OGN_K2F(3, false, 5) : OGN_K2F(3, true, 5);
    if (scalar) return OGN_K2F(k2f::GMAX, false, 3);
    return OGN_K2F(k2f::GMAX, true, 3);
#undef OGN_K2F
}
