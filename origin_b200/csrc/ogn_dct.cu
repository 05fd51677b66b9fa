// step01: DCT continuum fit and standardisation
// (dct_residual lib_origin.py:150-240, DCTMAT :127-146, Preprocessing.run steps.py:431-465).
//
// Everything that feeds the subtraction raw - cont is FP64: the continuum can be
// 10^3 times the noise, so an FP32 fit would leave residuals above the 1e-5
// parity bound.
//
// Parallelisation: a thread owns one spaxel (threads of a warp = consecutive
// spaxels: coalesced) and ONE OF DCT_NSEG WAVELENGTH SEGMENTS (blockIdx.y), so a
// 320x320 field gives 1.6 M threads instead of 102 400, and the loads of the
// next wavelengths are requested ahead of the arithmetic.
//
//   K5a dct_accum_kernel  per (spaxel, segment): the sums the weighted normal
//                         equations need.  With D0[z][i] = s_i cos(i theta_z) the Gram
//                         matrix D^T W D is G_ij = s_i s_j (C_|i-j| + C_i+j) / 2 with
//                         C_m = sum_z w_z cos(m theta_z): 2M-1 = 21 sums instead of 66,
//                         plus R_i = sum w v cos(i theta), B_i = sum v cos(i theta) and
//                         the "any masked voxel" flag
//       dct_solve_kernel  per spaxel: adds the segments in fixed order, builds G,
//                         Cholesky solve in registers -> order+1 coefficients
//       dct_fit_generic_kernel  any order (one thread per spaxel, Gram matrix in memory)
//   K5b dct_synth_kernel  cont = D0 coef; optionally data = raw - cont (f32) and
//                         per-wavelength partial sums / counts of unmasked data
//   K5c standardise_kernel (data - mean)/sqrt(var), cont/sqrt(var) and per-segment
//                         partial sums of the four per-spaxel reductions, reduce_maps_kernel
#include <math.h>

#include <vector>

#include "ogn_common.cuh"
#include "ogn_tma.cuh"

template <typename T>
__device__ __forceinline__ double ld_as_f64(const T *p, size_t i) { return (double)p[i]; }

constexpr int DCT_NSEG = 16;   // wavelength segments per spaxel

// part[seg][q][S], q = 0 .. 4M-2: C_0..C_{2M-2}, R_0..R_{M-1}, B_0..B_{M-1}; anym[seg][S]
template <int M, typename T>
__global__ void __launch_bounds__(128)
dct_accum_kernel(const T *__restrict__ raw, const T *__restrict__ var, const uint8_t *__restrict__ mask,
                 const double *__restrict__ ctab,  // [nz][2M-1]: cos(m theta_z)
                 int nz, size_t S, int approx, int zseg, double *__restrict__ part, uint8_t *__restrict__ anym) {
    constexpr int NC = 2 * M - 1, NQ = NC + 2 * M;
    extern __shared__ double tab_sm[];   // this block's wavelength segment of the cosine table (broadcast LDS)
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    for (int i = threadIdx.x; i < (z1 - z0) * NC; i += blockDim.x) tab_sm[i] = ctab[(size_t)z0 * NC + i];
    __syncthreads();
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double C[NC], R[M], B[M];
#pragma unroll
    for (int i = 0; i < NC; ++i) C[i] = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { R[i] = 0.0; B[i] = 0.0; }
    bool any_masked = false;
    // software pipeline: the samples of the next wavelength are in flight during the arithmetic
    T nv = T(0), nw = T(1);
    uint8_t nm = 0;
    if (z0 < z1) {
        const size_t o = (size_t)z0 * S + s;
        nv = raw[o];
        if (!approx) { nw = var[o]; nm = mask[o]; }
    }
    for (int z = z0; z < z1; ++z) {
        const double v = (double)nv, vr = (double)nw;
        any_masked |= nm != 0;
        if (z + 1 < z1) {
            const size_t o = (size_t)(z + 1) * S + s;
            nv = raw[o];
            if (!approx) { nw = var[o]; nm = mask[o]; }
        }
        const double *t = tab_sm + (size_t)(z - z0) * NC;
        if (approx) {
#pragma unroll
            for (int i = 0; i < M; ++i) B[i] = fma(t[i], v, B[i]);
        } else {
            const double w = 1.0 / vr, vw = v * w;
#pragma unroll
            for (int m = 0; m < NC; ++m) {
                const double c = t[m];
                C[m] = fma(c, w, C[m]);
                if (m < M) {
                    R[m] = fma(c, vw, R[m]);
                    B[m] = fma(c, v, B[m]);
                }
            }
        }
    }
    double *dst = part + (size_t)blockIdx.y * NQ * S + s;
#pragma unroll
    for (int m = 0; m < NC; ++m) dst[(size_t)m * S] = C[m];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        dst[(size_t)(NC + i) * S] = R[i];
        dst[(size_t)(NC + M + i) * S] = B[i];
    }
    anym[(size_t)blockIdx.y * S + s] = any_masked ? 1 : 0;
}

// coef layout: [M][S] (spaxel fastest) so that warps read/write it coalesced.
template <int M>
__global__ void __launch_bounds__(128)
dct_solve_kernel(const double *__restrict__ part, const uint8_t *__restrict__ anym, int nseg, size_t S, int nz,
                 int approx, double *__restrict__ coef) {
    constexpr int NC = 2 * M - 1, NQ = NC + 2 * M, NG = M * (M + 1) / 2;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double Q[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) Q[q] = 0.0;
    bool any_masked = false;
    for (int seg = 0; seg < nseg; ++seg) {   // fixed order: deterministic sums
        const double *src = part + (size_t)seg * NQ * S + s;
#pragma unroll
        for (int q = 0; q < NQ; ++q) Q[q] += src[(size_t)q * S];
        any_masked |= anym[(size_t)seg * S + s] != 0;
    }
    // DCTMAT scales (lib_origin.py:143-145): s_0 = sqrt(1/nz), s_i = sqrt(2/nz)
    const double s0 = sqrt(1.0 / nz), s1 = sqrt(2.0 / nz);
    double c[M];
    if (approx || any_masked) {
        // unweighted projection, D0 has orthonormal columns (lib_origin.py:192, :237)
#pragma unroll
        for (int i = 0; i < M; ++i) c[i] = (i ? s1 : s0) * Q[NC + M + i];
    } else {
        double G[NG], bw[M];
        auto at = [](int i, int j) { return i * M - i * (i - 1) / 2 + (j - i); };  // i <= j
#pragma unroll
        for (int i = 0; i < M; ++i) {
            bw[i] = (i ? s1 : s0) * Q[NC + i];
#pragma unroll
            for (int j = i; j < M; ++j)
                G[at(i, j)] = 0.5 * (i ? s1 : s0) * (j ? s1 : s0) * (Q[j - i] + Q[i + j]);
        }
        // Cholesky G = L L^T on the packed upper triangle, then two triangular solves
        // (the reference inverts G, lib_origin.py:233-235; same solution to round-off)
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

// cont[z][s] = sum_i d0[z][i] coef[i][s].  One thread per (spaxel, wavelength segment = blockIdx.y).
//   cont_out (f64 or f32, may be NULL), data_out = raw - cont (f32, NaN->excluded by mask),
//   lambda_sum / lambda_cnt: per-wavelength sums over the unmasked voxels of this launch.
//   MT > 0: compile-time order + 1 (coefficients in registers), MT = 0: any order
template <typename T, typename TO, int MT>
__global__ void __launch_bounds__(128)
dct_synth_kernel(const T *__restrict__ raw, const uint8_t *__restrict__ mask, const double *__restrict__ d0, int Mrt,
                 int nz, size_t S, int zseg, const double *__restrict__ coef, TO *__restrict__ cont_out,
                 double *__restrict__ cont64, float *__restrict__ data_out, double *__restrict__ lambda_sum,
                 double *__restrict__ lambda_cnt, int nx, int wy0, int wy1, int wx0, int wx1) {
    extern __shared__ double seg_acc[];   // [2][zseg] per-wavelength accumulators, then [zseg][M] DCTMAT segment
    double *seg_sum = seg_acc, *seg_cnt = seg_acc + zseg, *d0_sm = seg_acc + 2 * zseg;
    const int M = MT > 0 ? MT : Mrt;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < S;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    for (int i = threadIdx.x; i < 2 * zseg; i += blockDim.x) seg_acc[i] = 0.0;
    for (int i = threadIdx.x; i < (z1 - z0) * M; i += blockDim.x) d0_sm[i] = d0[(size_t)z0 * M + i];
    __syncthreads();
    // only spaxels inside the owned window contribute to the per-wavelength sums (multi-GPU tiles)
    const int sy = (int)(s / nx), sx = (int)(s - (size_t)sy * nx);
    const bool counted = live && sy >= wy0 && sy < wy1 && sx >= wx0 && sx < wx1;
    double c[MT > 0 ? MT : DCT_MAXM];
#pragma unroll
    for (int i = 0; i < (MT > 0 ? MT : DCT_MAXM); ++i) c[i] = (live && i < M) ? coef[(size_t)i * S + s] : 0.0;
    T nr = T(0);
    uint8_t nm = 0;
    const bool need_raw = data_out != nullptr || lambda_sum != nullptr;
    if (live && need_raw && z0 < z1) {
        nr = raw[(size_t)z0 * S + s];
        if (counted) nm = mask[(size_t)z0 * S + s];
    }
    for (int z = z0; z < z1; ++z) {
        const double *d = d0_sm + (size_t)(z - z0) * M;
        const double rv = (double)nr;
        const bool masked = nm != 0;
        if (live && need_raw && z + 1 < z1) {
            nr = raw[(size_t)(z + 1) * S + s];
            if (counted) nm = mask[(size_t)(z + 1) * S + s];
        }
        double cont = 0.0;
        if (MT > 0) {
#pragma unroll
            for (int i = 0; i < MT; ++i) cont = fma(d[i], c[i], cont);
        } else {
            for (int i = 0; i < M; ++i) cont = fma(d[i], c[i], cont);
        }
        double v = 0.0, n = 0.0;
        if (live) {
            const size_t o = (size_t)z * S + s;
            if (cont_out) cont_out[o] = (TO)cont;
            if (cont64) cont64[o] = cont;
            if (need_raw) {
                const double r = rv - cont;
                if (data_out) data_out[o] = (float)r;
                if (counted && !masked) { v = r; n = 1.0; }
            }
        }
        if (lambda_sum) {
            // warp: shuffle sum of the data, ballot count; block: shared-memory accumulators of its
            // wavelength segment; one global atomic per wavelength and block at the end
            const unsigned cnt = __popc(__ballot_sync(0xffffffffu, n > 0.0));
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && cnt) {
                atomicAdd(&seg_sum[z - z0], v);
                atomicAdd(&seg_cnt[z - z0], (double)cnt);
            }
        }
    }
    if (lambda_sum) {
        __syncthreads();
        for (int i = threadIdx.x; i < z1 - z0; i += blockDim.x)
            if (seg_cnt[i] > 0.0) {
                atomicAdd(lambda_sum + z0 + i, seg_sum[i]);
                atomicAdd(lambda_cnt + z0 + i, seg_cnt[i]);
            }
    }
}

// steps.py:439-450, :463-465, :472, :480 for one (spaxel, wavelength segment) per thread.  The continuum is
// re-synthesised from the spaxel's coefficients (order + 1 DFMAs per voxel) instead of being stored as a
// float64 cube between the two phases; the four per-spaxel reductions are written as per-segment
// partial sums part4[seg][4][S].
template <typename T, int MT>
__global__ void __launch_bounds__(128)
standardise_kernel(const T *__restrict__ raw, const T *__restrict__ var, const uint8_t *__restrict__ mask,
                   const double *__restrict__ d0, int Mrt, const double *__restrict__ coef,
                   const double *__restrict__ mean, int nz, size_t S, int zseg,
                   float *__restrict__ cube_std, float *__restrict__ cont_dct, double *__restrict__ part4) {
    extern __shared__ double d0_seg[];   // [zseg][M] DCTMAT segment
    const int M = MT > 0 ? MT : Mrt;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    for (int i = threadIdx.x; i < (z1 - z0) * M; i += blockDim.x) d0_seg[i] = d0[(size_t)z0 * M + i];
    __syncthreads();
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double c[MT > 0 ? MT : DCT_MAXM];
#pragma unroll
    for (int i = 0; i < (MT > 0 ? MT : DCT_MAXM); ++i) c[i] = i < M ? coef[(size_t)i * S + s] : 0.0;
    double a_std = 0, a_dct = 0, a_c2 = 0, a_o2 = 0;
    T nr = T(0), nv = T(1);
    uint8_t nm = 0;
    if (z0 < z1) {
        const size_t o = (size_t)z0 * S + s;
        nr = raw[o]; nv = var[o]; nm = mask[o];
    }
    for (int z = z0; z < z1; ++z) {
        const size_t o = (size_t)z * S + s;
        const double rv = (double)nr, vr = (double)nv;
        const bool masked = nm != 0;
        if (z + 1 < z1) {
            const size_t on = o + S;
            nr = raw[on]; nv = var[on]; nm = mask[on];
        }
        const double *d = d0_seg + (size_t)(z - z0) * M;
        double cont = 0.0;
        if (MT > 0) {
#pragma unroll
            for (int i = 0; i < MT; ++i) cont = fma(d[i], c[i], cont);
        } else {
            for (int i = 0; i < M; ++i) cont = fma(d[i], c[i], cont);
        }
        const double sd = sqrt(vr);
        double v = 0.0;
        if (!masked) v = ((rv - cont) - mean[z]) / sd;
        const float cd = (float)(cont / sd);
        if (cube_std) cube_std[o] = (float)v;
        if (cont_dct) cont_dct[o] = cd;
        a_std += v;
        a_o2 += v * v;
        a_dct += (double)cd;
        a_c2 += (double)cd * (double)cd;
    }
    double *dst = part4 + (size_t)blockIdx.y * 4 * S + s;
    dst[0] = a_std;
    dst[S] = a_o2;
    dst[2 * S] = a_dct;
    dst[3 * S] = a_c2;
}

// ima_std = mean_z cube_std, o2map = mean_z cube_std^2, ima_dct = mean_z cont_dct, cont_sumsq = sum_z cont_dct^2
__global__ void reduce_maps_kernel(const double *__restrict__ part4, int nseg, size_t S, int nz,
                                   double *__restrict__ ima_std, double *__restrict__ ima_dct,
                                   double *__restrict__ cont_sumsq, double *__restrict__ o2map) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double a[4] = {0, 0, 0, 0};
    for (int seg = 0; seg < nseg; ++seg)   // fixed order
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] += part4[((size_t)seg * 4 + q) * S + s];
    if (ima_std) ima_std[s] = a[0] / nz;
    if (o2map) o2map[s] = a[1] / nz;
    if (ima_dct) ima_dct[s] = a[2] / nz;
    if (cont_sumsq) cont_sumsq[s] = a[3];
}

// -------------------------------------------------------------------------------------------
// Streamed variants (float32 cubes, order 10, spaxel count a multiple of 16: every MUSE-shaped cube).
//
// The kernels above request the samples of ONE wavelength ahead of the arithmetic; with ~16 resident warps per SM
// that keeps ~5 KB in flight per SM - 0.9 TB/s by Little's law, which is what ncu measured (dct_accum: DRAM
// 0.87 TB/s, FP64 pipe 26 %, top stall long_scoreboard).  Here a block of 128 threads (= 128 consecutive
// spaxels) walks its wavelength segment in stages of ZT = 8 planes: one elected thread issues 2-D TMA loads of
// the {128 spaxels x 8 planes} boxes of raw / var / mask plus a bulk copy of the 8 table rows (cosines or DCTMAT)
// into a ring of NST = 4 shared-memory stages guarded by mbarriers, so ~28 KB per block are in flight without
// holding a register, and every thread reads its own column of the stage (conflict-free) and the table rows
// as broadcasts.  The arithmetic stays FP64 where the subtraction raw - cont needs it; the weights 1/var, the
// final division by sqrt(var) and the square root itself are FP32 (their inputs are float32 and the outputs
// are stored as float32: 2e-7 relative, against a 1e-5 bound).
// The per-wavelength sums of raw - cont (np.nanmean over all spaxels, steps.py:442) are reduced in a FIXED
// order - shuffle tree in a warp, the four warps of a block in order, the blocks in order by
// lambda_reduce_kernel - so cube_std is bit-reproducible from run to run (the atomics above are not).
// -------------------------------------------------------------------------------------------
namespace k5s {
using namespace tma;
constexpr int SP = 128;       // spaxels per block = threads
constexpr int ZT = 8;         // wavelength planes per stage
constexpr int NST = 4;        // stages in the ring
constexpr int CTP = 22;       // cosine table row: 2M-1 = 21 doubles padded to a 16-byte multiple
constexpr int DTP = 12;       // DCTMAT row: M = 11 doubles + the per-wavelength mean in column 11
constexpr int F32_B = ZT * SP * 4, U8_B = ZT * SP;

template <int NTP, bool VAR>
struct Stage {
    static constexpr int TAB_B = (ZT * NTP * 8 + 127) / 128 * 128;
    static constexpr int RAW = 0, VARO = F32_B, MASK = VAR ? 2 * F32_B : F32_B, TAB = MASK + U8_B;
    static constexpr int BYTES = TAB + TAB_B;
    static constexpr uint32_t TX = (VAR ? 2 : 1) * F32_B + U8_B + ZT * NTP * 8;
};

// one elected thread: all loads of one stage arrive on `bar`
template <int NTP, bool VAR>
__device__ __forceinline__ void issue_stage(unsigned char *st, uint64_t *bar, const CUtensorMap *raw_map,
                                            const CUtensorMap *var_map, const CUtensorMap *mask_map,
                                            const double *tab, int s0, int z) {
    using L = Stage<NTP, VAR>;
    mbar_expect_tx(bar, L::TX);
    tma_load_2d(st + L::RAW, raw_map, bar, s0, z);
    if (VAR) tma_load_2d(st + L::VARO, var_map, bar, s0, z);
    tma_load_2d(st + L::MASK, mask_map, bar, s0, z);
    bulk_load(st + L::TAB, tab + (size_t)z * NTP, ZT * NTP * 8, bar);   // the table is padded to a multiple of ZT rows
}

struct Ring {
    unsigned char *base;
    uint64_t *bars;
};
__device__ __forceinline__ Ring ring_init(unsigned char *smem, int stage_bytes) {
    Ring r{smem, reinterpret_cast<uint64_t *>(smem + NST * stage_bytes)};
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) mbar_init(&r.bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    return r;
}

// K5a: the sums of the weighted normal equations (see dct_accum_kernel), part[seg][q][S], anym[seg][S]
template <int M>
__global__ void __launch_bounds__(SP)
accum_stream_kernel(const __grid_constant__ CUtensorMap raw_map, const __grid_constant__ CUtensorMap var_map,
                    const __grid_constant__ CUtensorMap mask_map, const double *__restrict__ ctab, int nz, size_t S,
                    int approx, int zseg, double *__restrict__ part, uint8_t *__restrict__ anym) {
    constexpr int NC = 2 * M - 1, NQ = NC + 2 * M;
    static_assert(NC <= CTP, "table row too short");
    using L = Stage<CTP, true>;
    extern __shared__ __align__(128) unsigned char smem[];
    Ring ring = ring_init(smem, L::BYTES);
    const int tid = threadIdx.x, s0 = blockIdx.x * SP;
    const size_t s = (size_t)s0 + tid;
    const bool live = s < S;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    const int nit = (z1 - z0 + ZT - 1) / ZT;
    if (tid == 0)
        for (int i = 0; i < min(NST, nit); ++i)
            issue_stage<CTP, true>(ring.base + i * L::BYTES, &ring.bars[i], &raw_map, &var_map, &mask_map, ctab, s0, z0 + i * ZT);
    double C[NC], R[M], B[M];
#pragma unroll
    for (int i = 0; i < NC; ++i) C[i] = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { R[i] = 0.0; B[i] = 0.0; }
    bool any_masked = false;
    for (int it = 0; it < nit; ++it) {
        const int slot = it % NST;
        mbar_wait(&ring.bars[slot], (it / NST) & 1);
        const unsigned char *st = ring.base + slot * L::BYTES;
        const float *rawp = reinterpret_cast<const float *>(st + L::RAW) + tid;
        const float *varp = reinterpret_cast<const float *>(st + L::VARO) + tid;
        const unsigned char *mp = st + L::MASK + tid;
        const double *tabp = reinterpret_cast<const double *>(st + L::TAB);
        const int nrow = min(ZT, z1 - (z0 + it * ZT));
        if (live) {
#pragma unroll 2
            for (int zz = 0; zz < nrow; ++zz) {
                const double v = (double)rawp[zz * SP];
                const double *t = tabp + zz * CTP;
                if (approx) {
#pragma unroll
                    for (int i = 0; i < M; ++i) B[i] = fma(t[i], v, B[i]);
                } else {
                    any_masked |= mp[zz * SP] != 0;
                    const double w = (double)(1.0f / varp[zz * SP]);   // var = +inf under the mask: weight 0
                    const double vw = v * w;
#pragma unroll
                    for (int m = 0; m < NC; ++m) {
                        const double c = t[m];
                        C[m] = fma(c, w, C[m]);
                        if (m < M) {
                            R[m] = fma(c, vw, R[m]);
                            B[m] = fma(c, v, B[m]);
                        }
                    }
                }
            }
        }
        __syncthreads();   // every thread is done with the slot
        if (tid == 0 && it + NST < nit)
            issue_stage<CTP, true>(ring.base + slot * L::BYTES, &ring.bars[slot], &raw_map, &var_map, &mask_map, ctab, s0,
                                   z0 + (it + NST) * ZT);
    }
    if (!live) return;
    double *dst = part + (size_t)blockIdx.y * NQ * S + s;
#pragma unroll
    for (int m = 0; m < NC; ++m) dst[(size_t)m * S] = C[m];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        dst[(size_t)(NC + i) * S] = R[i];
        dst[(size_t)(NC + M + i) * S] = B[i];
    }
    anym[(size_t)blockIdx.y * S + s] = any_masked ? 1 : 0;
}

// K5b: per-wavelength partial sums / counts of the unmasked data = raw - cont over the block's spaxels,
// psum[blockIdx.x][z], pcnt[blockIdx.x][z] (fixed summation order inside the block)
template <int M>
__global__ void __launch_bounds__(SP)
sums_stream_kernel(const __grid_constant__ CUtensorMap raw_map, const __grid_constant__ CUtensorMap mask_map,
                   const double *__restrict__ d0p, int nz, size_t S, int zseg, const double *__restrict__ coef,
                   double *__restrict__ psum, int *__restrict__ pcnt, int nx, int wy0, int wy1, int wx0, int wx1) {
    using L = Stage<DTP, false>;
    extern __shared__ __align__(128) unsigned char smem[];
    Ring ring = ring_init(smem, L::BYTES);
    // the block's per-plane sums: every thread parks its 8 values in shared memory, then 16 threads per plane add
    // them in a fixed order (8 sequential adds each + a 4-level shuffle tree) - 8 shuffles per thread and stage
    // instead of the 80 of a per-plane warp reduction
    double *vbuf = reinterpret_cast<double *>(smem + NST * L::BYTES + 64);   // [2][ZT][SP]
    int *wcnt = reinterpret_cast<int *>(vbuf + 2 * ZT * SP);                 // [2][SP/32][ZT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, s0 = blockIdx.x * SP;
    const size_t s = (size_t)s0 + tid;
    const bool live = s < S;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    const int nit = (z1 - z0 + ZT - 1) / ZT;
    if (tid == 0)
        for (int i = 0; i < min(NST, nit); ++i)
            issue_stage<DTP, false>(ring.base + i * L::BYTES, &ring.bars[i], &raw_map, nullptr, &mask_map, d0p, s0, z0 + i * ZT);
    // only spaxels inside the owned window contribute (multi-GPU tiles: the halo belongs to the neighbours)
    const int sy = (int)(s / nx), sx = (int)(s - (size_t)sy * nx);
    const bool counted = live && sy >= wy0 && sy < wy1 && sx >= wx0 && sx < wx1;
    double c[M];
#pragma unroll
    for (int i = 0; i < M; ++i) c[i] = live ? coef[(size_t)i * S + s] : 0.0;
    const int rz = tid >> 4, rp = tid & 15;   // reduction role: plane rz of the stage, partial rp of 16
    for (int it = 0; it < nit; ++it) {
        const int slot = it % NST;
        mbar_wait(&ring.bars[slot], (it / NST) & 1);
        const unsigned char *st = ring.base + slot * L::BYTES;
        const float *rawp = reinterpret_cast<const float *>(st + L::RAW) + tid;
        const unsigned char *mp = st + L::MASK + tid;
        const double *tabp = reinterpret_cast<const double *>(st + L::TAB);
        const int nrow = min(ZT, z1 - (z0 + it * ZT));
        double *vb = vbuf + (size_t)(it & 1) * ZT * SP;
        int *wc = wcnt + ((it & 1) * (SP / 32) + warp) * ZT;
#pragma unroll
        for (int zz = 0; zz < ZT; ++zz) {
            double v = 0.0;
            bool hit = false;
            if (zz < nrow && counted && mp[zz * SP] == 0) {
                const double *d = tabp + zz * DTP;
                double cont = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) cont = fma(d[i], c[i], cont);
                v = (double)rawp[zz * SP] - cont;
                hit = true;
            }
            vb[zz * SP + tid] = v;
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) wc[zz] = __popc(b);
        }
        __syncthreads();   // the stage's samples are consumed and the partial values parked
        if (tid == 0 && it + NST < nit)
            issue_stage<DTP, false>(ring.base + slot * L::BYTES, &ring.bars[slot], &raw_map, nullptr, &mask_map, d0p, s0,
                                    z0 + (it + NST) * ZT);
        {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < SP / 16; ++j) a += vb[rz * SP + rp + 16 * j];      // fixed order
            for (int o = 8; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);    // fixed tree over the 16 partials
            if (rp == 0 && rz < nrow) {
                int n = 0;
                for (int w = 0; w < SP / 32; ++w) n += wcnt[((it & 1) * (SP / 32) + w) * ZT + rz];
                const size_t o = (size_t)blockIdx.x * nz + z0 + it * ZT + rz;
                psum[o] = a;
                pcnt[o] = n;
            }
        }
    }
}

// lambda_sum[z] = sum over the blocks, in block order; optionally the mean (NaN for a fully masked plane, like
// np.nanmean) into column M of the padded DCTMAT table the standardisation reads
__global__ void lambda_reduce_kernel(const double *__restrict__ psum, const int *__restrict__ pcnt, int nblocks, int nz,
                                     double *__restrict__ lambda_sum, double *__restrict__ lambda_cnt) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nz) return;
    double a = 0.0;
    long long n = 0;
    for (int b = 0; b < nblocks; ++b) {
        a += psum[(size_t)b * nz + z];
        n += pcnt[(size_t)b * nz + z];
    }
    lambda_sum[z] = a;
    lambda_cnt[z] = (double)n;
}
__global__ void mean_column_kernel(const double *__restrict__ lambda_sum, const double *__restrict__ lambda_cnt,
                                   const double *__restrict__ mean_in, int nz, double *__restrict__ d0p,
                                   double *__restrict__ mean_out) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nz) return;
    const double m = mean_in ? mean_in[z] : lambda_sum[z] / lambda_cnt[z];   // 0 / 0 = NaN
    if (d0p) d0p[(size_t)z * DTP + DTP - 1] = m;
    if (mean_out) mean_out[z] = m;
}

// K5c: standardisation + the four per-spaxel reductions (see standardise_kernel); the per-wavelength mean sits in
// column 11 of the table rows
template <int M>
__global__ void __launch_bounds__(SP)
std_stream_kernel(const __grid_constant__ CUtensorMap raw_map, const __grid_constant__ CUtensorMap var_map,
                  const __grid_constant__ CUtensorMap mask_map, const double *__restrict__ d0p, int nz, size_t S, int zseg,
                  const double *__restrict__ coef, float *__restrict__ cube_std, float *__restrict__ cont_dct,
                  double *__restrict__ part4) {
    static_assert(M < DTP, "no room for the mean column");
    using L = Stage<DTP, true>;
    extern __shared__ __align__(128) unsigned char smem[];
    Ring ring = ring_init(smem, L::BYTES);
    const int tid = threadIdx.x, s0 = blockIdx.x * SP;
    const size_t s = (size_t)s0 + tid;
    const bool live = s < S;
    const int z0 = blockIdx.y * zseg, z1 = min(nz, z0 + zseg);
    const int nit = (z1 - z0 + ZT - 1) / ZT;
    if (tid == 0)
        for (int i = 0; i < min(NST, nit); ++i)
            issue_stage<DTP, true>(ring.base + i * L::BYTES, &ring.bars[i], &raw_map, &var_map, &mask_map, d0p, s0, z0 + i * ZT);
    double c[M];
#pragma unroll
    for (int i = 0; i < M; ++i) c[i] = live ? coef[(size_t)i * S + s] : 0.0;
    double a_std = 0, a_dct = 0, a_c2 = 0, a_o2 = 0;
    for (int it = 0; it < nit; ++it) {
        const int slot = it % NST;
        mbar_wait(&ring.bars[slot], (it / NST) & 1);
        const unsigned char *st = ring.base + slot * L::BYTES;
        const float *rawp = reinterpret_cast<const float *>(st + L::RAW) + tid;
        const float *varp = reinterpret_cast<const float *>(st + L::VARO) + tid;
        const unsigned char *mp = st + L::MASK + tid;
        const double *tabp = reinterpret_cast<const double *>(st + L::TAB);
        const int nrow = min(ZT, z1 - (z0 + it * ZT));
        if (live) {
#pragma unroll 4
            for (int zz = 0; zz < nrow; ++zz) {
                const double *d = tabp + zz * DTP;
                double cont = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) cont = fma(d[i], c[i], cont);
                const float sd = sqrtf(varp[zz * SP]);
                // (data - mean) / std in FP64 up to the cancellation, FP32 for the division (steps.py:439-446)
                const float v = mp[zz * SP] ? 0.f : (float)(((double)rawp[zz * SP] - cont) - d[DTP - 1]) / sd;
                const float cd = (float)cont / sd;                  // cont /= std, astype(float32) (steps.py:440, :463)
                const size_t o = (size_t)(z0 + it * ZT + zz) * S + s;
                if (cube_std) cube_std[o] = v;
                if (cont_dct) cont_dct[o] = cd;
                a_std += (double)v;
                a_o2 = fma((double)v, (double)v, a_o2);
                a_dct += (double)cd;
                a_c2 = fma((double)cd, (double)cd, a_c2);
            }
        }
        __syncthreads();
        if (tid == 0 && it + NST < nit)
            issue_stage<DTP, true>(ring.base + slot * L::BYTES, &ring.bars[slot], &raw_map, &var_map, &mask_map, d0p, s0,
                                   z0 + (it + NST) * ZT);
    }
    if (!live) return;
    double *dst = part4 + (size_t)blockIdx.y * 4 * S + s;
    dst[0] = a_std;
    dst[S] = a_o2;
    dst[2 * S] = a_dct;
    dst[3 * S] = a_c2;
}
}  // namespace k5s

// -------------------------------------------------------------------------------------------

// The three tables of the DCT path, built on the host once per (nz, order) and kept in the context:
//   dctmat     [nz][M]            DCTMAT, lib_origin.py:143-145
//   dct_costab [nz][2M-1]         cos(m theta_z) (the products of two DCT atoms are sums of two of these)
//   dct_costab_p [nz + ZT][CTP], dct_d0p [nz + ZT][DTP]: the same rows padded to 16-byte multiples (and ZT
//                extra zero rows) for the bulk copies of the streamed kernels; column DTP-1 of dct_d0p receives
//                the per-wavelength mean of the call
struct DctTables {
    const double *d0 = nullptr, *ctab = nullptr, *ctab_p = nullptr;
    double *d0p = nullptr;
};
static int get_dct_tables(ogn_ctx *ctx, int nz, int M, DctTables *t) {
    const int NC = 2 * M - 1;
    const bool padded = M == 11;
    double *d0 = nullptr, *ct = nullptr, *ctp = nullptr, *d0p = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dctmat", (size_t)nz * M, &d0));
    OGN_TRY(ogn_scratch_t(ctx, "dct_costab", (size_t)nz * NC, &ct));
    if (padded) {
        OGN_TRY(ogn_scratch_t(ctx, "dct_costab_p", (size_t)(nz + k5s::ZT) * k5s::CTP, &ctp));
        OGN_TRY(ogn_scratch_t(ctx, "dct_d0p", (size_t)(nz + k5s::ZT) * k5s::DTP, &d0p));
    }
    if (ctx->dct_tab_nz != nz || ctx->dct_tab_M != M) {
        std::vector<double> h0((size_t)nz * M), hc((size_t)nz * NC);
        const double scale = sqrt(2.0 / nz), step = M_PI / nz;
        for (int z = 0; z < nz; ++z) {
            for (int j = 0; j < M; ++j) {
                double v = scale * cos((z + 0.5) * step * j);
                if (j == 0) v *= 1.0 / sqrt(2.0);
                h0[(size_t)z * M + j] = v;
            }
            for (int m = 0; m < NC; ++m) hc[(size_t)z * NC + m] = cos((z + 0.5) * step * m);
        }
        OGN_CUDA(cudaMemcpyAsync(d0, h0.data(), h0.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        OGN_CUDA(cudaMemcpyAsync(ct, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (padded) {
            std::vector<double> hp((size_t)(nz + k5s::ZT) * k5s::CTP, 0.0), hd((size_t)(nz + k5s::ZT) * k5s::DTP, 0.0);
            for (int z = 0; z < nz; ++z) {
                for (int m = 0; m < NC; ++m) hp[(size_t)z * k5s::CTP + m] = hc[(size_t)z * NC + m];
                for (int j = 0; j < M; ++j) hd[(size_t)z * k5s::DTP + j] = h0[(size_t)z * M + j];
            }
            OGN_CUDA(cudaMemcpyAsync(ctp, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            OGN_CUDA(cudaMemcpyAsync(d0p, hd.data(), hd.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        OGN_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope; first call only
        ctx->dct_tab_nz = nz;
        ctx->dct_tab_M = M;
    }
    t->d0 = d0; t->ctab = ct; t->ctab_p = ctp; t->d0p = d0p;
    return OGN_OK;
}

// streamed kernels: float32 cubes, order 10, 16-byte aligned rows for all three TMA maps
static bool dct_streamable(const void *raw, const void *var, const void *mask, int in_dtype, int M, size_t S) {
    static const bool off = getenv("OGN_DCT_NO_STREAM") != nullptr;
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return !off && in_dtype == OGN_F32 && M == 11 && S % 16 == 0 && raw && var && mask && al(raw) && al(var) && al(mask);
}
struct DctMaps { CUtensorMap raw, var, mask; };
static int make_dct_maps(ogn_ctx *ctx, const void *raw, const void *var, const void *mask, int nz, size_t S, DctMaps *m) {
    OGN_TRY(ogn_make_map_2d(ctx, &m->raw, raw, 4, nz, S, S * 4, k5s::SP, k5s::ZT));
    OGN_TRY(ogn_make_map_2d(ctx, &m->var, var, 4, nz, S, S * 4, k5s::SP, k5s::ZT));
    OGN_TRY(ogn_make_map_2d(ctx, &m->mask, mask, 1, nz, S, S, k5s::SP, k5s::ZT));
    return OGN_OK;
}

static inline int dct_segments(int nz) { return std::max(1, std::min(DCT_NSEG, nz / 32)); }
// segments for the kernels that stage `rows` doubles per wavelength in shared memory (<= 40 KB per block)
static inline int dct_segments_smem(int nz, int rows) {
    int nseg = dct_segments(nz);
    while ((size_t)ogn_div_up(nz, nseg) * rows * sizeof(double) > 40 * 1024) ++nseg;
    return nseg;
}

template <typename T>
static int run_fit(ogn_ctx *ctx, const T *raw, const T *var, const uint8_t *mask, const DctTables &tab, int M, int nz,
                   size_t S, int approx, double *coef, const DctMaps *maps) {
    const double *d0 = tab.d0;
    const int blocks = ogn_div_up((int64_t)S, 128);
    if (M == 11) {
        constexpr int NQ = 4 * 11 - 1;
        const int nseg = dct_segments_smem(nz, 21), zseg = ogn_div_up(nz, nseg);
        const double *ctab = tab.ctab;
        double *part = nullptr;
        uint8_t *anym = nullptr;
        OGN_TRY(ogn_scratch_t(ctx, "dct_part", (size_t)nseg * NQ * S, &part));
        OGN_TRY(ogn_scratch_t(ctx, "dct_anym", (size_t)nseg * S, &anym));
        if (maps) {
            ogn_timer t_(ctx, "k5a_dct_accum");
            using L = k5s::Stage<k5s::CTP, true>;
            const size_t sm = (size_t)k5s::NST * L::BYTES + 64;
            k5s::accum_stream_kernel<11><<<dim3(blocks, nseg), k5s::SP, sm, ctx->stream>>>(maps->raw, maps->var, maps->mask,
                                                                                          tab.ctab_p, nz, S, approx, zseg,
                                                                                          part, anym);
            OGN_LAUNCH_CHECK("accum_stream_kernel");
        } else {
            ogn_timer t_(ctx, "k5a_dct_accum");
            const size_t sm = (size_t)zseg * 21 * sizeof(double);
            auto kern = dct_accum_kernel<11, T>;
            kern<<<dim3(blocks, nseg), 128, sm, ctx->stream>>>(raw, var, mask, ctab, nz, S, approx, zseg, part, anym);
            OGN_LAUNCH_CHECK("dct_accum_kernel");
        }
        ogn_timer t_(ctx, "k5a_dct_solve");
        dct_solve_kernel<11><<<blocks, 128, 0, ctx->stream>>>(part, anym, nseg, S, nz, approx, coef);
        OGN_LAUNCH_CHECK("dct_solve_kernel");
    } else {
        double *ws = nullptr;
        OGN_TRY(ogn_scratch_t(ctx, "dct_gram_ws", S * (size_t)M * M, &ws));
        dct_fit_generic_kernel<T><<<blocks, 128, 0, ctx->stream>>>(raw, var, mask, d0, M, nz, S, approx, coef, ws);
        OGN_LAUNCH_CHECK("dct_fit_generic_kernel");
    }
    return OGN_OK;
}

// cont = D0 coef on (spaxel, wavelength-segment) threads; order 10 (M = 11) keeps the coefficients in registers
template <typename T, typename TO>
static int launch_synth(ogn_ctx *ctx, const T *raw, const uint8_t *mask, const double *d0, int M, int nz, size_t S,
                        const double *coef, TO *cont_out, double *cont64, float *data_out, double *lambda_sum,
                        double *lambda_cnt, int nx, int wy0, int wy1, int wx0, int wx1) {
    ogn_timer t_(ctx, "k5b_dct_synth");
    const int nseg = dct_segments_smem(nz, 2 + M), zseg = ogn_div_up(nz, nseg);
    const dim3 grid(ogn_div_up((int64_t)S, 128), nseg);
    const size_t sm = (size_t)(2 + M) * zseg * sizeof(double);
    if (M == 11)
        dct_synth_kernel<T, TO, 11><<<grid, 128, sm, ctx->stream>>>(raw, mask, d0, M, nz, S, zseg, coef, cont_out, cont64,
                                                                  data_out, lambda_sum, lambda_cnt, nx, wy0, wy1, wx0, wx1);
    else
        dct_synth_kernel<T, TO, 0><<<grid, 128, sm, ctx->stream>>>(raw, mask, d0, M, nz, S, zseg, coef, cont_out, cont64,
                                                                 data_out, lambda_sum, lambda_cnt, nx, wy0, wy1, wx0, wx1);
    OGN_LAUNCH_CHECK("dct_synth_kernel");
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
    DctTables tab;
    OGN_TRY(get_dct_tables(ctx, nz, M, &tab));
    double *coef = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dct_coef", S * M, &coef));
    void *d_cont = nullptr;
    OGN_TRY(ogn_output(ctx, "dct_cont_out", cont, vol * (out_dtype == OGN_F64 ? 8 : 4), &d_cont));
    DctMaps maps;
    const bool stream = dct_streamable(in.raw, in.var, in.mask, in_dtype, M, S);
    ctx->variants["step01"] = stream ? "tma-stream" : "column";
    if (stream) OGN_TRY(make_dct_maps(ctx, in.raw, in.var, in.mask, nz, S, &maps));
    if (in_dtype == OGN_F64) {
        OGN_TRY(run_fit<double>(ctx, (const double *)in.raw, (const double *)in.var, in.mask, tab, M, nz, S, approx, coef, nullptr));
    } else {
        OGN_TRY(run_fit<float>(ctx, (const float *)in.raw, (const float *)in.var, in.mask, tab, M, nz, S, approx, coef,
                               stream ? &maps : nullptr));
    }
    if (out_dtype == OGN_F64)
        OGN_TRY((launch_synth<float, double>(ctx, nullptr, nullptr, tab.d0, M, nz, S, coef, (double *)d_cont, nullptr, nullptr,
                                             nullptr, nullptr, nx, 0, ny, 0, nx)));
    else
        OGN_TRY((launch_synth<float, float>(ctx, nullptr, nullptr, tab.d0, M, nz, S, coef, (float *)d_cont, nullptr, nullptr,
                                            nullptr, nullptr, nx, 0, ny, 0, nx)));
    OGN_TRY(ogn_output_commit(ctx, cont, d_cont, vol * (out_dtype == OGN_F64 ? 8 : 4)));
    return ogn_finish_call(ctx);
}

// Phase 1 on device buffers: continuum fit + per-wavelength sums / counts into d_sum / d_cnt (device, [nz]).
static int preprocess_phase1(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype, const uint8_t *mask, int nz,
                             int ny, int nx, int order, int approx, const int *owned, double *d_sum, double *d_cnt) {
    const int wy0 = owned ? owned[0] : 0, wy1 = owned ? owned[1] : ny;
    const int wx0 = owned ? owned[2] : 0, wx1 = owned ? owned[3] : nx;
    if (wy0 < 0 || wy1 > ny || wx0 < 0 || wx1 > nx || wy0 > wy1 || wx0 > wx1)
        return ogn_fail(ctx, OGN_ERR_ARG, "owned window [%d,%d)x[%d,%d) outside the %dx%d field", wy0, wy1, wx0, wx1, ny, nx);
    if (!var || !mask) return ogn_fail(ctx, OGN_ERR_ARG, "var and mask are required");
    const size_t S = (size_t)ny * nx, vol = S * nz;
    const int M = order + 1;
    DctInputs in;
    OGN_TRY(stage_dct_inputs(ctx, raw, var, in_dtype, mask, vol, approx, &in));
    DctTables tab;
    OGN_TRY(get_dct_tables(ctx, nz, M, &tab));
    double *coef = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "dct_coef", S * M, &coef));
    const bool stream = dct_streamable(in.raw, in.var, in.mask, in_dtype, M, S);
    ctx->variants["step01"] = stream ? "tma-stream" : "column";
    DctMaps maps;
    if (stream) OGN_TRY(make_dct_maps(ctx, in.raw, in.var, in.mask, nz, S, &maps));
    if (in_dtype == OGN_F64)
        OGN_TRY(run_fit<double>(ctx, (const double *)in.raw, (const double *)in.var, in.mask, tab, M, nz, S, approx, coef, nullptr));
    else
        OGN_TRY(run_fit<float>(ctx, (const float *)in.raw, (const float *)in.var, in.mask, tab, M, nz, S, approx, coef,
                               stream ? &maps : nullptr));
    if (stream) {
        ogn_timer t_(ctx, "k5b_dct_sums");
        const int nblk = ogn_div_up((int64_t)S, k5s::SP);
        const int nseg = dct_segments(nz), zseg = ogn_div_up(nz, nseg);
        double *psum = nullptr;
        int *pcnt = nullptr;
        OGN_TRY(ogn_scratch_t(ctx, "prep_psum", (size_t)nblk * nz, &psum));
        OGN_TRY(ogn_scratch_t(ctx, "prep_pcnt", (size_t)nblk * nz, &pcnt));
        using L = k5s::Stage<k5s::DTP, false>;
        const size_t sm = (size_t)k5s::NST * L::BYTES + 64 + 2 * k5s::ZT * k5s::SP * sizeof(double) +
                          2 * (k5s::SP / 32) * k5s::ZT * sizeof(int);
        k5s::sums_stream_kernel<11><<<dim3(nblk, nseg), k5s::SP, sm, ctx->stream>>>(maps.raw, maps.mask, tab.d0p, nz, S, zseg, coef,
                                                                                   psum, pcnt, nx, wy0, wy1, wx0, wx1);
        OGN_LAUNCH_CHECK("sums_stream_kernel");
        k5s::lambda_reduce_kernel<<<ogn_div_up(nz, 128), 128, 0, ctx->stream>>>(psum, pcnt, nblk, nz, d_sum, d_cnt);
        OGN_LAUNCH_CHECK("lambda_reduce_kernel");
    } else {
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, d_sum, 0u, (size_t)nz * 8));
        OGN_TRY(ogn_fill_words(ctx, ctx->stream, d_cnt, 0u, (size_t)nz * 8));
        if (in_dtype == OGN_F64)
            OGN_TRY((launch_synth<double, double>(ctx, (const double *)in.raw, in.mask, tab.d0, M, nz, S, coef, nullptr, nullptr,
                                                  nullptr, d_sum, d_cnt, nx, wy0, wy1, wx0, wx1)));
        else
            OGN_TRY((launch_synth<float, double>(ctx, (const float *)in.raw, in.mask, tab.d0, M, nz, S, coef, nullptr, nullptr,
                                                 nullptr, d_sum, d_cnt, nx, wy0, wy1, wx0, wx1)));
    }
    ctx->prep.active = true;
    ctx->prep.nz = nz; ctx->prep.ny = ny; ctx->prep.nx = nx; ctx->prep.in_dtype = in_dtype;
    ctx->prep.raw = in.raw; ctx->prep.var = in.var; ctx->prep.mask = in.mask;
    ctx->prep.coef = coef; ctx->prep.d0 = tab.d0; ctx->prep.d0p = stream ? tab.d0p : nullptr; ctx->prep.M = M;
    return OGN_OK;
}

// Phase 2 on device buffers.  d_mean: per-wavelength mean [nz] (device), or NULL to take d_sum / d_cnt.
static int preprocess_phase2(ogn_ctx *ctx, const double *d_mean, const double *d_sum, const double *d_cnt,
                             double *d_mean_out, float *d_std, float *d_cd, double *d_is, double *d_id, double *d_c2,
                             double *d_o2) {
    const ogn_prep_state &st = ctx->prep;
    const int nz = st.nz;
    const size_t S = (size_t)st.ny * st.nx;
    const int blocks = ogn_div_up((int64_t)S, 128);
    double *part4 = nullptr;
    ogn_timer t_(ctx, "k5c_standardise");
    if (st.d0p) {
        const int nseg = dct_segments(nz), zseg = ogn_div_up(nz, nseg);
        OGN_TRY(ogn_scratch_t(ctx, "prep_part4", (size_t)nseg * 4 * S, &part4));
        k5s::mean_column_kernel<<<ogn_div_up(nz, 128), 128, 0, ctx->stream>>>(d_sum, d_cnt, d_mean, nz, st.d0p, d_mean_out);
        OGN_LAUNCH_CHECK("mean_column_kernel");
        DctMaps maps;
        OGN_TRY(make_dct_maps(ctx, st.raw, st.var, st.mask, nz, S, &maps));
        using L = k5s::Stage<k5s::DTP, true>;
        const size_t sm = (size_t)k5s::NST * L::BYTES + 64;
        k5s::std_stream_kernel<11><<<dim3(blocks, nseg), k5s::SP, sm, ctx->stream>>>(maps.raw, maps.var, maps.mask, st.d0p, nz, S,
                                                                                    zseg, st.coef, d_std, d_cd, part4);
        OGN_LAUNCH_CHECK("std_stream_kernel");
        reduce_maps_kernel<<<blocks, 128, 0, ctx->stream>>>(part4, nseg, S, nz, d_is, d_id, d_c2, d_o2);
        OGN_LAUNCH_CHECK("reduce_maps_kernel");
        return OGN_OK;
    }
    // general path (float64 cubes, other orders, unaligned rows)
    double *mean_dev = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "prep_mean_dev", (size_t)nz, &mean_dev));
    k5s::mean_column_kernel<<<ogn_div_up(nz, 128), 128, 0, ctx->stream>>>(d_sum, d_cnt, d_mean, nz, nullptr, mean_dev);
    OGN_LAUNCH_CHECK("mean_column_kernel");
    if (d_mean_out) OGN_CUDA(cudaMemcpyAsync(d_mean_out, mean_dev, (size_t)nz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    const int nseg = dct_segments_smem(nz, st.M), zseg = ogn_div_up(nz, nseg);
    OGN_TRY(ogn_scratch_t(ctx, "prep_part4", (size_t)nseg * 4 * S, &part4));
    const dim3 grid(blocks, nseg);
    const size_t sm = (size_t)st.M * zseg * sizeof(double);
#define OGN_STD(T_, MT_)                                                                                             \
    standardise_kernel<T_, MT_><<<grid, 128, sm, ctx->stream>>>((const T_ *)st.raw, (const T_ *)st.var, st.mask, st.d0, st.M, \
                                                               st.coef, mean_dev, nz, S, zseg, d_std, d_cd, part4)
    if (st.in_dtype == OGN_F64) { if (st.M == 11) OGN_STD(double, 11); else OGN_STD(double, 0); }
    else { if (st.M == 11) OGN_STD(float, 11); else OGN_STD(float, 0); }
#undef OGN_STD
    OGN_LAUNCH_CHECK("standardise_kernel");
    reduce_maps_kernel<<<blocks, 128, 0, ctx->stream>>>(part4, nseg, S, nz, d_is, d_id, d_c2, d_o2);
    OGN_LAUNCH_CHECK("reduce_maps_kernel");
    return OGN_OK;
}

extern "C" int ogn_preprocess_begin(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype,
                                    const uint8_t *mask, int nz, int ny, int nx, int order, int approx,
                                    const int *owned, double *lambda_sum, double *lambda_cnt) {
    OGN_TRY(check_dct_args(ctx, nz, ny, nx, order));
    if (!lambda_sum || !lambda_cnt) return ogn_fail(ctx, OGN_ERR_ARG, "lambda_sum / lambda_cnt are NULL");
    OGN_CUDA(cudaSetDevice(ctx->device));
    void *d_sum = nullptr, *d_cnt = nullptr;
    OGN_TRY(ogn_output(ctx, "prep_lsum", lambda_sum, (size_t)nz * 8, &d_sum));
    OGN_TRY(ogn_output(ctx, "prep_lcnt", lambda_cnt, (size_t)nz * 8, &d_cnt));
    OGN_TRY(preprocess_phase1(ctx, raw, var, in_dtype, mask, nz, ny, nx, order, approx, owned, (double *)d_sum, (double *)d_cnt));
    OGN_TRY(ogn_output_commit(ctx, lambda_sum, d_sum, (size_t)nz * 8));
    OGN_TRY(ogn_output_commit(ctx, lambda_cnt, d_cnt, (size_t)nz * 8));
    return ogn_finish_call(ctx);
}

struct PrepOutputs {
    void *std_ = nullptr, *cd = nullptr, *is = nullptr, *id = nullptr, *c2 = nullptr, *o2 = nullptr;
};
static int prep_outputs_begin(ogn_ctx *ctx, size_t vol, size_t S, float *cube_std, float *cont_dct, double *ima_std,
                              double *ima_dct, double *cont_sumsq, double *o2map, PrepOutputs *o) {
    if (cube_std) OGN_TRY(ogn_output(ctx, "prep_cube_std", cube_std, vol * 4, &o->std_));
    if (cont_dct) OGN_TRY(ogn_output(ctx, "prep_cont_dct", cont_dct, vol * 4, &o->cd));
    if (ima_std) OGN_TRY(ogn_output(ctx, "prep_ima_std", ima_std, S * 8, &o->is));
    if (ima_dct) OGN_TRY(ogn_output(ctx, "prep_ima_dct", ima_dct, S * 8, &o->id));
    if (cont_sumsq) OGN_TRY(ogn_output(ctx, "prep_c2", cont_sumsq, S * 8, &o->c2));
    if (o2map) OGN_TRY(ogn_output(ctx, "prep_o2", o2map, S * 8, &o->o2));
    return OGN_OK;
}
static int prep_outputs_commit(ogn_ctx *ctx, size_t vol, size_t S, float *cube_std, float *cont_dct, double *ima_std,
                               double *ima_dct, double *cont_sumsq, double *o2map, const PrepOutputs &o) {
    OGN_TRY(ogn_output_commit(ctx, cube_std, o.std_, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, cont_dct, o.cd, vol * 4));
    OGN_TRY(ogn_output_commit(ctx, ima_std, o.is, S * 8));
    OGN_TRY(ogn_output_commit(ctx, ima_dct, o.id, S * 8));
    OGN_TRY(ogn_output_commit(ctx, cont_sumsq, o.c2, S * 8));
    OGN_TRY(ogn_output_commit(ctx, o2map, o.o2, S * 8));
    return OGN_OK;
}

extern "C" int ogn_preprocess_finish(ogn_ctx *ctx, const double *lambda_mean, float *cube_std, float *cont_dct,
                                     double *ima_std, double *ima_dct, double *cont_sumsq, double *o2map) {
    if (!ctx) return OGN_ERR_ARG;
    if (!ctx->prep.active) return ogn_fail(ctx, OGN_ERR_ARG, "ogn_preprocess_finish without ogn_preprocess_begin");
    if (!lambda_mean) return ogn_fail(ctx, OGN_ERR_ARG, "lambda_mean is NULL");
    OGN_CUDA(cudaSetDevice(ctx->device));
    const int nz = ctx->prep.nz;
    const size_t S = (size_t)ctx->prep.ny * ctx->prep.nx, vol = S * nz;
    const void *d_mean = nullptr;
    OGN_TRY(ogn_input(ctx, "prep_mean", lambda_mean, (size_t)nz * 8, &d_mean));
    PrepOutputs o;
    OGN_TRY(prep_outputs_begin(ctx, vol, S, cube_std, cont_dct, ima_std, ima_dct, cont_sumsq, o2map, &o));
    OGN_TRY(preprocess_phase2(ctx, (const double *)d_mean, nullptr, nullptr, nullptr, (float *)o.std_, (float *)o.cd,
                              (double *)o.is, (double *)o.id, (double *)o.c2, (double *)o.o2));
    OGN_TRY(prep_outputs_commit(ctx, vol, S, cube_std, cont_dct, ima_std, ima_dct, cont_sumsq, o2map, o));
    ctx->prep.active = false;
    return ogn_finish_call(ctx);
}

// Both phases in one call for a whole field on one device: the per-wavelength mean never leaves the device, so
// a call whose cubes are device buffers returns without synchronising.
extern "C" int ogn_preprocess(ogn_ctx *ctx, const void *raw, const void *var, int in_dtype, const uint8_t *mask, int nz,
                              int ny, int nx, int order, int approx, double *lambda_mean, float *cube_std,
                              float *cont_dct, double *ima_std, double *ima_dct, double *cont_sumsq, double *o2map) {
    OGN_TRY(check_dct_args(ctx, nz, ny, nx, order));
    OGN_CUDA(cudaSetDevice(ctx->device));
    const size_t S = (size_t)ny * nx, vol = S * nz;
    double *d_sum = nullptr, *d_cnt = nullptr;
    OGN_TRY(ogn_scratch_t(ctx, "prep_lsum", (size_t)nz, &d_sum));
    OGN_TRY(ogn_scratch_t(ctx, "prep_lcnt", (size_t)nz, &d_cnt));
    void *d_mean_out = nullptr;
    if (lambda_mean) OGN_TRY(ogn_output(ctx, "prep_mean_out", lambda_mean, (size_t)nz * 8, &d_mean_out));
    OGN_TRY(preprocess_phase1(ctx, raw, var, in_dtype, mask, nz, ny, nx, order, approx, nullptr, d_sum, d_cnt));
    PrepOutputs o;
    OGN_TRY(prep_outputs_begin(ctx, vol, S, cube_std, cont_dct, ima_std, ima_dct, cont_sumsq, o2map, &o));
    OGN_TRY(preprocess_phase2(ctx, nullptr, d_sum, d_cnt, (double *)d_mean_out, (float *)o.std_, (float *)o.cd, (double *)o.is,
                              (double *)o.id, (double *)o.c2, (double *)o.o2));
    OGN_TRY(prep_outputs_commit(ctx, vol, S, cube_std, cont_dct, ima_std, ima_dct, cont_sumsq, o2map, o));
    OGN_TRY(ogn_output_commit(ctx, lambda_mean, d_mean_out, (size_t)nz * 8));
    ctx->prep.active = false;
    return ogn_finish_call(ctx);
}
