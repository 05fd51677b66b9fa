// Device helpers shared by the two spectral kernels (K2 in ogn_tglr.cu, K2f in ogn_tglr_fold.cu).
#pragma once

#include "ogn_tma.cuh"

// ---------------------------------------------------------------------------
// Edge classes.  With a single FSF the denominator of the GLR does not depend
// on the data: norm_fsf[z,y,x] = sum of K_z^2 over the part of the P x P
// footprint that falls inside the image (lib_origin.py:1039-1041 with
// weights=None).  Along an axis of length n there are min(n, P) distinct
// clippings ("classes"); class P/2 is the interior.
// ---------------------------------------------------------------------------
__host__ __device__ static inline int cls_of(int y, int n, int P) {
    int half = P / 2;
    if (n < P) return y;
    if (y < half) return y;
    if (y >= n - half) return P - (n - y);
    return half;
}

namespace ogn_dev {
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float *addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    const uint32_t d = tma::smem_u32(dst);
    const int bytes = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Packed FP32 (sm_100a): fma.rn.f32x2 -> SASS FFMA2 Rd.F32x2, Ra.F32x2, URb.F32 (scalar tap broadcast to both
// halves), Rc.F32x2: two FMAs of one lane per instruction and per issue slot (tools/ffma2_probe.cu).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
}  // namespace ogn_dev
