// FP32 FFMA peak of the device: the denominator of the K1/K2 roofline
// (MEASURED_PEAKS.json carries HBM and bf16-GEMM peaks only).
// Prints one JSON line: {"fp32_tflops": ..., "sm_mhz_est": ..., "sms": ...}
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void __launch_bounds__(256) fma_kernel(float *out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

// register-operand form: acc += x * y with x, y in registers (what K1/K2 issue)
template <int ILP>
__global__ void __launch_bounds__(256) fma_reg_kernel(float *out, int iters, const float *in) {
    float acc[ILP], x[4], y[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = in[threadIdx.x + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = in[threadIdx.x + 7 * i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = fmaf(x[i & 3], y[r], acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    float *out;
    cudaMalloc(&out, 4);
    constexpr int ILP = 16;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * ILP * 8 * (double)iters * threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    // sustained: ~2 s back to back
    cudaEventRecord(e0);
    int n = 0;
    float ms = 0;
    do {
        for (int k = 0; k < 20; ++k) fma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
        n += 20;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 2000.f);
    double sustained = 2.0 * ILP * 8 * (double)iters * threads * blocks * n / (ms * 1e-3) / 1e12;
    float *in;
    cudaMalloc(&in, 4096);
    cudaMemset(in, 0, 4096);
    double best_reg = 0;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0);
        fma_reg_kernel<ILP><<<blocks, threads>>>(out, iters, in);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        double tf = 2.0 * ILP * 8 * (double)iters * threads * blocks / (t * 1e-3) / 1e12;
        if (rep >= 2 && tf > best_reg) best_reg = tf;
    }
    printf("{\"fp32_tflops_reg_operands\": %.2f, ", best_reg);
    printf("\"fp32_tflops\": %.2f, \"fp32_tflops_sustained\": %.2f, \"sms\": %d, \"clock_khz_max\": %d, \"nominal_tflops\": %.2f}\n",
           best, sustained, prop.multiProcessorCount, prop.clockRate,
           prop.multiProcessorCount * 128 * 2 * (prop.clockRate * 1e3) / 1e12);
    return 0;
}
