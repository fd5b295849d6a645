// hl_bench.cu -- micro-benchmark for the FP32 roofline denominator.
// MEASURED_PEAKS.json carries HBM GB/s and bf16 TFLOP/s only; the collision kernel is
// bounded by the FP32 FMA pipe, so its peak is measured here: 8 independent FFMA
// chains per thread, 1024 threads/CTA, 2 CTAs/SM, timed with CUDA events.
#include "hl_common.cuh"

__global__ void __launch_bounds__(1024) k_ffma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;
}

extern "C" int hl_measure_fp32_peak(hl_ctx* ctx, double* h_tflops) {
    if (!ctx || !h_tflops) { hl_set_error("hl_measure_fp32_peak: bad arguments"); return 1; }
    if (hl_enter(ctx, nullptr, nullptr, "hl_measure_fp32_peak")) return 1;
    float* d = nullptr;
    HL_CUDA_OK(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    HL_CUDA_OK(cudaEventCreate(&e0));
    HL_CUDA_OK(cudaEventCreate(&e1));
    const int iters = 4096, grid = ctx->sm_count * 2;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        HL_CUDA_OK(cudaEventRecord(e0));
        k_ffma_peak<<<grid, 1024>>>(d, iters, 1.0000001f, 1e-9f);
        HL_CUDA_OK(cudaEventRecord(e1));
        HL_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        HL_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * iters * 1024.0 * grid;
        double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *h_tflops = best;
    return 0;
}
