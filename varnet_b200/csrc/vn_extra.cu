// vn_extra.cu — measurement helpers exported next to the engine: FP32-FMA peak microbenchmark
// (the roofline denominator SURVEY.md §8(d) asks to measure in the same run as the kernels).
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) vn_ffma_peak_kernel(float* out, int iters, float a, float b) {
    // 16 independent FMA chains per thread: enough ILP to cover the 4-cycle FFMA latency at 8 warps/SM-quadrant
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456f) out[0] = s;      // keep the chains alive without a store in the common case
}

// Returns the measured dense FP32 FMA throughput of `device` in TFLOP/s (2 flop per FMA), best of `reps`.
extern "C" int vn_fp32_peak_tflops(int device, int reps, double* tflops) {
    if (!tflops) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    float* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return -2;
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0);
        vn_ffma_peak_kernel<<<grid, block>>>(d, iters, 0.999f, 1e-4f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -2; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16 * 8 * (double)iters * (double)grid * block;
        if (r > 0 && ms > 0.f) best = flop / (ms * 1e-3) / 1e12 > best ? flop / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return 0;
}
