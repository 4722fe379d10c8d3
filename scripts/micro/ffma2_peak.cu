// Microbenchmark: FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) throughput.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), aa = *reinterpret_cast<const unsigned long long*>(&a), bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 x[16]; float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.999f);
    for (int k = 0; k < 16; ++k) x[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (MODE == 0) { x[k].x = fmaf(av.x, bv.x, x[k].x); x[k].y = fmaf(av.y, bv.y, x[k].y); }
                else ffma2(x[k], av, bv);
            }
    }
    float s = 0; for (int k = 0; k < 16; ++k) s += x[k].x + x[k].y;
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<grid, 256>>>(d, iters, 0.999f, 1e-4f); else k<1><<<grid, 256>>>(d, iters, 0.999f, 1e-4f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
        }
        const double flop = 2.0 * 2 * 16 * 4 * (double)iters * grid * 256;
        printf("%s: %.1f TFLOP/s (%.3f ms) %s\n", mode == 0 ? "FFMA  (scalar)" : "FFMA2 (f32x2) ", flop / (best * 1e-3) / 1e12, best, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
