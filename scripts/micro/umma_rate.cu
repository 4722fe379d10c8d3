// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8) on one SM for the accumulator patterns of vn_tc64.cu:
// does a chain of MMAs into ONE accumulator run slower than the same MMAs spread over independent accumulators?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr uint32_t SBO = 128, LBO = 2048;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// MN-major operand, LayoutType::SWIZZLE_128B_BASE32B: rows of 128 B (32 elements of M/N), k atoms of 4 rows, MN groups 16 KB apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((16384u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t aT, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(aT), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int PAT> __device__ __forceinline__ void issue_one(int i, uint32_t tmem, const uint64_t (&da)[8], const uint64_t (&db)[8]) {
    const int kb = i & 7;
    if (PAT == 0) mma_ss(tmem + 384, da[kb], db[kb], idesc(128), 1u);                                     // SS N128, one accumulator
    if (PAT == 1) mma_ss(tmem + 256 + 128 * (i & 1), da[kb], db[kb], idesc(128), 1u);                     // SS N128, two accumulators
    if (PAT == 2) mma_ts(tmem + 384, tmem + kb * 8, db[kb], idesc(128), 1u);                              // TS N128, one accumulator
    if (PAT == 3) mma_ts(tmem + 256 + 128 * (i & 1), tmem + kb * 8, db[kb], idesc(128), 1u);              // TS N128, two accumulators
    if (PAT == 4) mma_ts(tmem + 384, tmem + kb * 8, db[kb], idesc(64), 1u);                               // TS N64, one accumulator
    if (PAT == 5) mma_ts(tmem + 384 + 64 * (i & 1), tmem + kb * 8, db[kb], idesc(64), 1u);                // TS N64, two accumulators
    if (PAT == 6) mma_ts(tmem + 256 + 64 * (i & 3), tmem + kb * 8, db[kb], idesc(64), 1u);                // TS N64, four accumulators
    if (PAT == 7) { if (i & 1) mma_ts(tmem + 448, tmem + 64 + kb * 8, db[kb], idesc(64), 1u);             // forward pattern of vn_tc64
                    else mma_ts(tmem + 384, tmem + kb * 8, db[kb], idesc(128), 1u); }
    if (PAT == 8) mma_ts(tmem + ((i % 3) == 0 ? 256 : 320), tmem + kb * 8, db[kb], idesc(64), 1u);        // adjoint layer pattern (main, small, small)
    if (PAT == 9) mma_ss(tmem + 384, da[kb], db[kb], idesc(64), 1u);                                      // SS N64, one accumulator
    if (PAT == 10) mma_ts(tmem + 256, tmem + kb * 8, db[kb], idesc(256), 1u);                             // TS N256
    if (PAT == 11) mma_ts(tmem + 256, tmem + kb * 8, db[kb], idesc(192), 1u);                             // TS N192
    if (PAT == 12) mma_ts(tmem + 384, tmem + kb * 8, db[kb], idesc(32), 1u);                              // TS N32
    if (PAT == 13) mma_ss(tmem + 384, make_desc_mn((uint32_t)(da[0] & 0x3FFF) * 16 + kb * 1024), make_desc_mn((uint32_t)(db[0] & 0x3FFF) * 16 + kb * 1024), idesc(128) | (3u << 15), 1u);   // SS N128, A and B MN-major
    if (PAT == 14) mma_ss(tmem + 384, make_desc_mn((uint32_t)(da[0] & 0x3FFF) * 16 + kb * 1024), db[kb], idesc(128) | (1u << 15), 1u);   // SS N128, A MN-major
    if (PAT == 15) mma_ss(tmem + 384, da[kb], make_desc_mn((uint32_t)(db[0] & 0x3FFF) * 16 + kb * 1024), idesc(128) | (1u << 16), 1u);   // SS N128, B MN-major
    if (PAT == 16) mma_ts(tmem + 384, tmem + kb * 8, make_desc_mn((uint32_t)(db[0] & 0x3FFF) * 16 + kb * 1024), idesc(128) | (1u << 16), 1u);   // TS N128, B MN-major
    if (PAT == 17) mma_ts(tmem + 384, tmem + kb * 8, make_desc_mn((uint32_t)(db[0] & 0x3FFF) * 16 + kb * 1024), idesc(64) | (1u << 16), 1u);    // TS N64, B MN-major
}
template <int PAT> __device__ __forceinline__ void run_pat(long long* out, int n, uint32_t tmem, uint32_t sa, uint32_t sb, uint32_t bar, uint32_t& phase) {
    for (int rep = 0; rep < 3; ++rep) {
        long long t0 = 0, t1 = 0;
        if (threadIdx.x == 0) {
            uint64_t da[8], db[8];
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) { da[kb] = make_desc(sa + kb * 2 * LBO, LBO); db[kb] = make_desc(sb + kb * 2 * LBO, LBO); }
            t0 = clock64();
            for (int i0 = 0; i0 < n; i0 += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) issue_one<PAT>(u, tmem, da, db);
            }
            mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        if (threadIdx.x == 0) { t1 = clock64(); if (blockIdx.x == 0) out[PAT * 3 + rep] = t1 - t0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
}
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int npat, int n) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ __align__(8) unsigned long long barw;
    for (int i = threadIdx.x; i < 2 * 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    const uint32_t bar = smem_u32(&barw);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    uint32_t phase = 0;
    run_pat<0>(out, n, tmem, sa, sb, bar, phase); run_pat<1>(out, n, tmem, sa, sb, bar, phase); run_pat<2>(out, n, tmem, sa, sb, bar, phase);
    run_pat<3>(out, n, tmem, sa, sb, bar, phase); run_pat<4>(out, n, tmem, sa, sb, bar, phase); run_pat<5>(out, n, tmem, sa, sb, bar, phase);
    run_pat<6>(out, n, tmem, sa, sb, bar, phase); run_pat<7>(out, n, tmem, sa, sb, bar, phase); run_pat<8>(out, n, tmem, sa, sb, bar, phase);
    run_pat<9>(out, n, tmem, sa, sb, bar, phase); run_pat<10>(out, n, tmem, sa, sb, bar, phase); run_pat<11>(out, n, tmem, sa, sb, bar, phase);
    run_pat<12>(out, n, tmem, sa, sb, bar, phase);
    run_pat<13>(out, n, tmem, sa, sb, bar, phase); run_pat<14>(out, n, tmem, sa, sb, bar, phase); run_pat<15>(out, n, tmem, sa, sb, bar, phase);
    run_pat<16>(out, n, tmem, sa, sb, bar, phase); run_pat<17>(out, n, tmem, sa, sb, bar, phase);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    const int npat = 18;
    long long* d; cudaMalloc(&d, npat * 3 * 2 * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536);
    long long h[2][npat * 3];
    const int ns[2] = {48, 240};
    for (int k = 0; k < 2; ++k) {
        rate_kernel<<<148, 128, 2 * 65536>>>(d, npat, ns[k]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h[k], d, npat * 3 * sizeof(long long), cudaMemcpyDeviceToHost);
    }
    const char* names[npat] = {"SS N128 one acc", "SS N128 two accs", "TS N128 one acc", "TS N128 two accs", "TS N64 one acc", "TS N64 two accs",
                               "TS N64 four accs", "fwd pattern (TS N128 + TS N64)", "adj pattern (main,small,small N64)", "SS N64 one acc",
                               "TS N256", "TS N192", "TS N32", "SS N128 A,B MN-major (BASE32B)", "SS N128 A MN-major", "SS N128 B MN-major",
                               "TS N128 B MN-major", "TS N64 B MN-major"};
    for (int p = 0; p < npat; ++p) {
        const double per = (double)(h[1][p * 3 + 2] - h[0][p * 3 + 2]) / (ns[1] - ns[0]);
        printf("%-38s  n=48: %6lld cyc  n=240: %6lld cyc  -> %.1f cyc/MMA (148 CTAs, all SMs busy)\n", names[p], h[0][p * 3 + 2], h[1][p * 3 + 2], per);
    }
    return 0;
}
