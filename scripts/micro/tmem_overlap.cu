// Microbenchmark: do tensor-memory loads of the worker warps proceed while tcgen05.mma instructions of the same CTA are in flight?
//   one thread issues `nmma` MMAs (kind::tf32, 128x128x8, SS, accumulator columns [384,512)) and commits them to an mbarrier;
//   then all 16 warps read `nld` x (tcgen05.ld.32x32b.x16 + wait::ld) from columns [0,256) (not touched by the MMAs).
// Printed: cycles until the loads are done / until the MMAs are done, for (MMAs only), (loads only), (both).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_overlap tmem_overlap.cu && ./tmem_overlap
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
__device__ __forceinline__ uint32_t idesc(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t aT, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(aT), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

__global__ void __launch_bounds__(512, 1) overlap_kernel(long long* out, int nmma, int nld, int ts, int stores) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ __align__(8) unsigned long long barw;
    for (int i = threadIdx.x; i < 2 * 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    const uint32_t bar = smem_u32(&barw);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    const uint32_t tq = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 16 * (warp >> 2);
    uint32_t phase = 0;
    float accv = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (threadIdx.x == 0 && nmma > 0) {
            for (int i = 0; i < nmma; ++i) {
                const int kb = i & 7;
                if (ts) mma_ts(tmem + 384, tmem + 256 + kb * 8, make_desc(sb + kb * 2 * LBO, LBO), idesc(128), 1u);
                else mma_ss(tmem + 384, make_desc(sa + kb * 2 * LBO, LBO), make_desc(sb + kb * 2 * LBO, LBO), idesc(128), 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        for (int it = 0; it < nld; ++it) {
            uint32_t u[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                           "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                         : "r"(tq + 64 * (it & 3)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            accv += __uint_as_float(u[0] ^ u[7] ^ u[15]);
            if (stores) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(tq + 64 * (it & 3)), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                               "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        const long long t1 = clock64();
        if (nmma > 0) { while (!mbar_test(bar, phase)) {} phase ^= 1u; }
        const long long t2 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 511)) { out[(threadIdx.x ? 6 : 0) + rep * 2] = t1 - t0; out[(threadIdx.x ? 6 : 0) + rep * 2 + 1] = t2 - t0; }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (accv == 123.456f) out[15] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 16 * sizeof(long long));
    cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536);
    struct Case { int nmma, nld, ts, st; const char* name; };
    const Case cases[] = {
        {40, 0, 0, 0, "40 SS MMAs only"}, {40, 0, 1, 0, "40 TS MMAs only"},
        {0, 8, 0, 0, "8 x (ld x16 + wait) per warp only"}, {0, 8, 0, 1, "8 x (ld + st x16) per warp only"},
        {40, 8, 0, 0, "40 SS MMAs + 8 ld per warp"}, {40, 8, 1, 0, "40 TS MMAs + 8 ld per warp"},
        {40, 8, 0, 1, "40 SS MMAs + 8 (ld + st) per warp"}, {40, 8, 1, 1, "40 TS MMAs + 8 (ld + st) per warp"},
        {40, 2, 1, 0, "40 TS MMAs + 2 ld per warp"}, {40, 32, 1, 0, "40 TS MMAs + 32 ld per warp"},
    };
    for (const Case& c : cases) {
        cudaMemset(d, 0, 16 * sizeof(long long));
        overlap_kernel<<<148, 512, 2 * 65536>>>(d, c.nmma, c.nld, c.ts, c.st);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-40s thread 0 (issuer): loads done %6lld, MMAs done %6lld | thread 511: loads done %6lld, MMAs done %6lld cycles  (%s)\n",
               c.name, h[4], h[5], h[10], h[11], cudaGetErrorString(e));
    }
    return 0;
}
