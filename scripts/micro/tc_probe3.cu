// tc_probe3.cu — MN-major TF32 with LayoutType::SWIZZLE_128B_BASE32B (Swizzle<2,5,2>, 4-row k atoms), derived from tc_probe2.cu — round-2 probes for the width-64 tile kernel (vn_tc64.cu), sm_100a.
//
// (1) Does tcgen05.mma kind::tf32 accept MN-major (transposed) shared-memory operands, and in which canonical layout?
//     The weight-gradient GEMM gW = A^T Zbar contracts over POINTS, while every thread of the tile kernel owns one point
//     (a row): a K-major operand needs a transposing scatter (64 scalar STS per thread and step), an MN-major operand is
//     the natural row-major [point][neuron] tile written with 16-byte stores.  Round 1 found the NO-swizzle MN-major
//     form to return zeros; this probe tries the swizzled canonical layouts of cute/atom/mma_traits_sm100.hpp
//     (make_umma_desc<Major::MN>):  in 16-byte units
//         SW128: Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO))      32 elements of MN contiguous, 8 k-rows of 128 B
//         SW64 : Swizzle<2,4,3> o ((4,n),(8,k)):((1,LBO),(4,SBO))
//         SW32 : Swizzle<1,4,3> o ((2,n),(8,k)):((1,LBO),(2,SBO))
//     D[128 x 64] = A[128 x 64] * B, B given as W[k][n]; also both operands MN-major (the gW shape: D = At^T * Bt).
// (2) Tensor-memory load throughput per SM (tcgen05.ld.32x32b.x16 from 4 / 8 / 16 warps), the other quantity that
//     sizes the epilogues of the tile kernel.
// All mbarrier waits are bounded.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe2 tc_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (mn, k) of an MN-major operand: rows of `rowB` bytes (k index), chunks XOR-swizzled inside a row
__host__ __device__ inline uint32_t mn_off(int mn, int k, int rowB, int Ktot, int swzBits, int atomK, int swzBase) {
    const int epr = rowB / 4;                                   // elements of MN per row
    const int g = mn / epr, e = mn % epr;
    const uint32_t sbo = atomK * rowB, lbo = (Ktot / atomK) * sbo;
    uint32_t off = g * lbo + (k / atomK) * sbo + (k % atomK) * rowB + e * 4;
    // Swizzle<B, base, 7 - base>: XOR address bits [base, base+B) with bits [7, 7+B)
    const uint32_t mask = ((1u << swzBits) - 1u);
    off ^= ((off >> 7) & mask) << swzBase;
    return off;
}

struct Mode { int aMN, bMN, rowB, swzBits, layout, atomK, swzBase, swapLS; const char* name; };

__global__ void __launch_bounds__(128) probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ Dg,
                                            Mode md, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                       // 32 KB
    unsigned char* sB = smem + 32768;               // 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 49152 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t A_SBO = 128, A_LBO = (M / 8) * 128, BK_SBO = 128, BK_LBO = (N / 8) * 128;
    // A: Ag is [M][K] (value A[m][k]); B: Bg is W[K][N] (value B[k][n])
    for (int idx = tid; idx < M * K; idx += 128) {
        const int m = idx / K, k = idx % K;
        const uint32_t off = md.aMN ? mn_off(m, k, md.rowB, K, md.swzBits, md.atomK, md.swzBase) : (uint32_t)((k / 4) * A_LBO + (m / 8) * A_SBO + (m % 8) * 16 + (k % 4) * 4);
        *reinterpret_cast<float*>(sA + off) = Ag[idx];
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int k = idx / N, n = idx % N;
        const uint32_t off = md.bMN ? mn_off(n, k, md.rowB, K, md.swzBits, md.atomK, md.swzBase) : (uint32_t)((k / 4) * BK_LBO + (n / 8) * BK_SBO + (n % 8) * 16 + (k % 4) * 4);
        *reinterpret_cast<float*>(sB + off) = Bg[idx];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(md.aMN, md.bMN, N);
        const uint32_t sbo = md.atomK * md.rowB, lbo = (K / md.atomK) * sbo, kstep = 8 * md.rowB;
        const uint32_t dl = md.swapLS ? sbo : lbo, ds = md.swapLS ? lbo : sbo;
        for (int kb = 0; kb < K / 8; ++kb) {
            // MN-major: LBO = stride between MN groups, SBO = stride between 8-row k groups (swizzled forms); the no-swizzle
            // form swaps the roles (make_umma_desc).  One MMA covers one k group.
            uint64_t da, db;
            if (md.aMN) da = md.layout == 0 ? make_desc(smem_u32(sA) + kb * kstep, sbo, lbo, 0) : make_desc(smem_u32(sA) + kb * kstep, dl, ds, md.layout);
            else da = make_desc(smem_u32(sA) + kb * 2 * A_LBO, A_LBO, A_SBO, 0);
            if (md.bMN) db = md.layout == 0 ? make_desc(smem_u32(sB) + kb * kstep, sbo, lbo, 0) : make_desc(smem_u32(sB) + kb * kstep, dl, ds, md.layout);
            else db = make_desc(smem_u32(sB) + kb * 2 * BK_LBO, BK_LBO, BK_SBO, 0);
            const uint32_t acc = kb ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    if (!done && tid == 0) status[0] = -1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        uint32_t r[64];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 64; c += 16)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[c + 0]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7]),
                           "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]), "=r"(r[c + 12]), "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                         : "r"(taddr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = warp * 32 + lane;
        for (int n = 0; n < N; ++n) Dg[m * N + n] = __uint_as_float(r[n]);
        if (tid == 0) status[0] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// ---- tensor-memory load throughput: `nwarps` warps each issue `iters` x (tcgen05.ld.32x32b.x16 [+ wait]) on one SM
__global__ void __launch_bounds__(512) tmem_bw(int iters, int batch, long long* cycles, float* sink) {
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tq = tslot + ((uint32_t)(32 * (warp & 3)) << 16) + 16 * (warp >> 2);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int b = 0; b < batch; ++b) {
            uint32_t u[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                           "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                         : "r"(tq + 64 * ((it + b) & 3)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(u[0] ^ u[7] ^ u[15]);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    float *hA = new float[M * K], *hW = new float[K * N], *hD = new float[M * N];
    srand(7);
    for (int i = 0; i < M * K; ++i) hA[i] = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (int i = 0; i < K * N; ++i) hW[i] = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hW, N * K * 4, cudaMemcpyHostToDevice);
    const size_t smem = 49152 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const Mode modes[] = {
        {0, 0, 16, 0, 0, 8, 4, 0, "A K-major, B K-major (control)"},
        {0, 1, 128, 3, 2, 8, 4, 0, "B MN-major, SWIZZLE_128B (16 B base)"},
        {0, 1, 128, 2, 1, 4, 5, 0, "B MN-major, SWIZZLE_128B_BASE32B"},
        {1, 0, 128, 2, 1, 4, 5, 0, "A MN-major, SWIZZLE_128B_BASE32B"},
        {1, 1, 128, 2, 1, 4, 5, 0, "A and B MN-major, SWIZZLE_128B_BASE32B (weight-gradient shape)"},
        {1, 1, 128, 2, 1, 4, 5, 1, "A and B MN-major, SWIZZLE_128B_BASE32B, LBO/SBO swapped"},
        {1, 1, 128, 2, 1, 8, 5, 0, "A and B MN-major, BASE32B with 8-row k atoms"},
    };
    for (const Mode& md : modes) {
        cudaMemset(dD, 0, M * N * 4); cudaMemset(dS, 0, 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, md, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        double err = 0, refmax = 0, dmax = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) {
                // K-major control: the B buffer is indexed (n, k) <- W[k][n] as well, so every mode computes A * W
                const float a = md.aMN ? hA[m * K + k] : hA[m * K + k];
                s += (double)tf32_trunc(a) * tf32_trunc(hW[k * N + n]);
            }
            err = fmax(err, fabs(hD[m * N + n] - s)); refmax = fmax(refmax, fabs(s)); dmax = fmax(dmax, fabs((double)hD[m * N + n]));
        }
        printf("%-70s cuda=%s status=%d  max|D-ref|=%.3e  max|D|=%.3f max|ref|=%.3f  %s\n", md.name, cudaGetErrorString(e), st, err, dmax, refmax,
               err < 1e-3 ? "OK" : "MISMATCH");
        if (e != cudaSuccess) break;
    }
    return 0;
}
