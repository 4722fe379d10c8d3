// tc_probe.cu — round-2 groundwork: validates hand-built tcgen05 descriptors on sm_100a.
//   D[128x64] (fp32, TMEM) = A[128x64] * B[64x64]^T-or-not, kind::tf32, cta_group::1, M=128, N=64, K=8 per MMA.
// Operands live in shared memory in the NO-SWIZZLE canonical layouts (cute/atom/mma_traits_sm100.hpp:
// make_umma_desc): 8x16B core matrices,
//   K-major : addr(r,k) = (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4          (r = M or N index)
//   MN-major: addr(n,k) = (n/4)*SBO + (k/8)*LBO + (k%8)*16 + (n%4)*4
// so one buffer written as "K-major over (k, n)" is the MN-major operand of the transposed product with
// LBO/SBO swapped — the trick DESIGN.md §7 relies on to share W between A*W and Zbar*W^T.
// Modes: 0 = K-major A, K-major B (B given as [N][K]);  1 = K-major A, MN-major B (B given as W[K][N]);
//        2 = mode 1 with the 3xTF32 split (hi*hi + hi*lo + lo*hi): FP32-level accuracy check.
// All mbarrier waits are bounded; the host prints max errors against an FP64 reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;        // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;        // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version = 1 (Blackwell) [46,48)
    // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int a_major_mn, int b_major_mn) {
    return (1u << 4)            // c_format = F32
         | (2u << 7)            // a_format = TF32
         | (2u << 10)           // b_format = TF32
         | ((uint32_t)a_major_mn << 15) | ((uint32_t)b_major_mn << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128) tc_probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ Dg,
                                               int mode, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                       // 32 KB  (hi)
    float* sB = reinterpret_cast<float*>(smem + 32768);               // 16 KB  (hi)
    float* sAl = reinterpret_cast<float*>(smem + 49152);              // 32 KB  (lo, mode 2)
    float* sBl = reinterpret_cast<float*>(smem + 81920);              // 16 KB  (lo, mode 2)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 98304 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    constexpr uint32_t A_SBO = 128, A_LBO = (M / 8) * 128;            // K-major A: 16 row groups per 16B K chunk
    constexpr uint32_t BK_SBO = 128, BK_LBO = (N / 8) * 128;          // K-major B
    constexpr uint32_t BM_LBO = 128, BM_SBO = (K / 8) * 128;          // MN-major B: (n/4)*SBO + (k/8)*LBO + (k%8)*16 + (n%4)*4

    auto split = [](float x, float& hi, float& lo) {
        hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);      // what the tensor core keeps of an FP32 bit pattern
        lo = x - hi;
    };
    for (int idx = tid; idx < M * K; idx += 128) {
        const int m = idx / K, k = idx % K;
        float hi = Ag[idx], lo = 0.f;
        if (mode == 2) split(Ag[idx], hi, lo);
        const int off = ((k / 4) * A_LBO + (m / 8) * A_SBO + (m % 8) * 16 + (k % 4) * 4) / 4;
        sA[off] = hi; sAl[off] = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        int n, k;
        if (mode == 0 || mode == 2) { n = idx / K; k = idx % K; } else { k = idx / N; n = idx % N; }      // Bg is [N][K] or W[K][N]
        float hi = Bg[idx], lo = 0.f;
        if (mode == 2) split(Bg[idx], hi, lo);
        const int off = ((mode == 0 || mode == 2) ? ((k / 4) * BK_LBO + (n / 8) * BK_SBO + (n % 8) * 16 + (k % 4) * 4)
                                   : ((n / 4) * BM_SBO + (k / 8) * BM_LBO + (k % 8) * 16 + (n % 4) * 4)) / 4;
        sB[off] = hi; sBl[off] = lo;
    }
    if (mode >= 3) {
        // address-mapping probe: A selects k = m % 64 (identity), the B buffer encodes its own location:
        // mode 3 -> index of the 16-byte unit, mode 4 -> position inside the unit; D[m][n] then shows which
        // shared-memory word the MN-major descriptor maps (n, k=m) to.
        __syncthreads();
        for (int idx = tid; idx < M * K; idx += 128) {
            const int m = idx / K, k = idx % K;
            sA[((k / 4) * A_LBO + (m / 8) * A_SBO + (m % 8) * 16 + (k % 4) * 4) / 4] = (k == (m % 64)) ? 1.f : 0.f;
        }
        // the pattern covers all shared memory behind sA (64 KB), so reads outside the 16 KB operand still show up
        for (int w = tid; w < (98304 - 32768) / 4; w += 128) sB[w] = (mode == 3 || mode == 5) ? (float)(w / 4 + 1) : (float)(w % 4 + 1);   // +1: zero means "nothing read"
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const bool kmaj = (mode == 0 || mode == 2);
        const uint32_t idesc = make_idesc(0, kmaj ? 0 : 1);
        const int nterm = mode == 2 ? 3 : 1;
        int first = 1;
        for (int term = 0; term < nterm; ++term) {
            const float* a = (term == 2) ? sAl : sA;                  // hi*hi, hi*lo, lo*hi
            const float* b = (term == 1) ? sBl : sB;
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t da = make_desc(smem_u32(a) + kb * 2 * A_LBO, A_LBO, A_SBO);
                const uint64_t db = kmaj ? make_desc(smem_u32(b) + kb * 2 * BK_LBO, BK_LBO, BK_SBO)
                                  : (mode >= 5 ? make_desc(smem_u32(b) + kb * BM_LBO, BM_SBO, BM_LBO)      // swapped roles
                                               : make_desc(smem_u32(b) + kb * BM_LBO, BM_LBO, BM_SBO));
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // bounded wait on phase 0
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) status[0] = -1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        uint32_t r[64];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[c + 0]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]),
                           "=r"(r[c + 6]), "=r"(r[c + 7]), "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]),
                           "=r"(r[c + 12]), "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                         : "r"(taddr + c));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = warp * 32 + lane;
        for (int n = 0; n < N; ++n) Dg[m * N + n] = __uint_as_float(r[n]);
        if (tid == 0) status[0] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    float *hA = new float[M * K], *hB = new float[N * K], *hW = new float[K * N], *hD = new float[M * N];
    srand(7);
    for (int i = 0; i < M * K; ++i) hA[i] = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float v = (rand() / (float)RAND_MAX) * 2.f - 1.f; hB[n * K + k] = v; hW[k * N + n] = v; }
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
    const size_t smem = 98304 + 64;
    cudaFuncSetAttribute(tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 7; ++mode) {
        cudaMemcpy(dB, (mode == 0 || mode == 2) ? hB : hW, N * K * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, M * N * 4); cudaMemset(dS, 0, 4);
        tc_probe<<<1, 128, smem>>>(dA, dB, dD, mode, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        double errTrunc = 0, errExact = 0, ref_max = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double st32 = 0, sx = 0;
            for (int k = 0; k < K; ++k) { st32 += (double)tf32_trunc(hA[m * K + k]) * tf32_trunc(hB[n * K + k]); sx += (double)hA[m * K + k] * hB[n * K + k]; }
            errTrunc = fmax(errTrunc, fabs(hD[m * N + n] - st32)); errExact = fmax(errExact, fabs(hD[m * N + n] - sx)); ref_max = fmax(ref_max, fabs(sx));
        }
        if (mode >= 3) {
            printf("mode %d (%s of the word read for (n, k)%s):\n", mode, (mode == 3 || mode == 5) ? "16B-unit index+1" : "position in unit+1", mode >= 5 ? ", LBO/SBO swapped" : "");
            for (int k = 0; k < 10; ++k) { printf("  k=%2d:", k); for (int n = 0; n < 10; ++n) printf(" %5.0f", hD[k * N + n]); printf("\n"); }
            continue;
        }
        printf("mode %d: cuda=%s status=%d  max|D-ref_tf32trunc|=%.3e  max|D-ref_exact|=%.3e  (max|ref|=%.2f)\n", mode, cudaGetErrorString(e), st, errTrunc, errExact, ref_max);
        if (e != cudaSuccess) break;
    }
    return 0;
}
