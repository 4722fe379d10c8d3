// vn_tc64.cu — resident-tile tensor-core kernel for hidden widths 33..64 (see vn_tc64.h).
//
// Tensor memory map (512 columns x 128 lanes, lane = quadrature point of the tile):
//   region s = [128 s, 128 s + 128), s < S <= 3 : operand of stream s, hi TF32 half in columns +0..63, lo half in +64..127
//                                                 (forward: a_{l,s}; adjoint: zbar_{l,s}); between two adjoint layers its
//                                                 first 64 columns park abar_{l-1,s} (the drained layer-GEMM result)
//   work     = [384, 512)                        : accumulators of the running GEMM: hi*hi in +0..63 (a chain of 8 MMAs),
//                                                 hi*lo + lo*hi in +64..127; the weight-gradient GEMM uses all 128 columns
// The tensor core accumulates with truncation (vn_tc.cu): chains of hi*hi products are kept at <= 16 MMAs and the
// 2^-11-sized cross products go to their own columns; the sets are summed in FP32 round-to-nearest by the drain.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "vn_tc64.h"

namespace {

constexpr int TP = 128;            // points per tile == TMEM lanes
constexpr int W = 64;              // padded hidden width
constexpr int NH = 4;              // neuron groups: warp w -> lane quarter w & 3, neuron group w >> 2
constexpr int CPT = W / NH;        // neurons (columns) per thread
constexpr int NT = 128 * NH;       // threads
constexpr int FOLD = 512;          // tiles accumulated in the FP32 window slab before the FP64 fold (VARNET_B200_TC64_FOLD)
constexpr uint32_t SBO = 128;      // bytes between 8-row groups of a canonical K-major operand
constexpr uint32_t W_LBO = 2048;   // weight images: 128 rows x 16 B per 4-wide K unit
constexpr uint32_t G_LBO = 2064;   // weight-gradient operands (K = points): + 16 B pad, transposing stores hit 32 banks
constexpr int WIMG_FLOATS = 8192, WIMG_BYTES = 32768;
constexpr int G_BYTES = 32 * G_LBO;
constexpr int OFF_WST = 0;                         // two weight stages
constexpr int OFF_GA = 2 * WIMG_BYTES;             // [a_hi^T ; a_lo^T]     128 rows x 128 points
constexpr int OFF_GB = OFF_GA + G_BYTES;           // [zbar_hi^T ; zbar_lo^T]
constexpr int OFF_PAR = OFF_GB + G_BYTES;
// MN-major weight-gradient operands (GWM > 0): row = point (128 B = 32 neurons), LayoutType::SWIZZLE_128B_BASE32B
// (Swizzle<2,5,2>: the 32-byte chunk index of a row is XORed with point & 3; k atoms of 4 points = 512 B), four MN groups
// of 32 rows of the 128-row operand (hi 0..31, hi 32..63, lo 0..31, lo 32..63), 16 KB each.  This is the only canonical
// layout in which tcgen05.mma kind::tf32 accepts MN-major operands (scripts/micro/tc_probe3.cu; CUTLASS sm100_common.inl:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout").
constexpr uint32_t MN_GROUP = 16384, MN_ATOM = 512, MN_KSTEP = 1024;
constexpr int OFF_MGA = 2 * WIMG_BYTES;            // 64 KB
constexpr int OFF_MGB = OFF_MGA + 4 * MN_GROUP;    // 64 KB
constexpr int PAR_FLOATS = 1152 + NH * 384 + 640 + 512 + 2 * VN_KIN * 128;  // W0[8][64] | bias[8][64] | wout[64] | bout.. | usP[NH][3][128] | us[3][128] | I[128] | R[128] | coef[4][128] | x[2][VN_KIN][128]
constexpr int OFF_BAR = OFF_PAR + PAR_FLOATS * 4;
constexpr int SMEM_BYTES = OFF_BAR + 64;
constexpr uint32_t COL_WORK = 384;

struct Tc64Args { TileArgs t; const float* wimg; int* err; long long* timing; int fwdOnly; int fold; int coefPre; };      // timing: optional [16] phase cycle counters of CTA 0 (VARNET_B200_TC64_TIMING); fwdOnly: loss / lossVec / R only (v1 schedule)

struct SlabLayout { int vecOff, boutOff, psz, nkind; };
__host__ __device__ inline SlabLayout slab_layout(int L, int inpDim) {
    SlabLayout s;
    s.vecOff = (L - 1) * (TP * W);                 // gw blocks of layers 1..L-1: [64/4 column quads][128 rows][4] (see gw_slot)
    s.nkind = L + 1 + inpDim;                      // gb_0..gb_{L-1} | g(w_out) | gW_0 rows
    s.boutOff = s.vecOff + s.nkind * 4 * W;        // vec slots: [kind][lane quarter][64]
    s.psz = s.boutOff + 4;
    return s;
}

// slot of gW-block element (row r of the 128 accumulator rows, column j): quad-major, so that the 32 threads of a warp
// (consecutive rows, the same column quad) touch 512 contiguous bytes per 16-byte reduction
__host__ __device__ inline int gw_slot(int l, int r, int j) { return (l - 1) * (TP * W) + (((j >> 2) * TP + r) << 2) + (j & 3); }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;                                      // no swizzle
}
// D = F32, A = B = TF32, K-major, M = 128
template <int N> struct IDesc { static constexpr uint32_t v = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((MN_GROUP >> 4) & 0x3FFF) << 16;        // leading byte offset: between MN groups of 32 elements
    d |= (uint64_t)((MN_ATOM >> 4) & 0x3FFF) << 32;         // stride byte offset: between k atoms of 4 points
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                 // SWIZZLE_128B_BASE32B
    return d;
}
constexpr uint32_t IDESC_GW_MN = IDesc<128>::v | (1u << 15) | (1u << 16);     // A and B MN-major
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t aT, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(aT), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a lost commit must not hang the GPU
    uint32_t done = 0;
    long long t0 = 0;
    for (int spin = 0; !done; ++spin) {
        // suspend-time hint (ns): a waiting warp sleeps in the barrier unit instead of spinning through the issue slots of
        // the three working warps that share its scheduler; it is woken as soon as the phase completes
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (!done && (spin & 63) == 63) {                      // ~0.1 s at 2 GHz, measured on the clock (a try_wait may itself suspend)
            const long long t = clock64();
            if (t0 == 0) t0 = t; else if (t - t0 > 200000000ll) break;
        }
    }
    return done != 0;
}
// one lane of a converged warp (elect.sync): ptxas then knows the MMA operands are uniform and emits back-to-back UTCHMMA instead
// of an ELECT / BRA.U.ANY loop around each of them
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// activation of this kernel class: tanh as 1 - 2 / (exp(2|x|) + 1) on the special-function unit (7 instructions against ~25
// for tanhf; absolute error ~1e-7, the size of the FP32 rounding of the pre-activation itself); sigmoid as in vn_tile.cuh
template <int ACT> __device__ __forceinline__ float act64(float z) {
    if (ACT == VN_SIGMOID) return act_f<VN_SIGMOID>(z);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(z) * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), z);
}
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// ------------------------------------------------------------------ per-thread operand movement (thread = point p x neurons c0..c0+CPT-1)
// hi/lo split of v -> the thread's columns of operand region `reg` (address of the lane quarter, first column of the region)
__device__ __forceinline__ void put_operand(uint32_t reg, int c0, const float (&v)[CPT]) {
#pragma unroll
    for (int hf = 0; hf < CPT / 16; ++hf) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { hi[i] = tf32_rn(v[16 * hf + i]); lo[i] = v[16 * hf + i] - hi[i]; }
        tmem_st16(reg + c0 + 16 * hf, hi);
        tmem_st16(reg + 64 + c0 + 16 * hf, lo);
    }
}
// hi + lo == the FP32 value exactly
__device__ __forceinline__ void get_operand(uint32_t reg, int c0, float (&v)[CPT]) {
#pragma unroll
    for (int hf = 0; hf < CPT / 16; ++hf) {
        float hi[16], lo[16];
        tmem_ld16(reg + c0 + 16 * hf, hi);
        tmem_ld16(reg + 64 + c0 + 16 * hf, lo);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 * hf + i] = hi[i] + lo[i];
    }
}
__device__ __forceinline__ void put_plain(uint32_t taddr, const float (&v)[CPT]) {
    float t[16];
#pragma unroll
    for (int hf = 0; hf < CPT / 16; ++hf) {
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = v[16 * hf + i];
        tmem_st16(taddr + 16 * hf, t);
    }
}
__device__ __forceinline__ void get_plain(uint32_t taddr, float (&v)[CPT]) {
#pragma unroll
    for (int hf = 0; hf < CPT / 16; ++hf) {
        float t[16];
        tmem_ld16(taddr + 16 * hf, t);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 * hf + i] = t[i];
    }
}
// two accumulator sets `a` and `b` (addresses of the thread's first column) -> their FP32 sum
__device__ __forceinline__ void drain_sum2(uint32_t a, uint32_t b, float (&z)[CPT]) {
#pragma unroll
    for (int hf = 0; hf < CPT / 16; ++hf) {
        float m[16], s[16];
        tmem_ld16(a + 16 * hf, m);
        tmem_ld16(b + 16 * hf, s);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) z[16 * hf + i] = m[i] + s[i];
    }
}
// transposing store: rows = neurons (hi rows 0..63, lo rows 64..127), K = points; a warp's 32 points of one neuron
// fall into 32 different banks (G_LBO = 2048 + 16)
__device__ __forceinline__ void put_transposed(unsigned char* G, int p, int c0, const float (&v)[CPT]) {
    unsigned char* base = G + (p >> 2) * G_LBO + (p & 3) * 4 + (c0 >> 3) * SBO;
#pragma unroll
    for (int jj = 0; jj < CPT; ++jj) {
        const float hi = tf32_rn(v[jj]);
        *reinterpret_cast<float*>(base + (jj >> 3) * SBO + (jj & 7) * 16) = hi;
        *reinterpret_cast<float*>(base + (8 + (jj >> 3)) * SBO + (jj & 7) * 16) = v[jj] - hi;
    }
}
// MN-major store of the same operand: the thread's 16 neurons of point p are 64 contiguous bytes of row p (two 32-byte
// chunks, swizzled with p & 3); hi image in MN group h >> 1, lo image two groups further.  With 16-byte stores the lanes p
// and p + 4 of a quarter warp would hit the same bank group (their rows are 512 B apart and carry the same swizzle), so
// for GWM == 2 the lanes with p & 4 hold their register quads pairwise exchanged (`w` quad u = neuron quad u ^ 1: `v` is
// exchanged in place by quad_swap, the stash rows are loaded that way) and store quad slot u to the other half of its
// chunk: every store instruction then covers 8 different 16-byte bank groups per quarter warp.
__device__ __forceinline__ void quad_swap(float (&v)[CPT], bool sw) {
#pragma unroll
    for (int cc = 0; cc < CPT / 8; ++cc)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[8 * cc + i], b = v[8 * cc + 4 + i];
            v[8 * cc + i] = sw ? b : a;
            v[8 * cc + 4 + i] = sw ? a : b;
        }
}
__device__ __forceinline__ void put_mn(unsigned char* G, int p, int h, const float (&w)[CPT], bool sw) {
    unsigned char* row = G + (h >> 1) * MN_GROUP + p * 128;
    const uint32_t x = (uint32_t)(p & 3) << 5, o = 64u * (h & 1), f = sw ? 16u : 0u;
#pragma unroll
    for (int u = 0; u < CPT / 4; ++u) {
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { hi[i] = tf32_rn(w[4 * u + i]); lo[i] = w[4 * u + i] - hi[i]; }
        unsigned char* c = row + (((o + 32u * (u >> 1)) ^ x) + ((16u * (u & 1)) ^ f));
        *reinterpret_cast<float4*>(c) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(c + 2 * MN_GROUP) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}
// stash: [(l*S + s)][neuron/4][point][4].  The 148 per-CTA slabs (57 MB at 4x64) are rewritten by every tile: they are
// written and read with an L2 evict_last policy (and the streamed point table with evict_first) so that they stay in the
// 126 MB L2 instead of being written back to HBM tile after tile.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ float ldg_stream(const float* ptr, uint64_t pol) {
    float v; asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(pol)); return v;
}
__device__ __forceinline__ void stash_put(float* st, int slab, int p, int c0, const float (&v)[CPT], uint64_t pol) {
    float4* b = reinterpret_cast<float4*>(st) + ((size_t)slab * 16 + (c0 >> 2)) * TP + p;
#pragma unroll
    for (int u = 0; u < CPT / 4; ++u)
        asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(b + u * TP), "f"(v[4 * u]), "f"(v[4 * u + 1]),
                     "f"(v[4 * u + 2]), "f"(v[4 * u + 3]), "l"(pol) : "memory");
}
__device__ __forceinline__ void stash_get(const float* st, int slab, int p, int c0, float (&v)[CPT], uint64_t pol) {
    const float4* b = reinterpret_cast<const float4*>(st) + ((size_t)slab * 16 + (c0 >> 2)) * TP + p;
#pragma unroll
    for (int u = 0; u < CPT / 4; ++u)
        asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v[4 * u]), "=f"(v[4 * u + 1]), "=f"(v[4 * u + 2]),
                     "=f"(v[4 * u + 3]) : "l"(b + u * TP), "l"(pol) : "memory");
}
// the same rows with the quads pairwise exchanged (register quad u <- neuron quad u ^ 1) where `sw` (put_mn)
__device__ __forceinline__ void stash_get_sw(const float* st, int slab, int p, int c0, float (&v)[CPT], uint64_t pol, bool sw) {
    const float4* b = reinterpret_cast<const float4*>(st) + ((size_t)slab * 16 + (c0 >> 2)) * TP + p;
    const int d = sw ? TP : 0;
#pragma unroll
    for (int u = 0; u < CPT / 4; ++u)
        asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v[4 * u]), "=f"(v[4 * u + 1]), "=f"(v[4 * u + 2]),
                     "=f"(v[4 * u + 3]) : "l"(b + u * TP + ((u & 1) ? -d : d)), "l"(pol) : "memory");
}
// sum over the 32 lanes of v[c], c < CPT: the lanes with index >> COLSUM_SHIFT == c return column c
__device__ __forceinline__ float warp_colsum(float (&v)[CPT], int lane) {
    int n = CPT;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            const bool up = (lane & off) != 0;
            const int hn = n >> 1;
#pragma unroll
            for (int i = 0; i < CPT / 2; ++i) {
                if (i < hn) {
                    const float send = up ? v[i] : v[i + hn];
                    const float keep = up ? v[i + hn] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            n = hn;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    return v[0];
}
constexpr int COLSUM_SHIFT = (CPT == 32) ? 0 : (CPT == 16 ? 1 : 2);      // column of the returned sum = lane >> COLSUM_SHIFT

// ------------------------------------------------------------------ the kernel
// Tensor-memory columns (see the header comment).  Forward sweep: the S operand regions + one work accumulator.  Adjoint
// sweep: one operand region, the parked abar_{l-1,s} (= the hi*hi accumulators of the layer GEMM, fixed up in place
// with the cross products), the cross-product columns and the weight-gradient accumulator, so that the layer GEMM and
// the weight-gradient GEMM of a step are issued together and waited for once.
constexpr uint32_t COL_OP = 0;          // adjoint: zbar_{l,s} hi | lo
constexpr uint32_t COL_PARK = 128;      //          + 64 s: abar_{l-1,s}
constexpr uint32_t COL_SMALL = 320;     //          hi*lo + lo*hi of the running layer GEMM
constexpr uint32_t COL_GW = 384;        //          weight-gradient accumulator (128 columns)

template <int S, int ACT, int VAR>
__global__ void __launch_bounds__(NT, 1) tc64_var_kernel(const __grid_constant__ Tc64Args K) {
    constexpr int GWM = VAR == 3 ? 0 : VAR;         // weight-gradient operand layout (gw_mode)
    constexpr bool EARLY = VAR == 3;                // [a_hi ; a_lo] of the NEXT step is written under the running layer GEMM
    extern __shared__ __align__(1024) unsigned char smem[];
    const TileArgs& A = K.t;
    const NetDesc& net = A.net;
    const int L = net.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, h = warp >> 2;
    const int p = 32 * q + lane;                    // point of the tile == TMEM lane
    const int c0 = CPT * h;                         // first of this thread's CPT neurons
    const bool sw = GWM == 2 && (p & 4);            // this lane keeps the weight-gradient operands quad-exchanged (put_mn)

    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    float* W0s = par;                               // [8][64]
    float* bs = par + 512;                          // [8][64]
    float* wout = par + 1024;                       // [64]
    float* misc = par + 1088;                       // [0] = b_out
    float* usP = par + 1152;                        // [NH][3][128] output-layer partial dots of the neuron groups
    float* us = usP + NH * 384;                          // [3][128] u_s, then the adjoint seeds
    float* Ish = us + 384;                          // [128]
    float* Rsh = Ish + 128;                         // [128]
    float* coefS = Rsh + 128;                       // [4][128] integrand coefficients of the tile's points (gcoef_0, gcoef_1, dNt, source*N)
    float* xS = coefS + 512;                        // [2][VN_KIN][128] MLP inputs of this tile's and of the CTA's next tile's points
    const bool xPre = (K.coefPre & 2) != 0;
    const uint32_t bar = smem_u32(smem + OFF_BAR);            // forward GEMMs / adjoint layer GEMM
    const uint32_t barGw = bar + 8;                             // weight-gradient GEMM
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
    unsigned char* GA = smem + (GWM ? OFF_MGA : OFF_GA);
    unsigned char* GB = smem + (GWM ? OFF_MGB : OFF_GB);

    if (*reinterpret_cast<volatile int*>(K.err)) return;
    for (int i = tid; i < PAR_FLOATS; i += NT) par[i] = 0.f;
    __syncthreads();
    {
        const float* __restrict__ th = A.theta;
        const int w0 = net.width[0];
        for (int idx = tid; idx < net.inpDim * w0; idx += NT) { const int c = idx / w0, j = idx - c * w0; W0s[c * W + j] = th[net.woff[0] + idx]; }
        for (int l = 0; l < L; ++l)
            for (int j = tid; j < net.width[l]; j += NT) bs[l * W + j] = th[net.boff[l] + j];
        for (int j = tid; j < net.width[L - 1]; j += NT) wout[j] = th[net.woff[L] + j];
        if (tid == 0) {
            misc[0] = th[net.boff[L]];
            mbar_init(bar, 1);
            mbar_init(barGw, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16);           // this warp's lane quarter

    const uint64_t polLast = policy_evict_last(), polFirst = policy_evict_first();
    const SlabLayout sl = slab_layout(L, net.inpDim);
    float* part = A.part32 + (size_t)blockIdx.x * sl.psz;
    double* part64 = A.part + (size_t)blockIdx.x * sl.psz;
    float* stash = A.stash + (size_t)blockIdx.x * A.stashFloats;
    const int nImg = K.fwdOnly ? (L - 1) : 2 * (L - 1);      // forward-only launches cycle through the forward images only
    uint32_t phase = 0, phaseGw = 0;
    bool ok = true;
    int imgNext = 0;                                 // next weight image of the tile sequence to be consumed
    double lossAcc = 0.0;
    bool first = true, firstFold = !A.accumulate;
    int win = 0;

    // weight image n of the per-tile sequence: forward layers 1..L-1, then adjoint layers L-1..1
    auto image_of = [&](int n) { return n < L - 1 ? 2 * n : 2 * (2 * L - 3 - n) + 1; };
    int stage = 0;                                   // shared-memory stage of the image in flight (the stages alternate; with an
                                                     // odd image count per tile - forward-only launches - n & 1 would not)
    auto prefetch_image = [&](int n, int st) {
        const float* src = K.wimg + (size_t)image_of(n) * WIMG_FLOATS;
        float* dst = reinterpret_cast<float*>(smem + OFF_WST + st * WIMG_BYTES);
#pragma unroll
        for (int i = 0; i < WIMG_BYTES / 16 / NT; ++i) cp_async16(dst + 4 * (i * NT + tid), src + 4 * (i * NT + tid));
        cp_async_commit();
    };
    // the image of this layer is complete (issued one layer earlier); put the next one in flight into the other stage,
    // whose last readers (the MMAs of the previous layer) have completed
    auto acquire_image = [&]() -> uint32_t {
        cp_async_wait_all();
        fence_async_smem();
        const int n = imgNext, st = stage;
        imgNext = (n + 1 == nImg) ? 0 : n + 1;
        stage ^= 1;
        prefetch_image(imgNext, stage);
        return smem_u32(smem + OFF_WST + st * WIMG_BYTES);
    };
    // publish this thread's TMEM / shared-memory writes and finished TMEM reads, then let thread 0 issue
    auto sync_for_issue = [&]() {
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    auto wait_mma = [&]() {
        if (ok) ok = mbar_wait(bar, phase);
        phase ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    auto wait_gw = [&]() {
        if (ok) ok = mbar_wait(barGw, phaseGw);
        phaseGw ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    // forward layer GEMM of stream s: work = A_s x [W_hi | W_lo]^T (hi*hi | hi*lo), then the second half += A_s,lo x W_hi
    auto issue_fwd = [&](int s, uint32_t wst) {
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t aHi = tmem + 128 * s, aLo = aHi + 64, d = tmem + COL_WORK;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                const uint64_t db = make_desc(wst + kb * 2 * W_LBO, W_LBO);
                mma_ts(d, aHi + kb * 8, db, IDesc<128>::v, kb ? 1u : 0u);
                mma_ts(d + 64, aLo + kb * 8, db, IDesc<64>::v, 1u);
            }
            mma_commit(bar);
        }
    };
    // adjoint step of stream s: abar_{l-1,s} = zbar_{l,s} W_l^T (hi*hi -> park s, cross products -> small) and
    // gW_l (+)= [a_hi ; a_lo]^T-rows x [zbar_hi ; zbar_lo]^T-rows over the 128 points of the tile
    // (the weight-gradient GEMM goes first and has its own barrier: its accumulators are drained under the layer GEMM)
    auto issue_adj = [&](int s, uint32_t wst, uint32_t gwAcc, bool gwLast) {
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t ga = smem_u32(GA), gb = smem_u32(GB);
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {
                if (GWM) mma_ss(tmem + COL_GW, make_desc_mn(ga + kb * MN_KSTEP), make_desc_mn(gb + kb * MN_KSTEP), IDESC_GW_MN, kb ? 1u : gwAcc);
                else mma_ss(tmem + COL_GW, make_desc(ga + kb * 2 * G_LBO, G_LBO), make_desc(gb + kb * 2 * G_LBO, G_LBO), IDesc<128>::v, kb ? 1u : gwAcc);
            }
            if (gwLast || EARLY) mma_commit(barGw);     // drained after the last stream only; earlier streams are covered by `bar`
            const uint32_t aHi = tmem + COL_OP, aLo = aHi + 64, dMain = tmem + COL_PARK + 64 * s, dSmall = tmem + COL_SMALL;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                const uint64_t dbh = make_desc(wst + kb * 2 * W_LBO, W_LBO);
                const uint64_t dbl = make_desc(wst + 8 * SBO + kb * 2 * W_LBO, W_LBO);          // rows 64..127: the lo half
                mma_ts(dMain, aHi + kb * 8, dbh, IDesc<64>::v, kb ? 1u : 0u);
                mma_ts(dSmall, aHi + kb * 8, dbl, IDesc<64>::v, kb ? 1u : 0u);
                mma_ts(dSmall, aLo + kb * 8, dbh, IDesc<64>::v, 1u);
            }
            mma_commit(bar);
        }
    };
    // this warp's sum over its 32 points of column c0+lane -> the warp's vec slot of `kind`
    auto vec_add = [&](int kind, float v, bool overwrite) {
        if (lane & ((1 << COLSUM_SHIFT) - 1)) return;
        float* slot = part + sl.vecOff + (kind * 4 + q) * W + c0 + (lane >> COLSUM_SHIFT);
        if (overwrite) __stcg(slot, v); else atomicAdd(slot, v);
    };

    prefetch_image(0, 0);

    // MLP inputs of the 128 points of `tile` -> xS[buf]: streamed tables with 4-byte cp.async (no register, no wait), everything else
    // (extra inputs, in-kernel generated coordinates, index-list batches) computed and stored; called one tile ahead
    auto stage_inputs = [&](int tile, int buf) {
        for (int idx = tid; idx < net.inpDim * TP; idx += NT) {
            const int c = idx / TP, pp = idx - c * TP;
            const unsigned int g = (unsigned int)(A.tile0 + tile) * TP + pp;
            float* dst = xS + (buf * VN_KIN + c) * TP + pp;
            if (c >= A.nxTable) { *dst = __ldg(A.extraX + (c - A.nxTable)); continue; }
            if (A.useGen) {
                float val = 0.f;
                if (g < A.P) {
                    const unsigned int bb = g / A.integNum;
                    const int q2 = (int)(g - bb * A.integNum);
                    const long long i2 = A.gen.tf0 + (long long)table_tf(A, bb);
                    const long long s2 = i2 / A.gen.nTime, j2 = i2 - s2 * A.gen.nTime;
                    const double ctr = c < A.gen.dim ? __ldg(A.gen.coord + s2 * A.gen.dim + c) : __ldg(A.gen.tcoord + j2);
                    val = __double2float_rn(__dadd_rn(ctr, __ldg(A.gen.hd + (size_t)c * A.gen.q + q2)));
                }
                *dst = val;
                continue;
            }
            const float* col = A.cols + (size_t)(A.colX + c) * A.pstride;
            if (!A.tfIndex) {                           // zero padded table
                asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;\n" ::"r"(smem_u32(dst)), "l"(col + g), "l"(polFirst) : "memory");   // streamed once: evict_first, the stash slabs keep the L2
                continue;
            }
            *dst = g < A.P ? __ldg(col + table_row(A, g)) : 0.f;
        }
        cp_async_commit();                              // cp_async_wait_all() waits for committed groups only
    };
    int xbuf = 0;
    if (xPre && (int)blockIdx.x < A.ntiles) stage_inputs(blockIdx.x, 0);

    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x, xbuf ^= 1) {
        if (xPre) {                                     // this tile's inputs (staged one tile ago) have landed and are visible
            cp_async_wait_all();
            __syncthreads();
        }
        const unsigned int base = (unsigned int)(A.tile0 + tile) * TP;
        const unsigned int gp = base + p;
        const bool valid = gp < A.P;
        const size_t row = (valid && !A.useGen) ? table_row(A, gp) : 0;
        // in-kernel table generation: this point's test function (space node gs, time node gj) and Gauss index gq
        long long gs = 0, gj = 0;
        int gq = 0;
        if (A.useGen && valid) {
            const unsigned int b = gp / A.integNum;
            gq = (int)(gp - b * A.integNum);
            if (!xPre) {                                 // with staged inputs only the Gauss index is needed here
                const long long i = A.gen.tf0 + (long long)table_tf(A, b);
                gs = i / A.gen.nTime; gj = i - gs * A.gen.nTime;
            }
        }
        auto coefv = [&](int k) -> float {               // 0, 1: gcoef; 2: dNt; 3: source*N
            if (A.useGen) return __ldg(A.gen.coef + gq * 4 + k);
            const int col = k < 2 ? A.colG + k : (k == 2 ? A.colT : A.colS);
            if (!A.tfIndex) return ldg_stream(A.cols + (size_t)col * A.pstride + row, polFirst);      // streamed once
            return __ldg(A.cols + (size_t)col * A.pstride + row);
        };
        auto input = [&](int c) -> float {
            if (xPre) return xS[(xbuf * VN_KIN + c) * TP + p];
            if (c >= A.nxTable) return __ldg(A.extraX + (c - A.nxTable));
            if (A.useGen) {
                if (!valid) return 0.f;
                const double ctr = c < A.gen.dim ? __ldg(A.gen.coord + gs * A.gen.dim + c) : __ldg(A.gen.tcoord + gj);
                return __double2float_rn(__dadd_rn(ctr, __ldg(A.gen.hd + (size_t)c * A.gen.q + gq)));
            }
            if (!A.tfIndex) return ldg_stream(A.cols + (size_t)(A.colX + c) * A.pstride + gp, polFirst);  // zero padded table
            return valid ? __ldg(A.cols + (size_t)(A.colX + c) * A.pstride + row) : 0.f;
        };

        // table columns of the CTA's next tile -> L2 (the table streams from HBM once; 24 lines of 128 B per tile)
        if (!A.tfIndex && !A.useGen && tile + (int)gridDim.x < A.ntiles) {
            const int nc = A.nxTable + (S - 1) + (A.colT >= 0 ? 1 : 0) + (A.colS >= 0 ? 1 : 0);
            if (tid < nc * 4) {
                const float* nx = A.cols + (size_t)(A.colX + (tid >> 2)) * A.pstride + (size_t)(A.tile0 + tile + (int)gridDim.x) * TP + (tid & 3) * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            }
        }

        // integrand coefficient h of point p: requested now, parked in shared memory at the end of the forward sweep, so that the
        // integrand / seed phases (128 threads, the tensor core idle) do not wait for global memory
        float cfr = 0.f;
        if ((K.coefPre & 1) && (h < S - 1 || (h == 2 && A.timeDependent) || (h == 3 && A.isSource))) cfr = coefv(h);

        // ---- inputs and layer 0 (K = inpDim: FP32 FMA).  Stream 1+k is seeded with the unit vector e_k.
        float d1[CPT];                                // act'(z_l) of the value stream, kept for the tangent streams
        {
            float v[CPT];
#pragma unroll
            for (int jj = 0; jj < CPT; ++jj) v[jj] = bs[c0 + jj];
            for (int c = 0; c < net.inpDim; ++c) {
                const float xc = input(c);
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) v[jj] = fmaf(xc, W0s[c * W + c0 + jj], v[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < CPT; ++jj) { v[jj] = act64<ACT>(v[jj]); d1[jj] = act_d1<ACT>(v[jj]); }
            put_operand(tq, c0, v);
            if (!K.fwdOnly) stash_put(stash, 0, p, c0, v, polLast);
            for (int k = 0; k < S - 1; ++k) {
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) v[jj] = d1[jj] * W0s[k * W + c0 + jj];
                put_operand(tq + 128 * (1 + k), c0, v);
                if (!K.fwdOnly) stash_put(stash, 1 + k, p, c0, v, polLast);
            }
        }

        // ---- hidden layers, forward: the GEMM of the next (layer, stream) step is issued before this step's epilogue
        // arithmetic, so the tensor core runs under the activation functions
        {
            const int nfs = (L - 1) * S;
            uint32_t wst = acquire_image();
            sync_for_issue();
            issue_fwd(0, wst);
            // inputs of the CTA's next tile -> the other shared-memory buffer, while the first GEMM of this tile runs
            if (xPre && tile + (int)gridDim.x < A.ntiles) stage_inputs(tile + (int)gridDim.x, xbuf ^ 1);
            int l = 1, s = 0;
            for (int k = 0; k < nfs; ++k) {
                wait_mma();
                float v[CPT];
                drain_sum2(tq + COL_WORK + c0, tq + COL_WORK + 64 + c0, v);
                int ln = l, sn = s + 1;
                if (sn == S) { sn = 0; ln = l + 1; }
                if (k + 1 < nfs) {
                    if (sn == 0) wst = acquire_image();
                    sync_for_issue();
                    issue_fwd(sn, wst);
                }
                if (s == 0) {
#pragma unroll
                    for (int jj = 0; jj < CPT; ++jj) {
                        v[jj] = act64<ACT>(v[jj] + bs[l * W + c0 + jj]);
                        d1[jj] = act_d1<ACT>(v[jj]);
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < CPT; ++jj) v[jj] *= d1[jj];
                }
                put_operand(tq + 128 * s, c0, v);
                if (!K.fwdOnly) stash_put(stash, l * S + s, p, c0, v, polLast);
                if (l == L - 1) {
                    // output layer (Dense(1)): partial dot over this thread's neurons
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int jj = 0; jj < CPT; jj += 2) { a0 = fmaf(v[jj], wout[c0 + jj], a0); a1 = fmaf(v[jj + 1], wout[c0 + jj + 1], a1); }
                    usP[(h * 3 + s) * TP + p] = a0 + a1;
                }
                l = ln; s = sn;
            }
        }
        if (K.coefPre & 1) coefS[h * TP + p] = cfr;
        tmem_wait_st();
        __syncthreads();

        // ---- u_s, integrand, R_i of the test functions of this tile, loss, adjoint seeds (TFModel.py:653-664)
        if (tid < TP) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float u = 0.f;
#pragma unroll
                for (int g = 0; g < NH; ++g) u += usP[(g * 3 + s) * TP + p];
                if (s == 0) u += misc[0];
                us[s * TP + p] = u;
            }
            float I = 0.f;
            if (valid) {
#pragma unroll
                for (int k = 0; k < S - 1; ++k) I = fmaf(us[(1 + k) * TP + p], (K.coefPre & 1) ? coefS[k * TP + p] : coefv(k), I);
                if (A.timeDependent) I -= us[p] * ((K.coefPre & 1) ? coefS[2 * TP + p] : coefv(2));
                if (A.isSource) I -= ((K.coefPre & 1) ? coefS[3 * TP + p] : coefv(3));
                if (A.integW) I *= __ldg(A.integW + (gp % A.integNum));
            }
            Ish[p] = I;
        }
        __syncthreads();
        {
            const int nf = TP / (int)A.integNum;
            for (int f = warp; f < nf; f += NT / 32) {
                float r = 0.f;
                for (int qq = lane; qq < (int)A.integNum; qq += 32) r += Ish[f * A.integNum + qq];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                if (lane == 0) {
                    const unsigned int i = base / A.integNum + f;
                    Rsh[f] = r;
                    if (i * A.integNum < A.P) {
                        const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                        const float r2 = r * r;
                        A.R[i] = r;
                        A.lossVec[i] = dj * r2;
                        lossAcc += A.detJvec ? (double)dj * (double)r2 : (double)r2;
                    }
                }
            }
        }
        __syncthreads();
        if (K.fwdOnly) continue;                       // loss-only pass (splitLoss, trainWeight): R_i, lossVec and the loss partial are done
        if (tid < TP) {
            float lam = 0.f;
            if (valid) {
                const unsigned int i = gp / A.integNum, qq = gp - i * A.integNum;
                const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                const float wq = A.integW ? __ldg(A.integW + qq) : 1.f;
                lam = 2.f * __ldg(A.wts + 2) * dj * wq * Rsh[p / A.integNum];
            }
            us[p] = A.timeDependent ? -lam * ((K.coefPre & 1) ? coefS[2 * TP + p] : coefv(2)) : 0.f;
#pragma unroll
            for (int k = 0; k < S - 1; ++k) us[(1 + k) * TP + p] = lam * ((K.coefPre & 1) ? coefS[k * TP + p] : coefv(k));
        }
        __syncthreads();

        // ---- output layer gradients: g(b_out) = sum_p ubar_0, g(w_out)[j] = sum_{s,p} a_{L-1,s}[p][j] ubar_s[p]
        // (a_{L-1,s} is still in the forward operand regions)
        if (h == 0) {
            float sb = us[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
            if (lane == 0) { float* slot = part + sl.boutOff + q; if (first) __stcg(slot, sb); else atomicAdd(slot, sb); }
        }
        {
            float gwo[CPT];
#pragma unroll
            for (int jj = 0; jj < CPT; ++jj) gwo[jj] = 0.f;
            for (int s = 0; s < S; ++s) {
                float a[CPT];
                get_operand(tq + 128 * s, c0, a);
                const float ub = us[s * TP + p];
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) gwo[jj] = fmaf(a[jj], ub, gwo[jj]);
            }
            const float r = warp_colsum(gwo, lane);
            vec_add(L, r, first);
        }

        // ---- adjoint sweep: step (l, s) turns abar_{l,s} into zbar_{l,s} (tangent streams first: the value stream needs
        // their second-order term), then abar_{l-1,s} = zbar_{l,s} W_l^T and gW_l += a_{l-1,s}^T zbar_{l,s}
        {
            float a0[CPT], dpre[CPT], apre[CPT], cross[CPT];
            const bool L0MERGE = (K.coefPre & 4) != 0;
            bool gaStored = false;
            stash_get(stash, (L - 1) * S, p, c0, a0, polLast);
            stash_get(stash, (L - 1) * S + 1, p, c0, dpre, polLast);
            stash_get_sw(stash, (L - 2) * S + 1, p, c0, apre, polLast, sw);
            for (int l = L - 1; l >= 0; --l) {
                uint32_t wst = 0;
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) cross[jj] = 0.f;
                for (int si = 0; si < S; ++si) {
                    const int s = (si < S - 1) ? si + 1 : 0;
                    float v[CPT];
                    if (l == L - 1) {
                        const float ub = us[s * TP + p];
#pragma unroll
                        for (int jj = 0; jj < CPT; ++jj) v[jj] = ub * wout[c0 + jj];
                    } else {
                        get_plain(tq + COL_PARK + 64 * s + c0, v);
                    }
                    if (s > 0) {
#pragma unroll
                        for (int jj = 0; jj < CPT; ++jj) { cross[jj] = fmaf(v[jj], dpre[jj], cross[jj]); v[jj] *= act_d1<ACT>(a0[jj]); }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < CPT; ++jj) v[jj] = fmaf(v[jj], act_d1<ACT>(a0[jj]), act_d2r<ACT>(a0[jj]) * cross[jj]);
                    }
                    // next step (ln, sn)
                    int ln = l, sn = (si + 1 < S - 1) ? si + 2 : 0;
                    if (si + 1 == S) { ln = l - 1; sn = 1; }
                    if (l >= 1) {
                        if (si == 0) wst = acquire_image();
                        put_operand(tq + COL_OP, c0, v);
                        if (GWM) {
                            put_mn(GA, p, h, apre, sw);
                            if (GWM == 2) {
                                quad_swap(v, sw);
                                put_mn(GB, p, h, v, sw);
                                if (s == 0) quad_swap(v, sw);
                            } else {
                                put_mn(GB, p, h, v, false);
                            }
                        } else {
                            put_transposed(GB, p, c0, v);
                            if (!EARLY || !gaStored) put_transposed(GA, p, c0, apre);
                        }
                        if (s == 0) {
#pragma unroll
                            for (int jj = 0; jj < CPT; ++jj) a0[jj] = apre[jj];                  // a_{l-1,0}: the value activations of the next layer down
                            if (GWM == 2) quad_swap(a0, sw);                                     // apre holds it quad-exchanged
                            if (!EARLY) {
                                const float r = warp_colsum(v, lane);                            // g(b_l) = sum_p zbar_{l,0}
                                vec_add(l, r, first);
                            }
                        }
                        sync_for_issue();
                        issue_adj(s, wst, si > 0 ? 1u : 0u, si == S - 1);
                        if (EARLY && s == 0) {                                                   // registers only: under the weight-gradient GEMM
                            const float r = warp_colsum(v, lane);
                            vec_add(l, r, first);
                        }
                        // stash rows of the next step, in flight while the tensor core runs
                        if (ln >= 0) {
                            if (sn > 0) stash_get(stash, ln * S + sn, p, c0, dpre, polLast);
                            if (ln >= 1) stash_get_sw(stash, (ln - 1) * S + sn, p, c0, apre, polLast, sw);
                        }
                        if (EARLY) {
                            // the weight-gradient MMAs of this step have read both operands: the next step's a-operand (its rows are
                            // in registers by now) goes to shared memory while the layer GEMM still runs
                            wait_gw();
                            gaStored = ln >= 1;
                            if (ln >= 1) put_transposed(GA, p, c0, apre);
                        }
                        if (si == S - 1) {
                            if (!EARLY) wait_gw();
                            // the S streams of a layer are ONE K = S*128 contraction (a chain of 16 S <= 48 hi*hi MMAs), drained once:
                            // rows 0..63: a_hi (x) [zbar_hi | zbar_lo]; rows 64..127: a_lo (x) zbar_hi (lo x lo dropped)
                            float g[CPT];
                            if (q < 2) drain_sum2(tq + COL_GW + c0, tq + COL_GW + 64 + c0, g); else get_plain(tq + COL_GW + c0, g);
                            float* slot = part + gw_slot(l, p, c0);
                            if (first) {                                       // first write of this window: overwrite
#pragma unroll
                                for (int u = 0; u < CPT / 4; ++u) __stcg(reinterpret_cast<float4*>(slot) + u * TP, make_float4(g[4 * u], g[4 * u + 1], g[4 * u + 2], g[4 * u + 3]));
                            } else {
#pragma unroll
                                for (int u = 0; u < CPT / 4; ++u) red_add_v4(slot + 4 * u * TP, g[4 * u], g[4 * u + 1], g[4 * u + 2], g[4 * u + 3]);
                            }
                        }
                        wait_mma();
                        {
                            float t[CPT];
                            drain_sum2(tq + COL_PARK + 64 * s + c0, tq + COL_SMALL + c0, t);
                            put_plain(tq + COL_PARK + 64 * s + c0, t);                  // park abar_{l-1,s} in place
                        }
                        tmem_wait_st();
                    } else {
                        // layer 0: gb_0 = sum_p zbar_{0,0}; gW_0[c] = sum_p x_c zbar_{0,0} (+ sum_p zbar_{0,1+c} for the spatial inputs)
                        // a_{0,sn} of the next tangent stream: recomputed from a_{0,0} (in a0) exactly as the forward sweep computed it
                        // (act'(a_{0,0}) W_0[sn-1][:]: the same bits) instead of an exposed L2 round trip to the stash
                        if (ln >= 0 && sn > 0) {
                            if (L0MERGE) {
#pragma unroll
                                for (int jj = 0; jj < CPT; ++jj) dpre[jj] = act_d1<ACT>(a0[jj]) * W0s[(sn - 1) * W + c0 + jj];
                            } else {
                                stash_get(stash, sn, p, c0, dpre, polLast);
                            }
                        }
                        if (L0MERGE) {
                            // zbar of tangent stream 1 + c waits in the (idle) operand columns of tensor memory and joins the column sum
                            // of input row c: inpDim + 1 transposing warp reductions per tile instead of inpDim + S
                            if (s > 0) {
                                put_plain(tq + COL_OP + 64 * (s - 1) + c0, v);
                                tmem_wait_st();
                            } else {
                                for (int c = 0; c < net.inpDim; ++c) {
                                    const float xc = input(c);
                                    float t[CPT];
                                    if (c < S - 1) {
                                        get_plain(tq + COL_OP + 64 * c + c0, t);
#pragma unroll
                                        for (int jj = 0; jj < CPT; ++jj) t[jj] = fmaf(xc, v[jj], t[jj]);
                                    } else {
#pragma unroll
                                        for (int jj = 0; jj < CPT; ++jj) t[jj] = xc * v[jj];
                                    }
                                    const float r = warp_colsum(t, lane);
                                    vec_add(L + 1 + c, r, first);
                                }
                                const float r = warp_colsum(v, lane);
                                vec_add(0, r, first);
                            }
                        } else if (s > 0) {
                            const float r = warp_colsum(v, lane);
                            vec_add(L + 1 + (s - 1), r, first);
                        } else {
                            for (int c = 0; c < net.inpDim; ++c) {
                                const float xc = input(c);
                                float t[CPT];
#pragma unroll
                                for (int jj = 0; jj < CPT; ++jj) t[jj] = xc * v[jj];
                                const float r = warp_colsum(t, lane);
                                vec_add(L + 1 + c, r, first && c >= S - 1);
                            }
                            const float r = warp_colsum(v, lane);
                            vec_add(0, r, first);
                        }
                    }
                }
            }
        }
        tmem_wait_ld();

        first = false;
        if (++win == K.fold || tile + (int)gridDim.x >= A.ntiles) {
            // fold this thread's slots of the FP32 window into the FP64 slab (single writer per slot)
            auto put = [&](int idx) {
                const double v = (double)__ldcg(part + idx);
                if (firstFold) __stcg(part64 + idx, v); else part64[idx] += v;
            };
            for (int l = 1; l < L; ++l)
                for (int jj = 0; jj < CPT; ++jj) put(gw_slot(l, p, c0 + jj));
            if ((lane & ((1 << COLSUM_SHIFT) - 1)) == 0)
                for (int k = 0; k < sl.nkind; ++k) put(sl.vecOff + (k * 4 + q) * W + c0 + (lane >> COLSUM_SHIFT));
            if (h == 0 && lane == 0) put(sl.boutOff + q);
            firstFold = false; first = true; win = 0;
        }
    }
    cp_async_wait_all();
    if (lane == 0) {
        double* lp = A.lossPart + blockIdx.x * (NT / 32) + warp;
        *lp = A.accumulate ? *lp + lossAcc : lossAcc;
    }
    if (!ok) *K.err = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ==================================================================================================================
// v2: warp-specialised, software-pipelined version of the tile kernel (round 2).
//
//  * 17 warps: 16 worker warps (thread = point x 16 neurons, as above) + ONE issuing warp.  The workers never execute
//    MMA-issue code and there is no __syncthreads on the issue path: workers -> issuer hand-offs are mbarriers with one
//    arrival per worker warp, issuer -> workers hand-offs are tcgen05.commit -> mbarrier.
//  * every FP32 product still comes from three TF32 MMAs, but all three now accumulate into ONE 64-column accumulator,
//    cross products first: while only the 2^-11-sized terms have been added the accumulator is small and the tensor
//    core's truncating accumulate costs nothing; the chain of full-size hi*hi additions stays at 8 MMAs.  No separate
//    cross-product columns, no drain-and-add fix-up: a forward step reads 16 columns once, and the adjoint layer GEMM
//    leaves abar_{l-1,s} parked in tensor memory in its final form.
//  * forward: two accumulators ping-pong, so the GEMM of step k+1 runs under the epilogue of step k.
//  * adjoint: per step the weight-gradient GEMM is issued first, the layer GEMM second.  While the tensor core runs
//    step k the workers already compute zbar of step k+1 from the parked abar (under the weight-gradient GEMM), then
//    drain the weight-gradient accumulator and rewrite the transposed shared-memory operands (under the layer GEMM; the
//    operand pair is 129 KB, a second one does not fit), and store the tensor-memory operand once the layer GEMM is done.
//  * the weight images are brought in by the issuing warp with one 32 KB bulk copy (cp.async.bulk -> mbarrier) each;
//    the inputs and integrand coefficients of the CTA's next tile are prefetched into shared memory with cp.async.
//  * g(w_out) needs sum_s a_{L-1,s} * coef_s per point: it is accumulated in registers by the epilogues of the last
//    layer and scaled by lambda once R_i is known, instead of re-reading the three operands from tensor memory.
//
// Tensor-memory columns.  Forward: operand of stream s [128 s, 128 s + 128) (hi | lo), accumulators [384,448), [448,512).
// Adjoint: operand [0,128), parked abar_{l-1,s} [128 + 64 s, +64), weight-gradient accumulator [384,512).
namespace v2 {

constexpr int NWORK = 512;                         // worker threads (16 warps)
constexpr int NT2 = NWORK + 128;                   // + the issuing warp's warpgroup (registers are allocated per 4 warps: a lone
                                                   //   17th warp would cost as much as four and cap the kernel at 96 registers)
constexpr int NXIN = 12;                           // prefetched per-point values: <= 8 inputs | dNt | gcoef_0..1 | source*N
constexpr int OFF2_WST = 0;
constexpr int OFF2_G = 2 * WIMG_BYTES;             // GA | GB (64.5 KB each: hi and lo rows; no room for a second pair)
constexpr int OFF2_PAR = OFF2_G + 2 * G_BYTES;
static_assert(OFF2_PAR + (1152 + NH * 384 + 768 + 2 * NXIN * TP) * 4 + 160 + 128 <= 232448, "shared memory of the v2 tile kernel");
constexpr int PAR2_FLOATS = 1152 + NH * 384 + 768 + 2 * NXIN * TP;   // v1 block | lam[128] | xin[2][NXIN][128]
constexpr int OFF2_BAR = OFF2_PAR + PAR2_FLOATS * 4;
constexpr int OFF2_TIM = OFF2_BAR + 160;             // 16 x int64 phase timers of thread 0 (debug)
constexpr int SMEM2_BYTES = OFF2_TIM + 128;
constexpr uint32_t COL2_ACC = 384;                 // forward accumulators: + 64 (k & 1)
constexpr uint32_t COL2_CROSS = 320;               // adjoint: second-order term sum_k abar_k (.) adot_k of the running layer (FP32)
enum { B_STEPF = 0, B_STEPA = 2, B_ACC = 4, B_LAYER = 6, B_GW = 7, B_WFULL = 8, B_WFREE = 10, NBAR = 12 };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int S, int ACT>
__global__ void __launch_bounds__(NT2, 1) tc64_var_kernel_v2(const __grid_constant__ Tc64Args K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TileArgs& A = K.t;
    const NetDesc& net = A.net;
    const int L = net.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float* par = reinterpret_cast<float*>(smem + OFF2_PAR);
    float* W0s = par;                               // [8][64]
    float* bs = par + 512;                          // [8][64]
    float* wout = par + 1024;                       // [64]
    float* misc = par + 1088;                       // [0] = b_out
    float* usP = par + 1152;                        // [NH][3][128] output-layer partial dots of the neuron groups
    float* us = usP + NH * 384;                     // [3][128] u_s, then the adjoint seeds
    float* Ish = us + 384;                          // [128]
    float* Rsh = Ish + 128;                         // [128]
    float* lamS = Rsh + 128;                        // [128]
    float* xin = lamS + 128;                        // [2][NXIN][128]
    const uint32_t bar0 = smem_u32(smem + OFF2_BAR);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF2_BAR + 8 * NBAR);

    if (*reinterpret_cast<volatile int*>(K.err)) return;
    for (int i = tid; i < PAR2_FLOATS; i += NT2) par[i] = 0.f;
    __syncthreads();
    {
        const float* __restrict__ th = A.theta;
        const int w0 = net.width[0];
        for (int idx = tid; idx < net.inpDim * w0; idx += NT2) { const int c = idx / w0, j = idx - c * w0; W0s[c * W + j] = th[net.woff[0] + idx]; }
        for (int l = 0; l < L; ++l)
            for (int j = tid; j < net.width[l]; j += NT2) bs[l * W + j] = th[net.boff[l] + j];
        for (int j = tid; j < net.width[L - 1]; j += NT2) wout[j] = th[net.woff[L] + j];
        if (tid == 0) {
            misc[0] = th[net.boff[L]];
            for (int i = 0; i < NBAR; ++i) {
                const bool fromWorkers = i < B_ACC;
                mbar_init(BAR(i), fromWorkers ? 16u : 1u);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const int nfs = (L - 1) * S;                    // MMA steps of a sweep: (layer 1..L-1) x streams
    const int nImg = 2 * (L - 1);
    bool ok = true;

    if (warp >= NWORK / 32) {
        // ============================================================ issuing warp (one elected lane) + 3 idle warps of its warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");         // hand the registers to the workers
        if (warp == NWORK / 32 && elect_one()) {
            uint32_t phase = 0;                     // bit i: phase parity this thread waits for next on barrier i
            int nLoaded = 0, nUsed = 0, rLoad = 0;  // weight images of the whole launch: loaded / consumed (stage = n & 1); rLoad = nLoaded % nImg
            const int nTotal = ((A.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nImg;
            auto wait = [&](int i) { if (ok) ok = mbar_wait(BAR(i), (phase >> i) & 1u); phase ^= 1u << i; };
            auto load_next = [&]() {                // image nLoaded -> stage nLoaded & 1 (its previous user: image nLoaded - 2)
                if (nLoaded >= nTotal) return;
                const int st = nLoaded & 1;
                if (nLoaded >= 2) wait(B_WFREE + st);
                const int img = rLoad < L - 1 ? 2 * rLoad : 2 * (2 * L - 3 - rLoad) + 1;     // forward layers 1..L-1, then adjoint layers L-1..1
                mbar_expect_tx(BAR(B_WFULL + st), WIMG_BYTES);
                bulk_g2s(smem_u32(smem + OFF2_WST + st * WIMG_BYTES), K.wimg + (size_t)img * WIMG_FLOATS, WIMG_BYTES, BAR(B_WFULL + st));
                ++nLoaded;
                if (++rLoad == nImg) rLoad = 0;
            };
            auto acquire = [&]() -> uint32_t {      // image nUsed is in shared memory
                const int st = nUsed & 1;
                wait(B_WFULL + st);
                ++nUsed;
                return smem_u32(smem + OFF2_WST + st * WIMG_BYTES);
            };
            // d (64 columns) = A x W as 3xTF32, cross products first: lo*hi, hi*lo, then the chain of 8 hi*hi
            auto issue_layer = [&](uint32_t aHi, uint32_t wst, uint32_t d) {
                const uint32_t aLo = aHi + 64;
#pragma unroll
                for (int kb = 0; kb < 8; ++kb) mma_ts(d, aLo + kb * 8, make_desc(wst + kb * 2 * W_LBO, W_LBO), IDesc<64>::v, kb ? 1u : 0u);
#pragma unroll
                for (int kb = 0; kb < 8; ++kb) mma_ts(d, aHi + kb * 8, make_desc(wst + 8 * SBO + kb * 2 * W_LBO, W_LBO), IDesc<64>::v, 1u);
#pragma unroll
                for (int kb = 0; kb < 8; ++kb) mma_ts(d, aHi + kb * 8, make_desc(wst + kb * 2 * W_LBO, W_LBO), IDesc<64>::v, 1u);
            };
            load_next();
            for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
                // ---- forward: event j = 0: layer 0 done; j >= 1: epilogue of step j-1 done.  Step k needs event max(0, k-1).
                uint32_t wst = 0;
                for (int j = 0; j <= nfs; ++j) {
                    wait(B_STEPF + (j & 1));
                    tc_fence_after();
                    const int k0 = j == 0 ? 0 : j + 1, k1 = j == 0 ? 1 : j + 1;
                    for (int k = k0; k <= k1 && k < nfs; ++k) {
                        const int s = k % S;
                        if (s == 0) wst = acquire();
                        issue_layer(tmem + 128 * s, wst, tmem + COL2_ACC + 64 * (k & 1));
                        mma_commit(BAR(B_ACC + (k & 1)));
                        if (s == S - 1) mma_commit(BAR(B_WFREE + ((nUsed - 1) & 1)));
                        if (s == 0) load_next();
                    }
                }
                // ---- adjoint: step k = (layer L-1-k/S, stream order 1..S-1, 0).  The workers have drained the weight-gradient
                // accumulator of step k-1 and rewritten both operand sets before they signal step k.
                for (int k = 0; k < nfs; ++k) {
                    const int si = k % S, s = (si < S - 1) ? si + 1 : 0;
                    wait(B_STEPA + (k & 1));
                    tc_fence_after();
                    const uint32_t ga = smem_u32(smem + OFF2_G), gb = ga + G_BYTES;
#pragma unroll
                    for (int kb = 0; kb < 16; ++kb)
                        mma_ss(tmem + COL_GW, make_desc(ga + kb * 2 * G_LBO, G_LBO), make_desc(gb + kb * 2 * G_LBO, G_LBO), IDesc<128>::v, kb ? 1u : 0u);
                    mma_commit(BAR(B_GW));
                    if (si == 0) wst = acquire();
                    issue_layer(tmem + COL_OP, wst, tmem + COL_PARK + 64 * s);
                    mma_commit(BAR(B_LAYER));
                    if (si == S - 1) mma_commit(BAR(B_WFREE + ((nUsed - 1) & 1)));
                    if (si == 0) load_next();
                }
            }
            if (!ok) *K.err = 1;
        }
        __syncwarp();
    } else {
        // ============================================================ worker warps
        // the pool only holds what the issuing warpgroup gave back: 640 x 96 at launch = 512 x 112 + 128 x 24 + 1024 spare
        // (asking for 120 would need more than was released and block forever)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        const int q = warp & 3, h = warp >> 2;
        const int p = 32 * q + lane;                    // point of the tile == TMEM lane
        const int c0 = CPT * h;                         // first of this thread's CPT neurons
        const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16);
        const uint64_t polLast = policy_evict_last();
        const SlabLayout sl = slab_layout(L, net.inpDim);
        float* part = A.part32 + (size_t)blockIdx.x * sl.psz;
        double* part64 = A.part + (size_t)blockIdx.x * sl.psz;
        float* stash = A.stash + (size_t)blockIdx.x * A.stashFloats;
        uint32_t phase = 0;                             // bit i: phase parity this thread waits for next on barrier i
        double lossAcc = 0.0;
        bool first = true, firstFold = !A.accumulate;
        int win = 0;

        auto wait = [&](int i) {
            if (ok) ok = mbar_wait(BAR(i), (phase >> i) & 1u);
            phase ^= 1u << i;
            __syncwarp();
            tc_fence_after();
        };
        // publish this thread's TMEM / shared-memory writes and finished TMEM reads, one arrival per warp
        auto signal = [&](uint32_t b) {
            tmem_wait_st();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b);
        };
        auto vec_add = [&](int kind, float v, bool overwrite) {
            if (lane & ((1 << COLSUM_SHIFT) - 1)) return;
            float* slot = part + sl.vecOff + (kind * 4 + q) * W + c0 + (lane >> COLSUM_SHIFT);
            if (overwrite) __stcg(slot, v); else atomicAdd(slot, v);
        };
        // prefetched per-point values of a tile: xin[buf][c][p], c < nxTable: inputs; 8: dNt; 9, 10: gcoef; 11: source*N.
        // Thread (p, h) fetches columns h, h + 4, h + 8.
        const int XT = 8, XG = 9, XS = 11;
        auto prefetch_inputs = [&](int tile, int buf) {
            const unsigned int gp = (unsigned int)(A.tile0 + tile) * TP + p;
            const bool valid = gp < A.P;
            const size_t row = valid ? table_row(A, gp) : 0;
            float* dst = xin + (size_t)buf * NXIN * TP + p;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int c = h + 4 * u;
                int col = -1;
                if (c < A.nxTable) col = A.colX + c;
                else if (c == XT) col = A.colT;
                else if (c >= XG && c < XG + S - 1) col = A.colG + (c - XG);
                else if (c == XS) col = A.colS;
                if (col >= 0) cp_async4(dst + c * TP, A.cols + (size_t)col * A.pstride + (A.tfIndex ? row : (size_t)gp));
            }
            cp_async_commit();
        };
        prefetch_inputs(blockIdx.x, 0);
        int xb = 0;
        // optional phase timing of one thread (CTA 0, thread 0): slot i accumulates the cycles since the previous mark
        long long* timS = reinterpret_cast<long long*>(smem + OFF2_TIM);
        const bool timOn = K.timing != nullptr && blockIdx.x == 0 && tid == 0;
        if (timOn) { for (int i = 0; i < 15; ++i) timS[i] = 0; timS[15] = clock64(); }
        auto TIM = [&](int i) { if (timOn) { const long long t = clock64(); timS[i] += t - timS[15]; timS[15] = t; } };

        for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
            const unsigned int base = (unsigned int)(A.tile0 + tile) * TP;
            const unsigned int gp = base + p;
            const bool valid = gp < A.P;
            const float* xv = xin + (size_t)xb * NXIN * TP + p;
            cp_async_wait_all();
            worker_bar();                                  // every worker's prefetched values of this tile are in shared memory
            TIM(0);
            auto input = [&](int c) -> float {
                if (c >= A.nxTable) return __ldg(A.extraX + (c - A.nxTable));
                return valid ? xv[c * TP] : 0.f;
            };
            // coefficient of stream s in the integrand I = sum_k du_k gcoef_k - u dNt (TFModel.py:653-657)
            float cf[S];
            cf[0] = (A.timeDependent && valid) ? -xv[XT * TP] : 0.f;
#pragma unroll
            for (int k = 0; k < S - 1; ++k) cf[1 + k] = valid ? xv[(XG + k) * TP] : 0.f;

            // ---- inputs and layer 0 (K = inpDim: FP32 FMA).  Stream 1+k is seeded with the unit vector e_k.
            float d1[CPT];                                // act'(z_l) of the value stream, kept for the tangent streams
            {
                float v[CPT];
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) v[jj] = bs[c0 + jj];
                for (int c = 0; c < net.inpDim; ++c) {
                    const float xc = input(c);
#pragma unroll
                    for (int jj = 0; jj < CPT; ++jj) v[jj] = fmaf(xc, W0s[c * W + c0 + jj], v[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) { v[jj] = act64<ACT>(v[jj]); d1[jj] = act_d1<ACT>(v[jj]); }
                put_operand(tq, c0, v);
                stash_put(stash, 0, p, c0, v, polLast);
                for (int k = 0; k < S - 1; ++k) {
#pragma unroll
                    for (int jj = 0; jj < CPT; ++jj) v[jj] = d1[jj] * W0s[k * W + c0 + jj];
                    put_operand(tq + 128 * (1 + k), c0, v);
                    stash_put(stash, 1 + k, p, c0, v, polLast);
                }
            }
            signal(BAR(B_STEPF));                          // forward event 0
            TIM(1);
            if (tile + (int)gridDim.x < A.ntiles) prefetch_inputs(tile + (int)gridDim.x, xb ^ 1);

            // ---- hidden layers, forward: epilogue of step k while the tensor core runs step k+1
            float tw[CPT];                                 // sum_s a_{L-1,s} coef_s: g(w_out) before the scaling by lambda
#pragma unroll
            for (int jj = 0; jj < CPT; ++jj) tw[jj] = 0.f;
            {
                int l = 1, s = 0;
                for (int k = 0; k < nfs; ++k) {
                    TIM(2);
                    wait(B_ACC + (k & 1));
                    TIM(3);
                    float v[CPT];
                    get_plain(tq + COL2_ACC + 64 * (k & 1) + c0, v);
                    if (s == 0) {
#pragma unroll
                        for (int jj = 0; jj < CPT; ++jj) {
                            v[jj] = act64<ACT>(v[jj] + bs[l * W + c0 + jj]);
                            d1[jj] = act_d1<ACT>(v[jj]);
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < CPT; ++jj) v[jj] *= d1[jj];
                    }
                    put_operand(tq + 128 * s, c0, v);
                    signal(BAR(B_STEPF + ((k + 1) & 1)));      // forward event k+1: accumulator drained, operand written
                    stash_put(stash, l * S + s, p, c0, v, polLast);
                    if (l == L - 1) {
                        // output layer (Dense(1)): partial dot over this thread's neurons
                        float a0 = 0.f, a1 = 0.f;
                        const float cs = s == 0 ? cf[0] : (s == 1 ? cf[1] : cf[S - 1]);
#pragma unroll
                        for (int jj = 0; jj < CPT; jj += 2) {
                            a0 = fmaf(v[jj], wout[c0 + jj], a0); a1 = fmaf(v[jj + 1], wout[c0 + jj + 1], a1);
                            tw[jj] = fmaf(v[jj], cs, tw[jj]); tw[jj + 1] = fmaf(v[jj + 1], cs, tw[jj + 1]);
                        }
                        usP[(h * 3 + s) * TP + p] = a0 + a1;
                    }
                    if (++s == S) { s = 0; ++l; }
                }
            }
            TIM(2);
            worker_bar();

            // ---- u_s, integrand, R_i of the test functions of this tile, loss, adjoint seeds (TFModel.py:653-664)
            if (tid < TP) {
                float u[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    float a = 0.f;
#pragma unroll
                    for (int g = 0; g < NH; ++g) a += usP[(g * 3 + s) * TP + p];
                    u[s] = a;
                }
                u[0] += misc[0];
                float I = 0.f;
                if (valid) {
#pragma unroll
                    for (int k = 0; k < S - 1; ++k) I = fmaf(u[1 + k], cf[1 + k], I);
                    if (A.timeDependent) I = fmaf(u[0], cf[0], I);
                    if (A.isSource) I -= xv[XS * TP];
                    if (A.integW) I *= __ldg(A.integW + (gp % A.integNum));
                }
                Ish[p] = I;
            }
            worker_bar();
            {
                const int nf = TP / (int)A.integNum;
                for (int f = warp; f < nf; f += NWORK / 32) {
                    float r = 0.f;
                    for (int qq = lane; qq < (int)A.integNum; qq += 32) r += Ish[f * A.integNum + qq];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                    if (lane == 0) {
                        const unsigned int i = base / A.integNum + f;
                        Rsh[f] = r;
                        if (i * A.integNum < A.P) {
                            const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                            const float r2 = r * r;
                            A.R[i] = r;
                            A.lossVec[i] = dj * r2;
                            lossAcc += A.detJvec ? (double)dj * (double)r2 : (double)r2;
                        }
                    }
                }
            }
            worker_bar();
            if (tid < TP) {
                float lam = 0.f;
                if (valid) {
                    const unsigned int i = gp / A.integNum, qq = gp - i * A.integNum;
                    const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                    const float wq = A.integW ? __ldg(A.integW + qq) : 1.f;
                    lam = 2.f * __ldg(A.wts + 2) * dj * wq * Rsh[p / A.integNum];
                }
                lamS[p] = lam;
#pragma unroll
                for (int s = 0; s < S; ++s) us[s * TP + p] = lam * cf[s];        // adjoint seeds ubar_s
            }
            worker_bar();

            // ---- output layer gradients: g(b_out) = sum_p ubar_0, g(w_out)[j] = sum_p lambda_p sum_s a_{L-1,s}[p][j] coef_s[p]
            if (h == 0) {
                float sb = us[p];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
                if (lane == 0) { float* slot = part + sl.boutOff + q; if (first) __stcg(slot, sb); else atomicAdd(slot, sb); }
            }
            {
                const float lam = lamS[p];
#pragma unroll
                for (int jj = 0; jj < CPT; ++jj) tw[jj] *= lam;
                const float r = warp_colsum(tw, lane);
                vec_add(L, r, first);
            }

            // ---- adjoint sweep.  Step k: zbar_{l,s} from the parked abar_{l,s} (tangent streams first: the value stream needs
            // their second-order term); the tensor core then computes abar_{l-1,s} = zbar_{l,s} W_l^T into the park of s and
            // gW_l += a_{l-1,s}^T zbar_{l,s}.  The preparation of step k overlaps the MMAs of step k-1.
            {
                // the second-order term of a layer is accumulated over its tangent steps in spare tensor-memory columns (16 registers
                // that would otherwise be spilled: a0, dpre, apre, the step's zbar and the drained gradient tile are live together)
                float a0[CPT], dpre[CPT], apre[CPT];
                stash_get(stash, (L - 1) * S, p, c0, a0, polLast);
                stash_get(stash, (L - 1) * S + 1, p, c0, dpre, polLast);
                stash_get(stash, (L - 2) * S + 1, p, c0, apre, polLast);
                unsigned char* GA = smem + OFF2_G;
                unsigned char* GB = GA + G_BYTES;
                TIM(4);
                // drain of the weight-gradient accumulator of a step of layer ld (fd: first stream of the layer) into the FP32 window slab
                auto drain_gw = [&](int ld, bool fd) {
                    TIM(5);
                    wait(B_GW);
                    TIM(6);
                    // rows 0..63: a_hi (x) [zbar_hi | zbar_lo]; rows 64..127: a_lo (x) zbar_hi (lo x lo dropped)
                    float g[CPT];
                    if (q < 2) drain_sum2(tq + COL_GW + c0, tq + COL_GW + 64 + c0, g); else get_plain(tq + COL_GW + c0, g);
                    float* slot = part + gw_slot(ld, p, c0);
                    if (first && fd) {                                 // first write of this window: overwrite
#pragma unroll
                        for (int u = 0; u < CPT / 4; ++u) __stcg(reinterpret_cast<float4*>(slot) + u * TP, make_float4(g[4 * u], g[4 * u + 1], g[4 * u + 2], g[4 * u + 3]));
                    } else {
#pragma unroll
                        for (int u = 0; u < CPT / 4; ++u) red_add_v4(slot + 4 * u * TP, g[4 * u], g[4 * u + 1], g[4 * u + 2], g[4 * u + 3]);
                    }
                };
                int k = 0;
                for (int l = L - 1; l >= 0; --l) {
                    for (int si = 0; si < S; ++si, ++k) {
                        const int s = (si < S - 1) ? si + 1 : 0;
                        // next step (ln, sn)
                        int ln = l, sn = (si + 1 < S - 1) ? si + 2 : 0;
                        if (si + 1 == S) { ln = l - 1; sn = 1; }
                        if (k == nfs) {                                          // entering layer 0: finish the last MMA step
                            drain_gw(1, false);
                            wait(B_LAYER);
                            TIM(9);
                        }
                        float v[CPT];
                        if (l == L - 1) {
                            const float ub = us[s * TP + p];
#pragma unroll
                            for (int jj = 0; jj < CPT; ++jj) v[jj] = ub * wout[c0 + jj];
                        } else {
                            get_plain(tq + COL_PARK + 64 * s + c0, v);           // abar_{l,s}: its layer GEMM (step k-S) has completed
                        }
                        if (s > 0) {
                            float cross[CPT];
                            if (si == 0) {
#pragma unroll
                                for (int jj = 0; jj < CPT; ++jj) cross[jj] = v[jj] * dpre[jj];
                            } else {
                                get_plain(tq + COL2_CROSS + c0, cross);
#pragma unroll
                                for (int jj = 0; jj < CPT; ++jj) cross[jj] = fmaf(v[jj], dpre[jj], cross[jj]);
                            }
                            put_plain(tq + COL2_CROSS + c0, cross);              // tcgen05.wait::st: in signal() below, before the next step reads it
#pragma unroll
                            for (int jj = 0; jj < CPT; ++jj) v[jj] *= act_d1<ACT>(a0[jj]);
                        } else {
                            float cross[CPT];
                            get_plain(tq + COL2_CROSS + c0, cross);
#pragma unroll
                            for (int jj = 0; jj < CPT; ++jj) v[jj] = fmaf(v[jj], act_d1<ACT>(a0[jj]), act_d2r<ACT>(a0[jj]) * cross[jj]);
                        }
                        // tangent activations of the next step: in flight from here on
                        if (ln >= 0 && sn > 0) stash_get(stash, ln * S + sn, p, c0, dpre, polLast);
                        if (l >= 1) {
                            if (k >= 1) drain_gw(L - 1 - (k - 1) / S, ((k - 1) % S) == 0);     // step k-1; its operands in shared memory are free now
                            put_transposed(GB, p, c0, v);
                            put_transposed(GA, p, c0, apre);
                            if (s == 0) {
#pragma unroll
                                for (int jj = 0; jj < CPT; ++jj) a0[jj] = apre[jj];      // a_{l-1,0}: the value activations of the next layer down
                            }
                            if (ln >= 1) stash_get(stash, (ln - 1) * S + sn, p, c0, apre, polLast);
                            TIM(7);
                            if (k >= 1) wait(B_LAYER);                           // layer GEMM of step k-1 done: the operand region is free
                            TIM(8);
                            put_operand(tq + COL_OP, c0, v);
                            signal(BAR(B_STEPA + (k & 1)));
                            if (s == 0) {
                                const float r = warp_colsum(v, lane);            // g(b_l) = sum_p zbar_{l,0}
                                vec_add(l, r, first);
                            }
                        } else {
                            // layer 0: gb_0 = sum_p zbar_{0,0}; gW_0[c] = sum_p x_c zbar_{0,0} (+ sum_p zbar_{0,1+c} for the spatial inputs)
                            if (s > 0) {
                                const float r = warp_colsum(v, lane);
                                vec_add(L + 1 + (s - 1), r, first);
                            } else {
                                for (int c = 0; c < net.inpDim; ++c) {
                                    const float xc = input(c);
                                    float t[CPT];
#pragma unroll
                                    for (int jj = 0; jj < CPT; ++jj) t[jj] = xc * v[jj];
                                    const float r = warp_colsum(t, lane);
                                    vec_add(L + 1 + c, r, first && c >= S - 1);
                                }
                                const float r = warp_colsum(v, lane);
                                vec_add(0, r, first);
                            }
                        }
                    }
                }
            }
            tmem_wait_ld();
            TIM(10);

            first = false;
            if (++win == FOLD || tile + (int)gridDim.x >= A.ntiles) {
                // fold this thread's slots of the FP32 window into the FP64 slab (single writer per slot)
                auto put = [&](int idx) {
                    const double v = (double)__ldcg(part + idx);
                    if (firstFold) __stcg(part64 + idx, v); else part64[idx] += v;
                };
                for (int l = 1; l < L; ++l)
                    for (int jj = 0; jj < CPT; ++jj) put(gw_slot(l, p, c0 + jj));
                if ((lane & ((1 << COLSUM_SHIFT) - 1)) == 0)
                    for (int kk = 0; kk < sl.nkind; ++kk) put(sl.vecOff + (kk * 4 + q) * W + c0 + (lane >> COLSUM_SHIFT));
                if (h == 0 && lane == 0) put(sl.boutOff + q);
                firstFold = false; first = true; win = 0;
            }
            xb ^= 1;
            TIM(11);
        }
        if (timOn) for (int i = 0; i < 15; ++i) K.timing[i] = timS[i];
        cp_async_wait_all();
        if (lane == 0) {
            double* lp = A.lossPart + blockIdx.x * (NWORK / 32) + warp;
            *lp = A.accumulate ? *lp + lossAcc : lossAcc;
        }
        if (!ok) *K.err = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int S, int ACT> cudaError_t launch_t(const Tc64Args& k, int grid, size_t smem, cudaStream_t st) {
    tc64_var_kernel_v2<S, ACT><<<grid, NT2, smem, st>>>(k);
    return cudaGetLastError();
}
template <int S, int ACT> cudaError_t prepare_t(size_t smem) {
    return cudaFuncSetAttribute(tc64_var_kernel_v2<S, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace v2

// ------------------------------------------------------------------ weight images
// image 2(l-1)   (forward, B rows n = output neuron, K = input):  n < 64: hi(W_l[k][n]),  n >= 64: lo(W_l[k][n-64])
// image 2(l-1)+1 (adjoint, B rows n = input neuron,  K = output): n < 64: hi(W_l[n][k]),  n >= 64: lo(W_l[n-64][k])
__global__ void tc64_prep_kernel(NetDesc net, const float* __restrict__ theta, float* __restrict__ wimg) {
    const int img = blockIdx.y, l = 1 + (img >> 1), dir = img & 1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // (n, k)
    if (e >= 128 * 64) return;
    const int n = e >> 6, k = e & 63, nn = n & 63;
    const int wi = net.width[l - 1], wo = net.width[l];
    const int i = dir ? nn : k, j = dir ? k : nn;
    const float w = (i < wi && j < wo) ? theta[net.woff[l] + i * wo + j] : 0.f;
    const float hi = tf32_rn(w);
    wimg[(size_t)img * WIMG_FLOATS + ((k >> 2) * W_LBO + (n >> 3) * SBO + (n & 7) * 16 + (k & 3) * 4) / 4] = n < 64 ? hi : w - hi;
}

// ------------------------------------------------------------------ cross-CTA reduction (fixed order)
__global__ void tc64_reduce_kernel(NetDesc net, const double* __restrict__ slab, int psz, int nCta, double* __restrict__ flat) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= net.nparam) return;
    const SlabLayout sl = slab_layout(net.L, net.inpDim);
    int l = 0, isBias = 0, i = 0, j = 0;
    for (l = 0; l <= net.L; ++l) {
        const int wi = l == 0 ? net.inpDim : net.width[l - 1];
        const int wo = l == net.L ? 1 : net.width[l];
        if (idx >= net.woff[l] && idx < net.woff[l] + wi * wo) { i = (idx - net.woff[l]) / wo; j = (idx - net.woff[l]) - i * wo; break; }
        if (idx >= net.boff[l] && idx < net.boff[l] + wo) { isBias = 1; j = idx - net.boff[l]; break; }
    }
    int slot[4], nslot = 4;
    if (l == net.L && isBias) { for (int qq = 0; qq < 4; ++qq) slot[qq] = sl.boutOff + qq; }
    else if (l == net.L) { for (int qq = 0; qq < 4; ++qq) slot[qq] = sl.vecOff + (net.L * 4 + qq) * W + i; }
    else if (isBias) { for (int qq = 0; qq < 4; ++qq) slot[qq] = sl.vecOff + (l * 4 + qq) * W + j; }
    else if (l == 0) { for (int qq = 0; qq < 4; ++qq) slot[qq] = sl.vecOff + ((net.L + 1 + i) * 4 + qq) * W + j; }
    else { nslot = 2; slot[0] = gw_slot(l, i, j); slot[1] = gw_slot(l, 64 + i, j); }
    double s = 0.0;
    for (int c = lane; c < nCta; c += 32)
        for (int k = 0; k < nslot; ++k) s += __ldcg(slab + (size_t)c * psz + slot[k]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) flat[idx] = s;
}

// weight-gradient operands of the default schedule (VARNET_B200_TC64_GW), all bit-identical in their results:
//   3 (default) = K-major operands written by transposing 4-byte stores, with [a_hi ; a_lo] of the NEXT adjoint step (its stash rows
//       are in registers by then) written as soon as the weight-gradient MMAs of the running step have completed (their own
//       mbarrier on every step), i.e. under the layer GEMM instead of in front of the next issue; the bias column sums moved behind
//       the issue as well.  Same-box A/B on 1/4 of cfg 4 (gpurun_out/r2bo_ab.log, r2bp_ab.log): 38.43 -> 37.56 ms (+2.3 %);
//   0 = the same operands, all written in front of the issue (round 2's first schedule);
//   1 = MN-major operands (SWIZZLE_128B_BASE32B), plain 16-byte stores; 2 = MN-major, quad-exchanged lanes (bank-conflict-free
//       16-byte stores).  Measured (gpurun_out/r2be_ab.log): 38.4 (mode 0) / 42.2 / 40.4 ms.  The MN-major MMAs run at the full rate
//       (scripts/micro/umma_rate.cu: 64.0 against 63.5 cycles per 128x128x8), but a thread owns a (point, 16 neurons) strip, so its
//       16-byte stores of one instruction hit only 4 of the 8 bank groups (mode 1: 2-way conflicts, 2x the store wavefronts)
//       unless half of the lanes exchange register quads first (mode 2: +7 % instructions); the scalar transposing stores have
//       neither problem and the same wavefront count.
int gw_mode() {
    const char* e = getenv("VARNET_B200_TC64_GW");         // read per launch: a test switches it inside one process
    const int m = e ? atoi(e) : 3;
    return (m < 0 || m > 3) ? 3 : m;
}
template <int S, int ACT> cudaError_t launch_t(const Tc64Args& k, int grid, size_t smem, cudaStream_t st) {
    switch (gw_mode()) {
        case 0: tc64_var_kernel<S, ACT, 0><<<grid, NT, smem, st>>>(k); break;
        case 1: tc64_var_kernel<S, ACT, 1><<<grid, NT, smem, st>>>(k); break;
        case 2: tc64_var_kernel<S, ACT, 2><<<grid, NT, smem, st>>>(k); break;
        default: tc64_var_kernel<S, ACT, 3><<<grid, NT, smem, st>>>(k); break;
    }
    return cudaGetLastError();
}
template <int S, int ACT> cudaError_t prepare_t(size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(tc64_var_kernel<S, ACT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc64_var_kernel<S, ACT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc64_var_kernel<S, ACT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc64_var_kernel<S, ACT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e;
}

}  // namespace

static long long* g_timBuf = nullptr;       // phase cycle counters of the last v2 launch (CTA 0, thread 0), VARNET_B200_TC64_TIMING=1
int vn_tc64_read_timing(long long out[16]) {
    if (!g_timBuf) return 0;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, g_timBuf, 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 1 : 0;
}
// Two schedules of the same tile algorithm.  Default: the round-1 kernel (512 threads, MMAs issued by a worker thread between
// __syncthreads).  VARNET_B200_TC64=v2 selects the warp-specialised, software-pipelined kernel of round 2: parity-green and
// bitwise reproducible too, and measured within 1 % of the default on the headline workload (167.3 vs 165.9 ms per 6.4e7-point
// step, same box): with the MMA waits overlapped, a tile is bound by the workers' own instruction stream
// (profiles/r2_tc64_v2.md has the phase table).
static bool use_v1() {
    static const int v = [] { const char* e = getenv("VARNET_B200_TC64"); return (e && !strcmp(e, "v2")) ? 0 : 1; }();
    return v != 0;
}
bool vn_tc64_supported(const NetDesc& net, int S) {
    if (S < 2 || S > 3 || net.L < 2 || net.L > 6 || net.inpDim > VN_KIN) return false;
    int wmax = 0;
    for (int l = 0; l < net.L; ++l) wmax = net.width[l] > wmax ? net.width[l] : wmax;
    return wmax > 32 && wmax <= W;
}
void vn_tc64_geometry(const NetDesc& net, int S, Tc64Geom* g) {
    const SlabLayout sl = slab_layout(net.L, net.inpDim);
    g->psz = sl.psz;
    g->smemBytes = use_v1() ? SMEM_BYTES : v2::SMEM2_BYTES;
    g->stashFloats = (long long)net.L * S * TP * W;
    g->nImages = 2 * (net.L - 1);
    g->lossSlots = NT / 32;
}
cudaError_t vn_tc64_prepare(int S, int act, size_t smem) {
    if (use_v1()) {
        if (S == 2) return act == VN_SIGMOID ? prepare_t<2, VN_SIGMOID>(smem) : prepare_t<2, VN_TANH>(smem);
        return act == VN_SIGMOID ? prepare_t<3, VN_SIGMOID>(smem) : prepare_t<3, VN_TANH>(smem);
    }
    if (S == 2) return act == VN_SIGMOID ? v2::prepare_t<2, VN_SIGMOID>(smem) : v2::prepare_t<2, VN_TANH>(smem);
    return act == VN_SIGMOID ? v2::prepare_t<3, VN_SIGMOID>(smem) : v2::prepare_t<3, VN_TANH>(smem);
}
cudaError_t vn_tc64_stage_weights(const NetDesc& net, const float* theta, float* wimg, cudaStream_t st) {
    tc64_prep_kernel<<<dim3(32, 2 * (net.L - 1)), 256, 0, st>>>(net, theta, wimg);
    return cudaGetLastError();
}
bool vn_tc64_forward_only_available() { return use_v1(); }
cudaError_t vn_tc64_launch(int S, int act, const TileArgs& a, const float* wimg, int* err, int grid, size_t smem, cudaStream_t st, int fwdOnly) {
    Tc64Args k;
    k.t = a; k.wimg = wimg; k.err = err; k.timing = nullptr; k.fwdOnly = fwdOnly;
    static const int fold = [] { const char* e = getenv("VARNET_B200_TC64_FOLD"); const int v = e ? atoi(e) : FOLD; return v > 0 ? v : FOLD; }();
    k.fold = fold;
    // bit 0: integrand coefficients requested at tile start; bit 1: MLP inputs staged in shared memory one tile ahead
    { const char* ev = getenv("VARNET_B200_TC64_COEF"); const int m = ev ? atoi(ev) : 7; k.coefPre = (m < 0 || m > 7) ? 7 : m; }       // bit 0: coefficients, bit 1: inputs, bit 2: merged layer-0 column sums
    if (fwdOnly && !use_v1()) return cudaErrorNotSupported;
    static const bool timWanted = getenv("VARNET_B200_TC64_TIMING") != nullptr;
    if (timWanted && !use_v1()) {
        if (!g_timBuf) { cudaMalloc(&g_timBuf, 16 * sizeof(long long)); cudaMemset(g_timBuf, 0, 16 * sizeof(long long)); }
        k.timing = g_timBuf;
    }
    if (use_v1()) {
        if (S == 2) return act == VN_SIGMOID ? launch_t<2, VN_SIGMOID>(k, grid, smem, st) : launch_t<2, VN_TANH>(k, grid, smem, st);
        return act == VN_SIGMOID ? launch_t<3, VN_SIGMOID>(k, grid, smem, st) : launch_t<3, VN_TANH>(k, grid, smem, st);
    }
    if (S == 2) return act == VN_SIGMOID ? v2::launch_t<2, VN_SIGMOID>(k, grid, smem, st) : v2::launch_t<2, VN_TANH>(k, grid, smem, st);
    return act == VN_SIGMOID ? v2::launch_t<3, VN_SIGMOID>(k, grid, smem, st) : v2::launch_t<3, VN_TANH>(k, grid, smem, st);
}
cudaError_t vn_tc64_reduce(const NetDesc& net, const double* slab64, int psz, int nCta, double* flat, cudaStream_t st) {
    tc64_reduce_kernel<<<(net.nparam + 3) / 4, 128, 0, st>>>(net, slab64, psz, nCta, flat);
    return cudaGetLastError();
}
