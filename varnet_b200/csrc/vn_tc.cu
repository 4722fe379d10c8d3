// vn_tc.cu — tensor-core kernel class (see vn_tc.h): tcgen05.mma kind::tf32 with the 3xTF32 split and accumulators
// in TMEM.  Layer GEMMs: activation operand written to tensor memory by its loaders, weight tiles in shared memory;
// weight-gradient GEMM: both operands in shared memory.  Shared-memory operands use the no-swizzle K-major
// canonical layout (8 rows x 16 bytes core matrices; validated by scripts/micro/tc_probe.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <algorithm>
#include "vn_tc.h"

namespace {

constexpr int TM = 128;         // rows (points, or input neurons in tc_gw) per CTA tile == TMEM lanes
constexpr int TN = 128;         // columns (neurons) per CTA tile
constexpr int NTHR = 256;        // loader / epilogue threads (8 warps)
constexpr int NTHR_ALL = NTHR + 32;   // + one MMA-issuing warp (tc_gw_kernel and the all-shared-memory GEMM variant)
constexpr int KC = 32;           // K floats per pipeline stage (one 128-byte row segment per operand row)
constexpr uint32_t TILE_SBO = 128;                  // bytes between 8-row groups
// bytes between 16-byte K units: 128 rows x 16 B, plus one 16-byte pad so that the eight K units of one row
// fall into eight different bank groups (the loader writes a row's 128-byte segment with one quarter-warp)
constexpr uint32_t TILE_LBO = (TM / 8) * 128 + 16;
constexpr int TILE_BYTES = (KC / 4) * TILE_LBO;
constexpr int NST = 3;                               // pipeline stages
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A hi, A lo, B hi, B lo
constexpr int BAR_OFF = NST * STAGE_BYTES;         // mbarriers: full[NST] (256 loader arrivals), empty[NST] and done (tcgen05.commit)
constexpr int SMEM_BYTES = BAR_OFF + 128;
// A-from-TMEM variant of the layer GEMM (tc_gemm_kernel<..., TS = true>): the activation operand is written by its
// loader threads straight into tensor memory, so the MMAs read only the weight tiles from shared memory
// (48 KB per K chunk instead of 96 KB) and the loaders store only those (32 KB instead of 64 KB): the main loop is
// no longer shared-memory bound.  TMEM: two main sets + the small-term set (384 columns) + two A stages of
// (hi 32 | lo 32) columns.
constexpr int TS_NST = 2;
constexpr int TS_STAGE_BYTES = 2 * TILE_BYTES;       // B hi, B lo
constexpr int TS_BAR_OFF = TS_NST * TS_STAGE_BYTES;
constexpr int TS_SMEM_BYTES = TS_BAR_OFF + 128;
constexpr uint32_t TS_ACOL = 3 * 128;
constexpr uint32_t TMEM_COLS = 512;                  // 128-column accumulator sets (main sets + the small terms) [+ the A stages]

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                  // start address        [0,14)
    d |= (uint64_t)((TILE_LBO >> 4) & 0x3FFF) << 16;         // leading byte offset  [16,30)
    d |= (uint64_t)((TILE_SBO >> 4) & 0x3FFF) << 32;         // stride byte offset   [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version 1 (sm_100)
    return d;                                                // no swizzle, base offset 0
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t dTmem, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dTmem), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
// MN-major operands of the weight-gradient GEMM (rows of the operand = points = K): LayoutType::SWIZZLE_128B_BASE32B, the only
// canonical layout in which kind::tf32 takes MN-major shared-memory operands (scripts/micro/tc_probe3.cu).  One tile =
// 128 operand rows (neurons) x KC points: four MN groups of 32 neurons, each KC rows of 128 B (one point, 32 neurons), the
// 32-byte chunk index of a row XORed with point & 3; k atoms of 4 points (512 B).
constexpr uint32_t MN_ROW = 128, MN_ATOM = 4 * MN_ROW, MN_GROUP = KC * MN_ROW, MN_TILE = 4 * MN_GROUP, MN_KSTEP = 8 * MN_ROW;
constexpr int MN_STAGE_BYTES = 4 * MN_TILE;            // A hi, A lo, B hi, B lo
constexpr int MN_BAR_OFF = NST * MN_STAGE_BYTES;
constexpr int MN_SMEM_BYTES = MN_BAR_OFF + 128;
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((MN_GROUP >> 4) & 0x3FFF) << 16;        // leading byte offset: between MN groups
    d |= (uint64_t)((MN_ATOM >> 4) & 0x3FFF) << 32;         // stride byte offset: between k atoms
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                 // SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ void mma_tf32_mn(uint32_t dTmem, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dTmem), "l"(da), "l"(db), "r"(IDESC | (1u << 15) | (1u << 16)), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {      // 16-byte vector reduction (sm_90+)
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// bounded wait (a lost commit must not hang the GPU): false on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 18) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
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
// A operand from tensor memory (lane = row, one 32-bit K element per column), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t dTmem, uint32_t aTmem, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(dTmem), "r"(aTmem), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t NCOL>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOL) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOL>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NCOL) : "memory");
}

// ------------------------------------------------------------------ operand tiles
// One tile = 128 rows x KC floats, loaded by 256 threads.  Every activation array of a chunk is stored once, in the
// quad-major layout [index/4][point][4] (16-byte units of four consecutive neurons of one point, points contiguous):
//   LAY_QM — layer GEMMs (rows = points, K = neurons): the global order IS the order of the shared-memory canonical
//            layout, so a warp reads 512 contiguous bytes and writes four whole core matrices, and the GEMM epilogues
//            (thread = point) store and reload it fully coalesced;
//   LAY_QT — weight-gradient GEMM (rows = neurons, K = points): lane = point, so a warp request is again 512
//            contiguous bytes (one neuron quad x 32 points); the transposition happens in the shared-memory store:
//            each of the four neurons of a unit is written as a 4-byte scalar into its (neuron, 4 points) K unit —
//            bank = lane thanks to the 16-byte pad on TILE_LBO, so these stores are conflict free too.
// No second (neuron-major) copy of the activations exists.
//   LAY_MN — weight-gradient GEMM with MN-major operands: the 16-byte quad-major units go to shared memory as they are (two
//            16-byte stores per unit: hi and lo image) — no transposition.  A quarter warp covers 4 points x the 2 quads of
//            one 32-byte chunk, i.e. 8 different 16-byte bank groups (the swizzle separates the 4 points), and its global
//            request is two runs of 64 contiguous bytes.
enum { LAY_QM = 0, LAY_QT = 1, LAY_MN = 2 };
struct Opnd { const float* p; size_t ld; };     // p = array + (first row / 4 for QT, first row for QM) term, ld = points of the array
struct TileRegs { float4 v[TM * (KC / 4) / NTHR]; };

template <int LAY>
__device__ __forceinline__ void tile_load(const Opnd& o, int k0, TileRegs& r, int tid) {
    constexpr int U = TM * (KC / 4) / NTHR;
    if (LAY == LAY_QM) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int u = i * NTHR + tid, row = u % TM, k4 = u / TM;
            r.v[i] = __ldg(reinterpret_cast<const float4*>(o.p + ((size_t)(k0 / 4 + k4) * o.ld + row) * 4));
        }
    } else if (LAY == LAY_MN) {
        static_assert(U == 4 && KC == 32, "warp = two neuron quads x 16 points per request");
        const int warp = tid >> 5, lane = tid & 31;
        const int pt = 4 * (lane >> 3) + (lane & 3), qb = (lane >> 2) & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = 4 * warp + i, quad = 2 * (m >> 1) + qb, point = 16 * (m & 1) + pt;
            r.v[i] = __ldg(reinterpret_cast<const float4*>(o.p + ((size_t)quad * o.ld + k0 + point) * 4));
        }
} else {
        static_assert(U == 4 && KC == 32, "warp = four neuron quads x 32 points");
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            r.v[i] = __ldg(reinterpret_cast<const float4*>(o.p + ((size_t)(4 * warp + i) * o.ld + k0 + lane) * 4));
    }
}
// round to the 10-bit TF32 mantissa (the tensor core itself truncates: a truncated split leaves a one-sided
// 2^-21 bias per product that grows with K; rounding the high half makes the residual sign-symmetric)
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void store_split(float4 v, unsigned char* hiTile, unsigned char* loTile, uint32_t off) {
    const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    *reinterpret_cast<float4*>(hiTile + off) = h;
    *reinterpret_cast<float4*>(loTile + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);   // truncated by the tensor core: unbiased here
}
template <int LAY>
__device__ __forceinline__ void tile_store_split(const TileRegs& r, unsigned char* hiTile, unsigned char* loTile, int tid) {
    constexpr int U = TM * (KC / 4) / NTHR;
    if (LAY == LAY_QM) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int u = i * NTHR + tid, row = u % TM, k4 = u / TM;
            store_split(r.v[i], hiTile, loTile, k4 * TILE_LBO + row * 16);
        }
    } else if (LAY == LAY_MN) {
        const int warp = tid >> 5, lane = tid & 31;
        const int pt = 4 * (lane >> 3) + (lane & 3), qb = (lane >> 2) & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = 4 * warp + i, quad = 2 * (m >> 1) + qb, point = 16 * (m & 1) + pt;
            const uint32_t off = (quad >> 3) * MN_GROUP + point * MN_ROW + (((uint32_t)((quad & 7) >> 1) ^ (uint32_t)(point & 3)) << 5) + (quad & 1) * 16;
            store_split(r.v[i], hiTile, loTile, off);
        }
} else {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = r.v[i];                                         // neurons 4q..4q+3 (q = 4 warp + i) of point `lane`
            const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
            const uint32_t off = (lane >> 2) * TILE_LBO + (16 * warp + 4 * i) * 16 + (lane & 3) * 4;
            *reinterpret_cast<float*>(hiTile + off) = h.x;       *reinterpret_cast<float*>(loTile + off) = v.x - h.x;
            *reinterpret_cast<float*>(hiTile + off + 16) = h.y;  *reinterpret_cast<float*>(loTile + off + 16) = v.y - h.y;
            *reinterpret_cast<float*>(hiTile + off + 32) = h.z;  *reinterpret_cast<float*>(loTile + off + 32) = v.z - h.z;
            *reinterpret_cast<float*>(hiTile + off + 48) = h.w;  *reinterpret_cast<float*>(loTile + off + 48) = v.w - h.w;
        }
    }
}

// MMAs of one K chunk.  The tensor core accumulates with truncation, a one-sided error of ~0.25 ulp of the
// accumulator per MMA (measured: error grows linearly with the chain length), so the chains are kept short:
// the two small products lo*hi, hi*lo go to their own column set (magnitude 2^-11: their truncation is
// harmless) and hi*hi of K block kb goes to the main set chosen by `mainSet(kb, fresh)`; the sets are summed
// in FP32 (round to nearest) by the epilogue.
template <class F>
__device__ __forceinline__ void issue_chunk_mn(uint32_t tmem, uint32_t stageAddr, bool firstChunk, F&& mainSet) {
    const uint32_t aHi = stageAddr, aLo = stageAddr + MN_TILE, bHi = stageAddr + 2 * MN_TILE, bLo = stageAddr + 3 * MN_TILE;
#pragma unroll
    for (int kb = 0; kb < KC / 8; ++kb) {
        const uint64_t dAh = make_desc_mn(aHi + kb * MN_KSTEP), dAl = make_desc_mn(aLo + kb * MN_KSTEP);
        const uint64_t dBh = make_desc_mn(bHi + kb * MN_KSTEP), dBl = make_desc_mn(bLo + kb * MN_KSTEP);
        mma_tf32_mn(tmem + 3 * TN, dAl, dBh, (firstChunk && kb == 0) ? 0u : 1u);
        mma_tf32_mn(tmem + 3 * TN, dAh, dBl, 1u);
        bool fresh = false;
        const int set = mainSet(kb, fresh);
        mma_tf32_mn(tmem + set * TN, dAh, dBh, fresh ? 0u : 1u);
    }
}
template <class F>
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint32_t stageAddr, bool firstChunk, F&& mainSet) {
    const uint32_t aHi = stageAddr, aLo = stageAddr + TILE_BYTES, bHi = stageAddr + 2 * TILE_BYTES, bLo = stageAddr + 3 * TILE_BYTES;
#pragma unroll
    for (int kb = 0; kb < KC / 8; ++kb) {
        const uint64_t dAh = make_desc(aHi + kb * 2 * TILE_LBO), dAl = make_desc(aLo + kb * 2 * TILE_LBO);
        const uint64_t dBh = make_desc(bHi + kb * 2 * TILE_LBO), dBl = make_desc(bLo + kb * 2 * TILE_LBO);
        mma_tf32(tmem + 3 * TN, dAl, dBh, (firstChunk && kb == 0) ? 0u : 1u);
        mma_tf32(tmem + 3 * TN, dAh, dBl, 1u);
        bool fresh = false;
        const int set = mainSet(kb, fresh);
        mma_tf32(tmem + set * TN, dAh, dBh, fresh ? 0u : 1u);
    }
}

// barriers: full[b] = all 256 loader threads stored (and proxy-fenced) their part of stage b; empty[b] = the MMAs
// that read stage b completed (tcgen05.commit); done = every MMA of the tile completed
struct Bars {
    uint32_t full, empty, done;         // shared-space addresses; full + 8*b, empty + 8*b
};
__device__ __forceinline__ uint32_t pipe_setup(unsigned char* smem, int tid, int warp, Bars& bars, int barOff = BAR_OFF) {
    uint64_t* bp = reinterpret_cast<uint64_t*>(smem + barOff);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + barOff + 8 * (2 * NST + 1));
    bars.full = smem_u32(bp); bars.empty = smem_u32(bp + NST); bars.done = smem_u32(bp + 2 * NST);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) { mbar_init(bars.full + 8 * i, NTHR); mbar_init(bars.empty + 8 * i, 1); }
        mbar_init(bars.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<TMEM_COLS>(tslot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *tslot;
}

// Loader side of one K chunk (256 threads): wait until the MMAs that read this stage three chunks ago are done,
// write the split tiles, put the chunk after next in flight, publish the stage to the MMA warp.
struct ChunkSrc { Opnd a, b; int k0; };
template <int LAY, class F>
__device__ __forceinline__ void loader_step(unsigned char* smem, const Bars& bars, int it, int nIt, TileRegs& ra, TileRegs& rb,
                                            F&& src, int tid, bool& ok) {
    const int b = it % NST;
    constexpr int stageBytes = LAY == LAY_MN ? MN_STAGE_BYTES : STAGE_BYTES, tileBytes = LAY == LAY_MN ? (int)MN_TILE : TILE_BYTES;
    unsigned char* stage = smem + b * stageBytes;
    if (it >= NST && ok) ok = mbar_wait(bars.empty + 8 * b, (uint32_t)((it / NST) - 1) & 1u);
    tile_store_split<LAY>(ra, stage, stage + tileBytes, tid);
    tile_store_split<LAY>(rb, stage + 2 * tileBytes, stage + 3 * tileBytes, tid);
    if (it + 2 < nIt) {
        const ChunkSrc c = src(it + 2);
        tile_load<LAY>(c.a, c.k0, ra, tid);
        tile_load<LAY>(c.b, c.k0, rb, tid);
    }
    fence_async_smem();                 // generic-proxy stores -> visible to the tensor core (async proxy)
    mbar_arrive(bars.full + 8 * b);
}
// MMA warp (one elected lane): consume the stages in order
template <bool MN = false, class F>
__device__ __forceinline__ bool mma_warp_loop(unsigned char* smem, const Bars& bars, uint32_t tmem, int nIt, F&& mainSetOf) {
    bool ok = true;
    for (int it = 0; it < nIt; ++it) {
        const int b = it % NST;
        if (ok) ok = mbar_wait(bars.full + 8 * b, (uint32_t)(it / NST) & 1u);
        tc_fence_after();
        if (MN) issue_chunk_mn(tmem, smem_u32(smem + b * MN_STAGE_BYTES), it == 0, [&](int kb, bool& fresh) { return mainSetOf(it, kb, fresh); });
        else issue_chunk(tmem, smem_u32(smem + b * STAGE_BYTES), it == 0, [&](int kb, bool& fresh) { return mainSetOf(it, kb, fresh); });
        mma_commit(bars.empty + 8 * b);
    }
    mma_commit(bars.done);
    return ok;
}

// ------------------------------------------------------------------ layer GEMM (forward / adjoint), one stream per launch
enum { EPI_FWD_VALUE = 0, EPI_FWD_TANGENT = 1, EPI_ADJ_TANGENT = 2, EPI_ADJ_VALUE = 3 };

struct GemmArgs {
    const float* A;                         // this stream's operand, quad-major [K/4][rows][4]
    const float* B; int rowsB;              // weights, quad-major [K/4][rowsB][4], zero padded
    unsigned int rows;                      // rows (points) of every quad-major activation array of the chunk
    int K;                                  // multiple of KC
    int nTilesN;
    const float* bias; int widthOut;        // FWD_VALUE
    const float* val;                       // FWD_TANGENT: a of this layer; ADJ_*: a of the layer below   (quad-major)
    const float* tan;                       // ADJ_TANGENT: tangent activation of the layer below, this stream
    float* cross; int crossMode;            // ADJ_TANGENT: 0 = write, 1 = accumulate; ADJ_VALUE: 1 = read, 0 = no tangent streams
    float* outQm;                           // quad-major [N/4][rows][4]
    int* err;
};

// 16 consecutive columns n..n+15 of row `prow` of a quad-major array: four 16-byte units, coalesced across the warp
__device__ __forceinline__ void ld16(const float* __restrict__ base, unsigned int rows, size_t prow, int n, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + ((size_t)(n / 4 + q) * rows + prow) * 4));
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
__device__ __forceinline__ float4* qm_ptr(float* base, unsigned int rows, size_t prow, int n) {
    return reinterpret_cast<float4*>(base + ((size_t)(n / 4) * rows + prow) * 4);
}

// TS kernels are built for two co-resident CTAs per SM (66 KB of shared memory, <= 112 registers): each CTA needs all
// 512 TMEM columns, so `tcgen05.alloc` of the second CTA blocks until the first one has drained its accumulators into
// registers and released them — its prologue (barrier set-up, first global loads) overlaps the other CTA's main loop,
// and its epilogue arithmetic / stores overlap the other CTA's next main loop.
template <int EPI, int ACT, bool TS>
__global__ void __launch_bounds__(TS ? NTHR : NTHR_ALL, TS ? 2 : 1) tc_gemm_kernel(const GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (*reinterpret_cast<volatile int*>(a.err)) return;                      // an earlier launch lost a barrier: do not spin again
    const int mt = blockIdx.x / a.nTilesN, nt = blockIdx.x - mt * a.nTilesN;
    const size_t m0 = (size_t)mt * TM;
    const int n0 = nt * TN;
    const int nIt = a.K / KC;
    const bool loader = TS || warp < NTHR / 32;                               // TS: 8 warps, thread 0 also issues the MMAs (two CTAs per SM need <= 128 registers x 256 threads)
    const Opnd oa{a.A + m0 * 4, a.rows}, ob{a.B + (size_t)n0 * 4, (size_t)a.rowsB};
    // TS: thread = its own row (TMEM lane) x 16 K elements (warps 0-3: K 0..15 of the chunk, warps 4-7: K 16..31)
    const int arow = (warp & 3) * 32 + lane, ahalf = (warp >> 2) & 1;
    auto loadA = [&](int it, TileRegs& r) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            r.v[j] = __ldg(reinterpret_cast<const float4*>(oa.p + ((size_t)(it * (KC / 4) + ahalf * 4 + j) * oa.ld + arow) * 4));
    };
    TileRegs ra[2], rb[2];                                                    // two K chunks in flight
    if (loader) {                                                             // first loads before TMEM is even allocated
        if (TS) loadA(0, ra[0]); else tile_load<LAY_QM>(oa, 0, ra[0], tid);
        tile_load<LAY_QM>(ob, 0, rb[0], tid);
        if (nIt > 1) {
            if (TS) loadA(1, ra[1]); else tile_load<LAY_QM>(oa, KC, ra[1], tid);
            tile_load<LAY_QM>(ob, KC, rb[1], tid);
        }
    }
    Bars bars;
    const uint32_t tmem = pipe_setup(smem, tid, warp, bars, TS ? TS_BAR_OFF : BAR_OFF);

    const int row = (warp & 3) * 32 + lane, half = (warp >> 2) & 1;
    const size_t prow = m0 + row;
    const int nb = n0 + half * 64;                                            // first of this thread's 64 columns
    float zs[64];                                                             // this thread's accumulator sums
    if (!loader) {
        // ---- MMA warp: the main chain rotates over three (TS: two) column sets
        if (lane == 0) {
            bool ok = true;
            if constexpr (TS) {
                (void)ok;
            } else {
                ok = mma_warp_loop(smem, bars, tmem, nIt, [&](int it, int kb, bool& fresh) {
                    const int kbg = it * (KC / 8) + kb;
                    fresh = kbg < 3;
                    return kbg % 3;
                });
            }
            if (!ok) *a.err = 1;
        }
    } else {
        // ---- loaders
        bool ok = true;
        if constexpr (TS) {
            const uint32_t tA = tmem + ((uint32_t)((warp & 3) * 32) << 16) + TS_ACOL + ahalf * 16;
            auto step = [&](int it, TileRegs& rA, TileRegs& rB) {
                const int b = it % TS_NST;
                unsigned char* stage = smem + b * TS_STAGE_BYTES;
                if (it >= TS_NST && ok) ok = mbar_wait(bars.empty + 8 * b, (uint32_t)((it / TS_NST) - 1) & 1u);
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = rA.v[j];
                    hi[4 * j] = tf32_rn(v.x); hi[4 * j + 1] = tf32_rn(v.y); hi[4 * j + 2] = tf32_rn(v.z); hi[4 * j + 3] = tf32_rn(v.w);
                    lo[4 * j] = v.x - hi[4 * j]; lo[4 * j + 1] = v.y - hi[4 * j + 1]; lo[4 * j + 2] = v.z - hi[4 * j + 2]; lo[4 * j + 3] = v.w - hi[4 * j + 3];
                }
                tc_fence_after();
                tmem_st16(tA + b * 64, hi);
                tmem_st16(tA + b * 64 + 32, lo);
                tile_store_split<LAY_QM>(rB, stage, stage + TILE_BYTES, tid);
                if (it + 2 < nIt) { loadA(it + 2, rA); tile_load<LAY_QM>(ob, (it + 2) * KC, rB, tid); }
                tmem_wait_st();
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(bars.full + 8 * b);
                if (tid == 0) {
                    // MMAs of this chunk: small products into set 2, hi*hi alternating between sets 0 and 1
                    if (ok) ok = mbar_wait(bars.full + 8 * b, (uint32_t)(it / TS_NST) & 1u);
                    tc_fence_after();
                    const uint32_t bHi = smem_u32(stage), bLo = bHi + TILE_BYTES;
                    const uint32_t aHi = tmem + TS_ACOL + b * 64, aLo = aHi + 32;
#pragma unroll
                    for (int kb = 0; kb < KC / 8; ++kb) {
                        const uint64_t dBh = make_desc(bHi + kb * 2 * TILE_LBO), dBl = make_desc(bLo + kb * 2 * TILE_LBO);
                        mma_tf32_ts(tmem + 2 * TN, aLo + kb * 8, dBh, (it == 0 && kb == 0) ? 0u : 1u);
                        mma_tf32_ts(tmem + 2 * TN, aHi + kb * 8, dBl, 1u);
                        const int kbg = it * (KC / 8) + kb;
                        mma_tf32_ts(tmem + (kbg & 1) * TN, aHi + kb * 8, dBh, kbg < 2 ? 0u : 1u);
                    }
                    mma_commit(bars.empty + 8 * b);
                }
            };
#pragma unroll 1
            for (int it0 = 0; it0 < nIt; it0 += 2) {
                step(it0, ra[0], rb[0]);
                if (it0 + 1 < nIt) step(it0 + 1, ra[1], rb[1]);
            }
            if (tid == 0) mma_commit(bars.done);
        } else {
            auto src = [&](int it) { return ChunkSrc{oa, ob, it * KC}; };
#pragma unroll 1
            for (int it0 = 0; it0 < nIt; it0 += 2) {
                loader_step<LAY_QM>(smem, bars, it0, nIt, ra[0], rb[0], src, tid, ok);
                if (it0 + 1 < nIt) loader_step<LAY_QM>(smem, bars, it0 + 1, nIt, ra[1], rb[1], src, tid, ok);
            }
        }
        // ---- drain: thread = one point (TMEM lane), warps 0-3 / 4-7 take the two column halves; the accumulator sets are
        // summed in FP32 (round to nearest) into registers so that tensor memory can be released before the epilogue
        if (ok) ok = mbar_wait(bars.done, 0u);
        if (!ok) *a.err = 1;
        tc_fence_after();
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * 64;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            float z[16], t1[16], t2[16], t3[16];
            tmem_ld16(tlane + cb * 16, z);
            tmem_ld16(tlane + TN + cb * 16, t1);
            tmem_ld16(tlane + 2 * TN + cb * 16, t2);
            if (!TS) tmem_ld16(tlane + 3 * TN + cb * 16, t3);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) zs[cb * 16 + c] = TS ? (z[c] + t1[c]) + t2[c] : (z[c] + t1[c]) + (t2[c] + t3[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
    if (!loader) return;

    // ---- epilogue from registers (operands are quad-major: every access of a warp is a contiguous 512-byte run)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        const int n = nb + cb * 16;
        float z[16], av[16], xv[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) z[c] = zs[cb * 16 + c];
        if (EPI != EPI_FWD_VALUE) ld16(a.val, a.rows, prow, n, av);
        if (EPI == EPI_ADJ_TANGENT) ld16(a.tan, a.rows, prow, n, xv);
        if (EPI == EPI_ADJ_VALUE && a.crossMode) ld16(a.cross, a.rows, prow, n, xv);
        if (EPI == EPI_FWD_VALUE) {
            // a = act(z + b)   (App. A.2)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bv = (n + c < a.widthOut) ? __ldg(a.bias + n + c) : 0.f;
                z[c] = act_f<ACT>(z[c] + bv);
            }
        } else if (EPI == EPI_FWD_TANGENT) {
            // tangent stream: act'(z) * zdot, act' from the value a of this layer
#pragma unroll
            for (int c = 0; c < 16; ++c) z[c] *= act_d1<ACT>(av[c]);
        } else if (EPI == EPI_ADJ_TANGENT) {
            // dabar_k -> dzbar_k = dabar_k act'; its share of the second-order term: cross += dabar_k * adot_k   (App. A.3)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float* cp = reinterpret_cast<float*>(qm_ptr(a.cross, a.rows, prow, n + 4 * q));
                const float c0 = z[4 * q] * xv[4 * q], c1 = z[4 * q + 1] * xv[4 * q + 1];
                const float c2 = z[4 * q + 2] * xv[4 * q + 2], c3 = z[4 * q + 3] * xv[4 * q + 3];
                if (a.crossMode) red_add_v4(cp, c0, c1, c2, c3);
                else *reinterpret_cast<float4*>(cp) = make_float4(c0, c1, c2, c3);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) z[c] *= act_d1<ACT>(av[c]);
        } else {
            // abar -> zbar = abar act' + act''/act' * cross
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float d1 = act_d1<ACT>(av[c]);
                z[c] = a.crossMode ? fmaf(z[c], d1, act_d2r<ACT>(av[c]) * xv[c]) : z[c] * d1;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) *qm_ptr(a.outQm, a.rows, prow, n + 4 * q) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
    }
}

// ------------------------------------------------------------------ weight-gradient GEMM (split-K over points)
struct GwArgs {
    const float* A; size_t aStream;     // quad-major activations of layer l-1: [s][i/4][ld][4]
    const float* B; size_t bStream;     // quad-major zbar of layer l:          [s][j/4][ld][4]
    unsigned int ld;                    // points of the arrays
    int S;
    unsigned int nPts, kPts;            // valid (padded to 128) points of the chunk; points per split (multiple of KC)
    int tilesJ, nsplit;
    double* g; int wi, wo;              // g[i*wo + j] += ...   for i < wi, j < wo
    double* gb;                         // bias gradient g(b_l)[j] += sum_p zbar_{l,0}[p][j] (taken from the B operand by the CTAs of row tile 0)
    int* err;
};

// K runs over thousands of points: the main accumulator rotates over three column sets in epochs of GW_EPOCH
// chunks and each set is drained into FP32 registers (round to nearest) two epochs later, when its MMAs have
// long completed, so no truncating chain is longer than GW_EPOCH * KC / 8 MMAs and the pipeline never stalls.
constexpr int GW_EPOCH = 4;

template <bool MN>
__global__ void __launch_bounds__(NTHR_ALL, 1) tc_gw_kernel(const GwArgs a) {
    constexpr int LAY = MN ? LAY_MN : LAY_QT;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (*reinterpret_cast<volatile int*>(a.err)) return;
    const int split = blockIdx.x % a.nsplit, tile = blockIdx.x / a.nsplit;
    const int it_ = tile / a.tilesJ, jt = tile - it_ * a.tilesJ;
    const unsigned int pBeg = (unsigned int)split * a.kPts;
    if (pBeg >= a.nPts) return;
    const unsigned int pEnd = min(pBeg + a.kPts, a.nPts);
    const int chunks = (int)((pEnd - pBeg) / KC);
    const int nIt = a.S * chunks;
    Bars bars;
    const uint32_t tmem = pipe_setup(smem, tid, warp, bars, MN ? MN_BAR_OFF : BAR_OFF);

    if (warp == NTHR / 32) {
        if (lane == 0) {
            const bool ok = mma_warp_loop<MN>(smem, bars, tmem, nIt, [&](int it, int kb, bool& fresh) {
                fresh = (it % GW_EPOCH == 0) && kb == 0;
                return (it / GW_EPOCH) % 3;
            });
            if (!ok) *a.err = 1;
        }
    } else {
        const float* Ag = a.A + (size_t)(it_ * TM / 4) * a.ld * 4;
        const float* Bg = a.B + (size_t)(jt * TN / 4) * a.ld * 4;
        auto src = [&](int it) {
            const int s = it / chunks, pc = it - s * chunks;
            return ChunkSrc{Opnd{Ag + s * a.aStream, a.ld}, Opnd{Bg + s * a.bStream, a.ld}, (int)pBeg + pc * KC};
        };
        const int half = warp >> 2;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * 64;
        float acc[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.f;
        // bias gradient: the loader of the zbar operand sees every (point, neuron) of stream 0 exactly once per row tile
        const bool doBias = a.gb != nullptr && it_ == 0;
        float bs[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) bs[c] = 0.f;
        auto biasAcc = [&](int it, const TileRegs& r) {
            if (doBias && it < chunks) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int o = MN ? 4 * (i >> 1) : 4 * i;       // LAY_MN: units 0, 1 are one quad (4 warp + qb), units 2, 3 the quad two further
                    bs[o] += r.v[i].x; bs[o + 1] += r.v[i].y; bs[o + 2] += r.v[i].z; bs[o + 3] += r.v[i].w;
                }
            }
        };
        auto drain = [&](int set) {
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float z[16];
                tmem_ld16(tlane + set * TN + cb * 16, z);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[cb * 16 + c] += z[c];
            }
        };
        TileRegs ra[2], rb[2];
        { const ChunkSrc c = src(0); tile_load<LAY>(c.a, c.k0, ra[0], tid); tile_load<LAY>(c.b, c.k0, rb[0], tid); }
        if (nIt > 1) { const ChunkSrc c = src(1); tile_load<LAY>(c.a, c.k0, ra[1], tid); tile_load<LAY>(c.b, c.k0, rb[1], tid); }
        bool ok = true;
#pragma unroll 1
        for (int it0 = 0; it0 < nIt; it0 += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int it = it0 + h;
                if (it < nIt) {
                    if (it % GW_EPOCH == 0 && it >= 2 * GW_EPOCH) {
                        // epoch e-2 is complete once the commit of chunk it-3 has been observed (GW_EPOCH + 1 >= NST)
                        if (ok) ok = mbar_wait(bars.empty + 8 * (it % NST), (uint32_t)((it / NST) - 1) & 1u);
                        tc_fence_after();
                        drain(((it / GW_EPOCH) - 2) % 3);
                        tc_fence_before();          // the set is overwritten by epoch e+1, whose first stage this thread publishes later
                    }
                    biasAcc(it, rb[h]);
                    loader_step<LAY>(smem, bars, it, nIt, ra[h], rb[h], src, tid, ok);
                }
            }
        }
        if (ok) ok = mbar_wait(bars.done, 0u);
        if (!ok) *a.err = 1;
        tc_fence_after();
        const int nEp = (nIt + GW_EPOCH - 1) / GW_EPOCH;
        for (int e = max(0, nEp - 2); e < nEp; ++e) drain(e % 3);           // the last two epochs were not drained in the loop
        drain(3);                                                            // small terms

        if (doBias && MN) {
            // LAY_MN loader mapping: lane bit 2 selects the quad of a pair (quads 4 warp + qb and 4 warp + 2 + qb); the other lane
            // bits run over points
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float t = bs[c];
                t += __shfl_xor_sync(0xffffffffu, t, 1); t += __shfl_xor_sync(0xffffffffu, t, 2);
                t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 16);
                const int qb = (lane >> 2) & 1;
                const int j = jt * TN + 4 * (4 * warp + 2 * (c >> 2) + qb) + (c & 3);
                if ((lane & 27) == 0 && j < a.wo) atomicAdd(a.gb + j, (double)t);
            }
        }
        if (doBias && !MN) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float t = bs[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                const int j = jt * TN + 16 * warp + c;                       // quad 4 warp + c / 4, neuron c % 4 (LAY_QT loader mapping)
                if (lane == 0 && j < a.wo) atomicAdd(a.gb + j, (double)t);
            }
        }
        const int i = it_ * TM + (warp & 3) * 32 + lane;
        if (ok && i < a.wi) {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int j = jt * TN + half * 64 + c;
                if (j < a.wo) atomicAdd(a.g + (size_t)i * a.wo + j, (double)acc[c]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------ FP32 kernels around the GEMMs
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;       // valid in thread 0
}

// zero-padded quad-major copies [K/4][WP][4] of the hidden kernels as B operands (B[n][k]):
//   forward  Wt: n = output neuron, k = input neuron  -> W_l[k][n]
//   adjoint  Wn: n = input neuron,  k = output neuron -> W_l[n][k]
__global__ void tc_stage_weights_kernel(NetDesc net, int WP, const float* __restrict__ theta, float* __restrict__ Wn, float* __restrict__ Wt) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)WP * WP;
    if (idx >= per * (net.L - 1)) return;
    const int l = (int)(idx / per) + 1;
    const int r = (int)(idx % per);
    const int k = (r / (4 * WP)) * 4 + (r & 3), n = (r >> 2) % WP;
    const int wi = net.width[l - 1], wo = net.width[l];
    Wt[idx] = (k < wi && n < wo) ? theta[net.woff[l] + k * wo + n] : 0.f;
    Wn[idx] = (n < wi && k < wo) ? theta[net.woff[l] + n * wo + k] : 0.f;
}

struct L0Args {
    TileArgs in; unsigned int base; int WP;
    float* X; float* Aqm; size_t sQm; unsigned int cap;
};
// layer 0 (K = inpDim): a = act(X W0 + b0), tangent k: act'(z) W0[k,:]; also keeps X of the chunk for g(W0).
// thread = one point x four neurons: coalesced 16-byte quad-major stores
template <int S, int ACT>
__global__ void __launch_bounds__(256) tc_layer0_kernel(const L0Args a) {
    const TileArgs& A = a.in;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int p = blockIdx.x * 32 + lane;
    const int n4 = blockIdx.y * 8 + warp, n0 = 4 * n4;
    const int inpDim = A.net.inpDim, w0 = A.net.width[0];
    const unsigned int gp = a.base + p;
    float x[VN_KIN];
#pragma unroll
    for (int k = 0; k < VN_KIN; ++k) {
        x[k] = 0.f;
        if (k < inpDim && gp < A.P)
            x[k] = (k >= A.nxTable) ? __ldg(A.extraX + (k - A.nxTable)) : __ldg(A.cols + (size_t)(A.colX + k) * A.pstride + table_row(A, gp));
        if (k < inpDim && n4 == 0) a.X[(size_t)k * a.cap + p] = x[k];
    }
    float o[S][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int n = n0 + e;
        float w[VN_KIN], z = (n < w0) ? __ldg(A.theta + A.net.boff[0] + n) : 0.f;
#pragma unroll
        for (int k = 0; k < VN_KIN; ++k) {
            w[k] = (k < inpDim && n < w0) ? __ldg(A.theta + A.net.woff[0] + k * w0 + n) : 0.f;
            z = fmaf(x[k], w[k], z);
        }
        const float av = act_f<ACT>(z), d1 = act_d1<ACT>(av);
        o[0][e] = av;
#pragma unroll
        for (int s = 1; s < S; ++s) o[s][e] = d1 * w[s - 1];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        *reinterpret_cast<float4*>(a.Aqm + s * a.sQm + ((size_t)n4 * a.cap + p) * 4) = make_float4(o[s][0], o[s][1], o[s][2], o[s][3]);
    }
}

struct OutArgs {
    TileArgs in; unsigned int base, nPts; int WP; bool needGrad;
    const float* Alast; size_t sQm;
    float* seeds; unsigned int cap; double* g64;
};
// output layer (Dense(1), linear): u_s = A_{L-1,s} . w_out (+ b_out); then per mode the integrand
// (TFModel.py:653-660), the BC/IC residual and its seed (TFModel.py:643-650), or the model value.
// thread = one point (quad-major rows: coalesced 16-byte loads, w_out broadcast)
template <int S, int MODE>
__global__ void __launch_bounds__(256) tc_out_kernel(const OutArgs a) {
    __shared__ double sh[8];
    const TileArgs& A = a.in;
    const unsigned int p = blockIdx.x * 256 + threadIdx.x;
    const int L = A.net.L, wlast = A.net.width[L - 1];
    const float* wout = A.theta + A.net.woff[L];
    float seedv = 0.f;
    if (p < a.nPts) {
        float u[S];
#pragma unroll
        for (int s = 0; s < S; ++s) u[s] = 0.f;
        for (int n = 0; n < wlast; n += 4) {
            const float w0 = __ldg(wout + n), w1 = n + 1 < wlast ? __ldg(wout + n + 1) : 0.f;
            const float w2 = n + 2 < wlast ? __ldg(wout + n + 2) : 0.f, w3 = n + 3 < wlast ? __ldg(wout + n + 3) : 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.Alast + s * a.sQm + ((size_t)(n >> 2) * a.cap + p) * 4));
                u[s] = fmaf(v.x, w0, fmaf(v.y, w1, fmaf(v.z, w2, fmaf(v.w, w3, u[s]))));
            }
        }
        u[0] += __ldg(A.theta + A.net.boff[L]);
        const unsigned int gp = a.base + p;
        if (MODE == TC_VAR) {
            if (gp < A.P) {
                float I = 0.f;
                const size_t row = table_row(A, gp);
#pragma unroll
                for (int k = 0; k < S - 1; ++k) I = fmaf(u[1 + k], __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + row), I);
                if (A.timeDependent) I -= u[0] * __ldg(A.cols + (size_t)A.colT * A.pstride + row);
                if (A.isSource) I -= __ldg(A.cols + (size_t)A.colS * A.pstride + row);
                if (A.integW) I *= __ldg(A.integW + (gp % A.integNum));
                A.Iw[gp] = I;
            }
        } else if (MODE == TC_BIC) {
            if (gp < A.P) {
                const float r = u[0] - __ldg(A.label + gp);
                A.cj[gp] = A.biDimVal * r * r;
                float sc;                                   // mean over boundary rows / initial rows (TFModel.py:644-648)
                if (gp < A.bDof) sc = __ldg(A.wts + 0) / (float)A.bDof;
                else sc = A.timeDependent ? __ldg(A.wts + 1) / (float)(A.P - A.bDof) : 0.f;
                seedv = 2.f * A.biDimVal * r * sc;
            }
            a.seeds[p] = seedv;
        } else {
            if (gp < A.P) A.uout[gp] = u[0];
        }
    }
    if (MODE == TC_BIC && a.needGrad) {
        const double t = block_sum_d((double)seedv, sh);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(a.g64 + A.net.boff[L], t);       // g(b_out) = sum of seeds
    }
}

struct SeedArgs {
    TileArgs in; unsigned int tf0, nTf, base; bool needGrad;
    float* seeds; unsigned int cap; double* lossAcc; double* g64;
};
// per test function: R_i = sum_q I_iq, lossVec_i = detJ_i R_i^2 (TFModel.py:659-668); adjoint seeds
// lambda = 2 w2 detJ_i w_q R_i, ubar = -lambda dNt, ubar_k = lambda gcoef_k (App. A.3)
template <int S>
__global__ void __launch_bounds__(256) tc_seed_kernel(const SeedArgs a) {
    __shared__ double sh[8];
    const TileArgs& A = a.in;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int b = blockIdx.x * 8 + warp;
    double lossv = 0.0, sb = 0.0;
    if (b < a.nTf) {
        const unsigned int i = a.tf0 + b;
        float r = 0.f;
        for (unsigned int q = lane; q < A.integNum; q += 32) r += A.Iw[(size_t)i * A.integNum + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
        if (lane == 0) {
            const float r2 = r * r;
            A.R[i] = r;
            A.lossVec[i] = dj * r2;
            lossv = A.detJvec ? (double)dj * (double)r2 : (double)r2;
        }
        if (a.needGrad) {
            const float w2 = __ldg(A.wts + 2);
            for (unsigned int q = lane; q < A.integNum; q += 32) {
                const unsigned int gp = i * A.integNum + q, p = gp - a.base;
                const size_t row = table_row(A, gp);
                const float wq = A.integW ? __ldg(A.integW + q) : 1.f;
                const float lam = 2.f * w2 * dj * wq * r;
                const float s0 = A.timeDependent ? -lam * __ldg(A.cols + (size_t)A.colT * A.pstride + row) : 0.f;
                a.seeds[p] = s0;
                sb += (double)s0;
#pragma unroll
                for (int k = 0; k < S - 1; ++k)
                    a.seeds[(size_t)(1 + k) * a.cap + p] = lam * __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + row);
            }
        }
    }
    const double tl = block_sum_d(lossv, sh);
    if (threadIdx.x == 0) atomicAdd(a.lossAcc, tl);
    if (a.needGrad) {
        const double tb = block_sum_d(sb, sh);
        if (threadIdx.x == 0) atomicAdd(a.g64 + A.net.boff[A.net.L], tb);
    }
}

struct TopArgs {
    int WP, wlast; const float* wout; unsigned int nPts;
    const float* Aqm; size_t sQm; const float* seeds; unsigned int cap;
    float* Dqm; double* gwout;
};
// top of the adjoint: zbar_{L-1} from the seeds (outer product with w_out), g(w_out).
// thread = one point x four neurons; a warp walks four 32-point groups and reduces its g(w_out) share once
template <int S, int ACT>
__global__ void __launch_bounds__(256) tc_top_kernel(const TopArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n4 = blockIdx.y * 8 + warp, n0 = 4 * n4;
    float wv[4], gacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) wv[e] = n0 + e < a.wlast ? __ldg(a.wout + n0 + e) : 0.f;
    for (int g = 0; g < 4; ++g) {
        const unsigned int p = (blockIdx.x * 4 + g) * 32 + lane;
        if (p >= a.nPts) break;
        const size_t q = ((size_t)n4 * a.cap + p) * 4;
        const float4 a4 = *reinterpret_cast<const float4*>(a.Aqm + q);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        const float s0 = a.seeds[p];
        float d1[4], cross[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) { d1[e] = act_d1<ACT>(av[e]); gacc[e] = fmaf(av[e], s0, gacc[e]); }
#pragma unroll
        for (int s = 1; s < S; ++s) {
            const float4 d4 = *reinterpret_cast<const float4*>(a.Aqm + s * a.sQm + q);
            const float da[4] = {d4.x, d4.y, d4.z, d4.w};
            const float sd = a.seeds[(size_t)s * a.cap + p];
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float ab = sd * wv[e];
                gacc[e] = fmaf(da[e], sd, gacc[e]);
                cross[e] = fmaf(ab, da[e], cross[e]);
                o[e] = ab * d1[e];
            }
            *reinterpret_cast<float4*>(a.Dqm + s * a.sQm + q) = make_float4(o[0], o[1], o[2], o[3]);
        }
        float zb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) zb[e] = fmaf(s0 * wv[e], d1[e], act_d2r<ACT>(av[e]) * cross[e]);
        *reinterpret_cast<float4*>(a.Dqm + q) = make_float4(zb[0], zb[1], zb[2], zb[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float t = gacc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && n0 + e < a.wlast) atomicAdd(a.gwout + n0 + e, (double)t);
    }
}

struct RowArgs {
    const float* Dqm; size_t sQm; unsigned int cap, nPts, len;   // len: points per split (multiple of 128)
    int quads, width, splits, first, streams; const float* X; int inpDim;
    double* gb; double* gw0;        // gb[j]; gw0[k*width + j] (layer 0 only)
};
// bias gradients g(b_l)[j] = sum_p zbar_l[p][j]; for layer 0 also g(W_0)[k][j] = sum_p X[p][k] zbar_0[p][j]
// plus, for the tangent streams, sum_p dzbar_0^k[p][j] added to row k   (App. A.3, l = 0).
// warp = (stream, neuron quad, point split); lanes walk the points (coalesced 16-byte quad-major units)
__global__ void __launch_bounds__(256) tc_rowsum_kernel(const RowArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= a.streams * a.quads * a.splits) return;
    const int split = w % a.splits, qi = w / a.splits;
    const int s = qi / a.quads, j4 = qi - s * a.quads;
    const unsigned int pBeg = (unsigned int)split * a.len, pEnd = min(pBeg + a.len, a.nPts);
    if (pBeg >= pEnd) return;
    const float* d = a.Dqm + s * a.sQm + (size_t)j4 * a.cap * 4;
    const bool dots = a.first && s == 0;
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, dot[VN_KIN][4];
#pragma unroll
    for (int k = 0; k < VN_KIN; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) dot[k][e] = 0.f;
    for (unsigned int p0 = pBeg + lane; p0 < pEnd; p0 += 128) {               // len is a multiple of 128: four independent loads in flight
        float4 vv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) vv[q] = *reinterpret_cast<const float4*>(d + (size_t)(p0 + 32 * q) * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
        const unsigned int p = p0 + 32 * q;
        const float4 v = vv[q];
        sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
        if (dots) {
#pragma unroll
            for (int k = 0; k < VN_KIN; ++k)
                if (k < a.inpDim) {
                    const float x = a.X[(size_t)k * a.cap + p];
                    dot[k][0] = fmaf(v.x, x, dot[k][0]); dot[k][1] = fmaf(v.y, x, dot[k][1]);
                    dot[k][2] = fmaf(v.z, x, dot[k][2]); dot[k][3] = fmaf(v.w, x, dot[k][3]);
                }
        }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[e] += __shfl_xor_sync(0xffffffffu, sum[e], o);
    if (dots) {
#pragma unroll
        for (int k = 0; k < VN_KIN; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot[k][e] += __shfl_xor_sync(0xffffffffu, dot[k][e], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            if (j >= a.width) break;
            if (s == 0) atomicAdd(a.gb + j, (double)sum[e]);
            else atomicAdd(a.gw0 + (size_t)(s - 1) * a.width + j, (double)sum[e]);
            if (dots)
                for (int k = 0; k < a.inpDim; ++k) atomicAdd(a.gw0 + (size_t)k * a.width + j, (double)dot[k][e]);
        }
    }
}

// ------------------------------------------------------------------ strong-form residual (monitoring path)
// res = -u_t + kappa sum_k d2u/dx_k^2 - sum_k (vel_k - dkappa/dx_k) du/dx_k + s   (TFModel.py:718-772) for wide networks.
// Called every few epochs on a test grid, so this is a plain FP32 kernel: thread = one point x four output neurons,
// six streams (value | tangents of the dim+1 inputs | second derivatives in x_k) kept neuron-major in global scratch.
struct ResLayerArgs {
    TileArgs in; unsigned int base, n, ld;      // chunk: first point, points, scratch row stride
    int l, nT, dim;                             // layer; tangent streams (dim + time); second-derivative streams
    const float* src; float* dst;               // [6][WP][ld] neuron-major, stream stride = WPs*ld
    size_t streamStride;
};
template <int ACT>
__global__ void __launch_bounds__(128) tc_res_layer_kernel(const ResLayerArgs a) {
    const TileArgs& A = a.in;
    const NetDesc& net = A.net;
    const unsigned int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= a.n) return;
    const int j0 = blockIdx.y * 4, l = a.l, wo = net.width[l];
    const int wi = l == 0 ? net.inpDim : net.width[l - 1];
    const float* W = A.theta + net.woff[l];
    float acc[VN_S_RES][4];
#pragma unroll
    for (int s = 0; s < VN_S_RES; ++s)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[s][e] = 0.f;
    if (l == 0) {
        const unsigned int gp = a.base + p;
        for (int k = 0; k < wi; ++k) {
            const float x = __ldg(A.cols + (size_t)(A.colX + k) * A.pstride + gp);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float w = j0 + e < wo ? __ldg(W + k * wo + j0 + e) : 0.f;
                acc[0][e] = fmaf(x, w, acc[0][e]);
                if (k < a.nT) acc[1 + k][e] = w;          // zdot_k = W0[k][j]; zddot = 0 at the input layer
            }
        }
    } else {
        for (int i = 0; i < wi; ++i) {
            float w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = j0 + e < wo ? __ldg(W + i * wo + j0 + e) : 0.f;
#pragma unroll
            for (int s = 0; s < VN_S_RES; ++s) {
                if (s == 0 || (s <= 3 && s - 1 < a.nT) || (s >= 4 && s - 4 < a.dim)) {
                    const float v = a.src[s * a.streamStride + (size_t)i * a.ld + p];
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[s][e] = fmaf(v, w[e], acc[s][e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = j0 + e;
        if (j >= wo) break;
        const float av = act_f<ACT>(acc[0][e] + __ldg(A.theta + net.boff[l] + j)), d1 = act_d1<ACT>(av);
        a.dst[(size_t)j * a.ld + p] = av;
#pragma unroll
        for (int s = 1; s < VN_S_RES; ++s) {
            float o = 0.f;
            if (s <= 3) { if (s - 1 < a.nT) o = d1 * acc[s][e]; }
            else if (s - 4 < a.dim) { const float zd = acc[s - 3][e]; o = d1 * fmaf(act_d2r<ACT>(av) * zd, zd, acc[s][e]); }
            a.dst[s * a.streamStride + (size_t)j * a.ld + p] = o;
        }
    }
}
__global__ void __launch_bounds__(128) tc_res_out_kernel(const ResLayerArgs a) {
    const TileArgs& A = a.in;
    const NetDesc& net = A.net;
    const unsigned int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= a.n) return;
    const int L = net.L, wl = net.width[L - 1];
    const float* wout = A.theta + net.woff[L];
    float u[VN_S_RES];
#pragma unroll
    for (int s = 0; s < VN_S_RES; ++s) u[s] = 0.f;
    for (int i = 0; i < wl; ++i) {
        const float w = __ldg(wout + i);
#pragma unroll
        for (int s = 0; s < VN_S_RES; ++s)
            if (s == 0 || (s <= 3 && s - 1 < a.nT) || (s >= 4 && s - 4 < a.dim)) u[s] = fmaf(a.src[s * a.streamStride + (size_t)i * a.ld + p], w, u[s]);
    }
    u[0] += __ldg(A.theta + net.boff[L]);
    const unsigned int gp = a.base + p;
    float res = A.timeDependent ? -u[1 + A.dim] : 0.f, lap = 0.f, adv = 0.f;
    for (int k = 0; k < A.dim; ++k) {
        lap += u[4 + k];
        const float vd = __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + gp) - __ldg(A.cols + (size_t)(A.colDD + k) * A.pstride + gp);
        adv = fmaf(vd, u[1 + k], adv);
    }
    res = fmaf(__ldg(A.cols + (size_t)A.colD * A.pstride + gp), lap, res) - adv;
    res += __ldg(A.cols + (size_t)A.colS * A.pstride + gp);
    A.Iw[gp] = res;
    A.uout[gp] = u[0];
}

__global__ void tc_grad_out_kernel(const double* __restrict__ g, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)g[i];
}

// ------------------------------------------------------------------ host side
struct Work {
    float *Wn, *Wt, *X, *Aqm, *Dqm, *seeds, *cross;
    size_t sQm, layerStride, bytes;          // stream stride of the quad-major arrays, layer stride
};
Work carve(void* base, int L, int S, int WP, unsigned int cap) {
    Work w;
    size_t off = 0;
    auto take = [&](size_t nfloats) { float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off); off += (nfloats * 4 + 255) & ~(size_t)255; return p; };
    w.sQm = (size_t)cap * WP; w.layerStride = (size_t)S * w.sQm;
    w.Wn = take((size_t)std::max(L - 1, 1) * WP * WP);
    w.Wt = take((size_t)std::max(L - 1, 1) * WP * WP);
    w.X = take((size_t)VN_KIN * cap);
    w.Aqm = take((size_t)L * w.layerStride);
    w.Dqm = take(2 * w.layerStride);
    w.seeds = take((size_t)S * cap);
    w.cross = take(w.sQm);
    w.bytes = off;
    return w;
}

bool g_ts = true;        // A-from-TMEM variant of the layer GEMMs (VARNET_B200_TC_TS=0 selects the all-shared-memory pipeline)
template <int EPI> cudaError_t launch_gemm(int act, const GemmArgs& g, int grid, cudaStream_t st) {
    if (g_ts) {
        if (act == VN_SIGMOID) tc_gemm_kernel<EPI, VN_SIGMOID, true><<<grid, NTHR, TS_SMEM_BYTES, st>>>(g);
        else tc_gemm_kernel<EPI, VN_TANH, true><<<grid, NTHR, TS_SMEM_BYTES, st>>>(g);
    } else {
        if (act == VN_SIGMOID) tc_gemm_kernel<EPI, VN_SIGMOID, false><<<grid, NTHR_ALL, SMEM_BYTES, st>>>(g);
        else tc_gemm_kernel<EPI, VN_TANH, false><<<grid, NTHR_ALL, SMEM_BYTES, st>>>(g);
    }
    return cudaGetLastError();
}
template <int EPI> cudaError_t prep_gemm() {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<EPI, VN_SIGMOID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_gemm_kernel<EPI, VN_TANH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_gemm_kernel<EPI, VN_SIGMOID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_gemm_kernel<EPI, VN_TANH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES);
}

template <int S> cudaError_t launch_layer0(int act, const L0Args& a, dim3 grid, cudaStream_t st) {
    if (act == VN_SIGMOID) tc_layer0_kernel<S, VN_SIGMOID><<<grid, 256, 0, st>>>(a);
    else tc_layer0_kernel<S, VN_TANH><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}
template <int S> cudaError_t launch_top(int act, const TopArgs& a, dim3 grid, cudaStream_t st) {
    if (act == VN_SIGMOID) tc_top_kernel<S, VN_SIGMOID><<<grid, 256, 0, st>>>(a);
    else tc_top_kernel<S, VN_TANH><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

#define TC_S_SWITCH(S, CALL)                \
    ((S) == 1 ? CALL(1) : ((S) == 2 ? CALL(2) : CALL(3)))

unsigned int gcd_u(unsigned int a, unsigned int b) { while (b) { const unsigned int t = a % b; a = b; b = t; } return a; }

// weight-gradient GEMM operands: round 1's K-major tiles written by transposing 4-byte stores (default), or MN-major tiles
// (VARNET_B200_TC_GW=mn: the quad-major 16-byte units stored as they are, a quarter of the store instructions, no bank
// conflicts).  Same-box A/B (scripts/ab_wide.py, gpurun_out/r2bl_ab.log): 4x256 237.0 / 237.3 ms K-major against 238.5 / 239.0 ms
// MN-major, 4x128 88.3 / 87.8 against 87.8 / 88.7 ms, identical loss: the kernel is bound by the delivery of its operands
// from L2 (32 KB per 12 MMAs), not by the loader's stores.
bool gw_mn() {
    const char* e = getenv("VARNET_B200_TC_GW");           // read per launch: a test switches it inside one process
    return e && e[0] == 'm';
}
}  // namespace

#define TCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

bool vn_tc_geometry(const NetDesc& net, int S, int numSMs, TcGeom* g) {
    int wmax = 0;
    for (int l = 0; l < net.L; ++l) wmax = std::max(wmax, net.width[l]);
    if (wmax > 256 || S < 1 || S > 3) return false;
    g->WP = wmax <= 128 ? 128 : 256;
    int waves = 8;                                                   // 128-point tiles per SM and chunk
    if (const char* w = getenv("VARNET_B200_TC_WAVES")) waves = std::max(1, std::min(16, atoi(w)));
    g->capPts = (unsigned int)numSMs * TM * waves;
    g->workBytes = carve(nullptr, net.L, S, g->WP, g->capPts).bytes;
    g->smemGemm = TS_SMEM_BYTES;         // layer GEMMs: A operand in tensor memory, weight tiles in shared memory (VARNET_B200_TC_TS=0: SMEM_BYTES)
    g->smemGw = gw_mn() ? MN_SMEM_BYTES : SMEM_BYTES;
    return true;
}

cudaError_t vn_tc_prepare(int S, int act) {
    (void)S; (void)act;
    if (const char* t = getenv("VARNET_B200_TC_TS")) g_ts = atoi(t) != 0;

    cudaError_t e = cudaFuncSetAttribute(tc_gw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MN_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if ((e = prep_gemm<EPI_FWD_VALUE>()) != cudaSuccess) return e;
    if ((e = prep_gemm<EPI_FWD_TANGENT>()) != cudaSuccess) return e;
    if ((e = prep_gemm<EPI_ADJ_TANGENT>()) != cudaSuccess) return e;
    if ((e = prep_gemm<EPI_ADJ_VALUE>()) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t vn_tc_stage_weights(const NetDesc& net, const TcGeom& g, const float* theta, void* work, cudaStream_t st) {
    if (net.L < 2) return cudaSuccess;
    const Work w = carve(work, net.L, 1, g.WP, g.capPts);             // Wn / Wt come first: independent of S
    const long long n = (long long)(net.L - 1) * g.WP * g.WP;
    tc_stage_weights_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(net, g.WP, theta, w.Wn, w.Wt);
    return cudaGetLastError();
}

cudaError_t vn_tc_residual(const TileArgs& A, int act, const TcGeom& g, void* work, cudaStream_t st, long long* launches) {
    const NetDesc& net = A.net;
    const int L = net.L, nT = A.dim + (A.timeDependent ? 1 : 0);
    int wmax = 0;
    for (int l = 0; l < L; ++l) wmax = std::max(wmax, net.width[l]);
    // two ping-pong buffers of six streams inside the chunk workspace
    const size_t avail = g.workBytes / sizeof(float);
    unsigned int chunk = (unsigned int)std::min<size_t>(g.capPts, avail / ((size_t)2 * VN_S_RES * wmax));
    chunk = chunk / 128 * 128;
    if (chunk == 0) return cudaErrorInvalidValue;
    float* buf0 = reinterpret_cast<float*>(work);
    float* buf1 = buf0 + (size_t)VN_S_RES * wmax * chunk;
    for (unsigned long long c0 = 0; c0 < A.P; c0 += chunk) {
        ResLayerArgs a; a.in = A; a.base = (unsigned int)c0; a.n = (unsigned int)std::min<unsigned long long>(chunk, A.P - c0); a.ld = chunk;
        a.nT = nT; a.dim = A.dim; a.streamStride = (size_t)wmax * chunk;
        float* src = buf1; float* dst = buf0;
        for (int l = 0; l < L; ++l) {
            a.l = l; a.src = src; a.dst = dst;
            const dim3 grid((a.n + 127) / 128, (net.width[l] + 3) / 4);
            if (act == VN_SIGMOID) tc_res_layer_kernel<VN_SIGMOID><<<grid, 128, 0, st>>>(a);
            else tc_res_layer_kernel<VN_TANH><<<grid, 128, 0, st>>>(a);
            TCK(cudaGetLastError());
            std::swap(src, dst);
            ++*launches;
        }
        a.src = src;
        tc_res_out_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a);
        TCK(cudaGetLastError());
        ++*launches;
    }
    return cudaSuccess;
}

cudaError_t vn_tc_grad_out(const double* g64, float* gbuf, int n, cudaStream_t st) {
    tc_grad_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(g64, gbuf, n);
    return cudaGetLastError();
}

cudaError_t vn_tc_run(TcJob& j) {
    const TileArgs& A = j.in;
    const NetDesc& net = A.net;
    const int L = net.L, S = j.S, WP = j.geom.WP, act = j.act;
    const unsigned int cap = j.geom.capPts;
    const Work w = carve(j.work, L, S, WP, cap);
    // the staged weights are carved with S = 1 by vn_tc_stage_weights: same offsets (they precede every S-dependent block)
    cudaStream_t st = j.st;
    j.launches = 0;

    // chunk size: whole test functions, whole 128-point tiles
    unsigned int chunk = cap;
    if (j.mode == TC_VAR) {
        const unsigned int base = A.integNum / gcd_u(A.integNum, TM) * TM;       // lcm(integNum, 128)
        if (base > cap) return cudaErrorInvalidValue;
        chunk = cap / base * base;
    }
    for (unsigned long long c0 = 0; c0 < A.P; c0 += chunk) {
        const unsigned int base = (unsigned int)c0;
        const unsigned int valid = (unsigned int)std::min<unsigned long long>(chunk, A.P - c0);
        const unsigned int nPts = (valid + TM - 1) / TM * TM;
        const int mTiles = (int)(nPts / TM);
        if (j.waitRows) TCK(j.waitRows(c0 + valid));

        // ---- layer 0
        {
            L0Args a; a.in = A; a.base = base; a.WP = WP; a.X = w.X; a.Aqm = w.Aqm; a.sQm = w.sQm; a.cap = cap;
            const dim3 grid(nPts / 32, WP / 32);
#define CALL(SS) launch_layer0<SS>(act, a, grid, st)
            TCK(TC_S_SWITCH(S, CALL));
#undef CALL
            j.launches++;
        }
        // ---- hidden layers 1..L-1 (forward): value stream first, the tangent streams need act' of its output
        for (int l = 1; l < L; ++l) {
            for (int sidx = 0; sidx < S; ++sidx) {
                GemmArgs g{};
                g.A = w.Aqm + (size_t)(l - 1) * w.layerStride + sidx * w.sQm; g.rows = cap;
                g.B = w.Wt + (size_t)(l - 1) * WP * WP; g.rowsB = WP;
                g.K = (net.width[l - 1] + KC - 1) / KC * KC;
                g.nTilesN = (net.width[l] + TN - 1) / TN;
                g.bias = A.theta + net.boff[l]; g.widthOut = net.width[l];
                g.val = w.Aqm + (size_t)l * w.layerStride;
                g.outQm = w.Aqm + (size_t)l * w.layerStride + sidx * w.sQm;
                g.err = j.err;
                if (sidx == 0) TCK(launch_gemm<EPI_FWD_VALUE>(act, g, mTiles * g.nTilesN, st));
                else TCK(launch_gemm<EPI_FWD_TANGENT>(act, g, mTiles * g.nTilesN, st));
                j.launches++;
            }
        }
        // ---- output layer + integrand / BC-IC residual / value
        {
            OutArgs a; a.in = A; a.base = base; a.nPts = nPts; a.WP = WP; a.needGrad = j.needGrad;
            a.Alast = w.Aqm + (size_t)(L - 1) * w.layerStride; a.sQm = w.sQm; a.seeds = w.seeds; a.cap = cap; a.g64 = j.g64;
            const unsigned int grid = (nPts + 255) / 256;
            if (j.mode == TC_VAR) {
#define CALL(SS) (tc_out_kernel<SS, TC_VAR><<<grid, 256, 0, st>>>(a), cudaGetLastError())
                TCK(TC_S_SWITCH(S, CALL));
#undef CALL
            } else if (j.mode == TC_BIC) {
                tc_out_kernel<1, TC_BIC><<<grid, 256, 0, st>>>(a);
                TCK(cudaGetLastError());
            } else {
                tc_out_kernel<1, TC_EVAL><<<grid, 256, 0, st>>>(a);
                TCK(cudaGetLastError());
            }
            j.launches++;
        }
        if (j.mode == TC_EVAL) continue;
        if (j.mode == TC_VAR) {
            if (j.needGrad) TCK(cudaMemsetAsync(w.seeds, 0, (size_t)S * cap * sizeof(float), st));
            SeedArgs a; a.in = A; a.tf0 = base / A.integNum; a.nTf = valid / A.integNum; a.base = base; a.needGrad = j.needGrad;
            a.seeds = w.seeds; a.cap = cap; a.lossAcc = j.lossAcc; a.g64 = j.g64;
            const unsigned int grid = (a.nTf + 7) / 8;
#define CALL(SS) (tc_seed_kernel<SS><<<grid, 256, 0, st>>>(a), cudaGetLastError())
            TCK(TC_S_SWITCH(S, CALL));
#undef CALL
            j.launches++;
        }
        if (!j.needGrad) continue;

        // ---- top of the adjoint: D_{L-1}
        int cur = 0;
        {
            TopArgs a; a.WP = WP; a.wlast = net.width[L - 1]; a.wout = A.theta + net.woff[L];
            a.Aqm = w.Aqm + (size_t)(L - 1) * w.layerStride; a.sQm = w.sQm; a.seeds = w.seeds; a.cap = cap; a.nPts = nPts;
            a.Dqm = w.Dqm; a.gwout = j.g64 + net.woff[L];
            const dim3 grid((nPts / 32 + 3) / 4, WP / 32);
#define CALL(SS) launch_top<SS>(act, a, grid, st)
            TCK(TC_S_SWITCH(S, CALL));
#undef CALL
            j.launches++;
        }
        auto rowsum = [&](int l, int curBuf) -> cudaError_t {
            RowArgs r;
            r.Dqm = w.Dqm + (size_t)curBuf * w.layerStride; r.sQm = w.sQm; r.cap = cap; r.nPts = nPts;
            r.first = (l == 0); r.streams = (l == 0 ? S : 1); r.quads = (net.width[l] + 3) / 4; r.width = net.width[l];
            const int rows = r.streams * r.quads;
            r.splits = std::max(1, std::min<int>((int)(nPts / 1024), (8 * j.numSMs * 8) / std::max(rows, 1)));
            r.len = ((nPts + r.splits - 1) / r.splits + TM - 1) / TM * TM;
            r.X = w.X; r.inpDim = net.inpDim;
            r.gb = j.g64 + net.boff[l]; r.gw0 = j.g64 + net.woff[0];
            const int warps = rows * r.splits;
            tc_rowsum_kernel<<<(warps + 7) / 8, 256, 0, st>>>(r);
            j.launches++;
            return cudaGetLastError();
        };
        for (int l = L - 1; l >= 1; --l) {
            // gW_l from (A_{l-1}, D_l)
            {
                GwArgs g{};
                g.A = w.Aqm + (size_t)(l - 1) * w.layerStride; g.aStream = w.sQm;
                g.B = w.Dqm + (size_t)cur * w.layerStride; g.bStream = w.sQm;
                g.ld = cap; g.S = S; g.nPts = nPts;
                const int tilesI = (net.width[l - 1] + TM - 1) / TM;
                g.tilesJ = (net.width[l] + TN - 1) / TN;
                const int tiles = tilesI * g.tilesJ;
                g.nsplit = std::max(1, std::min<int>(j.numSMs / tiles, (int)(nPts / 256)));
                g.kPts = ((nPts + g.nsplit - 1) / g.nsplit + KC - 1) / KC * KC;
                g.g = j.g64 + net.woff[l]; g.wi = net.width[l - 1]; g.wo = net.width[l]; g.gb = j.g64 + net.boff[l];
                g.err = j.err;
                if (gw_mn()) tc_gw_kernel<true><<<tiles * g.nsplit, NTHR_ALL, MN_SMEM_BYTES, st>>>(g);
                else tc_gw_kernel<false><<<tiles * g.nsplit, NTHR_ALL, SMEM_BYTES, st>>>(g);
                TCK(cudaGetLastError());
                j.launches++;
            }
            // D_{l-1} = through act'/act'' of layer l-1 of (D_l W_l^T): tangent streams first (they build the
            // second-order term `cross`), then the value stream
            for (int sidx = S - 1; sidx >= 0; --sidx) {
                GemmArgs g{};
                g.A = w.Dqm + (size_t)cur * w.layerStride + sidx * w.sQm; g.rows = cap;
                g.B = w.Wn + (size_t)(l - 1) * WP * WP; g.rowsB = WP;
                g.K = (net.width[l] + KC - 1) / KC * KC;
                g.nTilesN = (net.width[l - 1] + TN - 1) / TN;
                g.val = w.Aqm + (size_t)(l - 1) * w.layerStride;
                g.tan = w.Aqm + (size_t)(l - 1) * w.layerStride + sidx * w.sQm;
                g.cross = w.cross;
                g.outQm = w.Dqm + (size_t)(cur ^ 1) * w.layerStride + sidx * w.sQm;
                g.err = j.err;
                if (sidx > 0) { g.crossMode = (sidx == S - 1) ? 0 : 1; TCK(launch_gemm<EPI_ADJ_TANGENT>(act, g, mTiles * g.nTilesN, st)); }
                else { g.crossMode = S > 1 ? 1 : 0; TCK(launch_gemm<EPI_ADJ_VALUE>(act, g, mTiles * g.nTilesN, st)); }
                j.launches++;
            }
            cur ^= 1;
        }
        TCK(rowsum(0, cur));
    }
    return cudaSuccess;
}
