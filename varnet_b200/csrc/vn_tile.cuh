// vn_tile.cuh — fused FP32 tile kernels for the weak-form residual and its adjoint.
//
// One CTA owns a tile of TP quadrature points and walks the whole MLP for them:
//   stream 0      : a_l        = act(a_{l-1} W_l + b_l)                (model(Input),   TFModel.py:625)
//   stream 1..DIM : da_l/dx_k  = act'(z_l) * (da_{l-1}/dx_k W_l)       (tf.gradients(model(Input), Input)[:, :dim], TFModel.py:536-541,
//                                                                        evaluated forward-mode: same numbers, no second sweep)
// Each layer is a [S*TP x Kin] x [Kin x W] register-tiled FP32-FMA GEMM whose A operand
// (activations, point-contiguous) and B operand (weights) are read from shared memory with
// 128-bit loads; the epilogue applies the activation and writes the next layer's operand.
//
// vn_fwd_kernel : forward only (loss / evaluation), two ping-pong operand buffers.
// vn_adj_kernel : forward + adjoint in one pass over the tile.  Shared memory holds
//   three operand buffers; the activations of layers 0..L-3 are stashed in a per-CTA global slab that
//   stays L2-resident (148 CTAs x <=2 x 52 KB) and are brought back with cp.async while the
//   abar = zbar W^T GEMM of the layer above runs.  Per layer the adjoint does two GEMMs:
//   abar_{l-1} = zbar_l W_l^T and gW_l += [a_{l-1}; da_{l-1}]^T [zbar_l; dzbar_l].
//   When integNum divides TP the per-test-function residual R_i = sum_q w_q I_iq (TFModel.py:659-661)
//   is reduced inside the tile (MODE_VAR_FUSED) so the step needs no separate forward pass.
// Weight-gradient tiles are accumulated per CTA in FP64 partial slabs (thread-private slots, no
// atomics => bitwise deterministic) and summed across CTAs by vn_finalize_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vn_pdl.cuh"

#define VN_MAX_LAYERS 8
#define VN_KIN 8          // padded MLP input rows (== VN_MAX_INPDIM)

enum { VN_SIGMOID = 0, VN_TANH = 1 };

enum {
    MODE_VAR_FWD = 0,   // I_p per point                       (TFModel.py:653-660)
    MODE_EVAL = 1,      // u = model(X)                        (VarNetUtility.py:1127)
    MODE_BIC_FWD = 2,   // biDimVal*(model(biInput)-biLabel)^2 (TFModel.py:643)
    MODE_VAR_ADJ = 3,   // d(w2*varLoss)/d theta, R_i read from global (any integNum)
    MODE_BIC_ADJ = 4,   // d(w0*bCs+w1*iCs)/d theta
    MODE_VAR_FUSED = 5, // loss + d(w2*varLoss)/d theta, R_i reduced in the tile (integNum | TP)
    MODE_RESIDUAL = 6   // strong-form residual -u_t + kappa Lap u - (vel - grad kappa).grad u + s (TFModel.py:718-772)
};
#define VN_S_RES 6      // residual streams: value | 3 input tangents (x.., t) | 2 second derivatives (x_0, x_1)

struct NetDesc {
    int L;                          // hidden layers
    int inpDim;
    int width[VN_MAX_LAYERS];       // hidden widths
    int wpad[VN_MAX_LAYERS];        // rounded up to a multiple of 4
    int woff[VN_MAX_LAYERS + 1];    // flat offset of kernel l (l == L: output layer)
    int boff[VN_MAX_LAYERS + 1];    // flat offset of bias l
    int nparam;
};

// In-kernel table generation: for a uniform space-time mesh with constant coefficients the point table is
// periodic in the Gauss index q, so a kernel can rebuild a row from the centre of its test function instead of reading it:
//   X[i,q,d] = coord[s][d] + hd[d][q]  (d < dim),  X[i,q,dim] = tcoord[j] + hd[dim][q],  i = tf0 + table test function = s*nTime + j
// in float64 exactly as the host does (VarNet.py:576-586), rounded to float32; gcoef / dNt / source*N come from a q-indexed table.
struct GenTab {
    const double* coord;            // [nSpace][dim] centres
    const double* tcoord;           // [nTime] or nullptr
    const double* hd;               // [feDim][q]: h[d] * delta[d][q]
    const float* coef;              // [q][4]: gcoef_0, gcoef_1, dNt, source*N (float32, rounded like the feed cast)
    long long nTime, tf0;
    int q, dim, feDim;
};

struct TileArgs {
    NetDesc net;
    const float* theta;             // flat parameters (device)
    const float* cols;              // SoA point table: column c at cols + c*pstride
    long long pstride;              // multiple of the tile size, zero padded
    int colX, colG, colT, colS;     // first column of X, gcoef (residual: vel), dNt, source*N (residual: source); -1 = absent
    int colD, colDD;                // residual: diffusivity column, first grad(kappa) column
    int nxTable;                    // MLP input columns stored in the table; inputs [nxTable, inpDim) are constants
    const float* extraX;            // [inpDim - nxTable] constant trailing inputs (MOR parameters), device
    int useGen;                     // 1: no materialised table, rows are regenerated from `gen` (tensor-core tile kernel only)
    GenTab gen;
    const int* tfIndex;             // mini-batch: table test-function index of batch slot b, or nullptr (identity)
    int dim;                        // spatial dimension (residual kernel: runtime stream roles)
    unsigned int P;                 // valid points (rows)
    int ntiles;
    int tile0;                      // first tile of this launch (adjoint kernels launched per uploaded chunk, vn_loss_grad_fed)
    int accumulate;                 // 1: add to the partial slabs / loss partials of the previous launches of this step
    int timeDependent, isSource;
    // variational term
    unsigned int integNum;
    const float* integW;            // [integNum] or nullptr
    const float* detJ;              // [1] or [nb]
    int detJvec;
    float* R;                       // [nb] per-test-function residuals (read: VAR_ADJ, written: VAR_FUSED)
    float* lossVec;                 // [nb] detJ_i R_i^2 (written by VAR_FUSED)
    const float* wts;               // [3] loss weights (device)
    float* Iw;                      // [P] weighted integrand out (VAR_FWD)
    // boundary / initial rows
    const float* label;             // [nbi]
    unsigned int bDof;
    float biDimVal;
    float* cj;                      // [nbi] biDimVal*(u-label)^2 out
    // evaluation
    float* uout;                    // [P]
    // adjoint
    double* part;                   // [gridDim.x][psz] FP64 partial gradients (folded from part32 every VN_FOLD tiles)
    float* part32;                  // [gridDim.x][psz] FP32 window accumulators (same slot layout)
    int psz;
    float* stash;                   // [gridDim.x][stashFloats] activations of layers 0..L-3
    long long stashFloats;
    double* lossPart;               // [gridDim.x][NT/32] partial sums of (detJ_i) R_i^2 (VAR_FUSED)
};

template <int S_, int WP_, int TP_, int TN_, int ACT_>
struct TileCfg {
    static constexpr int S = S_;            // streams: value + DIM input tangents
    static constexpr int WP = WP_;          // padded hidden width class
    static constexpr int TP = TP_;          // points per tile
    static constexpr int TN = TN_;          // neurons per thread in the layer GEMMs
    static constexpr int ACT = ACT_;
    static constexpr int TPS = TP + 4;      // padded point stride  (bank-conflict-free row walks)
    static constexpr int WS = WP + 4;       // padded weight row stride
    static constexpr int NPG = TP / 4;      // point groups (4 points per thread)
    static constexpr int NNG = WP / TN;     // neuron groups
    static constexpr int NT = NPG * NNG;    // threads per CTA
    static constexpr int NW = NT / 32;
    static constexpr int NPGW = NPG / 8;    // warps along the point axis
    static constexpr int KIN = VN_KIN;
    // weight-gradient GEMM: 128 tile owners (8 row groups x 16 column groups) x KS point slices
    static constexpr int KS = NT / 128;
    static constexpr int TI = WP / 8;       //   rows per thread    (i = ig + 8 t)
    static constexpr int TJ = WP / 16;      //   columns per thread (j = jg + 16 u)
    static constexpr int BUF = S * WP * TPS;
    static constexpr int BUF0 = S * KIN * TPS;
    static_assert(NPG % 8 == 0 && NNG % 4 == 0, "warp mapping needs 8 point groups x 4 neuron groups per warp");
    static_assert(NT % 128 == 0 && TP % (4 * KS) == 0, "weight-gradient tiling");
    static_assert(NT % WP == 0 && (TP / (NT / WP)) % 4 == 0, "output-layer gradient mapping");
};

// layout of one CTA's FP64 partial-gradient slab (shared by the adjoint kernel and finalize)
struct PartLayout {
    int NT, KS, TI, TJ, WP;
    int off_gw[VN_MAX_LAYERS];
    int off_gb[VN_MAX_LAYERS];
    int off_wout, off_bout, psz;
};

template <class C>
__host__ __device__ inline PartLayout make_part_layout(int L) {
    PartLayout p;
    p.NT = C::NT; p.KS = C::KS; p.TI = C::TI; p.TJ = C::TJ; p.WP = C::WP;
    int o = 0;
    for (int l = 0; l < VN_MAX_LAYERS; ++l) { p.off_gw[l] = 0; p.off_gb[l] = 0; }
    for (int l = 0; l < L; ++l) {
        p.off_gw[l] = o; o += C::NT * (((l == 0 ? C::TJ : C::TI * C::TJ) + 1) & ~1);     // pair-interleaved: even slot count
        p.off_gb[l] = o; o += C::KS * C::WP;
    }
    p.off_wout = o; o += C::NT;
    p.off_bout = o; o += 1;
    p.psz = (o + 1) & ~1;
    return p;
}

template <class C>
__host__ __device__ inline size_t tile_smem_floats(int L, bool adj) {
    size_t n = 0;
    n += C::KIN * C::WS;                    // W0
    n += (size_t)(L - 1) * C::WP * C::WS;   // W1..W_{L-1}
    n += (size_t)L * C::WP;                 // biases
    n += C::WP + 4;                         // wout | bout
    n += C::BUF0;                           // layer "-1": inputs + unit tangent rows
    n += (size_t)(adj ? 3 : 2) * C::BUF;    // operand buffers
    n += 4 * C::TP;                         // integrand / per-tile scratch
    n += (size_t)C::S * C::TP;              // u / seeds
    return n;
}
template <class C>
__host__ __device__ inline long long tile_stash_floats(int L) { return L > 2 ? (long long)(L - 2) * C::BUF : 0; }

// ------------------------------------------------------------------ activations
template <int ACT> __device__ __forceinline__ float act_f(float z) {
    if (ACT == VN_SIGMOID) return __fdividef(1.0f, 1.0f + expf(-z));
    return tanhf(z);
}
template <int ACT> __device__ __forceinline__ float act_d1(float a) {      // act'(z) in terms of a
    return ACT == VN_SIGMOID ? a * (1.0f - a) : fmaf(-a, a, 1.0f);
}
template <int ACT> __device__ __forceinline__ float act_d2r(float a) {     // act''(z)/act'(z)
    return ACT == VN_SIGMOID ? fmaf(-2.0f, a, 1.0f) : -2.0f * a;
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float f4get(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

#define VN_FOLD 32   // tiles accumulated in the FP32 window slab before it is folded into the FP64 slab

// ---- packed FP32 FMA (Blackwell fma.rn.f32x2 -> SASS FFMA2): two IEEE-rn FMAs per issued instruction.
// The GEMM cores keep their accumulators as point pairs so that the activation fragment (a float4 of four
// consecutive points) is used directly as two packed operands; only the weight is duplicated into a pair.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ ulonglong2 lds2x64(const float* p) { return *reinterpret_cast<const ulonglong2*>(p); }

// out[s][j][p] = sum_i in[s][i][p] * W[i][j]; thread tile: 4 points x S streams x TN neurons
template <class C, int KD>
__device__ __forceinline__ void fwd_gemm(const float* __restrict__ Bin, const float* __restrict__ Wm,
                                         int Kin, int p0, int j0, float (&acc)[C::S][4][C::TN]) {
    u64 acc2[C::S][2][C::TN];
#pragma unroll
    for (int s = 0; s < C::S; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int t = 0; t < C::TN; ++t) acc2[s][h][t] = 0ull;
#pragma unroll 8
    for (int i = 0; i < Kin; ++i) {
        u64 w2[C::TN];
        if (C::TN % 4 == 0) {
#pragma unroll
            for (int t4 = 0; t4 < C::TN / 4; ++t4) {
                float4 v = lds4(Wm + i * C::WS + j0 + 4 * t4);
                w2[4 * t4 + 0] = pack2(v.x, v.x); w2[4 * t4 + 1] = pack2(v.y, v.y);
                w2[4 * t4 + 2] = pack2(v.z, v.z); w2[4 * t4 + 3] = pack2(v.w, v.w);
            }
        } else {
#pragma unroll
            for (int t = 0; t < C::TN; ++t) { const float v = Wm[i * C::WS + j0 + t]; w2[t] = pack2(v, v); }
        }
#pragma unroll
        for (int s = 0; s < C::S; ++s) {
            const ulonglong2 a = lds2x64(Bin + (s * KD + i) * C::TPS + p0);
#pragma unroll
            for (int t = 0; t < C::TN; ++t) {
                ffma2(acc2[s][0][t], a.x, w2[t]);
                ffma2(acc2[s][1][t], a.y, w2[t]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < C::S; ++s)
#pragma unroll
        for (int t = 0; t < C::TN; ++t) {
            unpack2(acc2[s][0][t], acc[s][0][t], acc[s][1][t]);
            unpack2(acc2[s][1][t], acc[s][2][t], acc[s][3][t]);
        }
}

// out[s][i_t][p] = sum_j D[s][j][p] * W[i_t][j];  thread owns interleaved rows i_t = ng + NNG*t
template <class C>
__device__ __forceinline__ void adj_gemm(const float* __restrict__ Dm, const float* __restrict__ Wm,
                                         int Kout, int p0, int ng, float (&acc)[C::S][4][C::TN]) {
    u64 acc2[C::S][2][C::TN];
#pragma unroll
    for (int s = 0; s < C::S; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int t = 0; t < C::TN; ++t) acc2[s][h][t] = 0ull;
#pragma unroll 2
    for (int j0 = 0; j0 < Kout; j0 += 4) {
        float4 w[C::TN];
#pragma unroll
        for (int t = 0; t < C::TN; ++t) w[t] = lds4(Wm + (ng + C::NNG * t) * C::WS + j0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            u64 w2[C::TN];
#pragma unroll
            for (int t = 0; t < C::TN; ++t) { const float wv = f4get(w[t], jj); w2[t] = pack2(wv, wv); }
#pragma unroll
            for (int s = 0; s < C::S; ++s) {
                const ulonglong2 d = lds2x64(Dm + (s * C::WP + j0 + jj) * C::TPS + p0);
#pragma unroll
                for (int t = 0; t < C::TN; ++t) {
                    ffma2(acc2[s][0][t], d.x, w2[t]);
                    ffma2(acc2[s][1][t], d.y, w2[t]);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < C::S; ++s)
#pragma unroll
        for (int t = 0; t < C::TN; ++t) {
            unpack2(acc2[s][0][t], acc[s][0][t], acc[s][1][t]);
            unpack2(acc2[s][1][t], acc[s][2][t], acc[s][3][t]);
        }
}

// gW[i][j] += sum_{s,p} Bprev[s][i][p] * D[s][j][p];  i = ig + 8t, j = jg + 16u; the CTA's KS point
// slices accumulate into separate partial slots (no cross-thread reduction).  ig == 0 threads also
// produce gb[j] = sum_p D[0][j][p].  `first` overwrites instead of accumulating.
template <class C, int KD, int TIK>
__device__ __forceinline__ void gw_gemm(const float* __restrict__ Bprev, const float* __restrict__ Dm,
                                        int ig, int jg, int kslice, float* __restrict__ pgw,
                                        float* __restrict__ pgb, bool first) {
    // packed accumulators: (sum over even points, sum over odd points) per (i, j); added at the end
    u64 acc2[TIK][C::TJ];
    float bacc[C::TJ];
#pragma unroll
    for (int t = 0; t < TIK; ++t)
#pragma unroll
        for (int u = 0; u < C::TJ; ++u) acc2[t][u] = 0ull;
#pragma unroll
    for (int u = 0; u < C::TJ; ++u) bacc[u] = 0.f;
    // FP32 window slab, pair-interleaved so that a warp's accesses are contiguous: element r = t*TJ+u of
    // thread `tid` lives at pgw[(r>>1)*2*NT + (r&1)] (pgw already includes 2*tid).  Every slot has exactly one
    // writer, so accumulating with fire-and-forget reductions (RED.ADD.F32: no load, no conversion, no
    // scoreboard wait) stays bitwise deterministic; the first tile of a window overwrites instead.
    constexpr int PSL = C::TP / C::KS;
    const int pbeg = kslice * PSL;
#pragma unroll
    for (int s = 0; s < C::S; ++s) {
#pragma unroll 2
        for (int p = pbeg; p < pbeg + PSL; p += 4) {
            ulonglong2 a[TIK], d[C::TJ];
#pragma unroll
            for (int t = 0; t < TIK; ++t) a[t] = lds2x64(Bprev + (s * KD + ig + 8 * t) * C::TPS + p);
#pragma unroll
            for (int u = 0; u < C::TJ; ++u) d[u] = lds2x64(Dm + (s * C::WP + jg + 16 * u) * C::TPS + p);
            // two passes so that consecutive FFMA2s never hit the same accumulator back to back
#pragma unroll
            for (int t = 0; t < TIK; ++t)
#pragma unroll
                for (int u = 0; u < C::TJ; ++u) ffma2(acc2[t][u], a[t].x, d[u].x);
#pragma unroll
            for (int t = 0; t < TIK; ++t)
#pragma unroll
                for (int u = 0; u < C::TJ; ++u) ffma2(acc2[t][u], a[t].y, d[u].y);
            if (s == 0 && ig == 0) {
#pragma unroll
                for (int u = 0; u < C::TJ; ++u) {
                    float d0, d1, d2, d3;
                    unpack2(d[u].x, d0, d1); unpack2(d[u].y, d2, d3);
                    bacc[u] += (d0 + d1) + (d2 + d3);
                }
            }
        }
    }
    float acc[TIK][C::TJ];
#pragma unroll
    for (int t = 0; t < TIK; ++t)
#pragma unroll
        for (int u = 0; u < C::TJ; ++u) { float lo, hi; unpack2(acc2[t][u], lo, hi); acc[t][u] = lo + hi; }
    constexpr int NE = TIK * C::TJ;
#pragma unroll
    for (int r = 0; r < NE; r += 2) {
        float* q = pgw + (r >> 1) * (2 * C::NT);
        const float v0 = acc[r / C::TJ][r % C::TJ];
        const float v1 = (r + 1 < NE) ? acc[(r + 1 < NE ? r + 1 : r) / C::TJ][(r + 1 < NE ? r + 1 : r) % C::TJ] : 0.f;
        if (first) __stcg(reinterpret_cast<float2*>(q), make_float2(v0, v1));
        else { atomicAdd(q, v0); if (r + 1 < NE) atomicAdd(q + 1, v1); }
    }
    if (ig == 0) {
#pragma unroll
        for (int u = 0; u < C::TJ; ++u) {
            if (first) __stcg(pgb + u, bacc[u]);
            else atomicAdd(pgb + u, bacc[u]);
        }
    }
}

// Fold this thread's slots of the FP32 window slab into the FP64 slab (first fold overwrites).
template <class C>
__device__ __forceinline__ void fold_slab(const PartLayout& pl, int L, const float* __restrict__ p32,
                                          double* __restrict__ p64, int tid, int ig, int jg, int kslice, bool firstFold) {
    auto put = [&](int idx) {
        const double v = (double)__ldcg(p32 + idx);
        if (firstFold) __stcg(p64 + idx, v); else atomicAdd(p64 + idx, v);
    };
    for (int l = 0; l < L; ++l) {
        const int n = l == 0 ? C::TJ : C::TI * C::TJ;
        for (int r = 0; r < n; ++r) put(pl.off_gw[l] + (r >> 1) * (2 * C::NT) + 2 * tid + (r & 1));
        if (ig == 0)
            for (int u = 0; u < C::TJ; ++u) put(pl.off_gb[l] + kslice * C::WP + jg * C::TJ + u);
    }
    put(pl.off_wout + tid);
    if (tid == 0) put(pl.off_bout);
}

// ---- pieces shared by the two kernels -------------------------------------------------------------
template <class C>
struct SmemMap {
    float *W0, *Wl, *bias, *wout, *Bm1, *A, *coef, *us;
    __device__ SmemMap(float* smem, int L, int nbuf) {
        W0 = smem;
        Wl = W0 + C::KIN * C::WS;
        bias = Wl + (L - 1) * C::WP * C::WS;
        wout = bias + L * C::WP;
        Bm1 = wout + C::WP + 4;
        A = Bm1 + C::BUF0;
        coef = A + nbuf * C::BUF;
        us = coef + 4 * C::TP;
    }
};

// zero shared memory, stage the weights (zero padded), write the unit tangent rows of layer "-1"
template <class C>
__device__ __forceinline__ void stage_network(const TileArgs& A, const SmemMap<C>& m, float* smem, int total,
                                              int nunit = C::S - 1) {
    const NetDesc& net = A.net;
    const int tid = threadIdx.x, L = net.L;
    constexpr int NT = C::NT, WS = C::WS, WP = C::WP;
    for (int i = tid; i < total; i += NT) smem[i] = 0.f;
    // k-steps-in-one-graph path (vn_pdl.cuh): the weights below were written by the previous step's reduction kernel
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    const float* th = A.theta;
    auto ld = [](const float* q) { return __ldcg(q); };                 // coherent loads: never ld.global.nc behind pdl_wait
    for (int idx = tid; idx < net.inpDim * net.width[0]; idx += NT) {
        int i = idx / net.width[0], j = idx - i * net.width[0];
        m.W0[i * WS + j] = ld(th + net.woff[0] + idx);
    }
    for (int l = 1; l < L; ++l) {
        const int wi = net.width[l - 1], wo = net.width[l];
        float* Wm = m.Wl + (l - 1) * WP * WS;
        for (int idx = tid; idx < wi * wo; idx += NT) {
            int i = idx / wo, j = idx - i * wo;
            Wm[i * WS + j] = ld(th + net.woff[l] + idx);
        }
    }
    for (int l = 0; l < L; ++l)
        for (int j = tid; j < net.width[l]; j += NT) m.bias[l * WP + j] = ld(th + net.boff[l] + j);
    for (int j = tid; j < net.width[L - 1]; j += NT) m.wout[j] = ld(th + net.woff[L] + j);
    if (tid == 0) m.wout[WP] = ld(th + net.boff[L]);
    // d x_c / d x_k = delta_ck  (stream 1+k seeds input column k)
    for (int idx = tid; idx < nunit * C::TP; idx += NT) {
        int k = idx / C::TP, p = idx - k * C::TP;
        m.Bm1[((1 + k) * C::KIN + k) * C::TPS + p] = 1.f;
    }
    __syncthreads();
}

// one forward layer: GEMM + bias + activation -> Bout (and optionally the global stash)
template <class C, bool RES = false>
__device__ __forceinline__ void forward_layer(const NetDesc& net, const SmemMap<C>& m, int l, const float* Bin,
                                              float* Bout, float* stash, int p0, int ng) {
    constexpr int S = C::S, WP = C::WP, TN = C::TN, TPS = C::TPS;
    const int j0 = TN * ng;
    if (j0 >= net.wpad[l]) return;
    float acc[S][4][TN];
    if (l == 0) fwd_gemm<C, C::KIN>(m.Bm1, m.W0, (net.inpDim + 3) & ~3, p0, j0, acc);
    else fwd_gemm<C, WP>(Bin, m.Wl + (l - 1) * WP * C::WS, net.wpad[l - 1], p0, j0, acc);
#pragma unroll
    for (int t = 0; t < TN; ++t) {
        const float b = m.bias[l * WP + j0 + t];
        float a[4], d1[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            a[p] = act_f<C::ACT>(acc[0][p][t] + b);
            d1[p] = act_d1<C::ACT>(a[p]);
        }
        const int off0 = (j0 + t) * TPS + p0;
        const float4 v0 = make_float4(a[0], a[1], a[2], a[3]);
        sts4(Bout + off0, v0);
        if (stash) __stcg(reinterpret_cast<float4*>(stash + off0), v0);
        if constexpr (RES) {
            // streams 1..3: first derivatives d1*zdot_k; streams 4,5: second derivatives
            // d2a/dx_k^2 = act''(z) zdot_k^2 + act'(z) zddot_k   (act'' = act' * act''/act')
#pragma unroll
            for (int s = 1; s < S; ++s) {
                float o[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (s <= 3) o[p] = d1[p] * acc[s][p][t];
                    else {
                        const float zd = acc[s - 3][p][t];
                        o[p] = d1[p] * fmaf(act_d2r<C::ACT>(a[p]) * zd, zd, acc[s][p][t]);
                    }
                }
                sts4(Bout + (s * WP + j0 + t) * TPS + p0, make_float4(o[0], o[1], o[2], o[3]));
            }
        } else {
#pragma unroll
            for (int s = 1; s < S; ++s) {
                const int off = (s * WP + j0 + t) * TPS + p0;
                const float4 v = make_float4(d1[0] * acc[s][0][t], d1[1] * acc[s][1][t], d1[2] * acc[s][2][t],
                                             d1[3] * acc[s][3][t]);
                sts4(Bout + off, v);
                if (stash) __stcg(reinterpret_cast<float4*>(stash + off), v);
            }
        }
    }
}

// batch point -> table row: mini-batches address the resident table through a test-function index list
__device__ __forceinline__ size_t table_row(const TileArgs& A, unsigned int gp) {
    if (!A.tfIndex) return gp;
    const unsigned int b = gp / A.integNum;
    return (size_t)__ldg(A.tfIndex + b) * A.integNum + (gp - b * A.integNum);
}
__device__ __forceinline__ unsigned int table_tf(const TileArgs& A, unsigned int b) {
    return A.tfIndex ? (unsigned int)__ldg(A.tfIndex + b) : b;
}

// stage the tile's input columns into layer "-1" (coalesced 128-bit loads from the SoA table; groups of 4
// points never straddle a test function in indexed mode because integNum % 4 == 0 there)
template <class C>
__device__ __forceinline__ void load_inputs(const TileArgs& A, const SmemMap<C>& m, unsigned int base) {
    for (int idx = threadIdx.x; idx < A.net.inpDim * C::NPG; idx += C::NT) {
        const int c = idx / C::NPG, q = idx - c * C::NPG;
        const unsigned int gp = base + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= A.nxTable) {
            const float x = __ldg(A.extraX + (c - A.nxTable));
            v = make_float4(x, x, x, x);
        } else if (!A.tfIndex) {
            v = __ldg(reinterpret_cast<const float4*>(A.cols + (size_t)(A.colX + c) * A.pstride + gp));   // padded table
        } else if (gp < A.P) {
            v = __ldg(reinterpret_cast<const float4*>(A.cols + (size_t)(A.colX + c) * A.pstride + table_row(A, gp)));
        }
        sts4(m.Bm1 + c * C::TPS + 4 * q, v);
    }
}

// output layer (Dense(1), linear): us[s][p] = u (s=0) and du/dx_k (s=1+k)
template <class C>
__device__ __forceinline__ void output_layer(const SmemMap<C>& m, const float* Blast, int wlast) {
    for (int idx = threadIdx.x; idx < C::S * C::TP; idx += C::NT) {
        const int s = idx / C::TP, p = idx - s * C::TP;
        float acc0 = 0.f, acc1 = 0.f;
        for (int i = 0; i < wlast; i += 2) {
            acc0 = fmaf(Blast[(s * C::WP + i) * C::TPS + p], m.wout[i], acc0);
            acc1 = fmaf(Blast[(s * C::WP + i + 1) * C::TPS + p], m.wout[i + 1], acc1);
        }
        float acc = acc0 + acc1;
        if (s == 0) acc += m.wout[C::WP];
        m.us[idx] = acc;
    }
}

// integrand I = sum_k u_k gcoef_k - u dNt - source N  (TFModel.py:653-657), times integW_q (:660)
template <class C>
__device__ __forceinline__ float integrand(const TileArgs& A, const SmemMap<C>& m, int p, unsigned int gp) {
    float I = 0.f;
    const size_t row = table_row(A, gp);
#pragma unroll
    for (int k = 0; k < C::S - 1; ++k)
        I = fmaf(m.us[(1 + k) * C::TP + p], __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + row), I);
    if (A.timeDependent) I -= m.us[p] * __ldg(A.cols + (size_t)A.colT * A.pstride + row);
    if (A.isSource) I -= __ldg(A.cols + (size_t)A.colS * A.pstride + row);
    if (A.integW) I *= __ldg(A.integW + (gp % A.integNum));
    return I;
}

// ------------------------------------------------------------------ forward-only kernel
template <class C, int MODE>
__global__ void __launch_bounds__(C::NT, 1) vn_fwd_kernel(const __grid_constant__ TileArgs A) {
    constexpr int TP = C::TP, NT = C::NT, BUF = C::BUF;
    extern __shared__ __align__(16) float smem[];
    const NetDesc& net = A.net;
    const int L = net.L;
    const SmemMap<C> m(smem, L, 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = (lane & 7) + 8 * (warp % C::NPGW);
    const int ng = (lane >> 3) + 4 * (warp / C::NPGW);
    const int p0 = 4 * pg;
    constexpr bool RES = (MODE == MODE_RESIDUAL);
    // residual: unit tangent rows for the dim spatial inputs and, if time dependent, the time input (column dim)
    stage_network<C>(A, m, smem, (int)tile_smem_floats<C>(L, false), RES ? A.dim + (A.timeDependent ? 1 : 0) : C::S - 1);

    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
        const unsigned int base = (unsigned int)(A.tile0 + tile) * TP;      // tile0: forward launched per uploaded chunk (vn_loss_grad_fed, two-pass class)
        load_inputs<C>(A, m, base);
        __syncthreads();
        for (int l = 0; l < L; ++l) {
            forward_layer<C, RES>(net, m, l, m.A + ((l - 1) & 1) * BUF, m.A + (l & 1) * BUF, nullptr, p0, ng);
            __syncthreads();
        }
        output_layer<C>(m, m.A + ((L - 1) & 1) * BUF, net.wpad[L - 1]);
        __syncthreads();
        for (int p = tid; p < TP; p += NT) {
            const unsigned int gp = base + p;
            if (gp >= A.P) continue;
            if (MODE == MODE_VAR_FWD) {
                A.Iw[gp] = integrand<C>(A, m, p, gp);
            } else if (MODE == MODE_RESIDUAL) {
                // res = -u_t + kappa * sum_k d2u/dx_k^2 - sum_k (vel_k - dkappa/dx_k) du/dx_k + s   (TFModel.py:750-754)
                float res = A.timeDependent ? -m.us[(1 + A.dim) * TP + p] : 0.f;
                float lap = 0.f, adv = 0.f;
                for (int k = 0; k < A.dim; ++k) {
                    lap += m.us[(4 + k) * TP + p];
                    const float vd = __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + gp) -
                                     __ldg(A.cols + (size_t)(A.colDD + k) * A.pstride + gp);
                    adv = fmaf(vd, m.us[(1 + k) * TP + p], adv);
                }
                res = fmaf(__ldg(A.cols + (size_t)A.colD * A.pstride + gp), lap, res) - adv;
                res += __ldg(A.cols + (size_t)A.colS * A.pstride + gp);
                A.Iw[gp] = res;
                A.uout[gp] = m.us[p];
            } else if (MODE == MODE_EVAL) {
                A.uout[gp] = m.us[p];
            } else {
                const float r = m.us[p] - __ldg(A.label + gp);
                A.cj[gp] = A.biDimVal * r * r;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ forward + adjoint kernel
template <class C, int MODE>
__global__ void __launch_bounds__(C::NT, 1) vn_adj_kernel(const __grid_constant__ TileArgs A) {
    constexpr bool BIC = (MODE == MODE_BIC_ADJ);
    constexpr bool FUSED = (MODE == MODE_VAR_FUSED);
    constexpr int S = C::S, WP = C::WP, TP = C::TP, TN = C::TN, TPS = C::TPS, WS = C::WS, NT = C::NT;
    constexpr int KIN = C::KIN, BUF = C::BUF;
    extern __shared__ __align__(16) float smem[];
    const NetDesc& net = A.net;
    const int L = net.L;
    const SmemMap<C> m(smem, L, 3);
    float* us = m.us;                                        // u_s[p] (forward) then the adjoint seeds
    float* Ish = m.coef;                                     // integrand per point (FUSED)
    float* Rsh = m.coef + TP;                                // R per test function of the tile (FUSED)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = (lane & 7) + 8 * (warp % C::NPGW);
    const int ng = (lane >> 3) + 4 * (warp / C::NPGW);
    const int p0 = 4 * pg;
    // weight-gradient GEMM mapping: 128 owners x KS point slices
    const int own = tid & 127, kslice = tid >> 7;
    const int ig = own & 7, jg = ((own & 31) >> 3) + 4 * (own >> 5);

    stage_network<C>(A, m, smem, (int)tile_smem_floats<C>(L, true));

    const PartLayout pl = make_part_layout<C>(L);
    double* part64 = A.part + (size_t)blockIdx.x * pl.psz;
    float* part = A.part32 + (size_t)blockIdx.x * pl.psz;   // FP32 window accumulators
    float* stash = A.stash + (size_t)blockIdx.x * A.stashFloats;
    double lossAcc = 0.0;                                    // lane 0 of each warp (FUSED)
    bool first = true;                                       // first tile of the current FP32 window
    bool firstFold = !A.accumulate;
    int win = 0;

    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
        const unsigned int base = (unsigned int)(A.tile0 + tile) * TP;
        load_inputs<C>(A, m, base);
        if (MODE == MODE_VAR_ADJ) {
            // seeds from the globally reduced R_i: lambda = 2 w2 detJ_i w_q R_i ; ubar = -lambda dNt ; ubar_k = lambda gcoef_k
            for (int p = tid; p < TP; p += NT) {
                const unsigned int gp = base + p;
                float lam = 0.f;
                size_t row = 0;
                if (gp < A.P) {
                    const unsigned int i = gp / A.integNum, q = gp - i * A.integNum;
                    const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                    const float wq = A.integW ? __ldg(A.integW + q) : 1.f;
                    lam = 2.f * __ldg(A.wts + 2) * dj * wq * __ldg(A.R + i);
                    row = table_row(A, gp);
                }
                us[p] = A.timeDependent ? -lam * __ldg(A.cols + (size_t)A.colT * A.pstride + row) : 0.f;
#pragma unroll
                for (int k = 0; k < S - 1; ++k)
                    us[(1 + k) * TP + p] = lam * __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + row);
            }
        }
        __syncthreads();

        // ---- forward sweep: B_l -> operand buffer (l & 1); layers 0..L-3 also go to the L2 stash
        for (int l = 0; l < L; ++l) {
            forward_layer<C>(net, m, l, m.A + ((l - 1) & 1) * BUF, m.A + (l & 1) * BUF,
                             (l < L - 2) ? stash + (size_t)l * BUF : nullptr, p0, ng);
            __syncthreads();
        }
        const int cur = (L - 1) & 1;
        const float* Blast = m.A + cur * BUF;
        const int wlast = net.wpad[L - 1];

        if (FUSED || BIC) {
            output_layer<C>(m, Blast, wlast);
            __syncthreads();
        }
        if (FUSED) {
            // R_i = sum_q w_q I_iq over the test functions of this tile, lossVec_i = detJ_i R_i^2
            for (int p = tid; p < TP; p += NT) {
                const unsigned int gp = base + p;
                Ish[p] = gp < A.P ? integrand<C>(A, m, p, gp) : 0.f;
            }
            __syncthreads();
            const int nf = TP / (int)A.integNum;
            for (int f = warp; f < nf; f += C::NW) {
                float r = 0.f;
                for (int q = lane; q < (int)A.integNum; q += 32) r += Ish[f * A.integNum + q];
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
            __syncthreads();
            for (int p = tid; p < TP; p += NT) {
                const unsigned int gp = base + p;
                float lam = 0.f;
                size_t row = 0;
                if (gp < A.P) {
                    const unsigned int i = gp / A.integNum, q = gp - i * A.integNum;
                    const float dj = A.detJvec ? __ldg(A.detJ + table_tf(A, i)) : __ldg(A.detJ);
                    const float wq = A.integW ? __ldg(A.integW + q) : 1.f;
                    lam = 2.f * __ldg(A.wts + 2) * dj * wq * Rsh[p / A.integNum];
                    row = table_row(A, gp);
                }
                us[p] = A.timeDependent ? -lam * __ldg(A.cols + (size_t)A.colT * A.pstride + row) : 0.f;
#pragma unroll
                for (int k = 0; k < S - 1; ++k)
                    us[(1 + k) * TP + p] = lam * __ldg(A.cols + (size_t)(A.colG + k) * A.pstride + row);
            }
            __syncthreads();
        }
        if (BIC) {
            for (int p = tid; p < TP; p += NT) {
                const unsigned int gp = base + p;
                float seed = 0.f;
                if (gp < A.P) {
                    const float r = us[p] - __ldg(A.label + gp);
                    A.cj[gp] = A.biDimVal * r * r;
                    // mean over boundary rows / initial rows (TFModel.py:644-648)
                    float sc;
                    if (gp < A.bDof) sc = __ldg(A.wts + 0) / (float)A.bDof;
                    else sc = A.timeDependent ? __ldg(A.wts + 1) / (float)(A.P - A.bDof) : 0.f;
                    seed = 2.f * A.biDimVal * r * sc;
                }
                us[p] = seed;
            }
            __syncthreads();
        }

        // ---- top of the adjoint: zbar_{L-1} from ubar (outer product with w_out) -> buffer 2
        float* X = m.A + 2 * BUF;                            // D_l
        float* Y = m.A + (cur ^ 1) * BUF;                    // B_{l-1}: B_{L-2} is still resident from the forward sweep
        float* Z = m.A + cur * BUF;                          // D_{l-1} (overwrites B_{L-1} once it is consumed)
#pragma unroll
        for (int t = 0; t < TN; ++t) {
            const int i = ng + C::NNG * t;
            if (i < wlast) {
                const float wv = m.wout[i];
                const float4 a4 = lds4(Blast + i * TPS + p0);
                const float4 u0 = lds4(us + p0);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float ub[4] = {u0.x * wv, u0.y * wv, u0.z * wv, u0.w * wv};
                float zb[4], cross[4] = {0.f, 0.f, 0.f, 0.f}, d1[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) d1[p] = act_d1<C::ACT>(a[p]);
#pragma unroll
                for (int s = 1; s < S; ++s) {
                    const float4 us4 = lds4(us + s * TP + p0);
                    const float4 da4 = lds4(Blast + (s * WP + i) * TPS + p0);
                    const float dab[4] = {us4.x * wv, us4.y * wv, us4.z * wv, us4.w * wv};
                    const float da[4] = {da4.x, da4.y, da4.z, da4.w};
                    float o[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) { cross[p] = fmaf(dab[p], da[p], cross[p]); o[p] = dab[p] * d1[p]; }
                    sts4(X + (s * WP + i) * TPS + p0, make_float4(o[0], o[1], o[2], o[3]));
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) zb[p] = fmaf(ub[p], d1[p], act_d2r<C::ACT>(a[p]) * cross[p]);
                sts4(X + i * TPS + p0, make_float4(zb[0], zb[1], zb[2], zb[3]));
            }
        }
        {
            // g(w_out)[i] = sum_{s,p} B_last[s][i][p] * ubar_s[p]   (thread: neuron i, point slice)
            constexpr int NPART = NT / WP, PSL = TP / NPART;
            const int i = tid % WP, prt = tid / WP;
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s)
                for (int p = prt * PSL; p < (prt + 1) * PSL; p += 4) {
                    const float4 b4 = lds4(Blast + (s * WP + i) * TPS + p);
                    const float4 u4 = lds4(us + s * TP + p);
                    acc = fmaf(b4.x, u4.x, acc); acc = fmaf(b4.y, u4.y, acc);
                    acc = fmaf(b4.z, u4.z, acc); acc = fmaf(b4.w, u4.w, acc);
                }
            float* pw = part + pl.off_wout + tid;
            if (first) __stcg(pw, acc); else atomicAdd(pw, acc);
            if (tid == 0) {
                float sb = 0.f;
                for (int p = 0; p < TP; ++p) sb += us[p];
                float* pb = part + pl.off_bout;
                if (first) __stcg(pb, sb); else atomicAdd(pb, sb);
            }
        }
        __syncthreads();

        // ---- backward sweep.  Roles: X = zbar_l, Y = activations of layer l-1, Z = zbar_{l-1}.
        for (int l = L - 1; l >= 1; --l) {
            if (l < L - 1) {
                // bring B_{l-1} back from the L2 stash while the abar GEMM below runs
                const float* src = stash + (size_t)(l - 1) * BUF;
                const int rows = net.wpad[l - 1], chunks = rows * (TPS / 4);
                for (int s = 0; s < S; ++s)
                    for (int c = tid; c < chunks; c += NT)
                        cp_async16(Y + s * WP * TPS + 4 * c, src + s * WP * TPS + 4 * c);
                cp_async_commit();
            }
            float acc[S][4][TN];
            const bool active = ng < net.wpad[l - 1];
            if (active) adj_gemm<C>(X, m.Wl + (l - 1) * WP * WS, net.wpad[l], p0, ng, acc);
            if (l < L - 1) {
                cp_async_wait_all();
                __syncthreads();
            }
            // through act' / act'' of layer l-1: zbar_{l-1}, dzbar_{l-1} -> Z
            if (active) {
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    const int i = ng + C::NNG * t;
                    if (i < net.wpad[l - 1]) {
                        const float4 a4 = lds4(Y + i * TPS + p0);
                        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                        float d1[4], cross[4] = {0.f, 0.f, 0.f, 0.f}, zb[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) d1[p] = act_d1<C::ACT>(a[p]);
#pragma unroll
                        for (int s = 1; s < S; ++s) {
                            const float4 da4 = lds4(Y + (s * WP + i) * TPS + p0);
                            const float da[4] = {da4.x, da4.y, da4.z, da4.w};
                            float o[4];
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                cross[p] = fmaf(acc[s][p][t], da[p], cross[p]);
                                o[p] = acc[s][p][t] * d1[p];
                            }
                            sts4(Z + (s * WP + i) * TPS + p0, make_float4(o[0], o[1], o[2], o[3]));
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            zb[p] = fmaf(acc[0][p][t], d1[p], act_d2r<C::ACT>(a[p]) * cross[p]);
                        sts4(Z + i * TPS + p0, make_float4(zb[0], zb[1], zb[2], zb[3]));
                    }
                }
            }
            // gW_l, gb_l from (B_{l-1}, zbar_l)
            gw_gemm<C, WP, C::TI>(Y, X, ig, jg, kslice, part + pl.off_gw[l] + 2 * tid,
                                  part + pl.off_gb[l] + kslice * WP + jg * C::TJ, first);
            __syncthreads();
            float* t = X; X = Z; Z = Y; Y = t;               // zbar_{l-1} becomes the operand; the other two are free
        }
        // layer 0: gW_0, gb_0 from the inputs (layer "-1": X rows + unit tangent rows)
        gw_gemm<C, KIN, 1>(m.Bm1, X, ig, jg, kslice, part + pl.off_gw[0] + 2 * tid,
                           part + pl.off_gb[0] + kslice * WP + jg * C::TJ, first);
        __syncthreads();
        first = false;
        if (++win == VN_FOLD || tile + (int)gridDim.x >= A.ntiles) {
            fold_slab<C>(pl, L, part, part64, tid, ig, jg, kslice, firstFold);
            firstFold = false; first = true; win = 0;
        }
    }
    if (FUSED && lane == 0) {
        double* lp = A.lossPart + blockIdx.x * C::NW + warp;
        *lp = A.accumulate ? *lp + lossAcc : lossAcc;
    }
}

// ------------------------------------------------------------------ small kernels
// R_i = sum_q Iw[i*integNum+q]; lossVec_i = detJ_i R_i^2; block partial of sum_i (detJ_i) R_i^2 in FP64.
struct SegArgs {
    const float* Iw; unsigned int nb, integNum; const float* detJ; int detJvec;
    const int* tfIndex;             // batch slot -> table test function (per-test-function detJ), or nullptr
    float* R; float* lossVec; double* blockSum;
};
__global__ void vn_segreduce_kernel(SegArgs A);

struct FinalArgs {
    NetDesc net; PartLayout pl;
    const double* partVar; int nVar;        // per-CTA slabs of the variational adjoint kernel
    const double* partBic; int nBic;        // ... of the boundary/initial adjoint kernel
    const double* flat;                     // [nparam] already reduced variational gradient (vn_tc64 kernel), or nullptr
    const double* slab; const int* slabSlot; int slabStride, nSlab;   // or: per-CTA slabs [nSlab][slabStride] with parameter idx at slabSlot[idx] (vn_tpp kernel)
    const int* slotParam;                   // slab mode, [nSlab slots]: parameter of a slot or -1 -> one block per 32-slot patch (coalesced slab reads)
    const double* segSum; int nSeg;         // partial sums of (detJ_i) R_i^2
    const float* detJ; int detJvec;
    const float* cj; unsigned int nbi, bDof; int timeDependent;
    const float* wts;
    float* gbuf;                            // [nparam | loss, BCloss, ICloss, varLoss]
    int needGrad;
    // optimizer fused into the reduction (single-GPU vn_train_step): 0 = none, 1 = Adam, 2 = RMSProp
    int fuseOpt; float lr;
    float* theta; float* m; float* v;
    long long* step; double* corr;          // device step counter and Adam bias-correction factor (advanced by the last block)
    unsigned int* ticket;                   // zero-initialised block counter
    const int* err;                         // tensor-core classes: non-zero = an mbarrier wait expired in this step (or nullptr)
    float* lossRing; int ringSize;          // fused optimizer: loss of step t is also left in lossRing[t % ringSize] (vn_train_steps)
};
__global__ void vn_finalize_kernel(FinalArgs A);
