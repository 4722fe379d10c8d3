// vn_tpp.h — thread-per-point FP32 kernel for NARROW networks (every hidden width <= 32): the reference's own operator
// configurations ([20], [10, 20], [10, 20, 30]; Operator_1Dt.py:140, Operator_2Dt.py:139, Operator_1DtMOR.py:178).
//
// The FMA tile class (vn_tile.cuh) pads every layer of such a network to its 32-wide register tiling and runs one CTA of 256
// threads per SM through a dozen CTA-wide barriers per tile; at [10, 20] it executes 3.2x the algorithmic FMAs at 32 % FMA-pipe
// utilisation.  Here one thread owns one quadrature point: the forward sweep (value + `dim` forward tangents), the integrand,
// the adjoint sweep and z-bar live in that thread's column of a [row][128 points] shared-memory array, looped over the ACTUAL
// widths in blocks of 8 neurons (weights are warp-uniform 128-bit broadcasts).  Only the weight gradients need the other
// threads' points: per layer, warps take 8x4 patches of [a_{l-1}; 1]^T [zbar_l] and contract them over the tile's 128 points
// (two points per lane and step: the 64-bit loads are the packed operands of fma.rn.f32x2), a transposing warp butterfly leaves one patch entry per lane, and those are
// accumulated in FP64 in the CTA's own global slab (fire-and-forget reductions, one writer per slot).  z-bar_l overwrites a_l in place once the patches that need
// a_l are done, so a CTA needs (sum_l S w_l + inpDim + S + 2) * 512 B: 3-5 CTAs per SM instead of one.
// Only the variational term (MODE_VAR_FUSED semantics, integNum | 128) runs here; boundary/initial rows and loss-only passes stay
// on the FMA tile class.
#pragma once
#include "vn_tile.cuh"

struct TppLayout {
    int L, S, inpDim;
    int w[VN_MAX_LAYERS];          // hidden widths
    int wp[VN_MAX_LAYERS];         // rounded up to a multiple of 8 (row stride of the weight images)
    int rowA[VN_MAX_LAYERS];       // first shared-memory row of layer l (stream s at + s * w[l])
    int rowX, rowU, rowOne, rowZero, nrows;
    int offW[VN_MAX_LAYERS + 1];   // W_l as [in][wp[l]] (l == L: w_out[wp[L-1]])
    int offWT[VN_MAX_LAYERS];      // W_l^T as [out][wpin] (l >= 1), wpin = wp[l-1]
    int offB[VN_MAX_LAYERS + 1];   // biases (l == L: b_out)
    int wfloats;                   // floats of the weight region (multiple of 4)
    int patch0[VN_MAX_LAYERS + 2]; // first 8x4 patch of gradient block l = 0..L ([a_{l-1}; 1]^T zbar_l; block L: the output layer)
    int ncb[VN_MAX_LAYERS + 1];    // column blocks of block l
    int nrb[VN_MAX_LAYERS + 1];    // row blocks of block l (rows: a_{l-1} then the bias row)
    int tabA[VN_MAX_LAYERS + 1], tabZ[VN_MAX_LAYERS + 1], ntab;   // operand-row offset tables (ints, in shared memory)
    int npatch;
    size_t smemBytes;
};

bool vn_tpp_supported(const NetDesc& net, int S);
void vn_tpp_layout(const NetDesc& net, int S, TppLayout* lay);
cudaError_t vn_tpp_prepare(int S, int act, size_t smemBytes, int* ctasPerSM);
// a: as for vn_adj_kernel<MODE_VAR_FUSED>; a.part = [grid][npatch * 32] FP64 patch slabs, a.lossPart = [grid][4]
cudaError_t vn_tpp_launch(int S, int act, const TileArgs& a, const TppLayout& lay, int grid, cudaStream_t st);
// slots[nparam]: position of every flat parameter (reference variable order) in a CTA's patch slab; vn_finalize_kernel sums
// the slabs of all CTAs in fixed order (bitwise reproducible)
void vn_tpp_param_slots(const NetDesc& net, const TppLayout& lay, int* slots);
