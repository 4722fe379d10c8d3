// vn_tc64.h — resident-tile tensor-core kernel for hidden widths 33..64 (the headline 4x64 network).
//
// The FMA tile class (vn_tile.cuh) needs ~113 k cycles per 64 points at width 64: its operand delivery from
// shared memory, not the FMA pipe, is the limiter.  This class keeps one 128-point tile per CTA (TMEM lane =
// quadrature point) and runs every hidden-layer contraction on the 5th-generation tensor cores as 3xTF32:
//   forward   Z_s = A_{l-1,s} W_l         (TFModel.py:208-242 Dense layers, value + `dim` forward tangents)
//   adjoint   abar_s = zbar_{l,s} W_l^T    
//   gradient  gW_l = sum_{s,p} a_{l-1,s}^T zbar_{l,s}
// The activation operand never leaves the SM: the epilogue threads (thread = point x 32 neurons) read the FP32
// accumulators from tensor memory, apply bias/activation/derivative factors and write the hi/lo TF32 split of the
// next operand straight back into tensor memory (A-from-TMEM MMAs); the weights come as pre-split canonical
// shared-memory images (one 32 KB cp.async per layer and direction, L2 resident); the weight-gradient GEMM takes
// point-contiguous operands that the same threads transpose into shared memory with conflict-free scalar stores.
// Layer 0 (K = inpDim), the output layer, the integrand / per-test-function residual, the adjoint seeds and all
// bias / layer-0 gradients are FP32 in the same kernel (warp transpose-reductions).  Only the variational term
// (MODE_VAR_FUSED semantics, integNum | 128) runs here; boundary/initial rows stay on the FMA class.
#pragma once
#include "vn_tile.cuh"

struct Tc64Geom {
    int psz;                // floats (FP32 window) / doubles (FP64) of one CTA's gradient slab
    size_t smemBytes;
    long long stashFloats;  // per-CTA activation stash (all hidden layers, all streams)
    int nImages;            // staged weight images (2 per hidden-to-hidden layer), 8192 floats each
    int lossSlots;          // loss partials written per CTA (A.lossPart)
};

bool vn_tc64_supported(const NetDesc& net, int S);
void vn_tc64_geometry(const NetDesc& net, int S, Tc64Geom* g);
cudaError_t vn_tc64_prepare(int S, int act, size_t smemBytes);
// theta -> hi/lo canonical K-major images [W_hi | W_lo] (forward) and [W^T_hi | W^T_lo] (adjoint) per layer
cudaError_t vn_tc64_stage_weights(const NetDesc& net, const float* theta, float* wimg, cudaStream_t st);
// a: as for vn_adj_kernel<MODE_VAR_FUSED> (part/part32/psz/stash/stashFloats/lossPart sized from Tc64Geom)
// fwdOnly: loss-only pass (R, lossVec, loss partials; no stash, no gradients) — available in the default schedule only
cudaError_t vn_tc64_launch(int S, int act, const TileArgs& a, const float* wimg, int* err, int grid, size_t smemBytes,
                           cudaStream_t st, int fwdOnly = 0);
bool vn_tc64_forward_only_available();
// fixed-order sum of the per-CTA FP64 slabs -> flat[nparam] in reference variable order
cudaError_t vn_tc64_reduce(const NetDesc& net, const double* slab64, int psz, int nCta, double* flat, cudaStream_t st);
// debug: phase cycle counters of CTA 0 / thread 0 of the last v2 launch (only filled when VARNET_B200_TC64_TIMING is set)
int vn_tc64_read_timing(long long out[16]);
