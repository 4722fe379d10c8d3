// vn_tc.h — tensor-core kernel class (hidden width 65..256): tcgen05 3xTF32 layer GEMMs.
//
// The resident-tile FMA classes (vn_tile.cuh) keep the weights and three operand buffers of a point tile in
// shared memory; at width 128/256 neither fits (one 256x256 layer is 256 KB).  At those widths the hidden
// layers are real dense contractions (TFModel.py:208-242 Dense layers; 3.5 MFLOP per quadrature point for
// 4x256), so this class runs them on the 5th-generation tensor cores and streams the activations of a chunk
// of points through global memory (L2/HBM: ~150 flop/B, far above the ridge) in one quad-major layout
// [neuron/4][point][4].  One chunk = 8 waves of 128-point tiles; per chunk, layer and stream the step is
//   forward   Z_s = A_{l-1,s} W_l             tc_gemm<FWD_VALUE | FWD_TANGENT>   (epilogue: bias, act, act' * tangent)
//   adjoint   abar_s = D_{l,s} W_l^T          tc_gemm<ADJ_TANGENT | ADJ_VALUE>   (epilogue: through act'/act'' of layer l-1)
//   gradient  gW_l = sum_{s,p} A_{l-1,s}^T D_{l,s}   tc_gw   (split-K over points, FP64 atomics; also g(b_l))
// with FP32 accuracy from the 3xTF32 split x = hi + lo (lo*hi + hi*lo + hi*hi) and short accumulation chains
// (the tensor core accumulates with truncation: several TMEM accumulator sets, summed in FP32 by the epilogue).
// Layer 0 (K = inpDim <= 8), the output layer, the integrand / per-test-function residual and the layer-0
// gradients are plain FP32 kernels.  Math: SURVEY.md App. A (TFModel.py:515-714).
#pragma once
#include <functional>
#include "vn_tile.cuh"

enum { TC_VAR = 0, TC_BIC = 1, TC_EVAL = 2 };

struct TcGeom {
    int WP;                 // padded hidden width (128 or 256)
    unsigned int capPts;    // chunk capacity in points (multiple of 128)
    size_t workBytes;       // workspace for one chunk (quad-major activations of every layer, adjoint ping-pong, staged weights)
    size_t smemGemm, smemGw;
};

// false when the network is outside this class (width > 256)
bool vn_tc_geometry(const NetDesc& net, int S, int numSMs, TcGeom* g);
cudaError_t vn_tc_prepare(int S, int act);          // opt-in shared memory attributes of the kernels used

struct TcJob {
    int S, act, mode;
    TcGeom geom;
    TileArgs in;            // rows to process: table columns, P, integNum, tfIndex, extraX, labels, outputs (R, lossVec, Iw, cj, uout)
    bool needGrad;
    void* work;             // geom.workBytes
    double* g64;            // [nparam] FP64 gradient accumulator, reference variable order (caller zeroes it once per step)
    double* lossAcc;        // sum_i (detJ_i) R_i^2 (TC_VAR; caller zeroes it)
    int* err;               // device flag: set to 1 when a bounded mbarrier wait expired
    cudaStream_t st;
    int numSMs;
    long long launches;     // out: kernels launched
    // optional (fed steps): called before a chunk is processed with the number of leading table rows it needs; makes the stream
    // wait for the uploads that cover them (vn_loss_grad_fed on the tensor-core class)
    std::function<cudaError_t(unsigned long long)> waitRows;
};

// theta -> zero-padded quad-major copies of the hidden kernels (both orientations) inside the workspace
cudaError_t vn_tc_stage_weights(const NetDesc& net, const TcGeom& g, const float* theta, void* work, cudaStream_t st);
cudaError_t vn_tc_run(TcJob& j);
// strong-form residual of the evaluation rows described by A (cols: X | kappa | vel | grad kappa | source; outputs A.uout, A.Iw)
cudaError_t vn_tc_residual(const TileArgs& A, int act, const TcGeom& g, void* work, cudaStream_t st, long long* launches);
cudaError_t vn_tc_grad_out(const double* g64, float* gbuf, int n, cudaStream_t st);
