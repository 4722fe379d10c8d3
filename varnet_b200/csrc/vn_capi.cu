// vn_capi.cu — C ABI (include/varnet_b200.h) + the small reduction / optimizer / packing kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "../../include/varnet_b200.h"
#include "vn_dispatch.h"
#include "vn_tc.h"
#include "vn_tc64.h"
#include "vn_tpp.h"
#include "vn_pdl.cuh"
#include <utility>

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(VN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* vn_last_error(void) { return g_err; }

// ------------------------------------------------------------------ NCCL, bound at run time
// The engine does not link NCCL: the few entry points it needs are resolved with dlopen/dlsym from the library that is
// already in the process (torch's bundled libnccl.so.2, RTLD_NOLOAD first) or from a path given by the caller, so one
// build serves single-GPU hosts without NCCL and 8-GPU boxes alike.  Types/constants as in nccl.h (2.x ABI).
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;              // ncclUniqueId
    typedef void* Comm;                                           // ncclComm_t
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;
    bool ok() const { return GetUniqueId && CommInitRank && AllReduce && CommDestroy; }
};
static NcclApi g_nccl;
static int nccl_load(const char* path) {
    if (g_nccl.ok()) return VN_OK;
    void* h = nullptr;
    if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already loaded by the host framework
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(VN_E_STATE, "NCCL library not found: %s", dlerror());
    g_nccl.handle = h;
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(NcclApi::UniqueId*)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(NcclApi::Comm*, int, NcclApi::UniqueId, int)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclApi::Comm, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(NcclApi::Comm)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.ok()) return fail(VN_E_STATE, "NCCL library lacks the expected entry points");
    return VN_OK;
}
static const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"; }
enum { kNcclFloat32 = 7, kNcclSum = 0 };

// ------------------------------------------------------------------ dispatch over kernel classes
bool vn_tile_geometry(int S, int cls, int act, int mode, int L, TileGeom* g) {
    if (cls == 16) return vn_geom_c16(S, act, mode, L, g);
    if (cls == 32) return vn_geom_c32(S, act, mode, L, g);
    if (cls == 64) return vn_geom_c64(S, act, mode, L, g);
    if (cls == 164) return vn_geom_c164(S, act, mode, L, g);
    return false;
}
cudaError_t vn_tile_launch(int S, int cls, int act, int mode, const TileArgs& a, int grid, size_t smem,
                           cudaStream_t st) {
    if (cls == 16) return vn_launch_c16(S, act, mode, a, grid, smem, st);
    if (cls == 32) return vn_launch_c32(S, act, mode, a, grid, smem, st);
    if (cls == 64) return vn_launch_c64(S, act, mode, a, grid, smem, st);
    if (cls == 164) return vn_launch_c164(S, act, mode, a, grid, smem, st);
    return cudaErrorInvalidValue;
}
cudaError_t vn_tile_prepare(int S, int cls, int act, int mode, size_t smem) {
    if (cls == 16) return vn_prepare_c16(S, act, mode, smem);
    if (cls == 32) return vn_prepare_c32(S, act, mode, smem);
    if (cls == 64) return vn_prepare_c64(S, act, mode, smem);
    if (cls == 164) return vn_prepare_c164(S, act, mode, smem);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ small kernels
// One thread per test function: R_i = sum_q Iw[i,q] (TFModel.py:659-661), lossVec_i = detJ_i R_i^2
// (:668); FP64 block partials of sum_i detJ_i R_i^2 (detJvec) or sum_i R_i^2 (scalar detJ, :662-664).
__global__ void vn_segreduce_kernel(SegArgs A) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    double term = 0.0;
    if (i < A.nb) {
        const float* row = A.Iw + (size_t)i * A.integNum;
        float r = 0.f;
        if ((A.integNum & 3) == 0) {
            for (unsigned int q = 0; q < A.integNum; q += 4) {
                float4 v = __ldg(reinterpret_cast<const float4*>(row + q));
                r += v.x; r += v.y; r += v.z; r += v.w;
            }
        } else {
            for (unsigned int q = 0; q < A.integNum; ++q) r += __ldg(row + q);
        }
        const float r2 = r * r;
        const float dj = A.detJvec ? __ldg(A.detJ + (A.tfIndex ? (unsigned int)__ldg(A.tfIndex + i) : i)) : __ldg(A.detJ);
        A.R[i] = r;
        A.lossVec[i] = dj * r2;
        term = A.detJvec ? (double)dj * (double)r2 : (double)r2;
    }
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) A.blockSum[blockIdx.x] = v;
    }
}

__device__ static double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;   // valid in thread 0
}

// TF-1.x Adam (adam.py _apply_dense): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).
// `corr` = sqrt(1-b2^t)/(1-b1^t) for the step being applied, maintained by vn_advance_kernel.
__device__ __forceinline__ void adam_update(float* theta, float* m, float* v, int i, float gi, float lr, const double* corr) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float lr_t = lr * (float)corr[0];
    theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
__global__ void vn_adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                               const float* __restrict__ g, int n, float lr, const double* __restrict__ corr,
                               const int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (err && *err)) return;         // poisoned step (expired mbarrier wait): keep the weights
    adam_update(theta, m, v, i, g[i], lr, corr);
}
// TF-1.x RMSProp (decay .9, momentum 0, eps 1e-10, ms initialised to ones): ms = .9 ms + .1 g^2;
// mom = lr*g/sqrt(ms+eps); theta -= mom.
__device__ __forceinline__ void rmsprop_update(float* theta, float* ms, int i, float gi, float lr) {
    const float s = 0.9f * ms[i] + 0.1f * gi * gi;
    ms[i] = s;
    theta[i] -= lr * gi / sqrtf(s + 1e-10f);
}
__global__ void vn_rmsprop_kernel(float* __restrict__ theta, float* __restrict__ ms, const float* __restrict__ g,
                                  int n, float lr, const int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (err && *err)) return;
    rmsprop_update(theta, ms, i, g[i], lr);
}
__device__ __forceinline__ void advance_step(long long* step, double* corr) {
    const long long t = step[0] + 1;       // step just applied
    step[0] = t;
    const double tn = (double)(t + 1);     // bias correction of the NEXT step
    corr[0] = sqrt(1.0 - pow(0.999, tn)) / (1.0 - pow(0.9, tn));
}
// the last k entries of the loss ring in step order (the step counter already counts them)
__global__ void vn_ring_gather_kernel(const float* __restrict__ ring, int ringSize, const long long* __restrict__ step, int k, float* __restrict__ out) {
    const long long s0 = step[0] - k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = ring[(unsigned long long)(s0 + i) % (unsigned long long)ringSize];
}
__global__ void vn_advance_kernel(long long* step, double* corr, const int* err, float* lossRing, int ringSize, const float* lossPtr) {
    if (lossRing) lossRing[(unsigned long long)step[0] % (unsigned long long)ringSize] = *lossPtr;       // loss of the step being applied
    if (!(err && *err)) advance_step(step, corr);
}

// Blocks [0, gridDim.x-1): one WARP per flat parameter sums the per-CTA FP64 slabs: lane c takes slabs
// c, c+32, ... in order and the lanes are combined by a fixed xor tree (deterministic replacement for
// TF's gradient accumulation; a serial per-thread walk over ~300 slabs was latency-bound).
// Last block: the loss scalars loss = w0*bCs + w1*iCs + w2*varLoss (TFModel.py:643-666).
// With fuseOpt the optimizer update is applied by the same warps and the last block to finish advances the step.
__global__ void vn_finalize_kernel(FinalArgs A) {
    const NetDesc& net = A.net;
    const PartLayout& pl = A.pl;
    // k-steps-in-one-graph path: the next step's kernels may become resident now (they block in their own pdl_wait until this
    // grid has completed); everything below reads what the step's point kernels wrote
    pdl_launch_dependents();
    pdl_wait();
    // a tensor-core kernel of this step gave up on an mbarrier wait: its partial slabs are stale.  Publish NaN and leave
    // the weights alone, so the caller sees the failure in the very step it happened (the host also reads the flag).
    const bool bad = A.err != nullptr && *reinterpret_cast<const volatile int*>(A.err) != 0;
    __shared__ double sh[32];
    __shared__ double red[32][32];
    if (blockIdx.x + 1 < gridDim.x) {
        const int lane = threadIdx.x & 31;
        const bool patchMode = A.slotParam != nullptr;        // thread-per-point class: block = one 32-slot patch of the per-CTA slabs
        const int wrp = threadIdx.x >> 5, nwrp = blockDim.x >> 5;
        const int pslot = blockIdx.x * 32 + lane;
        const int idx = patchMode ? __ldg(A.slotParam + pslot) : blockIdx.x * nwrp + wrp;
        double s = 0.0;
        // patch mode: warp w sums slabs w, w + nwrp, ... of the block's patch, one slot per lane: 256 contiguous bytes per load
        // (one warp per parameter read 32 slabs per load, one 32-byte sector each)
        if (patchMode && A.needGrad)
            for (int c = wrp; c < A.nSlab; c += 8 * nwrp) {           // 8 loads in flight per lane, added in slab order
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = c + u * nwrp < A.nSlab ? __ldcg(A.slab + (size_t)(c + u * nwrp) * A.slabStride + pslot) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) s += t[u];
            }
        if (A.needGrad && idx >= 0 && idx < net.nparam) {
        // locate (layer, kind, i, j)
        int l = 0, isBias = 0, i = 0, j = 0;
        for (l = 0; l <= net.L; ++l) {
            const int wi = l == 0 ? net.inpDim : net.width[l - 1];
            const int wo = l == net.L ? 1 : net.width[l];
            if (idx >= net.woff[l] && idx < net.woff[l] + wi * wo) { i = (idx - net.woff[l]) / wo; j = (idx - net.woff[l]) - i * wo; break; }
            if (idx >= net.boff[l] && idx < net.boff[l] + wo) { isBias = 1; j = idx - net.boff[l]; break; }
        }
        int slot[16], nslot = 1;
        if (l == net.L) {
            if (isBias) slot[0] = pl.off_bout;
            else { nslot = pl.NT / pl.WP; for (int p = 0; p < nslot; ++p) slot[p] = pl.off_wout + i + pl.WP * p; }
        } else if (isBias) {
            nslot = pl.KS;
            for (int h = 0; h < pl.KS; ++h) slot[h] = pl.off_gb[l] + h * pl.WP + (j & 15) * pl.TJ + (j >> 4);
        } else {
            // owner thread of gW[i][j]: i = ig + 8t, j = jg + 16u (vn_tile.cuh gw_gemm); one slot per point slice
            const int ig = i & 7, t = i >> 3, jg = j & 15, u = j >> 4;
            const int own = (jg >> 2) * 32 + ig + 8 * (jg & 3);
            nslot = pl.KS;
            for (int h = 0; h < pl.KS; ++h) {
                const int tid = h * 128 + own, r = (l == 0 ? 0 : t * pl.TJ) + u;
                slot[h] = pl.off_gw[l] + (r >> 1) * (2 * pl.NT) + 2 * tid + (r & 1);   // pair-interleaved slab
            }
        }
        if (patchMode) {
            for (int c = wrp; c < A.nBic; c += nwrp)
                for (int p = 0; p < nslot; ++p) s += __ldcg(A.partBic + (size_t)c * pl.psz + slot[p]);
        } else {
        for (int c = lane; c < A.nVar; c += 32)
            for (int p = 0; p < nslot; ++p) s += __ldcg(A.partVar + (size_t)c * pl.psz + slot[p]);
        for (int c = lane; c < A.nBic; c += 32)
            for (int p = 0; p < nslot; ++p) s += __ldcg(A.partBic + (size_t)c * pl.psz + slot[p]);
        if (A.slab) {
            const int sl = __ldg(A.slabSlot + idx);
            for (int c = lane; c < A.nSlab; c += 32) s += __ldcg(A.slab + (size_t)c * A.slabStride + sl);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        }
        if (patchMode) {                                       // warps of the block combined in fixed order by warp 0
            red[wrp][lane] = s;
            __syncthreads();
            s = 0.0;
            if (wrp == 0)
                for (int k = 0; k < nwrp; ++k) s += red[k][lane];
        }
        if (A.needGrad && idx >= 0 && idx < net.nparam) {
        if (patchMode ? wrp == 0 : lane == 0) {
            if (A.flat) s += A.flat[idx];
            const float g = bad ? __int_as_float(0x7fc00000) : (float)s;
            A.gbuf[idx] = g;
            // single-GPU training step: apply_gradients fused into the reduction (TFModel.py:313)
            if (bad) { }
            else if (A.fuseOpt == 1) adam_update(A.theta, A.m, A.v, idx, g, A.lr, A.corr);
            else if (A.fuseOpt == 2) rmsprop_update(A.theta, A.v, idx, g, A.lr);
        }
        }
    } else {
    double v = 0.0;
    for (int k = threadIdx.x; k < A.nSeg; k += blockDim.x) v += __ldcg(A.segSum + k);
    const double segTot = block_sum(v, sh);
    double vb = 0.0, vi = 0.0;
    for (unsigned int k = threadIdx.x; k < A.nbi; k += blockDim.x) {
        const double c = (double)__ldcg(A.cj + k);
        if (k < A.bDof) vb += c; else vi += c;
    }
    const double sb = block_sum(vb, sh);
    const double si = block_sum(vi, sh);
    if (threadIdx.x == 0) {
        const float varLoss = A.detJvec ? (float)segTot : A.detJ[0] * (float)segTot;   // detJ moved outside the sum (:664)
        const float bCs = (float)(sb / (double)A.bDof);                                  // 0/0 -> nan like tf.reduce_mean
        const float iCs = A.timeDependent ? (float)(si / (double)(A.nbi - A.bDof)) : 0.f;
        const float loss = bad ? __int_as_float(0x7fc00000) : A.wts[0] * bCs + A.wts[1] * iCs + A.wts[2] * varLoss;
        float* o = A.gbuf + net.nparam;
        o[0] = loss; o[1] = bCs; o[2] = iCs; o[3] = varLoss;
        // loss history of back-to-back steps (vn_train_steps): the step counter is advanced only after every block of this
        // grid has taken its ticket, i.e. after this store
        if (A.fuseOpt && A.lossRing) A.lossRing[(unsigned long long)A.step[0] % (unsigned long long)A.ringSize] = loss;
    }
    }
    if (A.fuseOpt) {
        // the last block to get here advances the step counter / Adam bias correction: every parameter warp has read
        // `corr` before its block took a ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(A.ticket, 1u);
            if (t == gridDim.x - 1) {
                if (!bad) advance_step(A.step, A.corr);
                *A.ticket = 0u;
            }
        }
    }
}


__global__ void vn_fill_kernel(float* p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Row-major feed arrays -> SoA column table (float32, zero-padded to `pstride`).  Casting a float64
// feed with __double2float_rn reproduces the float32 placeholder rounding bit-for-bit.
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<double>(double v) { return __double2float_rn(v); }

template <typename T>
__global__ void vn_pack_kernel(const T* __restrict__ X, int inpDim, const T* __restrict__ G, int dim,
                               const T* __restrict__ dNt, const T* __restrict__ src, const T* __restrict__ N,
                               float* __restrict__ cols, long long pstride, long long off, long long n,
                               int colX, int colG, int colT, int colS) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long gp = off + r;
    for (int c = 0; c < inpDim; ++c) cols[(size_t)(colX + c) * pstride + gp] = to_f32<T>(X[r * inpDim + c]);
    if (G) for (int c = 0; c < dim; ++c) cols[(size_t)(colG + c) * pstride + gp] = to_f32<T>(G[r * dim + c]);
    if (dNt && colT >= 0) cols[(size_t)colT * pstride + gp] = to_f32<T>(dNt[r]);
    if (src && N && colS >= 0) cols[(size_t)colS * pstride + gp] = to_f32<T>(src[r]) * to_f32<T>(N[r]);   // float32 product, as tf.multiply(source, N) (:657)
}
// Point table of a uniform space-time mesh with constant coefficients, generated in place (vn_generate_table_f64).
// Same float64 operation order as the host code it replaces (VarNet.py:576-586,837; no FMA contraction), rounded to
// float32 like the feed cast, so the table is bit-identical to an uploaded one.
struct GenArgs {
    const double* coord; const double* tcoord; long long nTime;
    double h[VN_MAX_INPDIM];            // element sizes he[d], then ht
    const double* delta;                // [feDim][q]
    const double* N; const double* dN;  // [q], [q][feDim]
    double diff, vel[3], source;
    long long tf0, n; int q, dim, feDim;
    float* cols; long long pstride; int colX, colG, colT, colS;
};
__global__ void vn_generate_kernel(const GenArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const long long i = a.tf0 + r / a.q;
    const int qq = (int)(r % a.q);
    const long long s = i / a.nTime, j = i - s * a.nTime;
    for (int d = 0; d < a.dim; ++d)
        a.cols[(size_t)(a.colX + d) * a.pstride + r] =
            __double2float_rn(__dadd_rn(a.coord[s * a.dim + d], __dmul_rn(a.h[d], a.delta[(size_t)d * a.q + qq])));
    if (a.feDim > a.dim)
        a.cols[(size_t)(a.colX + a.dim) * a.pstride + r] =
            __double2float_rn(__dadd_rn(a.tcoord[j], __dmul_rn(a.h[a.dim], a.delta[(size_t)a.dim * a.q + qq])));
    const double Nq = a.N[qq];
    for (int k = 0; k < a.dim; ++k)
        a.cols[(size_t)(a.colG + k) * a.pstride + r] =
            __double2float_rn(__dadd_rn(__dmul_rn(a.diff, a.dN[(size_t)qq * a.feDim + k]), __dmul_rn(a.vel[k], Nq)));
    if (a.colT >= 0) a.cols[(size_t)a.colT * a.pstride + r] = __double2float_rn(a.dN[(size_t)qq * a.feDim + a.dim]);
    if (a.colS >= 0) a.cols[(size_t)a.colS * a.pstride + r] = __double2float_rn(a.source) * __double2float_rn(Nq);
}
// residual inputs: X | diff | vel[dim] | diff_dx[dim] | source  ->  SoA columns
template <typename T>
__global__ void vn_pack_res_kernel(const T* __restrict__ X, int inpDim, const T* __restrict__ diff,
                                   const T* __restrict__ vel, const T* __restrict__ ddx, const T* __restrict__ src,
                                   int dim, float* __restrict__ cols, long long pstride, long long n) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int c = 0;
    for (int k = 0; k < inpDim; ++k) cols[(size_t)(c++) * pstride + r] = to_f32<T>(X[r * inpDim + k]);
    cols[(size_t)(c++) * pstride + r] = to_f32<T>(diff[r]);
    for (int k = 0; k < dim; ++k) cols[(size_t)(c++) * pstride + r] = to_f32<T>(vel[r * dim + k]);
    for (int k = 0; k < dim; ++k) cols[(size_t)(c++) * pstride + r] = to_f32<T>(ddx[r * dim + k]);
    cols[(size_t)(c++) * pstride + r] = to_f32<T>(src[r]);
}
template <typename T>
__global__ void vn_cast_kernel(const T* __restrict__ in, float* __restrict__ out, long long n) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = to_f32<T>(in[r]);
}

// ------------------------------------------------------------------ host staging of pageable feeds
// cudaMemcpyAsync out of pageable memory is staged by the driver with one host thread (~11 GB/s measured for the 3 GB float64
// feed of a cfg-4 step).  The fed step instead copies each sub-chunk of the caller's arrays into one of two engine-owned pinned
// buffers with a few host threads and lets the DMA engine take it from there, so the host copy of sub-chunk k+1 overlaps the
// transfer of sub-chunk k.  A tiny persistent pool: run(n, f) executes f(0..n-1) on the workers and the caller.
class HostPool {
public:
    explicit HostPool(int nthreads) {
        for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 0) return;
        { std::lock_guard<std::mutex> lk(m_); f_ = &f; next_ = 0; total_ = n; done_ = 0; ++gen_; }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        cvDone_.wait(lk, [this] { return done_ == total_; });
        f_ = nullptr;
    }
    int size() const { return (int)workers_.size() + 1; }
private:
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            { std::lock_guard<std::mutex> lk(m_); if (!f_ || next_ >= total_) return; i = next_++; f = f_; }
            (*f)(i);
            { std::lock_guard<std::mutex> lk(m_); if (++done_ == total_) cvDone_.notify_all(); }
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> lk(m_); cv_.wait(lk, [&] { return stop_ || gen_ != seen; }); if (stop_) return; seen = gen_; }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, cvDone_;
    const std::function<void(int)>* f_ = nullptr;
    int next_ = 0, total_ = 0, done_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};
static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// ------------------------------------------------------------------ engine
enum { PK_VAR_FWD = 0, PK_SEG = 1, PK_VAR_ADJ = 2, PK_BIC = 3, PK_FINAL = 4, PK_OPT = 5 };
// Bumped whenever a live device buffer is freed and re-allocated: captured step graphs bake device pointers in, so a graph
// captured in an older epoch must not be replayed (vn_train_step compares PointSet::graphEpoch with it).
static long long g_reallocEpoch = 0;
struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) { cudaFree(p); ++g_reallocEpoch; }
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) { cudaFree(p); ++g_reallocEpoch; } p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// one uploaded point table (SoA columns + per-test-function detJ + Gauss weights) and its captured step graph
struct PointSet {
    DevBuf cols, integW, detJ;
    long long pstride = 0; unsigned int rows = 0, nbTab = 0, integNum = 0; int detJvec = 0, hasIntegW = 0;
    int colX = 0, colG = 0, colT = -1, colS = -1, ncols = 0, nx = 0;
    bool loaded = false;
    // in-kernel generation (vn_generate_table_f64 on an engine whose variational passes all run on the tensor-core tile kernel):
    // no materialised columns, only the mesh centres and the q-periodic tables
    bool inKernel = false; DevBuf genBuf; GenTab genTab{};
    cudaGraphExec_t graph = nullptr; float graphLr = -1.f; int graphLaunches = 0; unsigned int graphNb = 0; bool graphIndexed = false;
    long long graphEpoch = -1;   // g_reallocEpoch at capture time
    // k consecutive steps as ONE graph (vn_train_steps / vn_train_batches): no graph launch between the steps
    cudaGraphExec_t graphK = nullptr; int graphKn = 0; float graphKLr = -1.f; int graphKLaunches = 0; unsigned int graphKNb = 0;
    bool graphKIndexed = false, graphKSeq = false; long long graphKEpoch = -1;
};

struct vn_engine {
    vn_config cfg;
    NetDesc net;
    int S = 0, wclass = 0, numSMs = 0;
    // tensor-core class (wclass == 256): chunk workspace, FP64 gradient accumulator [nparam | loss sum], barrier-timeout flag
    TcGeom tcGeom{};
    DevBuf tcWork, tcAcc, tcErr;
    // width-64 tensor-core tile kernel (vn_tc64.h) for the variational term of the 64-wide FMA class
    bool tc64 = false;           // network and build allow it
    bool useTc64 = false;        // ... and the current batch does (integNum | 128)
    Tc64Geom tc64Geom{};
    DevBuf tc64Img, tc64Flat;
    // thread-per-point kernel (vn_tpp.h) for the variational term of narrow networks (every hidden width <= 32)
    bool tpp = false;            // network and build allow it (>= 2 CTAs per SM)
    bool useTpp = false;         // ... and the current batch does (integNum | 128)
    TppLayout tppLay{};
    int tppCtas = 0;             // resident CTAs per SM
    DevBuf tppSlot;              // [nparam] slot of each parameter in a CTA's patch slab
    // fed steps (vn_loss_grad_fed_*): copy stream + one event per uploaded chunk
    cudaStream_t copyStream = nullptr;
    std::vector<cudaEvent_t> fedEvents;
    // pageable feeds: two pinned bounce buffers, the event of the last transfer out of each, and the copy threads
    void* pin[2] = {nullptr, nullptr}; size_t pinBytes = 0; cudaEvent_t pinEv[2] = {nullptr, nullptr}; int pinNext = 0;
    HostPool* pool = nullptr;
    // boundary/initial adjoint kernel runs concurrently with the variational one (fork/join, also inside the step graph)
    cudaStream_t auxStream = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    DevBuf ticket;               // block counter of the fused finalize + optimizer kernel
    bool fused = false;          // per-test-function residual reduced inside the adjoint kernel (integNum | TP)
    cudaStream_t stream = nullptr;      // engine-owned blocking stream (ordered w.r.t. the legacy default stream) or the caller's
    cudaStream_t ownStream = nullptr;
    int64_t launches = 0;
    int pdlGraphs = 0;           // 1: the last k-step graph was instantiated with programmatic kernel -> kernel edges
    bool graphOK = true;         // CUDA-graph capture of vn_train_step available (one graph per table slot)
    // table slots: several uploaded point tables can be resident; mini-batches select test functions of the
    // current table through a device index list (vn_select_table / vn_set_batch)
    std::vector<PointSet*> slots;
    PointSet* t = nullptr;
    DevBuf batchIdx, extraX, batchSeq, lossOut;
    const int* idxOverride = nullptr;   // capture of a mini-batch epoch: step i reads its index list in batchSeq directly
    // vn_train_batches_begin / _end: up to two calls in flight, each with its own pinned staging (index lists, the small
    // uploads that precede the call: extra inputs, BC/IC rows) and pinned result slots
    struct Pend {
        void* seq = nullptr; size_t seqBytes = 0; float* loss = nullptr; int* err = nullptr; cudaEvent_t ev = nullptr; int k = 0;
        unsigned char* aux = nullptr; size_t auxOff = 0;
    } pend[2];
    int pendNext = 0;            // slot of the next vn_train_batches_begin
    bool indexed = false;        // batch = index list into the table (else: the whole table in order)
    int nExtra = 0;              // trailing MLP inputs supplied as per-call constants (MOR parameters)
    // parameters + optimizer
    DevBuf theta, m, v, gbuf, wts, stepbuf, corrbuf, lossRing;
    // interior points of the current batch (P = nb * integNum)
    DevBuf Iw, R, lossVec, segSum;
    unsigned int P = 0, nb = 0;
    // boundary / initial rows
    DevBuf bcols, blabel, cj;
    long long bstride = 0; unsigned int nbi = 0, bDof = 0; float biDimVal = 0.f;
    // adjoint partial slabs
    DevBuf partVar, partBic, part32Var, part32Bic, stashVar, stashBic, lossPart; int gridVar = 0, gridBic = 0;
    // scratch
    DevBuf stage, evalCols, evalOut;
    // geometry
    TileGeom gVarFwd, gVarAdj, gBicFwd, gBicAdj, gEval, gRes;
    bool weightsSet = false;
    bool resOK = true;           // strong-form residual kernel available for this depth/width
    // multi-GPU: this tower's own NCCL communicator (vn_comm_init); the step's all-reduce runs on the engine stream
    void* comm = nullptr; int commRank = 0, commWorld = 1;
    void* l2WindowPtr = nullptr; size_t l2WindowBytes = 0; cudaStream_t l2WindowStream = nullptr;   // persisting-L2 window (set_stash_window)
    // optional per-kernel CUDA-event timing (vn_profile_enable / vn_profile_read)
    bool profOn = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profPending[VN_PROF_SLOTS];
    double profMs[VN_PROF_SLOTS] = {0};
    int64_t profCnt[VN_PROF_SLOTS] = {0};
};

// RAII bracket: records an event pair around a kernel launch on the engine's stream when profiling is on
struct ProfScope {
    vn_engine* e; int slot; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(vn_engine* e_, int slot_) : e(e_), slot(slot_) {
        if (!e->profOn) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, e->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, e->stream);
        e->profPending[slot].push_back({a, b});
    }
};

static const int kPad = 256;
static const int kLossRing = 4096;      // loss history kept on the device for back-to-back steps (vn_train_steps)

static void drop_graph(PointSet* t) {
    if (t && t->graph) { cudaGraphExecDestroy(t->graph); t->graph = nullptr; }
    if (t && t->graphK) { cudaGraphExecDestroy(t->graphK); t->graphK = nullptr; }
    if (t) { t->graphLr = -1.f; t->graphKLr = -1.f; t->graphKn = 0; }
}
static void drop_graph(vn_engine* e) {            // everything captured so far is stale (stream / BC-IC table changed)
    for (PointSet* t : e->slots) drop_graph(t);
}    // point-table padding: multiple of every tile size

static int build_net(const vn_config& c, NetDesc* n) {
    memset(n, 0, sizeof(*n));
    n->L = c.nLayers; n->inpDim = c.inpDim;
    int off = 0;
    for (int l = 0; l <= c.nLayers; ++l) {
        const int wi = l == 0 ? c.inpDim : c.widths[l - 1];
        const int wo = l == c.nLayers ? 1 : c.widths[l];
        n->woff[l] = off; off += wi * wo;
        n->boff[l] = off; off += wo;
        if (l < c.nLayers) { n->width[l] = c.widths[l]; n->wpad[l] = (c.widths[l] + 3) & ~3; }
    }
    n->nparam = off;
    return 0;
}

extern "C" int vn_create(const vn_config* cfg, vn_engine** out) {
    if (!cfg || !out) return fail(VN_E_INVALID, "null argument");
    if (cfg->dim < 1 || cfg->dim > 2) return fail(VN_E_INVALID, "dim must be 1 or 2 (got %d)", cfg->dim);
    if (cfg->nLayers < 1 || cfg->nLayers > VN_MAX_LAYERS) return fail(VN_E_INVALID, "nLayers must be in [1,%d]", VN_MAX_LAYERS);
    const int minInp = cfg->dim + (cfg->timeDependent ? 1 : 0);
    if (cfg->inpDim < minInp || cfg->inpDim > VN_MAX_INPDIM) return fail(VN_E_INVALID, "inpDim must be in [%d,%d]", minInp, VN_MAX_INPDIM);
    if (cfg->act != VN_ACT_SIGMOID && cfg->act != VN_ACT_TANH) return fail(VN_E_INVALID, "unknown activation id %d", cfg->act);
    if (cfg->optimizer != VN_OPT_ADAM && cfg->optimizer != VN_OPT_RMSPROP) return fail(VN_E_INVALID, "unknown optimizer requested!");
    int wmax = 0;
    for (int l = 0; l < cfg->nLayers; ++l) {
        if (cfg->widths[l] < 1) return fail(VN_E_INVALID, "layer width must be positive");
        wmax = std::max(wmax, (int)cfg->widths[l]);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(VN_E_CUDA, "no CUDA device visible: the varnet_b200 engine has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(VN_E_INVALID, "requested processor %d is unavailable!", cfg->device);
    CK(cudaSetDevice(cfg->device));

    vn_engine* e = new (std::nothrow) vn_engine();
    if (!e) return fail(VN_E_INVALID, "out of host memory");
    e->cfg = *cfg;
    if (cudaStreamCreate(&e->ownStream) == cudaSuccess) e->stream = e->ownStream;
    if (cudaStreamCreateWithFlags(&e->auxStream, cudaStreamNonBlocking) != cudaSuccess) e->auxStream = nullptr;
    if (e->auxStream && (cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming) != cudaSuccess ||
                         cudaEventCreateWithFlags(&e->evJoin, cudaEventDisableTiming) != cudaSuccess)) {
        cudaStreamDestroy(e->auxStream); e->auxStream = nullptr;
    }
    e->slots.push_back(new PointSet());
    e->t = e->slots[0];
    build_net(*cfg, &e->net);
    e->S = 1 + cfg->dim;
    if (wmax > 256) { delete e; return fail(VN_E_UNSUPPORTED, "hidden width %d exceeds the compiled kernel families (<=256)", wmax); }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    e->numSMs = prop.multiProcessorCount;
    const int L = cfg->nLayers, act = cfg->act;
    const char* forceCls = getenv("VARNET_B200_CLASS");            // "tc": tensor-core class for any width (crossover measurements)
    if (wmax > 64 || (forceCls && !strcmp(forceCls, "tc"))) {
        // tensor-core class: tcgen05 3xTF32 layer GEMMs, activations of a chunk streamed through global memory (vn_tc.h)
        e->wclass = 256;
        if (!vn_tc_geometry(e->net, e->S, e->numSMs, &e->tcGeom)) { delete e; return fail(VN_E_UNSUPPORTED, "no compiled kernel for this configuration"); }
        if (e->tcGeom.smemGw > prop.sharedMemPerBlockOptin) {
            const size_t need = e->tcGeom.smemGw; delete e; return fail(VN_E_UNSUPPORTED, "tensor-core class needs %zu B of shared memory per CTA", need);
        }
        cudaError_t ce = vn_tc_prepare(e->S, act);
        if (ce != cudaSuccess) { delete e; return fail(VN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); }
        e->graphOK = false;          // thousands of launches per step: plain stream launches
        CK(e->tcWork.ensure(e->tcGeom.workBytes));
        CK(e->tcAcc.ensure((size_t)(e->net.nparam + 2) * sizeof(double)));
        CK(e->tcErr.ensure(sizeof(int)));
        CK(cudaMemset(e->tcErr.p, 0, sizeof(int)));
    } else {
    e->wclass = wmax <= 16 ? 16 : (wmax <= 32 ? 32 : 64);
    if (e->wclass == 64) {       // deep 64-wide networks: the weights of all layers no longer fit next to 64-point tiles
        TileGeom probe;
        if (vn_tile_geometry(e->S, 64, act, MODE_VAR_ADJ, L, &probe) && probe.smemBytes > prop.sharedMemPerBlockOptin)
            e->wclass = 164;
    }
    bool ok = vn_tile_geometry(e->S, e->wclass, act, MODE_VAR_FWD, L, &e->gVarFwd) &&
              vn_tile_geometry(e->S, e->wclass, act, MODE_VAR_ADJ, L, &e->gVarAdj) &&
              vn_tile_geometry(1, e->wclass, act, MODE_BIC_FWD, L, &e->gBicFwd) &&
              vn_tile_geometry(1, e->wclass, act, MODE_BIC_ADJ, L, &e->gBicAdj) &&
              vn_tile_geometry(1, e->wclass, act, MODE_EVAL, L, &e->gEval) &&
              vn_tile_geometry(VN_S_RES, e->wclass, act, MODE_RESIDUAL, L, &e->gRes);
    if (!ok) { delete e; return fail(VN_E_UNSUPPORTED, "no compiled kernel for this configuration"); }
    const size_t smemMax = prop.sharedMemPerBlockOptin;
    const TileGeom* gs[7] = {&e->gVarFwd, &e->gVarAdj, &e->gBicFwd, &e->gBicAdj, &e->gEval, &e->gVarAdj, &e->gRes};
    const int modes[7] = {MODE_VAR_FWD, MODE_VAR_ADJ, MODE_BIC_FWD, MODE_BIC_ADJ, MODE_EVAL, MODE_VAR_FUSED, MODE_RESIDUAL};
    for (int k = 0; k < 7; ++k) {
        if (gs[k]->smemBytes > smemMax && k == 6) { e->resOK = false; continue; }   // evaluation-only kernel: not fatal
        if (gs[k]->smemBytes > smemMax) {
            const size_t need = gs[k]->smemBytes;
            delete e;
            return fail(VN_E_UNSUPPORTED, "network needs %zu B of shared memory per CTA (limit %zu): depth/width outside the resident-tile kernel family", need, smemMax);
        }
        const int S = (k < 2 || k == 5) ? e->S : (k == 6 ? VN_S_RES : 1);
        cudaError_t ce = vn_tile_prepare(S, e->wclass, act, modes[k], gs[k]->smemBytes);
        if (ce != cudaSuccess) { delete e; return fail(VN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); }
    }
    }
    {
        // 64-wide class: the variational term runs on the tensor-core tile kernel unless VARNET_B200_CLASS=fma
        const bool wantTc64 = !(forceCls && !strcmp(forceCls, "fma"));
        if ((e->wclass == 64 || e->wclass == 164) && wantTc64 && vn_tc64_supported(e->net, e->S)) {
            vn_tc64_geometry(e->net, e->S, &e->tc64Geom);
            // no silent fall-back to the 2.5x slower FMA tiles: a build whose tensor-core kernel cannot be configured is an error
            if (e->tc64Geom.smemBytes > prop.sharedMemPerBlockOptin) {
                const size_t need = e->tc64Geom.smemBytes; delete e;
                return fail(VN_E_UNSUPPORTED, "width-64 tensor-core tile kernel needs %zu B of shared memory per CTA (set VARNET_B200_CLASS=fma for the FMA tiles)", need);
            }
            cudaError_t ce64 = vn_tc64_prepare(e->S, act, e->tc64Geom.smemBytes);
            if (ce64 != cudaSuccess) { delete e; return fail(VN_E_CUDA, "width-64 tensor-core tile kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(ce64)); }
            {
                e->tc64 = true;
                CK(e->tc64Img.ensure((size_t)e->tc64Geom.nImages * 8192 * sizeof(float)));
                CK(e->tc64Flat.ensure((size_t)e->net.nparam * sizeof(double)));
                CK(e->tcErr.ensure(sizeof(int)));
                CK(cudaMemset(e->tcErr.p, 0, sizeof(int)));
            }
        }
    }
    {
        // narrow networks: the variational term runs on the thread-per-point kernel unless VARNET_B200_CLASS=fma; networks whose
        // per-point rows leave room for fewer than two CTAs per SM stay on the FMA tiles (a class choice, both are CUDA kernels)
        const bool wantTpp = !(forceCls && !strcmp(forceCls, "fma"));
        if ((e->wclass == 16 || e->wclass == 32) && wantTpp && vn_tpp_supported(e->net, e->S)) {
            vn_tpp_layout(e->net, e->S, &e->tppLay);
            if (e->tppLay.smemBytes * 2 + 2048 <= prop.sharedMemPerMultiprocessor && e->tppLay.smemBytes <= prop.sharedMemPerBlockOptin) {
                int ctas = 0;
                cudaError_t cet = vn_tpp_prepare(e->S, act, e->tppLay.smemBytes, &ctas);
                if (cet != cudaSuccess) { delete e; return fail(VN_E_CUDA, "thread-per-point kernel: %s", cudaGetErrorString(cet)); }
                if (ctas >= 2) {
                    e->tpp = true; e->tppCtas = ctas;
                    std::vector<int> slots((size_t)e->net.nparam);
                    vn_tpp_param_slots(e->net, e->tppLay, slots.data());
                    // behind them the inverse map [npatch * 32]: parameter of a slab slot, -1 for padding (vn_finalize_kernel, patch mode)
                    const size_t np0 = slots.size();
                    slots.resize(np0 + (size_t)e->tppLay.npatch * 32, -1);
                    for (size_t i = 0; i < np0; ++i) slots[np0 + slots[i]] = (int)i;
                    CK(e->tppSlot.ensure(slots.size() * sizeof(int)));
                    CK(cudaMemcpy(e->tppSlot.p, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice));
                }
            }
        }
    }
    const int np = e->net.nparam;
    CK(e->theta.ensure(np * sizeof(float)));
    CK(e->m.ensure(np * sizeof(float)));
    CK(e->v.ensure(np * sizeof(float)));
    CK(e->gbuf.ensure((np + 4) * sizeof(float)));
    CK(e->wts.ensure(4 * sizeof(float)));
    CK(e->stepbuf.ensure(sizeof(long long)));
    CK(e->corrbuf.ensure(sizeof(double)));
    CK(e->ticket.ensure(sizeof(unsigned int)));
    CK(e->lossRing.ensure(kLossRing * sizeof(float)));
    CK(cudaMemset(e->ticket.p, 0, sizeof(unsigned int)));
    CK(cudaMemset(e->theta.p, 0, np * sizeof(float)));
    CK(cudaMemset(e->gbuf.p, 0, (np + 4) * sizeof(float)));
    const float w1[4] = {1.f, 1.f, 1.f, 0.f};
    CK(cudaMemcpy(e->wts.p, w1, sizeof(w1), cudaMemcpyHostToDevice));
    *out = e;
    const float* nullf = nullptr;
    return vn_set_optimizer_state(e, nullf, nullf, np, 0);
}

extern "C" int vn_destroy(vn_engine* e) {
    if (!e) return VN_OK;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    drop_graph(e);
    if (e->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(e->comm); e->comm = nullptr; }
    if (e->ownStream) cudaStreamDestroy(e->ownStream);
    if (e->copyStream) cudaStreamDestroy(e->copyStream);
    if (e->auxStream) cudaStreamDestroy(e->auxStream);
    if (e->evFork) cudaEventDestroy(e->evFork);
    if (e->evJoin) cudaEventDestroy(e->evJoin);
    for (cudaEvent_t ev : e->fedEvents) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) { if (e->pin[i]) cudaFreeHost(e->pin[i]); if (e->pinEv[i]) cudaEventDestroy(e->pinEv[i]); }
    for (auto& p : e->pend) {
        if (p.seq) cudaFreeHost(p.seq);
        if (p.loss) cudaFreeHost(p.loss);
        if (p.err) cudaFreeHost(p.err);
        if (p.aux) cudaFreeHost(p.aux);
        if (p.ev) cudaEventDestroy(p.ev);
    }
    delete e->pool;
    for (PointSet* t : e->slots) { t->cols.release(); t->integW.release(); t->detJ.release(); t->genBuf.release(); delete t; }
    DevBuf* bufs[] = {&e->theta, &e->m, &e->v, &e->gbuf, &e->wts, &e->stepbuf, &e->corrbuf, &e->batchIdx, &e->extraX,
                      &e->Iw, &e->R, &e->lossVec, &e->segSum, &e->bcols, &e->blabel, &e->cj, &e->partVar,
                      &e->partBic, &e->part32Var, &e->part32Bic, &e->stashVar, &e->stashBic, &e->lossPart, &e->stage, &e->evalCols, &e->evalOut,
                      &e->tcWork, &e->tcAcc, &e->tcErr, &e->ticket, &e->tc64Img, &e->tc64Flat, &e->tppSlot, &e->lossRing, &e->batchSeq, &e->lossOut};
    for (DevBuf* b : bufs) b->release();
    delete e;
    return VN_OK;
}

extern "C" int vn_set_stream(vn_engine* e, void* s) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    drop_graph(e);
    e->stream = s ? reinterpret_cast<cudaStream_t>(s) : e->ownStream;
    return VN_OK;
}
extern "C" int vn_synchronize(vn_engine* e) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_param_count(const vn_engine* e, int64_t* n) {
    if (!e || !n) return fail(VN_E_INVALID, "null argument");
    *n = e->net.nparam;
    return VN_OK;
}

extern "C" int vn_set_optimizer_state(vn_engine* e, const float* m, const float* v, int64_t n, int64_t step) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (n != e->net.nparam) return fail(VN_E_INVALID, "parameter count mismatch: got %lld, expected %d", (long long)n, e->net.nparam);
    CK(cudaSetDevice(e->cfg.device));
    const int np = e->net.nparam;
    if (m) CK(cudaMemcpyAsync(e->m.p, m, np * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    else CK(cudaMemsetAsync(e->m.p, 0, np * sizeof(float), e->stream));
    if (v) CK(cudaMemcpyAsync(e->v.p, v, np * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    else if (e->cfg.optimizer == VN_OPT_RMSPROP) {      // TF initialises the rms slot to ones
        vn_fill_kernel<<<(np + 255) / 256, 256, 0, e->stream>>>(e->v.as<float>(), 1.0f, np);
        CK(cudaGetLastError());
    } else CK(cudaMemsetAsync(e->v.p, 0, np * sizeof(float), e->stream));
    const long long st = step;
    const double tn = (double)(step + 1);
    const double corr = sqrt(1.0 - pow(0.999, tn)) / (1.0 - pow(0.9, tn));
    CK(cudaMemcpyAsync(e->stepbuf.p, &st, sizeof(st), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->corrbuf.p, &corr, sizeof(corr), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_get_optimizer_state(vn_engine* e, float* m, float* v, int64_t n, int64_t* step) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (n != e->net.nparam) return fail(VN_E_INVALID, "parameter count mismatch");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    if (m) CK(cudaMemcpy(m, e->m.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (v) CK(cudaMemcpy(v, e->v.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (step) { long long st = 0; CK(cudaMemcpy(&st, e->stepbuf.p, sizeof(st), cudaMemcpyDeviceToHost)); *step = st; }
    return VN_OK;
}
extern "C" int vn_set_params(vn_engine* e, const float* theta, int64_t n) {
    if (!e || !theta) return fail(VN_E_INVALID, "null argument");
    if (n != e->net.nparam) return fail(VN_E_INVALID, "parameter count mismatch: got %lld, expected %d", (long long)n, e->net.nparam);
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(e->theta.p, theta, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return vn_set_optimizer_state(e, nullptr, nullptr, n, 0);
}
extern "C" int vn_get_params(vn_engine* e, float* theta, int64_t n) {
    if (!e || !theta) return fail(VN_E_INVALID, "null argument");
    if (n != e->net.nparam) return fail(VN_E_INVALID, "parameter count mismatch");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(theta, e->theta.p, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_set_weights(vn_engine* e, const float w[3]) {
    if (!e || !w) return fail(VN_E_INVALID, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(e->wts.p, w, 3 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->weightsSet = true;
    return VN_OK;
}

// ------------------------------------------------------------------ uploads
static const long long kChunk = 1 << 22;    // rows per staging chunk

// The per-CTA activation stash is rewritten by every tile and read back a few microseconds later: it belongs in L2.  The
// kernels already tag these accesses evict_last, but that is only a hint (round 1: 15.4 GB of stash write-back per 6.4e7-point
// launch).  Here the stash is additionally declared a PERSISTING access-policy window of the engine's stream, with the L2
// set-aside sized for it, and the rest of the step's traffic (the streamed point table) as streaming.  Best effort: a device
// that refuses any of the calls simply keeps the hints.  VARNET_B200_L2_PERSIST=0 disables it (A/B measurements).
static void set_stash_window(vn_engine* e) {
    static const bool off = [] { const char* v = getenv("VARNET_B200_L2_PERSIST"); return v && !strcmp(v, "0"); }();
    if (off || !e->stashVar.p || e->stashVar.bytes < (1u << 20) || !e->stream) return;
    if (e->l2WindowPtr == e->stashVar.p && e->l2WindowBytes == e->stashVar.bytes && e->l2WindowStream == e->stream) return;
    int maxPersist = 0, maxWindow = 0;
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, e->cfg.device);
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, e->cfg.device);
    if (maxPersist <= 0 || maxWindow <= 0) { cudaGetLastError(); return; }
    const size_t want = std::min<size_t>(e->stashVar.bytes, (size_t)maxPersist);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = e->stashVar.p;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>(e->stashVar.bytes, (size_t)maxWindow);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) { cudaGetLastError(); return; }
    e->l2WindowPtr = e->stashVar.p; e->l2WindowBytes = e->stashVar.bytes; e->l2WindowStream = e->stream;
    drop_graph(e);          // captured kernel nodes carry the window they were captured with
}

// size the per-batch work buffers and pick the single-pass (fused residual) path when integNum | TP
static int ensure_work(vn_engine* e) {
    PointSet* t = e->t;
    const long long P = (long long)e->nb * t->integNum;
    e->P = (unsigned int)P;
    CK(e->Iw.ensure((size_t)std::max<long long>(P, 1) * sizeof(float)));
    CK(e->R.ensure((size_t)std::max<unsigned>(e->nb, 1) * sizeof(float)));
    CK(e->lossVec.ensure((size_t)std::max<unsigned>(e->nb, 1) * sizeof(float)));
    const int nSeg = (int)((e->nb + 255) / 256);
    CK(e->segSum.ensure((size_t)std::max(nSeg, 1) * sizeof(double)));
    if (e->wclass == 256) { e->fused = false; e->gridVar = 0; return VN_OK; }     // tensor-core class: chunk workspace is fixed at creation
    e->useTc64 = e->tc64 && t->integNum > 0 && (128 % t->integNum) == 0;
    if (e->useTc64) {
        const long long tiles = (P + 127) / 128;
        e->gridVar = (int)std::max<long long>(1, std::min<long long>(tiles, e->numSMs));
        CK(e->partVar.ensure((size_t)e->numSMs * e->tc64Geom.psz * sizeof(double)));
        CK(e->part32Var.ensure((size_t)e->numSMs * e->tc64Geom.psz * sizeof(float)));
        CK(e->stashVar.ensure(std::max<size_t>(16, (size_t)e->numSMs * e->tc64Geom.stashFloats * sizeof(float))));
        CK(e->lossPart.ensure((size_t)e->numSMs * e->tc64Geom.lossSlots * sizeof(double)));
        e->fused = true;
        set_stash_window(e);
        return VN_OK;
    }
    e->useTpp = e->tpp && t->integNum > 0 && (128 % t->integNum) == 0;
    if (e->useTpp) {
        // balanced persistent grid: every CTA gets the same number of 128-point tiles (+-1)
        const long long tiles = std::max<long long>(1, (P + 127) / 128), cap = (long long)e->numSMs * e->tppCtas;
        const long long rounds = (tiles + cap - 1) / cap;
        e->gridVar = (int)((tiles + rounds - 1) / rounds);
        CK(e->partVar.ensure((size_t)cap * e->tppLay.npatch * 32 * sizeof(double)));
        CK(e->lossPart.ensure((size_t)cap * 4 * sizeof(double)));
        e->fused = true;
        return VN_OK;
    }
    const long long tilesAdj = (P + e->gVarAdj.TP - 1) / e->gVarAdj.TP;
    e->gridVar = (int)std::max<long long>(1, std::min<long long>(tilesAdj, e->numSMs));
    CK(e->partVar.ensure((size_t)e->numSMs * e->gVarAdj.pl.psz * sizeof(double)));
    CK(e->part32Var.ensure((size_t)e->numSMs * e->gVarAdj.pl.psz * sizeof(float)));
    CK(e->stashVar.ensure(std::max<size_t>(16, (size_t)e->numSMs * e->gVarAdj.stashFloats * sizeof(float))));
    CK(e->lossPart.ensure((size_t)e->numSMs * (e->gVarAdj.NT / 32) * sizeof(double)));
    e->fused = (e->gVarAdj.TP % t->integNum) == 0;
    set_stash_window(e);
    return VN_OK;
}

// Upload one point table into the current slot.  X has `nx` columns; the remaining inpDim-nx MLP inputs are
// per-call constants (vn_set_extra_inputs).  The batch is reset to "whole table, in order".
// A fed step uploads the table chunk by chunk on the copy stream and launches the adjoint kernel once per chunk
// on the engine stream as soon as that chunk is packed (cudaStreamWaitEvent): copies overlap the step's kernels.
// tile size in which a fed step (vn_loss_grad_fed_*) hands uploaded chunks to the kernels of this engine
static int fed_tile(const vn_engine* e, long long integNum) {
    if (e->wclass == 256) return 128;                                                   // tensor-core class: 128-point tiles
    if ((e->tc64 || e->tpp) && integNum > 0 && (128 % integNum) == 0) return 128;       // resident-tile tensor-core / thread-per-point kernels
    if (integNum > 0 && (e->gVarAdj.TP % integNum) == 0) return e->gVarAdj.TP;          // fused FMA tiles
    return e->gVarFwd.TP;                                                               // two-pass class: the forward pass runs per chunk
}
struct FedPlan {
    struct Sub { int tile0, ntiles; cudaEvent_t ev; };
    // chunk k of the table: staged, copied and packed on the copy stream when run_loss asks for it, so that the host-side staging
    // of chunk k+1 (pageable feeds) runs while the kernel of chunk k is already executing.  Returns 1 (produced), 0 (done), < 0 (error).
    std::function<int(size_t, Sub*)> produce;
};

static const long long kSubChunk = 1 << 20;      // rows per pinned bounce buffer
static int ensure_stager(vn_engine* e, size_t bytes) {
    if (!e->pool) {
        int n = 8;
        if (const char* v = getenv("VARNET_B200_STAGE_THREADS")) n = atoi(v);
        const int hw = (int)std::thread::hardware_concurrency();
        n = std::max(1, std::min(n, hw > 0 ? hw : 1));
        e->pool = new HostPool(n - 1);
    }
    if (bytes > e->pinBytes) {
        for (int i = 0; i < 2; ++i) {
            if (e->pinEv[i]) CK(cudaEventSynchronize(e->pinEv[i]));
            if (e->pin[i]) { cudaFreeHost(e->pin[i]); e->pin[i] = nullptr; }
            CK(cudaHostAlloc(&e->pin[i], bytes, cudaHostAllocDefault));
            if (!e->pinEv[i]) CK(cudaEventCreateWithFlags(&e->pinEv[i], cudaEventDisableTiming));
        }
        e->pinBytes = bytes;
    }
    return VN_OK;
}
// rows [off, off + n) of the caller's pageable arrays -> the device staging blocks, through the pinned bounce buffers.  The host
// threads also do the feed cast: float64 arrays are rounded to float32 (round-to-nearest-even, the same rounding as the
// placeholder cast and as __double2float_rn) while they are copied, which halves the pinned traffic and the PCIe transfer.
template <typename T>
static int stage_pageable(vn_engine* e, cudaStream_t us, const T* X, int nx, const T* G, int dim, const T* dNt, const T* src, const T* N,
                          long long off, long long n, float* sX, float* sG, float* sT, float* sS, float* sN) {
    const int rowVals = nx + dim + (dNt ? 1 : 0) + (src ? 2 : 0);
    int rc = ensure_stager(e, (size_t)kSubChunk * rowVals * sizeof(float));
    if (rc) return rc;
    for (long long r0 = 0; r0 < n; r0 += kSubChunk) {
        const long long m = std::min(kSubChunk, n - r0);
        const int b = e->pinNext;
        e->pinNext ^= 1;
        CK(cudaEventSynchronize(e->pinEv[b]));           // the previous transfer out of this buffer has finished
        float* pX = reinterpret_cast<float*>(e->pin[b]);
        float* pG = pX + m * nx;
        float* pT = pG + m * dim;
        float* pS = pT + (dNt ? m : 0);
        float* pN = pS + (src ? m : 0);
        struct Seg { float* dst; const T* srcp; size_t count; };
        Seg segs[5] = {{pX, X + (off + r0) * nx, (size_t)m * nx}, {pG, G + (off + r0) * dim, (size_t)m * dim},
                       {pT, dNt ? dNt + off + r0 : nullptr, dNt ? (size_t)m : 0}, {pS, src ? src + off + r0 : nullptr, src ? (size_t)m : 0},
                       {pN, (N && src) ? N + off + r0 : nullptr, (N && src) ? (size_t)m : 0}};
        const int parts = e->pool->size();
        e->pool->run(parts, [&](int i) {
            for (const Seg& sg : segs) {
                if (!sg.count) continue;
                const size_t lo = sg.count * (size_t)i / parts / 16 * 16, hi = (i + 1 == parts) ? sg.count : sg.count * (size_t)(i + 1) / parts / 16 * 16;
                float* __restrict__ d = sg.dst;
                const T* __restrict__ sp = sg.srcp;
                for (size_t j = lo; j < hi; ++j) d[j] = (float)sp[j];
            }
        });
        CK(cudaMemcpyAsync(sX + r0 * nx, pX, (size_t)m * nx * sizeof(float), cudaMemcpyHostToDevice, us));
        CK(cudaMemcpyAsync(sG + r0 * dim, pG, (size_t)m * dim * sizeof(float), cudaMemcpyHostToDevice, us));
        if (dNt) CK(cudaMemcpyAsync(sT + r0, pT, (size_t)m * sizeof(float), cudaMemcpyHostToDevice, us));
        if (src) {
            CK(cudaMemcpyAsync(sS + r0, pS, (size_t)m * sizeof(float), cudaMemcpyHostToDevice, us));
            CK(cudaMemcpyAsync(sN + r0, pN, (size_t)m * sizeof(float), cudaMemcpyHostToDevice, us));
        }
        CK(cudaEventRecord(e->pinEv[b], us));
    }
    return VN_OK;
}

template <typename T>
static int upload_table(vn_engine* e, const T* X, int nx, const T* G, const T* src, const T* N, const T* dNt,
                        int64_t nb, int32_t integNum, const T* integW, const T* detJ, int32_t detJvec, FedPlan* plan = nullptr) {
    if (!e || !X || !G || !detJ) return fail(VN_E_INVALID, "Input, gcoef and detJ are required");
    if (nb < 1 || integNum < 1) return fail(VN_E_INVALID, "intShape must be positive");
    const vn_config& c = e->cfg;
    if (nx < c.dim + (c.timeDependent ? 1 : 0) || nx > c.inpDim)
        return fail(VN_E_INVALID, "the table must hold between %d and %d input columns", c.dim + (c.timeDependent ? 1 : 0), c.inpDim);
    const long long P = (long long)nb * integNum;
    if (P >= (1ll << 31) - kPad) return fail(VN_E_UNSUPPORTED, "more than 2^31 quadrature points per engine; shard the test functions");
    if (c.timeDependent && !dNt) return fail(VN_E_INVALID, "dNt is required for time-dependent problems");
    if (c.isSource && (!src || !N)) return fail(VN_E_INVALID, "source and N are required when lossOpt['isSource'] is set");
    if (c.integWflag && !integW) return fail(VN_E_INVALID, "integW is required when lossOpt['integWflag'] is set");
    CK(cudaSetDevice(c.device));
    PointSet* t = e->t;
    drop_graph(t);
    t->inKernel = false;
    int col = 0;
    t->nx = nx;
    t->colX = col; col += nx;
    t->colG = col; col += c.dim;
    t->colT = c.timeDependent ? col++ : -1;
    t->colS = c.isSource ? col++ : -1;
    t->ncols = col;
    t->pstride = (P + kPad - 1) / kPad * kPad;
    t->rows = (unsigned int)P; t->nbTab = (unsigned int)nb; t->integNum = (unsigned int)integNum; t->detJvec = detJvec ? 1 : 0;
    CK(t->cols.ensure((size_t)t->ncols * t->pstride * sizeof(float)));
    const int rowVals = nx + c.dim + 3;
    const long long chunk = std::min<long long>(kChunk, P);
    const long long nd = detJvec ? nb : 1;
    CK(e->stage.ensure(std::max((size_t)chunk * rowVals, (size_t)std::max<long long>(nd, integNum)) * sizeof(T)));
    if (plan) {
        // small tables first (they share the staging buffer), then the chunks on the copy stream without host syncs
        if (!e->copyStream) CK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
        t->hasIntegW = (c.integWflag && integW) ? 1 : 0;
        CK(t->detJ.ensure(nd * sizeof(float)));
        CK(cudaMemcpyAsync(e->stage.p, detJ, nd * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        vn_cast_kernel<T><<<(unsigned)((nd + 255) / 256), 256, 0, e->stream>>>(e->stage.as<T>(), t->detJ.as<float>(), nd);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        if (t->hasIntegW) {
            CK(t->integW.ensure(integNum * sizeof(float)));
            CK(cudaMemcpyAsync(e->stage.p, integW, integNum * sizeof(T), cudaMemcpyHostToDevice, e->stream));
            vn_cast_kernel<T><<<(integNum + 255) / 256, 256, 0, e->stream>>>(e->stage.as<T>(), t->integW.as<float>(), integNum);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(e->stream));
        }
        e->launches += 1 + t->hasIntegW;
        cudaStream_t us = e->copyStream;
        // only the zero padding behind the last row has to be cleared: every other element is overwritten by the packs
        if (t->pstride > P)
            for (int cc = 0; cc < t->ncols; ++cc)
                CK(cudaMemsetAsync(t->cols.as<float>() + (size_t)cc * t->pstride + P, 0, (size_t)(t->pstride - P) * sizeof(float), us));
        const int TP = fed_tile(e, integNum);
        // pageable caller arrays (what NumPy hands over) are staged through pinned bounce buffers by a few host threads
        static const bool stagerOff = [] { const char* v = getenv("VARNET_B200_STAGE_THREADS"); return v && atoi(v) <= 0; }();
        const bool pageable = !stagerOff && is_pageable(X) && is_pageable(G);
        // the first kernel launch of the step waits for the first chunk: chunks 0, 1, 2 are 1/4, 1/4 and 1/2 of the regular size
        // (1 M, 1 M, 2 M rows, then 4 M each), so the kernels start after 24 MB instead of 96 MB have crossed the bus — what
        // the end-to-end step loses against the resident one, and most of it when the ranks of a multi-GPU job feed 1/N each
        const long long q = chunk / 4;
        const bool ramp = chunk == kChunk && q % TP == 0 && q / TP >= e->numSMs;
        plan->produce = [=](size_t k, FedPlan::Sub* out) -> int {
            long long off = (long long)k * chunk, len = chunk;
            if (ramp) {
                if (k < 2) { off = (long long)k * q; len = q; }
                else if (k == 2) { off = 2 * q; len = 2 * q; }
                else off = (long long)(k - 2) * chunk;
            }
            if (off >= P) return 0;
            const long long n = std::min(len, P - off);
            T* sX = e->stage.as<T>();
            T* sG = sX + n * nx;
            T* sT = sG + n * c.dim;
            T* sS = sT + n;
            T* sN = sS + n;
            if (pageable) {
                // staged and cast to float32 by the host threads: the device blocks and the pack kernel are float
                float* fX = e->stage.as<float>();
                float* fG = fX + n * nx;
                float* fT = fG + n * c.dim;
                float* fS = fT + n;
                float* fN = fS + n;
                int rcs = stage_pageable<T>(e, us, X, nx, G, c.dim, t->colT >= 0 ? dNt : nullptr, t->colS >= 0 ? src : nullptr,
                                            t->colS >= 0 ? N : nullptr, off, n, fX, fG, fT, fS, fN);
                if (rcs) return rcs;
                vn_pack_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, us>>>(
                    fX, nx, fG, c.dim, t->colT >= 0 ? fT : nullptr, t->colS >= 0 ? fS : nullptr,
                    t->colS >= 0 ? fN : nullptr, t->cols.as<float>(), t->pstride, off, n, t->colX, t->colG, t->colT, t->colS);
            } else {
                CK(cudaMemcpyAsync(sX, X + off * nx, n * nx * sizeof(T), cudaMemcpyHostToDevice, us));
                CK(cudaMemcpyAsync(sG, G + off * c.dim, n * c.dim * sizeof(T), cudaMemcpyHostToDevice, us));
                if (t->colT >= 0) CK(cudaMemcpyAsync(sT, dNt + off, n * sizeof(T), cudaMemcpyHostToDevice, us));
                if (t->colS >= 0) {
                    CK(cudaMemcpyAsync(sS, src + off, n * sizeof(T), cudaMemcpyHostToDevice, us));
                    CK(cudaMemcpyAsync(sN, N + off, n * sizeof(T), cudaMemcpyHostToDevice, us));
                }
                vn_pack_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, us>>>(
                    sX, nx, sG, c.dim, t->colT >= 0 ? sT : nullptr, t->colS >= 0 ? sS : nullptr,
                    t->colS >= 0 ? sN : nullptr, t->cols.as<float>(), t->pstride, off, n, t->colX, t->colG, t->colT, t->colS);
            }
            CK(cudaGetLastError());
            e->launches++;
            if (e->fedEvents.size() <= k) {
                cudaEvent_t ev;
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                e->fedEvents.push_back(ev);
            }
            CK(cudaEventRecord(e->fedEvents[k], us));
            *out = {(int)(off / TP), (int)((n + TP - 1) / TP), e->fedEvents[k]};
            return 1;
        };
        t->loaded = true;
        e->indexed = false;
        e->nb = t->nbTab;
        return ensure_work(e);
    }
    CK(cudaMemsetAsync(t->cols.p, 0, (size_t)t->ncols * t->pstride * sizeof(float), e->stream));
    for (long long off = 0; off < P; off += chunk) {
        const long long n = std::min(chunk, P - off);
        T* sX = e->stage.as<T>();
        T* sG = sX + n * nx;
        T* sT = sG + n * c.dim;
        T* sS = sT + n;
        T* sN = sS + n;
        CK(cudaMemcpyAsync(sX, X + off * nx, n * nx * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(sG, G + off * c.dim, n * c.dim * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        if (t->colT >= 0) CK(cudaMemcpyAsync(sT, dNt + off, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        if (t->colS >= 0) {
            CK(cudaMemcpyAsync(sS, src + off, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
            CK(cudaMemcpyAsync(sN, N + off, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        }
        vn_pack_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(
            sX, nx, sG, c.dim, t->colT >= 0 ? sT : nullptr, t->colS >= 0 ? sS : nullptr,
            t->colS >= 0 ? sN : nullptr, t->cols.as<float>(), t->pstride, off, n, t->colX, t->colG, t->colT, t->colS);
        CK(cudaGetLastError());
        e->launches++;
        CK(cudaStreamSynchronize(e->stream));      // staging buffer is reused by the next chunk
    }
    // small tables
    t->hasIntegW = (c.integWflag && integW) ? 1 : 0;
    CK(t->detJ.ensure(nd * sizeof(float)));
    CK(cudaMemcpyAsync(e->stage.p, detJ, nd * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    vn_cast_kernel<T><<<(unsigned)((nd + 255) / 256), 256, 0, e->stream>>>(e->stage.as<T>(), t->detJ.as<float>(), nd);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    if (t->hasIntegW) {
        CK(t->integW.ensure(integNum * sizeof(float)));
        CK(cudaMemcpyAsync(e->stage.p, integW, integNum * sizeof(T), cudaMemcpyHostToDevice, e->stream));
        vn_cast_kernel<T><<<(integNum + 255) / 256, 256, 0, e->stream>>>(e->stage.as<T>(), t->integW.as<float>(), integNum);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
    }
    e->launches += 1 + t->hasIntegW;
    t->loaded = true;
    e->indexed = false;
    e->nb = t->nbTab;
    return ensure_work(e);
}

template <typename T>
static int upload_points(vn_engine* e, const T* X, const T* G, const T* src, const T* N, const T* dNt, int64_t nb,
                         int32_t integNum, const T* integW, const T* detJ, int32_t detJvec) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    e->nExtra = 0;
    return upload_table<T>(e, X, e->cfg.inpDim, G, src, N, dNt, nb, integNum, integW, detJ, detJvec);
}

extern "C" int vn_generate_table_f64(vn_engine* e, const double* coord, int64_t nSpace, const double* tcoord, int64_t nTime,
                                     const double* hVec, const double* delta, const double* N, const double* dN,
                                     double diff, const double* vel, double source, int64_t tf0, int64_t nb,
                                     int32_t integNum, const double* integW, double detJ) {
    if (!e || !coord || !hVec || !delta || !N || !dN || !vel) return fail(VN_E_INVALID, "null argument");
    const vn_config& c = e->cfg;
    const int feDim = c.dim + (c.timeDependent ? 1 : 0);
    if (c.timeDependent && (!tcoord || nTime < 1)) return fail(VN_E_INVALID, "time coordinates are required for time-dependent problems");
    if (!c.timeDependent) nTime = 1;
    if (nb < 1 || integNum < 1 || nSpace < 1) return fail(VN_E_INVALID, "intShape must be positive");
    if (tf0 < 0 || tf0 + nb > nSpace * nTime) return fail(VN_E_INVALID, "test-function range [%lld, %lld) outside the mesh (%lld)", (long long)tf0, (long long)(tf0 + nb), (long long)(nSpace * nTime));
    if (c.integWflag && !integW) return fail(VN_E_INVALID, "integW is required when lossOpt['integWflag'] is set");
    const long long P = (long long)nb * integNum;
    if (P >= (1ll << 31) - kPad) return fail(VN_E_UNSUPPORTED, "more than 2^31 quadrature points per engine; shard the test functions");
    CK(cudaSetDevice(c.device));
    PointSet* t = e->t;
    drop_graph(t);
    int col = 0;
    t->nx = feDim;
    t->colX = col; col += feDim;
    t->colG = col; col += c.dim;
    t->colT = c.timeDependent ? col++ : -1;
    t->colS = c.isSource ? col++ : -1;
    t->ncols = col;
    t->pstride = (P + kPad - 1) / kPad * kPad;
    t->rows = (unsigned int)P; t->nbTab = (unsigned int)nb; t->integNum = (unsigned int)integNum; t->detJvec = 0;
    // Engines whose variational passes (loss and loss + gradient) all run on the tensor-core tile kernel never read a materialised
    // table: the kernel regenerates each row from the centre of its test function (GenTab).  Nothing of size nT is allocated.
    static const bool inKernelOff = [] { const char* v = getenv("VARNET_B200_INKERNEL_GEN"); return v && !strcmp(v, "0"); }();
    t->inKernel = !inKernelOff && e->tc64 && vn_tc64_forward_only_available() && (128 % integNum) == 0 && c.dim <= 2 && integNum <= 4096;
    if (t->inKernel) {
        const size_t nC = (size_t)nSpace * c.dim, nTm = c.timeDependent ? (size_t)nTime : 0, nH = (size_t)feDim * integNum;
        std::vector<double> hd(nH);
        for (int d = 0; d < feDim; ++d)
            for (int q = 0; q < integNum; ++q) { volatile double pr = hVec[d] * delta[(size_t)d * integNum + q]; hd[(size_t)d * integNum + q] = pr; }
        std::vector<float> coef((size_t)integNum * 4, 0.f);
        for (int q = 0; q < integNum; ++q) {
            for (int k = 0; k < c.dim; ++k) {
                volatile double p1 = diff * dN[(size_t)q * feDim + k];      // separate roundings, as NumPy evaluates diff*dNx + vel*N
                volatile double p2 = vel[k] * N[q];
                volatile double sm = p1 + p2;
                coef[(size_t)q * 4 + k] = (float)sm;
            }
            if (c.timeDependent) coef[(size_t)q * 4 + 2] = (float)dN[(size_t)q * feDim + c.dim];
            coef[(size_t)q * 4 + 3] = (float)source * (float)N[q];
        }
        const size_t bytes = (nC + nTm + nH) * sizeof(double) + coef.size() * sizeof(float);
        CK(t->genBuf.ensure(bytes));
        double* dC = t->genBuf.as<double>();
        double* dT = dC + nC; double* dH = dT + nTm; float* dF = reinterpret_cast<float*>(dH + nH);
        CK(cudaMemcpyAsync(dC, coord, nC * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        if (nTm) CK(cudaMemcpyAsync(dT, tcoord, nTm * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(dH, hd.data(), nH * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(dF, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));          // hd / coef are locals
        t->genTab.coord = dC; t->genTab.tcoord = nTm ? dT : nullptr; t->genTab.hd = dH; t->genTab.coef = dF;
        t->genTab.nTime = nTime; t->genTab.tf0 = tf0; t->genTab.q = integNum; t->genTab.dim = c.dim; t->genTab.feDim = feDim;
        t->cols.release();
        t->hasIntegW = (c.integWflag && integW) ? 1 : 0;
        CK(t->detJ.ensure(sizeof(float)));
        const float djk = (float)detJ;
        CK(cudaMemcpyAsync(t->detJ.p, &djk, sizeof(float), cudaMemcpyHostToDevice, e->stream));
        if (t->hasIntegW) {
            std::vector<float> wf(integNum);
            for (int q = 0; q < integNum; ++q) wf[q] = (float)integW[q];
            CK(t->integW.ensure(integNum * sizeof(float)));
            CK(cudaMemcpyAsync(t->integW.p, wf.data(), integNum * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        }
        CK(cudaStreamSynchronize(e->stream));
        t->loaded = true;
        e->indexed = false;
        e->nb = t->nbTab;
        if (c.inpDim == feDim) e->nExtra = 0;
        return ensure_work(e);
    }
    CK(t->cols.ensure((size_t)t->ncols * t->pstride * sizeof(float)));
    CK(cudaMemsetAsync(t->cols.p, 0, (size_t)t->ncols * t->pstride * sizeof(float), e->stream));
    // small host tables -> staging (doubles): coord | tcoord | delta | N | dN | integW
    const size_t nC = (size_t)nSpace * c.dim, nT = c.timeDependent ? (size_t)nTime : 0, nD = (size_t)feDim * integNum;
    const size_t nTot = nC + nT + nD + (size_t)integNum + nD + (size_t)integNum + 1;
    CK(e->stage.ensure(nTot * sizeof(double)));
    double* sC = e->stage.as<double>();
    double* sT = sC + nC; double* sD = sT + nT; double* sN = sD + nD; double* sdN = sN + integNum; double* sW = sdN + nD;
    CK(cudaMemcpyAsync(sC, coord, nC * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (nT) CK(cudaMemcpyAsync(sT, tcoord, nT * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sD, delta, nD * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sN, N, (size_t)integNum * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sdN, dN, nD * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    GenArgs g;
    memset(&g, 0, sizeof(g));
    g.coord = sC; g.tcoord = sT; g.nTime = nTime; g.delta = sD; g.N = sN; g.dN = sdN;
    for (int d = 0; d < feDim; ++d) g.h[d] = hVec[d];
    g.diff = diff; for (int k = 0; k < c.dim; ++k) g.vel[k] = vel[k];
    g.source = source; g.tf0 = tf0; g.n = P; g.q = integNum; g.dim = c.dim; g.feDim = feDim;
    g.cols = t->cols.as<float>(); g.pstride = t->pstride; g.colX = t->colX; g.colG = t->colG; g.colT = t->colT; g.colS = t->colS;
    vn_generate_kernel<<<(unsigned)((P + 255) / 256), 256, 0, e->stream>>>(g);
    CK(cudaGetLastError());
    t->hasIntegW = (c.integWflag && integW) ? 1 : 0;
    CK(t->detJ.ensure(sizeof(float)));
    const float dj = (float)detJ;
    CK(cudaMemcpyAsync(t->detJ.p, &dj, sizeof(float), cudaMemcpyHostToDevice, e->stream));
    if (t->hasIntegW) {
        CK(t->integW.ensure(integNum * sizeof(float)));
        CK(cudaMemcpyAsync(sW, integW, (size_t)integNum * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        vn_cast_kernel<double><<<(integNum + 255) / 256, 256, 0, e->stream>>>(sW, t->integW.as<float>(), integNum);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(e->stream));
    e->launches += 1 + t->hasIntegW;
    t->loaded = true;
    e->indexed = false;
    e->nb = t->nbTab;
    if (c.inpDim == feDim) e->nExtra = 0;  // otherwise the trailing MLP inputs (MOR parameters) come from vn_set_extra_inputs
    return ensure_work(e);
}

// Host-to-device copy of a caller's (pageable) array that does not wait for the stream: the bytes are staged in the pinned
// arena of the next vn_train_batches_begin call, so the uploads that precede a call (extra inputs, BC/IC rows of the next
// MOR batch) queue up behind the steps of the call still running.  *staged = false: plain copy, the caller synchronises.
static const size_t kAuxBytes = 4u << 20;
static cudaError_t h2d_staged(vn_engine* e, void* dst, const void* src, size_t bytes, bool* staged) {
    vn_engine::Pend& p = e->pend[e->pendNext];
    *staged = false;
    if (p.k == 0 && bytes <= kAuxBytes) {
        if (!p.aux) { cudaError_t ce = cudaHostAlloc(reinterpret_cast<void**>(&p.aux), kAuxBytes, cudaHostAllocDefault); if (ce != cudaSuccess) return ce; }
        size_t off = (p.auxOff + 15) & ~(size_t)15;
        if (off + bytes > kAuxBytes) {                  // arena used up without a call in between: everything staged so far must land first
            cudaError_t ce = cudaStreamSynchronize(e->stream);
            if (ce != cudaSuccess) return ce;
            off = 0;
        }
        memcpy(p.aux + off, src, bytes);
        p.auxOff = off + bytes;
        *staged = true;
        return cudaMemcpyAsync(dst, p.aux + off, bytes, cudaMemcpyHostToDevice, e->stream);
    }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream);
}

extern "C" int vn_select_table(vn_engine* e, int32_t slot) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (slot < 0 || slot >= VN_MAX_TABLES) return fail(VN_E_INVALID, "table slot must be in [0,%d)", VN_MAX_TABLES);
    while ((int)e->slots.size() <= slot) e->slots.push_back(new PointSet());
    e->t = e->slots[slot];
    if (e->t->loaded) {                     // default batch of a resident table: all of it, in order
        e->indexed = false;
        e->nb = e->t->nbTab;
        return ensure_work(e);
    }
    e->nb = 0; e->P = 0;
    return VN_OK;
}
extern "C" int vn_table_loaded(const vn_engine* e, int32_t slot) {
    return (e && slot >= 0 && slot < (int)e->slots.size() && e->slots[slot]->loaded) ? 1 : 0;
}
extern "C" int vn_free_table(vn_engine* e, int32_t slot) {
    if (!e || slot < 0 || slot >= (int)e->slots.size()) return fail(VN_E_INVALID, "bad table slot");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    PointSet* t = e->slots[slot];
    drop_graph(t);
    t->cols.release(); t->integW.release(); t->detJ.release(); t->genBuf.release(); t->inKernel = false;
    t->loaded = false; t->rows = t->nbTab = 0;
    if (t == e->t) { e->nb = 0; e->P = 0; }
    return VN_OK;
}
// Mini-batch = the listed test functions of the current table, in that order (the reference gathers
// `integInd[batchInd[n0:n1]]` on the host for every batch, VarNetUtility.py:833-844); NULL = whole table.
extern "C" int vn_set_batch(vn_engine* e, const int32_t* tf_index, int64_t nb) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (!e->t->loaded) return fail(VN_E_STATE, "vn_upload_points must be called first to construct training tables!");
    CK(cudaSetDevice(e->cfg.device));
    if (!tf_index) {
        e->indexed = false; e->nb = e->t->nbTab;
        return ensure_work(e);
    }
    if (nb < 1) return fail(VN_E_INVALID, "a batch needs at least one test function");
    if (e->t->integNum % 4 != 0) return fail(VN_E_UNSUPPORTED, "indexed batches need integNum to be a multiple of 4");
    CK(e->batchIdx.ensure((size_t)nb * sizeof(int32_t)));
    CK(cudaMemcpyAsync(e->batchIdx.p, tf_index, (size_t)nb * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->indexed = true; e->nb = (unsigned int)nb;
    return ensure_work(e);
}
// Constant values of the trailing MLP inputs (MOR parameters): the table then only stores the space-time
// columns, instead of a re-tiled copy per parameter batch (VarNet.py:843-851, VarNetUtility.py:725-729).
extern "C" int vn_set_extra_inputs(vn_engine* e, const float* vals, int32_t n) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (n < 0 || n > e->cfg.inpDim) return fail(VN_E_INVALID, "bad number of extra inputs");
    CK(cudaSetDevice(e->cfg.device));
    CK(e->extraX.ensure(VN_MAX_INPDIM * sizeof(float)));
    if (n > 0) {
        if (!vals) return fail(VN_E_INVALID, "null argument");
        bool staged = false;
        CK(h2d_staged(e, e->extraX.p, vals, n * sizeof(float), &staged));
        if (!staged) CK(cudaStreamSynchronize(e->stream));
    }
    e->nExtra = n;
    return VN_OK;
}
extern "C" int vn_upload_table_f64(vn_engine* e, const double* X, int32_t nx, const double* G, const double* s,
                                   const double* N, const double* dNt, int64_t nb, int32_t integNum, const double* iw,
                                   const double* dj, int32_t djv) {
    return upload_table<double>(e, X, nx, G, s, N, dNt, nb, integNum, iw, dj, djv);
}
extern "C" int vn_upload_table_f32(vn_engine* e, const float* X, int32_t nx, const float* G, const float* s,
                                   const float* N, const float* dNt, int64_t nb, int32_t integNum, const float* iw,
                                   const float* dj, int32_t djv) {
    return upload_table<float>(e, X, nx, G, s, N, dNt, nb, integNum, iw, dj, djv);
}

template <typename T>
static int upload_bic(vn_engine* e, const T* bX, const T* bL, int64_t nbi, int64_t bDof, double biDimVal) {
    if (!e || !bX || !bL) return fail(VN_E_INVALID, "biInput and biLabel are required");
    if (nbi < 1 || bDof < 0 || bDof > nbi) return fail(VN_E_INVALID, "need 0 <= bDof <= nbi and nbi >= 1");
    // the captured step graphs carry the row counts and biDimVal as kernel arguments: only new CONTENT of the same shape (a
    // MOR parameter batch re-uploads its BC/IC rows before every epoch of mini-batches, VarNet.py:843-851) keeps them valid;
    // reallocated buffers are caught by the graphs' allocation epoch
    if ((unsigned int)nbi != e->nbi || (unsigned int)bDof != e->bDof || (float)biDimVal != e->biDimVal) {
        cudaStreamSynchronize(e->stream);               // a replay may still be running (vn_train_batches_begin)
        drop_graph(e);
    }
    const vn_config& c = e->cfg;
    CK(cudaSetDevice(c.device));
    e->bstride = (nbi + kPad - 1) / kPad * kPad;
    e->nbi = (unsigned int)nbi; e->bDof = (unsigned int)bDof; e->biDimVal = (float)biDimVal;
    CK(e->bcols.ensure((size_t)c.inpDim * e->bstride * sizeof(float)));
    CK(cudaMemsetAsync(e->bcols.p, 0, (size_t)c.inpDim * e->bstride * sizeof(float), e->stream));
    CK(e->blabel.ensure((size_t)e->bstride * sizeof(float)));
    CK(e->cj.ensure((size_t)e->bstride * sizeof(float)));
    CK(e->stage.ensure((size_t)nbi * (c.inpDim + 1) * sizeof(T)));
    T* sX = e->stage.as<T>();
    T* sL = sX + nbi * c.inpDim;
    bool stagedX = false, stagedL = false;
    CK(h2d_staged(e, sX, bX, nbi * c.inpDim * sizeof(T), &stagedX));
    CK(h2d_staged(e, sL, bL, nbi * sizeof(T), &stagedL));
    vn_pack_kernel<T><<<(unsigned)((nbi + 255) / 256), 256, 0, e->stream>>>(
        sX, c.inpDim, nullptr, 0, nullptr, nullptr, nullptr, e->bcols.as<float>(), e->bstride, 0, nbi, 0, 0, -1, -1);
    CK(cudaGetLastError());
    vn_cast_kernel<T><<<(unsigned)((nbi + 255) / 256), 256, 0, e->stream>>>(sL, e->blabel.as<float>(), nbi);
    CK(cudaGetLastError());
    e->launches += 2;
    if (!(stagedX && stagedL)) CK(cudaStreamSynchronize(e->stream));      // the caller's arrays are read until then
    if (e->wclass == 256) { e->gridBic = 0; return VN_OK; }
    const long long tiles = e->bstride / e->gBicAdj.TP;
    e->gridBic = (int)std::min<long long>(tiles, e->numSMs);
    CK(e->partBic.ensure((size_t)e->gridBic * e->gBicAdj.pl.psz * sizeof(double)));
    CK(e->part32Bic.ensure((size_t)e->gridBic * e->gBicAdj.pl.psz * sizeof(float)));
    CK(e->stashBic.ensure(std::max<size_t>(16, (size_t)e->gridBic * e->gBicAdj.stashFloats * sizeof(float))));
    return VN_OK;
}

extern "C" int vn_upload_points_f32(vn_engine* e, const float* X, const float* G, const float* s, const float* N,
                                    const float* dNt, int64_t nb, int32_t integNum, const float* iw, const float* dj,
                                    int32_t djv) {
    return upload_points<float>(e, X, G, s, N, dNt, nb, integNum, iw, dj, djv);
}
extern "C" int vn_upload_points_f64(vn_engine* e, const double* X, const double* G, const double* s, const double* N,
                                    const double* dNt, int64_t nb, int32_t integNum, const double* iw, const double* dj,
                                    int32_t djv) {
    return upload_points<double>(e, X, G, s, N, dNt, nb, integNum, iw, dj, djv);
}
extern "C" int vn_upload_bic_f32(vn_engine* e, const float* bX, const float* bL, int64_t nbi, int64_t bDof, float bdv) {
    return upload_bic<float>(e, bX, bL, nbi, bDof, (double)bdv);
}
extern "C" int vn_upload_bic_f64(vn_engine* e, const double* bX, const double* bL, int64_t nbi, int64_t bDof, double bdv) {
    return upload_bic<double>(e, bX, bL, nbi, bDof, bdv);
}

// ------------------------------------------------------------------ hot path
static void base_args(const vn_engine* e, TileArgs* a) {
    memset(a, 0, sizeof(*a));
    a->net = e->net;
    a->theta = e->theta.as<float>();
    a->timeDependent = e->cfg.timeDependent;
    a->isSource = e->cfg.isSource;
    a->dim = e->cfg.dim;
    a->nxTable = e->cfg.inpDim; a->tfIndex = nullptr; a->extraX = nullptr;
    a->wts = e->wts.as<float>();
}
static void var_args(const vn_engine* e, TileArgs* a) {
    base_args(e, a);
    const PointSet* t = e->t;
    a->cols = t->cols.as<float>(); a->pstride = t->pstride;
    a->colX = t->colX; a->colG = t->colG; a->colT = t->colT; a->colS = t->colS;
    a->useGen = t->inKernel ? 1 : 0; a->gen = t->genTab;
    a->nxTable = t->nx; a->extraX = e->extraX.as<float>();
    a->tfIndex = e->indexed ? (e->idxOverride ? e->idxOverride : e->batchIdx.as<int>()) : nullptr;
    a->P = e->P;
    a->integNum = t->integNum;
    a->integW = t->hasIntegW ? t->integW.as<float>() : nullptr;
    a->detJ = t->detJ.as<float>(); a->detJvec = t->detJvec;
    a->R = e->R.as<float>(); a->Iw = e->Iw.as<float>(); a->lossVec = e->lossVec.as<float>();
}
static void bic_args(const vn_engine* e, TileArgs* a) {
    base_args(e, a);
    a->cols = e->bcols.as<float>(); a->pstride = e->bstride;
    a->colX = 0; a->colG = 0; a->colT = -1; a->colS = -1;
    a->nxTable = e->cfg.inpDim; a->tfIndex = nullptr;
    a->P = e->nbi; a->label = e->blabel.as<float>(); a->bDof = e->bDof; a->biDimVal = e->biDimVal;
    a->cj = e->cj.as<float>();
}

// tensor-core class: chunked layer pipeline (vn_tc.cu), FP64 accumulation of the gradient across chunks
static int run_loss_tc(vn_engine* e, bool needGrad, const FedPlan* plan = nullptr) {
    const vn_config& c = e->cfg;
    cudaStream_t st = e->stream;
    const int np = e->net.nparam;
    double* acc = e->tcAcc.as<double>();
    CK(cudaMemsetAsync(acc, 0, (size_t)(np + 2) * sizeof(double), st));
    CK(vn_tc_stage_weights(e->net, e->tcGeom, e->theta.as<float>(), e->tcWork.p, st));
    e->launches++;
    TcJob j;
    j.act = c.act; j.geom = e->tcGeom; j.needGrad = needGrad; j.work = e->tcWork.p; j.g64 = acc; j.lossAcc = acc + np;
    j.err = e->tcErr.as<int>(); j.st = st; j.numSMs = e->numSMs;
    {
        j.S = e->S; j.mode = TC_VAR;
        var_args(e, &j.in);
        size_t pk = 0; long long covered = 0; bool done = false; int prc = 0;
        if (plan)
            j.waitRows = [&](unsigned long long need) -> cudaError_t {      // uploads run ahead on the copy stream, chunk by chunk
                while (!done && covered < (long long)need) {
                    FedPlan::Sub sub;
                    const int pr = plan->produce(pk, &sub);
                    if (pr < 0) { prc = pr; return cudaErrorUnknown; }
                    if (pr == 0) { done = true; break; }
                    const cudaError_t we = cudaStreamWaitEvent(st, sub.ev, 0);
                    if (we != cudaSuccess) return we;
                    covered = ((long long)sub.tile0 + sub.ntiles) * 128; ++pk;
                }
                return cudaSuccess;
            };
        ProfScope ps(e, PK_VAR_ADJ);
        cudaError_t ce = vn_tc_run(j);
        j.waitRows = nullptr;
        if (prc) return prc;
        if (ce != cudaSuccess) return fail(VN_E_CUDA, "tensor-core pipeline (variational term): %s", cudaGetErrorString(ce));
        e->launches += j.launches;
    }
    {
        j.S = 1; j.mode = TC_BIC;
        bic_args(e, &j.in);
        ProfScope ps(e, PK_BIC);
        cudaError_t ce = vn_tc_run(j);
        if (ce != cudaSuccess) return fail(VN_E_CUDA, "tensor-core pipeline (boundary/initial rows): %s", cudaGetErrorString(ce));
        e->launches += j.launches;
    }
    ProfScope ps(e, PK_FINAL);
    if (needGrad) { CK(vn_tc_grad_out(acc, e->gbuf.as<float>(), np, st)); e->launches++; }
    FinalArgs f;
    memset(&f, 0, sizeof(f));
    f.net = e->net;
    f.segSum = acc + np; f.nSeg = 1;
    f.detJ = e->t->detJ.as<float>(); f.detJvec = e->t->detJvec;
    f.cj = e->cj.as<float>(); f.nbi = e->nbi; f.bDof = e->bDof; f.timeDependent = c.timeDependent;
    f.wts = e->wts.as<float>(); f.gbuf = e->gbuf.as<float>(); f.needGrad = 0;
    f.err = e->tcErr.as<int>();
    vn_finalize_kernel<<<1, 128, 0, st>>>(f);         // loss scalars only
    CK(cudaGetLastError());
    e->launches++;
    return VN_OK;
}

static int run_loss(vn_engine* e, bool needGrad, const FedPlan* plan = nullptr, float fuseLr = -1.f) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (!e->P || !e->t->loaded) return fail(VN_E_STATE, "vn_upload_points must be called first to construct training tables!");
    if (!e->nbi) return fail(VN_E_STATE, "vn_upload_bic must be called first to construct training tables!");
    if (e->t->nx + e->nExtra != e->cfg.inpDim)
        return fail(VN_E_STATE, "table holds %d input columns and %d extra inputs are set, the MLP needs %d", e->t->nx, e->nExtra, e->cfg.inpDim);
    const vn_config& c = e->cfg;
    CK(cudaSetDevice(c.device));
    cudaStream_t st = e->stream;
    if (e->wclass == 256) return run_loss_tc(e, needGrad, plan);
    TileArgs a;
    // boundary / initial rows: independent of the variational kernels until the final reduction, so they run on the
    // auxiliary stream next to them (fork here, join before vn_finalize_kernel); sequential when profiling
    auto launch_bic = [&](cudaStream_t bs) -> int {
        TileArgs b;
        bic_args(e, &b);
        const int mode = needGrad ? MODE_BIC_ADJ : MODE_BIC_FWD;
        const TileGeom& g = needGrad ? e->gBicAdj : e->gBicFwd;
        b.ntiles = (int)(e->bstride / g.TP);
        b.part = e->partBic.as<double>(); b.part32 = e->part32Bic.as<float>(); b.psz = g.pl.psz;
        b.stash = e->stashBic.as<float>(); b.stashFloats = g.stashFloats;
        const int grid = needGrad ? e->gridBic : std::min(b.ntiles, 2 * e->numSMs);
        ProfScope ps(e, PK_BIC);
        CK(vn_tile_launch(1, e->wclass, c.act, mode, b, grid, g.smemBytes, bs));
        e->launches++;
        return VN_OK;
    };
    const bool concurrentBic = !e->profOn && e->auxStream != nullptr;
    if (concurrentBic) {
        CK(cudaEventRecord(e->evFork, st));
        CK(cudaStreamWaitEvent(e->auxStream, e->evFork, 0));
        int rc = launch_bic(e->auxStream);
        if (rc) return rc;
        CK(cudaEventRecord(e->evJoin, e->auxStream));
    }
    var_args(e, &a);
    int nSeg = (int)((e->nb + 255) / 256);
    const double* segPtr = e->segSum.as<double>();
    const bool tcFwd = !needGrad && e->fused && e->useTc64 && vn_tc64_forward_only_available();   // loss-only pass on the tensor-core tile kernel
    if ((needGrad && e->fused) || tcFwd) {
        // single pass: forward, in-tile residual reduction, adjoint (MODE_VAR_FUSED)
        const TileGeom& g = e->gVarAdj;
        a.ntiles = (int)(((long long)e->P + g.TP - 1) / g.TP);
        a.part = e->partVar.as<double>(); a.part32 = e->part32Var.as<float>(); a.psz = g.pl.psz;
        a.stash = e->stashVar.as<float>(); a.stashFloats = g.stashFloats;
        a.lossPart = e->lossPart.as<double>();
        const bool tc = e->useTc64, tpp = e->useTpp && needGrad;
        if (tpp) { a.ntiles = (int)(((long long)e->P + 127) / 128); a.psz = e->tppLay.npatch * 32; }
        if (tc) {
            a.ntiles = (int)(((long long)e->P + 127) / 128);
            a.psz = e->tc64Geom.psz; a.stashFloats = e->tc64Geom.stashFloats;
            CK(vn_tc64_stage_weights(e->net, e->theta.as<float>(), e->tc64Img.as<float>(), st));
            e->launches++;
        }
        auto launch_var = [&]() -> cudaError_t {
            if (tpp) return vn_tpp_launch(e->S, c.act, a, e->tppLay, e->gridVar, st);
            if (tc) return vn_tc64_launch(e->S, c.act, a, e->tc64Img.as<float>(), e->tcErr.as<int>(), e->gridVar, e->tc64Geom.smemBytes, st, tcFwd ? 1 : 0);
            return vn_tile_launch(e->S, e->wclass, c.act, MODE_VAR_FUSED, a, e->gridVar, g.smemBytes, st);
        };
        {
        ProfScope ps(e, PK_VAR_ADJ);
        if (plan) {
            // one launch per uploaded chunk, each waiting for its own pack kernel on the copy stream
            for (size_t k = 0;; ++k) {
                FedPlan::Sub sub;
                const int pr = plan->produce(k, &sub);
                if (pr < 0) return pr;
                if (pr == 0) break;
                CK(cudaStreamWaitEvent(st, sub.ev, 0));
                a.tile0 = sub.tile0; a.ntiles = sub.ntiles; a.accumulate = k > 0 ? 1 : 0;
                CK(launch_var());
                e->launches++;
            }
            a.tile0 = 0; a.accumulate = 0;
        } else {
            CK(launch_var());
            e->launches++;
        }
        }
        if (tc && needGrad) {
            CK(vn_tc64_reduce(e->net, e->partVar.as<double>(), e->tc64Geom.psz, e->gridVar, e->tc64Flat.as<double>(), st));
            e->launches++;
        }
        nSeg = e->gridVar * (tc ? e->tc64Geom.lossSlots : (tpp ? 4 : g.NT / 32));
        segPtr = e->lossPart.as<double>();
    } else {
        // 1. forward over all quadrature points -> weighted integrand
        {
            const TileGeom& g = e->gVarFwd;
            a.ntiles = (int)(((long long)e->P + g.TP - 1) / g.TP);
            ProfScope ps(e, PK_VAR_FWD);
            if (plan) {
                // fed step: the forward pass takes every uploaded chunk as soon as it has been packed
                for (size_t k = 0;; ++k) {
                    FedPlan::Sub sub;
                    const int pr = plan->produce(k, &sub);
                    if (pr < 0) return pr;
                    if (pr == 0) break;
                    CK(cudaStreamWaitEvent(st, sub.ev, 0));
                    a.tile0 = sub.tile0; a.ntiles = sub.ntiles;
                    CK(vn_tile_launch(e->S, e->wclass, c.act, MODE_VAR_FWD, a, std::min(a.ntiles, 2 * e->numSMs), g.smemBytes, st));
                    if (k) e->launches++;
                }
                a.tile0 = 0;
            } else {
                CK(vn_tile_launch(e->S, e->wclass, c.act, MODE_VAR_FWD, a, std::min(a.ntiles, 2 * e->numSMs), g.smemBytes, st));
            }
        }
        // 2. per-test-function residuals R_i, lossVec, block partials of the variational loss
        {
            SegArgs s;
            s.Iw = e->Iw.as<float>(); s.nb = e->nb; s.integNum = e->t->integNum; s.detJ = e->t->detJ.as<float>();
            s.detJvec = e->t->detJvec; s.tfIndex = e->indexed ? (e->idxOverride ? e->idxOverride : e->batchIdx.as<int>()) : nullptr; s.R = e->R.as<float>(); s.lossVec = e->lossVec.as<float>();
            s.blockSum = e->segSum.as<double>();
            ProfScope ps(e, PK_SEG);
            vn_segreduce_kernel<<<nSeg, 256, 0, st>>>(s);
            CK(cudaGetLastError());
        }
        e->launches += 2;
        if (needGrad) {
            // 3. adjoint over quadrature points (forward recomputed per tile, seeds from R_i)
            const TileGeom& g = e->gVarAdj;
            a.ntiles = (int)(((long long)e->P + g.TP - 1) / g.TP);
            a.part = e->partVar.as<double>(); a.part32 = e->part32Var.as<float>(); a.psz = g.pl.psz;
            a.stash = e->stashVar.as<float>(); a.stashFloats = g.stashFloats;
            ProfScope ps(e, PK_VAR_ADJ);
            CK(vn_tile_launch(e->S, e->wclass, c.act, MODE_VAR_ADJ, a, e->gridVar, g.smemBytes, st));
            e->launches++;
        }
    }
    // 4. boundary / initial rows (already running on the auxiliary stream unless profiling)
    if (concurrentBic) CK(cudaStreamWaitEvent(st, e->evJoin, 0));
    else { int rc = launch_bic(st); if (rc) return rc; }
    // 5. deterministic cross-CTA reduction + loss scalars
    {
        FinalArgs f;
        memset(&f, 0, sizeof(f));
        f.net = e->net; f.pl = e->gVarAdj.pl;
        f.partVar = e->partVar.as<double>(); f.nVar = e->gridVar;
        if (needGrad && e->fused && e->useTc64) { f.nVar = 0; f.flat = e->tc64Flat.as<double>(); }
        if (needGrad && e->fused && e->useTpp) {          // the reduction kernel sums the patch slabs itself (no separate launch)
            f.nVar = 0; f.slab = e->partVar.as<double>(); f.slabSlot = e->tppSlot.as<int>();
            f.slabStride = e->tppLay.npatch * 32; f.nSlab = e->gridVar;
            f.slotParam = e->tppSlot.as<int>() + e->net.nparam;             // inverse map behind the slots: block = patch
        }
        f.partBic = e->partBic.as<double>(); f.nBic = e->gridBic;
        f.segSum = segPtr; f.nSeg = nSeg;
        f.detJ = e->t->detJ.as<float>(); f.detJvec = e->t->detJvec;
        f.cj = e->cj.as<float>(); f.nbi = e->nbi; f.bDof = e->bDof; f.timeDependent = c.timeDependent;
        f.wts = e->wts.as<float>(); f.gbuf = e->gbuf.as<float>(); f.needGrad = needGrad ? 1 : 0;
        f.err = e->tc64 ? e->tcErr.as<int>() : nullptr;
        if (needGrad && fuseLr >= 0.f) {        // vn_train_step on one GPU: optimizer update and step advance inside the reduction
            f.fuseOpt = c.optimizer == VN_OPT_ADAM ? 1 : 2; f.lr = fuseLr;
            f.theta = e->theta.as<float>(); f.m = e->m.as<float>(); f.v = e->v.as<float>();
            f.step = e->stepbuf.as<long long>(); f.corr = e->corrbuf.as<double>(); f.ticket = e->ticket.as<unsigned int>();
            f.lossRing = e->lossRing.as<float>(); f.ringSize = kLossRing;
        }
        const int nb = needGrad ? (e->net.nparam + 3) / 4 : 0;          // one warp per parameter, 4 warps per block
        ProfScope ps(e, PK_FINAL);
        if (f.slotParam) vn_finalize_kernel<<<e->tppLay.npatch + 1, 1024, 0, st>>>(f);     // one 32-warp block per 32-slot patch
        else vn_finalize_kernel<<<nb + 1, 128, 0, st>>>(f);
        CK(cudaGetLastError());
        e->launches++;
    }
    return VN_OK;
}

static int read_scalars(vn_engine* e, float out[4]) {
    CK(cudaMemcpyAsync(out, e->gbuf.as<float>() + e->net.nparam, 4 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    int tcErr = 0;
    if (e->wclass == 256 || e->tc64) CK(cudaMemcpyAsync(&tcErr, e->tcErr.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (tcErr) {
        cudaMemsetAsync(e->tcErr.p, 0, sizeof(int), e->stream);       // report once; the next call starts clean
        return fail(VN_E_CUDA, "tensor-core pipeline: an mbarrier wait expired (results of this call are invalid)");
    }
    return VN_OK;
}

extern "C" int vn_loss(vn_engine* e, float out[4], float* lossVec) {
    int rc = run_loss(e, false);
    if (rc) return rc;
    if (lossVec) CK(cudaMemcpyAsync(lossVec, e->lossVec.p, (size_t)e->nb * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (out) return read_scalars(e, out);
    if (lossVec) CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_loss_grad(vn_engine* e, float out[4]) {
    int rc = run_loss(e, true);
    if (rc) return rc;
    if (out) return read_scalars(e, out);
    return VN_OK;
}
// Loss + gradient of a step whose point table arrives with the call (the reference feeds every array on every
// sess.run, VarNetUtility.py:1044): same result as vn_upload_points_* followed by vn_loss_grad, but the chunked
// host-to-device copies overlap the step's kernels.  The caller's arrays are no longer read when the call returns.
template <typename T>
static int loss_grad_fed(vn_engine* e, const T* X, const T* G, const T* src, const T* N, const T* dNt, int64_t nb,
                         int32_t integNum, const T* integW, const T* detJ, int32_t detJvec, float out[4]) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    e->nExtra = 0;
    const long long P = (long long)nb * integNum;
    // every kernel family takes the chunks as they arrive: the fused single-pass kernels per chunk, the two-pass class (integNum does
    // not divide the tile) its forward pass per chunk, the tensor-core class its own point chunks once the uploads cover them
    const int fedTP = fed_tile(e, integNum);
    const bool overlap = e->nbi > 0 && P > kChunk && integNum > 0 && fedTP > 0 && kChunk % fedTP == 0 && kChunk / fedTP >= e->numSMs && !e->profOn;
    if (!overlap) {
        int rc = upload_table<T>(e, X, e->cfg.inpDim, G, src, N, dNt, nb, integNum, integW, detJ, detJvec);
        if (rc) return rc;
        rc = run_loss(e, true);
        if (rc) return rc;
        return out ? read_scalars(e, out) : VN_OK;
    }
    FedPlan plan;
    int rc = upload_table<T>(e, X, e->cfg.inpDim, G, src, N, dNt, nb, integNum, integW, detJ, detJvec, &plan);
    if (!rc) rc = run_loss(e, true, &plan);
    cudaError_t ce = cudaStreamSynchronize(e->copyStream);          // every copy out of the caller's arrays has completed
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(VN_E_CUDA, "copy stream: %s", cudaGetErrorString(ce));
    return out ? read_scalars(e, out) : VN_OK;
}
extern "C" int vn_loss_grad_fed_f32(vn_engine* e, const float* X, const float* G, const float* s, const float* N, const float* dNt,
                                    int64_t nb, int32_t integNum, const float* iw, const float* dj, int32_t djv, float out[4]) {
    return loss_grad_fed<float>(e, X, G, s, N, dNt, nb, integNum, iw, dj, djv, out);
}
extern "C" int vn_loss_grad_fed_f64(vn_engine* e, const double* X, const double* G, const double* s, const double* N, const double* dNt,
                                    int64_t nb, int32_t integNum, const double* iw, const double* dj, int32_t djv, float out[4]) {
    return loss_grad_fed<double>(e, X, G, s, N, dNt, nb, integNum, iw, dj, djv, out);
}
extern "C" int vn_grad_buffer(vn_engine* e, void** p, int64_t* n) {
    if (!e || !p || !n) return fail(VN_E_INVALID, "null argument");
    *p = e->gbuf.p; *n = e->net.nparam + 4;
    return VN_OK;
}
extern "C" int vn_get_grad(vn_engine* e, float* grad, int64_t n, float out[4]) {
    if (!e || !grad) return fail(VN_E_INVALID, "null argument");
    if (n != e->net.nparam) return fail(VN_E_INVALID, "parameter count mismatch");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(grad, e->gbuf.p, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (out) return read_scalars(e, out);
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
// ---- multi-GPU: one handle per GPU, one NCCL communicator over the towers.
extern "C" int vn_comm_unique_id(const char* nccl_lib, void* id128) {
    if (!id128) return fail(VN_E_INVALID, "null argument");
    int rc = nccl_load(nccl_lib);
    if (rc) return rc;
    NcclApi::UniqueId id;
    const int nr = g_nccl.GetUniqueId(&id);
    if (nr != 0) return fail(VN_E_CUDA, "ncclGetUniqueId: %s", nccl_err(nr));
    memcpy(id128, &id, sizeof(id));
    return VN_OK;
}
extern "C" int vn_comm_init(vn_engine* e, const char* nccl_lib, const void* id128, int32_t rank, int32_t world) {
    if (!e || !id128) return fail(VN_E_INVALID, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(VN_E_INVALID, "rank %d outside the world of %d towers", rank, world);
    int rc = nccl_load(nccl_lib);
    if (rc) return rc;
    CK(cudaSetDevice(e->cfg.device));
    if (e->comm) { g_nccl.CommDestroy(e->comm); e->comm = nullptr; }
    NcclApi::UniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclApi::Comm c = nullptr;
    const int nr = g_nccl.CommInitRank(&c, world, id, rank);
    if (nr != 0) return fail(VN_E_CUDA, "ncclCommInitRank: %s", nccl_err(nr));
    e->comm = c; e->commRank = rank; e->commWorld = world;
    drop_graph(e);
    return VN_OK;
}
extern "C" int vn_comm_world(const vn_engine* e) { return (e && e->comm) ? e->commWorld : 1; }
// SUM of the gradient buffer [grad | loss, BCloss, ICloss, varLoss] over the towers, on the engine's stream, in place
// (TFNN.sum_grads, TFModel.py:342-377).  Ordered after the kernels of vn_loss_grad and before vn_optimizer_step by stream order.
extern "C" int vn_allreduce_grad(vn_engine* e) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (!e->comm) return fail(VN_E_STATE, "vn_comm_init has not been called on this tower");
    CK(cudaSetDevice(e->cfg.device));
    const int nr = g_nccl.AllReduce(e->gbuf.p, e->gbuf.p, (size_t)e->net.nparam + 4, kNcclFloat32, kNcclSum, e->comm, e->stream);
    if (nr != 0) return fail(VN_E_CUDA, "ncclAllReduce: %s", nccl_err(nr));
    e->launches++;
    return VN_OK;
}
extern "C" int vn_host_register(const void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(VN_E_INVALID, "null argument");
    if (!is_pageable(ptr)) return 1;              // cudaHostAlloc / torch pinned memory / registered by someone else: usable as it is
    const cudaError_t ce = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
    if (ce == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 1; }       // page-locked by its owner: nothing to undo later
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(VN_E_CUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(ce)); }
    return VN_OK;
}
extern "C" int vn_host_unregister(const void* ptr) {
    if (!ptr) return fail(VN_E_INVALID, "null argument");
    const cudaError_t ce = cudaHostUnregister(const_cast<void*>(ptr));
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(VN_E_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(ce)); }
    return VN_OK;
}
extern "C" int vn_get_scalars(vn_engine* e, float out[4]) {
    if (!e || !out) return fail(VN_E_INVALID, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    return read_scalars(e, out);
}
extern "C" int vn_get_lossvec(vn_engine* e, float* lossVec, int64_t nb) {
    if (!e || !lossVec) return fail(VN_E_INVALID, "null argument");
    if (nb != (int64_t)e->nb || !e->lossVec.p) return fail(VN_E_STATE, "lossVec holds %u values (asked for %lld); call vn_loss / vn_loss_grad first", e->nb, (long long)nb);
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(lossVec, e->lossVec.p, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_debug_tc64_timing(int64_t out[16]) {
    long long t[16] = {0};
    const int ok = vn_tc64_read_timing(t);
    for (int i = 0; i < 16; ++i) out[i] = t[i];
    return ok ? VN_OK : VN_E_STATE;
}
extern "C" int vn_check_error(vn_engine* e) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    int tcErr = 0;
    if (e->tcErr.p) CK(cudaMemcpyAsync(&tcErr, e->tcErr.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (tcErr) {
        cudaMemsetAsync(e->tcErr.p, 0, sizeof(int), e->stream);
        return fail(VN_E_CUDA, "tensor-core pipeline: an mbarrier wait expired (results since the last check are invalid)");
    }
    return VN_OK;
}
extern "C" int vn_optimizer_step(vn_engine* e, float lr) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (lr < 0.f) return fail(VN_E_INVALID, "learning rate must be positive!");
    CK(cudaSetDevice(e->cfg.device));
    const int np = e->net.nparam;
    ProfScope ps(e, PK_OPT);
    if (e->cfg.optimizer == VN_OPT_ADAM)
        vn_adam_kernel<<<(np + 255) / 256, 256, 0, e->stream>>>(e->theta.as<float>(), e->m.as<float>(), e->v.as<float>(),
                                                                e->gbuf.as<float>(), np, lr, e->corrbuf.as<double>(), e->tcErr.as<int>());
    else
        vn_rmsprop_kernel<<<(np + 255) / 256, 256, 0, e->stream>>>(e->theta.as<float>(), e->v.as<float>(),
                                                                   e->gbuf.as<float>(), np, lr, e->tcErr.as<int>());
    CK(cudaGetLastError());
    vn_advance_kernel<<<1, 1, 0, e->stream>>>(e->stepbuf.as<long long>(), e->corrbuf.as<double>(), e->tcErr.as<int>(),
                                              e->lossRing.as<float>(), kLossRing, e->gbuf.as<float>() + np);
    CK(cudaGetLastError());
    e->launches += 2;
    return VN_OK;
}
static int train_step_enqueue(vn_engine* e, float lr) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    if (lr < 0.f) return fail(VN_E_INVALID, "learning rate must be positive!");
    CK(cudaSetDevice(e->cfg.device));
    const bool useGraph = e->graphOK && !e->profOn && e->stream != nullptr;
    // resident-tile classes: the optimizer update is applied inside vn_finalize_kernel (one kernel less per step);
    // the tensor-core class and profiling runs keep the separate optimizer kernels
    // With a communicator (vn_comm_init) the gradient buffer is all-reduced between the reduction and the update, on the
    // same stream, and {kernels, ncclAllReduce, optimizer} are captured into the step graph together.
    const bool fuse = e->wclass != 256 && !e->profOn && !e->comm;
    auto step_once = [&]() -> int {
        int rc = run_loss(e, true, nullptr, fuse ? lr : -1.f);
        if (!rc && e->comm) rc = vn_allreduce_grad(e);
        if (!rc && !fuse) rc = vn_optimizer_step(e, lr);
        return rc;
    };
    PointSet* t = e->t;
    // a captured step stays valid while the table, the batch size / kind and lr are unchanged (the index list,
    // the extra inputs and the loss weights are read from device memory at replay time)
    if (useGraph && t->graph && t->graphLr == lr && t->graphNb == e->nb && t->graphIndexed == e->indexed &&
        t->graphEpoch == g_reallocEpoch) {
        CK(cudaGraphLaunch(t->graph, e->stream));
        e->launches += t->graphLaunches;
    } else if (useGraph) {
        // capture the step once per (tables, lr): replay removes the per-kernel launch gaps that dominate
        // the small operator configurations (5 kernels of a few tens of microseconds)
        drop_graph(t);
        if (!e->P || !t->loaded) return fail(VN_E_STATE, "vn_upload_points must be called first to construct training tables!");
        if (!e->nbi) return fail(VN_E_STATE, "vn_upload_bic must be called first to construct training tables!");
        if (t->nx + e->nExtra != e->cfg.inpDim) return fail(VN_E_STATE, "table input columns + extra inputs do not match the MLP input size");
        const int64_t l0 = e->launches;
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
        int rc = VN_OK;
        if (ce == cudaSuccess) {
            rc = step_once();
            ce = cudaStreamEndCapture(e->stream, &g);
        }
        if (ce != cudaSuccess || rc || !g || cudaGraphInstantiate(&t->graph, g, 0) != cudaSuccess) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            t->graph = nullptr; e->graphOK = false;             // fall back to plain launches for this engine
            e->launches = l0;
            rc = step_once();
            if (rc) return rc;
        } else {
            cudaGraphDestroy(g);
            t->graphLr = lr; t->graphLaunches = (int)(e->launches - l0);
            t->graphNb = e->nb; t->graphIndexed = e->indexed; t->graphEpoch = g_reallocEpoch;
            CK(cudaGraphLaunch(t->graph, e->stream));
        }
    } else {
        int rc = step_once();
        if (rc) return rc;
    }
    return VN_OK;
}
// Kernel -> kernel edges of a captured step sequence become PROGRAMMATIC dependencies (vn_pdl.cuh): the downstream grid is scheduled
// while the upstream one still runs and blocks in griddepcontrol.wait until it has completed, which takes the grid launch latency
// (a few microseconds per dependency, two per step) off the critical path of the launch-bound operator configurations.  Only for
// graphs whose kernel nodes are all {tpp_var_kernel, vn_adj_kernel (boundary/initial rows), vn_finalize_kernel}: these three carry
// the wait.  Returns false (graph unchanged or partly converted: the caller discards it) when the runtime refuses.
static bool pdl_edges(cudaGraph_t g) {
    size_t n = 0;
    if (cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &n) != cudaSuccess) { cudaGetLastError(); return false; }
    if (!n) return true;
    std::vector<cudaGraphNode_t> from(n), to(n);
    std::vector<cudaGraphEdgeData> ed(n);
    if (cudaGraphGetEdges_v2(g, from.data(), to.data(), ed.data(), &n) != cudaSuccess) { cudaGetLastError(); return false; }
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType tf, tt;
        if (cudaGraphNodeGetType(from[i], &tf) != cudaSuccess || cudaGraphNodeGetType(to[i], &tt) != cudaSuccess) { cudaGetLastError(); return false; }
        if (tf != cudaGraphNodeTypeKernel || tt != cudaGraphNodeTypeKernel || ed[i].type != cudaGraphDependencyTypeDefault) continue;
        if (cudaGraphRemoveDependencies_v2(g, &from[i], &to[i], &ed[i], 1) != cudaSuccess) { cudaGetLastError(); return false; }
        cudaGraphEdgeData pe;
        memset(&pe, 0, sizeof(pe));
        pe.from_port = cudaGraphKernelNodePortProgrammatic;
        pe.type = cudaGraphDependencyTypeProgrammatic;
        if (cudaGraphAddDependencies_v2(g, &from[i], &to[i], &pe, 1) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    return true;
}
static bool pdl_wanted(const vn_engine* e, bool fuse) {
    static const bool off = [] { const char* v = getenv("VARNET_B200_PDL"); return v && v[0] == '0'; }();
    return !off && fuse && e->useTpp && e->fused && !e->comm && e->auxStream != nullptr;
}

// k consecutive optimizer steps as ONE captured graph: for the launch-bound operator configurations a step is three kernels of a few
// tens of microseconds, and a graph launch per step leaves the GPU idle for a few microseconds between them.  `seqBytes` > 0: step i
// first copies index list i of e->batchSeq into e->batchIdx (vn_train_batches); 0: the same batch k times (vn_train_steps).
// Returns 1 if the steps were enqueued this way, 0 if the caller has to enqueue them one by one, < 0 on error.
static const int kStepsPerGraph = 16;
static int train_steps_one_graph(vn_engine* e, float lr, int k, size_t seqBytes) {
    static const bool off = [] { const char* v = getenv("VARNET_B200_MULTISTEP_GRAPH"); return v && v[0] == '0'; }();
    if (off || k < 2 || !e->graphOK || e->profOn || e->stream == nullptr) return 0;
    PointSet* t = e->t;
    if (!e->P || !t->loaded || !e->nbi || t->nx + e->nExtra != e->cfg.inpDim) return 0;      // let the single-step path report it
    // only small steps: a captured sequence of long steps gains nothing and the capture itself costs k steps of launch work
    if ((double)e->P * e->net.nparam > 4.0e9) return 0;
    const bool seq = seqBytes > 0;
    if (!(t->graphK && t->graphKn == k && t->graphKLr == lr && t->graphKNb == e->nb && t->graphKIndexed == e->indexed &&
          t->graphKSeq == seq && t->graphKEpoch == g_reallocEpoch)) {
        if (t->graphK) { cudaGraphExecDestroy(t->graphK); t->graphK = nullptr; }
        const bool fuse = e->wclass != 256 && !e->comm;
        const int64_t l0 = e->launches;
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
        int rc = VN_OK;
        if (ce == cudaSuccess) {
            // step i reads its index list where the upload put it (no copy between two steps: the step kernels stay adjacent
            // kernel nodes); the engine is left on the last batch by one copy behind the last step
            for (int i = 0; i < k && !rc; ++i) {
                if (seq) e->idxOverride = reinterpret_cast<const int*>(e->batchSeq.as<char>() + seqBytes * i);
                rc = run_loss(e, true, nullptr, fuse ? lr : -1.f);
                if (!rc && e->comm) rc = vn_allreduce_grad(e);
                if (!rc && !fuse) rc = vn_optimizer_step(e, lr);
            }
            e->idxOverride = nullptr;
            if (seq && !rc && cudaMemcpyAsync(e->batchIdx.p, e->batchSeq.as<char>() + seqBytes * (k - 1), seqBytes, cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess) rc = VN_E_CUDA;
            ce = cudaStreamEndCapture(e->stream, &g);
        }
        const int64_t captured = e->launches - l0;
        e->launches = l0;
        if (ce != cudaSuccess || rc || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            t->graphK = nullptr;
            return 0;
        }
        // programmatic kernel -> kernel edges on a clone, so that a refusal leaves the captured graph as it was
        e->pdlGraphs = 0;
        if (pdl_wanted(e, fuse)) {
            cudaGraph_t gp = nullptr;
            if (cudaGraphClone(&gp, g) == cudaSuccess && pdl_edges(gp) && cudaGraphInstantiate(&t->graphK, gp, 0) == cudaSuccess) e->pdlGraphs = 1;
            else { cudaGetLastError(); t->graphK = nullptr; }
            if (gp) cudaGraphDestroy(gp);
        }
        if (!t->graphK && cudaGraphInstantiate(&t->graphK, g, 0) != cudaSuccess) {
            cudaGraphDestroy(g);
            cudaGetLastError();
            t->graphK = nullptr;
            return 0;
        }
        cudaGraphDestroy(g);
        t->graphKn = k; t->graphKLr = lr; t->graphKNb = e->nb; t->graphKIndexed = e->indexed; t->graphKSeq = seq;
        t->graphKEpoch = g_reallocEpoch; t->graphKLaunches = (int)captured;
    }
    if (cudaGraphLaunch(t->graphK, e->stream) != cudaSuccess) { cudaGetLastError(); return 0; }
    e->launches += t->graphKLaunches;
    return 1;
}

extern "C" int vn_train_step(vn_engine* e, float lr, float* loss_out) {
    int rc = train_step_enqueue(e, lr);
    if (rc) return rc;
    if (loss_out) {
        // the loss read is also where an expired tensor-core barrier wait of this (or an earlier unfetched) step surfaces:
        // the reduction kernel published NaN and skipped the optimizer update, here the caller gets VN_E_CUDA
        float sc[4];
        rc = read_scalars(e, sc);
        *loss_out = sc[0];
        if (rc) return rc;
    }
    return VN_OK;
}
// k optimizer steps on the current batch back to back, one host round trip: the captured step graph is replayed k times and
// the k losses come back together (the reference reads the loss of every step, VarNetUtility.py:1044, but only acts on it
// per epoch: tolerance test VarNet.py:1378, bookkeeping every saveFreq epochs).  For the launch-bound operator configurations
// this removes the per-step synchronisation and readback.  k <= 4096.
extern "C" int vn_train_steps(vn_engine* e, float lr, int32_t k, float* losses) {
    if (!e || !losses) return fail(VN_E_INVALID, "null argument");
    if (k < 1 || k > kLossRing) return fail(VN_E_INVALID, "k must be in [1, %d]", kLossRing);
    CK(cudaSetDevice(e->cfg.device));
    long long step0 = 0;
    CK(cudaMemcpyAsync(&step0, e->stepbuf.p, sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    // chunks of kStepsPerGraph steps go out as one graph each (the caller's k varies from call to call: VarNet.train never crosses a
    // saveFreq boundary), the remainder step by step
    int done = 0;
    while (k - done >= kStepsPerGraph) {
        const int og = train_steps_one_graph(e, lr, kStepsPerGraph, 0);
        if (og < 0) return og;
        if (!og) break;
        done += kStepsPerGraph;
    }
    for (int i = done; i < k; ++i) {
        int rc = train_step_enqueue(e, lr);
        if (rc) return rc;
    }
    const int a = (int)(step0 % kLossRing), n1 = std::min(k, kLossRing - a);
    CK(cudaMemcpyAsync(losses, e->lossRing.as<float>() + a, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (k > n1) CK(cudaMemcpyAsync(losses + n1, e->lossRing.as<float>(), (size_t)(k - n1) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    float sc[4];
    return read_scalars(e, sc);              // synchronises; reports an expired tensor-core barrier wait
}

// One optimizer step per mini-batch for k mini-batches of the current table, one host round trip: row i of `tf_index`
// ([k][nb], host) is the index list of step i (what k x {vn_set_batch, vn_train_step} would do: ManageTrainData.optimIter,
// VarNetUtility.py:1021-1047, one sess.run per mini-batch).  The k lists are uploaded once; between two replays of the captured
// step graph only a device-to-device copy of the next list runs on the engine's stream.  The engine is left on the last batch.
extern "C" int vn_train_batches_begin(vn_engine* e, float lr, const int32_t* tf_index, int64_t nb, int32_t k) {
    if (!e || !tf_index) return fail(VN_E_INVALID, "null argument");
    if (k < 1 || k > kLossRing) return fail(VN_E_INVALID, "k must be in [1, %d]", kLossRing);
    if (nb < 1) return fail(VN_E_INVALID, "a batch needs at least one test function");
    vn_engine::Pend& p = e->pend[e->pendNext];
    if (p.k) return fail(VN_E_STATE, "two vn_train_batches_begin calls are in flight: vn_train_batches_end must collect the older one first");
    if (!e->t->loaded) return fail(VN_E_STATE, "vn_upload_points must be called first to construct training tables!");
    if (e->t->integNum % 4 != 0) return fail(VN_E_UNSUPPORTED, "indexed batches need integNum to be a multiple of 4");
    CK(cudaSetDevice(e->cfg.device));
    const size_t one = (size_t)nb * sizeof(int32_t);
    if (one * k > p.seqBytes) {                          // this slot's previous call was collected: nothing reads the old buffer
        if (p.seq) { cudaFreeHost(p.seq); p.seq = nullptr; p.seqBytes = 0; }
        CK(cudaHostAlloc(&p.seq, one * k, cudaHostAllocDefault));
        p.seqBytes = one * k;
    }
    if (!p.loss) CK(cudaHostAlloc(reinterpret_cast<void**>(&p.loss), kLossRing * sizeof(float), cudaHostAllocDefault));
    if (!p.err) CK(cudaHostAlloc(reinterpret_cast<void**>(&p.err), sizeof(int), cudaHostAllocDefault));
    if (!p.ev) CK(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
    CK(e->batchSeq.ensure(one * k));
    CK(e->batchIdx.ensure(one));
    CK(e->lossOut.ensure(kLossRing * sizeof(float)));
    memcpy(p.seq, tf_index, one * k);
    CK(cudaMemcpyAsync(e->batchSeq.p, p.seq, one * k, cudaMemcpyHostToDevice, e->stream));
    e->indexed = true; e->nb = (unsigned int)nb;
    int rc = ensure_work(e);
    if (rc) return rc;
    const int og = train_steps_one_graph(e, lr, k, one);
    if (og < 0) return og;
    for (int i = 0; i < k && !og; ++i) {
        CK(cudaMemcpyAsync(e->batchIdx.p, e->batchSeq.as<char>() + one * i, one, cudaMemcpyDeviceToDevice, e->stream));
        rc = train_step_enqueue(e, lr);
        if (rc) return rc;
    }
    vn_ring_gather_kernel<<<1, 256, 0, e->stream>>>(e->lossRing.as<float>(), kLossRing, e->stepbuf.as<long long>(), k, e->lossOut.as<float>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p.loss, e->lossOut.p, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    *p.err = 0;
    if (e->wclass == 256 || e->tc64) CK(cudaMemcpyAsync(p.err, e->tcErr.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaEventRecord(p.ev, e->stream));
    p.k = k;
    e->pendNext ^= 1;
    return VN_OK;
}
// collects the OLDER of the calls in flight
extern "C" int vn_train_batches_end(vn_engine* e, float* losses, int32_t k) {
    if (!e || !losses) return fail(VN_E_INVALID, "null argument");
    const int slot = e->pend[e->pendNext].k ? e->pendNext : (e->pendNext ^ 1);
    vn_engine::Pend& p = e->pend[slot];
    if (!p.k) return fail(VN_E_STATE, "no vn_train_batches_begin call is in flight");
    if (k != p.k) return fail(VN_E_INVALID, "the call in flight has %d mini-batches", p.k);
    CK(cudaSetDevice(e->cfg.device));
    p.k = 0;
    CK(cudaEventSynchronize(p.ev));
    p.auxOff = 0;                                        // everything staged for this call has landed
    memcpy(losses, p.loss, (size_t)k * sizeof(float));
    if (*p.err) {
        cudaMemsetAsync(e->tcErr.p, 0, sizeof(int), e->stream);       // report once; the next call starts clean
        return fail(VN_E_CUDA, "tensor-core pipeline: an mbarrier wait expired (results of this call are invalid)");
    }
    return VN_OK;
}
extern "C" int vn_train_batches(vn_engine* e, float lr, const int32_t* tf_index, int64_t nb, int32_t k, float* losses) {
    if (!e || !losses) return fail(VN_E_INVALID, "null argument");
    if (e->pend[0].k || e->pend[1].k) return fail(VN_E_STATE, "vn_train_batches_end must collect the calls in flight first");
    int rc = vn_train_batches_begin(e, lr, tf_index, nb, k);
    if (rc) return rc;
    return vn_train_batches_end(e, losses, k);
}

// ------------------------------------------------------------------ evaluation
template <typename T>
static int eval_impl(vn_engine* e, const T* X, int64_t n, float* u) {
    if (!e || !X || !u) return fail(VN_E_INVALID, "null argument");
    if (n < 1) return fail(VN_E_INVALID, "need at least one evaluation point");
    if (n >= (1ll << 31) - kPad) return fail(VN_E_UNSUPPORTED, "too many evaluation points in one call");
    const vn_config& c = e->cfg;
    CK(cudaSetDevice(c.device));
    const long long stride = (n + kPad - 1) / kPad * kPad;
    CK(e->evalCols.ensure((size_t)c.inpDim * stride * sizeof(float)));
    CK(e->evalOut.ensure((size_t)stride * sizeof(float)));
    CK(cudaMemsetAsync(e->evalCols.p, 0, (size_t)c.inpDim * stride * sizeof(float), e->stream));
    CK(e->stage.ensure((size_t)n * c.inpDim * sizeof(T)));
    CK(cudaMemcpyAsync(e->stage.p, X, (size_t)n * c.inpDim * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    vn_pack_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(
        e->stage.as<T>(), c.inpDim, nullptr, 0, nullptr, nullptr, nullptr, e->evalCols.as<float>(), stride, 0, n, 0, 0, -1, -1);
    CK(cudaGetLastError());
    TileArgs a;
    base_args(e, &a);
    a.cols = e->evalCols.as<float>(); a.pstride = stride; a.colX = 0; a.colT = -1; a.colS = -1;
    a.P = (unsigned int)n; a.uout = e->evalOut.as<float>();
    if (e->wclass == 256) {
        CK(vn_tc_stage_weights(e->net, e->tcGeom, e->theta.as<float>(), e->tcWork.p, e->stream));
        TcJob j;
        j.S = 1; j.act = c.act; j.mode = TC_EVAL; j.geom = e->tcGeom; j.in = a; j.needGrad = false; j.work = e->tcWork.p;
        j.g64 = e->tcAcc.as<double>(); j.lossAcc = j.g64 + e->net.nparam; j.err = e->tcErr.as<int>(); j.st = e->stream; j.numSMs = e->numSMs;
        cudaError_t ce = vn_tc_run(j);
        if (ce != cudaSuccess) return fail(VN_E_CUDA, "tensor-core pipeline (evaluation): %s", cudaGetErrorString(ce));
        e->launches += j.launches + 2;
    } else {
    const TileGeom& g = e->gEval;
    a.ntiles = (int)(stride / g.TP);
    CK(vn_tile_launch(1, e->wclass, c.act, MODE_EVAL, a, std::min(a.ntiles, 2 * e->numSMs), g.smemBytes, e->stream));
    e->launches += 2;
    }
    CK(cudaMemcpyAsync(u, e->evalOut.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}
extern "C" int vn_eval_f32(vn_engine* e, const float* X, int64_t n, float* u) { return eval_impl<float>(e, X, n, u); }
extern "C" int vn_eval_f64(vn_engine* e, const double* X, int64_t n, float* u) { return eval_impl<double>(e, X, n, u); }

extern "C" int vn_residual_f64(vn_engine* e, const double* X, const double* diff, const double* vel,
                               const double* diff_dx, const double* source, int64_t n, float* u, float* res) {
    if (!e || !X || !diff || !vel || !diff_dx || !source || !res) return fail(VN_E_INVALID, "null argument");
    if (n < 1) return fail(VN_E_INVALID, "need at least one evaluation point");
    if (n >= (1ll << 31) - kPad) return fail(VN_E_UNSUPPORTED, "too many evaluation points in one call");
    if (!e->resOK) return fail(VN_E_UNSUPPORTED, "strong-form residual kernel does not fit shared memory for this network depth");
    const vn_config& c = e->cfg;
    CK(cudaSetDevice(c.device));
    const int ncol = c.inpDim + 2 + 2 * c.dim;
    const long long stride = (n + kPad - 1) / kPad * kPad;
    CK(e->evalCols.ensure((size_t)ncol * stride * sizeof(float)));
    CK(e->evalOut.ensure((size_t)2 * stride * sizeof(float)));
    CK(cudaMemsetAsync(e->evalCols.p, 0, (size_t)ncol * stride * sizeof(float), e->stream));
    const size_t rowVals = (size_t)c.inpDim + 2 + 2 * c.dim;
    CK(e->stage.ensure((size_t)n * rowVals * sizeof(double)));
    double* sX = e->stage.as<double>();
    double* sD = sX + n * c.inpDim;
    double* sV = sD + n;
    double* sG = sV + n * c.dim;
    double* sS = sG + n * c.dim;
    CK(cudaMemcpyAsync(sX, X, (size_t)n * c.inpDim * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sD, diff, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sV, vel, (size_t)n * c.dim * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sG, diff_dx, (size_t)n * c.dim * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(sS, source, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    vn_pack_res_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(sX, c.inpDim, sD, sV, sG, sS, c.dim,
                                                                                  e->evalCols.as<float>(), stride, n);
    CK(cudaGetLastError());
    TileArgs a;
    base_args(e, &a);
    a.cols = e->evalCols.as<float>(); a.pstride = stride;
    a.colX = 0; a.colD = c.inpDim; a.colG = c.inpDim + 1; a.colDD = c.inpDim + 1 + c.dim; a.colS = c.inpDim + 1 + 2 * c.dim;
    a.colT = -1; a.dim = c.dim;
    a.P = (unsigned int)n;
    a.uout = e->evalOut.as<float>(); a.Iw = e->evalOut.as<float>() + stride;
    if (e->wclass == 256) {
        long long n_l = 0;
        cudaError_t ce = vn_tc_residual(a, c.act, e->tcGeom, e->tcWork.p, e->stream, &n_l);
        if (ce != cudaSuccess) return fail(VN_E_CUDA, "strong-form residual (wide networks): %s", cudaGetErrorString(ce));
        e->launches += n_l + 1;
    } else {
    const TileGeom& g = e->gRes;
    a.ntiles = (int)(stride / g.TP);
    CK(vn_tile_launch(VN_S_RES, e->wclass, c.act, MODE_RESIDUAL, a, std::min(a.ntiles, 2 * e->numSMs), g.smemBytes, e->stream));
    e->launches += 2;
    }
    if (u) CK(cudaMemcpyAsync(u, e->evalOut.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(res, e->evalOut.as<float>() + stride, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return VN_OK;
}

extern "C" int vn_profile_enable(vn_engine* e, int on) {
    if (!e) return fail(VN_E_INVALID, "null engine");
    e->profOn = on != 0;
    return VN_OK;
}
// Accumulated device time (ms) and launch counts per kernel slot since the last read:
// 0 var forward, 1 segmented reduce, 2 var adjoint, 3 boundary/initial, 4 finalize, 5 optimizer.
extern "C" int vn_profile_read(vn_engine* e, double ms[VN_PROF_SLOTS], int64_t counts[VN_PROF_SLOTS]) {
    if (!e || !ms || !counts) return fail(VN_E_INVALID, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < VN_PROF_SLOTS; ++k) {
        for (auto& pr : e->profPending[k]) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) { e->profMs[k] += t; e->profCnt[k]++; }
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        e->profPending[k].clear();
        ms[k] = e->profMs[k]; counts[k] = e->profCnt[k];
        e->profMs[k] = 0.0; e->profCnt[k] = 0;
    }
    return VN_OK;
}

extern "C" int vn_kernel_info(const vn_engine* e, char* buf, size_t n) {
    if (!e || !buf) return fail(VN_E_INVALID, "null argument");
    if (e->wclass == 256) {
        snprintf(buf, n, "family=tcgen05-3xtf32 class=256 S=%d L=%d WP=%d chunk=%u points gemm(tile=128x128,smem=%zu) gw(smem=%zu) "
                 "workspace=%zuB nparam=%d SMs=%d", e->S, e->net.L, e->tcGeom.WP, e->tcGeom.capPts, e->tcGeom.smemGemm,
                 e->tcGeom.smemGw, e->tcGeom.workBytes, e->net.nparam, e->numSMs);
        return VN_OK;
    }
    if (e->useTc64) {
        snprintf(buf, n, "family=tcgen05-3xtf32-tile64 class=64 S=%d L=%d var_adj(TP=128,NT=512,smem=%zu,grid=%d,fused-R single pass,"
                 "stash=%lldB/CTA,A-from-TMEM%s) bic_adj(fp32-fma-tile,TP=%d,smem=%zu,grid=%d) nparam=%d SMs=%d",
                 e->S, e->net.L, e->tc64Geom.smemBytes, e->gridVar, (long long)(e->tc64Geom.stashFloats * 4),
                 (e->t && e->t->inKernel) ? ",table=in-kernel generation" : "", e->gBicAdj.TP,
                 e->gBicAdj.smemBytes, e->gridBic, e->net.nparam, e->numSMs);
        return VN_OK;
    }
    if (e->useTpp) {
        snprintf(buf, n, "family=fp32-thread-per-point class=%d S=%d L=%d var_adj(TP=128,NT=128,smem=%zu,grid=%d,%d CTAs/SM,fused-R single pass,"
                 "%d gradient patches) var_fwd(fp32-fma-tile,TP=%d) bic_adj(fp32-fma-tile,TP=%d,smem=%zu,grid=%d) nparam=%d SMs=%d pdl=%d",
                 e->wclass, e->S, e->net.L, e->tppLay.smemBytes, e->gridVar, e->tppCtas, e->tppLay.npatch, e->gVarFwd.TP, e->gBicAdj.TP,
                 e->gBicAdj.smemBytes, e->gridBic, e->net.nparam, e->numSMs, e->pdlGraphs);
        return VN_OK;
    }
    snprintf(buf, n,
             "family=fp32-fma-tile class=%d S=%d L=%d var_fwd(TP=%d,NT=%d,smem=%zu) var_adj(TP=%d,NT=%d,smem=%zu,grid=%d,%s,"
             "stash=%lldB/CTA) bic_adj(TP=%d,smem=%zu,grid=%d) nparam=%d SMs=%d",
             e->wclass, e->S, e->net.L, e->gVarFwd.TP, e->gVarFwd.NT, e->gVarFwd.smemBytes, e->gVarAdj.TP, e->gVarAdj.NT,
             e->gVarAdj.smemBytes, e->gridVar, e->fused ? "fused-R single pass" : "two-pass",
             (long long)(e->gVarAdj.stashFloats * 4), e->gBicAdj.TP, e->gBicAdj.smemBytes, e->gridBic, e->net.nparam, e->numSMs);
    return VN_OK;
}
extern "C" int64_t vn_launch_count(const vn_engine* e) { return e ? e->launches : 0; }
