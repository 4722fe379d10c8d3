// vn_dispatch.h — table of compiled tile-kernel families (one translation unit per width class).
#pragma once
#include "vn_tile.cuh"

struct TileGeom {
    int TP, NT;             // points per tile, threads per CTA
    size_t smemBytes;       // dynamic shared memory for this (L, adj)
    long long stashFloats;  // per-CTA global activation stash (adjoint kernels)
    PartLayout pl;          // FP64 partial-slab layout (adjoint kernels)
};

// Kernel classes compiled in: 16 (hidden width <= 16, 256-point tiles), 32 (hidden width <= 32), 64 (width <= 64, up to 4 hidden layers resident),
// 164 (width <= 64, deep networks: smaller tiles).  Returns false when (S, cls, act, mode) has no kernel.
bool vn_tile_geometry(int S, int cls, int act, int mode, int L, TileGeom* g);
cudaError_t vn_tile_launch(int S, int cls, int act, int mode, const TileArgs& a, int grid,
                           size_t smemBytes, cudaStream_t st);
cudaError_t vn_tile_prepare(int S, int cls, int act, int mode, size_t smemBytes);   // opt-in smem attribute

// per-class entry points (defined in vn_inst_*.cu)
#define VN_DECL_CLASS(W)                                                                                   \
    bool vn_geom_c##W(int S, int act, int mode, int L, TileGeom* g);                                        \
    cudaError_t vn_launch_c##W(int S, int act, int mode, const TileArgs& a, int grid, size_t smem,         \
                               cudaStream_t st);                                                            \
    cudaError_t vn_prepare_c##W(int S, int act, int mode, size_t smem);
VN_DECL_CLASS(16)
VN_DECL_CLASS(32)
VN_DECL_CLASS(64)
VN_DECL_CLASS(164)
