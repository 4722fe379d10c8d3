// vn_dispatch.h — table of compiled tile-kernel families (one translation unit per width class).
#pragma once
#include "vn_tile.cuh"

struct TileGeom {
    int TP, NT;             // points per tile, threads per CTA
    size_t smemBytes;       // dynamic shared memory for this (L, adj)
    PartLayout pl;          // FP64 partial-slab layout (adjoint kernels)
};

// Width classes compiled in. A network uses the smallest class >= max hidden width.
// Returns false when (S, wclass, act, mode) has no compiled kernel.
bool vn_tile_geometry(int S, int wclass, int act, int mode, int L, TileGeom* g);
cudaError_t vn_tile_launch(int S, int wclass, int act, int mode, const TileArgs& a, int grid,
                           size_t smemBytes, cudaStream_t st);
cudaError_t vn_tile_prepare(int S, int wclass, int act, int mode, size_t smemBytes);   // opt-in smem attribute

// per-class entry points (defined in vn_inst_w*.cu)
#define VN_DECL_CLASS(W)                                                                                   \
    bool vn_geom_w##W(int S, int act, int mode, int L, TileGeom* g);                                        \
    cudaError_t vn_launch_w##W(int S, int act, int mode, const TileArgs& a, int grid, size_t smem,         \
                               cudaStream_t st);                                                            \
    cudaError_t vn_prepare_w##W(int S, int act, int mode, size_t smem);
VN_DECL_CLASS(32)
VN_DECL_CLASS(64)
