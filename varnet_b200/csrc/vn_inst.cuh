// vn_inst.cuh — instantiates the tile kernels of one width class.
// Included by vn_inst_w32.cu / vn_inst_w64.cu after defining
//   VN_W        hidden-width class
//   VN_TP_ADJ   points per tile of the adjoint kernels (all layers resident in smem)
//   VN_TP_FWD   points per tile of the forward-only kernels (ping-pong buffers)
//   VN_TN       neurons per thread
#include "vn_dispatch.h"

namespace {

template <int S, int ACT, int MODE> struct CfgOf {
    static constexpr bool adj = (MODE == MODE_VAR_ADJ || MODE == MODE_BIC_ADJ);
    using type = TileCfg<S, VN_W, adj ? VN_TP_ADJ : VN_TP_FWD, VN_TN, ACT>;
};

template <int S, int ACT, int MODE> bool geom(int L, TileGeom* g) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    constexpr bool adj = CfgOf<S, ACT, MODE>::adj;
    g->TP = C::TP; g->NT = C::NT;
    g->smemBytes = tile_smem_floats<C>(L, adj) * sizeof(float);
    g->pl = make_part_layout<C>(L);
    return true;
}
template <int S, int ACT, int MODE> cudaError_t launch(const TileArgs& a, int grid, size_t smem, cudaStream_t st) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    vn_tile_kernel<C, MODE><<<grid, C::NT, smem, st>>>(a);
    return cudaGetLastError();
}
template <int S, int ACT, int MODE> cudaError_t prepare(size_t smem) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    return cudaFuncSetAttribute(vn_tile_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// (S, act, mode) -> template instance
#define VN_FOR_EACH(OP, ...)                                                                                \
    if (S == 1 && mode == MODE_EVAL)    { if (act == VN_SIGMOID) return OP<1, VN_SIGMOID, MODE_EVAL>(__VA_ARGS__);    else return OP<1, VN_TANH, MODE_EVAL>(__VA_ARGS__); }    \
    if (S == 1 && mode == MODE_BIC_FWD) { if (act == VN_SIGMOID) return OP<1, VN_SIGMOID, MODE_BIC_FWD>(__VA_ARGS__); else return OP<1, VN_TANH, MODE_BIC_FWD>(__VA_ARGS__); } \
    if (S == 1 && mode == MODE_BIC_ADJ) { if (act == VN_SIGMOID) return OP<1, VN_SIGMOID, MODE_BIC_ADJ>(__VA_ARGS__); else return OP<1, VN_TANH, MODE_BIC_ADJ>(__VA_ARGS__); } \
    if (S == 2 && mode == MODE_VAR_FWD) { if (act == VN_SIGMOID) return OP<2, VN_SIGMOID, MODE_VAR_FWD>(__VA_ARGS__); else return OP<2, VN_TANH, MODE_VAR_FWD>(__VA_ARGS__); } \
    if (S == 2 && mode == MODE_VAR_ADJ) { if (act == VN_SIGMOID) return OP<2, VN_SIGMOID, MODE_VAR_ADJ>(__VA_ARGS__); else return OP<2, VN_TANH, MODE_VAR_ADJ>(__VA_ARGS__); } \
    if (S == 3 && mode == MODE_VAR_FWD) { if (act == VN_SIGMOID) return OP<3, VN_SIGMOID, MODE_VAR_FWD>(__VA_ARGS__); else return OP<3, VN_TANH, MODE_VAR_FWD>(__VA_ARGS__); } \
    if (S == 3 && mode == MODE_VAR_ADJ) { if (act == VN_SIGMOID) return OP<3, VN_SIGMOID, MODE_VAR_ADJ>(__VA_ARGS__); else return OP<3, VN_TANH, MODE_VAR_ADJ>(__VA_ARGS__); }

}  // namespace

#define VN_CAT2(a, b) a##b
#define VN_CAT(a, b) VN_CAT2(a, b)

bool VN_CAT(vn_geom_w, VN_W)(int S, int act, int mode, int L, TileGeom* g) {
    VN_FOR_EACH(geom, L, g)
    return false;
}
cudaError_t VN_CAT(vn_launch_w, VN_W)(int S, int act, int mode, const TileArgs& a, int grid, size_t smem,
                                      cudaStream_t st) {
    VN_FOR_EACH(launch, a, grid, smem, st)
    return cudaErrorInvalidValue;
}
cudaError_t VN_CAT(vn_prepare_w, VN_W)(int S, int act, int mode, size_t smem) {
    VN_FOR_EACH(prepare, smem)
    return cudaErrorInvalidValue;
}
