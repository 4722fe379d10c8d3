// vn_inst.cuh — instantiates the tile kernels of one kernel class.
// Included by vn_inst_*.cu after defining
//   VN_CLS      class id (suffix of the exported entry points)
//   VN_W        hidden-width class
//   VN_TP_ADJ   points per tile of the forward+adjoint kernels
//   VN_TP_FWD   points per tile of the forward-only kernels
//   VN_TP_RES   points per tile of the strong-form residual kernel (6 streams)
//   VN_TN       neurons per thread
#include "vn_dispatch.h"

namespace {

constexpr bool is_adj(int mode) { return mode == MODE_VAR_ADJ || mode == MODE_BIC_ADJ || mode == MODE_VAR_FUSED; }

template <int S, int ACT, int MODE> struct CfgOf {
    using type = TileCfg<S, VN_W, is_adj(MODE) ? VN_TP_ADJ : (MODE == MODE_RESIDUAL ? VN_TP_RES : VN_TP_FWD), VN_TN, ACT>;
};

template <int S, int ACT, int MODE> bool geom(int L, TileGeom* g) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    g->TP = C::TP; g->NT = C::NT;
    g->smemBytes = tile_smem_floats<C>(L, is_adj(MODE)) * sizeof(float);
    g->stashFloats = is_adj(MODE) ? tile_stash_floats<C>(L) : 0;
    g->pl = make_part_layout<C>(L);
    return true;
}
template <int S, int ACT, int MODE> cudaError_t launch(const TileArgs& a, int grid, size_t smem, cudaStream_t st) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    if constexpr (is_adj(MODE)) vn_adj_kernel<C, MODE><<<grid, C::NT, smem, st>>>(a);
    else vn_fwd_kernel<C, MODE><<<grid, C::NT, smem, st>>>(a);
    return cudaGetLastError();
}
template <int S, int ACT, int MODE> cudaError_t prepare(size_t smem) {
    using C = typename CfgOf<S, ACT, MODE>::type;
    if constexpr (is_adj(MODE))
        return cudaFuncSetAttribute(vn_adj_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else
        return cudaFuncSetAttribute(vn_fwd_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// (S, act, mode) -> template instance
#define VN_CASE(SS, MM, OP, ...)                                                                           \
    if (S == SS && mode == MM) {                                                                          \
        if (act == VN_SIGMOID) return OP<SS, VN_SIGMOID, MM>(__VA_ARGS__);                                \
        else return OP<SS, VN_TANH, MM>(__VA_ARGS__);                                                     \
    }
#define VN_FOR_EACH(OP, ...)                                                                              \
    VN_CASE(1, MODE_EVAL, OP, __VA_ARGS__)                                                                \
    VN_CASE(1, MODE_BIC_FWD, OP, __VA_ARGS__)                                                             \
    VN_CASE(1, MODE_BIC_ADJ, OP, __VA_ARGS__)                                                             \
    VN_CASE(2, MODE_VAR_FWD, OP, __VA_ARGS__)                                                             \
    VN_CASE(2, MODE_VAR_ADJ, OP, __VA_ARGS__)                                                             \
    VN_CASE(2, MODE_VAR_FUSED, OP, __VA_ARGS__)                                                           \
    VN_CASE(3, MODE_VAR_FWD, OP, __VA_ARGS__)                                                             \
    VN_CASE(3, MODE_VAR_ADJ, OP, __VA_ARGS__)                                                             \
    VN_CASE(3, MODE_VAR_FUSED, OP, __VA_ARGS__)                                                           \
    VN_CASE(VN_S_RES, MODE_RESIDUAL, OP, __VA_ARGS__)

}  // namespace

#define VN_CAT2(a, b) a##b
#define VN_CAT(a, b) VN_CAT2(a, b)

bool VN_CAT(vn_geom_c, VN_CLS)(int S, int act, int mode, int L, TileGeom* g) {
    VN_FOR_EACH(geom, L, g)
    return false;
}
cudaError_t VN_CAT(vn_launch_c, VN_CLS)(int S, int act, int mode, const TileArgs& a, int grid, size_t smem,
                                        cudaStream_t st) {
    VN_FOR_EACH(launch, a, grid, smem, st)
    return cudaErrorInvalidValue;
}
cudaError_t VN_CAT(vn_prepare_c, VN_CLS)(int S, int act, int mode, size_t smem) {
    VN_FOR_EACH(prepare, smem)
    return cudaErrorInvalidValue;
}
