// kern_stats.cu -- instantiations of the mu +- 3 sigma statistics kernel.
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T>
KernelFn pick(bool vec, int group) {
    constexpr int V = ElemTraits<T>::VEC;
#define LSQ_S(VEC_, G_) lsq_stats_kernel<T, VEC_, G_, kThreads, kUnrollStats, kLd, kMinBlocksFwd>
    if (group == 32) return vec ? LSQ_S(V, 32) : LSQ_S(1, 32);
    return vec ? LSQ_S(V, kThreads) : LSQ_S(1, kThreads);
#undef LSQ_S
}
}  // namespace
KernelFn get_stats_kernel(int xdtype, bool vec, int group) {
    if (xdtype == DT_F32) return pick<float>(vec, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16>(vec, group);
    return pick<__half>(vec, group);
}
}  // namespace lsqb200
