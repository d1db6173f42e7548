// kern_stats.cu -- instantiations of the mu +- 3 sigma statistics kernel.
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int NW>
KernelFn pick_g(int group) {
#define LSQ_S(G_) lsq_stats_kernel<T, NW, G_, kThreads, unroll_for_stats(kUnrollStats, NW, G_), kLd, kMinBlocksStats>
    return group == 32 ? LSQ_S(32) : LSQ_S(kThreads);
#undef LSQ_S
}
template <typename T>
KernelFn pick(int nw, int group) {
    switch (nw) {
        case 8: return pick_g<T, 8>(group);
        case 4: return pick_g<T, 4>(group);
        case 2: return pick_g<T, 2>(group);
        default: return pick_g<T, 0>(group);
    }
}
}  // namespace
KernelFn get_stats_kernel(int xdtype, int nw, int group) {
    if (xdtype == DT_F32) return pick<float>(nw, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16>(nw, group);
    return pick<__half>(nw, group);
}
}  // namespace lsqb200
