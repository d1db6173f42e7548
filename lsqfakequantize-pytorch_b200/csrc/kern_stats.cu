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
namespace {
template <int U, int MB>
KernelFn rowstats(int xdtype) {
    if (xdtype == DT_F32) return lsq_rowstats_kernel<float, kThreads, U, kLd, MB>;
    if (xdtype == DT_BF16) return lsq_rowstats_kernel<__nv_bfloat16, kThreads, U, kLd, MB>;
    return lsq_rowstats_kernel<__half, kThreads, U, kLd, MB>;
}
}  // namespace
// variant (Tuning::rowstats): 1 = four units in flight per lane, 4 CTAs/SM; 2 = two units, 6 CTAs/SM; 3 = one unit, 8 CTAs/SM
KernelFn get_rowstats_kernel(int xdtype, int variant) {
    if (variant == 2) return rowstats<2, 6>(xdtype);
    if (variant == 3) return rowstats<1, 8>(xdtype);
    return rowstats<kRowStatsUnroll, kRowStatsMinBlocks>(xdtype);
}
KernelFn get_stats_kernel(int xdtype, int nw, int group) {
    if (xdtype == DT_F32) return pick<float>(nw, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16>(nw, group);
    return pick<__half>(nw, group);
}
}  // namespace lsqb200
