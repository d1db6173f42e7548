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
template <int U, int MB>
KernelFn rowstats3(int xdtype) {
    if (xdtype == DT_F32) return lsq_rowstats3_kernel<float, kThreads, U, kLd, MB>;
    if (xdtype == DT_BF16) return lsq_rowstats3_kernel<__nv_bfloat16, kThreads, U, kLd, MB>;
    return lsq_rowstats3_kernel<__half, kThreads, U, kLd, MB>;
}
}  // namespace
// variant (Tuning::rowstats): 1 = four units in flight per lane, 4 CTAs/SM; 2 = two units, 6 CTAs/SM; 3 = one unit, 8 CTAs/SM;
// row-entry table + pivot by shuffle (lsq_rowstats3_kernel, plans only): 4 = two units, 6 CTAs/SM; 5 = one unit, 8 CTAs/SM; 6 = four units, 4 CTAs/SM
KernelFn get_rowstats_kernel(int xdtype, int variant) {
    if (variant == 4) return rowstats3<2, 6>(xdtype);
    if (variant == 5) return rowstats3<1, 8>(xdtype);
    if (variant == 6) return rowstats3<4, 4>(xdtype);
    if (variant == 2) return rowstats<2, 6>(xdtype);
    if (variant == 3) return rowstats<1, 8>(xdtype);
    return rowstats<kRowStatsUnroll, kRowStatsMinBlocks>(xdtype);
}
// bulk-copy ring (Tuning::rowstats = 7): rows of whole 32-byte units, one resident CTA per SM, 128 KB ring
KernelFn get_rowstats_ring_kernel(int xdtype) {
    if (xdtype == DT_F32) return lsq_rowstats_ring_kernel<float>;
    if (xdtype == DT_BF16) return lsq_rowstats_ring_kernel<__nv_bfloat16>;
    return lsq_rowstats_ring_kernel<__half>;
}
KernelFn get_stats_kernel(int xdtype, int nw, int group) {
    if (xdtype == DT_F32) return pick<float>(nw, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16>(nw, group);
    return pick<__half>(nw, group);
}
}  // namespace lsqb200
