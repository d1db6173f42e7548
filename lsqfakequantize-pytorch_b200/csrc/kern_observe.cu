// kern_observe.cu -- instantiations of the fused observer-step kernel.
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int NW>
KernelFn pick_g(int group) {
#define LSQ_O(G_) lsq_observe_kernel<T, NW, G_, kThreads, unroll_for(8, NW, 256), kLd, 4>   // read-only stream: 4 x 256-bit units in flight per thread
    return group == 32 ? LSQ_O(32) : LSQ_O(kThreads);
#undef LSQ_O
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
KernelFn get_observe_kernel(int xdtype, int nw, int group) {
    if (xdtype == DT_F32) return pick<float>(nw, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16>(nw, group);
    return pick<__half>(nw, group);
}
}  // namespace lsqb200
