// kern_fwd.cu -- instantiations of the forward kernel family.
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE, int NW>
KernelFn pick_g(bool init, int group) {
#define LSQ_F(INIT_, G_) lsq_fwd_kernel<T, MODE, NW, INIT_, G_, kThreads, unroll_for(kUnrollFwd, NW, G_), kLd, kSt, minb_for(kMinBlocksFwd, G_)>
    if (group == 32) return init ? LSQ_F(true, 32) : LSQ_F(false, 32);
    return init ? LSQ_F(true, kThreads) : LSQ_F(false, kThreads);
#undef LSQ_F
}
template <typename T, int MODE>
KernelFn pick(int nw, bool init, int group) {
    switch (nw) {
        case 8: return pick_g<T, MODE, 8>(init, group);
        case 4: return pick_g<T, MODE, 4>(init, group);
        case 2: return pick_g<T, MODE, 2>(init, group);
        default: return pick_g<T, MODE, 0>(init, group);
    }
}
}  // namespace
KernelFn get_fwd_kernel(int xdtype, int mode, int nw, bool init, int group) {
    if (mode_relu(mode) || mode_add(mode)) return get_fwd_kernel_pre(mode, xdtype, nw, init, group);
    if (xdtype == DT_F64) return get_fwd_kernel_f64(nw, init, group);
    if (xdtype == DT_F32) return pick<float, M_FP32>(nw, init, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16, M_FP32>(nw, init, group);
    if (mode == M_HALF_EXACT) return pick<__half, M_HALF_EXACT>(nw, init, group);
    return pick<__half, M_FP32>(nw, init, group);
}
}  // namespace lsqb200
