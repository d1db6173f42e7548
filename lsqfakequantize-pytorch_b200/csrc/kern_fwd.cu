// kern_fwd.cu -- instantiations of the forward kernel family.
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE>
KernelFn pick(bool vec, bool init, int group) {
    constexpr int V = ElemTraits<T>::VEC;
#define LSQ_F(VEC_, INIT_, G_) lsq_fwd_kernel<T, MODE, VEC_, INIT_, G_, kThreads, kUnrollFwd, kLd, kSt, kMinBlocksFwd>
    if (group == 32) {
        if (vec) return init ? LSQ_F(V, true, 32) : LSQ_F(V, false, 32);
        return init ? LSQ_F(1, true, 32) : LSQ_F(1, false, 32);
    }
    if (vec) return init ? LSQ_F(V, true, kThreads) : LSQ_F(V, false, kThreads);
    return init ? LSQ_F(1, true, kThreads) : LSQ_F(1, false, kThreads);
#undef LSQ_F
}
}  // namespace
KernelFn get_fwd_kernel(int xdtype, int mode, bool vec, bool init, int group) {
    if (xdtype == DT_F32) return pick<float, M_FP32>(vec, init, group);
    if (xdtype == DT_BF16) return pick<__nv_bfloat16, M_FP32>(vec, init, group);
    if (mode == M_HALF_EXACT) return pick<__half, M_HALF_EXACT>(vec, init, group);
    return pick<__half, M_FP32>(vec, init, group);
}
}  // namespace lsqb200
