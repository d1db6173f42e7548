// kern_bwd_f32.cu -- instantiations of the backward kernel family (fp32 tensors).
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE, int VEC_, int G_>
KernelFn pick_b(int bmode) {
#define LSQ_B(B_) lsq_bwd_kernel<T, MODE, VEC_, B_, G_, kThreads, kUnrollBwd, kLd, kSt, kMinBlocksBwd>
    switch (bmode) {
        case B_NORMAL: return LSQ_B(B_NORMAL);
        case B_INIT: return LSQ_B(B_INIT);
        case B_EVAL: return LSQ_B(B_EVAL);
        default: return LSQ_B(B_EVAL_INIT);
    }
#undef LSQ_B
}
template <typename T, int MODE>
KernelFn pick(bool vec, int bmode, int group) {
    constexpr int V = ElemTraits<T>::VEC;
    if (group == 32) return vec ? pick_b<T, MODE, V, 32>(bmode) : pick_b<T, MODE, 1, 32>(bmode);
    return vec ? pick_b<T, MODE, V, kThreads>(bmode) : pick_b<T, MODE, 1, kThreads>(bmode);
}
}  // namespace
KernelFn get_bwd_kernel_f32(int mode, bool vec, int bmode, int group) {
    (void)mode;
    return pick<float, M_FP32>(vec, bmode, group);
}
}  // namespace lsqb200
