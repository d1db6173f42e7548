// kern_bwd_f32.cu -- instantiations of the backward kernel family (fp32 tensors).
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE, int NW, int G_>
KernelFn pick_b(int bmode) {
#define LSQ_B(B_) lsq_bwd_kernel<T, MODE, NW, B_, G_, kThreads, unroll_for(kUnrollBwd, NW, G_), kLd, kSt, minb_for(kMinBlocksBwd, G_)>
    switch (bmode) {
        case B_NORMAL: return LSQ_B(B_NORMAL);
        case B_INIT: return LSQ_B(B_INIT);
        case B_EVAL: return LSQ_B(B_EVAL);
        default: return LSQ_B(B_EVAL_INIT);
    }
#undef LSQ_B
}
template <typename T, int MODE, int NW>
KernelFn pick_g(int bmode, int group) {
    return group == 32 ? pick_b<T, MODE, NW, 32>(bmode) : pick_b<T, MODE, NW, kThreads>(bmode);
}
template <typename T, int MODE>
KernelFn pick(int nw, int bmode, int group) {
    switch (nw) {
        case 8: return pick_g<T, MODE, 8>(bmode, group);
        case 4: return pick_g<T, MODE, 4>(bmode, group);
        case 2: return pick_g<T, MODE, 2>(bmode, group);
        default: return pick_g<T, MODE, 0>(bmode, group);
    }
}
}  // namespace
KernelFn get_bwd_kernel_f32(int mode, int nw, int bmode, int group) {
    (void)mode;
    return pick<float, M_FP32>(nw, bmode, group);
}
}  // namespace lsqb200
