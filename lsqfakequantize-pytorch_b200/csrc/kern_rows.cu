// kern_rows.cu -- instantiations of the lean warp-per-row forward / backward kernels (aligned weight rows, fp32-internal arithmetic).
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T>
KernelFn pick_f(bool init) {
    return init ? lsq_rowfwd_kernel<T, M_FP32, true, kThreads, kRowUnrollFwd, kLd, kSt, kRowMinBlocksFwd>
                : lsq_rowfwd_kernel<T, M_FP32, false, kThreads, kRowUnrollFwd, kLd, kSt, kRowMinBlocksFwd>;
}
template <typename T>
KernelFn pick_b(int bmode) {
#define LSQ_RB(B_) lsq_rowbwd_kernel<T, M_FP32, B_, kThreads, kRowUnrollBwd, kLd, kSt, kRowMinBlocksBwd>
    switch (bmode) {
        case B_NORMAL: return LSQ_RB(B_NORMAL);
        case B_INIT: return LSQ_RB(B_INIT);
        case B_EVAL: return LSQ_RB(B_EVAL);
        default: return LSQ_RB(B_EVAL_INIT);
    }
#undef LSQ_RB
}
}  // namespace
KernelFn get_rowfwd_kernel(int xdtype, bool init) {
    if (xdtype == DT_F32) return pick_f<float>(init);
    if (xdtype == DT_BF16) return pick_f<__nv_bfloat16>(init);
    return pick_f<__half>(init);
}
KernelFn get_rowbwd_kernel(int xdtype, int bmode) {
    if (xdtype == DT_F32) return pick_b<float>(bmode);
    if (xdtype == DT_BF16) return pick_b<__nv_bfloat16>(bmode);
    return pick_b<__half>(bmode);
}
}  // namespace lsqb200
