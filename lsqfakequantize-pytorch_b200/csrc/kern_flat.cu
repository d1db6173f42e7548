// kern_flat.cu -- instantiations of the lean per-tensor kernels (one contiguous channel, 32-byte aligned buffers, single launches).
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T>
FlatKernelFn pick_f(bool init) {
    return init ? lsq_flatfwd_kernel<T, M_FP32, true, kThreads, kLd, kSt, kMinBlocksFwd>
                : lsq_flatfwd_kernel<T, M_FP32, false, kThreads, kLd, kSt, kMinBlocksFwd>;
}
template <typename T>
FlatKernelFn pick_b(int bmode) {
#define LSQ_FB(B_) lsq_flatbwd_kernel<T, M_FP32, B_, kThreads, kLd, kSt, kMinBlocksBwd>
    switch (bmode) {
        case B_NORMAL: return LSQ_FB(B_NORMAL);
        case B_INIT: return LSQ_FB(B_INIT);
        case B_EVAL: return LSQ_FB(B_EVAL);
        default: return LSQ_FB(B_EVAL_INIT);
    }
#undef LSQ_FB
}
}  // namespace
FlatKernelFn get_flatfwd_kernel(int xdtype, bool init) {
    if (xdtype == DT_F32) return pick_f<float>(init);
    if (xdtype == DT_BF16) return pick_f<__nv_bfloat16>(init);
    return pick_f<__half>(init);
}
FlatKernelFn get_flatbwd_kernel(int xdtype, int bmode) {
    if (xdtype == DT_F32) return pick_b<float>(bmode);
    if (xdtype == DT_BF16) return pick_b<__nv_bfloat16>(bmode);
    return pick_b<__half>(bmode);
}
}  // namespace lsqb200
