// kern_col.cu -- instantiations of the column-layout (short channel rows) kernels.
// Variants (unit words, rows in flight, min CTAs/SM fwd / bwd): 0 = (4, 2, 4/3), 1 = (4, 4, 3/2), 2 = (2, 4, 6/4), 3 = (2, 8, 4/3),
// 4 = (4, 2, 6/4), 5 = (4, 1, 6/4).  Round 2 also measured software-pipelined forms (next row group's loads issued before the current
// group is processed, two register buffers): within +-2 % of variant 1 on every short-row layout (profiles/r2_column_anatomy.md) - the
// launches are bounded by per-CTA prologue / epilogue latency and SM imbalance, not by load latency inside the loop - so they were dropped.
#include "lsq_column.cuh"
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE, int CNW, int U, int MBF, int MBB>
struct V {
    static ColKernelFn f(bool init) {
        return init ? lsq_col_fwd_kernel<T, MODE, true, CNW, U, MBF, kLd, kSt> : lsq_col_fwd_kernel<T, MODE, false, CNW, U, MBF, kLd, kSt>;
    }
    static ColKernelFn b(int bmode) {
        switch (bmode) {
            case B_NORMAL: return lsq_col_bwd_kernel<T, MODE, B_NORMAL, CNW, U, MBB, kLd, kSt>;
            case B_INIT: return lsq_col_bwd_kernel<T, MODE, B_INIT, CNW, U, MBB, kLd, kSt>;
            case B_EVAL: return lsq_col_bwd_kernel<T, MODE, B_EVAL, CNW, U, MBB, kLd, kSt>;
            default: return lsq_col_bwd_kernel<T, MODE, B_EVAL_INIT, CNW, U, MBB, kLd, kSt>;
        }
    }
};
template <typename T, int MODE>
ColKernelFn pick_f(bool init, int v) {
    switch (v) {
        case 1: return V<T, MODE, 4, 4, 3, 2>::f(init);
        case 2: return V<T, MODE, 2, 4, 6, 4>::f(init);
        case 3: return V<T, MODE, 2, 8, 4, 3>::f(init);
        case 4: return V<T, MODE, 4, 2, 6, 4>::f(init);
        case 5: return V<T, MODE, 4, 1, 6, 4>::f(init);
        default: return V<T, MODE, 4, 2, 4, 3>::f(init);
    }
}
template <typename T, int MODE>
ColKernelFn pick_b(int bmode, int v) {
    switch (v) {
        case 1: return V<T, MODE, 4, 4, 3, 2>::b(bmode);
        case 2: return V<T, MODE, 2, 4, 6, 4>::b(bmode);
        case 3: return V<T, MODE, 2, 8, 4, 3>::b(bmode);
        case 4: return V<T, MODE, 4, 2, 6, 4>::b(bmode);
        case 5: return V<T, MODE, 4, 1, 6, 4>::b(bmode);
        default: return V<T, MODE, 4, 2, 4, 3>::b(bmode);
    }
}
}  // namespace
namespace {
template <typename T, int MODE, int R, int S>
ColKernelFn pick_tma_b(int bmode) {
    switch (bmode) {
        case B_NORMAL: return lsq_col_bwd_tma_kernel<T, MODE, B_NORMAL, R, S, kSt>;
        case B_INIT: return lsq_col_bwd_tma_kernel<T, MODE, B_INIT, R, S, kSt>;
        case B_EVAL: return lsq_col_bwd_tma_kernel<T, MODE, B_EVAL, R, S, kSt>;
        default: return lsq_col_bwd_tma_kernel<T, MODE, B_EVAL_INIT, R, S, kSt>;
    }
}
template <typename T, int MODE>
ColKernelFn pick_tma(int bmode, int tv, int* smem) {
    switch (tv) {
        case 2: *smem = 4 * 2 * 2 * kTmaRowBytes; return pick_tma_b<T, MODE, 2, 4>(bmode);
        case 3: *smem = 3 * 2 * 8 * kTmaRowBytes; return pick_tma_b<T, MODE, 8, 3>(bmode);
        default: *smem = 3 * 2 * 4 * kTmaRowBytes; return pick_tma_b<T, MODE, 4, 3>(bmode);
    }
}
}  // namespace
ColKernelFn get_col_bwd_tma_kernel(int xdtype, int mode, int bmode, int tv, int* smem) {
    if (xdtype == DT_F32) return pick_tma<float, M_FP32>(bmode, tv, smem);
    if (xdtype == DT_BF16) return pick_tma<__nv_bfloat16, M_FP32>(bmode, tv, smem);
    return mode == M_HALF_EXACT ? pick_tma<__half, M_HALF_EXACT>(bmode, tv, smem) : pick_tma<__half, M_FP32>(bmode, tv, smem);
}
ColKernelFn get_col_fwd_kernel(int xdtype, int mode, bool init, int v) {
    if (mode == M_FP32_RELU) return get_col_fwd_kernel_pre_relu(xdtype, init);
    if (mode == M_FP32_ADD_RELU) return get_col_fwd_kernel_pre_addrelu(xdtype, init);
    if (mode == M_FP32_ADD) return get_col_fwd_kernel_pre_add(xdtype, init);
    if (xdtype == DT_F32) return pick_f<float, M_FP32>(init, v);
    if (xdtype == DT_BF16) return pick_f<__nv_bfloat16, M_FP32>(init, v);
    return mode == M_HALF_EXACT ? pick_f<__half, M_HALF_EXACT>(init, v) : pick_f<__half, M_FP32>(init, v);
}
ColKernelFn get_col_bwd_kernel(int xdtype, int mode, int bmode, int v) {
#define LSQ_PICK_COL(P) (xdtype == DT_F32 ? get_col_bwd_kernel_pre_##P##_f32(bmode) \
                         : (xdtype == DT_F16 ? get_col_bwd_kernel_pre_##P##_f16(bmode) : get_col_bwd_kernel_pre_##P##_bf16(bmode)))
    if (mode == M_FP32_RELU) return LSQ_PICK_COL(relu);
    if (mode == M_FP32_ADD_RELU) return LSQ_PICK_COL(addrelu);
    if (mode == M_FP32_ADD) return LSQ_PICK_COL(add);
#undef LSQ_PICK_COL
    if (xdtype == DT_F32) return pick_b<float, M_FP32>(bmode, v);
    if (xdtype == DT_BF16) return pick_b<__nv_bfloat16, M_FP32>(bmode, v);
    return mode == M_HALF_EXACT ? pick_b<__half, M_HALF_EXACT>(bmode, v) : pick_b<__half, M_FP32>(bmode, v);
}
}  // namespace lsqb200
