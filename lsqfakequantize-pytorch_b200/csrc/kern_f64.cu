// kern_f64.cu -- instantiations of the float64 forward / backward kernels (lsq_f64.cuh).
#include "lsq_host.h"
#include "lsq_f64.cuh"
namespace lsqb200 {
namespace {
constexpr int kMinBlocksF64 = 3, kMinBlocksF64Bwd = 2;   // register caps 85 / 128: no spills
template <int NW, int G_>
KernelFn pick_f(bool init) {
#define LSQ_F(INIT_) lsq_fwd_f64_kernel<NW, INIT_, G_, kThreads, unroll_for(kUnrollFwd, NW, G_), kLd, kSt, kMinBlocksF64>
    return init ? LSQ_F(true) : LSQ_F(false);
#undef LSQ_F
}
template <int NW, int G_>
KernelFn pick_b(int bmode) {
#define LSQ_B(B_) lsq_bwd_f64_kernel<NW, B_, G_, kThreads, unroll_for(kUnrollBwd, NW, G_), kLd, kSt, kMinBlocksF64Bwd>
    switch (bmode) {
        case B_NORMAL: return LSQ_B(B_NORMAL);
        case B_INIT: return LSQ_B(B_INIT);
        case B_EVAL: return LSQ_B(B_EVAL);
        default: return LSQ_B(B_EVAL_INIT);
    }
#undef LSQ_B
}
}  // namespace
KernelFn get_fwd_kernel_f64(int nw, bool init, int group) {
    switch (nw) {
        case 8: return group == 32 ? pick_f<8, 32>(init) : pick_f<8, kThreads>(init);
        case 4: return group == 32 ? pick_f<4, 32>(init) : pick_f<4, kThreads>(init);
        case 2: return group == 32 ? pick_f<2, 32>(init) : pick_f<2, kThreads>(init);
        default: return nullptr;   // a double is 8-byte aligned or the call is rejected
    }
}
KernelFn get_bwd_kernel_f64(int nw, int bmode, int group) {
    switch (nw) {
        case 8: return group == 32 ? pick_b<8, 32>(bmode) : pick_b<8, kThreads>(bmode);
        case 4: return group == 32 ? pick_b<4, 32>(bmode) : pick_b<4, kThreads>(bmode);
        case 2: return group == 32 ? pick_b<2, 32>(bmode) : pick_b<2, kThreads>(bmode);
        default: return nullptr;
    }
}
}  // namespace lsqb200
