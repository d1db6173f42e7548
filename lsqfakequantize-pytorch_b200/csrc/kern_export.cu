// kern_export.cu -- instantiations of the integer export kernels (quantize / dequantize / qparams).
#include "lsq_export.cuh"
#include "lsq_host.h"
namespace lsqb200 {
namespace {
template <typename T, int MODE, int NW, int SEM>
KernelFn pick_g(int dir, int group) {
#define LSQ_E(DIR_, G_) lsq_export_kernel<T, MODE, NW, SEM, DIR_, G_, kThreads, unroll_for(kUnrollFwd, NW, G_), kLd, minb_for(kMinBlocksFwd, G_)>
    if (group == 32) return dir == DIR_QUANT ? LSQ_E(DIR_QUANT, 32) : LSQ_E(DIR_DEQUANT, 32);
    return dir == DIR_QUANT ? LSQ_E(DIR_QUANT, kThreads) : LSQ_E(DIR_DEQUANT, kThreads);
#undef LSQ_E
}
template <typename T, int MODE, int SEM>
KernelFn pick(int nw, int dir, int group) {
    switch (nw) {
        case 8: return pick_g<T, MODE, 8, SEM>(dir, group);
        case 4: return pick_g<T, MODE, 4, SEM>(dir, group);
        case 2: return pick_g<T, MODE, 2, SEM>(dir, group);
        default: return pick_g<T, MODE, 0, SEM>(dir, group);
    }
}
template <typename T, int MODE>
KernelFn pick_sem(int nw, int sem, int dir, int group) {
    if (sem == SEM_TORCH) return pick<T, MODE, SEM_TORCH>(nw, dir, group);
    if (sem == SEM_TORCH_CPU) return pick<T, MODE, SEM_TORCH_CPU>(nw, dir, group);
    return pick<T, MODE, SEM_LSQ>(nw, dir, group);
}
}  // namespace
KernelFn get_export_kernel(int xdtype, int mode, int nw, int sem, int dir, int group) {
    if (xdtype == DT_F32) return pick_sem<float, M_FP32>(nw, sem, dir, group);
    if (xdtype == DT_BF16) return pick_sem<__nv_bfloat16, M_FP32>(nw, sem, dir, group);
    if (mode == M_HALF_EXACT && sem == SEM_LSQ) return pick<__half, M_HALF_EXACT, SEM_LSQ>(nw, dir, group);
    return pick_sem<__half, M_FP32>(nw, sem, dir, group);
}
int launch_qparams(const void* scale, const void* shift, float* scale_out, long long* zp_out, long long n, int pdt,
                   float tmin, float tmax, bool pdl, cudaStream_t st) {
    if (n <= 0) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, lsq_qparams_kernel, scale, shift, scale_out, zp_out, n, pdt, tmin, tmax);
}
}  // namespace lsqb200
