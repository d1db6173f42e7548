// kern_optim.cu -- fused scale / shift parameter update on the flat buffers.
//
// After the backward kernels have written every site's grad_scale / grad_shift into one flat fp32 buffer and the
// data-parallel all-reduce has summed it, ONE launch updates every LSQ parameter of the model (27 702 floats for
// ResNet-50) instead of torch's per-parameter optimizer kernels.  The arithmetic follows torch.optim's own
// single-tensor SGD / Adam (third-party: torch 2.11, torch/optim/sgd.py `_single_tensor_sgd`, torch/optim/adam.py
// `_single_tensor_adam`, non-capturable, no amsgrad) operation by operation in fp32, with `grad_mul` folding DDP's
// 1/world averaging into the same pass.
#include "lsq_host.h"
#include "../../include/lsq_b200.h"

namespace lsqb200 {
namespace {

struct OptimArgs {
    float lr, momentum, one_minus_dampening, weight_decay, grad_mul;
    float beta1_w /* 1 - beta1 */, beta2, one_minus_beta2, eps, step_size /* lr / (1 - beta1^t) */, inv_bc2_sqrt;
    int kind, nesterov, first_step, has_momentum;
};

__device__ __forceinline__ void optim_update(long long i, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                                             float* __restrict__ s2, const OptimArgs& a, bool first_step, float step_size, float inv_bc2_sqrt) {
    float grad = g[i];
    if (a.grad_mul != 1.0f) grad = __fmul_rn(grad, a.grad_mul);
    const float param = p[i];
    if (a.weight_decay != 0.0f) grad = __fmaf_rn(a.weight_decay, param, grad);            // grad.add(param, alpha=wd)
    if (a.kind == 0) {                                                                     // ---- SGD
        if (a.has_momentum) {
            float buf;
            if (first_step) buf = grad;                                                    // buf = clone(grad)
            else buf = __fmaf_rn(a.one_minus_dampening, grad, __fmul_rn(s1[i], a.momentum));   // buf.mul_(mu).add_(grad, alpha=1-damp)
            s1[i] = buf;
            grad = a.nesterov ? __fmaf_rn(a.momentum, buf, grad) : buf;
        }
        p[i] = __fmaf_rn(-a.lr, grad, param);                                              // param.add_(grad, alpha=-lr)
    } else {                                                                               // ---- Adam
        const float m0 = s1[i], v0 = s2[i];
        const float m = __fmaf_rn(a.beta1_w, __fsub_rn(grad, m0), m0);                     // exp_avg.lerp_(grad, 1 - beta1)
        const float v = __fmaf_rn(__fmul_rn(a.one_minus_beta2, grad), grad, __fmul_rn(v0, a.beta2));   // mul_(beta2).addcmul_(g, g, 1-beta2)
        s1[i] = m; s2[i] = v;
        const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), inv_bc2_sqrt), a.eps);      // sqrt(v) / sqrt(1 - beta2^t) + eps
        p[i] = __fmaf_rn(-step_size, __fdiv_rn(m, denom), param);                          // param.addcdiv_(m, denom, value=-step_size)
    }
}

__global__ void __launch_bounds__(256) lsq_flat_optim_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                                                             float* __restrict__ s2, long long n, const OptimArgs a) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    optim_update(i, p, g, s1, s2, a, a.first_step != 0, a.step_size, a.inv_bc2_sqrt);
}

// Per-element step counts (torch.optim keeps `state['step']` per parameter and skips parameters without a gradient): an element
// whose `active` byte is 0 is left alone - parameter, state and count -, every other element advances its own count and takes the
// bias corrections of THAT count.  A quantizer whose scale only starts learning after its observer window (requires_grad False
// until then, observers.py:455-456) so gets the same first Adam step torch.optim gives it.
__global__ void __launch_bounds__(256) lsq_flat_optim_sites_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                                                                   float* __restrict__ s2, int* __restrict__ steps,
                                                                   const unsigned char* __restrict__ active, long long n, const OptimArgs a,
                                                                   double lr, double beta1, double beta2) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (active && !active[i]) return;
    const int t = steps[i] + 1;
    steps[i] = t;
    float step_size = 0.f, inv_bc2_sqrt = 0.f;
    if (a.kind == 1) {
        const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
        step_size = (float)(lr / bc1);
        inv_bc2_sqrt = __fdiv_rn(1.0f, (float)sqrt(bc2));
    }
    optim_update(i, p, g, s1, s2, a, t == 1, step_size, inv_bc2_sqrt);
}

}  // namespace

namespace {
OptimArgs make_args(const lsqb200_optim_args* o) {
    OptimArgs a{};
    a.kind = o->kind; a.nesterov = o->nesterov; a.first_step = o->step <= 1; a.has_momentum = o->momentum != 0.0;
    a.lr = (float)o->lr; a.momentum = (float)o->momentum; a.one_minus_dampening = (float)(1.0 - o->dampening);
    a.weight_decay = (float)o->weight_decay; a.grad_mul = (float)o->grad_mul;
    if (o->kind == 1) {
        const double t = (double)(o->step < 1 ? 1 : o->step);
        const double bc1 = 1.0 - std::pow(o->beta1, t), bc2 = 1.0 - std::pow(o->beta2, t);
        a.beta1_w = (float)(1.0 - o->beta1); a.beta2 = (float)o->beta2; a.one_minus_beta2 = (float)(1.0 - o->beta2);
        a.eps = (float)o->eps; a.step_size = (float)(o->lr / bc1); a.inv_bc2_sqrt = 1.0f / (float)std::sqrt(bc2);
    }
    return a;
}
cudaLaunchConfig_t optim_cfg(long long n, bool pdl, cudaStream_t st, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cfg;
}
}  // namespace

int launch_flat_optim(float* p, const float* g, float* s1, float* s2, long long n, const lsqb200_optim_args* o, bool pdl, cudaStream_t st) {
    if (n <= 0) return 0;
    const OptimArgs a = make_args(o);
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = optim_cfg(n, pdl, st, attr);
    return (int)cudaLaunchKernelEx(&cfg, lsq_flat_optim_kernel, p, g, s1, s2, n, a);
}

int launch_flat_optim_sites(float* p, const float* g, float* s1, float* s2, int* steps, const unsigned char* active, long long n,
                            const lsqb200_optim_args* o, bool pdl, cudaStream_t st) {
    if (n <= 0) return 0;
    const OptimArgs a = make_args(o);
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = optim_cfg(n, pdl, st, attr);
    return (int)cudaLaunchKernelEx(&cfg, lsq_flat_optim_sites_kernel, p, g, s1, s2, steps, active, n, a, o->lr, o->beta1, o->beta2);
}
}  // namespace lsqb200
