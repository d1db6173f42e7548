/*
 * lsq_b200.h -- C ABI of libtorchlsq_b200.so: the B200-native (sm_100a) LSQ+ fake-quantize path.
 *
 * This is the drop-in boundary for the ONE hot path of torchlsq 2.1 (DeadAt0m/LSQFakeQuantize-
 * PyTorch): what the reference's `torchlsq/_C.so` does behind `torch.ops.torchlsq.*` for CUDA
 * tensors.  Plain pointers and sizes only -- no torch types.  Citations are file:line under
 * /root/reference/torchlsq/.
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer to a contiguous buffer viewed as
 *     (outer, C, inner): per-tensor ops use C == 1; per-channel `axis` splits the shape into
 *     outer = prod(shape[:axis]), C = shape[axis], inner = prod(shape[axis+1:]).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - functions never allocate, never synchronise, never throw.  Return 0 on success, a
 *     negative LSQB200_ERR_* for argument errors, or a positive cudaError_t.
 *     lsqb200_last_error() returns a thread-local message for the last non-zero return.
 *   - scale/shift are read ON DEVICE (the reference syncs the host with scale[0].item(),
 *     csrc/ops/cuda/lsq_cuda.cu:52-53,120-121); |scale| is clamped to eps and 1/s formed
 *     in-kernel with IEEE ops (csrc/ops/cuda/lsq_cuda.cu:54-55, csrc/ops/kernels/lsq_kernel.h:157-159).
 *   - arithmetic contract: csrc/ops/kernels/lsq_kernel.h:6-145 as compiled by nvcc for the
 *     reference CUDA build (x*inv_s+zp and (r-zp)*s-x are single FMAs); see DESIGN.md section 3.
 */
#ifndef LSQ_B200_H
#define LSQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSQB200_ABI_VERSION 3   /* 2: lsqb200_segment gained `prologue` (was reserved) and a trailing `x2`; lsqb200_*_pre entry points
                                 * 3: lsqb200_plan_rebind */

#if defined(__GNUC__)
#define LSQB200_API __attribute__((visibility("default")))
#else
#define LSQB200_API
#endif

/* element types of x / y / grad / grad_x, and of scale / shift / grad_scale / grad_shift */
#define LSQB200_F32 0
#define LSQB200_F16 1
#define LSQB200_BF16 2
/* float64 tensors with float64 scale / shift (the reference dispatches double too:
 * AT_DISPATCH_FLOATING_TYPES_AND_HALF, csrc/ops/cuda/lsq_cuda.cu:45,113,186,266); accepted by the
 * fwd / bwd calls and by plans, not by the statistics / observer / export calls.  The arithmetic
 * is the reference CUDA build's: clamps go through ::fminf / ::fmaxf (csrc/ops/global_scope.h:51-52),
 * i.e. through float -- see csrc/lsq_f64.cuh. */
#define LSQB200_F64 3

#define LSQB200_OK 0
#define LSQB200_ERR_ARG (-1)        /* null pointer, negative size, bad flag */
#define LSQB200_ERR_DTYPE (-2)      /* unsupported (x dtype, param dtype) pair */
#define LSQB200_ERR_WORKSPACE (-3)  /* workspace missing / too small / misaligned */
#define LSQB200_ERR_PLAN (-4)       /* bad plan handle or segment list */

/* Scalar arguments of the four backend ops, in schema order
 * (csrc/ops/lsq.cpp:138-145: quant_min, quant_max, type_min, type_max, use_grad_scaling,
 * grad_scaler, sym, eval_mode, init_mode). */
typedef struct lsqb200_qargs {
    int64_t quant_min, quant_max, type_min, type_max;
    double grad_scaler;
    int32_t use_grad_scaling; /* gs = grad_scaler / sqrt(numel * quant_max), lsq_cuda.cu:124,274 */
    int32_t sym;              /* symmetric: grad_shift == 0, lsq_kernel.h:85 */
    int32_t eval_mode;        /* grad_scale = grad_shift = 0, lsq_kernel.h:126-145 */
    int32_t init_mode;        /* learned init: y = x, gx = g, g' = 2(xfq - x), lsq_kernel.h:13,35,57 */
} lsqb200_qargs;

/* ---- library info ------------------------------------------------------------------------ */
/* LSQB200_ABI_VERSION the library was built with. */
LSQB200_API int lsqb200_abi_version(void);
/* CUDA_VERSION the kernels were compiled with; replaces quantops::cuda_version(),
 * csrc/torchlsq.cpp:25-31 (`torch.ops.torchlsq._cuda_version`). */
LSQB200_API int64_t lsqb200_cuda_version(void);
LSQB200_API const char* lsqb200_last_error(void);
/* Fixed size of the reduction workspace every backward / stats call needs.  The caller
 * allocates it once per (device, stream), ZERO-FILLS IT ONCE, and never touches it again:
 * kernels leave it zeroed.  Two calls may not use the same workspace concurrently. */
LSQB200_API size_t lsqb200_workspace_bytes(void);

/* ---- per-tensor: replaces lsq_forward_per_tensor_impl / lsq_backward_per_tensor_impl,
 *      csrc/ops/cuda/lsq_cuda.cu:18-61 and :64-143 ------------------------------------------ */
LSQB200_API int lsqb200_fwd_tensor(const void* x, void* y, const void* scale, const void* shift,
                       int64_t numel, int xdtype, int pdtype,
                       const lsqb200_qargs* q, void* stream);
/* gx may be NULL (x needs no grad: 2 reads, no write).  gscale/gshift: 1 element of pdtype. */
LSQB200_API int lsqb200_bwd_tensor(const void* grad, const void* x, void* gx,
                       const void* scale, const void* shift, void* gscale, void* gshift,
                       int64_t numel, int xdtype, int pdtype,
                       const lsqb200_qargs* q, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---- per-channel: replaces lsq_forward_per_channel_impl / lsq_backward_per_channel_impl,
 *      csrc/ops/cuda/lsq_cuda.cu:147-199 and :202-297.  scale/shift/gscale/gshift: C elements.
 *      gs uses the WHOLE tensor's numel (CUDA formula, lsq_cuda.cu:274). ---------------------- */
LSQB200_API int lsqb200_fwd_channel(const void* x, void* y, const void* scale, const void* shift,
                        int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                        const lsqb200_qargs* q, void* stream);
LSQB200_API int lsqb200_bwd_channel(const void* grad, const void* x, void* gx,
                        const void* scale, const void* shift, void* gscale, void* gshift,
                        int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                        const lsqb200_qargs* q, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- prologue fusion (not in the reference, where the ops in front of a fake-quant site are
 *      separate ATen passes: add writes x + x2, relu reads it and writes relu(..) to HBM, csrc/ops/cuda/lsq_cuda.cu:18-61
 *      reads that back, and the backward makes the same trips in reverse).  With a prologue the site's input is formed
 *      in registers, x' = x (NONE) | max(x, 0) (RELU) | max(x + x2, 0) (ADD_RELU: the residual join of a ResNet block)
 *      | x + x2 (ADD), where the sum is rounded to the tensor type exactly as ATen's add stores it and max is torch.relu
 *      (NaN stays NaN, -0 -> +0):
 *        forward   y  = lsq_forward(x')
 *        backward  gx = (relu and x' <= 0) ? 0 : lsq_grad_x(grad, x');  grad_scale / grad_shift as the plain call on x'
 *      i.e. bit for bit what the separate passes + autograd return (gx is the gradient of x and, for the ADD
 *      prologues, of x2 as well), in 2 / 3 tensor trips (RELU: instead of 4 / 6) or 3 / 4 (ADD_RELU: instead of 7 / 6).
 *      With init_mode (learned init) y = x' and gx = relu-masked grad.
 *      x2: second addend, same dtype and (outer, C, inner) layout as x; ignored (may be NULL) unless the prologue adds.
 *      Supported for float32 / float16 / bfloat16 tensors with float32 (or, for bfloat16, bfloat16) scale / shift;
 *      other pairs return LSQB200_ERR_DTYPE.  All other arguments as in the calls above. ---------------------------- */
#define LSQB200_PRE_NONE 0
#define LSQB200_PRE_RELU 1
#define LSQB200_PRE_ADD_RELU 2
#define LSQB200_PRE_ADD 3
LSQB200_API int lsqb200_fwd_tensor_pre(const void* x, const void* x2, void* y, const void* scale, const void* shift,
                           int64_t numel, int xdtype, int pdtype,
                           const lsqb200_qargs* q, int prologue, void* stream);
LSQB200_API int lsqb200_bwd_tensor_pre(const void* grad, const void* x, const void* x2, void* gx,
                           const void* scale, const void* shift, void* gscale, void* gshift,
                           int64_t numel, int xdtype, int pdtype,
                           const lsqb200_qargs* q, int prologue, void* workspace, size_t workspace_bytes,
                           void* stream);
LSQB200_API int lsqb200_fwd_channel_pre(const void* x, const void* x2, void* y, const void* scale, const void* shift,
                            int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                            const lsqb200_qargs* q, int prologue, void* stream);
LSQB200_API int lsqb200_bwd_channel_pre(const void* grad, const void* x, const void* x2, void* gx,
                            const void* scale, const void* shift, void* gscale, void* gshift,
                            int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                            const lsqb200_qargs* q, int prologue, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- mu +- 3 sigma weight initialisation: replaces the torch.mean / torch.std passes of
 *      LSQFakeQuantizer._init_weights, quantized/modules/observers.py:329-337.
 *      scale_out: float[C]; C == 1 gives the whole-tensor statistic. ------------------------- */
LSQB200_API int lsqb200_weight_init_stats(const void* w, float* scale_out,
                              int64_t outer, int64_t C, int64_t inner, int xdtype,
                              int64_t quant_min, int64_t quant_max,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused observer step (the module's default `init_mode='observer'`): replaces, for one forward,
 *      `self.activation_post_process(x)` + `calculate_qparams()` + `_set_weights(scale, zero_point)`,
 *      quantized/modules/observers.py:446-449 with torch's MinMaxObserver / MovingAverageMinMaxObserver
 *      (and their PerChannel variants).  One read of x, no host sync, no fp32 copy of x.
 *      min_val / max_val: running state, float[C] or [1], updated in place (+inf / -inf = never observed);
 *      scale_out / shift_out (may be NULL): the LSQ parameters, shift = -zero_point * scale. ---------- */
typedef struct lsqb200_observer_args {
    int64_t quant_min, quant_max;   /* the observer's range (after reduce_range) */
    double averaging_constant;      /* EMA constant; ignored unless moving_average */
    double eps;                     /* lower bound of scale (observer.eps) */
    int32_t moving_average;         /* 0: running min / max, 1: exponential moving average */
    int32_t symmetric;              /* per_tensor_symmetric / per_channel_symmetric */
    int32_t zero_point_sym;         /* zero point of the symmetric scheme: 0 (qint8), 128 or (qmin+qmax)//2 (quint8) */
    int32_t reserved;
} lsqb200_observer_args;
LSQB200_API int lsqb200_observe(const void* x, int64_t outer, int64_t C, int64_t inner, int xdtype, int per_channel,
                    float* min_val, float* max_val, float* scale_out, float* shift_out,
                    const lsqb200_observer_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* ---- integer export, the step after QAT: real uint8 / int8 tensors and
 *      torch-style qparams made on the device.  `codes` is uint8 (codes_signed == 0, quint8) or int8
 *      (codes_signed == 1, qint8), same (outer, C, inner) box as x, 1 byte per element.
 *      LSQB200_SEM_LSQ:   the integer the training forward forms (csrc/ops/kernels/lsq_kernel.h:12-13):
 *                         rint(clamp(x/s + zp, quant_min, quant_max)); dequantize(codes) == lsq forward, bit for bit.
 *      LSQB200_SEM_TORCH: LSQFakeQuantizer.calculate_qparams() (quantized/modules/observers.py:378-422:
 *                         scale = max(scale, eps), zero_point = clamp(round(-shift/scale), type range)) followed by
 *                         torch.quantize_per_tensor / quantize_per_channel on CUDA - what
 *                         torch.quantization.convert does with this module.  Needs float32 scale / shift. ------- */
#define LSQB200_SEM_LSQ 0
#define LSQB200_SEM_TORCH 1      /* torch's CUDA quantizers: nearbyint(double(x) / double(scale)) + zero_point */
#define LSQB200_SEM_TORCH_CPU 2  /* torch's CPU quantizers (fbgemm / quantize_val): nearbyint(x * (1.0f / scale)) + zero_point */
LSQB200_API int lsqb200_quantize(const void* x, void* codes, const void* scale, const void* shift,
                     int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype, int per_channel,
                     const lsqb200_qargs* q, int codes_signed, int semantics, void* stream);
/* y = (code - zp) * s with zp, s as defined by `semantics` */
LSQB200_API int lsqb200_dequantize(const void* codes, void* y, const void* scale, const void* shift,
                       int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype, int per_channel,
                       const lsqb200_qargs* q, int codes_signed, int semantics, void* stream);
/* calculate_qparams on the device: scale_out[i] = max(scale[i], eps) (float), zero_point_out[i] (int64, may be
 * NULL) = clamp(round(-shift[i] / scale_out[i]), type_min, type_max).  No host round trip, no sync. */
LSQB200_API int lsqb200_qparams(const void* scale, const void* shift, float* scale_out, int64_t* zero_point_out,
                    int64_t n, int pdtype, int64_t type_min, int64_t type_max, void* stream);

/* ---- fused parameter step: ONE launch updates every LSQ scale / shift parameter of a model from
 *      the flat gradient buffer the backward kernels (and the data-parallel all-reduce) filled - torch.optim's SGD / Adam
 *      arithmetic (single-tensor path, fp32), with DDP's 1/world averaging folded in as grad_mul.  The reference leaves the
 *      update of its lazily created parameters to the user's optimizer (README.md:101). ------------------------------------ */
#define LSQB200_OPT_SGD 0
#define LSQB200_OPT_ADAM 1
typedef struct lsqb200_optim_args {
    double lr, weight_decay, grad_mul;        /* grad_mul: e.g. 1 / world_size (1 = plain sum) */
    double momentum, dampening;               /* SGD */
    double beta1, beta2, eps;                 /* Adam */
    int64_t step;                             /* 1-based count of this update (Adam bias correction, SGD momentum-buffer init) */
    int32_t kind;                             /* LSQB200_OPT_SGD / LSQB200_OPT_ADAM */
    int32_t nesterov;
} lsqb200_optim_args;
/* params / grads: float[n]; state1: momentum buffer (SGD with momentum) or exp_avg (Adam); state2: exp_avg_sq (Adam) or NULL. */
LSQB200_API int lsqb200_flat_optimizer_step(float* params, const float* grads, float* state1, float* state2, int64_t n,
                                const lsqb200_optim_args* args, void* stream);
/* The same update with torch.optim's PER-PARAMETER step counts (torch/optim/adam.py: `state['step']`, advanced only for parameters
 * that have a gradient): steps[n] (device, int32, zero-initialised by the caller) is each element's own count - it drives Adam's
 * bias corrections and the first-step initialisation of SGD's momentum buffer -, and active[n] (device bytes, or NULL = all)
 * marks the elements that take part in this step; the others keep parameter, state and count.  `args->step` is ignored.  A scale
 * that only becomes learnable after its observer window (observers.py:455-456) then gets the first step torch.optim gives it. */
LSQB200_API int lsqb200_flat_optimizer_step_sites(float* params, const float* grads, float* state1, float* state2, int32_t* steps,
                                      const uint8_t* active, int64_t n, const lsqb200_optim_args* args, void* stream);

/* ---- multi-tensor plans: one launch for many fake-quant sites (the 54 ResNet-50 weights, or
 *      every site of a step).  Semantics per segment are exactly those of the calls above. --- */
typedef struct lsqb200_segment {
    const void* x;      /* input */
    void* y;            /* forward output (may be NULL if the plan is only run backward) */
    const void* grad;   /* upstream gradient (backward) */
    void* gx;           /* grad_x output, may be NULL */
    const void* scale;  /* [C] (or [1] when per_channel == 0) */
    const void* shift;
    void* gscale;       /* [C] or [1], pdtype -- e.g. a slice of the data-parallel flat grad buffer */
    void* gshift;
    int64_t outer, C, inner; /* per-tensor: outer = 1, C = 1, inner = numel */
    int32_t xdtype, pdtype;
    int32_t per_channel;
    int32_t prologue;   /* LSQB200_PRE_*: see lsqb200_fwd_tensor_pre (0 = none) */
    lsqb200_qargs q;
    const void* x2;     /* second addend of the ADD prologues (same layout as x), else NULL */
} lsqb200_segment;

typedef struct lsqb200_plan lsqb200_plan; /* opaque */

/* Copies the segment table to the device and sizes a private workspace (this call allocates
 * and synchronises; the run calls below do neither).  All segments must share xdtype/pdtype
 * alignment class; mixed dtypes are allowed. */
LSQB200_API int lsqb200_plan_create(const lsqb200_segment* segs, int32_t nseg, lsqb200_plan** out);
/* Re-point the per-step tensors of every segment - outputs that are allocated afresh each training step (y, gx, the
 * parameter gradients) and the upstream gradient autograd hands over - while x, scale, shift and the layout stay as created.
 * Each argument is an array of `nseg` device pointers (plan_create's segment order) or NULL to keep the current ones; a new
 * pointer must be at least as aligned (up to 32 bytes) as the one it replaces.  The device-side table is patched by a tiny
 * kernel on `stream` (no allocation, no synchronisation, no host staging): use the plan on that one stream. */
LSQB200_API int lsqb200_plan_rebind(lsqb200_plan* plan, void* const* y, const void* const* grad, void* const* gx,
                        void* const* gscale, void* const* gshift, void* stream);
LSQB200_API int lsqb200_plan_forward(lsqb200_plan* plan, void* stream);
LSQB200_API int lsqb200_plan_backward(lsqb200_plan* plan, void* stream);
/* float scale_out per channel written to each segment's `gscale` slot is NOT used here:
 * stats go to `scale_out[offset_of_segment ...]`, offsets = running sum of C (or 1). */
LSQB200_API int lsqb200_plan_weight_init_stats(lsqb200_plan* plan, float* scale_out, void* stream);
LSQB200_API int lsqb200_plan_destroy(lsqb200_plan* plan);
/* number of kernel launches one plan_forward / plan_backward issues (for launch accounting) */
LSQB200_API int lsqb200_plan_launches(const lsqb200_plan* plan, int backward);

/* ---- introspection used by tests and bench.py ------------------------------------------- */
typedef struct lsqb200_launch_info {
    int32_t regime;       /* 0 = flat contiguous channel rows, 1 = strided rows */
    int32_t vec;          /* elements per 128-bit access (1 = scalar fallback) */
    int32_t threads;      /* CTA size */
    int32_t splits;       /* CTAs per channel */
    int64_t grid;         /* CTAs launched */
    int64_t units_per_split;
} lsqb200_launch_info;
/* Geometry the library would pick for a (outer, C, inner) tensor; pure host computation. */
LSQB200_API int lsqb200_query_launch(int64_t outer, int64_t C, int64_t inner, int xdtype, int backward,
                         int aligned16, lsqb200_launch_info* out);
/* Override launch-geometry knobs (NULL / "" restores defaults), e.g.
 * "fwd_tile_kb=32,bwd_tile_kb=256,interleave=1,max_unit_bytes=32".  For experiments
 * (also read once from $LSQB200_TUNE); not part of the reference surface.  Kernel-family switches, all with
 * bit-identical (flatkernels) or reduction-order-identical results: "flatkernels=0|1|2" (lean per-tensor forward [and
 * backward] for single launches; default 1), "rowkernels=0|1" (lean warp-per-row forward / backward for weight rows),
 * "rowstats=0..7" (warp-per-row statistics: 1-3 descriptor form, 4-6 row-entry form, 7 bulk-copy ring; default 4), "column_path=0|1", "col_tma=0..3", "pdl=0|1". */
LSQB200_API int lsqb200_set_tuning(const char* spec);

#ifdef __cplusplus
}
#endif
#endif /* LSQ_B200_H */
