// torch_binding.cpp -- the host side of the drop-in above the C ABI, in C++ because the reference's is:
// registers the reference's dispatcher surface (library `torchlsq`) and its autograd layer and forwards CUDA tensors to
// libtorchlsq_b200.so (include/lsq_b200.h).  No kernels here, no CPU implementation, no fallback.
//
// Replaces, for CUDA tensors (citations under /root/reference/torchlsq/csrc/):
//   torchlsq.cpp:35-39            library fragment: `_cuda_version`, `lsq`
//   ops/lsq.cpp:104-146           front op `quantops::ops::lsq` + the four backend schemas (verbatim)
//   ops/autograd/lsq_autograd.cpp:16-303   LSQPer{Tensor,Channel}Function, ...BackwardFunction, key Autograd
//   ops/cuda/lsq_cuda.cu:18-314   host wrappers of the CUDA backend (argument checks, output allocation)
// New surface: `torchlsq::lsq_pre` = fake-quant behind a fused relu / add / add+relu prologue.
//
// Built by csrc/Makefile with g++ against the installed torch headers -> torchlsq/_C.so (the reference's extension name),
// loaded with torch.ops.load_library by torchlsq/extension.py.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../include/lsq_b200.h"

namespace lsqb200_torch {
namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// ---- small helpers ---------------------------------------------------------------------------------------------------
inline int dtype_code(at::ScalarType t) {
    switch (t) {
        case at::kFloat: return LSQB200_F32;
        case at::kHalf: return LSQB200_F16;
        case at::kBFloat16: return LSQB200_BF16;
        case at::kDouble: return LSQB200_F64;
        default: return -1;
    }
}

inline void check_rc(int rc, const char* what) {
    if (rc != 0) {
        const char* msg = lsqb200_last_error();
        TORCH_CHECK(false, what, " failed (code ", rc, "): ", msg ? msg : "?");
    }
}

// same rules and messages as lsq_cuda.cu:31-35 (+ the fp32-parameter superset)
void check_common(const Tensor& x, const Tensor& scale, const Tensor& shift) {
    TORCH_CHECK(x.is_cuda(), "`input` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)");
    TORCH_CHECK(scale.is_cuda(), "`scale` tensor must be CUDA tensor");
    TORCH_CHECK(shift.is_cuda(), "`shift` tensor must be CUDA tensor");
    TORCH_CHECK(dtype_code(x.scalar_type()) >= 0, "`input` must be float64, float32, float16 or bfloat16 on the B200 path, got ",
                x.scalar_type());
    TORCH_CHECK(scale.scalar_type() == shift.scalar_type(), "`scale` and `shift` must have the same floating-point type");
    if (scale.scalar_type() != x.scalar_type()) {
        TORCH_CHECK(scale.scalar_type() == at::kFloat && x.scalar_type() != at::kDouble,
                    "`input` and `scale` must have the same floating-point type",
                    x.scalar_type() == at::kDouble ? "" : " (or float32 scale/shift)");
    }
    TORCH_CHECK(scale.is_contiguous() && shift.is_contiguous(), "`scale` and `shift` must be contiguous");
}

void check_channel(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis) {
    TORCH_CHECK(scale.dim() == 1, "scale should be a 1-D tensor");
    TORCH_CHECK(shift.dim() == 1, "shift should be a 1-D tensor");
    TORCH_CHECK(scale.numel() == shift.numel(), "scale and shift need to have the same dimensions");
    TORCH_CHECK(axis >= 0 && axis < x.dim(), "`axis` must be between 0 and number of dimensions of input");   // D12
    TORCH_CHECK(scale.numel() == x.size(axis), "dimensions of scale and shift are not consistent with input tensor");
}

void check_prologue(const Tensor& x, const Tensor& scale) {
    TORCH_CHECK(!(x.scalar_type() == at::kDouble || (x.scalar_type() == at::kHalf && scale.scalar_type() == at::kHalf)),
                "fused-prologue lsq needs float32 / float16 / bfloat16 input with float32 scale / shift "
                "(float64 and all-float16 calls mirror reference inputs and have no fused prologue)");
}

struct Box {
    Tensor t;                 // x itself when its memory is dense in some dimension order, else a contiguous copy
    int64_t outer, C, inner;  // memory-order view (outer, C, inner); per-tensor: (1, 1, numel)
};

// Any non-overlapping dense layout (contiguous, channels_last, permuted) is processed in memory order and the output
// keeps the strides (the reference's empty_like(MemoryFormat::Preserve), lsq_cuda.cu:38); other layouts are compacted.
Box dense_box(const Tensor& x, int64_t axis /* < 0: per tensor */) {
    if (x.numel() == 0) return {x, 0, axis >= 0 ? x.size(axis) : 1, 0};
    if (x.is_contiguous()) {
        if (axis < 0) return {x, 1, 1, x.numel()};
        int64_t outer = 1;
        for (int64_t d = 0; d < axis; ++d) outer *= x.size(d);
        const int64_t C = x.size(axis);
        return {x, outer, C, x.numel() / (outer * C)};
    }
    const int64_t nd = x.dim();
    std::vector<int64_t> order;
    for (int64_t d = 0; d < nd; ++d)
        if (x.size(d) != 1 || d == axis) order.push_back(d);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return x.stride(a) > x.stride(b); });
    bool dense = true;
    int64_t expect = 1;
    for (auto it = order.rbegin(); it != order.rend(); ++it) {
        if (x.stride(*it) != expect) { dense = false; break; }
        expect *= x.size(*it);
    }
    Tensor t = x;
    if (!dense) {
        t = x.contiguous();
        order.resize(nd);
        std::iota(order.begin(), order.end(), 0);
    }
    if (axis < 0) return {t, 1, 1, t.numel()};
    int64_t outer = 1, inner = 1;
    bool before = true;
    for (int64_t d : order) {
        if (d == axis) { before = false; continue; }
        (before ? outer : inner) *= t.size(d);
    }
    return {t, outer, t.size(axis), inner};
}

// upstream grads may be expanded / differently strided (y.sum().backward()); the kernels walk grad with x's index
Tensor match_layout(const Tensor& grad, const Tensor& xd) {
    if (grad.sizes() == xd.sizes() && grad.strides() == xd.strides()) return grad;
    Tensor out = at::empty_like(xd, at::MemoryFormat::Preserve);
    out.copy_(grad.sizes() == xd.sizes() ? grad : grad.expand_as(xd));
    return out;
}

const void* addend_ptr(const c10::optional<Tensor>& x2, const Tensor& xd) {
    if (!x2.has_value() || !x2->defined()) return nullptr;
    TORCH_CHECK(x2->sizes() == xd.sizes() && x2->strides() == xd.strides() && x2->scalar_type() == xd.scalar_type() &&
                    x2->device() == xd.device(),
                "the two addends must have the same shape, strides, dtype and device");
    return x2->data_ptr();
}

// one zero-initialised reduction workspace per (device, stream); the kernels leave it zeroed (include/lsq_b200.h)
struct Workspace { void* ptr; size_t bytes; };
Workspace workspace(c10::DeviceIndex dev, cudaStream_t stream) {
    static std::mutex mu;
    static auto& pool = *new std::map<std::pair<int, void*>, Tensor>();   // never destroyed: outlives the CUDA context at exit
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(int(dev), (void*)stream);
    auto it = pool.find(key);
    if (it == pool.end()) {
        const auto n = (int64_t)lsqb200_workspace_bytes();
        Tensor ws = at::zeros({n}, at::TensorOptions().dtype(at::kByte).device(at::Device(at::kCUDA, dev)));
        it = pool.emplace(key, std::move(ws)).first;
    }
    return {it->second.data_ptr(), (size_t)it->second.numel()};
}

struct Scalars {   // the nine trailing schema arguments (ops/lsq.cpp:138)
    int64_t quant_min, quant_max, type_min, type_max;
    bool use_grad_scaling;
    double grad_scaler;
    bool sym, eval_mode, init_mode;
    lsqb200_qargs q() const {
        lsqb200_qargs a;
        a.quant_min = quant_min; a.quant_max = quant_max; a.type_min = type_min; a.type_max = type_max;
        a.grad_scaler = grad_scaler; a.use_grad_scaling = use_grad_scaling; a.sym = sym; a.eval_mode = eval_mode;
        a.init_mode = init_mode;
        return a;
    }
};

// ---- launchers: (axis < 0: per tensor) ------------------------------------------------------------------------------------
Tensor forward_cuda(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis, const Scalars& s, int64_t prologue,
                    const c10::optional<Tensor>& x2) {
    check_common(x, scale, shift);
    if (axis >= 0) check_channel(x, scale, shift, axis);
    else TORCH_CHECK(scale.numel() >= 1 && shift.numel() >= 1, "scale and shift need at least one element");
    Box b = dense_box(x, axis);
    const void* x2p = addend_ptr(x2, b.t);
    Tensor y = at::empty_like(b.t, at::MemoryFormat::Preserve);
    if (x.numel() == 0) return y;
    const lsqb200_qargs q = s.q();
    c10::cuda::OptionalCUDAGuard guard(x.device());
    void* stream = (void*)c10::cuda::getCurrentCUDAStream(x.device().index()).stream();
    const int xd = dtype_code(x.scalar_type()), pd = dtype_code(scale.scalar_type());
    int rc;
    if (axis < 0) {
        rc = prologue ? lsqb200_fwd_tensor_pre(b.t.data_ptr(), x2p, y.data_ptr(), scale.data_ptr(), shift.data_ptr(), b.inner, xd, pd, &q,
                                               (int)prologue, stream)
                      : lsqb200_fwd_tensor(b.t.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), b.inner, xd, pd, &q, stream);
        check_rc(rc, "lsq_forward_per_tensor");
    } else {
        rc = prologue ? lsqb200_fwd_channel_pre(b.t.data_ptr(), x2p, y.data_ptr(), scale.data_ptr(), shift.data_ptr(), b.outer, b.C, b.inner,
                                                xd, pd, &q, (int)prologue, stream)
                      : lsqb200_fwd_channel(b.t.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), b.outer, b.C, b.inner, xd, pd,
                                            &q, stream);
        check_rc(rc, "lsq_forward_per_channel");
    }
    return y;
}

// need_gx = false: the input does not require grad -> the kernel reads x and grad and writes nothing but the two sums
std::tuple<Tensor, Tensor, Tensor> backward_cuda(const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                                 int64_t axis, const Scalars& s, int64_t prologue, const c10::optional<Tensor>& x2,
                                                 bool need_gx) {
    check_common(x, scale, shift);
    if (axis >= 0) check_channel(x, scale, shift, axis);
    TORCH_CHECK(grad.scalar_type() == x.scalar_type(), "`grad` and `input` must have the same floating-point type");
    TORCH_CHECK(grad.numel() == x.numel(), "`x` and `grad` are not the same size");
    Box b = dense_box(x, axis);
    const void* x2p = addend_ptr(x2, b.t);
    Tensor gd = match_layout(grad, b.t);
    Tensor gx = need_gx ? at::empty_like(b.t, at::MemoryFormat::Preserve) : Tensor();
    const int64_t nparam = axis >= 0 ? b.C : 1;
    Tensor gscale = at::empty({nparam}, scale.options());
    Tensor gshift = at::empty({nparam}, shift.options());
    if (x.numel() == 0) {      // D13: zeros, not the inputs
        gscale.zero_(); gshift.zero_();
        return {need_gx ? gx : at::empty_like(b.t), gscale, gshift};
    }
    const lsqb200_qargs q = s.q();
    c10::cuda::OptionalCUDAGuard guard(x.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(x.device().index()).stream();
    Workspace ws = workspace(x.device().index(), st);
    void* gxp = need_gx ? gx.data_ptr() : nullptr;
    const int xd = dtype_code(x.scalar_type()), pd = dtype_code(scale.scalar_type());
    int rc;
    if (axis < 0) {
        rc = prologue ? lsqb200_bwd_tensor_pre(gd.data_ptr(), b.t.data_ptr(), x2p, gxp, scale.data_ptr(), shift.data_ptr(),
                                               gscale.data_ptr(), gshift.data_ptr(), b.inner, xd, pd, &q, (int)prologue, ws.ptr, ws.bytes,
                                               (void*)st)
                      : lsqb200_bwd_tensor(gd.data_ptr(), b.t.data_ptr(), gxp, scale.data_ptr(), shift.data_ptr(), gscale.data_ptr(),
                                           gshift.data_ptr(), b.inner, xd, pd, &q, ws.ptr, ws.bytes, (void*)st);
        check_rc(rc, "lsq_backward_per_tensor");
    } else {
        rc = prologue ? lsqb200_bwd_channel_pre(gd.data_ptr(), b.t.data_ptr(), x2p, gxp, scale.data_ptr(), shift.data_ptr(),
                                                gscale.data_ptr(), gshift.data_ptr(), b.outer, b.C, b.inner, xd, pd, &q, (int)prologue,
                                                ws.ptr, ws.bytes, (void*)st)
                      : lsqb200_bwd_channel(gd.data_ptr(), b.t.data_ptr(), gxp, scale.data_ptr(), shift.data_ptr(), gscale.data_ptr(),
                                            gshift.data_ptr(), b.outer, b.C, b.inner, xd, pd, &q, ws.ptr, ws.bytes, (void*)st);
        check_rc(rc, "lsq_backward_per_channel");
    }
    return {gx, gscale, gshift};
}

// ---- dispatcher entries, key CUDA (lsq_cuda.cu:301-314) ---------------------------------------------------------------------
#define LSQ_TAIL_PARAMS int64_t quant_min, int64_t quant_max, int64_t type_min, int64_t type_max, bool use_grad_scaling, \
                        double grad_scaler, bool sym, bool eval_mode, bool init_mode
#define LSQ_TAIL_ARGS quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode
#define LSQ_SCALARS Scalars{quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode}

Tensor fwd_tensor_cuda(const Tensor& x, const Tensor& scale, const Tensor& shift, LSQ_TAIL_PARAMS) {
    return forward_cuda(x, scale, shift, -1, LSQ_SCALARS, 0, c10::nullopt);
}
std::tuple<Tensor, Tensor, Tensor> bwd_tensor_cuda(const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                                   LSQ_TAIL_PARAMS) {
    return backward_cuda(grad, x, scale, shift, -1, LSQ_SCALARS, 0, c10::nullopt, true);
}
Tensor fwd_channel_cuda(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis, LSQ_TAIL_PARAMS) {
    TORCH_CHECK(axis >= 0 && axis < std::max<int64_t>(x.dim(), 1), "`axis` must be between 0 and number of dimensions of input");
    return forward_cuda(x, scale, shift, axis, LSQ_SCALARS, 0, c10::nullopt);
}
std::tuple<Tensor, Tensor, Tensor> bwd_channel_cuda(const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                                    int64_t axis, LSQ_TAIL_PARAMS) {
    TORCH_CHECK(axis >= 0 && axis < std::max<int64_t>(x.dim(), 1), "`axis` must be between 0 and number of dimensions of input");
    return backward_cuda(grad, x, scale, shift, axis, LSQ_SCALARS, 0, c10::nullopt, true);
}

// ---- key CPU: there is no CPU implementation (north_star) -------------------------------------------------------------------
[[noreturn]] void no_cpu(const char* name) {
    TORCH_CHECK(false, "torchlsq::", name, ": `input` tensor must be CUDA tensor - the B200-native build has no CPU implementation "
                "(no CPU fallback by design)");
}
Tensor fwd_tensor_cpu(const Tensor&, const Tensor&, const Tensor&, LSQ_TAIL_PARAMS) { no_cpu("lsq_forward_per_tensor"); }
std::tuple<Tensor, Tensor, Tensor> bwd_tensor_cpu(const Tensor&, const Tensor&, const Tensor&, const Tensor&, LSQ_TAIL_PARAMS) {
    no_cpu("lsq_backward_per_tensor");
}
Tensor fwd_channel_cpu(const Tensor&, const Tensor&, const Tensor&, int64_t, LSQ_TAIL_PARAMS) { no_cpu("lsq_forward_per_channel"); }
std::tuple<Tensor, Tensor, Tensor> bwd_channel_cpu(const Tensor&, const Tensor&, const Tensor&, const Tensor&, int64_t, LSQ_TAIL_PARAMS) {
    no_cpu("lsq_backward_per_channel");
}

// ---- key Meta: shape functions only (models on device='meta', fake-tensor tracing) -------------------------------------------
Tensor fwd_tensor_meta(const Tensor& x, const Tensor&, const Tensor&, LSQ_TAIL_PARAMS) { return at::empty_like(x); }
std::tuple<Tensor, Tensor, Tensor> bwd_tensor_meta(const Tensor&, const Tensor& x, const Tensor& scale, const Tensor& shift, LSQ_TAIL_PARAMS) {
    return {at::empty_like(x), at::empty({1}, scale.options()), at::empty({1}, shift.options())};
}
Tensor fwd_channel_meta(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis, LSQ_TAIL_PARAMS) {
    check_channel(x, scale, shift, axis);
    return at::empty_like(x);
}
std::tuple<Tensor, Tensor, Tensor> bwd_channel_meta(const Tensor&, const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis,
                                                    LSQ_TAIL_PARAMS) {
    check_channel(x, scale, shift, axis);
    return {at::empty_like(x), at::empty({x.size(axis)}, scale.options()), at::empty({x.size(axis)}, shift.options())};
}

// ---- typed handles of the four backend ops (dispatcher re-entry, ops/lsq.cpp:24-29,48-53,69-74,93-98) -------------------------
using FwdT = Tensor(const Tensor&, const Tensor&, const Tensor&, int64_t, int64_t, int64_t, int64_t, bool, double, bool, bool, bool);
using BwdT = std::tuple<Tensor, Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const Tensor&, int64_t, int64_t, int64_t, int64_t,
                                                bool, double, bool, bool, bool);
using FwdC = Tensor(const Tensor&, const Tensor&, const Tensor&, int64_t, int64_t, int64_t, int64_t, int64_t, bool, double, bool, bool, bool);
using BwdC = std::tuple<Tensor, Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const Tensor&, int64_t, int64_t, int64_t, int64_t,
                                                int64_t, bool, double, bool, bool, bool);
#define LSQ_OP_HANDLE(fn, Sig, name)                                                                        \
    const c10::TypedOperatorHandle<Sig>& fn() {                                                             \
        static const auto op = c10::Dispatcher::singleton().findSchemaOrThrow(name, "").typed<Sig>();       \
        return op;                                                                                          \
    }
LSQ_OP_HANDLE(op_fwd_tensor, FwdT, "torchlsq::lsq_forward_per_tensor")
LSQ_OP_HANDLE(op_bwd_tensor, BwdT, "torchlsq::lsq_backward_per_tensor")
LSQ_OP_HANDLE(op_fwd_channel, FwdC, "torchlsq::lsq_forward_per_channel")
LSQ_OP_HANDLE(op_bwd_channel, BwdC, "torchlsq::lsq_backward_per_channel")

// ---- autograd layer (ops/autograd/lsq_autograd.cpp:16-210) ------------------------------------------------------------------
// One Function for all four cases: axis < 0 = per tensor, prologue != 0 = fused prologue.  Saves {x, scale, shift[, x2]} and
// the scalars, returns the three gradients; a backward under create_graph goes through the dispatcher's backward op, whose
// own autograd entry refuses the second differentiation, as the reference's does.
struct Packed {
    static c10::IValue pack(int64_t axis, int64_t prologue, bool has_x2, const Scalars& s) {
        int64_t bits;
        std::memcpy(&bits, &s.grad_scaler, sizeof bits);
        const int64_t flags = (s.use_grad_scaling ? 1 : 0) | (s.sym ? 2 : 0) | (s.eval_mode ? 4 : 0) | (s.init_mode ? 8 : 0) | (has_x2 ? 16 : 0);
        return c10::IValue(std::vector<int64_t>{axis, prologue, flags, s.quant_min, s.quant_max, s.type_min, s.type_max, bits});
    }
};

class LSQFunction : public torch::autograd::Function<LSQFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const Tensor& x, const c10::optional<Tensor>& x2, const Tensor& scale, const Tensor& shift,
                          int64_t axis, int64_t prologue, int64_t quant_min, int64_t quant_max, int64_t type_min, int64_t type_max,
                          bool use_grad_scaling, double grad_scaler, bool sym, bool eval_mode, bool init_mode) {
        const Scalars s = LSQ_SCALARS;
        Tensor out;
        {
            at::AutoDispatchBelowADInplaceOrView below;
            if (x.is_cuda()) out = forward_cuda(x, scale, shift, axis, s, prologue, x2);
            else if (axis < 0) out = op_fwd_tensor().call(x, scale, shift, LSQ_TAIL_ARGS);
            else out = op_fwd_channel().call(x, scale, shift, axis, LSQ_TAIL_ARGS);
        }
        const bool has_x2 = x2.has_value() && x2->defined();
        ctx->saved_data["a"] = Packed::pack(axis, prologue, has_x2, s);
        if (has_x2) ctx->save_for_backward({x, scale, shift, *x2});
        else ctx->save_for_backward({x, scale, shift});
        return out;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_output) {
        const auto a = ctx->saved_data["a"].toIntVector();
        const int64_t axis = a[0], prologue = a[1], flags = a[2];
        Scalars s;
        s.quant_min = a[3]; s.quant_max = a[4]; s.type_min = a[5]; s.type_max = a[6];
        std::memcpy(&s.grad_scaler, &a[7], sizeof(double));
        s.use_grad_scaling = flags & 1; s.sym = flags & 2; s.eval_mode = flags & 4; s.init_mode = flags & 8;
        const bool has_x2 = flags & 16;
        const auto saved = ctx->get_saved_variables();
        const Tensor &x = saved[0], &scale = saved[1], &shift = saved[2];
        c10::optional<Tensor> x2 = has_x2 ? c10::optional<Tensor>(saved[3]) : c10::nullopt;
        const Tensor& g = grad_output[0];
        Tensor gx, gs, gb;
        if (x.is_cuda() && !at::GradMode::is_enabled()) {
            const bool need_gx = ctx->needs_input_grad(0) || (has_x2 && ctx->needs_input_grad(1));
            std::tie(gx, gs, gb) = backward_cuda(g, x, scale, shift, axis, s, prologue, x2, need_gx);
        } else {
            TORCH_CHECK(prologue == 0, "double backwards on fused-prologue lsq not supported");
            if (axis < 0)
                std::tie(gx, gs, gb) = op_bwd_tensor().call(g, x, scale, shift, s.quant_min, s.quant_max, s.type_min, s.type_max,
                                                            s.use_grad_scaling, s.grad_scaler, s.sym, s.eval_mode, s.init_mode);
            else
                std::tie(gx, gs, gb) = op_bwd_channel().call(g, x, scale, shift, axis, s.quant_min, s.quant_max, s.type_min, s.type_max,
                                                             s.use_grad_scaling, s.grad_scaler, s.sym, s.eval_mode, s.init_mode);
        }
        variable_list out(15);
        out[0] = gx;
        if (has_x2) out[1] = gx;     // d(x + x2)/dx2 = 1: both addends share the gradient
        out[2] = gs;
        out[3] = gb;
        return out;
    }
};

// the backward ops are differentiable entries themselves so that a second differentiation fails with the reference's message
// (lsq_autograd.cpp:77-108, :176-210)
class LSQPerTensorBackwardFunction : public torch::autograd::Function<LSQPerTensorBackwardFunction> {
public:
    static variable_list forward(AutogradContext*, const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                 LSQ_TAIL_PARAMS) {
        at::AutoDispatchBelowADInplaceOrView below;
        auto r = op_bwd_tensor().call(grad, x, scale, shift, LSQ_TAIL_ARGS);
        return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
    }
    static variable_list backward(AutogradContext*, variable_list) {
        TORCH_CHECK(0, "double backwards on lsq_per_tensor not supported");
        return {};
    }
};
class LSQPerChannelBackwardFunction : public torch::autograd::Function<LSQPerChannelBackwardFunction> {
public:
    static variable_list forward(AutogradContext*, const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                 int64_t axis, LSQ_TAIL_PARAMS) {
        at::AutoDispatchBelowADInplaceOrView below;
        auto r = op_bwd_channel().call(grad, x, scale, shift, axis, LSQ_TAIL_ARGS);
        return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
    }
    static variable_list backward(AutogradContext*, variable_list) {
        TORCH_CHECK(0, "double backwards on lsq_per_channel not supported");
        return {};
    }
};

Tensor fwd_tensor_autograd(const Tensor& x, const Tensor& scale, const Tensor& shift, LSQ_TAIL_PARAMS) {
    return LSQFunction::apply(x, c10::optional<Tensor>(), scale, shift, (int64_t)-1, (int64_t)0, LSQ_TAIL_ARGS);
}
Tensor fwd_channel_autograd(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t axis, LSQ_TAIL_PARAMS) {
    TORCH_CHECK(axis >= 0, "`axis` must be between 0 and number of dimensions of input");
    return LSQFunction::apply(x, c10::optional<Tensor>(), scale, shift, axis, (int64_t)0, LSQ_TAIL_ARGS);
}
std::tuple<Tensor, Tensor, Tensor> bwd_tensor_autograd(const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                                       LSQ_TAIL_PARAMS) {
    auto r = LSQPerTensorBackwardFunction::apply(grad, x, scale, shift, LSQ_TAIL_ARGS);
    return {r[0], r[1], r[2]};
}
std::tuple<Tensor, Tensor, Tensor> bwd_channel_autograd(const Tensor& grad, const Tensor& x, const Tensor& scale, const Tensor& shift,
                                                        int64_t axis, LSQ_TAIL_PARAMS) {
    auto r = LSQPerChannelBackwardFunction::apply(grad, x, scale, shift, axis, LSQ_TAIL_ARGS);
    return {r[0], r[1], r[2]};
}

// ---- front op quantops::ops::lsq (ops/lsq.cpp:104-134): dim checks, per-channel broadcast, is_affine -> sym --------------------
Tensor lsq_front(const Tensor& x, const Tensor& scale, const Tensor& shift, int64_t quant_min, int64_t quant_max, int64_t type_min,
                 int64_t type_max, int64_t axis, bool use_grad_scaling, double grad_scaler, bool is_affine, bool is_perchannel,
                 bool eval_mode, bool init_mode) {
    TORCH_CHECK(scale.dim() == 1, "scale should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)");
    TORCH_CHECK(shift.dim() == 1, "shift should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)");
    const bool sym = !is_affine;
    if (is_perchannel) {
        const int64_t size = std::max(scale.size(0), shift.size(0));
        Tensor sc = size != scale.size(0) ? scale.repeat({size}) : scale;
        Tensor sh = size != shift.size(0) ? shift.repeat({size}) : shift;
        return op_fwd_channel().call(x, sc, sh, axis, LSQ_TAIL_ARGS);
    }
    return op_fwd_tensor().call(x, scale, shift, LSQ_TAIL_ARGS);
}

// fake_quant(pre(x[, x2])) in one pass; same checks and broadcast as the front op
Tensor lsq_pre_front(int64_t prologue, const Tensor& x_in, const c10::optional<Tensor>& x2_in, const Tensor& scale, const Tensor& shift,
                     int64_t quant_min, int64_t quant_max, int64_t type_min, int64_t type_max, int64_t axis, bool use_grad_scaling,
                     double grad_scaler, bool is_affine, bool is_perchannel, bool eval_mode, bool init_mode) {
    TORCH_CHECK(scale.dim() == 1, "scale should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)");
    TORCH_CHECK(shift.dim() == 1, "shift should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)");
    TORCH_CHECK(x_in.is_cuda(), "`input` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)");
    TORCH_CHECK(prologue >= LSQB200_PRE_RELU && prologue <= LSQB200_PRE_ADD, "unknown prologue");
    check_prologue(x_in, scale);
    Tensor x = x_in;
    c10::optional<Tensor> x2;
    if (x2_in.has_value() && x2_in->defined()) {
        TORCH_CHECK(x2_in->sizes() == x.sizes() && x2_in->scalar_type() == x.scalar_type() && x2_in->device() == x.device(),
                    "the two addends must have the same shape, dtype and device");
        // both addends must share one dense layout: the first operand decides, the second is copied only if it differs
        Box b = dense_box(x, is_perchannel ? axis : -1);
        if (b.t.data_ptr() != x.data_ptr() || b.t.strides() != x.strides()) x = b.t;
        Tensor t2 = *x2_in;
        if (t2.strides() != x.strides()) t2 = at::empty_like(x, at::MemoryFormat::Preserve).copy_(t2);
        x2 = t2;
    } else {
        TORCH_CHECK(prologue == LSQB200_PRE_RELU, "the add prologues need a second addend");
    }
    const bool sym = !is_affine;
    if (is_perchannel) {
        TORCH_CHECK(axis >= 0 && axis < x.dim(), "`axis` must be between 0 and number of dimensions of input");
        const int64_t size = std::max(scale.size(0), shift.size(0));
        Tensor sc = size != scale.size(0) ? scale.repeat({size}) : scale;
        Tensor sh = size != shift.size(0) ? shift.repeat({size}) : shift;
        return LSQFunction::apply(x, x2, sc, sh, axis, prologue, LSQ_TAIL_ARGS);
    }
    return LSQFunction::apply(x, x2, scale, shift, (int64_t)-1, prologue, LSQ_TAIL_ARGS);
}

// ---- grouped fake-quant: many sites (the conv / linear WEIGHTS of a model) behind ONE autograd node and one multi-tensor plan
//      launch per direction (include/lsq_b200.h: lsqb200_plan_*).  New surface; the reference runs one op, one autograd node and
//      1 + 3 + 2 kernels per site (ops/cuda/lsq_cuda.cu:56-58,128-141).  The plan (device-side descriptor table) is cached per
//      group id and rebuilt only when an input's storage moves; the per-step tensors - outputs allocated afresh, the upstream
//      gradients autograd hands over - are swapped in with lsqb200_plan_rebind (a tiny patch kernel, no sync).  Outputs and
//      gradients are views of three flat buffers, so AccumulateGrad takes them over without a copy.
struct GroupState {
    lsqb200_plan* plan = nullptr;
    std::vector<const void*> key;          // data pointers of xs, scales, shifts the plan was built for
    Scalars s{};
    int64_t axis = 0;
    bool per_channel = true;
    int64_t generation = 0, launches = 0;
    std::mutex mu;
    ~GroupState() { if (plan) lsqb200_plan_destroy(plan); }
};
std::mutex g_groups_mu;
std::map<int64_t, std::shared_ptr<GroupState>>& groups() {
    static auto& m = *new std::map<int64_t, std::shared_ptr<GroupState>>();
    return m;
}
std::shared_ptr<GroupState> group_state(int64_t id) {
    std::lock_guard<std::mutex> lock(g_groups_mu);
    auto& m = groups();
    auto it = m.find(id);
    if (it == m.end()) it = m.emplace(id, std::make_shared<GroupState>()).first;
    return it->second;
}

struct GroupLayout {      // element / parameter offsets of every site inside the flat buffers (32-byte aligned slices)
    std::vector<int64_t> off, poff, nparam;
    int64_t total = 0, ptotal = 0;
};
GroupLayout group_layout(at::TensorList xs, at::TensorList scales) {
    GroupLayout L;
    const size_t n = xs.size();
    L.off.resize(n); L.poff.resize(n); L.nparam.resize(n);
    for (size_t i = 0; i < n; i++) {
        L.off[i] = L.total;
        L.total += (xs[i].numel() + 15) / 16 * 16;
        L.poff[i] = L.ptotal;
        L.nparam[i] = scales[i].numel();
        L.ptotal += (scales[i].numel() + 7) / 8 * 8;
    }
    return L;
}

void group_check(at::TensorList xs, at::TensorList scales, at::TensorList shifts, int64_t axis, bool per_channel) {
    const size_t n = xs.size();
    TORCH_CHECK(n > 0 && scales.size() == n && shifts.size() == n, "lsq_group needs as many scale / shift tensors as inputs (and at least one)");
    for (size_t i = 0; i < n; i++) {
        check_common(xs[i], scales[i], shifts[i]);
        TORCH_CHECK(xs[i].scalar_type() == xs[0].scalar_type() && xs[i].device() == xs[0].device() && xs[i].is_contiguous(),
                    "lsq_group inputs must share dtype and device and be contiguous");
        TORCH_CHECK(scales[i].scalar_type() == scales[0].scalar_type(), "lsq_group scale / shift tensors must share one dtype");
        if (per_channel) check_channel(xs[i], scales[i], shifts[i], axis);
        else TORCH_CHECK(scales[i].numel() == 1 && shifts[i].numel() == 1, "per-tensor sites take one scale / shift element");
        TORCH_CHECK(xs[i].numel() > 0, "lsq_group inputs must not be empty");
    }
}

// (re)build the cached plan when an input's storage moved; placeholders for the per-step tensors carry the alignment class
void group_prepare(GroupState& g, at::TensorList xs, at::TensorList scales, at::TensorList shifts, const GroupLayout& L, int64_t axis,
                   bool per_channel, const Scalars& s, const Tensor& flat_like) {
    const size_t n = xs.size();
    std::vector<const void*> key;
    key.reserve(3 * n);
    for (const auto& t : xs) key.push_back(t.data_ptr());
    for (const auto& t : scales) key.push_back(t.data_ptr());
    for (const auto& t : shifts) key.push_back(t.data_ptr());
    const bool same_q = g.plan && g.axis == axis && g.per_channel == per_channel && g.s.quant_min == s.quant_min &&
                        g.s.quant_max == s.quant_max && g.s.type_min == s.type_min && g.s.type_max == s.type_max &&
                        g.s.use_grad_scaling == s.use_grad_scaling && g.s.grad_scaler == s.grad_scaler && g.s.sym == s.sym &&
                        g.s.eval_mode == s.eval_mode && g.s.init_mode == s.init_mode;
    if (g.plan && same_q && key == g.key) return;
    if (g.plan) { lsqb200_plan_destroy(g.plan); g.plan = nullptr; }
    // placeholders for the per-step tensors: never dereferenced (every run is preceded by a rebind of what it uses), they only
    // tell the plan which alignment class - 32 bytes - the real buffers will have
    Tensor pflat_like = at::empty({2 * L.ptotal}, scales[0].options());
    std::vector<lsqb200_segment> segs(n);
    const lsqb200_qargs q = s.q();
    const int xd = dtype_code(xs[0].scalar_type()), pd = dtype_code(scales[0].scalar_type());
    const int64_t es = xs[0].element_size(), ps = scales[0].element_size();
    char* fb = static_cast<char*>(flat_like.data_ptr());
    char* pb = static_cast<char*>(pflat_like.data_ptr());
    for (size_t i = 0; i < n; i++) {
        lsqb200_segment& sg = segs[i];
        std::memset(&sg, 0, sizeof sg);
        Box b = dense_box(xs[i], per_channel ? axis : -1);
        sg.x = xs[i].data_ptr();
        sg.y = fb + L.off[i] * es; sg.grad = sg.y; sg.gx = sg.y;
        sg.scale = scales[i].data_ptr(); sg.shift = shifts[i].data_ptr();
        sg.gscale = pb + L.poff[i] * ps; sg.gshift = pb + (L.ptotal + L.poff[i]) * ps;
        sg.outer = b.outer; sg.C = b.C; sg.inner = b.inner;
        sg.xdtype = xd; sg.pdtype = pd; sg.per_channel = per_channel ? 1 : 0; sg.prologue = LSQB200_PRE_NONE;
        sg.q = q;
    }
    check_rc(lsqb200_plan_create(segs.data(), (int32_t)n, &g.plan), "lsqb200_plan_create");
    g.key = std::move(key); g.s = s; g.axis = axis; g.per_channel = per_channel;
    g.generation++;
    g.launches = lsqb200_plan_launches(g.plan, 0) + lsqb200_plan_launches(g.plan, 1);
}

std::vector<Tensor> group_forward_cuda(at::TensorList xs, at::TensorList scales, at::TensorList shifts, int64_t axis, bool per_channel,
                                       const Scalars& s, int64_t group) {
    group_check(xs, scales, shifts, axis, per_channel);
    const size_t n = xs.size();
    const GroupLayout L = group_layout(xs, scales);
    c10::cuda::OptionalCUDAGuard guard(xs[0].device());
    void* stream = (void*)c10::cuda::getCurrentCUDAStream(xs[0].device().index()).stream();
    Tensor flat = at::empty({L.total}, xs[0].options());
    auto st = group_state(group);
    std::lock_guard<std::mutex> lock(st->mu);
    group_prepare(*st, xs, scales, shifts, L, axis, per_channel, s, flat);
    std::vector<Tensor> ys(n);
    std::vector<void*> yp(n);
    for (size_t i = 0; i < n; i++) {
        ys[i] = flat.as_strided(xs[i].sizes(), xs[i].strides(), L.off[i]);
        yp[i] = ys[i].data_ptr();
    }
    check_rc(lsqb200_plan_rebind(st->plan, yp.data(), nullptr, nullptr, nullptr, nullptr, stream), "lsqb200_plan_rebind");
    check_rc(lsqb200_plan_forward(st->plan, stream), "lsqb200_plan_forward");
    return ys;
}

std::vector<Tensor> group_backward_cuda(const variable_list& grads, at::TensorList xs, at::TensorList scales, at::TensorList shifts,
                                        int64_t axis, bool per_channel, const Scalars& s, int64_t group) {
    const size_t n = xs.size();
    const GroupLayout L = group_layout(xs, scales);
    c10::cuda::OptionalCUDAGuard guard(xs[0].device());
    void* stream = (void*)c10::cuda::getCurrentCUDAStream(xs[0].device().index()).stream();
    Tensor flat = at::empty({L.total}, xs[0].options());
    Tensor pflat = at::empty({2 * L.ptotal}, scales[0].options());
    auto st = group_state(group);
    std::lock_guard<std::mutex> lock(st->mu);
    group_prepare(*st, xs, scales, shifts, L, axis, per_channel, s, flat);
    std::vector<Tensor> keep(n), out(3 * n);
    std::vector<const void*> gp(n);
    std::vector<void*> gxp(n), gsp(n), gbp(n);
    for (size_t i = 0; i < n; i++) {
        Tensor g = grads[i];
        if (!g.defined()) g = at::zeros_like(xs[i]);
        else if (g.strides() != xs[i].strides() || g.scalar_type() != xs[i].scalar_type() || (reinterpret_cast<uintptr_t>(g.data_ptr()) & 31u))
            g = at::empty_like(xs[i]).copy_(g.sizes() == xs[i].sizes() ? g : g.expand_as(xs[i]));
        keep[i] = g;
        gp[i] = g.data_ptr();
        out[i] = flat.as_strided(xs[i].sizes(), xs[i].strides(), L.off[i]);
        out[n + i] = pflat.as_strided({L.nparam[i]}, {1}, L.poff[i]);
        out[2 * n + i] = pflat.as_strided({L.nparam[i]}, {1}, L.ptotal + L.poff[i]);
        gxp[i] = out[i].data_ptr(); gsp[i] = out[n + i].data_ptr(); gbp[i] = out[2 * n + i].data_ptr();
    }
    check_rc(lsqb200_plan_rebind(st->plan, nullptr, gp.data(), gxp.data(), gsp.data(), gbp.data(), stream), "lsqb200_plan_rebind");
    check_rc(lsqb200_plan_backward(st->plan, stream), "lsqb200_plan_backward");
    // `keep` may die here: the kernels are queued on the stream the caching allocator frees the gradients on
    return out;
}

class LSQGroupFunction : public torch::autograd::Function<LSQGroupFunction> {
public:
    static variable_list forward(AutogradContext* ctx, at::TensorList xs, at::TensorList scales, at::TensorList shifts, int64_t axis,
                                 bool per_channel, int64_t group, int64_t quant_min, int64_t quant_max, int64_t type_min, int64_t type_max,
                                 bool use_grad_scaling, double grad_scaler, bool sym, bool eval_mode, bool init_mode) {
        const Scalars s = LSQ_SCALARS;
        variable_list ys;
        {
            at::AutoDispatchBelowADInplaceOrView below;
            ys = group_forward_cuda(xs, scales, shifts, axis, per_channel, s, group);
        }
        int64_t bits;
        std::memcpy(&bits, &s.grad_scaler, sizeof bits);
        const int64_t flags = (s.use_grad_scaling ? 1 : 0) | (s.sym ? 2 : 0) | (s.eval_mode ? 4 : 0) | (s.init_mode ? 8 : 0) | (per_channel ? 16 : 0);
        ctx->saved_data["a"] = c10::IValue(std::vector<int64_t>{axis, group, flags, s.quant_min, s.quant_max, s.type_min, s.type_max, bits,
                                                                   (int64_t)xs.size()});
        variable_list saved;
        saved.reserve(3 * xs.size());
        for (const auto& t : xs) saved.push_back(t);
        for (const auto& t : scales) saved.push_back(t);
        for (const auto& t : shifts) saved.push_back(t);
        ctx->save_for_backward(saved);
        return ys;
    }
    static variable_list backward(AutogradContext* ctx, variable_list grad_output) {
        TORCH_CHECK(!at::GradMode::is_enabled(), "double backwards on grouped lsq not supported");
        const auto a = ctx->saved_data["a"].toIntVector();
        Scalars s;
        s.quant_min = a[3]; s.quant_max = a[4]; s.type_min = a[5]; s.type_max = a[6];
        std::memcpy(&s.grad_scaler, &a[7], sizeof(double));
        const int64_t flags = a[2];
        s.use_grad_scaling = flags & 1; s.sym = flags & 2; s.eval_mode = flags & 4; s.init_mode = flags & 8;
        const size_t n = (size_t)a[8];
        const auto saved = ctx->get_saved_variables();
        at::TensorList all(saved);
        variable_list out = group_backward_cuda(grad_output, all.slice(0, n), all.slice(n, n), all.slice(2 * n, n), a[0], (flags & 16) != 0, s, a[1]);
        out.resize(3 * n + 12);          // + the twelve non-tensor arguments
        return out;
    }
};

std::vector<Tensor> lsq_group(at::TensorList xs, at::TensorList scales, at::TensorList shifts, int64_t quant_min, int64_t quant_max,
                              int64_t type_min, int64_t type_max, int64_t axis, bool use_grad_scaling, double grad_scaler, bool is_affine,
                              bool is_perchannel, bool eval_mode, bool init_mode, int64_t group) {
    TORCH_CHECK(!xs.empty() && xs[0].is_cuda(), "`input` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)");
    const bool sym = !is_affine;
    return LSQGroupFunction::apply(xs, scales, shifts, axis, is_perchannel, group, LSQ_TAIL_ARGS);
}
void lsq_group_release(int64_t group) {
    std::lock_guard<std::mutex> lock(g_groups_mu);
    groups().erase(group);
}
// [plan generation (how often the device table was rebuilt), kernel launches of one forward + one backward]
std::vector<int64_t> lsq_group_info(int64_t group) {
    auto st = group_state(group);
    std::lock_guard<std::mutex> lock(st->mu);
    return {st->generation, st->launches};
}

int64_t cuda_version() { return lsqb200_cuda_version(); }
int64_t binding_abi() { return lsqb200_abi_version(); }

}  // namespace

#define LSQ_TAIL_SCHEMA "int quant_min, int quant_max, int type_min, int type_max, bool use_grad_scaling, float grad_scaler, " \
                        "bool sym, bool eval_mode, bool init_mode"

TORCH_LIBRARY(torchlsq, m) {
    m.def("_cuda_version() -> int", &cuda_version);
    m.def("_b200_abi_version() -> int", &binding_abi);
    // the reference lets the schema be inferred from the C++ signature: positional arguments _0 ... _13 (torchlsq.cpp:37)
    m.def("lsq(Tensor _0, Tensor _1, Tensor _2, int _3, int _4, int _5, int _6, int _7, bool _8, float _9, bool _10, bool _11, "
          "bool _12, bool _13) -> Tensor", &lsq_front);
    m.def("lsq_pre(int prologue, Tensor x, Tensor? x2, Tensor scale, Tensor shift, int quant_min, int quant_max, int type_min, "
          "int type_max, int axis, bool use_grad_scaling, float grad_scaler, bool is_affine, bool is_perchannel, bool eval_mode, "
          "bool init_mode) -> Tensor", &lsq_pre_front);
    m.def("lsq_group(Tensor[] xs, Tensor[] scales, Tensor[] shifts, int quant_min, int quant_max, int type_min, int type_max, int axis, "
          "bool use_grad_scaling, float grad_scaler, bool is_affine, bool is_perchannel, bool eval_mode, bool init_mode, int group) -> Tensor[]",
          &lsq_group);
    m.def("lsq_group_release(int group) -> ()", &lsq_group_release);
    m.def("lsq_group_info(int group) -> int[]", &lsq_group_info);
    m.def("lsq_forward_per_tensor(Tensor x, Tensor scale, Tensor shift, " LSQ_TAIL_SCHEMA ") -> Tensor");
    m.def("lsq_backward_per_tensor(Tensor grad, Tensor x, Tensor scale, Tensor shift, " LSQ_TAIL_SCHEMA ") -> (Tensor, Tensor, Tensor)");
    m.def("lsq_forward_per_channel(Tensor x, Tensor scale, Tensor shift, int axis, " LSQ_TAIL_SCHEMA ") -> Tensor");
    m.def("lsq_backward_per_channel(Tensor grad, Tensor x, Tensor scale, Tensor shift, int axis, " LSQ_TAIL_SCHEMA
          ") -> (Tensor, Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(torchlsq, CUDA, m) {
    m.impl("lsq_forward_per_tensor", &fwd_tensor_cuda);
    m.impl("lsq_backward_per_tensor", &bwd_tensor_cuda);
    m.impl("lsq_forward_per_channel", &fwd_channel_cuda);
    m.impl("lsq_backward_per_channel", &bwd_channel_cuda);
}

TORCH_LIBRARY_IMPL(torchlsq, CPU, m) {
    m.impl("lsq_forward_per_tensor", &fwd_tensor_cpu);
    m.impl("lsq_backward_per_tensor", &bwd_tensor_cpu);
    m.impl("lsq_forward_per_channel", &fwd_channel_cpu);
    m.impl("lsq_backward_per_channel", &bwd_channel_cpu);
}

TORCH_LIBRARY_IMPL(torchlsq, Meta, m) {
    m.impl("lsq_forward_per_tensor", &fwd_tensor_meta);
    m.impl("lsq_backward_per_tensor", &bwd_tensor_meta);
    m.impl("lsq_forward_per_channel", &fwd_channel_meta);
    m.impl("lsq_backward_per_channel", &bwd_channel_meta);
}

TORCH_LIBRARY_IMPL(torchlsq, Autograd, m) {
    m.impl("lsq_forward_per_tensor", &fwd_tensor_autograd);
    m.impl("lsq_backward_per_tensor", &bwd_tensor_autograd);
    m.impl("lsq_forward_per_channel", &fwd_channel_autograd);
    m.impl("lsq_backward_per_channel", &bwd_channel_autograd);
}

}  // namespace lsqb200_torch
