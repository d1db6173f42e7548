"""Loads the B200-native backend and registers the reference's operator surface.

Replaces torchlsq/extension.py + the torch-extension `_C` of the reference
(/root/reference/torchlsq/extension.py:12-98, csrc/torchlsq.cpp:35-39, csrc/ops/lsq.cpp:137-146,
csrc/ops/autograd/lsq_autograd.cpp:290-303, csrc/ops/cuda/lsq_cuda.cu:301-314): instead of
`torch.ops.load_library(_C.so)` we dlopen a C-ABI library of hand-written sm_100a kernels
(`libtorchlsq_b200.so`, include/lsq_b200.h) and define the same six dispatcher entries from
Python with `torch.library`:

    torchlsq::_cuda_version() -> int
    torchlsq::lsq(Tensor, Tensor, Tensor, int, int, int, int, int, bool, float, bool, bool, bool, bool) -> Tensor
    torchlsq::lsq_forward_per_tensor / lsq_backward_per_tensor
    torchlsq::lsq_forward_per_channel / lsq_backward_per_channel      (schemas verbatim)

Kept names: `_HAS_OPS`, `_has_ops`, `_assert_has_ops`, `_check_cuda_version`.
There is no CPU implementation (north_star): CPU tensors raise.
"""
import torch

from . import _cabi

_HAS_OPS = False
error_str = ''


def _has_ops():
    return False


_DT = {torch.float32: _cabi.F32, torch.float16: _cabi.F16, torch.bfloat16: _cabi.BF16}
# the four backend ops (and plans) also take float64 tensors with float64 scale / shift, like the reference's
# AT_DISPATCH_FLOATING_TYPES_AND_HALF (lsq_cuda.cu:45); statistics / observer / export stay 16- and 32-bit
_DT_OPS = dict(_DT)
_DT_OPS[torch.float64] = _cabi.F64
_workspaces = {}
_lib_handle = None   # torch.library.Library must stay alive


def _workspace(device: torch.device, stream_ptr: int):
    """One zero-initialised reduction workspace per (device, stream); kernels keep it zeroed."""
    key = (device.index, stream_ptr)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = _cabi.load().lsqb200_workspace_bytes()
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _dense_layout(x: torch.Tensor, axis=None):
    """View x's memory as a contiguous (outer, C, inner) box without copying when possible.

    Returns (x_dense, outer, C, inner).  Any non-overlapping dense layout (contiguous,
    channels_last, permuted) is processed in memory order - the op is elementwise, and the
    output is allocated with the same strides (the reference's
    empty_like(MemoryFormat::Preserve), csrc/ops/cuda/lsq_cuda.cu:38).  Other layouts are
    made contiguous first.
    """
    if x.numel() == 0:
        return x, 0, (x.shape[axis] if axis is not None else 1), 0
    if x.is_contiguous():                        # the common case, no Python loops over strides
        if axis is None:
            return x, 1, 1, x.numel()
        shape = x.shape
        outer = 1
        for d in range(axis):
            outer *= shape[d]
        C = shape[axis]
        return x, outer, C, x.numel() // (outer * C)
    dims = [d for d in range(x.dim()) if x.shape[d] != 1 or d == axis]
    order = sorted(dims, key=lambda d: (-x.stride(d), d))
    expect, dense = 1, True
    for d in reversed(order):
        if x.stride(d) != expect:
            dense = False
            break
        expect *= x.shape[d]
    if not dense:
        x = x.contiguous()
        order = list(range(x.dim()))
    if axis is None:
        return x, 1, 1, x.numel()
    pos = order.index(axis)
    outer = 1
    for d in order[:pos]:
        outer *= x.shape[d]
    inner = 1
    for d in order[pos + 1:]:
        inner *= x.shape[d]
    return x, outer, x.shape[axis], inner


def _check_common(x, scale, shift, who_x='input'):
    if not x.is_cuda:
        raise RuntimeError(f"`{who_x}` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)")
    if not scale.is_cuda:
        raise RuntimeError("`scale` tensor must be CUDA tensor")
    if not shift.is_cuda:
        raise RuntimeError("`shift` tensor must be CUDA tensor")
    if x.dtype not in _DT_OPS:
        raise RuntimeError(f"`{who_x}` must be float64, float32, float16 or bfloat16 on the B200 path, got {x.dtype}")
    if scale.dtype != shift.dtype:
        raise RuntimeError("`scale` and `shift` must have the same floating-point type")
    # reference: scale/shift dtype == x dtype (lsq_cuda.cu:34-35); superset: fp32 params with fp16 / bf16 x
    if scale.dtype != x.dtype and (scale.dtype != torch.float32 or x.dtype == torch.float64):
        raise RuntimeError(f"`{who_x}` and `scale` must have the same floating-point type"
                           + ("" if x.dtype == torch.float64 else " (or float32 scale/shift)"))


def _match_layout(grad, xd):
    """Upstream grads may be expanded / differently strided (e.g. y.sum().backward()); the
    kernels need grad in exactly x's dense layout."""
    if grad.shape == xd.shape and grad.stride() == xd.stride():
        return grad
    out = torch.empty_like(xd)
    out.copy_(grad if grad.shape == xd.shape else grad.expand_as(xd))
    return out


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream          # (device_index) -> cudaStream_t as int, ~0.3 us
except AttributeError:                                        # pragma: no cover
    _raw_stream = None


def _stream_ptr(device):
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOCTX = _NoCtx()


def _on_device(device):
    """`with torch.cuda.device(d)` costs ~10 us; it is only needed when d is not the current device."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NOCTX
    return torch.cuda.device(device)


_qargs_cache = {}


def _qargs(*scalars):
    """lsqb200_qargs structs are immutable inputs: one per distinct scalar tuple."""
    q = _qargs_cache.get(scalars)
    if q is None:
        if len(_qargs_cache) > 4096:
            _qargs_cache.clear()
        q = _qargs_cache[scalars] = _cabi.qargs(*scalars)
    return q


def _addend_ptr(x2, xd):
    """Device pointer of the second addend of an ADD prologue, which must share x's dense layout (the kernels walk both
    with one index); None without one."""
    if x2 is None:
        return None
    if x2.shape != xd.shape or x2.stride() != xd.stride() or x2.dtype != xd.dtype or x2.device != xd.device:
        raise RuntimeError("the two addends must have the same shape, strides, dtype and device")
    return x2.data_ptr()


def _check_prologue(x, scale):
    if x.dtype == torch.float64 or (x.dtype == torch.float16 and scale.dtype == torch.float16):
        raise RuntimeError("fused-prologue lsq needs float32 / float16 / bfloat16 input with float32 scale / shift "
                           "(float64 and all-float16 calls mirror reference inputs and have no fused prologue)")


def _fwd_tensor_cuda(x, scale, shift, quant_min, quant_max, type_min, type_max,
                     use_grad_scaling, grad_scaler, sym, eval_mode, init_mode, prologue=0, x2=None):
    _check_common(x, scale, shift)
    if scale.numel() < 1 or shift.numel() < 1:
        raise RuntimeError("scale and shift need at least one element")
    lib = _cabi.load()
    xd, _, _, n = _dense_layout(x)
    x2p = _addend_ptr(x2, xd)
    y = torch.empty_like(xd)
    if n == 0:
        return y
    q = _qargs(quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode)
    with _on_device(x.device):
        if prologue:
            rc = lib.lsqb200_fwd_tensor_pre(xd.data_ptr(), x2p, y.data_ptr(), scale.data_ptr(), shift.data_ptr(), n,
                                            _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q, prologue, _stream_ptr(x.device))
        else:
            rc = lib.lsqb200_fwd_tensor(xd.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), n,
                                        _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q, _stream_ptr(x.device))
    _cabi.check(rc, "lsq_forward_per_tensor")
    return y


def _bwd_tensor_cuda(grad, x, scale, shift, quant_min, quant_max, type_min, type_max,
                     use_grad_scaling, grad_scaler, sym, eval_mode, init_mode, prologue=0, x2=None):
    _check_common(x, scale, shift)
    if grad.dtype != x.dtype:
        raise RuntimeError("`grad` and `input` must have the same floating-point type")
    if grad.numel() != x.numel():
        raise RuntimeError("`x` and `grad` are not the same size")
    lib = _cabi.load()
    xd, _, _, n = _dense_layout(x)
    x2p = _addend_ptr(x2, xd)
    gd = _match_layout(grad, xd)
    gx = torch.empty_like(xd)
    gscale = torch.empty(1, dtype=scale.dtype, device=scale.device)
    gshift = torch.empty(1, dtype=shift.dtype, device=shift.device)
    q = _qargs(quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode)
    with _on_device(x.device):
        sp = _stream_ptr(x.device)
        ws = _workspace(x.device, sp)
        if prologue:
            rc = lib.lsqb200_bwd_tensor_pre(gd.data_ptr(), xd.data_ptr(), x2p, gx.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                            gscale.data_ptr(), gshift.data_ptr(), n, _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q,
                                            prologue, ws.data_ptr(), ws.numel(), sp)
        else:
            rc = lib.lsqb200_bwd_tensor(gd.data_ptr(), xd.data_ptr(), gx.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                        gscale.data_ptr(), gshift.data_ptr(), n, _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q,
                                        ws.data_ptr(), ws.numel(), sp)
    _cabi.check(rc, "lsq_backward_per_tensor")
    return gx, gscale, gshift


def _check_channel(x, scale, shift, axis):
    if scale.dim() != 1:
        raise RuntimeError("scale should be a 1-D tensor")
    if shift.dim() != 1:
        raise RuntimeError("shift should be a 1-D tensor")
    if scale.numel() != shift.numel():
        raise RuntimeError("scale and shift need to have the same dimensions")
    if not (0 <= axis < x.dim()):   # reference checks `<= dim` in forward (D12); a clean range check here
        raise RuntimeError("`axis` must be between 0 and number of dimensions of input")
    if scale.numel() != x.shape[axis]:
        raise RuntimeError("dimensions of scale and shift are not consistent with input tensor")


def _fwd_channel_cuda(x, scale, shift, axis, quant_min, quant_max, type_min, type_max,
                      use_grad_scaling, grad_scaler, sym, eval_mode, init_mode, prologue=0, x2=None):
    _check_common(x, scale, shift)
    _check_channel(x, scale, shift, axis)
    lib = _cabi.load()
    xd, outer, C, inner = _dense_layout(x, axis)
    x2p = _addend_ptr(x2, xd)
    y = torch.empty_like(xd)
    if x.numel() == 0:
        return y
    q = _qargs(quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode)
    with _on_device(x.device):
        if prologue:
            rc = lib.lsqb200_fwd_channel_pre(xd.data_ptr(), x2p, y.data_ptr(), scale.data_ptr(), shift.data_ptr(), outer, C, inner,
                                             _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q, prologue, _stream_ptr(x.device))
        else:
            rc = lib.lsqb200_fwd_channel(xd.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), outer, C, inner,
                                         _DT_OPS[x.dtype], _DT_OPS[scale.dtype], q, _stream_ptr(x.device))
    _cabi.check(rc, "lsq_forward_per_channel")
    return y


def _bwd_channel_cuda(grad, x, scale, shift, axis, quant_min, quant_max, type_min, type_max,
                      use_grad_scaling, grad_scaler, sym, eval_mode, init_mode, prologue=0, x2=None):
    _check_common(x, scale, shift)
    _check_channel(x, scale, shift, axis)
    if grad.dtype != x.dtype:
        raise RuntimeError("`grad` and `input` must have the same floating-point type")
    if grad.numel() != x.numel():
        raise RuntimeError("`x` and `grad` are not the same size")
    lib = _cabi.load()
    xd, outer, C, inner = _dense_layout(x, axis)
    x2p = _addend_ptr(x2, xd)
    gd = _match_layout(grad, xd)
    gx = torch.empty_like(xd)
    gscale = torch.empty(C, dtype=scale.dtype, device=scale.device)
    gshift = torch.empty(C, dtype=shift.dtype, device=shift.device)
    q = _qargs(quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode)
    with _on_device(x.device):
        sp = _stream_ptr(x.device)
        ws = _workspace(x.device, sp)
        if prologue:
            rc = lib.lsqb200_bwd_channel_pre(gd.data_ptr(), xd.data_ptr(), x2p, gx.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                             gscale.data_ptr(), gshift.data_ptr(), outer, C, inner, _DT_OPS[x.dtype],
                                             _DT_OPS[scale.dtype], q, prologue, ws.data_ptr(), ws.numel(), sp)
        else:
            rc = lib.lsqb200_bwd_channel(gd.data_ptr(), xd.data_ptr(), gx.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                         gscale.data_ptr(), gshift.data_ptr(), outer, C, inner, _DT_OPS[x.dtype],
                                         _DT_OPS[scale.dtype], q, ws.data_ptr(), ws.numel(), sp)
    _cabi.check(rc, "lsq_backward_per_channel")
    return gx, gscale, gshift


def _no_cpu(name):
    def impl(*args, **kwargs):
        raise RuntimeError(f"torchlsq::{name}: `input` tensor must be CUDA tensor - the B200-native build has no CPU "
                           f"implementation (no CPU fallback by design)")
    return impl


# ---- shape functions (dispatch key Meta): meta / fake tensors flow through the ops, so models can be built on
#      device='meta' and traced for shapes; no arithmetic happens here and nothing falls back to it ---------------
def _fwd_meta(x, scale, shift, *scalars):
    return torch.empty_like(x)


def _bwd_tensor_meta(grad, x, scale, shift, *scalars):
    return torch.empty_like(x), scale.new_empty(1), shift.new_empty(1)


def _bwd_channel_meta(grad, x, scale, shift, axis, *scalars):
    _check_channel(x, scale, shift, axis)
    return torch.empty_like(x), scale.new_empty(x.shape[axis]), shift.new_empty(x.shape[axis])


def _fwd_channel_meta(x, scale, shift, axis, *scalars):
    _check_channel(x, scale, shift, axis)
    return torch.empty_like(x)


# ---- autograd layer (replaces csrc/ops/autograd/lsq_autograd.cpp:16-210) -------------------------
class _LSQPerTensorBackwardFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad, x, scale, shift, *scalars):
        with torch._C._AutoDispatchBelowAutograd():
            return torch.ops.torchlsq.lsq_backward_per_tensor(grad, x, scale, shift, *scalars)

    @staticmethod
    def backward(ctx, *grads):
        raise RuntimeError("double backwards on lsq_per_tensor not supported")


class _LSQPerChannelBackwardFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad, x, scale, shift, axis, *scalars):
        with torch._C._AutoDispatchBelowAutograd():
            return torch.ops.torchlsq.lsq_backward_per_channel(grad, x, scale, shift, axis, *scalars)

    @staticmethod
    def backward(ctx, *grads):
        raise RuntimeError("double backwards on lsq_per_channel not supported")


class _LSQPerTensorFunction(torch.autograd.Function):
    """save {x, scale, shift} + 9 scalars; return 3 grads + 9 None (lsq_autograd.cpp:16-74)."""

    @staticmethod
    def forward(ctx, x, scale, shift, *scalars):
        if x.is_cuda:                                  # straight to the kernel launcher: no second dispatcher hop
            out = _fwd_tensor_cuda(x, scale, shift, *scalars)
        else:
            with torch._C._AutoDispatchBelowAutograd():
                out = torch.ops.torchlsq.lsq_forward_per_tensor(x, scale, shift, *scalars)
        ctx.scalars = scalars
        ctx.save_for_backward(x, scale, shift)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, scale, shift = ctx.saved_tensors
        if x.is_cuda and not torch.is_grad_enabled():
            gx, gs, gb = _bwd_tensor_cuda(grad_output, x, scale, shift, *ctx.scalars)
        else:   # create_graph=True: through the dispatcher, whose Autograd entry refuses the double backward
            gx, gs, gb = torch.ops.torchlsq.lsq_backward_per_tensor(grad_output, x, scale, shift, *ctx.scalars)
        return (gx, gs, gb) + (None,) * 9


class _LSQPerChannelFunction(torch.autograd.Function):
    """lsq_autograd.cpp:111-173."""

    @staticmethod
    def forward(ctx, x, scale, shift, axis, *scalars):
        if x.is_cuda:
            out = _fwd_channel_cuda(x, scale, shift, axis, *scalars)
        else:
            with torch._C._AutoDispatchBelowAutograd():
                out = torch.ops.torchlsq.lsq_forward_per_channel(x, scale, shift, axis, *scalars)
        ctx.axis = axis
        ctx.scalars = scalars
        ctx.save_for_backward(x, scale, shift)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, scale, shift = ctx.saved_tensors
        if x.is_cuda and not torch.is_grad_enabled():
            gx, gs, gb = _bwd_channel_cuda(grad_output, x, scale, shift, ctx.axis, *ctx.scalars)
        else:
            gx, gs, gb = torch.ops.torchlsq.lsq_backward_per_channel(grad_output, x, scale, shift, ctx.axis, *ctx.scalars)
        return (gx, gs, gb) + (None,) * 10


def _lsq_front(x, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling, grad_scaler,
               is_affine, is_perchannel, eval_mode, init_mode):
    """quantops::ops::lsq, csrc/ops/lsq.cpp:104-134."""
    if scale.dim() != 1:
        raise RuntimeError("scale should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)")
    if shift.dim() != 1:
        raise RuntimeError("shift should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)")
    if is_perchannel:
        size = max(scale.size(0), shift.size(0))
        _scale = scale if size == scale.size(0) else scale.repeat(size)
        _shift = shift if size == shift.size(0) else shift.repeat(size)
        if x.is_cuda:     # same autograd node the dispatcher's Autograd entry would build, minus two Python dispatcher hops
            return _LSQPerChannelFunction.apply(x, _scale, _shift, axis, quant_min, quant_max, type_min, type_max,
                                                use_grad_scaling, grad_scaler, not is_affine, eval_mode, init_mode)
        return torch.ops.torchlsq.lsq_forward_per_channel(x, _scale, _shift, axis, quant_min, quant_max, type_min,
                                                          type_max, use_grad_scaling, grad_scaler, not is_affine,
                                                          eval_mode, init_mode)
    if x.is_cuda:
        return _LSQPerTensorFunction.apply(x, scale, shift, quant_min, quant_max, type_min, type_max,
                                           use_grad_scaling, grad_scaler, not is_affine, eval_mode, init_mode)
    return torch.ops.torchlsq.lsq_forward_per_tensor(x, scale, shift, quant_min, quant_max, type_min, type_max,
                                                     use_grad_scaling, grad_scaler, not is_affine, eval_mode, init_mode)


# ---- prologue fusion: fake_quant(relu(x)), fake_quant(relu(a + b)), fake_quant(a + b) in one pass; new
#      surface, nothing in the reference to mirror -----------------------------------------------------------------------
class _LSQPrePerTensorFunction(torch.autograd.Function):
    """y = lsq(pre(x[, x2])); one backward pass gives grad_x (shared by both addends) and the parameter sums."""

    @staticmethod
    def forward(ctx, prologue, x, x2, scale, shift, *scalars):
        out = _fwd_tensor_cuda(x, scale, shift, *scalars, prologue=prologue, x2=x2)
        ctx.scalars, ctx.prologue, ctx.has_x2 = scalars, prologue, x2 is not None
        ctx.save_for_backward(*((x, x2, scale, shift) if x2 is not None else (x, scale, shift)))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        if torch.is_grad_enabled():
            raise RuntimeError("double backwards on fused-prologue lsq not supported")
        if ctx.has_x2:
            x, x2, scale, shift = ctx.saved_tensors
        else:
            (x, scale, shift), x2 = ctx.saved_tensors, None
        gx, gs, gb = _bwd_tensor_cuda(grad_output, x, scale, shift, *ctx.scalars, prologue=ctx.prologue, x2=x2)
        return (None, gx, gx if ctx.has_x2 else None, gs, gb) + (None,) * 9


class _LSQPrePerChannelFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prologue, x, x2, scale, shift, axis, *scalars):
        out = _fwd_channel_cuda(x, scale, shift, axis, *scalars, prologue=prologue, x2=x2)
        ctx.axis, ctx.scalars, ctx.prologue, ctx.has_x2 = axis, scalars, prologue, x2 is not None
        ctx.save_for_backward(*((x, x2, scale, shift) if x2 is not None else (x, scale, shift)))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        if torch.is_grad_enabled():
            raise RuntimeError("double backwards on fused-prologue lsq not supported")
        if ctx.has_x2:
            x, x2, scale, shift = ctx.saved_tensors
        else:
            (x, scale, shift), x2 = ctx.saved_tensors, None
        gx, gs, gb = _bwd_channel_cuda(grad_output, x, scale, shift, ctx.axis, *ctx.scalars, prologue=ctx.prologue, x2=x2)
        return (None, gx, gx if ctx.has_x2 else None, gs, gb) + (None,) * 10


def _lsq_pre_front(prologue, x, x2, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling,
                   grad_scaler, is_affine, is_perchannel, eval_mode, init_mode):
    """Same checks and broadcast as `_lsq_front` (csrc/ops/lsq.cpp:104-134), prologue-fused kernels behind it."""
    if scale.dim() != 1:
        raise RuntimeError("scale should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)")
    if shift.dim() != 1:
        raise RuntimeError("shift should be a 1-D tensor, even in per tensor case(please, avoid torch.Scalar too)")
    if not x.is_cuda:
        raise RuntimeError("`input` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)")
    _check_prologue(x, scale)
    if x2 is not None:
        if x2.shape != x.shape or x2.dtype != x.dtype or x2.device != x.device:
            raise RuntimeError("the two addends must have the same shape, dtype and device")
        # both addends must share one dense layout: the first operand decides, the second is copied only if it differs
        xd = _dense_layout(x, axis if is_perchannel else None)[0]
        if xd.data_ptr() != x.data_ptr() or xd.stride() != x.stride():
            x = xd
        if x2.stride() != x.stride():
            x2 = torch.empty_like(x).copy_(x2)
    if is_perchannel:
        size = max(scale.size(0), shift.size(0))
        _scale = scale if size == scale.size(0) else scale.repeat(size)
        _shift = shift if size == shift.size(0) else shift.repeat(size)
        return _LSQPrePerChannelFunction.apply(prologue, x, x2, _scale, _shift, axis, quant_min, quant_max, type_min, type_max,
                                               use_grad_scaling, grad_scaler, not is_affine, eval_mode, init_mode)
    return _LSQPrePerTensorFunction.apply(prologue, x, x2, scale, shift, quant_min, quant_max, type_min, type_max,
                                          use_grad_scaling, grad_scaler, not is_affine, eval_mode, init_mode)


_TAIL = ("int quant_min, int quant_max, int type_min, int type_max, bool use_grad_scaling, float grad_scaler, "
         "bool sym, bool eval_mode, bool init_mode")


def _register_extensions():
    global _lib_handle
    lib = _cabi.load()          # OSError / AttributeError when the native library is absent or stale
    if lib.lsqb200_abi_version() != _cabi.ABI_VERSION:
        raise ImportError("libtorchlsq_b200.so has an unexpected ABI version")
    L = torch.library.Library("torchlsq", "DEF")
    L.define("_cuda_version() -> int")
    L.define("lsq(Tensor _0, Tensor _1, Tensor _2, int _3, int _4, int _5, int _6, int _7, bool _8, float _9, "
             "bool _10, bool _11, bool _12, bool _13) -> Tensor")
    L.define(f"lsq_forward_per_tensor(Tensor x, Tensor scale, Tensor shift, {_TAIL}) -> Tensor")
    L.define(f"lsq_backward_per_tensor(Tensor grad, Tensor x, Tensor scale, Tensor shift, {_TAIL}) -> (Tensor, Tensor, Tensor)")
    L.define(f"lsq_forward_per_channel(Tensor x, Tensor scale, Tensor shift, int axis, {_TAIL}) -> Tensor")
    L.define(f"lsq_backward_per_channel(Tensor grad, Tensor x, Tensor scale, Tensor shift, int axis, {_TAIL}) -> (Tensor, Tensor, Tensor)")

    L.impl("_cuda_version", lambda: int(lib.lsqb200_cuda_version()), "CompositeExplicitAutograd")
    L.impl("lsq", _lsq_front, "CompositeImplicitAutograd")
    for name, fn in (("lsq_forward_per_tensor", _fwd_tensor_cuda), ("lsq_backward_per_tensor", _bwd_tensor_cuda),
                     ("lsq_forward_per_channel", _fwd_channel_cuda), ("lsq_backward_per_channel", _bwd_channel_cuda)):
        L.impl(name, fn, "CUDA")
        L.impl(name, _no_cpu(name), "CPU")
    for name, fn in (("lsq_forward_per_tensor", _fwd_meta), ("lsq_backward_per_tensor", _bwd_tensor_meta),
                     ("lsq_forward_per_channel", _fwd_channel_meta), ("lsq_backward_per_channel", _bwd_channel_meta)):
        L.impl(name, fn, "Meta")
    L.impl("lsq_forward_per_tensor", lambda *a: _LSQPerTensorFunction.apply(*a), "Autograd")
    L.impl("lsq_backward_per_tensor", lambda *a: _LSQPerTensorBackwardFunction.apply(*a), "Autograd")
    L.impl("lsq_forward_per_channel", lambda *a: _LSQPerChannelFunction.apply(*a), "Autograd")
    L.impl("lsq_backward_per_channel", lambda *a: _LSQPerChannelBackwardFunction.apply(*a), "Autograd")
    _lib_handle = L


try:
    _register_extensions()
    _HAS_OPS = True

    def _has_ops():
        return True
except (ImportError, OSError, AttributeError) as e:
    error_str = str(e)


def _assert_has_ops():
    if not _has_ops():
        raise RuntimeError(
            "Couldn't load the torchlsq B200 backend (libtorchlsq_b200.so). Build it with "
            "`python __graft_entry__.py build` or `make -C lsqfakequantize-pytorch_b200/csrc`; it targets "
            "sm_100a only and there is no CPU or eager fallback."
            f"\n\nImport error details:\n\t{error_str}"
        )


def _check_cuda_version():
    """
    Make sure that CUDA versions match between the pytorch install and torchlsq install
    (same rule as the reference, extension.py:71-96: majors equal, torch minor <= ours).
    """
    if not _HAS_OPS:
        return -1
    _version = torch.ops.torchlsq._cuda_version()
    if _version != -1 and torch.version.cuda is not None:
        ts_major, ts_minor = _version // 1000, (_version % 1000) // 10
        t_major, t_minor = (int(v) for v in torch.version.cuda.split('.')[:2])
        if t_major != ts_major or t_minor > ts_minor:
            raise RuntimeError("Detected that PyTorch and torchlsq were compiled with different CUDA versions. "
                               "PyTorch has CUDA Version={}.{} and torchlsq has CUDA Version={}.{}. "
                               "Please reinstall the torchlsq that matches your PyTorch install."
                               .format(t_major, t_minor, ts_major, ts_minor))
    return _version


_check_cuda_version()
