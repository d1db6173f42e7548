"""`torchlsq.functional.lsq` - same signature and argument meaning as the reference
(/root/reference/torchlsq/functional.py:8-19, :89-97); the work is done by the sm_100a kernels
behind `torch.ops.torchlsq.lsq` (see extension.py and include/lsq_b200.h)."""
import torch
from . import _cabi
from .extension import _assert_has_ops


Tensor = torch.Tensor
_ops = {}


def _op(name):
    """The resolved overload (`torch.ops.torchlsq.<name>.default`): skips the packet's overload resolution on every call."""
    f = _ops.get(name)
    if f is None:
        _assert_has_ops()
        f = _ops[name] = getattr(torch.ops.torchlsq, name).default
    return f


def lsq(x: Tensor, scale: Tensor, shift: Tensor,
        quant_min: int = 0,
        quant_max: int = 255,
        type_min: int = None,
        type_max: int = None,
        axis: int = 1,
        use_grad_scaling: bool = True,
        grad_scaler: float = 1.,
        is_affine: bool = True,
        is_perchannel: bool = False,
        eval_mode: bool = False,
        init_mode: bool = False) -> Tensor:
    """LSQ+ fake quantisation (quantize -> dequantize) with learnable `scale` and `shift`.

    What is computed (the reference's code is the contract, csrc/ops/kernels/lsq_kernel.h):

        s   = max(|scale|, eps)                        zp = rint(clamp(-shift / s, type_min, type_max))
        y   = (rint(clamp(x / s + zp, quant_min, quant_max)) - zp) * s
        dy/dx     = 1 where quant_min < x/s + zp < quant_max (un-rounded, strict), else 0
        dy/dscale = (y - x)/s inside the range; quant_min - zp / quant_max - zp at the borders
        dy/dshift = 0 inside the range, 1 at the borders (always 0 when `is_affine` is False)

    `scale` / `shift` gradients are summed over the tensor (or per channel along `axis`) and,
    with `use_grad_scaling`, multiplied by grad_scaler / sqrt(numel * quant_max).

    Args:
        x: CUDA tensor, float32 / float16 / bfloat16.
        scale, shift: 1-D tensors (one element per tensor, or one per channel), float32 or x's dtype.
        quant_min, quant_max: quantised range (e.g. 0..127).
        type_min, type_max: limits of the integer type holding the zero point; default to the
            quantised range.
        axis: channel dimension when `is_perchannel`.
        use_grad_scaling, grad_scaler: gradient scaling of the learnable parameters.
        is_affine: False selects symmetric quantisation (no shift gradient).
        is_perchannel: one (scale, shift) pair per slice along `axis`.
        eval_mode: plain fake-quant: parameters receive zero gradient.
        init_mode: learned initialisation - output is x itself, the input gradient passes
            through, and the parameters descend on ||y - x||^2.
    """
    if not is_affine:
        assert quant_min <= 0 <= quant_max, 'quantization range must be covered 0 in symmetric quantization'
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max

    return _op("lsq")(x, scale, shift, quant_min, quant_max, type_min, type_max,
                      axis, use_grad_scaling, grad_scaler, is_affine, is_perchannel,
                      eval_mode, init_mode)


_ARGS_DOC = """Arguments after the tensors as in `lsq`.  Inputs: CUDA float32 / float16 / bfloat16 with float32 scale / shift
    (bfloat16 also with bfloat16 parameters); float64 and all-float16 calls (reference-exact contracts) are rejected."""


def _pre(prologue, x, x2, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling, grad_scaler,
         is_affine, is_perchannel, eval_mode, init_mode):
    if not is_affine:
        assert quant_min <= 0 <= quant_max, 'quantization range must be covered 0 in symmetric quantization'
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max
    return _op("lsq_pre")(prologue, x, x2, scale, shift, quant_min, quant_max, type_min, type_max,
                          axis, use_grad_scaling, grad_scaler, is_affine, is_perchannel, eval_mode, init_mode)


def lsq_relu(x: Tensor, scale: Tensor, shift: Tensor, quant_min: int = 0, quant_max: int = 255, type_min: int = None,
             type_max: int = None, axis: int = 1, use_grad_scaling: bool = True, grad_scaler: float = 1.,
             is_affine: bool = True, is_perchannel: bool = False, eval_mode: bool = False, init_mode: bool = False) -> Tensor:
    """`lsq(torch.relu(x), ...)` in ONE pass over x (not in the reference).

    The reference's call sequence `relu` -> `lsq` writes relu(x) to memory and reads it back (4 tensor trips forward,
    6 backward); here the kernels apply max(x, 0) in registers in front of the fake-quant and mask the input gradient
    with x > 0 on the way back (2 and 3 trips).  Values are bit-identical to `lsq(torch.relu(x), ...)` followed by
    autograd through both ops - output, grad_x, grad_scale and grad_shift; NaN stays NaN (it quantises to quant_min as
    in `lsq`).
    """
    return _pre(_cabi.PRE_RELU, x, None, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling,
                grad_scaler, is_affine, is_perchannel, eval_mode, init_mode)


def lsq_add_relu(a: Tensor, b: Tensor, scale: Tensor, shift: Tensor, quant_min: int = 0, quant_max: int = 255,
                 type_min: int = None, type_max: int = None, axis: int = 1, use_grad_scaling: bool = True,
                 grad_scaler: float = 1., is_affine: bool = True, is_perchannel: bool = False, eval_mode: bool = False,
                 init_mode: bool = False) -> Tensor:
    """`lsq(torch.relu(a + b), ...)` - the residual join of a ResNet block and the activation quantizer behind it - in
    ONE pass: a and b are read once, y written once (3 tensor trips instead of 7: add 3, relu 2, lsq 2); the backward
    reads a, b, grad and writes the one gradient both addends share (4 trips instead of 6).  The sum is rounded to the
    tensors' dtype exactly as `a + b` would store it, so every result is bit-identical to the three-op sequence.
    a and b must have the same shape and dtype.
    """
    return _pre(_cabi.PRE_ADD_RELU, a, b, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling,
                grad_scaler, is_affine, is_perchannel, eval_mode, init_mode)


def lsq_add(a: Tensor, b: Tensor, scale: Tensor, shift: Tensor, quant_min: int = 0, quant_max: int = 255,
            type_min: int = None, type_max: int = None, axis: int = 1, use_grad_scaling: bool = True,
            grad_scaler: float = 1., is_affine: bool = True, is_perchannel: bool = False, eval_mode: bool = False,
            init_mode: bool = False) -> Tensor:
    """`lsq(a + b, ...)` in one pass (residual joins without an activation); see `lsq_add_relu`."""
    return _pre(_cabi.PRE_ADD, a, b, scale, shift, quant_min, quant_max, type_min, type_max, axis, use_grad_scaling,
                grad_scaler, is_affine, is_perchannel, eval_mode, init_mode)


for _f in (lsq_relu, lsq_add_relu, lsq_add):
    _f.__doc__ += "\n    " + _ARGS_DOC + "\n"
