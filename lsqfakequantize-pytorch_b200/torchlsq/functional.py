"""`torchlsq.functional.lsq` - same signature and argument meaning as the reference
(/root/reference/torchlsq/functional.py:8-19, :89-97); the work is done by the sm_100a kernels
behind `torch.ops.torchlsq.lsq` (see extension.py and include/lsq_b200.h)."""
import torch
from .extension import _assert_has_ops, _lsq_front


Tensor = torch.Tensor


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
    _assert_has_ops()
    if not is_affine:
        assert quant_min <= 0 <= quant_max, 'quantization range must be covered 0 in symmetric quantization'
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max

    if x.is_cuda:
        # what torch.ops.torchlsq.lsq (CompositeImplicit, extension._lsq_front) does, called directly: one Python
        # dispatcher hop less per call (tools/host_overhead.py)
        return _lsq_front(x, scale, shift, quant_min, quant_max, type_min, type_max,
                          axis, use_grad_scaling, grad_scaler, is_affine, is_perchannel, eval_mode, init_mode)
    return torch.ops.torchlsq.lsq(x, scale, shift, quant_min, quant_max, type_min, type_max,
                                  axis, use_grad_scaling, grad_scaler, is_affine, is_perchannel,
                                  eval_mode, init_mode)
