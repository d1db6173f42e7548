"""Integer export - the step after QAT.

The reference stops at `LSQFakeQuantizer.calculate_qparams()`
(/root/reference/torchlsq/quantized/modules/observers.py:378-422): it copies `scale` / `shift` to
the host, forms `zero_point = clamp(round(-shift / scale))` there, and leaves the actual
quantisation to `torch.quantization.convert` -> `torch.quantize_per_tensor / _per_channel`.
Here both steps run on the device through the C ABI (include/lsq_b200.h `lsqb200_qparams`,
`lsqb200_quantize`, `lsqb200_dequantize`): 1 byte written per element instead of a float, no
host round trip.

Two code semantics:
  'lsq'    the integer the training forward forms (csrc/ops/kernels/lsq_kernel.h:12-13):
           `dequantize(quantize(x)) == lsq(x)` bit for bit; codes lie in [quant_min, quant_max].
  'torch'  `calculate_qparams()` + torch's CUDA `quantize_per_tensor / quantize_per_channel`
           (what `convert` produces); codes lie in the integer type's range.
  'torch_cpu'  the same with torch's CPU quantizer arithmetic (x * (1.0f / scale)).
There is no CPU path: CPU tensors raise.
"""
from typing import Optional, Tuple

import torch

from . import _cabi
from .extension import _DT, _assert_has_ops, _check_common, _check_channel, _dense_layout, _stream_ptr

Tensor = torch.Tensor
_SEM = {'lsq': _cabi.SEM_LSQ, 'torch': _cabi.SEM_TORCH, 'torch_cpu': _cabi.SEM_TORCH_CPU}
_CODE_DT = {torch.uint8: 0, torch.int8: 1}


def _prep(scale: Tensor, shift: Tensor, is_perchannel: bool, x: Tensor, axis: int):
    if scale.dim() != 1 or shift.dim() != 1:
        raise RuntimeError("scale and shift should be 1-D tensors")
    if is_perchannel:
        size = max(scale.size(0), shift.size(0))           # the front op's broadcast, csrc/ops/lsq.cpp:124-126
        scale = scale if scale.size(0) == size else scale.repeat(size)
        shift = shift if shift.size(0) == size else shift.repeat(size)
        _check_channel(x, scale, shift, axis)
    elif scale.numel() < 1 or shift.numel() < 1:
        raise RuntimeError("scale and shift need at least one element")
    return scale.detach().contiguous(), shift.detach().contiguous()


def _code_dtype(code_dtype, type_min):
    if code_dtype is None:
        return torch.int8 if type_min < 0 else torch.uint8
    code_dtype = {torch.qint8: torch.int8, torch.quint8: torch.uint8}.get(code_dtype, code_dtype)
    if code_dtype not in _CODE_DT:
        raise RuntimeError("codes are torch.uint8 (quint8) or torch.int8 (qint8)")
    return code_dtype


def quantize(x: Tensor, scale: Tensor, shift: Tensor, quant_min: int = 0, quant_max: int = 255,
             type_min: Optional[int] = None, type_max: Optional[int] = None, axis: int = 1,
             is_perchannel: bool = False, code_dtype=None, semantics: str = 'lsq') -> Tensor:
    """x (float32 / float16 / bfloat16, CUDA) -> uint8 / int8 codes with x's shape and strides."""
    _assert_has_ops()
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max
    _check_common(x, scale, shift)
    if x.dtype not in _DT:
        raise RuntimeError(f"integer export takes float32, float16 or bfloat16 tensors, got {x.dtype}")
    scale, shift = _prep(scale, shift, is_perchannel, x, axis)
    cdt = _code_dtype(code_dtype, type_min if semantics != 'lsq' else quant_min)
    xd, outer, C, inner = _dense_layout(x.detach(), axis if is_perchannel else None)
    codes = torch.empty_like(xd, dtype=cdt)
    q = _cabi.qargs(quant_min, quant_max, type_min, type_max, False, 1.0, False, False, False)
    with torch.cuda.device(x.device):
        rc = _cabi.load().lsqb200_quantize(xd.data_ptr(), codes.data_ptr(), scale.data_ptr(), shift.data_ptr(), outer, C,
                                           inner, _DT[x.dtype], _DT[scale.dtype], int(is_perchannel), q, _CODE_DT[cdt],
                                           _SEM[semantics], _stream_ptr(x.device))
    _cabi.check(rc, "lsqb200_quantize")
    return codes


def dequantize(codes: Tensor, scale: Tensor, shift: Tensor, quant_min: int = 0, quant_max: int = 255,
               type_min: Optional[int] = None, type_max: Optional[int] = None, axis: int = 1,
               is_perchannel: bool = False, dtype=torch.float32, semantics: str = 'lsq') -> Tensor:
    """codes (uint8 / int8, CUDA) -> (code - zero_point) * scale in `dtype`."""
    _assert_has_ops()
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max
    if codes.dtype not in _CODE_DT:
        raise RuntimeError("codes must be torch.uint8 or torch.int8")
    if not codes.is_cuda:
        raise RuntimeError("`codes` tensor must be CUDA tensor (torchlsq-b200 has no CPU path)")
    cd, outer, C, inner = _dense_layout(codes, axis if is_perchannel else None)
    y = torch.empty_like(cd, dtype=dtype)
    _check_common(y, scale, shift, who_x='output')
    if dtype not in _DT:
        raise RuntimeError(f"integer export produces float32, float16 or bfloat16 tensors, got {dtype}")
    scale, shift = _prep(scale, shift, is_perchannel, codes, axis)
    q = _cabi.qargs(quant_min, quant_max, type_min, type_max, False, 1.0, False, False, False)
    with torch.cuda.device(codes.device):
        rc = _cabi.load().lsqb200_dequantize(cd.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), outer, C,
                                             inner, _DT[dtype], _DT[scale.dtype], int(is_perchannel), q,
                                             _CODE_DT[codes.dtype], _SEM[semantics], _stream_ptr(codes.device))
    _cabi.check(rc, "lsqb200_dequantize")
    return y


def qparams(scale: Tensor, shift: Tensor, type_min: int, type_max: int) -> Tuple[Tensor, Tensor]:
    """`calculate_qparams()` on the device (observers.py:378-422): (max(scale, eps) float32,
    clamp(round(-shift / scale), type range) int64), no host copy, no sync."""
    _assert_has_ops()
    if not scale.is_cuda or not shift.is_cuda:
        raise RuntimeError("`scale` and `shift` must be CUDA tensors (torchlsq-b200 has no CPU path)")
    if scale.dtype not in _DT or scale.dtype != shift.dtype or scale.numel() != shift.numel():
        raise RuntimeError("`scale` and `shift` must have the same floating-point type and size")
    scale, shift = scale.detach().contiguous(), shift.detach().contiguous()
    s_out = torch.empty(scale.shape, dtype=torch.float32, device=scale.device)
    zp_out = torch.empty(scale.shape, dtype=torch.int64, device=scale.device)
    with torch.cuda.device(scale.device):
        rc = _cabi.load().lsqb200_qparams(scale.data_ptr(), shift.data_ptr(), s_out.data_ptr(), zp_out.data_ptr(),
                                          scale.numel(), _DT[scale.dtype], int(type_min), int(type_max),
                                          _stream_ptr(scale.device))
    _cabi.check(rc, "lsqb200_qparams")
    return s_out, zp_out


def to_quantized_tensor(x: Tensor, fq, semantics: str = 'torch') -> Tensor:
    """Real torch quantized tensor (quint8 / qint8) of `x` under the learned parameters of the
    `LSQFakeQuantizer` `fq` - what `torch.quantization.convert` would make of it, built from
    device-side codes and qparams."""
    from .quantized.modules.observers import TYPES_RANGE_MAPPING
    tmin, tmax = TYPES_RANGE_MAPPING[fq.dtype]['range']
    codes = quantize(x, fq.scale, fq.shift, fq.quant_min, fq.quant_max, tmin, tmax, fq.ch_axis, fq.is_perchannel,
                     code_dtype=fq.dtype, semantics=semantics)
    s, zp = qparams(fq.scale, fq.shift, tmin, tmax)
    if fq.is_perchannel:
        return torch._make_per_channel_quantized_tensor(codes, s.to(torch.float64), zp, fq.ch_axis)
    s_h, zp_h = float(s[0]), int(zp[0])            # a per-tensor quantized tensor carries host scalars
    return torch._make_per_tensor_quantized_tensor(codes, s_h, zp_h)
