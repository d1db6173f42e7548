"""Loads the B200-native backend and the reference's operator surface.

Replaces torchlsq/extension.py + the torch-extension `_C` of the reference
(/root/reference/torchlsq/extension.py:12-98): `torch.ops.load_library(torchlsq/_C.so)` as there, but `_C.so` is a thin
C++ binding (csrc/torch_binding.cpp, no kernels) that registers the same dispatcher entries

    torchlsq::_cuda_version() -> int
    torchlsq::lsq(Tensor, Tensor, Tensor, int, int, int, int, int, bool, float, bool, bool, bool, bool) -> Tensor
    torchlsq::lsq_forward_per_tensor / lsq_backward_per_tensor
    torchlsq::lsq_forward_per_channel / lsq_backward_per_channel      (schemas verbatim, csrc/ops/lsq.cpp:138-145)

with their autograd layer (csrc/ops/autograd/lsq_autograd.cpp) and forwards CUDA tensors to the C-ABI library of
hand-written sm_100a kernels, `libtorchlsq_b200.so` (include/lsq_b200.h).  Python reaches that library directly
(ctypes, `_cabi`) only for what the reference does in Python: statistics, observer step, export, plans.

Kept names: `_HAS_OPS`, `_has_ops`, `_assert_has_ops`, `_check_cuda_version`.
There is no CPU implementation and no eager fallback (north_star): CPU tensors raise; a missing library makes
`_assert_has_ops()` raise.
"""
from pathlib import Path

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


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream          # (device_index) -> cudaStream_t as int, ~0.3 us
except AttributeError:                                        # pragma: no cover
    _raw_stream = None


def _stream_ptr(device):
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


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


BINDING_NAME = "_C.so"     # the reference's extension module name (setup.py:114 `torchlsq._C`)


def _register_extensions():
    lib = _cabi.load()          # OSError / AttributeError when the native library is absent or stale
    if lib.lsqb200_abi_version() != _cabi.ABI_VERSION:
        raise ImportError("libtorchlsq_b200.so has an unexpected ABI version")
    binding = Path(__file__).resolve().parent / BINDING_NAME
    if not binding.exists():
        raise ImportError(f"{binding} not found - build it with `python __graft_entry__.py build` "
                          f"(or `make -C lsqfakequantize-pytorch_b200/csrc`)")
    torch.ops.load_library(str(binding))
    if torch.ops.torchlsq._b200_abi_version() != _cabi.ABI_VERSION:
        raise ImportError("torchlsq/_C.so is linked against another libtorchlsq_b200.so ABI version")


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
            "Couldn't load the torchlsq B200 backend (libtorchlsq_b200.so + _C.so). Build it with "
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
