"""Multi-tensor plans: one kernel launch for many fake-quant sites.

New relative to the reference (it launches 1 forward + 3 backward element-wise kernels + 2
`at::sum` per site, csrc/ops/cuda/lsq_cuda.cu:56-58,128-141,194-196,277-295).  A plan holds a
device-side table of site descriptors (`lsqb200_plan_*` in include/lsq_b200.h); running it
costs one launch per (dtype, mode, unit-width) class, typically one for all conv / linear
weights of a model.  Results are bit-identical to the per-site calls.

The plan keeps references to every tensor it was built from: pointers must stay valid, so
buffers are allocated once and written in place (the usual static-buffer discipline of CUDA
graphs).
"""
import ctypes
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _cabi
from .extension import _DT_OPS as _DT, _dense_layout


@dataclass
class Site:
    """One fake-quant site.  `gscale` / `gshift` may be slices of a flat gradient buffer
    (see torchlsq.dp.FlatGradBuffer) so the data-parallel all-reduce needs no gather step."""
    x: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor
    y: Optional[torch.Tensor] = None
    grad: Optional[torch.Tensor] = None
    gx: Optional[torch.Tensor] = None
    gscale: Optional[torch.Tensor] = None
    gshift: Optional[torch.Tensor] = None
    quant_min: int = 0
    quant_max: int = 127
    type_min: Optional[int] = None
    type_max: Optional[int] = None
    axis: int = 1
    use_grad_scaling: bool = True
    grad_scaler: float = 1.0
    is_affine: bool = True
    is_perchannel: bool = False
    eval_mode: bool = False
    init_mode: bool = False
    fuse_relu: bool = False   # fake-quant of relu(x) in one pass (include/lsq_b200.h: LSQB200_PRE_RELU)
    x2: Optional[torch.Tensor] = None   # second addend: the site quantises x + x2 (relu(x + x2) with fuse_relu), LSQB200_PRE_ADD*


def _ptr(t):
    return None if t is None else t.data_ptr()


class LSQPlan:
    def __init__(self, sites: List[Site]):
        if not sites:
            raise ValueError("LSQPlan needs at least one site")
        self._lib = _cabi.load()
        self.sites = list(sites)
        dev = sites[0].x.device
        segs = (_cabi.Segment * len(sites))()
        self._keep = []
        nslots = 0
        for i, s in enumerate(sites):
            for t in (s.x, s.scale, s.shift):
                if not t.is_cuda or t.device != dev:
                    raise RuntimeError("all plan tensors must live on one CUDA device")
            if s.x.dtype not in _DT or s.scale.dtype not in _DT or (s.x.dtype == torch.float64) != (s.scale.dtype == torch.float64):
                raise RuntimeError(f"unsupported dtype in plan site {i}")
            xd, outer, C, inner = _dense_layout(s.x, s.axis if s.is_perchannel else None)
            if xd.data_ptr() != s.x.data_ptr():
                raise RuntimeError("plan tensors must be dense (contiguous in some dimension order)")
            for other in (s.y, s.grad, s.gx, s.x2):
                if other is not None and (other.shape != s.x.shape or other.stride() != s.x.stride() or other.dtype != s.x.dtype):
                    raise RuntimeError("y / grad / gx / x2 must match x in shape, strides and dtype")
            nparam = C if s.is_perchannel else 1
            if s.scale.numel() != nparam or s.shift.numel() != nparam:
                raise RuntimeError("scale / shift length does not match the channel count")
            tmin = s.quant_min if s.type_min is None else s.type_min
            tmax = s.quant_max if s.type_max is None else s.type_max
            seg = segs[i]
            seg.x, seg.y, seg.grad, seg.gx = _ptr(s.x), _ptr(s.y), _ptr(s.grad), _ptr(s.gx)
            seg.scale, seg.shift, seg.gscale, seg.gshift = _ptr(s.scale), _ptr(s.shift), _ptr(s.gscale), _ptr(s.gshift)
            seg.outer, seg.C, seg.inner = outer, C, inner
            seg.xdtype, seg.pdtype = _DT[s.x.dtype], _DT[s.scale.dtype]
            seg.per_channel = int(s.is_perchannel)
            if s.x2 is not None:
                seg.prologue = _cabi.PRE_ADD_RELU if s.fuse_relu else _cabi.PRE_ADD
                seg.x2 = _ptr(s.x2)
            else:
                seg.prologue = _cabi.PRE_RELU if s.fuse_relu else _cabi.PRE_NONE
            seg.q = _cabi.qargs(s.quant_min, s.quant_max, tmin, tmax, s.use_grad_scaling, s.grad_scaler,
                                not s.is_affine, s.eval_mode, s.init_mode)
            nslots += nparam
        self.num_param_slots = nslots
        self.device = dev
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(self._lib.lsqb200_plan_create(segs, len(sites), ctypes.byref(handle)), "lsqb200_plan_create")
        self._h = handle

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lsqb200_plan_forward(self._h, self._stream()), "lsqb200_plan_forward")

    def backward(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lsqb200_plan_backward(self._h, self._stream()), "lsqb200_plan_backward")

    def weight_init_stats(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mu +- 3 sigma scales of every site in one launch; float32 [sum of channels]."""
        if out is None:
            out = torch.empty(self.num_param_slots, dtype=torch.float32, device=self.device)
        if out.numel() < self.num_param_slots or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("`out` must be a contiguous float32 tensor with one slot per channel")
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lsqb200_plan_weight_init_stats(self._h, out.data_ptr(), self._stream()),
                        "lsqb200_plan_weight_init_stats")
        return out

    def launches(self, backward: bool) -> int:
        return int(self._lib.lsqb200_plan_launches(self._h, int(backward)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lsqb200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
