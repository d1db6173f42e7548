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
from .extension import _DT_OPS as _DT, _dense_layout, _stream_ptr


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
        self._bound = {}          # tensors handed to rebind(): kept alive while the device table points at them
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(self._lib.lsqb200_plan_create(segs, len(sites), ctypes.byref(handle)), "lsqb200_plan_create")
        self._h = handle

    def _stream(self):
        return _stream_ptr(self.device)

    def rebind(self, y=None, grad=None, gx=None, gscale=None, gshift=None):
        """Re-point the per-step tensors (lists with one tensor per site, or None to keep) - fresh outputs and the
        upstream gradients of a training step - without rebuilding the plan: a tiny patch kernel on the current stream,
        no allocation, no synchronisation.  Layout, dtype and alignment (32 bytes: any torch allocation) must be those
        of the tensors the plan was created with; the plan keeps the new tensors alive."""
        n = len(self.sites)
        arrs = []
        for name, ts in (("y", y), ("grad", grad), ("gx", gx), ("gscale", gscale), ("gshift", gshift)):
            if ts is None:
                arrs.append(None)
                continue
            if len(ts) != n:
                raise ValueError(f"rebind: {name} needs one tensor per site")
            arrs.append((ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]))
            self._bound[name] = ts
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lsqb200_plan_rebind(self._h, *arrs, self._stream()), "lsqb200_plan_rebind")

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


# ---- grouped fake-quant of many tensors behind ONE autograd node --------------------------------------------------------------
_group_ids = iter(range(1, 1 << 62))


class LSQGroup:
    """Many fake-quant sites whose inputs are all known before any of them is needed - the conv / linear WEIGHTS of a
    model - quantised by one multi-tensor launch per direction instead of one launch (and one trip through the dispatcher
    and the autograd engine) per site.  `group()` returns the list of fake-quantised tensors, differentiable with respect
    to every x, scale and shift: ONE autograd node (`torch.ops.torchlsq.lsq_group`, csrc/torch_binding.cpp) whose backward
    runs when autograd has the gradient of every output and hands back every grad_x / grad_scale / grad_shift as views of
    three flat buffers (AccumulateGrad takes them over without a copy).  Outputs and grad_x are bit-identical to n
    `torchlsq.functional.lsq` calls; so are grad_scale / grad_shift of weight rows (same kernels, same order), while a per-tensor
    site may be cut into other tiles inside a plan than by a single launch - the same terms summed in another fixed fp64 order,
    i.e. equal to ~1e-15 before the final rounding.
    All sites share dtype, device and the scalar arguments below; x_i must be contiguous.

    The plan (device-side descriptor table) is built at the first call and rebuilt only when an input's storage moves; the
    per-step tensors are swapped in by `lsqb200_plan_rebind` (include/lsq_b200.h)."""

    def __init__(self, xs, scales, shifts, quant_min=-128, quant_max=127, type_min=None, type_max=None, axis=0,
                 use_grad_scaling=True, grad_scaler=1.0, is_affine=False, is_perchannel=True, eval_mode=False, init_mode=False):
        from .extension import _assert_has_ops
        _assert_has_ops()
        self.xs, self.scales, self.shifts = list(xs), list(scales), list(shifts)
        self.n = len(self.xs)
        if self.n == 0 or len(self.scales) != self.n or len(self.shifts) != self.n:
            raise ValueError("LSQGroup needs as many scale / shift tensors as inputs (and at least one)")
        if not is_affine:
            assert quant_min <= 0 <= quant_max, 'quantization range must be covered 0 in symmetric quantization'
        self._args = (int(quant_min), int(quant_max), int(quant_min if type_min is None else type_min),
                      int(quant_max if type_max is None else type_max), int(axis), bool(use_grad_scaling), float(grad_scaler),
                      bool(is_affine), bool(is_perchannel), bool(eval_mode), bool(init_mode))
        x0 = self.xs[0]
        for x, sc, sh in zip(self.xs, self.scales, self.shifts):
            if x.dtype != x0.dtype or x.device != x0.device or not x.is_contiguous():
                raise RuntimeError("LSQGroup inputs must share dtype and device and be contiguous")
            if sc.dtype != self.scales[0].dtype or sh.dtype != self.scales[0].dtype:
                raise RuntimeError("LSQGroup scale / shift tensors must share one dtype")
        self._id = next(_group_ids)
        self._op = torch.ops.torchlsq.lsq_group

    def __call__(self):
        return self._op(self.xs, self.scales, self.shifts, *self._args, self._id)

    def info(self):
        """(how often the device-side plan was (re)built, kernel launches of one forward + one backward)."""
        gen, launches = torch.ops.torchlsq.lsq_group_info(self._id)
        return gen, launches

    def launches(self):
        return self.info()[1]

    def close(self):
        if getattr(self, "_id", None) is not None:
            try:
                torch.ops.torchlsq.lsq_group_release(self._id)
            except Exception:
                pass
            self._id = None

    def __del__(self):
        self.close()


def group_weight_quantizers(model: torch.nn.Module):
    """Model-level helper for `prepare_qat` models: route every weight quantizer (`<module>.weight_fake_quant`, an
    LSQFakeQuantizer applied to `<module>.weight`, as torch's QAT Conv / Linear modules do) through ONE multi-tensor
    launch per direction.  A forward pre-hook on `model` quantises all weights up front; each quantizer then hands out
    its share when its module calls it.  Same results bit for bit; modules whose quantizer is not in steady state yet
    (parameters not created, fake-quant disabled, debug mode) or that transform the weight first (fused Conv-BN) simply keep
    running on their own; the grouping is re-derived whenever a quantizer's state changes (enable_* / disable_* calls,
    parameter creation, load_state_dict), not per step.  Attributes the state machine does not own (`debug_mode`,
    `fuse_relu`, ranges) are read when the group is formed: call the helper again after changing them.
    Returns the hook handle (`.remove()` undoes the grouping)."""
    from .quantized.modules.observers import LSQFakeQuantizer
    pairs = []
    for m in model.modules():
        q = getattr(m, "weight_fake_quant", None)
        w = getattr(m, "weight", None)
        if isinstance(q, LSQFakeQuantizer) and isinstance(w, torch.nn.Parameter) and q.otype == 0:
            pairs.append((q, w))
    from .quantized.modules import observers as _obs
    state = {"group": None, "ready": (), "epoch": -1}

    def rebuild():
        """Quantizers in steady state with identical settings form the group; the rest keep running on their own."""
        state["group"], state["ready"] = None, ()
        ready = [(q, w) for q, w in pairs if q._groupable(w)]
        if len(ready) < 2:
            return
        q0, w0 = ready[0]
        same = lambda q: (q.quant_min, q.quant_max, q.ch_axis, q.use_grad_scaling, q.grad_scaler, q.is_affine, q.is_perchannel, q.dtype,
                          q._m_learn)
        ready = [(q, w) for q, w in ready if same(q) == same(q0) and w.dtype == w0.dtype and w.device == w0.device and w.is_contiguous()]
        if len(ready) < 2:
            return
        for q, _ in ready:
            q._prepare_params()
        tmin, tmax = q0._type_range()
        state["group"] = LSQGroup([w for _, w in ready], [q.scale for q, _ in ready], [q.shift for q, _ in ready], q0.quant_min,
                                  q0.quant_max, tmin, tmax, q0.ch_axis, q0.use_grad_scaling, q0.grad_scaler, q0.is_affine,
                                  q0.is_perchannel, eval_mode=not bool(q0._m_learn), init_mode=False)
        state["ready"] = tuple(ready)

    def pre_hook(_module, _args):
        # any flag change / parameter creation of any LSQFakeQuantizer bumps the epoch: only then is the grouping re-derived,
        # so the steady-state cost of the hook is one op call plus one dict store per quantizer
        if state["epoch"] != _obs.STATE_EPOCH[0]:
            for q, _ in pairs:
                q.__dict__["_group_out"] = None
            rebuild()
            state["epoch"] = _obs.STATE_EPOCH[0]
        group = state["group"]
        if group is None:
            return
        for (q, w), y in zip(state["ready"], group()):
            q.__dict__["_group_out"] = (w, y)

    return model.register_forward_pre_hook(pre_hook)
