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
class _GroupFunction(torch.autograd.Function):
    """y_i = lsq(x_i, scale_i, shift_i) for every site of a group in one launch per direction (per kernel class).

    Inputs: the group, then x_0..x_{n-1}, scale_0.., shift_0..; outputs y_0..y_{n-1}.  The backward runs once, when autograd has
    the gradient of every output, and returns every grad_x / grad_scale / grad_shift: fresh views of three flat buffers, so
    AccumulateGrad takes them over without a copy.  Bit-identical to n `torchlsq.functional.lsq` calls."""

    @staticmethod
    def forward(ctx, group, *tensors):
        n = group.n
        xs = tensors[:n]
        plan = group._plan_for(tensors)
        flat_y = torch.empty(group.total, dtype=group.dtype, device=group.device)
        ys = [flat_y[o:o + m].view(x.shape) for (o, m), x in zip(group.spans, xs)]
        plan.rebind(y=ys)
        plan.forward()
        ctx.group, ctx.plan = group, plan
        ctx.save_for_backward(*tensors)        # version-checked like any saved tensor; the plan reads their storage
        return tuple(ys)

    @staticmethod
    def backward(ctx, *grads):
        if torch.is_grad_enabled():
            raise RuntimeError("double backwards on grouped lsq not supported")
        group, plan = ctx.group, ctx.plan
        tensors = ctx.saved_tensors
        n = group.n
        xs = tensors[:n]
        gs_in = []
        for g, x in zip(grads, xs):
            if g is None:
                g = torch.zeros_like(x)
            elif g.stride() != x.stride() or g.dtype != x.dtype or g.data_ptr() % 32:
                g = torch.empty_like(x).copy_(g)
            gs_in.append(g)
        flat_gx = torch.empty(group.total, dtype=group.dtype, device=group.device)
        flat_gp = torch.empty(2 * group.nparam, dtype=group.pdtype, device=group.device)
        gxs = [flat_gx[o:o + m].view(x.shape) for (o, m), x in zip(group.spans, xs)]
        gsc = [flat_gp[o:o + c] for o, c in group.pspans]
        gsh = [flat_gp[group.nparam + o:group.nparam + o + c] for o, c in group.pspans]
        plan.rebind(grad=gs_in, gx=gxs, gscale=gsc, gshift=gsh)
        plan.backward()
        return (None, *gxs, *gsc, *gsh)


class LSQGroup:
    """Many fake-quant sites whose inputs are all known before any of them is needed - the conv / linear WEIGHTS of a
    model - quantised by one multi-tensor launch per direction instead of one launch (and one trip through the dispatcher
    and the autograd engine) per site.  `group()` returns the list of fake-quantised tensors, differentiable with respect
    to every x, scale and shift.  All sites share dtype, device and the scalar arguments below; x_i must be contiguous.

    The plan (device-side table) is built at the first call and rebuilt only when an input tensor's storage moves."""

    def __init__(self, xs, scales, shifts, quant_min=-128, quant_max=127, type_min=None, type_max=None, axis=0,
                 use_grad_scaling=True, grad_scaler=1.0, is_affine=False, is_perchannel=True, eval_mode=False, init_mode=False):
        self.xs, self.scales, self.shifts = list(xs), list(scales), list(shifts)
        self.n = len(self.xs)
        if self.n == 0 or len(self.scales) != self.n or len(self.shifts) != self.n:
            raise ValueError("LSQGroup needs as many scale / shift tensors as inputs (and at least one)")
        self.q = dict(quant_min=quant_min, quant_max=quant_max, type_min=type_min, type_max=type_max, axis=axis,
                      use_grad_scaling=use_grad_scaling, grad_scaler=grad_scaler, is_affine=is_affine,
                      is_perchannel=is_perchannel, eval_mode=eval_mode, init_mode=init_mode)
        x0 = self.xs[0]
        self.dtype, self.device, self.pdtype = x0.dtype, x0.device, self.scales[0].dtype
        self.spans, self.pspans = [], []
        off = poff = 0
        for x, sc, sh in zip(self.xs, self.scales, self.shifts):
            if x.dtype != self.dtype or x.device != self.device or not x.is_contiguous():
                raise RuntimeError("LSQGroup inputs must share dtype and device and be contiguous")
            if sc.dtype != self.pdtype or sh.dtype != self.pdtype:
                raise RuntimeError("LSQGroup scale / shift tensors must share one dtype")
            m = x.numel()
            self.spans.append((off, m))
            off += -(-m // 16) * 16                  # keep every view 32-byte aligned for 2- and 4-byte elements
            c = sc.numel()
            self.pspans.append((poff, c))
            poff += -(-c // 8) * 8
        self.total, self.nparam = off, poff
        self._plan, self._key = None, None

    def _plan_for(self, tensors):
        key = tuple(t.data_ptr() for t in tensors)
        if self._plan is None or key != self._key:
            if self._plan is not None:
                self._plan.close()
            n = self.n
            xs, scs, shs = tensors[:n], tensors[n:2 * n], tensors[2 * n:]
            # placeholders with the alignment of the per-step buffers (rebind() swaps the real ones in before every run)
            fy = torch.empty(self.total, dtype=self.dtype, device=self.device)
            fp = torch.empty(2 * self.nparam, dtype=self.pdtype, device=self.device)
            sites = []
            for (o, m), (po, c), x, sc, sh in zip(self.spans, self.pspans, xs, scs, shs):
                v = fy[o:o + m].view(x.shape)
                sites.append(Site(x=x.detach(), scale=sc.detach(), shift=sh.detach(), y=v, grad=v, gx=v, gscale=fp[po:po + c],
                                  gshift=fp[self.nparam + po:self.nparam + po + c], **self.q))
            self._plan, self._key = LSQPlan(sites), key
        return self._plan

    def __call__(self):
        return list(_GroupFunction.apply(self, *self.xs, *self.scales, *self.shifts))

    def launches(self):
        return 0 if self._plan is None else self._plan.launches(False) + self._plan.launches(True)


def group_weight_quantizers(model: torch.nn.Module):
    """Model-level helper for `prepare_qat` models: route every weight quantizer (`<module>.weight_fake_quant`, an
    LSQFakeQuantizer applied to `<module>.weight`, as torch's QAT Conv / Linear modules do) through ONE multi-tensor
    launch per direction.  A forward pre-hook on `model` quantises all weights up front; each quantizer then hands out
    its share when its module calls it.  Same results bit for bit; modules whose quantizer is not in steady state yet
    (parameters not created, fake-quant disabled, debug mode) or that transform the weight first (fused Conv-BN) simply keep
    running on their own.  Returns the hook handle (`.remove()` undoes the grouping)."""
    from .quantized.modules.observers import LSQFakeQuantizer
    pairs = []
    for m in model.modules():
        q = getattr(m, "weight_fake_quant", None)
        w = getattr(m, "weight", None)
        if isinstance(q, LSQFakeQuantizer) and isinstance(w, torch.nn.Parameter) and q.otype == 0:
            pairs.append((q, w))
    state = {"group": None, "sig": None}

    def pre_hook(_module, _args):
        ready = [(q, w) for q, w in pairs if q._groupable(w)]
        for q, _ in pairs:
            q._group_out = None
        if len(ready) < 2:
            return
        q0 = ready[0][0]
        sig = tuple((id(q), id(w), id(q.scale), id(q.shift), bool(q._m_learn)) for q, w in ready)
        if sig != state["sig"]:
            qs = [q for q, _ in ready]
            if any((q.quant_min, q.quant_max, q.ch_axis, q.use_grad_scaling, q.grad_scaler, q.is_affine, q.is_perchannel, q.dtype,
                    q._m_learn) != (q0.quant_min, q0.quant_max, q0.ch_axis, q0.use_grad_scaling, q0.grad_scaler, q0.is_affine,
                                    q0.is_perchannel, q0.dtype, q0._m_learn) for q in qs) or \
               any(w.dtype != ready[0][1].dtype or w.device != ready[0][1].device or not w.is_contiguous() for _, w in ready):
                return                                  # heterogeneous quantizers: leave them alone
            tmin, tmax = q0._type_range()
            state["group"] = LSQGroup([w for _, w in ready], [q.scale for q in qs], [q.shift for q in qs], q0.quant_min, q0.quant_max,
                                      tmin, tmax, q0.ch_axis, q0.use_grad_scaling, q0.grad_scaler, q0.is_affine, q0.is_perchannel,
                                      eval_mode=not bool(q0._m_learn), init_mode=False)
            state["sig"] = sig
        for q, _ in ready:
            q._prepare_params()
        for (q, w), y in zip(ready, state["group"]()):
            q._group_out = (w, y)

    return model.register_forward_pre_hook(pre_hook)
