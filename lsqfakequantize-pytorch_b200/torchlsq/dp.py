"""Data-parallel support for LSQ+ QAT: one flat gradient buffer, one all-reduce per step.

The reference has no distributed code; under DistributedDataParallel its
lazily created scale / shift parameters would be averaged like any other parameter.  Here the
only data that has to cross GPUs is tiny: every site's grad_scale / grad_shift (27 702 floats
for ResNet-50: 71 activation sites x 2 + 27 560 weight channels).  Activations shard by batch
and are never exchanged.

Design: the backward kernels write their reduced gradients STRAIGHT into slices of one flat
fp32 buffer (the C ABI takes any device pointer for gscale / gshift), so there is no gather or
copy step; one `all_reduce(SUM)` over that buffer (NCCL over NVLink/NVSwitch on GPUs, gloo in
the CPU tests) finishes the step, optionally on a side stream so it overlaps whatever follows.

Semantics: each rank scales by gs = grad_scaler / sqrt(numel_local * quant_max) with its LOCAL
numel, exactly as the reference op would inside DDP (csrc/ops/cuda/lsq_cuda.cu:124); the
all-reduced value is therefore sum_r grad(shard_r), optionally divided by the world size
(`average=True`, DDP's convention).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class FlatGradBuffer:
    """Flat fp32 buffer holding (grad_scale, grad_shift) of every site, addressed by name."""

    def __init__(self, slots: Sequence[Tuple[str, int]], device, dtype=torch.float32):
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, n in slots:
            if name in self.offsets:
                raise ValueError(f"duplicate site name {name!r}")
            if n <= 0:
                raise ValueError(f"site {name!r} needs a positive channel count")
            self.offsets[name] = (off, n)
            off += 2 * n
        self.numel = off
        self.flat = torch.zeros(off, dtype=dtype, device=device)
        self._stream: Optional[torch.cuda.Stream] = None

    def gscale(self, name: str) -> torch.Tensor:
        off, n = self.offsets[name]
        return self.flat[off:off + n]

    def gshift(self, name: str) -> torch.Tensor:
        off, n = self.offsets[name]
        return self.flat[off + n:off + 2 * n]

    def views(self, name: str):
        return self.gscale(name), self.gshift(name)

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average: bool = False, side_stream: bool = False):
        """SUM-all-reduce the flat buffer.  Returns a handle with .wait(); with `side_stream`
        on CUDA the collective is enqueued on a private stream that first waits for the
        producer (current) stream, and .wait() makes the current stream wait for it."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return _Done()
        world = dist.get_world_size(group)
        if side_stream and self.flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=self.flat.device)
            cur = torch.cuda.current_stream(self.flat.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                if average:
                    self.flat.div_(world)
            return _StreamHandle(self._stream, self.flat.device)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(world)
        return _Done()

    def scatter_to_params(self, named_params: Dict[str, Tuple[torch.nn.Parameter, Optional[torch.nn.Parameter]]]):
        """Point `.grad` of each (scale, shift) parameter pair at its slice (no copy)."""
        for name, (scale, shift) in named_params.items():
            gs, gb = self.views(name)
            scale.grad = gs.view_as(scale)
            if shift is not None:
                shift.grad = gb.view_as(shift)


class FlatLSQOptimizer:
    """One flat parameter buffer + one flat gradient buffer + ONE fused update launch for every LSQ scale / shift
    parameter of a model.

    Call it after the warm-up forward that creates the parameters (the reference's rule, README.md:101: "make a test
    forward pass BEFORE adding model parameters to optimizer").  Every `scale` / `shift` parameter is re-pointed at a
    slice of one flat fp32 buffer and its `.grad` at the matching slice of a `FlatGradBuffer`, so that

        opt.zero_grad(); loss.backward(); opt.step()

    costs one all-reduce (when torch.distributed is initialised; `average=True` folds DDP's 1/world into the update)
    and one kernel launch, whatever the number of fake-quant sites.  The update is torch.optim.SGD's / Adam's
    single-tensor arithmetic in fp32 (csrc/kern_optim.cu).  Parameters that do not require grad (symmetric shifts, static
    quantizers, scales inside their observer window) are skipped - value, optimizer state and step count - as torch.optim skips
    parameters without a gradient.  (A parameter that requires grad but was not used in this step's forward is NOT skipped: its
    slice holds zeros, which is a step with a zero gradient.)

    Gradient aliasing: every `.grad` is a VIEW of the flat gradient buffer, which is what lets autograd accumulate straight
    into it.  Use `opt.zero_grad()` (zeroes in place).  `model.zero_grad()` / another optimizer's `zero_grad()` default to
    `set_to_none=True` and detach the views - autograd then allocates fresh gradients the flat buffer never sees; `step()`
    detects that, copies such stray gradients into their slices and re-points `.grad`, so the update is still right (at
    the price of one small copy per affected parameter and step).

    Step counting is torch.optim's: every parameter has its own count (a device int32 per element), advanced only in steps in
    which the parameter takes part, and it drives Adam's bias corrections and the first-step initialisation of SGD's momentum
    buffer.  A parameter takes part while it `requires_grad` - torch.optim skips parameters without a gradient, and a quantizer in
    its observer init window has `scale.requires_grad == False` (observers.py:455-456) -, so a scale that starts learning at
    batch 1001 gets Adam's first step then, not its 1001st.  The participation mask is rebuilt only when a flag changes.
    """

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter, Optional[torch.nn.Parameter]]], kind: str = "sgd",
                 lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0, nesterov: bool = False, weight_decay: float = 0.0,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, average: bool = True, group=None):
        from . import _cabi
        if kind not in ("sgd", "adam"):
            raise ValueError("kind must be 'sgd' or 'adam'")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        named_params = list(named_params)
        if not named_params:
            raise ValueError("no LSQ parameters given (run one forward pass first: the modules create them lazily)")
        dev = named_params[0][1].device
        for name, scale, shift in named_params:
            for t in (scale, shift):
                if t is not None and (t.dtype != torch.float32 or t.device != dev or not t.is_cuda):
                    raise RuntimeError(f"site {name!r}: LSQ parameters must be float32 CUDA tensors on one device")
            if shift is not None and shift.numel() != scale.numel():
                raise RuntimeError(f"site {name!r}: scale and shift differ in length")
        self._cabi = _cabi
        self.kind, self.lr, self.momentum, self.dampening, self.nesterov = kind, lr, momentum, dampening, nesterov
        self.weight_decay, self.betas, self.eps, self.average, self.group = weight_decay, betas, eps, average, group
        self.grads = FlatGradBuffer([(name, scale.numel()) for name, scale, _ in named_params], dev)
        self.params = torch.zeros_like(self.grads.flat)
        self.sites = named_params
        for name, scale, shift in named_params:
            off, n = self.grads.offsets[name]
            ps, pb = self.params[off:off + n], self.params[off + n:off + 2 * n]
            with torch.no_grad():
                ps.copy_(scale.detach().reshape(-1))
                scale.data = ps.view_as(scale)
                scale.grad = self.grads.gscale(name).view_as(scale)
                if shift is not None:
                    pb.copy_(shift.detach().reshape(-1))
                    shift.data = pb.view_as(shift)
                    shift.grad = self.grads.gshift(name).view_as(shift)
        self._grad_views = []          # (parameter, its slice of the flat gradient buffer, viewed in the parameter's shape)
        for name, scale, shift in named_params:
            self._grad_views.append((scale, self.grads.gscale(name).view_as(scale)))
            if shift is not None:
                self._grad_views.append((shift, self.grads.gshift(name).view_as(shift)))
        self.state1 = torch.zeros_like(self.params) if (kind == "adam" or momentum != 0) else None
        self.state2 = torch.zeros_like(self.params) if kind == "adam" else None
        self.steps = 0                                                              # calls of step()
        self.step_counts = torch.zeros(self.params.numel(), dtype=torch.int32, device=dev)     # per element, as torch.optim's state['step']
        self._active = torch.ones(self.params.numel(), dtype=torch.uint8, device=dev)
        self._active_key = None

    @classmethod
    def from_model(cls, model: torch.nn.Module, **kw):
        """Collect every module that owns `scale` and `shift` Parameters (LSQFakeQuantizer), named by its module path."""
        found = []
        for name, m in model.named_modules():
            sc, sh = getattr(m, "scale", None), getattr(m, "shift", None)
            if isinstance(sc, torch.nn.Parameter) and hasattr(m, "fake_quant_enabled"):
                found.append((name or "root", sc, sh if isinstance(sh, torch.nn.Parameter) else None))
        return cls(found, **kw)

    def state_dict(self):
        """Optimizer state for checkpoint / resume (the parameters themselves live in the model's state_dict): the flat momentum /
        Adam moment buffers and the per-element step counts, keyed by site name so that a model with the same quantizers - in any
        order - can load it."""
        out = {"kind": self.kind, "steps": self.steps, "sites": {}}
        for name, (off, n) in self.grads.offsets.items():
            sl = slice(off, off + 2 * n)
            out["sites"][name] = {"step_counts": self.step_counts[sl].clone(),
                                  "state1": self.state1[sl].clone() if self.state1 is not None else None,
                                  "state2": self.state2[sl].clone() if self.state2 is not None else None}
        return out

    def load_state_dict(self, sd):
        if sd.get("kind") != self.kind:
            raise ValueError(f"optimizer kind differs: checkpoint {sd.get('kind')!r}, this optimizer {self.kind!r}")
        missing = [name for name in self.grads.offsets if name not in sd["sites"]]
        if missing:
            raise KeyError(f"sites missing from the optimizer checkpoint: {missing[:4]}")
        self.steps = int(sd.get("steps", 0))
        for name, (off, n) in self.grads.offsets.items():
            st = sd["sites"][name]
            sl = slice(off, off + 2 * n)
            if st["step_counts"].numel() != 2 * n:
                raise ValueError(f"site {name!r}: {st['step_counts'].numel() // 2} channels in the checkpoint, {n} here")
            self.step_counts[sl].copy_(st["step_counts"])
            for buf, key in ((self.state1, "state1"), (self.state2, "state2")):
                if buf is not None and st.get(key) is not None:
                    buf[sl].copy_(st[key])

    def zero_grad(self):
        """Zero the flat gradient buffer in place (autograd keeps accumulating into the same slices)."""
        self.grads.zero_()

    def _adopt_stray_grads(self):
        """`.grad` tensors that no longer alias the flat buffer (set_to_none zero_grad, manual assignment): fold them in."""
        for p, view in self._grad_views:
            g = p.grad
            if g is None:
                p.grad = view                      # the slice holds zeros (zero_grad) or this step's kernel output
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g.reshape(view.shape))  # autograd accumulated into a fresh tensor: that IS this step's gradient
                p.grad = view

    def _refresh_active(self):
        """Elements of parameters that require grad take part in the step (torch.optim: parameters with a gradient)."""
        key = tuple(p.requires_grad for p, _ in self._grad_views)
        if key == self._active_key:
            return
        mask = torch.zeros(self.params.numel(), dtype=torch.uint8)
        base = self.grads.flat.data_ptr()
        for (p, view), on in zip(self._grad_views, key):
            if on:
                off = (view.data_ptr() - base) // 4
                mask[off:off + view.numel()] = 1
        self._active.copy_(mask)
        self._active_key = key

    def step(self):
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        self._adopt_stray_grads()
        self._refresh_active()
        self.grads.all_reduce(group=self.group)            # SUM; the 1/world of DDP-style averaging is folded into the update
        self.steps += 1
        a = self._cabi.OptimArgs(float(self.lr), float(self.weight_decay), 1.0 / world if self.average else 1.0, float(self.momentum),
                                 float(self.dampening), float(self.betas[0]), float(self.betas[1]), float(self.eps), int(self.steps),
                                 self._cabi.OPT_ADAM if self.kind == "adam" else self._cabi.OPT_SGD, int(self.nesterov))
        dev = self.params.device
        with torch.cuda.device(dev):
            rc = self._cabi.load().lsqb200_flat_optimizer_step_sites(
                self.params.data_ptr(), self.grads.flat.data_ptr(), self.state1.data_ptr() if self.state1 is not None else None,
                self.state2.data_ptr() if self.state2 is not None else None, self.step_counts.data_ptr(), self._active.data_ptr(),
                self.params.numel(), a, torch.cuda.current_stream(dev).cuda_stream)
        self._cabi.check(rc, "lsqb200_flat_optimizer_step_sites")


class _Done:
    def wait(self):
        return None


class _StreamHandle:
    def __init__(self, stream, device):
        self.stream, self.device = stream, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


def shard_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's slice of a batch of n (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_bound(per_rank: List[torch.Tensor], rel: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor]:
    """(want, bound) for checking an fp32 SUM-all-reduce of W per-rank buffers element by element:
    want = the fp64 sum over ranks, bound = rel * |want| + W * 2^-24 * sum_r |shard_r|.

    An fp32 sum of W addends in ANY order is within (W - 1) * 2^-24 * sum_r |shard_r| of the exact sum; relative to the RESULT
    the error is unbounded once the shards cancel, which is why a plain `|got - want| / |want| < 1e-6` check (round 1) holds
    for W = 2 - one correctly rounded addition - and fails by construction for W >= 4."""
    stack = torch.stack([t.double() for t in per_rank])
    want = stack.sum(0)
    bound = rel * want.abs() + len(per_rank) * 2.0 ** -24 * stack.abs().sum(0)
    return want, bound


def expected_allreduced(per_rank: List[torch.Tensor], average: bool = False) -> torch.Tensor:
    """Reference semantics of the all-reduced buffer: sum over ranks (/ world if average)."""
    out = torch.stack([t.double() for t in per_rank]).sum(0)
    return out / len(per_rank) if average else out
