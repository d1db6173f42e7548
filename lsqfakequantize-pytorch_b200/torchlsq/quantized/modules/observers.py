"""`LSQFakeQuantizer` - the reference's observer/fake-quant module, restated on top of the
B200-native kernels.

API parity target: /root/reference/torchlsq/quantized/modules/observers.py:72-483 (constructor
kwargs :164-172, toggles :265-310, `_init_weights` :314-342, `_set_weights` :346-373,
`convert_shift_to_zp` :378-401, `calculate_qparams` :403-422, `forward` :424-462).

What is native here:
  * the fake-quant call itself (`torchlsq.functional.lsq` -> sm_100a kernels),
  * the mu +- 3 sigma weight initialisation: ONE fused pass (`lsqb200_weight_init_stats`) instead
    of `torch.mean` + `torch.std`,
  * the learned initialisation (`init_mode=True` through the same kernels),
  * the observer initialisation (`init_mode='observer'`, the default): one fused pass per forward
    (`lsqb200_observe`) instead of torch's `x.to(float32)` + `aminmax` + ~15 tiny kernels + host syncs.
Everything else is Python control flow with the reference's semantics.  Two deliberate fixes:
`with_args` works (the reference forgot `from functools import partial`), and the
state-machine tests (`current_batch <= n_batches`, ...) read Python-side mirrors of the buffers,
so a `.cuda()`-ed module does not stall the stream with device->host reads every forward.
"""
import inspect
from functools import partial
from math import ceil, copysign, log
from typing import Tuple

import torch

from ... import _cabi
from ...functional import lsq, lsq_add, lsq_add_relu, lsq_relu

Tensor = torch.Tensor

OTYPES = {'weight': 0, 'activation': 1}
TYPES_RANGE_MAPPING = {
    torch.qint8: {'range': (-128, 127), 'bitness': 8, 'unsigned': False},
    torch.quint8: {'range': (0, 255), 'bitness': 8, 'unsigned': True},
}
QSCHEMES = (torch.per_tensor_affine, torch.per_tensor_symmetric,
            torch.per_channel_affine, torch.per_channel_symmetric)


def _check_qscheme(qscheme):
    assert qscheme in QSCHEMES, f"Only following schemes supported {QSCHEMES} but recieved {qscheme}"


def IS_QSCHEME_PER_CHANNEL(qscheme):
    _check_qscheme(qscheme)
    return qscheme in (torch.per_channel_affine, torch.per_channel_symmetric)


def IS_QSCHEME_AFFINE(qscheme):
    _check_qscheme(qscheme)
    return qscheme in (torch.per_tensor_affine, torch.per_channel_affine)


def IS_QSCHEME_PER_TENSOR(qscheme):
    return not IS_QSCHEME_PER_CHANNEL(qscheme)


def IS_QSCHEME_SYMMETRIC(qscheme):
    return not IS_QSCHEME_AFFINE(qscheme)


class _PartialWrapper(object):
    """Picklable factory wrapper, so `LSQFakeQuantizer.with_args(...)` can sit in a QConfig."""

    def __init__(self, p):
        self.p = p

    def __call__(self, *args, **keywords):
        return self.p(*args, **keywords)

    def __repr__(self):
        return self.p.__repr__()

    def with_args(self, **kwargs):
        return _with_args(self, **kwargs)


def _with_args(cls_or_self, **kwargs):
    """Class factory: `Foo.with_args(a=1).with_args(b=2)()` builds a fresh Foo(a=1, b=2) each call."""
    return _PartialWrapper(partial(cls_or_self, **kwargs))


# bumped whenever any LSQFakeQuantizer's state-machine flags or parameters change; torchlsq.multi.group_weight_quantizers
# re-derives its grouping only when it moved
STATE_EPOCH = [0]


class ObserverBase(torch.quantization.observer.ObserverBase):
    with_args = classmethod(_with_args)


def weight_init_scale(w: Tensor, ch_axis: int, per_channel: bool, quant_min: int, quant_max: int) -> Tensor:
    """scale = max(|mu - 3 sigma|, |mu + 3 sigma|) / 2^bitness, bitness = ceil(log2(qmax - qmin)) - 1.

    Reference: observers.py:329-337 (torch.mean + unbiased torch.std over all dims but `ch_axis`).
    Here: one pass over `w` by the sm_100a statistics kernel; returns float32 [C] (or [1]).
    """
    from ...extension import _DT, _dense_layout, _stream_ptr, _workspace
    if not w.is_cuda:
        raise RuntimeError("weight initialisation needs a CUDA tensor (torchlsq-b200 has no CPU path)")
    if w.dtype not in _DT:
        raise RuntimeError(f"weights must be float32, float16 or bfloat16, got {w.dtype}")
    lib = _cabi.load()
    w = w.detach()
    if per_channel:
        wd, outer, C, inner = _dense_layout(w, ch_axis)
    else:
        wd, outer, C, inner = _dense_layout(w)
    out = torch.empty(C if per_channel else 1, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        sp = _stream_ptr(w.device)
        ws = _workspace(w.device, sp)
        rc = lib.lsqb200_weight_init_stats(wd.data_ptr(), out.data_ptr(), outer, C, inner, _DT[w.dtype],
                                           int(quant_min), int(quant_max), ws.data_ptr(), ws.numel(), sp)
    _cabi.check(rc, "lsqb200_weight_init_stats")
    return out


_NATIVE_OBSERVERS = {
    # observer class -> (per_channel, moving_average)
    torch.quantization.MinMaxObserver: (False, False),
    torch.quantization.MovingAverageMinMaxObserver: (False, True),
    torch.quantization.PerChannelMinMaxObserver: (True, False),
    torch.quantization.MovingAveragePerChannelMinMaxObserver: (True, True),
}
_SUPPORTED_OBS_QSCHEMES = (torch.per_tensor_affine, torch.per_tensor_symmetric, torch.per_channel_affine, torch.per_channel_symmetric)


def observer_step(obs, x: Tensor, scale: Tensor, shift: Tensor) -> bool:
    """One fused pass replacing `obs(x)`, `obs.calculate_qparams()` and `_set_weights(scale, zero_point)`
    (observers.py:446-449) for torch's MinMax / MovingAverageMinMax observers and their PerChannel variants:
    x is read ONCE in its own dtype (torch first materialises `x.to(float32)`, then runs `aminmax`, then ~15 tiny
    kernels and two host syncs for the qparams); the running min / max buffers of `obs` and the LSQ `scale` / `shift`
    parameters are updated in place with bit-identical fp32 arithmetic.  Returns False (and does nothing) when the
    observer is of another kind - the caller then runs the torch observer as the reference does."""
    from ...extension import _DT, _dense_layout, _stream_ptr, _workspace
    kind = _NATIVE_OBSERVERS.get(type(obs))
    if kind is None or not x.is_cuda or x.dtype not in _DT or x.numel() == 0:
        return False
    per_channel, moving = kind
    if obs.qscheme not in _SUPPORTED_OBS_QSCHEMES or obs.min_val.dtype != torch.float32:
        return False
    if scale.dtype != torch.float32 or shift.dtype != torch.float32 or scale.device != x.device:
        return False
    nparam = x.shape[obs.ch_axis] if per_channel else 1
    if scale.numel() != nparam or shift.numel() != nparam:
        return False                                  # per-tensor observer on a per-channel quantizer or vice versa
    if obs.min_val.device != x.device:
        obs.to(x.device)                              # state follows the data (the torch path would sync on every copy_)
    if obs.min_val.numel() != nparam or (per_channel and obs.min_val.dim() != 1):
        obs.min_val.resize_(nparam).fill_(float("inf"))      # torch's "never observed" state
        obs.max_val.resize_(nparam).fill_(float("-inf"))
    # the values baked into the cached argument struct; `eps` is a device buffer of the observer: its identity and version stand
    # for its value (reading it would be a device->host sync per step)
    eps = obs.eps
    key = (obs.quant_min, obs.quant_max, getattr(obs, "averaging_constant", 1.0), id(eps), eps._version if isinstance(eps, torch.Tensor) else eps,
           obs.qscheme, obs.dtype)
    cached = obs.__dict__.get("_lsqb200_args")
    cache = cached[1] if cached is not None and cached[0] == key else None
    if cache is None:
        sym = obs.qscheme in (torch.per_tensor_symmetric, torch.per_channel_symmetric)
        zp_sym = 0
        if obs.dtype in (torch.quint8, torch.uint8):
            zp_sym = (obs.quant_min + obs.quant_max) // 2 if obs.has_customized_qrange else 128
        cache = _cabi.ObserverArgs(int(obs.quant_min), int(obs.quant_max), float(getattr(obs, "averaging_constant", 1.0)),
                                   float(obs.eps), int(moving), int(sym), int(zp_sym), 0)
        obs.__dict__["_lsqb200_args"] = (key, cache)
    lib = _cabi.load()
    xd = x.detach()
    if per_channel:
        xd, outer, C, inner = _dense_layout(xd, obs.ch_axis)
    else:
        xd, outer, C, inner = _dense_layout(xd)
    with torch.cuda.device(x.device):
        sp = _stream_ptr(x.device)
        ws = _workspace(x.device, sp)
        rc = lib.lsqb200_observe(xd.data_ptr(), outer, C, inner, _DT[x.dtype], int(per_channel), obs.min_val.data_ptr(),
                                 obs.max_val.data_ptr(), scale.data_ptr(), shift.data_ptr(), cache, ws.data_ptr(), ws.numel(), sp)
    _cabi.check(rc, "lsqb200_observe")
    return True


class LSQFakeQuantizer(ObserverBase):
    """Fake-quantize module with learned step size (LSQ / LSQ+, arXiv:1902.08153, arXiv:2004.09576).

    The forward emulates quantize -> dequantize with a learnable `scale` and `shift`
    (see `torchlsq.functional.lsq` for the exact arithmetic).  `qint8` marks a WEIGHT
    quantizer (symmetric only), `quint8` an ACTIVATION quantizer.

    Parameter initialisation
      * weights: static, scale = max(|mu - 3 sigma|, |mu + 3 sigma|) / 2^(b-1) at the first call;
      * activations: during the first `init_batches` training batches either
          - `init_mode='learnable'`: scale/shift descend on ||x_r - x||^2 (the op's init mode), or
          - `init_mode='observer'`: a torch observer (MovingAverage(PerChannel)MinMaxObserver is
            the usual choice) supplies scale / zero_point, the module acts as a plain fake-quant.
    The first forward only creates the parameters and returns its input, so build the optimizer
    after one warm-up forward.

    Default ranges are 7-bit (qint8: [-64, 63], quint8: [0, 127]) to keep torch's quantized
    kernels from overflowing; pass `avoid_torch_overflow=False` for the full 8 bits.

    Args:
        observer: observer CLASS used when `init_mode='observer'`.
        otype: 'weight' or 'activation'.
        dtype: torch.quint8 (activation) or torch.qint8 (weight).
        qscheme: per_tensor_affine | per_tensor_symmetric | per_channel_affine | per_channel_symmetric.
        quant_min, quant_max: custom quantised range (must contain 0).
        init_scale, init_shift: starting values (activations / affine schemes).
        ch_axis: channel axis; default 0 for weights, 1 for activations.
        learn_params: learn scale/shift (True) or behave as a static fake-quant (False).
        init_batches: length of the activation initialisation window.
        init_mode: 'observer' or 'learnable'.
        use_grad_scaling, grad_scaler: 1/sqrt(numel * quant_max) gradient scaling and an extra factor.
        avoid_torch_overflow: 7-bit default ranges.
        debug_mode: forward is the identity.
        fuse_relu: (new, not in the reference) the module stands for `Sequential(ReLU(), LSQFakeQuantizer(...))`:
            every value it returns or observes is taken from relu(x), and once the parameters are initialised the
            ReLU runs inside the fake-quant kernels (`torchlsq.functional.lsq_relu`: one pass over x, forward and
            backward) instead of as a separate pass.  Results are bit-identical to the two-module sequence.
    """
    init_modes = ('learnable', 'observer')
    _group_out = None      # (weight, fake-quantised weight) handed over by torchlsq.multi.group_weight_quantizers for this step

    @staticmethod
    def sign(x):
        return copysign(1, x)

    def __init__(self, observer, otype,
                 dtype=torch.quint8,
                 qscheme=torch.per_tensor_affine,
                 quant_min=None, quant_max=None,
                 init_scale=1., init_shift=0.,
                 ch_axis=None, learn_params=True,
                 init_batches=1000, init_mode='observer',
                 use_grad_scaling=True, grad_scaler=1.,
                 avoid_torch_overflow=True, debug_mode=False, fuse_relu=False, **observer_kwargs):
        super().__init__(dtype)
        assert init_mode in self.init_modes, f'only following modes available: {self.init_modes}'
        self.activation_post_process = None
        if init_mode == 'observer':
            assert inspect.isclass(observer), 'awaited Observer class not instance or function wrapper'
            # hand the observer whichever of our own constructor arguments it also understands
            observer_kwargs['reduce_range'] = avoid_torch_overflow
            mine = dict(dtype=dtype, qscheme=qscheme, quant_min=quant_min, quant_max=quant_max, ch_axis=ch_axis)
            wanted = set(inspect.signature(observer.__init__).parameters) - {'self'}
            passed = {}
            for key in wanted:
                if key in mine:
                    passed[key] = mine[key]
                elif key in observer_kwargs:
                    passed[key] = observer_kwargs[key]
            if 'ch_axis' in passed and passed['ch_axis'] is None:
                passed['ch_axis'] = int(bool(OTYPES.get(otype, 1)))
            self.activation_post_process = observer(**passed)

        assert otype in OTYPES, f'otype must be on of {tuple(OTYPES.keys())}, but {otype} is given'
        self.otype = OTYPES[otype]
        assert self.dtype in TYPES_RANGE_MAPPING, \
            f"Default Observer only works for {tuple(TYPES_RANGE_MAPPING.keys())} data types"

        self.qscheme = qscheme
        self.ch_axis = int(bool(self.otype)) if ch_axis is None else ch_axis   # 0: weights, 1: activations
        self.init_mode = init_mode
        self.n_batches = init_batches
        self.use_grad_scaling = use_grad_scaling
        self.grad_scaler = grad_scaler
        self.debug_mode = debug_mode
        self.fuse_relu = bool(fuse_relu)
        self.is_perchannel = IS_QSCHEME_PER_CHANNEL(self.qscheme)
        self.is_affine = IS_QSCHEME_AFFINE(self.qscheme)
        self.init_scale = init_scale
        self.init_shift = init_shift
        self.native_observer = True          # set False to force torch's own observer kernels (debugging)
        self.quant_min, self.quant_max = self._verify_qmin_qmax(quant_min, quant_max, lowbit=avoid_torch_overflow)
        self.reset(learn_params=learn_params)

    # ------------------------------------------------------------------ ranges
    def _verify_qmin_qmax(self, quant_min: int, quant_max: int, lowbit=True) -> Tuple[int, int]:
        """Resolve the quantised range (observers.py:213-242) and, for symmetric schemes,
        the shift that centres it."""
        if self.otype == 0:
            assert not self.is_affine, 'We support only symmetric scheme for weight'
            assert self.dtype == torch.qint8, 'Pytorch quantized operations implementaion requires `qint8` type for weights'
        else:
            assert self.dtype == torch.quint8, 'Pytorch quantized operations implementaion requires `quint8` type for activation'
        bits = TYPES_RANGE_MAPPING[self.dtype]['bitness'] - int(lowbit)
        self.has_customized_qrange = (quant_min is not None) and (quant_max is not None)
        if self.has_customized_qrange:
            assert quant_min <= 0 <= quant_max, "User-specified quantization range must include 0."
            assert quant_min < quant_max, "qmin must be strictly less than qmax for user-specified quantization range."
            assert 0 < quant_max - quant_min + 1 <= 2 ** bits, \
                f"quantization range should be positive and not exceed the maximum bit range (=2^{bits})."
        else:
            quant_min, quant_max = 0, 2 ** bits - 1
            if not TYPES_RANGE_MAPPING[self.dtype]['unsigned']:
                quant_min, quant_max = quant_min - 2 ** (bits - 1), quant_max - 2 ** (bits - 1)
        if not self.is_affine:
            mid = quant_min + quant_max
            self.init_shift = -float(abs(mid) // 2) * self.sign(mid) * self.init_scale
        return quant_min, quant_max

    # ------------------------------------------------------------------ state
    @torch.jit.export
    def reset(self, learn_params=True) -> None:
        if self.otype == 0:
            self.n_batches = -1       # weights are initialised statically, no init window
        self._initialized = False
        self.register_parameter('scale', None)
        self.register_parameter('shift', None)
        self.register_buffer('fake_quant_enabled', torch.tensor([1], dtype=torch.uint8))
        self.register_buffer('observer_enabled', torch.tensor([1], dtype=torch.uint8))
        self.register_buffer('learning_enabled', torch.tensor([int(learn_params)], dtype=torch.uint8))
        self.register_buffer('current_batch', torch.tensor([0], dtype=torch.int64))
        # host mirrors of the four buffers: the state machine never reads device memory
        self._m_fq, self._m_obs, self._m_learn, self._m_batch = 1, 1, int(learn_params), 0
        STATE_EPOCH[0] += 1
        self.enable_observer()

    def sync_state(self) -> None:
        """Re-read the four state buffers (one device->host read each).  The forward's state machine runs on host mirrors
        of `fake_quant_enabled`, `observer_enabled`, `learning_enabled` and `current_batch`, kept current by the
        enable_* / disable_* methods and by `load_state_dict`; call this after writing the buffers any other way (the
        reference's style `mod.observer_enabled[0] = 0`, a DDP buffer broadcast, `load_state_dict(assign=True)`)."""
        self._sync_mirrors()

    def _sync_mirrors(self):
        self._m_fq = int(self.fake_quant_enabled[0])
        self._m_obs = int(self.observer_enabled[0])
        self._m_learn = int(self.learning_enabled[0])
        self._m_batch = int(self.current_batch[0])
        STATE_EPOCH[0] += 1

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._sync_mirrors()

    def _set_flag(self, name, mirror, value):
        getattr(self, name)[0] = value
        setattr(self, mirror, int(value))
        STATE_EPOCH[0] += 1

    def check_is_init_mode(self):
        return bool(self._m_learn) and self.otype != 0 and self._m_batch <= self.n_batches

    @torch.jit.export
    def enable_observer(self) -> None:
        on = 1
        if self._m_learn == 1:
            if self.otype == 0:
                on = 0                                   # learned weights never need the observer
            elif self.init_mode == 'learnable':
                on = 0                                   # initialised by back-prop instead
            elif self._m_batch > self.n_batches:
                on = 0                                   # observer window already over
        self._set_flag('observer_enabled', '_m_obs', on)

    @torch.jit.export
    def disable_observer(self) -> None:
        self._set_flag('observer_enabled', '_m_obs', 0)

    @torch.jit.export
    def enable_fake_quant(self) -> None:
        self._set_flag('fake_quant_enabled', '_m_fq', 1)

    @torch.jit.export
    def disable_fake_quant(self) -> None:
        self._set_flag('fake_quant_enabled', '_m_fq', 0)

    @torch.jit.export
    def enable_param_learning(self):
        """Learn scale/shift, stop observing; the init window is considered done."""
        self._set_flag('learning_enabled', '_m_learn', 1)
        self.disable_observer()
        self.n_batches = -1

    @torch.jit.export
    def enable_static_estimate(self):
        """Freeze learning, estimate scale/zero_point with the observer."""
        self._set_flag('learning_enabled', '_m_learn', 0)
        self.enable_observer()

    # ------------------------------------------------------------------ parameters
    def _init_weights(self, x: Tensor, _init_device=torch.device('cpu')) -> None:
        """Create `scale` / `shift` from the first tensor seen (observers.py:314-342).
        Parameters must be handed to the optimizer only after this first forward."""
        self._initialized = True
        STATE_EPOCH[0] += 1
        per_ch = self.is_perchannel and x is not None
        size = (x.shape[self.ch_axis] if per_ch else 1,)
        device = x.device if x is not None else _init_device
        if self.otype == 0 and x is not None:
            scale = weight_init_scale(x, self.ch_axis, per_ch, self.quant_min, self.quant_max)
        else:
            scale = torch.full(size, self.init_scale, dtype=torch.float32, device=device)
        shift = torch.full(size, self.init_shift, dtype=torch.float32, device=device)
        self.scale = torch.nn.Parameter(scale)
        self.shift = torch.nn.Parameter(shift)
        self.scale.requires_grad = bool(self._m_learn)
        self.shift.requires_grad = bool(self._m_learn) and self.is_affine

    def _set_weights(self, scale=None, shift=None, zero_point=None, _init_device=torch.device('cpu')):
        """Copy new values into the parameters; `zero_point` is converted to shift = -zp * scale
        (observers.py:346-373)."""
        if self.scale is None:
            self._init_weights(None, _init_device=_init_device)    # per-tensor placeholder
        with torch.no_grad():
            if scale is not None:
                self.scale.data.copy_(scale.to(device=self.scale.device, dtype=self.scale.dtype).reshape(self.scale.shape))
            if zero_point is not None:
                shift = -zero_point.to(self.scale.device) * self.scale.detach()
            if shift is not None:
                self.shift.data.copy_(shift.to(device=self.shift.device, dtype=self.shift.dtype).reshape(self.shift.shape))

    def set_weights(self, scale, zero_point=None, _init_device=torch.device('cpu')):
        self._set_weights(scale, shift=None, zero_point=zero_point, _init_device=_init_device)

    @staticmethod
    def convert_shift_to_zp(shift, scale, dtype):
        """zero_point = clamp(round(-shift / scale), type range) as int64 (observers.py:378-401)."""
        tmin, tmax = TYPES_RANGE_MAPPING[dtype]['range']
        with torch.no_grad():
            return (-shift / scale).round_().clamp_(min=tmin, max=tmax).to(torch.int64)

    @torch.jit.export
    def calculate_qparams(self, verbose=True, need_shift=False):
        if not self._initialized:
            if verbose:
                print("Scale and Zero Point are not initialized properly, because  LSQObserver was never called. "
                      "You must at least run model on random tensor, before calling convert! "
                      "Returned init_scale and init_zero_point")
            zp = self.convert_shift_to_zp(torch.tensor(self.init_shift), torch.tensor(self.init_scale), self.dtype).item()
            return (self.init_scale, self.init_shift, zp) if need_shift else (self.init_scale, zp)
        scale = torch.max(self.scale.detach().clone().cpu(), torch.tensor(torch.finfo(torch.float32).eps))
        shift = self.shift.detach().clone().cpu()
        zero_point = self.convert_shift_to_zp(shift, scale, self.dtype)
        return (scale, shift, zero_point) if need_shift else (scale, zero_point)

    # ------------------------------------------------------------------ forward
    def forward(self, x, x2=None, relu=None):
        """`module(x)` as in the reference.  The two optional arguments (new) are how graph rewrites hand the module the
        ops in front of it: `module(a, b)` stands for `module(a + b)`, `relu=True` for a ReLU in between (`relu=None`: the
        module's own `fuse_relu`).  See `forward_add` and `torchlsq.fusion`."""
        return self._run(x, x2, self.fuse_relu if relu is None else bool(relu))

    def forward_add(self, a, b, relu=True):
        """(new, not in the reference) The module applied to `relu(a + b)` (or `a + b`): what
        `FloatFunctional.add_relu` / `.add` do with this module as their `activation_post_process` - residual join,
        activation and fake-quant in one kernel pass once the parameters are initialised (`torchlsq.functional.lsq_add_relu`).
        Same state machine and bit-identical results as `self(torch.relu(a + b))`."""
        return self._run(a, b, bool(relu))

    # ---- hooks for torchlsq.multi.group_weight_quantizers (one launch for all weight quantizers of a model)
    def _groupable(self, w) -> bool:
        """Steady-state weight quantizer whose forward is exactly one `lsq` call on `w`."""
        return (self._initialized and not self.debug_mode and self._m_fq == 1 and self._m_obs == 0 and self.otype == 0
                and not self.fuse_relu and self.scale is not None and self.scale.dtype == torch.float32 and w.is_cuda
                and w.dtype in (torch.float32, torch.float16, torch.bfloat16) and self.scale.device == w.device
                and self.scale.numel() == (w.shape[self.ch_axis] if self.is_perchannel else 1))

    def _type_range(self):
        return TYPES_RANGE_MAPPING[self.dtype]['range']

    def _prepare_params(self):
        full_lsq = bool(self._m_learn)
        self.scale.requires_grad = full_lsq
        self.shift.requires_grad = full_lsq and self.is_affine

    def _run(self, x, x2, relu):
        grouped = self._group_out
        if grouped is not None:
            self.__dict__["_group_out"] = None
            if grouped[0] is x and x2 is None and not relu:
                return grouped[1]
        # `pending`: the prologue (residual add and / or ReLU) still has to be applied to whatever we return or observe
        pending = relu or x2 is not None

        def materialise(t):
            if x2 is not None:
                t = t + x2
            return torch.relu(t) if relu else t

        if self.debug_mode:
            return materialise(x) if pending else x
        if not self._initialized:
            if pending:
                x = materialise(x)
            self._init_weights(x)
            return x                     # the first call only creates the parameters
        backprop_init = False
        full_lsq = bool(self._m_learn)
        in_window = self._m_batch <= self.n_batches and self.training and self._m_learn == 1
        if in_window:
            if self.init_mode == 'observer':
                full_lsq = False         # plain fake-quant while the observer estimates the range
                if self._m_batch == self.n_batches:
                    full_lsq = True
                    self.disable_observer()
            else:                        # 'learnable': scale/shift descend on ||x_r - x||^2
                self.disable_observer()
                backprop_init = self._m_batch != self.n_batches
            self.current_batch[0] += 1
            self._m_batch += 1

        if self._m_obs == 1:
            if pending:                  # the observer must see the site's real input: materialise it while the range is being estimated
                x, pending = materialise(x), False
            # fused native step for torch's MinMax-family observers; any other observer runs as in the reference
            if not (self.native_observer and observer_step(self.activation_post_process, x, self.scale.data, self.shift.data)):
                self.activation_post_process(x.detach())
                scale, zero_point = self.activation_post_process.calculate_qparams()
                self._set_weights(scale=scale, zero_point=zero_point)

        if self._m_fq == 1:
            backprop_init = backprop_init and full_lsq
            tmin, tmax = TYPES_RANGE_MAPPING[self.dtype]['range']
            self.scale.requires_grad = full_lsq
            self.shift.requires_grad = full_lsq and self.is_affine
            tail = (self.scale, self.shift, self.quant_min, self.quant_max, tmin, tmax,
                    self.ch_axis, self.use_grad_scaling, self.grad_scaler, self.is_affine, self.is_perchannel)
            if pending:
                fusable = x.is_cuda and x.dtype != torch.float64 and self.scale.dtype == torch.float32
                if x2 is not None:       # broadcasting / type-promoting adds keep ATen's semantics: materialise them
                    fusable = fusable and x2.shape == x.shape and x2.dtype == x.dtype and x2.device == x.device
                if fusable:
                    # steady state: the prologue runs inside the fake-quant kernels
                    if x2 is None:
                        return lsq_relu(x, *tail, eval_mode=(not full_lsq), init_mode=bool(backprop_init))
                    return (lsq_add_relu if relu else lsq_add)(x, x2, *tail, eval_mode=(not full_lsq), init_mode=bool(backprop_init))
                x = materialise(x)
            return lsq(x, *tail, eval_mode=(not full_lsq), init_mode=bool(backprop_init))
        return materialise(x) if pending else x

    @torch.jit.export
    def extra_repr(self):
        if self.debug_mode:
            return 'Debug mode: ON, doing nothing.'
        scale, shift, zp = self.calculate_qparams(verbose=False, need_shift=True)
        head = '' if self._initialized else '(Uninitialized!) '
        if self.check_is_init_mode():
            head += (f'(Observer in parameter init mode: {self.init_mode}; '
                     f'{self._m_batch}/{self.n_batches} batches left) ')
        per_channel = f'Yes, channel axis - {self.ch_axis}' if self.is_perchannel else 'No'
        torch.set_printoptions(threshold=8)
        text = (f"{head}Observer for {'weights' if self.otype == 0 else 'activation'}; "
                f"Learnable:{bool(self._m_learn)}; Observer:{bool(self._m_obs)}; FakeQuant:{bool(self._m_fq)}; "
                f"Qtype:{self.dtype}, Affine:{self.is_affine}, PerChannel:{per_channel}, "
                f"Qrange:[{self.quant_min},{self.quant_max}], scale={scale}, zero_point={zp} (shift={shift}).")
        torch.set_printoptions(threshold=1000)
        if hasattr(self, 'recalibrated'):
            text += '\nModule was recalibrated!'
        return text
