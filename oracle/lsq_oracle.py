"""Python face of the CPU oracle (oracle/lsq_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (lsqfakequantize-pytorch_b200/torchlsq) never does.

numpy in / numpy out.  16-bit tensors travel as uint16 bit patterns (`to_bits` / `from_bits`
convert from / to torch tensors).  See the C file's header for the reference lines each
function follows and for how the oracle is pinned.
"""
import ctypes
import subprocess
from ctypes import POINTER, Structure, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
F32, F16, BF16 = 0, 1, 2
CONTRACT_CPU, CONTRACT_CUDA = 0, 3   # reference CPU build (no fusion) / reference CUDA build (v and d fused)


class Cfg(Structure):
    _fields_ = [("quant_min", c_int64), ("quant_max", c_int64), ("type_min", c_int64), ("type_max", c_int64),
                ("grad_scaler", c_double), ("use_grad_scaling", c_int32), ("sym", c_int32), ("eval_mode", c_int32),
                ("init_mode", c_int32), ("contract", c_int32), ("numel_div_c", c_int32), ("half_exact", c_int32),
                ("gs_per_term", c_int32)]


_lib = None
_ref = None


def build(force=False):
    so = HERE / "liblsq_oracle.so"
    if force or not so.exists() or so.stat().st_mtime < (HERE / "lsq_oracle.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "liblsq_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        _lib.lsq_oracle_forward.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                            c_int, POINTER(Cfg)]
        _lib.lsq_oracle_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(Cfg)]
        _lib.lsq_oracle_forward_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(Cfg)]
        _lib.lsq_oracle_backward_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(Cfg)]
        _lib.lsq_oracle_weight_init.argtypes = [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64]
        _lib.lsq_oracle_qparams.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]
        _lib.lsq_oracle_quantize.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                             c_int, POINTER(Cfg), c_int]
        _lib.lsq_oracle_dequantize.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                               c_int, POINTER(Cfg), c_int]
        _lib.lsq_oracle_h2f.restype = c_float
        _lib.lsq_oracle_h2f.argtypes = [ctypes.c_uint16]
        _lib.lsq_oracle_f2h.restype = ctypes.c_uint16
        _lib.lsq_oracle_f2h.argtypes = [c_float]
        _lib.lsq_oracle_f2bf.restype = ctypes.c_uint16
        _lib.lsq_oracle_f2bf.argtypes = [c_float]
    return _lib


def cfg(quant_min=0, quant_max=255, type_min=None, type_max=None, use_grad_scaling=True, grad_scaler=1.0,
        sym=False, eval_mode=False, init_mode=False, contract=CONTRACT_CUDA, numel_div_c=False, half_exact=False,
        gs_per_term=True):
    type_min = quant_min if type_min is None else type_min
    type_max = quant_max if type_max is None else type_max
    return Cfg(int(quant_min), int(quant_max), int(type_min), int(type_max), float(grad_scaler),
               int(use_grad_scaling), int(sym), int(eval_mode), int(init_mode), int(contract), int(numel_div_c),
               int(half_exact), int(gs_per_term))


def _dt_of(a: np.ndarray, dt):
    if dt is not None:
        return dt
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.float16:
        return F16
    raise TypeError("pass dt=BF16/F16 explicitly for uint16 bit patterns")


def _raw(a: np.ndarray):
    """fp16 numpy arrays are reinterpreted as their uint16 bit pattern."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint16) if a.dtype == np.float16 else a


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _params(scale, shift):
    s = np.ascontiguousarray(np.asarray(scale, dtype=np.float32).reshape(-1))
    b = np.ascontiguousarray(np.asarray(shift, dtype=np.float32).reshape(-1))
    return s, b


def _params64(scale, shift):
    s = np.ascontiguousarray(np.asarray(scale, dtype=np.float64).reshape(-1))
    b = np.ascontiguousarray(np.asarray(shift, dtype=np.float64).reshape(-1))
    return s, b


def forward(x, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None):
    """y with x's dtype / bit pattern.  x is viewed as contiguous (outer, C, inner).
    float64 x: float64 scale / shift; c.contract == CONTRACT_CPU restates the reference CPU build (all double),
    anything else the reference CUDA build (clamps through float, fused v and d)."""
    if x.dtype == np.float64:
        xr = np.ascontiguousarray(x)
        inner = xr.size // (outer * C) if inner is None else inner
        assert outer * C * inner == xr.size
        s, b = _params64(scale, shift)
        y = np.empty_like(xr)
        lib().lsq_oracle_forward_f64(_p(xr), _p(y), _p(s), _p(b), outer, C, inner, int(per_channel), ctypes.byref(c))
        return y
    dt = _dt_of(x, dt)
    xr = _raw(x)
    inner = xr.size // (outer * C) if inner is None else inner
    assert outer * C * inner == xr.size
    s, b = _params(scale, shift)
    y = np.empty_like(xr)
    lib().lsq_oracle_forward(_p(xr), _p(y), dt, _p(s), _p(b), outer, C, inner, int(per_channel), ctypes.byref(c))
    return y.view(x.dtype) if x.dtype == np.float16 else y


def backward(g, x, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, with_abs=False):
    """(gx, gscale float64[C|1], gshift float64[C|1]): gx in x's dtype; the sums are the exact
    (double) sums of the reference's fp32 per-element terms.  with_abs=True appends the sums of
    |terms| (x |gs|), the natural yardstick for the error of any fp32 summation order.
    float64 x: the reference's double terms summed in long double."""
    if x.dtype == np.float64:
        xr, gr = np.ascontiguousarray(x), np.ascontiguousarray(g, dtype=np.float64)
        inner = xr.size // (outer * C) if inner is None else inner
        assert outer * C * inner == xr.size == gr.size
        s, b = _params64(scale, shift)
        nslot = C if per_channel else 1
        gx = np.empty_like(xr)
        gs_, gb_ = np.zeros(nslot, np.float64), np.zeros(nslot, np.float64)
        as_, ab_ = np.zeros(nslot, np.float64), np.zeros(nslot, np.float64)
        lib().lsq_oracle_backward_f64(_p(gr), _p(xr), _p(gx), _p(s), _p(b), _p(gs_), _p(gb_), _p(as_), _p(ab_),
                                      outer, C, inner, int(per_channel), ctypes.byref(c))
        return (gx, gs_, gb_, as_, ab_) if with_abs else (gx, gs_, gb_)
    dt = _dt_of(x, dt)
    xr, gr = _raw(x), _raw(g)
    inner = xr.size // (outer * C) if inner is None else inner
    assert outer * C * inner == xr.size == gr.size
    s, b = _params(scale, shift)
    nslot = C if per_channel else 1
    gx = np.empty_like(xr)
    gs_, gb_ = np.zeros(nslot, np.float64), np.zeros(nslot, np.float64)
    as_, ab_ = np.zeros(nslot, np.float64), np.zeros(nslot, np.float64)
    lib().lsq_oracle_backward(_p(gr), _p(xr), _p(gx), dt, _p(s), _p(b), _p(gs_), _p(gb_), _p(as_), _p(ab_),
                              outer, C, inner, int(per_channel), ctypes.byref(c))
    gx = gx.view(x.dtype) if x.dtype == np.float16 else gx
    return (gx, gs_, gb_, as_, ab_) if with_abs else (gx, gs_, gb_)


# ---- prologue fusion: lsq(relu(x)) restated as the reference's own call sequence ---------------------------
#      torch.relu (ATen clamp_min(x, 0): NaN stays NaN, negatives and -0 become +0) -> reference lsq op -> autograd through
#      both (relu backward = threshold_backward: x <= 0 -> exact +0, otherwise the incoming gradient).  Pinned by
#      tests/golden/ref_cpu_relu.npz (reference CPU op behind torch.relu, tests/golden/make_golden_relu.py).
def _relu_masks(x, dt):
    """(is_nan, nonpositive) of x given as fp32 / fp16 values or 16-bit patterns."""
    xr = _raw(x)
    if xr.dtype == np.uint16:
        mag = xr & np.uint16(0x7FFF)
        expo = np.uint16(0x7C00 if dt == F16 else 0x7F80)
        nan = mag > expo
        nonpos = ~nan & ((mag == 0) | ((xr >> 15) == 1))
        return xr, nan, nonpos
    nan = np.isnan(xr)
    with np.errstate(invalid="ignore"):
        nonpos = ~nan & (xr <= 0)
    return xr, nan, nonpos


def relu(x, dt=None, keep_neg_zero=False):
    """torch.relu on values or 16-bit patterns; returns the same representation as `x`.
    -0 input: ATen's CUDA kernel (fmaxf -> FMNMX, -0 < +0) and the sm_100a kernels return +0; ATen's CPU kernel
    returns the -0 it was given (keep_neg_zero=True).  Only the learned-init forward (y = relu(x), a copy) can show
    the difference; everything behind the fake-quant's fma(x, 1/s, zp) is blind to the sign of zero."""
    dt = _dt_of(x, dt) if x.dtype != np.float64 else None
    xr, _, nonpos = _relu_masks(x, dt)
    out = xr.copy()
    if keep_neg_zero:
        nonpos = nonpos & ~((xr & np.uint16(0x7FFF)) == 0 if xr.dtype == np.uint16 else (xr == 0))
    out[nonpos] = 0          # +0 in every representation
    return out.view(x.dtype) if x.dtype == np.float16 else out


def add(x, x2, dt=None):
    """x + x2 as ATen's add stores it: fp32 sum, one rounding to the tensor type (values or 16-bit patterns in and out)."""
    if x.dtype in (np.float32, np.float64):
        with np.errstate(invalid="ignore", over="ignore"):
            return (x + x2).astype(x.dtype)
    dt = _dt_of(x, dt)
    import torch
    sh = np.shape(x)
    out, _ = to_bits(from_bits(_raw(x).reshape(-1), dt) + from_bits(_raw(x2).reshape(-1), dt))
    out = out.reshape(sh)
    return out.view(np.float16) if x.dtype == np.float16 else out


def forward_add(x, x2, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, with_relu=True):
    """lsq(relu(x + x2)) (with_relu) or lsq(x + x2): the reference op behind ATen's add [+ relu]."""
    s = add(x, x2, dt)
    return (forward_relu if with_relu else forward)(s, scale, shift, c, outer, C, inner, per_channel, dt)


def backward_add(g, x, x2, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, with_abs=False, with_relu=True):
    """Backward of forward_add; the returned gx is the gradient of x and of x2 (add's backward passes it to both)."""
    s = add(x, x2, dt)
    return (backward_relu if with_relu else backward)(g, s, scale, shift, c, outer, C, inner, per_channel, dt, with_abs)


def forward_relu(x, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None):
    return forward(relu(x, dt, keep_neg_zero=(c.contract == CONTRACT_CPU)), scale, shift, c, outer, C, inner, per_channel, dt)


def backward_relu(g, x, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, with_abs=False):
    out = backward(g, relu(x, dt), scale, shift, c, outer, C, inner, per_channel, dt, with_abs)
    dtc = _dt_of(x, dt) if x.dtype != np.float64 else None
    _, _, nonpos = _relu_masks(x, dtc)
    gx = _raw(out[0]).copy().reshape(-1)
    gx[nonpos.reshape(-1)] = 0
    gx = gx.view(x.dtype) if x.dtype == np.float16 else gx
    return (gx,) + tuple(out[1:])


def weight_init(w, quant_min, quant_max, outer=1, C=1, inner=None, dt=None):
    dt = _dt_of(w, dt)
    wr = _raw(w)
    inner = wr.size // (outer * C) if inner is None else inner
    out = np.empty(C, np.float32)
    lib().lsq_oracle_weight_init(_p(wr), dt, _p(out), outer, C, inner, int(quant_min), int(quant_max))
    return out


SEM_LSQ, SEM_TORCH_CUDA, SEM_TORCH_CPU = 0, 1, 2


def qparams(scale, shift, type_min, type_max):
    """LSQFakeQuantizer.calculate_qparams(): (max(scale, eps) float32[n], zero_point int64[n])."""
    s, b = _params(scale, shift)
    so, zo = np.empty_like(s), np.empty(s.size, np.int64)
    lib().lsq_oracle_qparams(_p(s), _p(b), _p(so), _p(zo), s.size, int(type_min), int(type_max))
    return so, zo


def quantize(x, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, sem=SEM_LSQ):
    """int32 codes, one per element of x viewed as contiguous (outer, C, inner)."""
    dt = _dt_of(x, dt)
    xr = _raw(x)
    inner = xr.size // (outer * C) if inner is None else inner
    assert outer * C * inner == xr.size
    s, b = _params(scale, shift)
    codes = np.empty(xr.size, np.int32)
    lib().lsq_oracle_quantize(_p(xr), dt, _p(codes), _p(s), _p(b), outer, C, inner, int(per_channel), ctypes.byref(c), int(sem))
    return codes


def dequantize(codes, like, scale, shift, c: Cfg, outer=1, C=1, inner=None, per_channel=False, dt=None, sem=SEM_LSQ):
    """(code - zp) * s stored in `like`'s dtype / bit pattern."""
    dt = _dt_of(like, dt)
    lr = _raw(like)
    codes = np.ascontiguousarray(codes, dtype=np.int32).reshape(-1)
    inner = codes.size // (outer * C) if inner is None else inner
    assert outer * C * inner == codes.size == lr.size
    s, b = _params(scale, shift)
    y = np.empty_like(lr)
    lib().lsq_oracle_dequantize(_p(codes), _p(y), dt, _p(s), _p(b), outer, C, inner, int(per_channel), ctypes.byref(c), int(sem))
    return y.view(like.dtype) if like.dtype == np.float16 else y


# ---- torch <-> bit-pattern helpers (tests only) ---------------------------------------------
def to_bits(t):
    """torch tensor (fp32 / fp16 / bf16, any device) -> (numpy array, dt code)."""
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy(), F32
    if t.dtype == torch.float16:
        return t.view(torch.int16).numpy().view(np.uint16), F16
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16), BF16
    raise TypeError(t.dtype)


def from_bits(a, dt, shape=None):
    import torch
    if dt == F32:
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.float16 if dt == F16 else torch.bfloat16)
    return t.reshape(shape) if shape is not None else t


# ---- the reference's own scalar templates, when oracle/_ref/libref_scalar.so was built ----------
def ref_scalar():
    """ctypes handle of oracle/_ref/libref_scalar.so (reference lsq_kernel.h compiled as it lies), or None."""
    global _ref
    if _ref is None:
        so = HERE / "_ref" / "libref_scalar.so"
        if not so.exists():
            return None
        _ref = ctypes.CDLL(str(so))
        _ref.ref_fwd_tensor_f32.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_float, c_int64, c_int64, c_int64,
                                            c_int64, c_int]
        _ref.ref_bwd_tensor_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                            c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_int]
        _ref.ref_fwd_channel_f32.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                             c_int64, c_int64, c_int64, c_int64, c_int]
        _ref.ref_bwd_channel_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                             c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_int]
    return _ref
