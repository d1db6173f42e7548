"""ctypes binding of libtorchlsq_b200.so (C ABI declared in include/lsq_b200.h).

This is the only place Python touches the native library.  There is NO fallback: if the
library is missing or a call fails, the caller gets an exception (north_star: "no CPU
fallback"; a silent eager path would void every parity claim).
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32,
                    c_int64, c_size_t, c_void_p)
from pathlib import Path

LIB_NAME = "libtorchlsq_b200.so"
ABI_VERSION = 3          # LSQB200_ABI_VERSION of include/lsq_b200.h these prototypes were written against
F32, F16, BF16, F64 = 0, 1, 2, 3
SEM_LSQ, SEM_TORCH, SEM_TORCH_CPU = 0, 1, 2
PRE_NONE, PRE_RELU, PRE_ADD_RELU, PRE_ADD = 0, 1, 2, 3   # fused prologue in front of the fake-quant (lsqb200_*_pre)


class QArgs(Structure):
    """lsqb200_qargs -- scalar arguments in the reference's schema order (csrc/ops/lsq.cpp:138)."""
    _fields_ = [("quant_min", c_int64), ("quant_max", c_int64), ("type_min", c_int64), ("type_max", c_int64),
                ("grad_scaler", c_double), ("use_grad_scaling", c_int32), ("sym", c_int32),
                ("eval_mode", c_int32), ("init_mode", c_int32)]


class Segment(Structure):
    """lsqb200_segment -- one fake-quant site of a multi-tensor plan."""
    _fields_ = [("x", c_void_p), ("y", c_void_p), ("grad", c_void_p), ("gx", c_void_p),
                ("scale", c_void_p), ("shift", c_void_p), ("gscale", c_void_p), ("gshift", c_void_p),
                ("outer", c_int64), ("C", c_int64), ("inner", c_int64),
                ("xdtype", c_int32), ("pdtype", c_int32), ("per_channel", c_int32), ("prologue", c_int32),
                ("q", QArgs), ("x2", c_void_p)]


class ObserverArgs(Structure):
    """lsqb200_observer_args -- what torch's MinMax / MovingAverageMinMax observers need."""
    _fields_ = [("quant_min", c_int64), ("quant_max", c_int64), ("averaging_constant", c_double), ("eps", c_double),
                ("moving_average", c_int32), ("symmetric", c_int32), ("zero_point_sym", c_int32), ("reserved", c_int32)]


class OptimArgs(Structure):
    """lsqb200_optim_args -- torch.optim SGD / Adam hyper-parameters of one fused flat-buffer step."""
    _fields_ = [("lr", c_double), ("weight_decay", c_double), ("grad_mul", c_double), ("momentum", c_double),
                ("dampening", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double),
                ("step", c_int64), ("kind", c_int32), ("nesterov", c_int32)]


OPT_SGD, OPT_ADAM = 0, 1


class LaunchInfo(Structure):
    _fields_ = [("regime", c_int32), ("vec", c_int32), ("threads", c_int32), ("splits", c_int32),
                ("grid", c_int64), ("units_per_split", c_int64)]


_PROTOTYPES = {
    "lsqb200_abi_version": (c_int, []),
    "lsqb200_cuda_version": (c_int64, []),
    "lsqb200_last_error": (c_char_p, []),
    "lsqb200_workspace_bytes": (c_size_t, []),
    "lsqb200_fwd_tensor": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                   POINTER(QArgs), c_void_p]),
    "lsqb200_bwd_tensor": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_int, c_int, POINTER(QArgs), c_void_p, c_size_t, c_void_p]),
    "lsqb200_fwd_channel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                    c_int, c_int, POINTER(QArgs), c_void_p]),
    "lsqb200_bwd_channel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int64, c_int, c_int, POINTER(QArgs),
                                    c_void_p, c_size_t, c_void_p]),
    "lsqb200_fwd_tensor_pre": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                       POINTER(QArgs), c_int, c_void_p]),
    "lsqb200_bwd_tensor_pre": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int, c_int, POINTER(QArgs), c_int, c_void_p, c_size_t, c_void_p]),
    "lsqb200_fwd_channel_pre": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                        c_int, c_int, POINTER(QArgs), c_int, c_void_p]),
    "lsqb200_bwd_channel_pre": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int64, c_int, c_int, POINTER(QArgs), c_int,
                                        c_void_p, c_size_t, c_void_p]),
    "lsqb200_weight_init_stats": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int,
                                          c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "lsqb200_observe": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                POINTER(ObserverArgs), c_void_p, c_size_t, c_void_p]),
    "lsqb200_quantize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                 POINTER(QArgs), c_int, c_int, c_void_p]),
    "lsqb200_dequantize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                   POINTER(QArgs), c_int, c_int, c_void_p]),
    "lsqb200_qparams": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_void_p]),
    "lsqb200_flat_optimizer_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(OptimArgs), c_void_p]),
    "lsqb200_flat_optimizer_step_sites": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(OptimArgs),
                                                  c_void_p]),
    "lsqb200_plan_create": (c_int, [POINTER(Segment), c_int32, POINTER(c_void_p)]),
    "lsqb200_plan_rebind": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                    POINTER(c_void_p), c_void_p]),
    "lsqb200_plan_forward": (c_int, [c_void_p, c_void_p]),
    "lsqb200_plan_backward": (c_int, [c_void_p, c_void_p]),
    "lsqb200_plan_weight_init_stats": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lsqb200_plan_destroy": (c_int, [c_void_p]),
    "lsqb200_plan_launches": (c_int, [c_void_p, c_int]),
    "lsqb200_query_launch": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, c_int, POINTER(LaunchInfo)]),
    "lsqb200_set_tuning": (c_int, [c_char_p]),
}

_lib = None


def lib_path() -> Path:
    env = os.environ.get("TORCHLSQ_B200_LIB")
    return Path(env) if env else Path(__file__).resolve().parent / LIB_NAME


def load():
    """dlopen the library and attach prototypes.  Raises OSError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise OSError(f"{path} not found - build it with `python __graft_entry__.py build` "
                      f"(or `make -C lsqfakequantize-pytorch_b200/csrc`)")
    lib = ctypes.CDLL(str(path))
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = _lib.lsqb200_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def qargs(quant_min, quant_max, type_min, type_max, use_grad_scaling, grad_scaler, sym, eval_mode, init_mode):
    return QArgs(int(quant_min), int(quant_max), int(type_min), int(type_max), float(grad_scaler),
                 int(bool(use_grad_scaling)), int(bool(sym)), int(bool(eval_mode)), int(bool(init_mode)))
