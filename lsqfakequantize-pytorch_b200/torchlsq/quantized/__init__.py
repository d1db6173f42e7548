"""Module-tree toggles, same names as /root/reference/torchlsq/quantized/__init__.py:5-35.
Use with `model.apply(torchlsq.disable_observer)` etc."""
import torch
from .modules.observers import LSQFakeQuantizer

_TORCH_FQ = torch.quantization.FakeQuantize


def _is_fq(mod):
    return isinstance(mod, (_TORCH_FQ, LSQFakeQuantizer))


def _is_lsq_of(mod, qdtype):
    # torch's own FakeQuantize always matches; an LSQ quantizer only with the given dtype
    return isinstance(mod, _TORCH_FQ) or (isinstance(mod, LSQFakeQuantizer) and mod.dtype == qdtype)


def disable_fake_quant(mod):
    if _is_fq(mod):
        mod.disable_fake_quant()


def enable_fake_quant(mod):
    if _is_fq(mod):
        mod.enable_fake_quant()


def disable_observer(mod):
    if _is_fq(mod):
        mod.disable_observer()


def enable_observer(mod):
    if _is_fq(mod):
        mod.enable_observer()


def disable_fake_quant_on_act(mod):
    if _is_lsq_of(mod, torch.quint8):
        mod.disable_fake_quant()


def enable_fake_quant_on_act(mod):
    if _is_lsq_of(mod, torch.quint8):
        mod.enable_fake_quant()


def disable_observer_on_weights(mod):
    if _is_lsq_of(mod, torch.qint8):
        mod.disable_observer()


def enable_observer_on_weights(mod):
    if _is_lsq_of(mod, torch.qint8):
        mod.enable_observer()


__all__ = ["LSQFakeQuantizer", "disable_fake_quant", "enable_fake_quant", "disable_observer", "enable_observer",
           "disable_fake_quant_on_act", "enable_fake_quant_on_act", "disable_observer_on_weights",
           "enable_observer_on_weights"]
