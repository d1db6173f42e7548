"""Module-tree toggles, same names and meaning as /root/reference/torchlsq/quantized/__init__.py:5-35.
Use with `model.apply(torchlsq.disable_observer)` etc.

Each toggle calls one method of a fake-quant module.  torch's own `FakeQuantize` always matches; an `LSQFakeQuantizer`
matches the `_on_act` / `_on_weights` variants only when it is an activation (quint8) / weight (qint8) quantizer."""
import torch

from .modules.observers import LSQFakeQuantizer


def _toggle(name, method, lsq_dtype=None):
    def apply(mod):
        ours = isinstance(mod, LSQFakeQuantizer) and (lsq_dtype is None or mod.dtype == lsq_dtype)
        if ours or isinstance(mod, torch.quantization.FakeQuantize):
            getattr(mod, method)()
    apply.__name__ = apply.__qualname__ = name
    apply.__doc__ = f"`model.apply({name})`: call `{method}()` on the matching fake-quant modules."
    return apply


_TABLE = {
    "disable_fake_quant": ("disable_fake_quant", None),
    "enable_fake_quant": ("enable_fake_quant", None),
    "disable_observer": ("disable_observer", None),
    "enable_observer": ("enable_observer", None),
    "disable_fake_quant_on_act": ("disable_fake_quant", torch.quint8),
    "enable_fake_quant_on_act": ("enable_fake_quant", torch.quint8),
    "disable_observer_on_weights": ("disable_observer", torch.qint8),
    "enable_observer_on_weights": ("enable_observer", torch.qint8),
}
for _name, (_method, _dtype) in _TABLE.items():
    globals()[_name] = _toggle(_name, _method, _dtype)

__all__ = ["LSQFakeQuantizer", *_TABLE]
