"""Prologue fusion for models prepared with PyTorch's eager-mode quantization flow.

The reference is used through `torch.quantization` (README.md:101-126 of the reference: a `QConfig` of
`LSQFakeQuantizer.with_args(...)`, `fuse_modules_qat`, `prepare_qat`).  In a model prepared that way an activation
quantizer sits
  * behind the ReLU of a fused QAT module (`torch.ao.nn.intrinsic.qat.ConvBnReLU2d`, `ConvReLU2d`, `LinearReLU`, ...):
    `forward` = `F.relu(conv_bn(x))`, then the module's forward hook calls `activation_post_process`;
  * behind a residual join (`torch.ao.nn.quantized.FloatFunctional.add_relu` / `.add`): `torch.add`, `F.relu`, then
    `activation_post_process`.
`fuse_prologues(model)` re-routes those two patterns so that the ReLU / add + ReLU run INSIDE the fake-quant kernels
(`LSQFakeQuantizer(fuse_relu=True)`, `LSQFakeQuantizer.forward_add`; `torchlsq.functional.lsq_relu / lsq_add_relu /
lsq_add`): one pass over the activation instead of two or three, forward and backward, bit-identical results.

Only instance attributes are touched (`module.forward`, `functional.add_relu`, `functional.add`, `quantizer.fuse_relu`):
module TYPES stay what `torch.ao.quantization.convert` expects, so the prepared model converts as before -
`unfuse_prologues(model)` restores the instances exactly (call it before `convert` if you rely on type-keyed hooks that
inspect `forward`; `convert` itself replaces the patched modules and needs nothing).
"""
import types
from typing import Dict

import torch

from .quantized.modules.observers import LSQFakeQuantizer

_MARK = "_lsqb200_fused_prologue"


def _relu_parent_forward(mod):
    """`forward` of the first base class that is the same module without its ReLU, for torch.ao's fused QAT modules
    (ConvBnReLU{1,2,3}d -> ConvBn{1,2,3}d, ConvReLU{1,2,3}d -> qat.Conv{1,2,3}d, LinearReLU -> qat.Linear); else None."""
    try:
        import torch.ao.nn.intrinsic.qat as nniqat
    except ImportError:      # pragma: no cover
        return None
    relu_types = tuple(getattr(nniqat, n) for n in ("ConvBnReLU1d", "ConvBnReLU2d", "ConvBnReLU3d", "ConvReLU1d", "ConvReLU2d",
                                                    "ConvReLU3d", "LinearReLU") if hasattr(nniqat, n))
    if type(mod) not in relu_types:          # exact types only: a user subclass may do more than add a ReLU
        return None
    return type(mod).__mro__[1].forward


def _has_observer_hook(mod):
    return any(getattr(h, "__name__", "") == "_observer_forward_hook" for h in mod._forward_hooks.values())


def fuse_prologues(model: torch.nn.Module, relu: bool = True, residual: bool = True) -> Dict[str, int]:
    """Fold the ReLU of fused QAT modules and the add [+ ReLU] of `FloatFunctional`s into the `LSQFakeQuantizer` behind
    them.  Returns how many sites of each kind were re-routed.  Idempotent."""
    from torch.ao.nn.quantized import FloatFunctional
    done = {"relu": 0, "residual": 0}
    for mod in model.modules():
        if getattr(mod, _MARK, None):
            continue
        app = getattr(mod, "activation_post_process", None)
        if not isinstance(app, LSQFakeQuantizer):
            continue
        if relu and _has_observer_hook(mod):
            parent_forward = _relu_parent_forward(mod)
            if parent_forward is not None and not app.fuse_relu:
                mod.forward = types.MethodType(parent_forward, mod)      # the module without its F.relu
                app.fuse_relu = True                                     # ... which now runs inside the quantizer's kernels
                object.__setattr__(mod, _MARK, "relu")
                done["relu"] += 1
                continue
        if residual and isinstance(mod, FloatFunctional):
            def add_relu(self, x, y):
                return self.activation_post_process.forward_add(x, y, relu=True)

            def add(self, x, y):
                return self.activation_post_process.forward_add(x, y, relu=False)
            mod.add_relu = types.MethodType(add_relu, mod)
            mod.add = types.MethodType(add, mod)
            object.__setattr__(mod, _MARK, "residual")
            done["residual"] += 1
    return done


def unfuse_prologues(model: torch.nn.Module) -> int:
    """Undo `fuse_prologues`: the instances get their class's own `forward` / `add_relu` / `add` back."""
    n = 0
    for mod in model.modules():
        kind = getattr(mod, _MARK, None)
        if not kind:
            continue
        if kind == "relu_fx":                    # graph-mode patch: the graph still passes relu=True, only a re-prepare undoes it
            continue
        if kind == "relu":
            mod.__dict__.pop("forward", None)
            mod.activation_post_process.fuse_relu = False
        else:
            mod.__dict__.pop("add_relu", None)
            mod.__dict__.pop("add", None)
        mod.__dict__.pop(_MARK, None)
        n += 1
    return n


# ---- FX graph mode (torch.ao.quantization.quantize_fx.prepare_qat_fx) --------------------------------------------------------
def _is_relu_node(node, modules):
    import torch.nn.functional as F
    if node.op == "call_function":
        return node.target in (torch.relu, F.relu, torch.relu_) and len(node.args) >= 1
    if node.op == "call_method":
        return node.target in ("relu", "relu_")
    if node.op == "call_module":
        return type(modules.get(node.target)) is torch.nn.ReLU
    return False


def _is_add_node(node):
    import operator
    inplace = False
    if node.op == "call_function" and node.target in (operator.add, operator.iadd, torch.add):
        ok = len(node.args) == 2 and not node.kwargs          # torch.add(a, b, alpha=...) is left alone
        inplace = node.target is operator.iadd
    elif node.op == "call_method" and node.target in ("add", "add_"):
        ok = len(node.args) == 2 and not node.kwargs
        inplace = node.target == "add_"
    else:
        return False
    if not (ok and all(isinstance(a, torch.fx.Node) for a in node.args)):
        return False
    # `a += b` / `a.add_(b)` mutate a: if anything else reads a, dropping the add would hand that reader the pre-add value
    return not inplace or len(node.args[0].users) == 1


def fuse_prologues_fx(gm: "torch.fx.GraphModule", relu: bool = True, residual: bool = True) -> Dict[str, int]:
    """`fuse_prologues` for a GraphModule returned by `prepare_qat_fx`.  There the quantizers are graph nodes
    (`activation_post_process_N`), so the two patterns are rewritten in the graph itself:

        fused ConvBnReLU2d / ConvReLU2d / LinearReLU -> quantizer   the module loses its F.relu (instance patch), quantizer(x, relu=True)
        [add ->] [relu ->] quantizer                                 quantizer(a, b, relu=True) / quantizer(x, relu=True) / quantizer(a, b, relu=False)

    Only single-user chains are touched (an in-place ReLU or add whose result something else reads stays as it is).  The
    rewritten graph trains with bit-identical results; `convert_fx` pattern-matches the ORIGINAL graph, so convert a freshly
    prepared copy loaded with the trained `state_dict()` (module names and parameters are unchanged by this pass)."""
    modules = dict(gm.named_modules(remove_duplicate=False))   # prepare_qat_fx shares one quantizer between several nodes (pool, relu ...)
    done = {"relu": 0, "residual": 0}
    graph = gm.graph
    for node in list(graph.nodes):
        if node.op != "call_module" or not isinstance(modules.get(node.target), LSQFakeQuantizer):
            continue
        if len(node.args) != 1 or node.kwargs or not isinstance(node.args[0], torch.fx.Node):
            continue
        app = modules[node.target]
        p = node.args[0]
        if len(p.users) != 1:
            continue
        if relu and p.op == "call_module":
            mod = modules.get(p.target)
            parent_forward = _relu_parent_forward(mod) if mod is not None else None
            if parent_forward is not None and not getattr(mod, _MARK, None):
                mod.forward = types.MethodType(parent_forward, mod)      # the module without its F.relu ...
                node.kwargs = {"relu": True}                             # ... which this CALL of the (possibly shared) quantizer applies
                object.__setattr__(mod, _MARK, "relu_fx")
                done["relu"] += 1
                continue
        has_relu = _is_relu_node(p, modules)
        src = p.args[0] if has_relu and p.args and isinstance(p.args[0], torch.fx.Node) else None
        if has_relu and (src is None or len(src.users) != 1):
            continue                                       # an in-place ReLU on a tensor something else reads: leave it
        q = src if has_relu else p
        if residual and _is_add_node(q) and len(q.users) == 1:
            a, b = q.args
            node.args = (a, b)
            node.kwargs = {"relu": bool(has_relu)}
            if has_relu:
                graph.erase_node(p)
            graph.erase_node(q)
            done["residual"] += 1
        elif relu and has_relu:
            node.args = (src,)
            node.kwargs = {"relu": True}
            graph.erase_node(p)
            done["relu"] += 1
    graph.lint()
    gm.recompile()
    return done
