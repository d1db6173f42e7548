"""Host-side logic that needs no GPU: functional.lsq argument handling, operator registration,
error behaviour on CPU tensors, LSQFakeQuantizer ranges / toggles / state machine against the
traces recorded from the reference module (tests/golden/ref_module_trace.json)."""
import pytest
import torch

import torchlsq
from torchlsq import LSQFakeQuantizer
from torchlsq.functional import lsq
from torchlsq.quantized.modules import observers as obs


def test_ops_registered_with_reference_schemas():
    assert torchlsq.extension._HAS_OPS and torchlsq.extension._has_ops()
    tail = ("int quant_min, int quant_max, int type_min, int type_max, bool use_grad_scaling, float grad_scaler, "
            "bool sym, bool eval_mode, bool init_mode")
    want = {
        "lsq_forward_per_tensor": f"torchlsq::lsq_forward_per_tensor(Tensor x, Tensor scale, Tensor shift, {tail}) -> Tensor",
        "lsq_backward_per_tensor": f"torchlsq::lsq_backward_per_tensor(Tensor grad, Tensor x, Tensor scale, Tensor shift, {tail}) -> (Tensor, Tensor, Tensor)",
        "lsq_forward_per_channel": f"torchlsq::lsq_forward_per_channel(Tensor x, Tensor scale, Tensor shift, int axis, {tail}) -> Tensor",
        "lsq_backward_per_channel": f"torchlsq::lsq_backward_per_channel(Tensor grad, Tensor x, Tensor scale, Tensor shift, int axis, {tail}) -> (Tensor, Tensor, Tensor)",
    }
    for name, schema in want.items():
        assert str(getattr(torch.ops.torchlsq, name).default._schema) == schema
    s = str(torch.ops.torchlsq.lsq.default._schema)
    assert s.count("Tensor") == 4 and s.count("int") == 5 and s.count("bool") == 5 and "float" in s
    assert torch.ops.torchlsq._cuda_version() // 1000 == 12
    assert torchlsq.extension._check_cuda_version() == torch.ops.torchlsq._cuda_version()


def test_no_cpu_fallback():
    x = torch.randn(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        lsq(x, torch.ones(1), torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.torchlsq.lsq_backward_per_tensor(x, x, torch.ones(1), torch.zeros(1), 0, 127, 0, 255, True, 1.0,
                                                   False, False, False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        lsq(x, torch.ones(3), torch.zeros(3), is_perchannel=True, axis=1)


def test_functional_argument_checks():
    x = torch.randn(4, 3)
    with pytest.raises(AssertionError, match="symmetric"):
        lsq(x, torch.ones(1), torch.zeros(1), quant_min=1, quant_max=5, is_affine=False)
    with pytest.raises(RuntimeError, match="1-D tensor"):
        lsq(x, torch.tensor(1.0), torch.zeros(1))
    with pytest.raises(RuntimeError, match="1-D tensor"):
        lsq(x, torch.ones(1), torch.zeros(1, 1))


def test_qranges_match_reference(golden_module):
    kw = {
        "act_default": dict(otype='activation'),
        "act_8bit": dict(otype='activation', avoid_torch_overflow=False),
        "act_sym": dict(otype='activation', qscheme=torch.per_tensor_symmetric),
        "act_custom": dict(otype='activation', quant_min=0, quant_max=15),
        "w_default": dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric),
        "w_8bit": dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric, avoid_torch_overflow=False),
        "w_custom": dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric, quant_min=-7, quant_max=8,
                         init_scale=0.5),
    }
    for tag, ref in golden_module["qranges"].items():
        m = LSQFakeQuantizer(None, init_mode='learnable', **kw[tag])
        got = dict(quant_min=m.quant_min, quant_max=m.quant_max, init_shift=m.init_shift, ch_axis=m.ch_axis,
                   n_batches=m.n_batches)
        assert got == ref, tag


def test_constructor_asserts():
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_tensor_affine, init_mode='learnable')
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(None, 'weight', dtype=torch.quint8, qscheme=torch.per_tensor_symmetric, init_mode='learnable')
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(None, 'activation', dtype=torch.qint8, init_mode='learnable')
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(None, 'activation', init_mode='nope')
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(None, 'activation', init_mode='learnable', quant_min=0, quant_max=255)   # > 7 bit
    with pytest.raises(AssertionError):
        LSQFakeQuantizer(torch.quantization.MinMaxObserver(), 'activation', init_mode='observer')   # instance, not class


def test_with_args_factory_works():
    """the reference raises NameError here (functools.partial never imported)."""
    f = LSQFakeQuantizer.with_args(observer=torch.quantization.MovingAverageMinMaxObserver, otype='activation',
                                   init_batches=5)
    a, b = f(), f()
    assert a is not b and a.n_batches == 5 and isinstance(a.activation_post_process,
                                                          torch.quantization.MovingAverageMinMaxObserver)
    g = f.with_args(grad_scaler=2.0)
    assert g().grad_scaler == 2.0


def test_toggles_and_state_dict_keys():
    m = LSQFakeQuantizer(torch.quantization.MovingAverageMinMaxObserver, 'activation')
    assert set(m.state_dict().keys()) >= {"fake_quant_enabled", "observer_enabled", "learning_enabled", "current_batch"}
    assert int(m.observer_enabled[0]) == 1
    m.enable_param_learning()
    assert int(m.learning_enabled[0]) == 1 and int(m.observer_enabled[0]) == 0 and m.n_batches == -1
    m.enable_static_estimate()
    assert int(m.learning_enabled[0]) == 0 and int(m.observer_enabled[0]) == 1
    m.disable_fake_quant()
    assert int(m.fake_quant_enabled[0]) == 0
    torchlsq.enable_fake_quant(m)
    assert int(m.fake_quant_enabled[0]) == 1
    torchlsq.disable_observer(m)
    assert int(m.observer_enabled[0]) == 0
    torchlsq.enable_observer_on_weights(m)       # quint8 module: not a weight quantizer -> untouched
    assert int(m.observer_enabled[0]) == 0
    torchlsq.disable_fake_quant_on_act(m)
    assert int(m.fake_quant_enabled[0]) == 0
    # mirrors follow load_state_dict
    sd = m.state_dict()
    sd["current_batch"] = torch.tensor([77])
    m.load_state_dict(sd)
    assert m._m_batch == 77 and m._m_fq == 0


def test_calculate_qparams_and_zp_conversion():
    m = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_scale=0.5, init_shift=-3.2)
    s, zp = m.calculate_qparams(verbose=False)
    assert s == 0.5 and zp == 6          # round(3.2 / 0.5) = round(6.4)
    zp = LSQFakeQuantizer.convert_shift_to_zp(torch.tensor([-1000.0, 0.3, 1.25]), torch.tensor([0.5, 0.1, 0.5]), torch.quint8)
    assert zp.tolist() == [255, 0, 0] and zp.dtype == torch.int64
    zp = LSQFakeQuantizer.convert_shift_to_zp(torch.tensor([1000.0, -1.25]), torch.tensor([0.5, 0.5]), torch.qint8)
    assert zp.tolist() == [-128, 2]      # round-half-even of 2.5


class _Spy:
    def __init__(self):
        self.calls = []

    def __call__(self, x, scale, shift, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, is_affine, is_perchannel,
                 eval_mode=False, init_mode=False):
        self.calls.append(dict(qmin=qmin, qmax=qmax, tmin=tmin, tmax=tmax, axis=axis, use_gs=use_gs, gscaler=gscaler,
                               is_affine=is_affine, is_perchannel=is_perchannel, eval_mode=bool(eval_mode),
                               init_mode=bool(init_mode), scale_rg=bool(scale.requires_grad),
                               shift_rg=bool(shift.requires_grad)))
        return x * 1.0


def _run_trace(monkeypatch, build, steps, train_flags=None, x_shape=(4, 6), weight=False):
    spy = _Spy()
    monkeypatch.setattr(obs, "lsq", spy)
    if weight:   # the mu+-3sigma kernel needs a GPU; the state machine does not care about the values
        monkeypatch.setattr(obs, "weight_init_scale",
                            lambda w, ax, pc, qmin, qmax: torch.full((w.shape[ax] if pc else 1,), 0.01))
    torch.manual_seed(7)
    m = build()
    rec = []
    for i in range(steps):
        m.train(train_flags[i] if train_flags else True)
        spy.calls.clear()
        x = torch.randn(*x_shape) + 0.5
        out = m(x)
        rec.append(dict(step=i, identity=bool(out is x), call=(dict(spy.calls[0]) if spy.calls else None),
                        observer_enabled=int(m.observer_enabled[0]), current_batch=int(m.current_batch[0]),
                        n_batches=int(m.n_batches)))
    return rec


def _compare(rec, ref):
    assert len(rec) == len(ref)
    for a, b in zip(rec, ref):
        for k in ("step", "identity", "observer_enabled", "current_batch", "n_batches"):
            assert a[k] == b[k], (a["step"], k, a[k], b[k])
        assert (a["call"] is None) == (b["call"] is None), a["step"]
        if a["call"]:
            for k, v in a["call"].items():
                assert v == b["call"][k], (a["step"], k, v, b["call"][k])


MA = torch.quantization.MovingAverageMinMaxObserver


@pytest.mark.parametrize("name,build,steps,flags,shape,weight", [
    ("act_learnable_n3", lambda: LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=3), 8, None, (4, 6), False),
    ("act_observer_n2", lambda: LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2), 7, None, (4, 6), False),
    ("act_observer_static", lambda: LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2, learn_params=False), 5, None, (4, 6), False),
    ("act_learnable_evalmix", lambda: LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=2), 7,
     [True, True, False, True, True, False, True], (4, 6), False),
    ("weight_sym", lambda: LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable'), 4, None, (6, 4, 3, 3), True),
    ("weight_sym_8bit_tensor", lambda: LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric, init_mode='learnable', avoid_torch_overflow=False), 3, None, (6, 4, 3, 3), True),
])
def test_state_machine_matches_reference_trace(monkeypatch, golden_module, name, build, steps, flags, shape, weight):
    rec = _run_trace(monkeypatch, build, steps, flags, shape, weight)
    _compare(rec, golden_module["traces"][name])


def test_observer_mode_copies_observer_qparams(monkeypatch, golden_module):
    """scale / shift written by `_set_weights` from the torch observer, step by step, vs the reference."""
    spy = _Spy()
    seen = []

    def spy2(x, scale, shift, *a, **k):
        seen.append(([float(v) for v in scale.detach().reshape(-1)[:4]], [float(v) for v in shift.detach().reshape(-1)[:4]]))
        return spy(x, scale, shift, *a, **k)

    monkeypatch.setattr(obs, "lsq", spy2)
    torch.manual_seed(7)
    m = LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2)
    for i in range(7):
        m.train(True)
        m(torch.randn(4, 6) + 0.5)
    ref = [r["call"] for r in golden_module["traces"]["act_observer_n2"] if r["call"]]
    assert len(seen) == len(ref)
    for (s, b), r in zip(seen, ref):
        assert s == pytest.approx(r["scale"], rel=1e-6) and b == pytest.approx(r["shift"], rel=1e-6, abs=1e-7)


def test_meta_tensors_flow_through_the_ops():
    """Shape functions only (dispatch key Meta): same shapes / strides / dtypes as the CUDA kernels produce."""
    import torch
    import torchlsq  # noqa: F401
    x = torch.empty(4, 6, 5, 5, device="meta").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    s = torch.empty(1, device="meta", requires_grad=True)
    b = torch.empty(1, device="meta", requires_grad=True)
    y = torch.ops.torchlsq.lsq(x, s, b, 0, 127, 0, 255, 1, True, 1.0, True, False, False, False)
    assert y.is_meta and y.shape == x.shape and y.stride() == x.stride()
    y.backward(torch.empty_like(y))
    assert x.grad.shape == x.shape and s.grad.shape == (1,) and b.grad.shape == (1,)
    sc = torch.empty(6, device="meta", requires_grad=True)
    bc = torch.empty(6, device="meta", requires_grad=True)
    yc = torch.ops.torchlsq.lsq(x.detach(), sc, bc, 0, 127, 0, 255, 1, True, 1.0, True, True, False, False)
    yc.backward(torch.empty_like(yc))
    assert sc.grad.shape == (6,) and bc.grad.shape == (6,)
    import pytest
    with pytest.raises(RuntimeError, match="not consistent"):
        torch.ops.torchlsq.lsq(x.detach(), torch.empty(5, device="meta"), torch.empty(5, device="meta"), 0, 127, 0, 255, 1, True, 1.0,
                               True, True, False, False)


# ---- prologue fusion, host side: everything that needs no kernel ---------------------------------------------
def test_fused_prologue_functions_have_no_cpu_path_and_check_arguments():
    from torchlsq.functional import lsq_add, lsq_add_relu, lsq_relu
    x, s, b = torch.randn(4, 8), torch.tensor([0.1]), torch.tensor([0.0])
    for fn, args in ((lsq_relu, (x,)), (lsq_add_relu, (x, x)), (lsq_add, (x, x))):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn(*args, s, b, 0, 127)
        with pytest.raises(RuntimeError, match="1-D"):
            fn(*args, torch.tensor(0.1), b, 0, 127)
    with pytest.raises(AssertionError):
        lsq_relu(x, s, b, 1, 127, is_affine=False)


def test_module_fuse_relu_paths_that_return_the_input():
    """debug mode, the parameter-creating first call and a disabled fake-quant hand back relu(x) (and relu(a + b) for
    forward_add) - exactly what Sequential(ReLU(), LSQFakeQuantizer()) would."""
    x, y = torch.randn(4, 6), torch.randn(4, 6)
    m = LSQFakeQuantizer(None, 'activation', init_mode='learnable', fuse_relu=True, debug_mode=True)
    assert torch.equal(m(x), torch.relu(x)) and torch.equal(m.forward_add(x, y), torch.relu(x + y))
    assert torch.equal(m.forward_add(x, y, relu=False), x + y)
    m = LSQFakeQuantizer(None, 'activation', init_mode='learnable', fuse_relu=True)
    assert torch.equal(m(x), torch.relu(x)) and m._initialized          # first call: creates scale / shift, returns relu(x)
    m.disable_fake_quant()
    m.disable_observer()
    assert torch.equal(m(x), torch.relu(x)) and torch.equal(m.forward_add(x, y), torch.relu(x + y))
    plain = LSQFakeQuantizer(None, 'activation', init_mode='learnable')
    assert plain.fuse_relu is False and torch.equal(plain(x), x)


def test_fuse_prologues_patches_instances_only():
    import copy
    import torch.ao.quantization as tq
    import torch.nn as nn
    from torch.ao.nn.quantized import FloatFunctional
    from torchlsq.fusion import fuse_prologues, unfuse_prologues

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1, self.bn1, self.relu1 = nn.Conv2d(4, 4, 3, padding=1, bias=False), nn.BatchNorm2d(4), nn.ReLU()
            self.conv2, self.bn2 = nn.Conv2d(4, 4, 3, padding=1, bias=False), nn.BatchNorm2d(4)
            self.fc, self.fc_relu = nn.Linear(4, 4), nn.ReLU()
            self.skip, self.other = FloatFunctional(), FloatFunctional()
            self.quant, self.dequant = tq.QuantStub(), tq.DeQuantStub()

        def forward(self, x):
            x = self.quant(x)
            y = self.bn2(self.conv2(self.relu1(self.bn1(self.conv1(x)))))
            return self.dequant(self.skip.add_relu(y, x))

    m = Block().train()
    m.qconfig = tq.QConfig(activation=LSQFakeQuantizer.with_args(observer=tq.MovingAverageMinMaxObserver, otype='activation'),
                           weight=LSQFakeQuantizer.with_args(observer=None, otype='weight', dtype=torch.qint8,
                                                             qscheme=torch.per_channel_symmetric, init_mode='learnable'))
    m.other.qconfig = tq.QConfig(activation=tq.FakeQuantize.with_args(observer=tq.MovingAverageMinMaxObserver),
                                 weight=tq.default_weight_fake_quant)           # not ours: must be left alone
    tq.fuse_modules_qat(m, [['conv1', 'bn1', 'relu1'], ['conv2', 'bn2'], ['fc', 'fc_relu']], inplace=True)
    tq.prepare_qat(m, inplace=True)
    types_before = [type(x) for x in m.modules()]
    assert fuse_prologues(m) == {"relu": 2, "residual": 1}                      # ConvBnReLU2d, LinearReLU; skip
    assert fuse_prologues(m) == {"relu": 0, "residual": 0}                      # idempotent
    assert [type(x) for x in m.modules()] == types_before                       # convert() keys on module types
    assert m.conv1.activation_post_process.fuse_relu and m.fc.activation_post_process.fuse_relu
    assert not m.conv2.activation_post_process.fuse_relu                        # ConvBn2d has no ReLU to fold
    assert "add_relu" in m.skip.__dict__ and "add_relu" not in m.other.__dict__
    # the patched forward is the parent's (no F.relu): negative outputs survive until the quantizer
    m.conv1.activation_post_process.debug_mode = True
    m2 = copy.deepcopy(m)                                                       # patches follow a deep copy and bind to the copy
    assert m2.conv1.forward.__self__ is m2.conv1 and m2.skip.add_relu.__self__ is m2.skip
    assert unfuse_prologues(m) == 3
    assert "forward" not in m.conv1.__dict__ and "add_relu" not in m.skip.__dict__ and not m.conv1.activation_post_process.fuse_relu
    assert fuse_prologues(m, relu=False) == {"relu": 0, "residual": 1}


def test_fuse_prologues_fx_rewrites_single_user_chains_only():
    import torch.ao.quantization as tq
    import torch.nn as nn
    from torch.ao.quantization import QConfigMapping
    from torch.ao.quantization.quantize_fx import prepare_qat_fx
    from torchlsq.fusion import fuse_prologues_fx

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1, self.bn1, self.relu1 = nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU()
            self.conv2, self.bn2 = nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8)
            self.pool, self.relu2 = nn.AdaptiveAvgPool2d(1), nn.ReLU(inplace=True)

        def forward(self, x):
            y = self.relu1(self.bn1(self.conv1(x)))
            y = self.bn2(self.conv2(y))
            z = torch.relu(y + x)                       # add -> relu -> quantizer: one call
            w = self.relu2(self.pool(z))                # relu behind a quantizer torch shares between pool and relu: per-call flag
            t = y + w                                   # `y` has a second user (the add above): this add is still single-user, fused
            return z + t

    qc = tq.QConfig(activation=LSQFakeQuantizer.with_args(observer=tq.MovingAverageMinMaxObserver, otype='activation'),
                    weight=LSQFakeQuantizer.with_args(observer=None, otype='weight', dtype=torch.qint8,
                                                      qscheme=torch.per_channel_symmetric, init_mode='learnable'))
    gm = prepare_qat_fx(Block().train(), QConfigMapping().set_global(qc), example_inputs=(torch.randn(1, 8, 8, 8),))
    before = {n.name for n in gm.graph.nodes}
    done = fuse_prologues_fx(gm)
    after = {n.name: n for n in gm.graph.nodes}
    assert done["relu"] == 2 and done["residual"] == 3
    assert not any("relu" in n or n.startswith("add") for n in after)          # every relu / add node is gone ...
    assert {n for n in before if "relu" in n or n.startswith("add")}           # ... and there were some
    calls = [n for n in gm.graph.nodes if n.op == "call_module" and n.kwargs]
    assert sorted((len(n.args), n.kwargs["relu"]) for n in calls) == [(1, True), (1, True), (2, False), (2, False), (2, True)]
    modules = dict(gm.named_modules(remove_duplicate=False))
    assert all(isinstance(modules[n.target], LSQFakeQuantizer) for n in calls)
    assert not any(m.fuse_relu for m in modules.values() if isinstance(m, LSQFakeQuantizer))   # graph mode keeps the flag per call
    assert "forward" in gm.conv1.__dict__                                       # ConvBnReLU2d lost its F.relu
    assert fuse_prologues_fx(gm) == {"relu": 0, "residual": 0}


def test_round2_binding_entries_on_cpu():
    """torchlsq/_C.so (csrc/torch_binding.cpp) also registers the fused-prologue op and the grouped op; without a GPU they must be
    present with their schemas, refuse CPU tensors loudly (no CPU path) and keep their bookkeeping calls harmless."""
    import torchlsq  # noqa: F401
    from torchlsq.multi import LSQGroup
    assert "Tensor[] xs, Tensor[] scales, Tensor[] shifts" in str(torch.ops.torchlsq.lsq_group.default._schema)
    assert str(torch.ops.torchlsq.lsq_pre.default._schema).startswith("torchlsq::lsq_pre(int prologue, Tensor x, Tensor? x2, Tensor scale, Tensor shift")
    assert torch.ops.torchlsq._b200_abi_version() == 3
    w = [torch.randn(4, 3), torch.randn(5, 3)]
    sc = [torch.ones(4), torch.ones(5)]
    sh = [torch.zeros(4), torch.zeros(5)]
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.torchlsq.lsq_group(w, sc, sh, -128, 127, -128, 127, 0, True, 1.0, False, True, False, False, 12345)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.torchlsq.lsq_pre(1, w[0], None, sc[0], sh[0], 0, 127, 0, 255, 1, True, 1.0, True, False, False, False)
    assert list(torch.ops.torchlsq.lsq_group_info(12345)) == [0, 0]
    torch.ops.torchlsq.lsq_group_release(12345)
    with pytest.raises(ValueError):
        LSQGroup([], [], [])
    with pytest.raises(RuntimeError, match="dtype"):
        LSQGroup([w[0], w[1].half()], sc, sh)
    with pytest.raises(AssertionError, match="symmetric"):
        LSQGroup(w, sc, sh, quant_min=1, quant_max=5, is_affine=False)
    g = LSQGroup(w, sc, sh)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g()
    g.close()


def test_state_epoch_moves_with_the_state_machine():
    """group_weight_quantizers re-derives its grouping only when this counter moved: every flag change and parameter creation must bump it."""
    from torchlsq import LSQFakeQuantizer
    from torchlsq.quantized.modules import observers as obs
    q = LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable')
    seen = [obs.STATE_EPOCH[0]]
    for action in (q.disable_fake_quant, q.enable_fake_quant, q.enable_static_estimate, q.enable_param_learning, q.sync_state,
                   lambda: q.load_state_dict(q.state_dict()), q.reset):
        action()
        assert obs.STATE_EPOCH[0] > seen[-1], action
        seen.append(obs.STATE_EPOCH[0])
