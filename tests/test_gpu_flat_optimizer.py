"""Fused LSQ parameter step (csrc/kern_optim.cu, torchlsq.dp.FlatLSQOptimizer) against torch.optim's own
SGD / Adam on the same parameters and gradients (third-party arithmetic: torch 2.11 single-tensor optimizers; the fused
kernel follows them operation by operation in fp32 - tolerance 2e-6 relative, the difference between ATen's and our
contraction choices)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as U
    from torchlsq import _cabi
    from torchlsq.dp import FlatLSQOptimizer


def _raw_step(p, g, s1, s2, **kw):
    a = _cabi.OptimArgs(kw.get("lr", 0.01), kw.get("weight_decay", 0.0), kw.get("grad_mul", 1.0), kw.get("momentum", 0.0),
                        kw.get("dampening", 0.0), kw.get("beta1", 0.9), kw.get("beta2", 0.999), kw.get("eps", 1e-8), kw["step"],
                        kw.get("kind", 0), int(kw.get("nesterov", False)))
    rc = _cabi.load().lsqb200_flat_optimizer_step(p.data_ptr(), g.data_ptr(), s1.data_ptr() if s1 is not None else None,
                                                  s2.data_ptr() if s2 is not None else None, p.numel(), a, U.stream())
    _cabi.check(rc, "flat step")


@pytest.mark.parametrize("cfg", [dict(), dict(momentum=0.9), dict(momentum=0.9, nesterov=True), dict(momentum=0.8, dampening=0.1, weight_decay=0.01),
                                 dict(weight_decay=0.05)])
def test_flat_sgd_matches_torch_optim(cfg):
    gen = torch.Generator().manual_seed(0)
    n = 27_702
    p0 = (torch.randn(n, generator=gen) * 0.05).to(U.DEV)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref], lr=0.02, foreach=False, **cfg)
    p = p0.clone()
    buf = torch.zeros_like(p) if cfg.get("momentum", 0) else None
    for step in range(1, 7):
        g = torch.randn(n, generator=gen).to(U.DEV)
        ref.grad = g.clone()
        opt.step()
        _raw_step(p, g, buf, None, lr=0.02, step=step, kind=0, **cfg)
        assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=1e-9), (step, float((p - ref.detach()).abs().max()))


@pytest.mark.parametrize("cfg", [dict(), dict(betas=(0.8, 0.99), eps=1e-6), dict(weight_decay=0.01)])
def test_flat_adam_matches_torch_optim(cfg):
    gen = torch.Generator().manual_seed(1)
    n = 27_702
    p0 = (torch.randn(n, generator=gen) * 0.05).to(U.DEV)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.003, foreach=False, fused=False, **cfg)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    b1, b2 = cfg.get("betas", (0.9, 0.999))
    for step in range(1, 7):
        g = torch.randn(n, generator=gen).to(U.DEV)
        ref.grad = g.clone()
        opt.step()
        _raw_step(p, g, m, v, lr=0.003, step=step, kind=1, beta1=b1, beta2=b2, eps=cfg.get("eps", 1e-8), weight_decay=cfg.get("weight_decay", 0.0))
        assert torch.allclose(p, ref.detach(), rtol=5e-6, atol=2e-9), (step, float((p - ref.detach()).abs().max()))


@pytest.mark.parametrize("kind,cfg", [("adam", dict()), ("adam", dict(betas=(0.8, 0.99), weight_decay=0.01)), ("sgd", dict(momentum=0.9, dampening=0.1)),
                                      ("sgd", dict(momentum=0.9, nesterov=True))])
def test_flat_step_sites_counts_steps_per_parameter_like_torch_optim(kind, cfg):
    """torch.optim keeps a step count per parameter and skips parameters without a gradient, so a scale that only starts
    learning after its observer window gets Adam's FIRST step (bias corrections of t = 1) and SGD's first-step momentum buffer then.
    `lsqb200_flat_optimizer_step_sites` does the same per element: B joins at step 4, sits out step 6."""
    gen = torch.Generator().manual_seed(5)
    nA, nB = 600, 411
    p0 = (torch.randn(nA + nB, generator=gen) * 0.05).to(U.DEV)
    A = p0[:nA].clone().requires_grad_(True)
    B = p0[nA:].clone().requires_grad_(True)
    if kind == "adam":
        opt = torch.optim.Adam([A, B], lr=0.003, foreach=False, fused=False, **cfg)
    else:
        opt = torch.optim.SGD([A, B], lr=0.02, foreach=False, **cfg)
    p = p0.clone()
    s1 = torch.zeros_like(p)
    s2 = torch.zeros_like(p) if kind == "adam" else None
    steps = torch.zeros(nA + nB, dtype=torch.int32, device=U.DEV)
    active = torch.ones(nA + nB, dtype=torch.uint8, device=U.DEV)
    b1, b2 = cfg.get("betas", (0.9, 0.999))
    lib = _cabi.load()
    for step in range(1, 10):
        g = torch.randn(nA + nB, generator=gen).to(U.DEV)
        b_on = step >= 4 and step != 6
        A.grad = g[:nA].clone()
        B.grad = g[nA:].clone() if b_on else None
        opt.step()
        active[nA:] = 1 if b_on else 0
        a = _cabi.OptimArgs(0.003 if kind == "adam" else 0.02, cfg.get("weight_decay", 0.0), 1.0, cfg.get("momentum", 0.0), cfg.get("dampening", 0.0),
                            b1, b2, 1e-8, 12345, 1 if kind == "adam" else 0, int(cfg.get("nesterov", False)))     # args.step is ignored
        rc = lib.lsqb200_flat_optimizer_step_sites(p.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr() if s2 is not None else None,
                                                   steps.data_ptr(), active.data_ptr(), p.numel(), a, U.stream())
        _cabi.check(rc, "flat step sites")
        ref = torch.cat([A.detach(), B.detach()])
        # updates are ~lr (3e-3 / 2e-2) per step: a wrong step count shows as 1e-3, contraction differences accumulate to ~1e-8
        assert torch.allclose(p, ref, rtol=5e-6, atol=5e-8), (step, float((p - ref).abs().max()))
    assert int(steps[0]) == 9 and int(steps[-1]) == 5
    # active == NULL: every element takes part
    rc = lib.lsqb200_flat_optimizer_step_sites(p.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr() if s2 is not None else None,
                                               steps.data_ptr(), None, p.numel(), a, U.stream())
    _cabi.check(rc, "flat step sites")
    assert int(steps[0]) == 10 and int(steps[-1]) == 6


def test_flat_optimizer_after_observer_window_matches_torch_adam():
    """The reference's default initialisation (init_mode='observer'): scale / shift do not require grad during the window
    (observers.py:455-456), torch.optim.Adam skips them and starts THEIR step count when they begin to learn.  FlatLSQOptimizer,
    built before the window as the README flow has it, must give the same parameters afterwards."""
    import torch.nn as nn
    import torch.nn.functional as F
    from torchlsq import LSQFakeQuantizer
    torch.backends.cudnn.deterministic = True

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(3, 8, 3, padding=1)
            self.fc = nn.Linear(8 * 8 * 8, 4)
            self.fq_in = LSQFakeQuantizer(torch.ao.quantization.MovingAverageMinMaxObserver, 'activation', init_mode='observer', init_batches=3)
            self.fq_w = LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable')

        def forward(self, x):
            x = self.fq_in(x)
            x = F.max_pool2d(F.relu(F.conv2d(x, self.fq_w(self.conv.weight), self.conv.bias, padding=1)), 2)
            return self.fc(x.flatten(1))

    results = []
    for use_flat in (False, True):
        torch.manual_seed(0)
        net = Net().to(U.DEV).train()
        gen = torch.Generator().manual_seed(3)
        data = [(torch.randn(16, 3, 16, 16, generator=gen).to(U.DEV), torch.randint(0, 4, (16,), generator=gen).to(U.DEV)) for _ in range(9)]
        net(data[0][0])                        # the warm-up forward that creates the LSQ parameters (reference README.md:101)
        lsq_params = [p for n, p in net.named_parameters() if n.endswith(".scale") or n.endswith(".shift")]
        other = [p for n, p in net.named_parameters() if not (n.endswith(".scale") or n.endswith(".shift"))]
        base = torch.optim.SGD(other, lr=0.01)
        flat = FlatLSQOptimizer.from_model(net, kind="adam", lr=0.002) if use_flat else torch.optim.Adam(lsq_params, lr=0.002, foreach=False)
        in_window, learning = [], []
        for x, t in data:
            loss = F.cross_entropy(net(x), t)
            learning.append(bool(net.fq_in.scale.requires_grad))
            observed = net.fq_in.scale.detach().clone()
            base.zero_grad()
            flat.zero_grad()
            loss.backward()
            base.step()
            flat.step()
            if not learning[-1]:                 # inside the window the optimizer must leave the observer's estimate alone
                assert torch.equal(net.fq_in.scale.detach(), observed)
                in_window.append(observed)
        results.append(([p.detach().clone() for p in lsq_params], in_window, learning, flat))
    (p0, w0, l0, _), (p1, w1, l1, flat) = results
    assert l0 == l1 and 0 < sum(l0) < len(l0)                       # the activation scale joined late, at the same batch in both runs
    assert len(w0) == len(w1) and all(torch.equal(a, b) for a, b in zip(w0, w1))      # the input quantizer only sees the data
    # step counts are per parameter: the late starter has as many steps as batches it was learning in, the weight scale all nine,
    # the symmetric weight shift (never learnable) none - torch.optim's state['step'] of the twin run says the same
    counts = flat.step_counts
    off_s = (flat.grads.gscale("fq_in").data_ptr() - flat.grads.flat.data_ptr()) // 4
    off_w = (flat.grads.gscale("fq_w").data_ptr() - flat.grads.flat.data_ptr()) // 4
    off_wb = (flat.grads.gshift("fq_w").data_ptr() - flat.grads.flat.data_ptr()) // 4
    assert int(counts[off_s]) == sum(l0) and int(counts[off_w]) == len(l0) and int(counts[off_wb]) == 0
    # the two trainings stay together (Adam's first steps are sign-like, so a rounding flip costs up to 2 * lr per step: loose bound)
    for a, b in zip(p0, p1):
        assert torch.isfinite(a).all() and torch.allclose(a, b, rtol=0.0, atol=0.01), float((a - b).abs().max())


def test_argument_errors():
    p = torch.zeros(8, device=U.DEV)
    lib = _cabi.load()
    a = _cabi.OptimArgs(0.1, 0.0, 1.0, 0.9, 0.0, 0.9, 0.999, 1e-8, 1, 0, 0)
    assert lib.lsqb200_flat_optimizer_step(p.data_ptr(), p.data_ptr(), None, None, 8, a, U.stream()) == -1     # momentum needs a buffer
    a.kind = 7
    assert lib.lsqb200_flat_optimizer_step(p.data_ptr(), p.data_ptr(), None, None, 8, a, U.stream()) == -1
    a.kind = 1
    assert lib.lsqb200_flat_optimizer_step(p.data_ptr(), p.data_ptr(), p.data_ptr(), None, 8, a, U.stream()) == -1   # Adam needs both states
    assert lib.lsqb200_flat_optimizer_step(None, None, None, None, 0, a, U.stream()) == 0


def _net_and_data():
    import torch.nn as nn
    import torch.nn.functional as F
    from torchlsq import LSQFakeQuantizer

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(3, 8, 3, padding=1)
            self.fc = nn.Linear(8 * 8 * 8, 4)
            self.fq_in = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=1, init_scale=0.05)
            self.fq_a = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=1, init_scale=0.05, qscheme=torch.per_channel_affine)
            self.fq_w = LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable')

        def forward(self, x):
            x = self.fq_in(x)
            x = F.relu(F.conv2d(x, self.fq_w(self.conv.weight), self.conv.bias, padding=1))
            x = F.max_pool2d(self.fq_a(x), 2)
            return self.fc(x.flatten(1))

    torch.manual_seed(0)
    net = Net().to(U.DEV).train()
    gen = torch.Generator().manual_seed(3)
    data = [(torch.randn(16, 3, 16, 16, generator=gen).to(U.DEV), torch.randint(0, 4, (16,), generator=gen).to(U.DEV)) for _ in range(6)]
    net(data[0][0])            # creates the LSQ parameters
    return net, data


@pytest.mark.parametrize("kind,kw,tkw", [("sgd", dict(lr=0.02, momentum=0.9), dict(lr=0.02, momentum=0.9)),
                                         ("adam", dict(lr=0.002), dict(lr=0.002))])
def test_module_training_with_flat_optimizer_matches_torch_optim(kind, kw, tkw):
    """Same model, same data: LSQ parameters stepped by the fused flat optimizer vs by torch.optim, all other parameters by torch.optim."""
    import torch.nn.functional as F
    torch.backends.cudnn.deterministic = True
    results = []
    for use_flat in (False, True):
        net, data = _net_and_data()
        lsq_params = [p for n, p in net.named_parameters() if n.endswith(".scale") or n.endswith(".shift")]
        other = [p for n, p in net.named_parameters() if not (n.endswith(".scale") or n.endswith(".shift"))]
        base = torch.optim.SGD(other, lr=0.01)
        if use_flat:
            flat = FlatLSQOptimizer.from_model(net, kind=kind, **kw)
            assert flat.params.numel() == 2 * (1 + 8 + 8)
            assert all(p.data_ptr() >= flat.params.data_ptr() for p in lsq_params)
        else:
            cls = torch.optim.SGD if kind == "sgd" else torch.optim.Adam
            flat = cls(lsq_params, foreach=False, **tkw)
        losses = []
        for x, t in data:
            loss = F.cross_entropy(net(x), t)
            base.zero_grad()
            flat.zero_grad()
            loss.backward()
            base.step()
            flat.step()
            losses.append(float(loss.detach()))
        results.append((losses, [p.detach().clone() for p in lsq_params]))
    (l0, p0), (l1, p1) = results
    assert torch.allclose(torch.tensor(l0), torch.tensor(l1), rtol=1e-4, atol=1e-5), (l0, l1)
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=2e-3, atol=1e-6), float((a - b).abs().max())


def test_flat_optimizer_state_dict_resumes_bit_for_bit():
    """Checkpoint / resume: model state_dict + FlatLSQOptimizer.state_dict() (moments and per-element step counts, keyed by site)
    loaded into a fresh model and optimizer continue the training with the same bits."""
    import copy
    import torch.nn.functional as F
    torch.backends.cudnn.deterministic = True

    def make():
        net, data = _net_and_data()
        other = [p for n, p in net.named_parameters() if not (n.endswith(".scale") or n.endswith(".shift"))]
        return net, data, torch.optim.SGD(other, lr=0.01, momentum=0.9), FlatLSQOptimizer.from_model(net, kind="adam", lr=0.002)

    def run(net, base, flat, batches):
        for x, t in batches:
            loss = F.cross_entropy(net(x), t)
            base.zero_grad()
            flat.zero_grad()
            loss.backward()
            base.step()
            flat.step()

    net, data, base, flat = make()
    run(net, base, flat, data[:3])
    ck = (copy.deepcopy(net.state_dict()), flat.state_dict(), copy.deepcopy(base.state_dict()))
    run(net, base, flat, data[3:])
    want = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net2, _, base2, flat2 = make()
    net2.load_state_dict(ck[0])
    flat2.load_state_dict(ck[1])
    base2.load_state_dict(ck[2])
    assert all(p.data_ptr() >= flat2.params.data_ptr() for n, p in net2.named_parameters() if n.endswith(".scale"))   # still aliasing the flat buffer
    run(net2, base2, flat2, data[3:])
    for k, v in net2.state_dict().items():
        assert torch.equal(v, want[k]), k
    assert torch.equal(flat2.step_counts, flat.step_counts)
    with pytest.raises(ValueError):
        FlatLSQOptimizer.from_model(net2, kind="sgd", lr=0.1).load_state_dict(ck[1])


def test_flat_optimizer_survives_set_to_none_zero_grad():
    """`model.zero_grad()` (set_to_none=True, torch's default) detaches the .grad views from the flat buffer; autograd
    then accumulates into fresh tensors.  step() must fold those back in - the parameters keep learning exactly as with
    `opt.zero_grad()`."""
    import torch.nn.functional as F
    torch.backends.cudnn.deterministic = True
    finals = []
    for style in ("flat_zero_grad", "model_zero_grad"):
        net, data = _net_and_data()
        lsq_params = [p for n, p in net.named_parameters() if n.endswith(".scale") or n.endswith(".shift")]
        others = [p for n, p in net.named_parameters() if not (n.endswith(".scale") or n.endswith(".shift"))]
        flat = FlatLSQOptimizer.from_model(net, kind="sgd", lr=0.02, momentum=0.9)
        before = [p.detach().clone() for p in lsq_params]
        for x, t in data:
            if style == "flat_zero_grad":
                flat.zero_grad()
                for p in others:
                    p.grad = None
            else:
                net.zero_grad()                              # every .grad -> None, LSQ parameters included
                flat.zero_grad()                             # the flat buffer itself still has to start from zero
            F.cross_entropy(net(x), t).backward()
            flat.step()
            assert all(p.grad is not None and p.grad.data_ptr() >= flat.grads.flat.data_ptr() for p in lsq_params if p.requires_grad)
        finals.append([p.detach().clone() for p in lsq_params])
        assert any(not torch.equal(a, b) for a, b in zip(before, finals[-1]))       # they did move
    for a, b in zip(*finals):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9), float((a - b).abs().max())
