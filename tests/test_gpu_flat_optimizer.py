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
