"""`torchlsq.fusion.fuse_prologues` on a model prepared with PyTorch's eager-mode QAT flow - the way the reference is used
(its README: QConfig of LSQFakeQuantizer.with_args, fuse_modules_qat, prepare_qat): two residual blocks with
ConvBnReLU2d / ConvBn2d / FloatFunctional.add_relu.  The fused model (ReLU and add + ReLU inside the fake-quant kernels)
and an untouched deep copy are trained side by side through the observer window, the hand-over and steady-state steps:
same logits and loss bit for bit, same learned parameters; `unfuse_prologues` restores the instances."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(init_mode):
    import torch.ao.quantization as tq
    from torch.ao.nn.quantized import FloatFunctional
    from torchlsq import LSQFakeQuantizer

    class Block(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv1, self.bn1, self.relu1 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU()
            self.conv2, self.bn2 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c)
            self.skip = FloatFunctional()

        def forward(self, x):
            y = self.relu1(self.bn1(self.conv1(x)))
            y = self.bn2(self.conv2(y))
            return self.skip.add_relu(y, x)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.quant, self.dequant = tq.QuantStub(), tq.DeQuantStub()
            self.stem, self.stem_bn, self.stem_relu = nn.Conv2d(3, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16), nn.ReLU()
            self.b1, self.b2 = Block(16), Block(16)
            self.join = FloatFunctional()
            self.fc = nn.Linear(16, 10)

        def forward(self, x):
            x = self.quant(x)
            x = self.stem_relu(self.stem_bn(self.stem(x)))
            a = self.b1(x)
            b = self.b2(a)
            x = self.join.add(a, b)                       # residual join without activation
            x = self.dequant(x)
            return self.fc(x.mean((2, 3)))

    torch.manual_seed(0)
    net = Net().train()
    MA = tq.MovingAverageMinMaxObserver
    act = LSQFakeQuantizer.with_args(observer=MA if init_mode == "observer" else None, otype="activation", init_mode=init_mode,
                                     init_batches=2, init_scale=0.05)
    wei = LSQFakeQuantizer.with_args(observer=None, otype="weight", dtype=torch.qint8, qscheme=torch.per_channel_symmetric,
                                     init_mode="learnable", avoid_torch_overflow=False)
    net.qconfig = tq.QConfig(activation=act, weight=wei)
    net.fc.qconfig = None
    fuse = [["stem", "stem_bn", "stem_relu"]]
    for b in ("b1", "b2"):
        fuse += [[f"{b}.conv1", f"{b}.bn1", f"{b}.relu1"], [f"{b}.conv2", f"{b}.bn2"]]
    tq.fuse_modules_qat(net, fuse, inplace=True)
    tq.prepare_qat(net, inplace=True)
    return net.to(DEV)


@pytest.mark.parametrize("init_mode", ["observer", "learnable"])
def test_fused_prologues_train_like_the_unfused_model(init_mode):
    from torchlsq import LSQFakeQuantizer
    from torchlsq.fusion import fuse_prologues, unfuse_prologues
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    plain = _build(init_mode)
    fused = copy.deepcopy(plain)
    types_before = [type(m) for m in fused.modules()]
    assert fuse_prologues(fused) == {"relu": 3, "residual": 3}
    assert fuse_prologues(fused) == {"relu": 0, "residual": 0}
    assert [type(m) for m in fused.modules()] == types_before            # convert() keys on types: untouched
    gen = torch.Generator().manual_seed(1)
    data = [(torch.randn(16, 3, 16, 16, generator=gen).to(DEV), torch.randint(0, 10, (16,), generator=gen).to(DEV)) for _ in range(8)]
    opts = None
    for step, (x, t) in enumerate(data):
        outs = []
        for i, net in enumerate((plain, fused)):
            logits = net(x)                                              # step 0 only creates the LSQ parameters (module contract)
            loss = F.cross_entropy(logits, t)
            outs.append((logits.detach().clone(), loss.detach().clone()))
            if opts is not None:
                opts[i].zero_grad()
                loss.backward()
                opts[i].step()
        if opts is None:
            opts = [torch.optim.SGD(n.parameters(), lr=0.02, momentum=0.9) for n in (plain, fused)]
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (init_mode, step)
        for (n0, p0), (n1, p1) in zip(plain.named_parameters(), fused.named_parameters()):
            assert n0 == n1 and torch.allclose(p0, p1, rtol=1e-5, atol=1e-7), (init_mode, step, n0)
    # steady state reached: the fused quantizers really took the fused path (observer off, parameters learnable)
    q = fused.b1.conv1.activation_post_process
    assert isinstance(q, LSQFakeQuantizer) and q.fuse_relu and int(q.observer_enabled[0]) == 0
    plain.eval(), fused.eval()
    with torch.no_grad():
        assert torch.allclose(plain(data[0][0]), fused(data[0][0]), rtol=1e-5, atol=1e-6)
    assert unfuse_prologues(fused) == 6
    assert "forward" not in fused.b1.conv1.__dict__ and "add_relu" not in fused.b1.skip.__dict__ and not q.fuse_relu
    with torch.no_grad():
        assert torch.allclose(plain(data[1][0]), fused(data[1][0]), rtol=1e-5, atol=1e-6)


def test_fx_graph_mode_fusion_trains_like_the_unfused_graph():
    """`fuse_prologues_fx` on a `prepare_qat_fx` GraphModule (the reference's quantizer works in FX graph mode as well): ConvBnReLU2d,
    add -> relu and a bare add in front of quantizers, two residual blocks; fused and unfused graphs train side by side with
    bit-identical logits and loss."""
    import torch.ao.quantization as tq
    from torch.ao.quantization import QConfigMapping
    from torch.ao.quantization.quantize_fx import prepare_qat_fx
    from torchlsq import LSQFakeQuantizer
    from torchlsq.fusion import fuse_prologues_fx
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False

    class Block(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv1, self.bn1, self.relu1 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU()
            self.conv2, self.bn2 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c)

        def forward(self, x):
            y = self.relu1(self.bn1(self.conv1(x)))
            y = self.bn2(self.conv2(y))
            return torch.relu(y + x)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem, self.stem_bn, self.stem_relu = nn.Conv2d(3, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16), nn.ReLU()
            self.b1, self.b2 = Block(16), Block(16)
            self.fc = nn.Linear(16, 10)

        def forward(self, x):
            x = self.stem_relu(self.stem_bn(self.stem(x)))
            a = self.b1(x)
            b = self.b2(a)
            return self.fc((a + b).mean((2, 3)))

    def build():
        torch.manual_seed(0)
        act = LSQFakeQuantizer.with_args(observer=tq.MovingAverageMinMaxObserver, otype="activation", init_mode="observer", init_batches=2)
        wei = LSQFakeQuantizer.with_args(observer=None, otype="weight", dtype=torch.qint8, qscheme=torch.per_channel_symmetric,
                                         init_mode="learnable", avoid_torch_overflow=False)
        gm = prepare_qat_fx(Net().train(), QConfigMapping().set_global(tq.QConfig(activation=act, weight=wei)),
                            example_inputs=(torch.randn(1, 3, 16, 16),))
        return gm.to(DEV)

    plain, fused = build(), build()
    done = fuse_prologues_fx(fused)
    assert done["relu"] >= 3 and done["residual"] >= 3, done
    gen = torch.Generator().manual_seed(1)
    data = [(torch.randn(16, 3, 16, 16, generator=gen).to(DEV), torch.randint(0, 10, (16,), generator=gen).to(DEV)) for _ in range(7)]
    opts = None
    for step, (x, t) in enumerate(data):
        outs = []
        for i, net in enumerate((plain, fused)):
            logits = net(x)
            loss = F.cross_entropy(logits, t)
            outs.append((logits.detach().clone(), loss.detach().clone()))
            if opts is not None:
                opts[i].zero_grad()
                loss.backward()
                opts[i].step()
        if opts is None:
            opts = [torch.optim.SGD(n.parameters(), lr=0.02, momentum=0.9) for n in (plain, fused)]
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), step
    for (n0, p0), (n1, p1) in zip(plain.named_parameters(), fused.named_parameters()):
        assert n0 == n1 and torch.allclose(p0, p1, rtol=1e-5, atol=1e-7), n0
