"""Grouped fake-quant (`torchlsq.multi.LSQGroup`, `group_weight_quantizers`): all weight quantizers of a model behind ONE
autograd node and one multi-tensor launch per direction, with the plan's per-step tensors re-pointed by
`lsqb200_plan_rebind`.  Contract: bit-identical to one `torchlsq.functional.lsq` call per site (which the other GPU tests pin
to the oracle), forward and every gradient, step after step with fresh buffers."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [(64, 3, 7, 7), (64, 64, 1, 1), (64, 64, 3, 3), (256, 64, 1, 1), (128, 128, 3, 3), (1000, 512), (17, 5, 3, 3), (8, 33)]


def _sites(dtype=torch.float32, seed=0):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    ws = [(torch.randn(s, device=DEV, generator=gen) * 0.05).to(dtype).requires_grad_(True) for s in SHAPES]
    sc = [(0.001 + 0.002 * torch.rand(s[0], device=DEV, generator=gen)).requires_grad_(True) for s in SHAPES]
    sh = [torch.zeros(s[0], device=DEV).requires_grad_(True) for s in SHAPES]
    return ws, sc, sh


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("learn", [True, False])
def test_group_matches_per_site_calls_bit_for_bit(dtype, learn):
    from torchlsq.functional import lsq
    from torchlsq.multi import LSQGroup
    ws, sc, sh = _sites(dtype)
    ws2, sc2, sh2 = ([t.detach().clone().requires_grad_(True) for t in ts] for ts in (ws, sc, sh))
    group = LSQGroup(ws, sc, sh, -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True, eval_mode=not learn)
    gen = torch.Generator(device=DEV).manual_seed(7)
    for step in range(3):                     # fresh outputs / upstream grads every step: the plan is re-pointed, not rebuilt
        gs = [torch.randn(s, device=DEV, generator=gen).to(dtype) for s in SHAPES]
        ys = group()
        torch.autograd.backward(ys, gs)
        ys2 = [lsq(w, s, b, -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True, eval_mode=not learn)
               for w, s, b in zip(ws2, sc2, sh2)]
        torch.autograd.backward(ys2, gs)
        for i in range(len(SHAPES)):
            assert torch.equal(ys[i], ys2[i]), (step, i)
            assert torch.equal(ws[i].grad, ws2[i].grad), (step, i)
            assert torch.equal(sc[i].grad, sc2[i].grad), (step, i)
            assert torch.equal(sh[i].grad, sh2[i].grad), (step, i)
            assert ws[i].grad.data_ptr() % 32 == 0
        for t in ws + sc + sh + ws2 + sc2 + sh2:
            t.grad = None
    assert group.launches() <= 6              # a handful of kernel classes, not 2 x len(SHAPES)
    assert group.info()[0] == 1               # one plan for the three steps: fresh outputs / grads were re-pointed, not rebuilt
    # an in-place update of the weights (optimizer step) keeps the plan; moving a weight's storage rebuilds it
    with torch.no_grad():
        ws[0].mul_(0.5); ws2[0].mul_(0.5)
    assert torch.equal(group()[0], lsq(ws2[0], sc2[0], sh2[0], -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True,
                                       eval_mode=not learn))
    assert group.info()[0] == 1
    ws[1].data = ws[1].data.clone()
    group()
    assert group.info()[0] == 2


def test_group_errors_and_no_grad():
    from torchlsq.multi import LSQGroup
    ws, sc, sh = _sites()
    with pytest.raises(ValueError):
        LSQGroup([], [], [])
    with pytest.raises(RuntimeError, match="contiguous"):
        LSQGroup([ws[0].permute(1, 0, 2, 3)], sc[:1], sh[:1])
    with pytest.raises(RuntimeError, match="dtype"):
        LSQGroup([ws[0], ws[1].half()], sc[:2], sh[:2])
    group = LSQGroup(ws, sc, sh)
    with torch.no_grad():
        ys = group()
    assert all(not y.requires_grad for y in ys)
    ys = group()
    with pytest.raises(RuntimeError, match="double backwards"):
        torch.autograd.grad([y.sum() for y in ys], ws[0], create_graph=True)


def _qat_net():
    import torch.ao.quantization as tq
    from torchlsq import LSQFakeQuantizer

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.quant, self.dequant = tq.QuantStub(), tq.DeQuantStub()
            self.c1, self.r1 = nn.Conv2d(3, 16, 3, padding=1), nn.ReLU()
            self.c2, self.r2 = nn.Conv2d(16, 32, 3, padding=1, stride=2), nn.ReLU()
            self.c3, self.r3 = nn.Conv2d(32, 32, 1), nn.ReLU()
            self.fc = nn.Linear(32, 10)

        def forward(self, x):
            x = self.quant(x)
            x = self.r1(self.c1(x))
            x = self.r2(self.c2(x))
            x = self.r3(self.c3(x))
            x = self.fc(x.mean((2, 3)))
            return self.dequant(x)

    torch.manual_seed(0)
    net = Net().train()
    act = LSQFakeQuantizer.with_args(observer=None, otype="activation", init_mode="learnable", init_batches=1, init_scale=0.05)
    wei = LSQFakeQuantizer.with_args(observer=None, otype="weight", dtype=torch.qint8, qscheme=torch.per_channel_symmetric,
                                     init_mode="learnable", avoid_torch_overflow=False)
    net.qconfig = tq.QConfig(activation=act, weight=wei)
    tq.prepare_qat(net, inplace=True)
    return net.to(DEV)


def test_grouped_weight_quantizers_train_like_the_plain_model():
    """prepare_qat model (QAT Conv2d / Linear call `self.weight_fake_quant(self.weight)`): with the model-level helper the four
    weight quantizers run as one launch per direction; logits, loss and every learned parameter stay bit-identical to the
    untouched deep copy over the warm-up call, the learned-init window and steady-state SGD steps."""
    import warnings
    from torchlsq import LSQFakeQuantizer
    from torchlsq import multi
    from torchlsq.multi import group_weight_quantizers
    calls = []
    orig_call = multi.LSQGroup.__call__

    def counting_call(self):
        calls.append(self.n)
        return orig_call(self)
    multi.LSQGroup.__call__ = counting_call
    try:
        _grouped_training_body(LSQFakeQuantizer, group_weight_quantizers, calls)
    finally:
        multi.LSQGroup.__call__ = orig_call


def _grouped_training_body(LSQFakeQuantizer, group_weight_quantizers, calls):
    import warnings
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plain = _qat_net()
    grouped = copy.deepcopy(plain)
    handle = group_weight_quantizers(grouped)
    gen = torch.Generator(device=DEV).manual_seed(3)
    x0 = torch.randn(8, 3, 16, 16, device=DEV, generator=gen)
    with torch.no_grad():                                   # the reference's rule: one forward creates the parameters
        assert torch.equal(plain(x0), grouped(x0))
    opts = [torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9) for m in (plain, grouped)]
    wq = [m for m in grouped.modules() if isinstance(m, LSQFakeQuantizer) and m.otype == 0]
    assert len(wq) == 4
    for step in range(5):
        x = torch.randn(8, 3, 16, 16, device=DEV, generator=gen)
        t = torch.randint(0, 10, (8,), device=DEV, generator=gen)
        losses = []
        for m, opt in zip((plain, grouped), opts):
            opt.zero_grad(set_to_none=True)
            out = m(x)
            loss = nn.functional.cross_entropy(out, t)
            loss.backward()
            opt.step()
            losses.append((out.detach(), loss.detach()))
        assert torch.equal(losses[0][0], losses[1][0]) and torch.equal(losses[0][1], losses[1][1]), step
    assert calls == [4] * 5, calls                          # every training step quantised the four weights in one group call
    for (n1, p1), (n2, p2) in zip(plain.named_parameters(), grouped.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    # eval: still grouped (parameters frozen by no_grad), same logits
    plain.eval(); grouped.eval()
    with torch.no_grad():
        assert torch.equal(plain(x0), grouped(x0))
    # a quantizer leaving steady state drops out of the group without changing results
    wq[0].disable_fake_quant()
    [m for m in plain.modules() if isinstance(m, LSQFakeQuantizer) and m.otype == 0][0].disable_fake_quant()
    with torch.no_grad():
        assert torch.equal(plain(x0), grouped(x0))
    assert calls[-1] == 3
    handle.remove()
    with torch.no_grad():
        assert torch.equal(plain(x0), grouped(x0))


def test_group_per_tensor_sites_and_awkward_upstream_grads():
    """Per-tensor sites (one scale / shift element each) in a group, and upstream gradients autograd may hand over in any form:
    a misaligned view, an expanded scalar (y.sum().backward()), a missing one (an output nobody used).  The group copies what the
    kernels cannot read in place and must still match per-site calls bit for bit."""
    from torchlsq.functional import lsq
    from torchlsq.multi import LSQGroup
    gen = torch.Generator(device=DEV).manual_seed(9)
    shapes = [(4, 16, 14, 14), (1000,), (3, 5, 7), (8, 64)]
    for dtype in (torch.float32, torch.bfloat16):
        xs = [torch.randn(s, device=DEV, generator=gen).to(dtype).requires_grad_(True) for s in shapes]
        sc = [torch.tensor([0.03 + 0.01 * i], device=DEV, requires_grad=True) for i in range(len(shapes))]
        sh = [torch.tensor([-0.5 * i], device=DEV, requires_grad=True) for i in range(len(shapes))]
        xs2, sc2, sh2 = ([t.detach().clone().requires_grad_(True) for t in ts] for ts in (xs, sc, sh))
        group = LSQGroup(xs, sc, sh, 0, 127, 0, 255, is_affine=True, is_perchannel=False)
        ys = group()
        ys2 = [lsq(x, s, b, 0, 127, 0, 255) for x, s, b in zip(xs2, sc2, sh2)]
        big = torch.randn(shapes[0][0] * shapes[0][1] * 196 + 1, device=DEV, generator=gen).to(dtype)
        g0 = big[1:].view(shapes[0])                                   # not 32-byte aligned
        g1 = torch.ones((), device=DEV, dtype=dtype).expand(shapes[1])   # what sum().backward() produces
        g2 = torch.randn(shapes[2], device=DEV, generator=gen).to(dtype)
        # output 3 gets no gradient at all
        torch.autograd.backward([ys[0], ys[1], ys[2]], [g0, g1, g2])
        torch.autograd.backward([ys2[0], ys2[1], ys2[2]], [g0, g1, g2])
        for i in range(3):
            assert torch.equal(ys[i], ys2[i])
            assert torch.equal(xs[i].grad, xs2[i].grad), (dtype, i)
            # a plan may cut a per-tensor site into other tiles than a single launch does: same terms, another fixed fp64 order
            assert torch.allclose(sc[i].grad, sc2[i].grad, rtol=1e-6, atol=0) and torch.allclose(sh[i].grad, sh2[i].grad, rtol=1e-6, atol=0), \
                (dtype, i, sc[i].grad, sc2[i].grad, sh[i].grad, sh2[i].grad)
        assert torch.equal(ys[3], ys2[3])
        assert float(xs[3].grad.abs().sum()) == 0.0 and float(sc[3].grad.abs().sum()) == 0.0       # zeros for the unused output
        group.close()
