"""GPU tests of the reference-facing surface: torch.ops.torchlsq.*, torchlsq.functional.lsq with
autograd, LSQFakeQuantizer, multi-tensor plans - all running on the sm_100a kernels, checked
against the CPU oracle (and, where the reference CUDA build is present, against the reference
itself in a subprocess)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import gpu_util as U
from conftest import GOLDEN, ROOT, geometry
from oracle import lsq_oracle as O

pytestmark = pytest.mark.gpu


def _oracle(x, g, scale, shift, per_channel=False, axis=1, **kw):
    xb, dt = O.to_bits(x)
    gb, _ = O.to_bits(g)
    if per_channel:
        outer, C, inner = geometry(tuple(x.shape), axis)
    else:
        outer, C, inner = 1, 1, x.numel()
    c = O.cfg(**kw)
    y = O.forward(xb.reshape(-1), scale.detach().float().cpu().numpy(), shift.detach().float().cpu().numpy(), c, outer, C, inner,
                  per_channel, dt=dt)
    gx, gs, gb_, ms, mb = O.backward(gb.reshape(-1), xb.reshape(-1), scale.detach().float().cpu().numpy(),
                                     shift.detach().float().cpu().numpy(), c, outer, C, inner, per_channel, dt=dt, with_abs=True)
    return y, gx, gs, gb_, ms, mb


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_functional_autograd_per_tensor(dtype):
    from torchlsq.functional import lsq
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(8, 16, 14, 14, generator=gen).to(dtype).to(U.DEV).requires_grad_(True)
    g = torch.randn(8, 16, 14, 14, generator=gen).to(dtype).to(U.DEV)
    s = torch.tensor([0.03], device=U.DEV, requires_grad=True)
    b = torch.tensor([-1.7], device=U.DEV, requires_grad=True)
    y = lsq(x, s, b, 0, 127, 0, 255)
    assert y.shape == x.shape and y.dtype == dtype and y.grad_fn is not None
    y.backward(g)
    oy, ogx, ogs, ogb, ms, mb = _oracle(x, g, s, b, quant_min=0, quant_max=127, type_min=0, type_max=255)
    assert U.same_bits(y, oy) and U.same_bits(x.grad, ogx)
    assert s.grad.shape == (1,) and s.grad.dtype == torch.float32
    U.assert_grads_close(s.grad, ogs, ms, 1e-5)
    U.assert_grads_close(b.grad, ogb, mb, 1e-5)


def test_functional_defaults_and_type_range():
    """quant range defaults [0, 255] and type_min/max default to the quant range (functional.py:92-93)."""
    from torchlsq.functional import lsq
    x = torch.linspace(-3, 9, 1001, device=U.DEV)
    s = torch.tensor([0.02], device=U.DEV)
    b = torch.tensor([-1.0], device=U.DEV)
    y = lsq(x, s, b)
    oy = O.forward(x.cpu().numpy(), [0.02], [-1.0], O.cfg(0, 255, 0, 255))
    assert U.same_bits(y, oy)
    y2 = lsq(x, s, b, 0, 15)                       # zero point 50 clamps to type_max = 15
    oy2 = O.forward(x.cpu().numpy(), [0.02], [-1.0], O.cfg(0, 15, 0, 15))
    assert U.same_bits(y2, oy2)


@pytest.mark.parametrize("axis,shape", [(1, (4, 12, 9, 9)), (0, (16, 8, 3, 3)), (2, (3, 5, 7)), (1, (32, 40))])
def test_functional_autograd_per_channel(axis, shape):
    from torchlsq.functional import lsq
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=gen).to(U.DEV).requires_grad_(True)
    g = torch.randn(*shape, generator=gen).to(U.DEV)
    C = shape[axis]
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV).requires_grad_(True)
    b = (-torch.rand(C, generator=gen)).to(U.DEV).requires_grad_(True)
    y = lsq(x, s, b, 0, 127, 0, 255, axis=axis, is_perchannel=True)
    y.backward(g)
    oy, ogx, ogs, ogb, ms, mb = _oracle(x, g, s, b, per_channel=True, axis=axis, quant_min=0, quant_max=127, type_min=0, type_max=255)
    assert U.same_bits(y, oy) and U.same_bits(x.grad, ogx)
    U.assert_grads_close(s.grad, ogs, ms, 1e-6)
    U.assert_grads_close(b.grad, ogb, mb, 1e-6)


def test_per_channel_broadcasts_single_scale():
    """lsq.cpp:124-126: a length-1 scale/shift is repeated to the channel count."""
    from torchlsq.functional import lsq
    x = torch.randn(2, 6, 5, device=U.DEV)
    s1 = torch.tensor([0.05], device=U.DEV)
    b6 = torch.linspace(-1, 0, 6, device=U.DEV)
    y = lsq(x, s1, b6, 0, 127, 0, 255, axis=1, is_perchannel=True)
    y_ref = lsq(x, s1.repeat(6), b6, 0, 127, 0, 255, axis=1, is_perchannel=True)
    assert torch.equal(y, y_ref)


def test_memory_formats_and_views():
    """channels_last / permuted / sliced inputs: same values as the contiguous tensor; dense layouts keep
    their strides (empty_like(Preserve), lsq_cuda.cu:38)."""
    from torchlsq.functional import lsq
    gen = torch.Generator().manual_seed(2)
    xc = torch.randn(4, 8, 6, 6, generator=gen).to(U.DEV)
    s = torch.tensor([0.04], device=U.DEV)
    b = torch.tensor([-0.5], device=U.DEV)
    y0 = lsq(xc, s, b, 0, 127, 0, 255)
    xl = xc.contiguous(memory_format=torch.channels_last)
    yl = lsq(xl, s, b, 0, 127, 0, 255)
    assert torch.equal(yl, y0) and yl.stride() == xl.stride()
    xp = xc.permute(0, 2, 3, 1)
    assert torch.equal(lsq(xp, s, b, 0, 127, 0, 255), y0.permute(0, 2, 3, 1))
    xs = xc[:, ::2]
    assert torch.equal(lsq(xs, s, b, 0, 127, 0, 255), y0[:, ::2])
    # per-channel on channels_last memory (channel is the fastest dimension there)
    sc = (0.02 + 0.01 * torch.arange(8)).to(U.DEV)
    bc = (-0.1 * torch.arange(8)).float().to(U.DEV)
    yc = lsq(xc, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True)
    ycl = lsq(xl, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True)
    assert torch.equal(ycl, yc)
    # backward with an expanded (stride-0) upstream gradient and a channels_last input
    xg = xl.clone().requires_grad_(True)
    sg = sc.clone().requires_grad_(True)
    bg = bc.clone().requires_grad_(True)
    lsq(xg, sg, bg, 0, 127, 0, 255, axis=1, is_perchannel=True).sum().backward()
    xr = xc.clone().requires_grad_(True)
    sr = sc.clone().requires_grad_(True)
    br = bc.clone().requires_grad_(True)
    lsq(xr, sr, br, 0, 127, 0, 255, axis=1, is_perchannel=True).backward(torch.ones_like(xr))
    assert torch.equal(xg.grad, xr.grad)
    assert torch.allclose(sg.grad, sr.grad, rtol=1e-6, atol=1e-7) and torch.allclose(bg.grad, br.grad, rtol=1e-6, atol=1e-7)


def test_requires_grad_subsets_eval_and_symmetric():
    from torchlsq.functional import lsq
    x = torch.randn(1000, device=U.DEV)
    s = torch.tensor([0.03], device=U.DEV, requires_grad=True)
    b = torch.tensor([0.0], device=U.DEV, requires_grad=True)
    y = lsq(x, s, b, -64, 63, -128, 127, is_affine=False)          # x needs no grad
    y.sum().backward()
    assert s.grad is not None and b.grad.item() == 0.0
    s.grad = b.grad = None
    xe = x.clone().requires_grad_(True)
    lsq(xe, s, b, 0, 127, 0, 255, eval_mode=True).sum().backward()
    assert s.grad.item() == 0.0 and b.grad.item() == 0.0 and xe.grad is not None
    xi = x.clone().requires_grad_(True)
    yi = lsq(xi, s, b, 0, 127, 0, 255, init_mode=True)
    assert torch.equal(yi, xi)
    yi.backward(torch.full_like(xi, 3.0))
    assert torch.equal(xi.grad, torch.full_like(xi, 3.0))


def test_errors_match_reference_behaviour():
    from torchlsq.functional import lsq
    x = torch.randn(4, 6, device=U.DEV)
    s = torch.ones(1, device=U.DEV)
    b = torch.zeros(1, device=U.DEV)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        lsq(x, s.cpu(), b)
    with pytest.raises(RuntimeError, match="same floating-point type"):
        lsq(x, s.half(), b.half())
    with pytest.raises(RuntimeError, match="float32, float16 or bfloat16"):
        lsq(x.to(torch.int32), s, b)
    assert lsq(x.double(), s.double(), b.double()).dtype == torch.float64      # float64 is served (tests/test_gpu_f64.py)
    with pytest.raises(RuntimeError, match="not consistent"):
        lsq(x, torch.ones(5, device=U.DEV), torch.zeros(5, device=U.DEV), is_perchannel=True, axis=1)
    with pytest.raises(RuntimeError, match="axis"):
        lsq(x, torch.ones(6, device=U.DEV), torch.zeros(6, device=U.DEV), is_perchannel=True, axis=2)
    with pytest.raises(RuntimeError, match="same dimensions"):
        torch.ops.torchlsq.lsq_forward_per_channel(x, torch.ones(6, device=U.DEV), torch.zeros(5, device=U.DEV), 1, 0, 127, 0, 255,
                                                   True, 1.0, False, False, False)
    with pytest.raises(RuntimeError, match="same size"):
        torch.ops.torchlsq.lsq_backward_per_tensor(x[:2], x, s, b, 0, 127, 0, 255, True, 1.0, False, False, False)
    # double backward is refused (lsq_autograd.cpp:106)
    xg = x.clone().requires_grad_(True)
    sg = s.clone().requires_grad_(True)
    y = lsq(xg, sg, b, 0, 127, 0, 255)
    (gx,) = torch.autograd.grad(y.sum(), xg, create_graph=True)
    with pytest.raises(RuntimeError, match="double backwards"):
        gx.sum().backward()


def test_backward_ops_are_callable_directly():
    x = torch.randn(3000, device=U.DEV)
    g = torch.randn(3000, device=U.DEV)
    s = torch.tensor([0.03], device=U.DEV)
    b = torch.tensor([-1.0], device=U.DEV)
    gx, gs, gb = torch.ops.torchlsq.lsq_backward_per_tensor(g, x, s, b, 0, 127, 0, 255, True, 1.0, False, False, False)
    ogx, ogs, ogb = O.backward(g.cpu().numpy(), x.cpu().numpy(), [0.03], [-1.0], O.cfg(0, 127, 0, 255))
    assert U.same_bits(gx, ogx) and gs.shape == (1,) and abs(gs.item() - ogs[0]) <= 1e-6 * abs(ogs[0]) + 1e-9


# ------------------------------------------------------------------------------------------------
# module
# ------------------------------------------------------------------------------------------------
def test_module_weight_quantizer_init_and_training_step(golden_module):
    from torchlsq import LSQFakeQuantizer
    tag = "conv_64x3x7x7"
    ref = golden_module["winit"][tag]
    w = torch.from_numpy(np.load(GOLDEN / f"ref_winit_{tag}.npz")[f"winit/{tag}"]).to(U.DEV)
    m = LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable').to(U.DEV)
    out = m(w)
    assert out is w                                     # the first call only initialises
    assert m.scale.shape == (64,) and m.scale.dtype == torch.float32 and m.scale.is_cuda
    assert np.allclose(m.scale.detach().cpu().numpy(), np.array(ref["scale"]), rtol=2e-6)   # reference module's mu+-3sigma values
    assert torch.count_nonzero(m.shift).item() == 0
    wq = w.clone().requires_grad_(True)
    y = m(wq)
    y.backward(torch.ones_like(y))
    oy = O.forward(w.cpu().numpy().reshape(-1), m.scale.detach().cpu().numpy(), m.shift.detach().cpu().numpy(),
                   O.cfg(-64, 63, -128, 127, sym=True), 1, 64, 147, True)
    assert U.same_bits(y, oy)
    assert m.scale.grad is not None and m.shift.grad is None and not m.shift.requires_grad


def test_module_activation_learnable_init_then_lsq():
    from torchlsq import LSQFakeQuantizer
    torch.manual_seed(0)
    m = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=3, init_scale=0.5).to(U.DEV)
    m.train()
    xs = [torch.randn(8, 32, 10, 10, device=U.DEV).relu() for _ in range(8)]
    assert m(xs[0]) is xs[0]
    opt = torch.optim.SGD([m.scale, m.shift], lr=1e-2)
    losses = []
    for i in range(1, 8):
        x = xs[i].clone().requires_grad_(True)
        y = m(x)
        if i <= 3:                                      # learned-init window: identity forward, grads from ||x_r - x||^2
            assert torch.equal(y, x)
        else:
            assert not torch.equal(y, x)
        opt.zero_grad()
        y.sum().backward()
        assert torch.isfinite(m.scale.grad).all() and torch.isfinite(m.shift.grad).all()
        losses.append(float(m.scale.grad))
        opt.step()
    assert int(m.current_batch[0]) == 4 and m._m_batch == 4


def test_module_observer_mode_on_gpu():
    from torchlsq import LSQFakeQuantizer
    m = LSQFakeQuantizer(torch.quantization.MovingAverageMinMaxObserver, 'activation', init_mode='observer', init_batches=2).to(U.DEV)
    m.train()
    x = torch.rand(16, 64, device=U.DEV) * 4
    m(x)
    for i in range(4):
        xi = (torch.rand(16, 64, device=U.DEV) * 4).requires_grad_(True)
        y = m(xi)
        y.sum().backward()
    assert 0.02 < float(m.scale) < 0.05                 # ~ 4 / 127
    assert m.scale.requires_grad and int(m.observer_enabled[0]) == 0
    sd = m.state_dict()
    m2 = LSQFakeQuantizer(torch.quantization.MovingAverageMinMaxObserver, 'activation', init_mode='observer', init_batches=2).to(U.DEV)
    m2(x)                                               # creates scale / shift so the checkpoint loads
    m2.load_state_dict(sd)
    assert torch.equal(m2.scale, m.scale) and m2._m_batch == m._m_batch
    m.eval(); m2.eval()
    assert torch.equal(m(x), m2(x))


# ------------------------------------------------------------------------------------------------
# multi-tensor plans
# ------------------------------------------------------------------------------------------------
def test_plan_matches_individual_calls_bitwise():
    from torchlsq.multi import LSQPlan, Site
    gen = torch.Generator().manual_seed(5)
    shapes = [(64, 3, 7, 7), (64, 64, 1, 1), (128, 128, 3, 3), (512, 2048, 1, 1), (1000, 2048), (256, 256, 3, 3)]
    sites, refs = [], []
    for shp in shapes:
        w = (torch.randn(*shp, generator=gen) * 0.05).to(U.DEV)
        g = torch.randn(*shp, generator=gen).to(U.DEV)
        s = (0.0005 + 0.001 * torch.rand(shp[0], generator=gen)).to(U.DEV)
        b = torch.zeros(shp[0], device=U.DEV)
        st = Site(x=w, y=torch.empty_like(w), grad=g, gx=torch.empty_like(w), scale=s, shift=b,
                  gscale=torch.empty(shp[0], device=U.DEV), gshift=torch.empty(shp[0], device=U.DEV),
                  quant_min=-128, quant_max=127, type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True)
        sites.append(st)
    # plus two per-tensor bf16 activation sites in the same plan
    for n in (100_003, 3_000_000):
        x = torch.randn(n, generator=gen).to(torch.bfloat16).to(U.DEV)
        g = torch.randn(n, generator=gen).to(torch.bfloat16).to(U.DEV)
        sites.append(Site(x=x, y=torch.empty_like(x), grad=g, gx=torch.empty_like(x), scale=torch.tensor([0.03], device=U.DEV),
                          shift=torch.tensor([-1.0], device=U.DEV), gscale=torch.empty(1, device=U.DEV), gshift=torch.empty(1, device=U.DEV)))
    plan = LSQPlan(sites)
    plan.forward()
    plan.backward()
    assert plan.launches(False) >= 2 and plan.launches(True) >= 2
    for st in sites:
        q = U.qa(st.quant_min, st.quant_max, st.quant_min if st.type_min is None else st.type_min,
                 st.quant_max if st.type_max is None else st.type_max, True, 1.0, not st.is_affine, False, False)
        if st.is_perchannel:
            outer, C, inner = 1, st.x.shape[0], st.x[0].numel()
        else:
            outer, C, inner = 1, 1, st.x.numel()
        y = U.fwd(st.x.reshape(-1), st.scale, st.shift, q, outer, C, inner, st.is_perchannel)
        gx, gs, gb = U.bwd(st.grad.reshape(-1), st.x.reshape(-1), st.scale, st.shift, q, outer, C, inner, st.is_perchannel)
        assert torch.equal(y, st.y.reshape(-1)) and torch.equal(gx, st.gx.reshape(-1))
        assert torch.equal(gs, st.gscale) and torch.equal(gb, st.gshift)
    # one-launch mu +- 3 sigma for the weight sites
    wplan = LSQPlan(sites[:6])
    out = wplan.weight_init_stats()
    off = 0
    for st in sites[:6]:
        C = st.x.shape[0]
        ref = O.weight_init(st.x.cpu().numpy().reshape(-1), -128, 127, 1, C, st.x[0].numel())
        assert np.allclose(out[off:off + C].cpu().numpy(), ref, rtol=2e-6)
        off += C


def test_flat_grad_buffer_receives_kernel_output_directly():
    """grad_scale / grad_shift are written straight into slices of the data-parallel flat buffer."""
    from torchlsq.dp import FlatGradBuffer
    flat = FlatGradBuffer([("a", 1), ("w", 8)], U.DEV)
    x, g = torch.randn(5000, device=U.DEV), torch.randn(5000, device=U.DEV)
    s = torch.tensor([0.03], device=U.DEV)
    b = torch.tensor([-1.0], device=U.DEV)
    from torchlsq import _cabi
    lib = _cabi.load()
    gs, gb = flat.views("a")
    ws = U.workspace()
    assert lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), None, s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(), 5000, 0, 0,
                                  U.qa(), ws.data_ptr(), ws.numel(), U.stream()) == 0
    _, rs, rb = U.bwd(g, x, s, b, U.qa())
    assert flat.flat[0].item() == rs.item() and flat.flat[1].item() == rb.item()
    assert torch.count_nonzero(flat.flat[2:]).item() == 0


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA op (oracle/_ref, built for sm_100a), run in a subprocess because both
# packages register the `torchlsq::` dispatcher namespace
# ------------------------------------------------------------------------------------------------
_REF_SCRIPT = r'''
import sys, json, torch
sys.path.insert(0, sys.argv[1])
import torchlsq
from torchlsq.functional import lsq
assert "oracle/_ref" in torchlsq.__file__
d = torch.load(sys.argv[2])
out = {}
for name, c in d.items():
    x = c["x"].cuda().requires_grad_(True); g = c["g"].cuda()
    s = c["s"].cuda().requires_grad_(True); b = c["b"].cuda().requires_grad_(True)
    y = lsq(x, s, b, *c["args"])
    y.backward(g)
    out[name] = dict(y=y.detach().cpu(), gx=x.grad.cpu(), gs=s.grad.cpu(), gb=(b.grad if b.grad is not None else torch.zeros_like(b)).cpu())
torch.save(out, sys.argv[3])
'''


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "torchlsq" / "_C.so").exists(), reason="reference CUDA build (oracle/_ref) not present")
def test_bit_exact_against_reference_cuda_op(tmp_path):
    from torchlsq.functional import lsq
    gen = torch.Generator().manual_seed(77)
    cases = {}

    def add(name, x, g, s, b, args):
        cases[name] = dict(x=x, g=g, s=s, b=b, args=args)

    x1 = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(1))
    g1 = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(2))
    add("config1_fp32", x1, g1, torch.tensor([0.03]), torch.tensor([-1.7]), (0, 127, 0, 255, 1, True, 1.0, True, False, False, False))
    add("fp32_sym", x1[:4], g1[:4], torch.tensor([0.02]), torch.tensor([0.0]), (-64, 63, -128, 127, 1, True, 1.0, False, False, False, False))
    add("fp32_init", x1[:4], g1[:4], torch.tensor([0.03]), torch.tensor([-1.7]), (0, 127, 0, 255, 1, False, 1.0, True, False, False, True))
    xc = torch.randn(8, 96, 28, 28, generator=gen)
    gc = torch.randn(8, 96, 28, 28, generator=gen)
    add("fp32_channel", xc, gc, 0.02 + 0.02 * torch.rand(96, generator=gen), -torch.rand(96, generator=gen),
        (0, 127, 0, 255, 1, False, 1.0, True, True, False, False))
    xh = torch.randn(16, 64, 28, 28, generator=gen).half()
    gh = torch.randn(16, 64, 28, 28, generator=gen).half()
    add("fp16_tensor", xh, gh, torch.tensor([0.03]).half(), torch.tensor([-1.7]).half(), (0, 127, 0, 255, 1, False, 1.0, True, False, False, False))
    add("fp16_channel", xh, gh, (0.02 + 0.02 * torch.rand(64, generator=gen)).half(), (-torch.rand(64, generator=gen)).half(),
        (0, 127, 0, 255, 1, False, 1.0, True, True, False, False))
    inp, outp = tmp_path / "in.pt", tmp_path / "out.pt"
    torch.save(cases, inp)
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "-c", _REF_SCRIPT, str(ROOT / "oracle" / "_ref"), str(inp), str(outp)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    ref = torch.load(outp)
    report = {}
    for name, c in cases.items():
        x = c["x"].to(U.DEV).requires_grad_(True)
        s = c["s"].to(U.DEV).requires_grad_(True)
        b = c["b"].to(U.DEV).requires_grad_(True)
        y = lsq(x, s, b, *c["args"])
        y.backward(c["g"].to(U.DEV))
        r_ = ref[name]
        # forward and grad_x: bit-exact with the reference CUDA op
        assert torch.equal(y.detach().cpu().view(torch.int16 if y.dtype == torch.float16 else torch.int32),
                           r_["y"].view(torch.int16 if y.dtype == torch.float16 else torch.int32)), name
        assert torch.equal(x.grad.cpu(), r_["gx"]), name
        gs, gb = s.grad.float().cpu(), (b.grad.float().cpu() if b.grad is not None else torch.zeros_like(r_["gb"]).float())
        if c["x"].dtype == torch.float32:
            # the reference sums in fp32 (at::sum); agree to a few 1e-6 of the summed magnitude
            assert torch.allclose(gs, r_["gs"].float(), rtol=2e-5, atol=2e-6 * float(r_["gs"].abs().max())), (name, gs[:3], r_["gs"][:3])
            assert torch.allclose(gb, r_["gb"].float(), rtol=2e-5, atol=2e-6 * float(r_["gb"].abs().max()) + 1e-12), name
        else:
            # fp16 reference: per-term half rounding + half result (use_grad_scaling off, D7)
            assert torch.allclose(gs, r_["gs"].float(), rtol=4e-3, atol=4e-3 * float(r_["gs"].float().abs().max())), (name, gs[:3], r_["gs"][:3])
        report[name] = dict(gs_mine=float(gs.flatten()[0]), gs_ref=float(r_["gs"].float().flatten()[0]))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "ref_cuda_parity.json").write_text(json.dumps(U.stamped(report), indent=1))


# ------------------------------------------------------------------------------------------------
# fused observer step (init_mode='observer') vs torch's own observers + the module's host logic
# ------------------------------------------------------------------------------------------------
def _torch_observer_path(obs, x, scale, shift):
    """what observers.py:446-449 + :346-373 do: observer forward, calculate_qparams, _set_weights"""
    obs(x.detach())
    s, zp = obs.calculate_qparams()
    with torch.no_grad():
        scale.copy_(s.to(scale.device).to(scale.dtype).reshape(scale.shape))
        shift.copy_((-zp.to(scale.device) * scale).to(shift.dtype).reshape(shift.shape))


@pytest.mark.parametrize("cls,kw,per_channel", [
    (torch.quantization.MovingAverageMinMaxObserver, dict(dtype=torch.quint8, qscheme=torch.per_tensor_affine, reduce_range=True), False),
    (torch.quantization.MovingAverageMinMaxObserver, dict(dtype=torch.quint8, qscheme=torch.per_tensor_affine, averaging_constant=0.3), False),
    (torch.quantization.MinMaxObserver, dict(dtype=torch.quint8, qscheme=torch.per_tensor_symmetric), False),
    (torch.quantization.MinMaxObserver, dict(dtype=torch.qint8, qscheme=torch.per_tensor_symmetric, quant_min=-64, quant_max=63), False),
    (torch.quantization.MovingAveragePerChannelMinMaxObserver, dict(dtype=torch.quint8, qscheme=torch.per_channel_affine, ch_axis=1), True),
    (torch.quantization.PerChannelMinMaxObserver, dict(dtype=torch.qint8, qscheme=torch.per_channel_symmetric, ch_axis=0), True),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_observer_step_bit_exact_vs_torch_observers(cls, kw, per_channel, dtype):
    import warnings
    from torchlsq.quantized.modules.observers import observer_step
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = cls(**kw).to(U.DEV)
        nat = cls(**kw).to(U.DEV)
    ax = kw.get("ch_axis", None)
    shape = (6, 24, 10, 10)
    n = shape[ax] if per_channel else 1
    s_ref, b_ref = torch.ones(n, device=U.DEV), torch.zeros(n, device=U.DEV)
    s_nat, b_nat = torch.ones(n, device=U.DEV), torch.zeros(n, device=U.DEV)
    gen = torch.Generator().manual_seed(3)
    for step in range(5):
        x = (torch.randn(*shape, generator=gen) * (1 + step) + 0.3 * step).to(dtype).to(U.DEV)
        if step == 3:
            x = x.abs()                       # all-positive batch: min_neg clamps at 0
        _torch_observer_path(ref, x, s_ref, b_ref)
        assert observer_step(nat, x, s_nat, b_nat)
        assert torch.equal(nat.min_val.reshape(-1), ref.min_val.reshape(-1)), (step, nat.min_val, ref.min_val)
        assert torch.equal(nat.max_val.reshape(-1), ref.max_val.reshape(-1)), step
        assert torch.equal(s_nat, s_ref), (step, s_nat, s_ref)
        assert torch.equal(b_nat, b_ref), (step, b_nat, b_ref)
    # qparams exported by the torch observer object agree (its state was updated in place)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s1, z1 = nat.calculate_qparams()
        s2, z2 = ref.calculate_qparams()
    assert torch.equal(s1, s2) and torch.equal(z1, z2)


def test_observer_step_large_tensor_and_nan():
    from torchlsq.quantized.modules.observers import observer_step
    obs = torch.quantization.MinMaxObserver(dtype=torch.quint8, qscheme=torch.per_tensor_affine).to(U.DEV)
    x = torch.empty(64 * 256 * 56 * 56, dtype=torch.bfloat16, device=U.DEV).normal_(0, 2, generator=torch.Generator(U.DEV).manual_seed(0))
    s, b = torch.ones(1, device=U.DEV), torch.zeros(1, device=U.DEV)
    assert observer_step(obs, x, s, b)
    mn, mx = torch.aminmax(x.float())
    assert obs.min_val.item() == mn.item() and obs.max_val.item() == mx.item()
    assert s.item() == pytest.approx((mx.item() - mn.item()) / 255.0, rel=1e-6)
    x[12345] = float("nan")                   # torch.aminmax propagates NaN; so does the fused step
    assert observer_step(obs, x, s, b)
    assert torch.isnan(obs.min_val).item() and torch.isnan(obs.max_val).item()
    # unsupported observer kinds are left to torch
    h = torch.quantization.HistogramObserver().to(U.DEV)
    assert observer_step(h, x[:1000], s, b) is False


def test_module_observer_mode_native_equals_torch_path():
    from torchlsq import LSQFakeQuantizer
    MA = torch.quantization.MovingAverageMinMaxObserver
    torch.manual_seed(0)
    a = LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=3).to(U.DEV)
    b = LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=3).to(U.DEV)
    b.native_observer = False
    a.train(); b.train()
    for i in range(7):
        x = (torch.randn(8, 16, 12, 12, device=U.DEV) * (1 + 0.2 * i)).to(torch.bfloat16)
        ya, yb = a(x), b(x)
        assert torch.equal(ya, yb), i
        if i >= 1:
            assert torch.equal(a.scale, b.scale) and torch.equal(a.shift, b.shift), i
    assert int(a.observer_enabled[0]) == 0 and int(b.observer_enabled[0]) == 0


def test_weight_init_stats_kernel_variants_bit_identical():
    """The statistics kernels of a plan - descriptor form (round 1), row-entry form (default) and the bulk-copy ring - share one
    arithmetic and one summation order: the 27 560 ResNet-50 scales must agree bit for bit, rows that are not whole 32-byte
    units (K = 147) included, and with the single-tensor call."""
    import bench as B
    from torchlsq import _cabi
    from torchlsq.multi import LSQPlan, Site
    from torchlsq.quantized.modules.observers import weight_init_scale
    lib = _cabi.load()
    gen = torch.Generator(device=U.DEV).manual_seed(5)
    try:
        for dtype in (torch.float32, torch.bfloat16):
            ws = [torch.empty(s, device=U.DEV).normal_(0.01, 0.05, generator=gen).to(dtype) for s in B.W_SHAPES[:12] + B.W_SHAPES[-3:]]
            sites = [Site(x=w, scale=torch.ones(w.shape[0], device=U.DEV), shift=torch.zeros(w.shape[0], device=U.DEV), quant_min=-128,
                          quant_max=127, type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True) for w in ws]
            outs = {}
            for v in (2, 4, 5, 7):
                lib.lsqb200_set_tuning(f"rowstats={v}".encode())
                plan = LSQPlan(sites)                      # the kernel family is chosen at plan creation
                a = plan.weight_init_stats()
                b = plan.weight_init_stats(torch.empty_like(a))          # another output buffer: entries are relative, still right
                assert torch.equal(a, b)
                outs[v] = a.clone()
                plan.close()
            for v in (4, 5, 7):
                assert torch.equal(outs[v].view(torch.int32), outs[2].view(torch.int32)), (dtype, v)
            single = torch.cat([weight_init_scale(w, 0, True, -128, 127) for w in ws])
            assert torch.equal(single.view(torch.int32), outs[2].view(torch.int32)), dtype
    finally:
        lib.lsqb200_set_tuning(b"")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_input_without_grad_skips_grad_x_but_not_the_parameter_sums(dtype):
    """The C++ autograd layer hands the kernels a NULL grad_x when the input needs no gradient (a network's first quantizer, frozen
    features): two reads, no write.  The parameter gradients must be exactly those of the call that also produces grad_x, per tensor,
    per channel (row-tiled and column layouts) and behind a fused ReLU; layouts that are dense but not contiguous included."""
    from torchlsq.functional import lsq, lsq_relu
    gen = torch.Generator().manual_seed(11)
    cases = [((4, 16, 28, 28), None), ((4, 16, 28, 28), 1), ((8, 32, 7, 7), 1), ((64, 24), 1), ((32, 16, 3, 3), 0)]
    for shape, axis in cases:
        for fn in (lsq, lsq_relu):
            C = 1 if axis is None else shape[axis]
            x0 = torch.randn(*shape, generator=gen).to(dtype).to(U.DEV)
            if len(shape) == 4 and axis == 1:
                x0 = x0.contiguous(memory_format=torch.channels_last)      # dense, not contiguous: processed in memory order
            g = torch.randn(*shape, generator=gen).to(dtype).to(U.DEV)
            grads = []
            for need_x in (True, False):
                x = x0.clone().requires_grad_(need_x)
                s = (0.02 + 0.02 * torch.rand(C, generator=torch.Generator().manual_seed(3))).to(U.DEV).requires_grad_(True)
                b = (-torch.rand(C, generator=torch.Generator().manual_seed(4))).to(U.DEV).requires_grad_(True)
                y = fn(x, s, b, 0, 127, 0, 255, axis=0 if axis is None else axis, is_perchannel=axis is not None)
                y.backward(g)
                assert (x.grad is not None) == need_x
                grads.append((y.detach(), s.grad.clone(), b.grad.clone()))
            assert torch.equal(grads[0][0], grads[1][0])
            # same kernel, same reduction order: bit-identical unless the layout takes the column path (fp64 atomics in arrival order)
            assert torch.allclose(grads[0][1], grads[1][1], rtol=1e-6, atol=0) and torch.allclose(grads[0][2], grads[1][2], rtol=1e-6, atol=0), (shape, axis, fn.__name__)
