"""GPU parity of the fused ReLU prologue (include/lsq_b200.h LSQB200_PRE_RELU).

Through the C ABI against the oracle's restatement of `torch.relu` -> reference op -> autograd (oracle.forward_relu /
backward_relu, pinned to the reference's own CPU results by tests/test_oracle_relu.py): forward and grad_x BIT-EXACT,
grad_scale / grad_shift within 1e-6 (fp32) / 1e-5 (fp16, bf16 tensors) relative.  Through the public op against the
unfused sequence `lsq(torch.relu(x))` running on the same GPU (torch's CUDA relu + the plain kernels + autograd through
both): y and x.grad bit-identical, parameter gradients equal to 1e-6.
"""
import numpy as np
import pytest
import torch

import gpu_util as U
from conftest import geometry
from torchlsq import _cabi

pytestmark = pytest.mark.gpu

DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
RELU = _cabi.PRE_RELU
SPECIALS = [0.0, -0.0, float("nan"), float("inf"), -float("inf"), 1e-30, -1e-30, 1e-45, -1e-45, 65504.0, -65504.0, 3.3e38, -3.3e38]


def _mk(n, dtype, seed, scale=1.5, specials=True):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=gen) * scale
    g = torch.randn(n, generator=gen)
    if specials and n >= 64:
        idx = torch.randperm(n, generator=gen)[:len(SPECIALS)]
        x[idx] = torch.tensor(SPECIALS)
    return x.to(dtype).to(U.DEV), g.to(dtype).to(U.DEV)


def _params(vals_s, vals_b, dtype=torch.float32):
    return (torch.tensor(vals_s, dtype=dtype, device=U.DEV).reshape(-1),
            torch.tensor(vals_b, dtype=dtype, device=U.DEV).reshape(-1))


def _check(x, g, s, b, q, outer=1, C=1, inner=None, per_channel=False, rel=1e-6):
    y = U.fwd(x, s, b, q, outer, C, inner, per_channel, prologue=RELU)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, per_channel, relu=True)), "forward"
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, per_channel, prologue=RELU)
    ogx, ogs, ogb, mag_s, mag_b = U.oracle_bwd(g, x, s, b, q, outer, C, inner, per_channel, relu=True)
    assert U.same_bits(gx, ogx), "grad_x"
    finite = np.isfinite(ogs) & np.isfinite(ogb)
    if finite.all():
        U.assert_grads_close(gs, ogs, mag_s, rel, "gscale")
        U.assert_grads_close(gb, ogb, mag_b, rel, "gshift")
    return y, gx, gs, gb


# ------------------------------------------------------------------------------------------------
# C ABI vs oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 255, 4099, (1 << 20) + 3])
@pytest.mark.parametrize("shift", [-1.7, 0.0, 0.9])
def test_relu_tensor_vs_oracle(dt, n, shift):
    # shift 0: zp == quant_min (the clamp already is the ReLU); -1.7: relu'd zeros land inside the range; 0.9: zp clamps to 0
    x, g = _mk(n, DT[dt], seed=n, specials=False)
    s, b = _params([0.03], [shift])
    _check(x, g, s, b, U.qa(), rel=1e-6 if dt == "f32" else 1e-5)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_relu_tensor_special_values_bitwise(dt):
    """-0, NaN, +-inf, denormals, huge values: forward and grad_x bit for bit (sums are NaN / inf there: not compared)."""
    x, g = _mk(70001, DT[dt], seed=5)
    s, b = _params([0.03], [-1.7])
    for mode in (dict(), dict(init_mode=True), dict(eval_mode=True), dict(eval_mode=True, init_mode=True)):
        _check(x, g, s, b, U.qa(**mode))


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("mode", ["init", "eval", "eval_init", "sym", "no_gs"])
def test_relu_tensor_modes(dt, mode):
    x, g = _mk(300_001, DT[dt], seed=11, specials=False)
    kw = dict(init=dict(init_mode=True), eval=dict(eval_mode=True), eval_init=dict(eval_mode=True, init_mode=True),
              sym=dict(qmin=-64, qmax=63, tmin=-128, tmax=127, sym=True), no_gs=dict(use_gs=False, gscaler=0.5))[mode]
    s, b = _params([0.03], [0.0 if mode == "sym" else -1.7])
    y, gx, gs, gb = _check(x, g, s, b, U.qa(**kw), rel=1e-6 if dt == "f32" else 1e-5)
    if mode in ("eval", "eval_init"):
        assert gs.item() == 0.0 and gb.item() == 0.0
    if mode in ("init", "eval_init"):           # learned init behind the prologue: y = relu(x), gx = x > 0 ? g : 0
        assert torch.equal(y, torch.relu(x)) and torch.equal(gx, torch.where(x > 0, g, torch.zeros_like(g)))


def test_relu_misaligned_and_without_gx():
    base_x, base_g = _mk(40_000, torch.bfloat16, seed=3, specials=False)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    for off in (1, 2, 3, 5, 8):
        x, g = base_x[off:off + 30_001], base_g[off:off + 30_001]
        _check(x, g, s, b, q, rel=1e-5)
    x, g = base_x[:30_000], base_g[:30_000]
    _, gs0, gb0 = U.bwd(g, x, s, b, q, prologue=RELU)
    none, gs1, gb1 = U.bwd(g, x, s, b, q, want_gx=False, prologue=RELU)
    assert none is None and torch.equal(gs0, gs1) and torch.equal(gb0, gb1)


@pytest.mark.parametrize("dt,shape,axis", [
    ("f32", (4, 16, 56, 56), 1),     # row-tiled, long rows
    ("f16", (8, 64, 14, 14), 1),     # column layout (392-byte rows)
    ("bf16", (8, 96, 7, 7), 1),      # column layout, rows that straddle units
    ("bf16", (32, 7, 7, 128), 3),    # channels-last style: inner == 1
    ("f32", (64, 32, 3, 3), 0),      # weight-like rows owned by warp groups
    ("f16", (3, 5, 7), 1),           # scalar fallback
])
def test_relu_channel_vs_oracle(dt, shape, axis):
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=sum(shape), specials=False)
    C = shape[axis]
    gen = torch.Generator().manual_seed(C)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen) * 2 + 0.5).to(U.DEV)      # both signs of shift: zp > 0 and zp clamped to 0
    outer, C_, inner = geometry(shape, axis)
    for mode in (dict(), dict(init_mode=True), dict(eval_mode=True)):
        _check(x, g, s, b, U.qa(**mode), outer, C_, inner, True, rel=1e-6 if dt == "f32" else 1e-5)


def test_relu_abi_errors_and_empty():
    lib = _cabi.load()
    x, g = _mk(1024, torch.float16, seed=1, specials=False)
    q = U.qa()
    sh, bh = _params([0.03], [-1.7], torch.float16)
    y = torch.empty_like(x)
    rc = lib.lsqb200_fwd_tensor_pre(x.data_ptr(), None, y.data_ptr(), sh.data_ptr(), bh.data_ptr(), x.numel(), _cabi.F16, _cabi.F16, q,
                                    RELU, U.stream())
    assert rc == -2 and b"prologue" in lib.lsqb200_last_error()          # c10::Half-exact contract has no fused prologue
    s, b = _params([0.03], [-1.7])
    rc = lib.lsqb200_fwd_tensor_pre(x.data_ptr(), None, y.data_ptr(), s.data_ptr(), b.data_ptr(), x.numel(), _cabi.F16, _cabi.F32, q,
                                    7, U.stream())
    assert rc == -1 and b"unknown prologue" in lib.lsqb200_last_error()
    xd = torch.zeros(8, dtype=torch.float64, device=U.DEV)
    sd = torch.ones(1, dtype=torch.float64, device=U.DEV)
    rc = lib.lsqb200_fwd_tensor_pre(xd.data_ptr(), None, xd.data_ptr(), sd.data_ptr(), sd.data_ptr(), 8, _cabi.F64, _cabi.F64, q,
                                    RELU, U.stream())
    assert rc == -2
    # prologue 0 through the _pre entry == the plain call
    y0 = U.fwd(x, s, b, q)
    rc = lib.lsqb200_fwd_tensor_pre(x.data_ptr(), None, y.data_ptr(), s.data_ptr(), b.data_ptr(), x.numel(), _cabi.F16, _cabi.F32, q,
                                    0, U.stream())
    assert rc == 0 and torch.equal(y, y0)
    # empty input: defined zero parameter gradients
    e = torch.empty(0, dtype=torch.float32, device=U.DEV)
    gx, gs, gb = U.bwd(e, e, s, b, q, prologue=RELU)
    assert gs.item() == 0.0 and gb.item() == 0.0
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# public op vs the unfused sequence on the same GPU
# ------------------------------------------------------------------------------------------------
def _run(fn, x, s, b, g, **kw):
    x = x.clone().requires_grad_(True)
    s = s.clone().requires_grad_(True)
    b = b.clone().requires_grad_(True)
    y = fn(x, s, b, **kw)
    y.backward(g)
    return y.detach(), x.grad, s.grad, (b.grad if b.grad is not None else torch.zeros_like(b))


def _bits(t):
    return t.contiguous().view(torch.int32 if t.dtype == torch.float32 else torch.int16)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("kind", ["tensor", "tensor_init", "channel_nchw", "channel_7x7", "channel_last"])
def test_lsq_relu_equals_relu_then_lsq_on_gpu(dt, kind):
    from torchlsq.functional import lsq, lsq_relu
    shape = dict(tensor=(8, 64, 28, 28), tensor_init=(8, 64, 28, 28), channel_nchw=(4, 32, 56, 56), channel_7x7=(16, 128, 7, 7),
                 channel_last=(16, 7, 7, 128))[kind]
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=len(kind))
    x, g = x.reshape(shape), g.reshape(shape)
    kw = dict(quant_min=0, quant_max=127, type_min=0, type_max=255)
    if kind.startswith("tensor"):
        s, b = _params([0.03], [-1.7])
        kw["init_mode"] = kind == "tensor_init"
    else:
        axis = 3 if kind == "channel_last" else 1
        C = shape[axis]
        gen = torch.Generator().manual_seed(C)
        s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
        b = (-torch.rand(C, generator=gen) * 2 + 0.5).to(U.DEV)
        kw.update(axis=axis, is_perchannel=True)
    y1, gx1, gs1, gb1 = _run(lsq_relu, x, s, b, g, **kw)
    y0, gx0, gs0, gb0 = _run(lambda x_, s_, b_, **k: lsq(torch.relu(x_), s_, b_, **k), x, s, b, g, **kw)
    nan = torch.isnan(y0)
    assert torch.equal(torch.isnan(y1), nan) and torch.equal(_bits(y1)[~nan], _bits(y0)[~nan])
    nang = torch.isnan(gx0)
    assert torch.equal(torch.isnan(gx1), nang) and torch.equal(_bits(gx1)[~nang], _bits(gx0)[~nang])
    # x holds NaN / inf on purpose: the sums are NaN in both; compare a clean copy for the values
    xc = torch.nan_to_num(x, nan=0.5, posinf=3.0, neginf=-3.0)
    _, _, gs1, gb1 = _run(lsq_relu, xc, s, b, g, **kw)
    _, _, gs0, gb0 = _run(lambda x_, s_, b_, **k: lsq(torch.relu(x_), s_, b_, **k), xc, s, b, g, **kw)
    assert torch.allclose(gs1, gs0, rtol=1e-6, atol=1e-9) and torch.allclose(gb1, gb0, rtol=1e-6, atol=1e-9)


def test_lsq_relu_rejects_reference_exact_inputs_and_cpu():
    from torchlsq.functional import lsq_relu
    x = torch.randn(64, device=U.DEV, dtype=torch.float16)
    s, b = _params([0.03], [-1.7], torch.float16)
    with pytest.raises(RuntimeError, match="fused-prologue lsq needs"):
        lsq_relu(x, s, b, 0, 127)
    xd = torch.randn(64, device=U.DEV, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="fused-prologue lsq needs"):
        lsq_relu(xd, s.double(), b.double(), 0, 127)
    with pytest.raises(RuntimeError, match="CUDA"):
        lsq_relu(x.cpu().float(), s.cpu().float(), b.cpu().float(), 0, 127)
    xs = torch.randn(64, device=U.DEV, requires_grad=True)
    s32, b32 = _params([0.03], [-1.7])
    y = lsq_relu(xs, s32.requires_grad_(True), b32, 0, 127)
    (gx,) = torch.autograd.grad(y.sum(), xs, create_graph=False)
    assert gx.shape == xs.shape
    with pytest.raises(RuntimeError, match="double backwards"):
        y2 = lsq_relu(xs, s32, b32, 0, 127)
        torch.autograd.grad(y2.sum(), xs, create_graph=True)


def test_module_fuse_relu_equals_relu_module_sequence():
    """LSQFakeQuantizer(fuse_relu=True) == Sequential(ReLU(), LSQFakeQuantizer()) through initialisation (observer and
    learned), the hand-over and steady-state training steps: outputs, input gradients and parameters bit-identical."""
    from torchlsq import LSQFakeQuantizer
    MA = torch.quantization.MovingAverageMinMaxObserver
    for init_mode in ("observer", "learnable"):
        def build(fuse):
            m = LSQFakeQuantizer(MA, "activation", init_batches=3, init_mode=init_mode, fuse_relu=fuse).to(U.DEV)
            return m
        fused, plain = build(True), build(False)
        relu = torch.nn.ReLU()
        opt = None
        gen = torch.Generator().manual_seed(9)
        for step in range(8):
            x = (torch.randn(4, 16, 14, 14, generator=gen) * 2).to(U.DEV)
            g = torch.randn(4, 16, 14, 14, generator=gen).to(U.DEV)
            xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
            ya, yb = fused(xa), plain(relu(xb))
            assert torch.equal(ya, yb), (init_mode, step)
            if ya.requires_grad:
                ya.backward(g)
                yb.backward(g)
                assert torch.equal(xa.grad, xb.grad), (init_mode, step)
            if opt is None:                      # parameters exist after the first forward
                opt = (torch.optim.SGD(fused.parameters(), lr=1e-3), torch.optim.SGD(plain.parameters(), lr=1e-3))
            else:
                for o in opt:
                    o.step()
                    o.zero_grad()
            assert torch.allclose(fused.scale, plain.scale, rtol=1e-6, atol=0) and torch.allclose(fused.shift, plain.shift, rtol=1e-6, atol=1e-9)
        fused.eval(), plain.eval()
        x = torch.randn(4, 16, 14, 14, generator=gen).to(U.DEV)
        assert torch.allclose(fused(x), plain(relu(x)), rtol=1e-5, atol=1e-6)


def test_plan_with_fused_relu_sites_matches_calls():
    from torchlsq.functional import lsq_relu
    from torchlsq.multi import LSQPlan, Site
    gen = torch.Generator().manual_seed(21)
    sites, refs = [], []
    for i, shape in enumerate([(4, 8, 28, 28), (2, 16, 14, 14), (4099,)]):
        x = (torch.randn(shape, generator=gen) * 1.5).to(torch.bfloat16).to(U.DEV)
        g = torch.randn(shape, generator=gen).to(torch.bfloat16).to(U.DEV)
        s, b = _params([0.03 + 0.01 * i], [-1.7 + i])
        fuse = i != 1
        st = Site(x=x, scale=s, shift=b, y=torch.empty_like(x), grad=g, gx=torch.empty_like(x), gscale=torch.empty_like(s),
                  gshift=torch.empty_like(b), quant_min=0, quant_max=127, type_min=0, type_max=255, fuse_relu=fuse)
        sites.append(st)
        xr = x.clone().requires_grad_(True)
        sr, br = s.clone().requires_grad_(True), b.clone().requires_grad_(True)
        fn = lsq_relu if fuse else __import__("torchlsq.functional", fromlist=["lsq"]).lsq
        y = fn(xr, sr, br, 0, 127, 0, 255)
        y.backward(g)
        refs.append((y.detach(), xr.grad, sr.grad, br.grad))
    plan = LSQPlan(sites)
    plan.forward()
    plan.backward()
    torch.cuda.synchronize()
    for st, (y, gx, gs, gb) in zip(sites, refs):
        assert torch.equal(st.y, y) and torch.equal(st.gx, gx)
        assert torch.equal(st.gscale, gs) and torch.equal(st.gshift, gb)
    plan.close()
