"""GPU parity of the fused residual-join prologues (LSQB200_PRE_ADD_RELU / LSQB200_PRE_ADD):
fake_quant(relu(a + b)) and fake_quant(a + b) from a and b in one pass.

Through the C ABI against the oracle's restatement of ATen add [-> relu] -> reference op -> autograd (oracle.forward_add /
backward_add, pinned to the reference's own CPU results by tests/test_oracle_relu.py): forward and grad_x BIT-EXACT
(the sum is rounded to the tensor type as `a + b` stores it), parameter sums within 1e-6 (fp32) / 1e-5 (16-bit tensors).
Through the public op against the unfused sequence on the same GPU: y, a.grad, b.grad bit-identical.
"""
import numpy as np
import pytest
import torch

import gpu_util as U
from conftest import geometry
from torchlsq import _cabi

pytestmark = pytest.mark.gpu

DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
PRE = {"add_relu": _cabi.PRE_ADD_RELU, "add": _cabi.PRE_ADD}
SPECIALS = [0.0, -0.0, float("nan"), float("inf"), -float("inf"), 1e-30, -1e-30, 1e-45, 65504.0, -65504.0, 3.3e38, -3.3e38]


def _mk(n, dtype, seed, specials=False):
    gen = torch.Generator().manual_seed(seed)
    a = torch.randn(n, generator=gen) * 1.2
    b = torch.randn(n, generator=gen) * 0.9
    g = torch.randn(n, generator=gen)
    if specials and n >= 64:
        idx = torch.randperm(n, generator=gen)[:2 * len(SPECIALS)]
        a[idx[:len(SPECIALS)]] = torch.tensor(SPECIALS)
        b[idx[len(SPECIALS) // 2:len(SPECIALS) // 2 + len(SPECIALS)]] = torch.tensor(SPECIALS[::-1])   # overlaps: inf - inf, -0 + 0 ...
    return a.to(dtype).to(U.DEV), b.to(dtype).to(U.DEV), g.to(dtype).to(U.DEV)


def _params(vals_s, vals_b, dtype=torch.float32):
    return (torch.tensor(vals_s, dtype=dtype, device=U.DEV).reshape(-1),
            torch.tensor(vals_b, dtype=dtype, device=U.DEV).reshape(-1))


def _check(kind, a, b2, g, s, b, q, outer=1, C=1, inner=None, per_channel=False, rel=1e-6):
    relu = kind == "add_relu"
    y = U.fwd(a, s, b, q, outer, C, inner, per_channel, prologue=PRE[kind], x2=b2)
    assert U.same_bits(y, U.oracle_fwd(a, s, b, q, outer, C, inner, per_channel, relu=relu, x2=b2)), "forward"
    gx, gs, gb = U.bwd(g, a, s, b, q, outer, C, inner, per_channel, prologue=PRE[kind], x2=b2)
    ogx, ogs, ogb, mag_s, mag_b = U.oracle_bwd(g, a, s, b, q, outer, C, inner, per_channel, relu=relu, x2=b2)
    assert U.same_bits(gx, ogx), "grad_x"
    if (np.isfinite(ogs) & np.isfinite(ogb)).all():
        U.assert_grads_close(gs, ogs, mag_s, rel, "gscale")
        U.assert_grads_close(gb, ogb, mag_b, rel, "gshift")
    return y, gx, gs, gb


@pytest.mark.parametrize("kind", ["add_relu", "add"])
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 255, 4099, (1 << 20) + 3])
def test_add_tensor_vs_oracle(kind, dt, n):
    a, b2, g = _mk(n, DT[dt], seed=n)
    s, b = _params([0.03], [-1.7])
    _check(kind, a, b2, g, s, b, U.qa(), rel=1e-6 if dt == "f32" else 1e-5)


@pytest.mark.parametrize("kind", ["add_relu", "add"])
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_add_special_values_and_modes_bitwise(kind, dt):
    a, b2, g = _mk(70001, DT[dt], seed=5, specials=True)
    s, b = _params([0.03], [-1.7])
    for mode in (dict(), dict(init_mode=True), dict(eval_mode=True), dict(eval_mode=True, init_mode=True)):
        _check(kind, a, b2, g, s, b, U.qa(**mode))


@pytest.mark.parametrize("kind", ["add_relu", "add"])
def test_add_learned_init_is_the_rounded_sum(kind):
    a, b2, g = _mk(300_001, torch.bfloat16, seed=12)
    s, b = _params([0.03], [-1.7])
    y, gx, gs, gb = _check(kind, a, b2, g, s, b, U.qa(init_mode=True), rel=1e-5)
    want = a + b2
    if kind == "add_relu":
        assert torch.equal(y, torch.relu(want)) and torch.equal(gx, torch.where(want > 0, g, torch.zeros_like(g)))
    else:
        assert torch.equal(y, want) and torch.equal(gx, g)


def test_add_misaligned_operands_and_missing_x2():
    lib = _cabi.load()
    A, B2, G = _mk(40_000, torch.float16, seed=3)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    for oa, ob in ((0, 1), (1, 0), (3, 5), (8, 16), (16, 2)):       # the narrowest operand alignment picks the unit width
        _check("add_relu", A[oa:oa + 20_001], B2[ob:ob + 20_001], G[:20_001], s, b, q, rel=1e-5)
    y = torch.empty_like(A)
    rc = lib.lsqb200_fwd_tensor_pre(A.data_ptr(), None, y.data_ptr(), s.data_ptr(), b.data_ptr(), A.numel(), _cabi.F16, _cabi.F32, q,
                                    _cabi.PRE_ADD_RELU, U.stream())
    assert rc == -1 and b"x2" in lib.lsqb200_last_error()
    torch.cuda.synchronize()


@pytest.mark.parametrize("kind", ["add_relu", "add"])
@pytest.mark.parametrize("dt,shape,axis", [
    ("f32", (4, 16, 56, 56), 1),     # long rows
    ("bf16", (8, 64, 14, 14), 1),    # short rows: no column-layout kernel for the two-operand prologues, row-tiled path
    ("f16", (8, 96, 7, 7), 1),       # rows that are not a unit multiple: scalar path
    ("f32", (64, 32, 3, 3), 0),      # warp-group rows
])
def test_add_channel_vs_oracle(kind, dt, shape, axis):
    n = int(np.prod(shape))
    a, b2, g = _mk(n, DT[dt], seed=sum(shape))
    C = shape[axis]
    gen = torch.Generator().manual_seed(C)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen) * 2 + 0.5).to(U.DEV)
    outer, C_, inner = geometry(shape, axis)
    for mode in (dict(), dict(init_mode=True)):
        _check(kind, a, b2, g, s, b, U.qa(**mode), outer, C_, inner, True, rel=1e-6 if dt == "f32" else 1e-5)


def _bits(t):
    return t.contiguous().view(torch.int32 if t.dtype == torch.float32 else torch.int16)


def _eq_bits(t1, t0):
    nan = torch.isnan(t0)
    return torch.equal(torch.isnan(t1), nan) and torch.equal(_bits(t1)[~nan], _bits(t0)[~nan])


@pytest.mark.parametrize("kind", ["add_relu", "add"])
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("layout", ["contiguous", "channels_last", "mixed", "per_channel"])
def test_public_op_equals_unfused_sequence_on_gpu(kind, dt, layout):
    from torchlsq.functional import lsq, lsq_add, lsq_add_relu
    fused = lsq_add_relu if kind == "add_relu" else lsq_add
    shape = (8, 32, 28, 28)
    a, b2, g = _mk(int(np.prod(shape)), DT[dt], seed=len(layout), specials=True)
    a, b2, g = a.reshape(shape), b2.reshape(shape), g.reshape(shape)
    if layout in ("channels_last", "mixed"):
        a = a.contiguous(memory_format=torch.channels_last)
    if layout == "channels_last":
        b2 = b2.contiguous(memory_format=torch.channels_last)
    kw = dict(quant_min=0, quant_max=127, type_min=0, type_max=255)
    if layout == "per_channel":
        gen = torch.Generator().manual_seed(1)
        s = (0.02 + 0.02 * torch.rand(32, generator=gen)).to(U.DEV)
        b = (-torch.rand(32, generator=gen) * 2 + 0.5).to(U.DEV)
        kw.update(axis=1, is_perchannel=True)
    else:
        s, b = _params([0.03], [-1.7])

    def run(fn, a_, b_):
        a_, b_ = a_.clone(memory_format=torch.preserve_format).requires_grad_(True), b_.clone(memory_format=torch.preserve_format).requires_grad_(True)
        s_, sh_ = s.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = fn(a_, b_, s_, sh_)
        y.backward(g)
        return y.detach(), a_.grad, b_.grad, s_.grad, sh_.grad

    def unfused(a_, b_, s_, sh_):
        t = a_ + b_
        return lsq(torch.relu(t) if kind == "add_relu" else t, s_, sh_, **kw)

    y1, ga1, gb1, _, _ = run(lambda a_, b_, s_, sh_: fused(a_, b_, s_, sh_, **kw), a, b2)
    y0, ga0, gb0, _, _ = run(unfused, a, b2)
    assert _eq_bits(y1, y0) and _eq_bits(ga1, ga0) and _eq_bits(gb1, gb0)
    ac, bc = torch.nan_to_num(a, nan=0.5, posinf=3.0, neginf=-3.0), torch.nan_to_num(b2, nan=-0.25, posinf=1.0, neginf=-1.0)
    *_, gs1, gsh1 = run(lambda a_, b_, s_, sh_: fused(a_, b_, s_, sh_, **kw), ac, bc)
    *_, gs0, gsh0 = run(unfused, ac, bc)
    assert torch.allclose(gs1, gs0, rtol=1e-6, atol=1e-9) and torch.allclose(gsh1, gsh0, rtol=1e-6, atol=1e-9)


def test_public_op_argument_errors():
    from torchlsq.functional import lsq_add_relu
    s, b = _params([0.03], [-1.7])
    a = torch.randn(4, 8, device=U.DEV)
    with pytest.raises(RuntimeError, match="same shape"):
        lsq_add_relu(a, torch.randn(4, 7, device=U.DEV), s, b, 0, 127)
    with pytest.raises(RuntimeError, match="same shape"):
        lsq_add_relu(a, a.half(), s, b, 0, 127)
    with pytest.raises(RuntimeError, match="fused-prologue lsq needs"):
        lsq_add_relu(a.double(), a.double(), s.double(), b.double(), 0, 127)
    # one addend without grad, the other with: only that one receives it
    a1, a2 = a.clone().requires_grad_(True), a.clone()
    lsq_add_relu(a1, a2, s, b, 0, 127).sum().backward()
    assert a1.grad is not None and a2.grad is None


def test_plan_with_residual_sites_matches_calls():
    from torchlsq.functional import lsq_add, lsq_add_relu
    from torchlsq.multi import LSQPlan, Site
    gen = torch.Generator().manual_seed(33)
    sites, refs = [], []
    for i, (shape, relu) in enumerate([((4, 8, 28, 28), True), ((2, 16, 14, 14), False), ((4099,), True)]):
        a = (torch.randn(shape, generator=gen) * 1.5).to(torch.bfloat16).to(U.DEV)
        b2 = torch.randn(shape, generator=gen).to(torch.bfloat16).to(U.DEV)
        g = torch.randn(shape, generator=gen).to(torch.bfloat16).to(U.DEV)
        s, b = _params([0.03 + 0.01 * i], [-1.7 + i])
        st = Site(x=a, x2=b2, scale=s, shift=b, y=torch.empty_like(a), grad=g, gx=torch.empty_like(a), gscale=torch.empty_like(s),
                  gshift=torch.empty_like(b), quant_min=0, quant_max=127, type_min=0, type_max=255, fuse_relu=relu)
        sites.append(st)
        ar, sr, br = a.clone().requires_grad_(True), s.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = (lsq_add_relu if relu else lsq_add)(ar, b2, sr, br, 0, 127, 0, 255)
        y.backward(g)
        refs.append((y.detach(), ar.grad, sr.grad, br.grad))
    plan = LSQPlan(sites)
    plan.forward()
    plan.backward()
    torch.cuda.synchronize()
    for st, (y, gx, gs, gb) in zip(sites, refs):
        assert torch.equal(st.y, y) and torch.equal(st.gx, gx)
        assert torch.equal(st.gscale, gs) and torch.equal(st.gshift, gb)
    plan.close()


def test_module_forward_add_falls_back_for_broadcast_or_promoting_adds():
    """forward_add with operands ATen would broadcast / type-promote keeps ATen's semantics (separate add) - same values as
    the module applied to relu(a + b); matching operands take the fused kernels and give the same bits."""
    from torchlsq import LSQFakeQuantizer
    m = LSQFakeQuantizer(None, "activation", init_mode="learnable", init_batches=1, init_scale=0.05).to(U.DEV)
    a = torch.randn(4, 8, 6, 6, device=U.DEV)
    m(a)                                                   # creates the parameters
    m(a)                                                   # leaves the init window
    m.eval()
    bias = torch.randn(1, 8, 1, 1, device=U.DEV)
    with torch.no_grad():
        assert torch.equal(m.forward_add(a, bias), m(torch.relu(a + bias)))
        assert torch.equal(m.forward_add(a, bias.expand_as(a).contiguous()), m(torch.relu(a + bias)))     # fused path, same bits
        h = a.half()
        assert torch.equal(m.forward_add(h, bias), m(torch.relu(h + bias)))                              # promotes to float32
        assert torch.equal(m.forward_add(a, bias, relu=False), m(a + bias))
