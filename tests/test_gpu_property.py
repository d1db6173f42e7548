"""Property-based GPU parity (hypothesis): random shapes, axes, quantisation ranges, parameter
values, dtypes and modes through the C ABI, always against the CPU oracle; plus CUDA-graph capture
of the kernels (no host sync, no allocation inside the calls => capturable)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import gpu_util as U
from conftest import geometry

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.float16, torch.bfloat16]


@st.composite
def tensor_case(draw):
    rank = draw(st.integers(1, 4))
    shape = tuple(draw(st.integers(1, 9 if rank > 2 else 40)) for _ in range(rank))
    axis = draw(st.integers(0, rank - 1))
    per_channel = draw(st.booleans())
    dt = draw(st.sampled_from(DT))
    signed = draw(st.booleans())
    bits = draw(st.integers(2, 8))
    if signed:
        qmin, qmax, tmin, tmax = -(1 << (bits - 1)), (1 << (bits - 1)) - 1, -128, 127
    else:
        qmin, qmax, tmin, tmax = 0, (1 << bits) - 1, 0, 255
    mode = draw(st.sampled_from(["normal", "normal", "init", "eval", "sym"]))
    use_gs = draw(st.booleans())
    gscaler = draw(st.sampled_from([1.0, 0.5, 3.0]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    return shape, axis, per_channel, dt, (qmin, qmax, tmin, tmax), mode, use_gs, gscaler, seed


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(tensor_case())
def test_random_cases_match_oracle(case):
    shape, axis, per_channel, dt, (qmin, qmax, tmin, tmax), mode, use_gs, gscaler, seed = case
    gen = torch.Generator().manual_seed(seed)
    n = int(np.prod(shape))
    spread = float(torch.rand(1, generator=gen)) * 4 + 0.1
    x = (torch.randn(n, generator=gen) * spread).to(dt).to(U.DEV)
    g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
    if per_channel:
        outer, C, inner = geometry(shape, axis)
    else:
        outer, C, inner = 1, 1, n
    nparam = C if per_channel else 1
    s = (0.005 + 0.1 * torch.rand(nparam, generator=gen)) * torch.where(torch.rand(nparam, generator=gen) < 0.1, -1.0, 1.0)
    b = torch.randn(nparam, generator=gen) * (0.0 if mode == "sym" else 1.0)
    s, b = s.to(U.DEV), b.to(U.DEV)
    q = U.qa(qmin, qmax, tmin, tmax, use_gs, gscaler, sym=(mode == "sym"), eval_mode=(mode == "eval"), init_mode=(mode == "init"))
    y = U.fwd(x, s, b, q, outer, C, inner, per_channel)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, per_channel))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, per_channel)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, per_channel)
    assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == torch.float32 else 1e-5
    U.assert_grads_close(gs, ogs, ms, rel, "gscale")
    U.assert_grads_close(gb, ogb, mb, rel, "gshift")
    if mode == "eval":
        assert torch.count_nonzero(gs).item() == 0 and torch.count_nonzero(gb).item() == 0
    if mode == "sym":
        assert torch.count_nonzero(gb).item() == 0


@settings(max_examples=90, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(tensor_case(), st.sampled_from(["relu", "add_relu", "add"]))
def test_random_prologue_cases_match_oracle(case, prologue):
    """Fused prologues over random shapes, axes, ranges, parameters, dtypes and modes: C ABI vs the oracle's
    restatement of the separate passes (torch.relu / a + b -> reference op -> autograd)."""
    from torchlsq import _cabi
    shape, axis, per_channel, dt, (qmin, qmax, tmin, tmax), mode, use_gs, gscaler, seed = case
    code = {"relu": _cabi.PRE_RELU, "add_relu": _cabi.PRE_ADD_RELU, "add": _cabi.PRE_ADD}[prologue]
    relu = prologue != "add"
    gen = torch.Generator().manual_seed(seed)
    n = int(np.prod(shape))
    spread = float(torch.rand(1, generator=gen)) * 4 + 0.1
    x = (torch.randn(n, generator=gen) * spread).to(dt).to(U.DEV)
    x2 = (torch.randn(n, generator=gen) * spread * 0.7).to(dt).to(U.DEV) if prologue != "relu" else None
    g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
    outer, C, inner = geometry(shape, axis) if per_channel else (1, 1, n)
    nparam = C if per_channel else 1
    s = (0.005 + 0.1 * torch.rand(nparam, generator=gen)) * torch.where(torch.rand(nparam, generator=gen) < 0.1, -1.0, 1.0)
    b = torch.randn(nparam, generator=gen) * (0.0 if mode == "sym" else 1.0)
    s, b = s.to(U.DEV), b.to(U.DEV)
    q = U.qa(qmin, qmax, tmin, tmax, use_gs, gscaler, sym=(mode == "sym"), eval_mode=(mode == "eval"), init_mode=(mode == "init"))
    y = U.fwd(x, s, b, q, outer, C, inner, per_channel, prologue=code, x2=x2)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, per_channel, relu=relu, x2=x2))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, per_channel, prologue=code, x2=x2)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, per_channel, relu=relu, x2=x2)
    assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == torch.float32 else 1e-5
    U.assert_grads_close(gs, ogs, ms, rel, "gscale")
    U.assert_grads_close(gb, ogb, mb, rel, "gshift")


def test_cuda_graph_capture_and_replay():
    """forward + backward of two sites captured into one CUDA graph and replayed on new data."""
    from torchlsq import _cabi
    lib = _cabi.load()
    n = 3_000_000
    x = torch.randn(n, device=U.DEV).to(torch.bfloat16)
    g = torch.randn(n, device=U.DEV).to(torch.bfloat16)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    w = torch.randn(256, 1152, device=U.DEV) * 0.05
    gw = torch.randn(256, 1152, device=U.DEV)
    yw, gxw = torch.empty_like(w), torch.empty_like(w)
    s, b = torch.tensor([0.03], device=U.DEV), torch.tensor([-1.0], device=U.DEV)
    sw, bw = torch.full((256,), 0.002, device=U.DEV), torch.zeros(256, device=U.DEV)
    grads = torch.zeros(2 + 512, device=U.DEV)
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=U.DEV)
    qa, qw = U.qa(), U.qa(-128, 127, -128, 127, sym=True)

    def step(stream):
        sp = stream.cuda_stream
        assert lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, qa, sp) == 0
        assert lib.lsqb200_fwd_channel(w.data_ptr(), yw.data_ptr(), sw.data_ptr(), bw.data_ptr(), 1, 256, 1152, 0, 0, qw, sp) == 0
        assert lib.lsqb200_bwd_channel(gw.data_ptr(), w.data_ptr(), gxw.data_ptr(), sw.data_ptr(), bw.data_ptr(), grads[2:258].data_ptr(),
                                       grads[258:].data_ptr(), 1, 256, 1152, 0, 0, qw, ws.data_ptr(), ws.numel(), sp) == 0
        assert lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), grads[0:1].data_ptr(),
                                      grads[1:2].data_ptr(), n, 2, 0, qa, ws.data_ptr(), ws.numel(), sp) == 0

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(side)                                         # warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step(torch.cuda.current_stream())
    for trial in range(3):
        x.copy_(torch.randn(n, device=U.DEV).to(torch.bfloat16) * (trial + 1))
        g.copy_(torch.randn(n, device=U.DEV).to(torch.bfloat16))
        s.fill_(0.02 * (trial + 1))
        graph.replay()
        torch.cuda.synchronize()
        assert U.same_bits(y, U.oracle_fwd(x, s, b, qa))
        ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, qa)
        assert U.same_bits(gx, ogx)
        U.assert_grads_close(grads[0:1], ogs, ms, 1e-5)
        U.assert_grads_close(grads[1:2], ogb, mb, 1e-5)
        assert U.same_bits(yw, U.oracle_fwd(w.reshape(-1), sw, bw, qw, 1, 256, 1152, True))
