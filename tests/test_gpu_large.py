"""Sizes past 2^31 elements and the largest BASELINE configs[4] site (2048 x 64 x 112 x 112 = 1.64 G elements):
64-bit indexing end to end.  The oracle cannot run these sizes in seconds, so parity is checked through
size-independent properties: windows of the output (start, across the 2^31 boundary, ragged tail) against the oracle
bit for bit, the straight-through mask everywhere, and linearity of the reductions
(grad of the whole == sum of the kernel's own grads over chunks, which the smaller tests pin to the oracle)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as U


def _windows(n, w=1 << 18):
    edge = 1 << 31
    out = [(0, w), (n - w, n)]
    if n > edge + w:
        out.append((edge - w // 2 - 3, edge + w // 2 + 5))
    return out


@pytest.mark.parametrize("n,dtype", [((1 << 31) + 12345, torch.bfloat16), (2048 * 64 * 112 * 112, torch.bfloat16),
                                     ((1 << 31) + 8, torch.float16)])
def test_per_tensor_beyond_int32(n, dtype):
    free, _ = torch.cuda.mem_get_info()
    if free < 5 * 2 * n:
        pytest.skip("not enough device memory")
    gen = torch.Generator(device=U.DEV).manual_seed(n % 1000)
    x = torch.empty(n, dtype=dtype, device=U.DEV).normal_(0, 1.5, generator=gen)
    g = torch.empty(n, dtype=dtype, device=U.DEV).normal_(0, 1.0, generator=gen)
    s = torch.tensor([0.03], device=U.DEV)
    b = torch.tensor([-1.7], device=U.DEV)
    q = U.qa(use_gs=False)
    y = U.fwd(x, s, b, q)
    gx, gs, gb = U.bwd(g, x, s, b, q)
    torch.cuda.synchronize()
    for lo, hi in _windows(n):
        xs, gsl = x[lo:hi].contiguous(), g[lo:hi].contiguous()
        assert U.same_bits(y[lo:hi].contiguous(), U.oracle_fwd(xs, s, b, q)), (lo, hi)
        ogx = U.oracle_bwd(gsl, xs, s, b, q)[0]
        assert U.same_bits(gx[lo:hi].contiguous(), ogx), (lo, hi)
    # straight-through mask at full size: grad_x is g or (signed) zero, and zero exactly where the forward saturates
    sat = (y == y.min()) | (y == y.max())
    assert bool(((gx == g) | (gx == 0)).all())
    assert float((gx[sat] != 0).float().mean()) < 0.2          # saturated values that sit exactly on a border keep their gradient
    del sat, y
    # linearity: per-chunk gradients from the same kernel (each pinned to the oracle at small sizes) add up to the whole
    tot_s, tot_b = 0.0, 0.0
    step = 1 << 28
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        _, cs, cb = U.bwd(g[lo:hi], x[lo:hi], s, b, q, want_gx=False)
        tot_s += float(cs.double())
        tot_b += float(cb.double())
    assert abs(float(gs) - tot_s) <= 5e-6 * abs(tot_s) + 0.05, (float(gs), tot_s)
    assert abs(float(gb) - tot_b) <= 5e-6 * abs(tot_b) + 0.05, (float(gb), tot_b)


def test_per_channel_beyond_int32():
    """axis-1 per-channel on (outer, C, inner) = (2740, 1024, 784) fp16 = 2.2 G elements: the 28 x 28 maps of config 4 at a batch
    that pushes element offsets past 2^31."""
    outer, C, inner = 2740, 1024, 784
    n = outer * C * inner
    assert n > (1 << 31)
    free, _ = torch.cuda.mem_get_info()
    if free < 5 * 2 * n:
        pytest.skip("not enough device memory")
    gen = torch.Generator(device=U.DEV).manual_seed(3)
    x = torch.empty(n, dtype=torch.float16, device=U.DEV).normal_(0, 1.0, generator=gen)
    g = torch.empty(n, dtype=torch.float16, device=U.DEV).normal_(0, 1.0, generator=gen)
    s = (0.02 + 0.02 * torch.rand(C, device=U.DEV, generator=gen))
    b = -torch.rand(C, device=U.DEV, generator=gen)
    q = U.qa(use_gs=False)
    y = U.fwd(x, s, b, q, outer, C, inner, True)
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    torch.cuda.synchronize()
    # whole images at the start, across the 2^31 boundary and at the end, against the oracle
    img = C * inner
    k = (1 << 31) // img
    for first in (0, k - 1, outer - 2):
        lo, hi = first * img, (first + 2) * img
        xs, gsl = x[lo:hi].contiguous(), g[lo:hi].contiguous()
        assert U.same_bits(y[lo:hi].contiguous(), U.oracle_fwd(xs, s, b, q, 2, C, inner, True)), first
        assert U.same_bits(gx[lo:hi].contiguous(), U.oracle_bwd(gsl, xs, s, b, q, 2, C, inner, True)[0]), first
    # linearity over batch chunks
    tot_s = torch.zeros(C, dtype=torch.float64, device=U.DEV)
    tot_b = torch.zeros(C, dtype=torch.float64, device=U.DEV)
    step = 548
    for o in range(0, outer, step):
        oo = min(step, outer - o)
        _, cs, cb = U.bwd(g[o * img:(o + oo) * img], x[o * img:(o + oo) * img], s, b, q, oo, C, inner, True, want_gx=False)
        tot_s += cs.double()
        tot_b += cb.double()
    assert torch.allclose(gs.double(), tot_s, rtol=3e-6, atol=1e-2), float((gs.double() - tot_s).abs().max())
    assert torch.allclose(gb.double(), tot_b, rtol=3e-6, atol=1e-2), float((gb.double() - tot_b).abs().max())
