"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
Sizes are tiny so the 10-50x slowdown stays in seconds; results are still checked against the oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gpu_util as U
from torchlsq.multi import LSQPlan, Site

def check(x, g, s, b, q, outer=1, C=1, inner=None, pc=False):
    y = U.fwd(x, s, b, q, outer, C, inner, pc)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, pc))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, pc)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, pc)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-5); U.assert_grads_close(gb, ogb, mb, 1e-5)

gen = torch.Generator().manual_seed(0)
for dt in (torch.float32, torch.bfloat16, torch.float16):
    for n in (5, 4099, 300_001):                       # scalar tail, single tile, split tiles + last-arriver reduce
        x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
        s = torch.tensor([0.03], device=U.DEV); b = torch.tensor([-1.7], device=U.DEV)
        for mode in (dict(), dict(init_mode=True), dict(eval_mode=True)):
            check(x, g, s, b, U.qa(**mode))
    for shape, axis in (((6, 4, 3, 3), 0), ((64, 3, 7, 7), 0), ((4, 32, 14, 14), 1), ((2, 16, 56, 56), 1), ((64, 40), 1), ((16, 64, 7, 7), 1), ((3, 5, 7), 2)):
        n = int(np.prod(shape)); outer = int(np.prod(shape[:axis])); C = shape[axis]; inner = n // (outer * C)
        x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
        s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV); b = (-torch.rand(C, generator=gen)).to(U.DEV)
        check(x, g, s, b, U.qa(), outer, C, inner, True)
xh = torch.randn(20_001, generator=gen).half().to(U.DEV); gh = torch.randn(20_001, generator=gen).half().to(U.DEV)
check(xh, gh, torch.tensor([0.03], device=U.DEV).half(), torch.tensor([-1.7], device=U.DEV).half(), U.qa(use_gs=False))
# float64 kernels (lsq_f64.cuh): per-tensor (peel, single tile, split tiles) and per-channel rows / strided rows
from oracle import lsq_oracle as O
from torchlsq import _cabi
lib = _cabi.load()
def check64(x, g, s, b, q, outer=1, C=1, inner=None, pc=False):
    inner = x.numel() // (outer * C) if inner is None else inner
    y, gx = torch.empty_like(x), torch.empty_like(x)
    n = C if pc else 1
    gs = torch.empty(n, dtype=torch.float64, device=U.DEV); gb = torch.empty(n, dtype=torch.float64, device=U.DEV)
    ws = U.workspace()
    if pc:
        assert lib.lsqb200_fwd_channel(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), outer, C, inner, 3, 3, q, U.stream()) == 0
        assert lib.lsqb200_bwd_channel(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                       outer, C, inner, 3, 3, q, ws.data_ptr(), ws.numel(), U.stream()) == 0
    else:
        assert lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), x.numel(), 3, 3, q, U.stream()) == 0
        assert lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                      x.numel(), 3, 3, q, ws.data_ptr(), ws.numel(), U.stream()) == 0
    cfg = U.ocfg(q)
    oy = O.forward(x.cpu().numpy().reshape(-1), s.cpu().numpy(), b.cpu().numpy(), cfg, outer, C, inner, pc)
    ogx, ogs, ogb = O.backward(g.cpu().numpy().reshape(-1), x.cpu().numpy().reshape(-1), s.cpu().numpy(), b.cpu().numpy(), cfg, outer, C, inner, pc)
    assert np.array_equal(y.cpu().numpy().reshape(-1), oy) and np.array_equal(gx.cpu().numpy().reshape(-1), ogx)
    assert np.allclose(gs.cpu().numpy(), ogs, rtol=1e-9, atol=1e-12)
for n in (5, 4099, 300_001):
    x = torch.randn(n, generator=gen, dtype=torch.float64).to(U.DEV); g = torch.randn(n, generator=gen, dtype=torch.float64).to(U.DEV)
    check64(x, g, torch.tensor([0.03], dtype=torch.float64, device=U.DEV), torch.tensor([-1.7], dtype=torch.float64, device=U.DEV), U.qa())
for shape, axis in (((64, 3, 7, 7), 0), ((4, 32, 14, 14), 1), ((3, 5, 7), 2)):
    n = int(np.prod(shape)); outer = int(np.prod(shape[:axis])); C = shape[axis]; inner = n // (outer * C)
    x = torch.randn(n, generator=gen, dtype=torch.float64).to(U.DEV); g = torch.randn(n, generator=gen, dtype=torch.float64).to(U.DEV)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen, dtype=torch.float64)).to(U.DEV); b = (-torch.rand(C, generator=gen, dtype=torch.float64)).to(U.DEV)
    check64(x, g, s, b, U.qa(), outer, C, inner, True)
sites = []
for shp in ((16, 3, 7, 7), (32, 16, 1, 1), (8, 8, 3, 3)):
    w = (torch.randn(*shp, generator=gen) * 0.05).to(U.DEV)
    sites.append(Site(x=w, y=torch.empty_like(w), grad=torch.randn(*shp, generator=gen).to(U.DEV), gx=torch.empty_like(w),
                      scale=torch.full((shp[0],), 0.002, device=U.DEV), shift=torch.zeros(shp[0], device=U.DEV),
                      gscale=torch.empty(shp[0], device=U.DEV), gshift=torch.empty(shp[0], device=U.DEV),
                      quant_min=-128, quant_max=127, type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True))
plan = LSQPlan(sites); plan.forward(); plan.backward(); out = plan.weight_init_stats()
torch.cuda.synchronize()
assert torch.isfinite(out).all()
# fused prologues (relu / add+relu / add): row-tiled (per-tensor, per-channel long rows, warp-group rows) and column-layout kernels
def check_pre(code, x, x2, g, s, b, q, outer=1, C=1, inner=None, pc=False):
    relu = code != _cabi.PRE_ADD
    y = U.fwd(x, s, b, q, outer, C, inner, pc, prologue=code, x2=x2)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, pc, relu=relu, x2=x2))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, pc, prologue=code, x2=x2)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, pc, relu=relu, x2=x2)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-5); U.assert_grads_close(gb, ogb, mb, 1e-5)
for code in (_cabi.PRE_RELU, _cabi.PRE_ADD_RELU, _cabi.PRE_ADD):
    for dt in (torch.float32, torch.bfloat16):
        for n in (5, 4099, 300_001):
            x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
            x2 = torch.randn(n, generator=gen).to(dt).to(U.DEV) if code != _cabi.PRE_RELU else None
            for mode in (dict(), dict(init_mode=True)):
                check_pre(code, x, x2, g, torch.tensor([0.03], device=U.DEV), torch.tensor([-1.7], device=U.DEV), U.qa(**mode))
        for shape, axis in (((64, 3, 7, 7), 0), ((4, 32, 14, 14), 1), ((2, 16, 56, 56), 1), ((16, 64, 7, 7), 1)):
            n = int(np.prod(shape)); outer = int(np.prod(shape[:axis])); C = shape[axis]; inner = n // (outer * C)
            x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
            x2 = torch.randn(n, generator=gen).to(dt).to(U.DEV) if code != _cabi.PRE_RELU else None
            s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV); b = (-torch.rand(C, generator=gen)).to(U.DEV)
            check_pre(code, x, x2, g, s, b, U.qa(), outer, C, inner, True)
# lean kernels against the general ones: flat per-tensor (incl. the backward twin behind the knob), weight rows with head / tail, row statistics
for spec in (b"flatkernels=2", b"flatkernels=0,rowkernels=0,rowstats=0", b""):
    assert lib.lsqb200_set_tuning(spec) == 0
    for dt in (torch.float32, torch.bfloat16):
        for n in (37, 4099 * 8, 300_001):
            x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
            check(x, g, torch.tensor([0.03], device=U.DEV), torch.tensor([-1.7], device=U.DEV), U.qa())
        for shape in ((64, 3, 7, 7), (33, 1, 5, 5), (48, 64, 3, 3), (7, 1)):
            n = int(np.prod(shape)); C = shape[0]
            x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
            s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV); b = (-torch.rand(C, generator=gen)).to(U.DEV)
            check(x, g, s, b, U.qa(), 1, C, n // C, True)
            so = torch.empty(C, device=U.DEV)
            assert lib.lsqb200_weight_init_stats(x.data_ptr(), so.data_ptr(), 1, C, n // C, _cabi.F32 if dt == torch.float32 else _cabi.BF16,
                                                 -128, 127, U.workspace().data_ptr(), U.workspace().numel(), U.stream()) == 0
            torch.cuda.synchronize()
# round 2: plan statistics in the row-entry form (default) and through the bulk-copy ring (mbarrier + cp.async.bulk), rows that are
# whole 32-byte units and rows that are not; plan_rebind (patch kernel) and the grouped op behind autograd
for spec in (b"rowstats=2", b"rowstats=4", b"rowstats=7"):
    assert lib.lsqb200_set_tuning(spec) == 0
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        ws_ = [(torch.randn(*shp, generator=gen) * 0.05).to(dt).to(U.DEV) for shp in ((16, 3, 7, 7), (40, 16, 1, 1), (24, 8, 3, 3), (9, 1024), (70, 64))]
        st_ = [Site(x=w, scale=torch.ones(w.shape[0], device=U.DEV), shift=torch.zeros(w.shape[0], device=U.DEV), quant_min=-128, quant_max=127,
                    type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True) for w in ws_]
        p_ = LSQPlan(st_)
        o_ = p_.weight_init_stats(); o2_ = p_.weight_init_stats(torch.empty_like(o_))
        torch.cuda.synchronize()
        assert torch.equal(o_, o2_) and torch.isfinite(o_).all()
        p_.close()
lib.lsqb200_set_tuning(b"")
from torchlsq.multi import LSQGroup
gw = [(torch.randn(*shp, generator=gen) * 0.05).to(U.DEV).requires_grad_(True) for shp in ((16, 3, 7, 7), (32, 16, 1, 1), (8, 8, 3, 3), (10, 64))]
gsc = [torch.full((w.shape[0],), 0.002, device=U.DEV, requires_grad=True) for w in gw]
gsh = [torch.zeros(w.shape[0], device=U.DEV, requires_grad=True) for w in gw]
grp = LSQGroup(gw, gsc, gsh, -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True)
from torchlsq.functional import lsq as _lsq
for _ in range(2):
    ys = grp()
    torch.autograd.backward(ys, [torch.randn(*w.shape, generator=gen).to(U.DEV) for w in gw])
    for w, a, b_, y in zip(gw, gsc, gsh, ys):
        assert torch.equal(y, _lsq(w.detach(), a.detach(), b_.detach(), -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True))
        w.grad = a.grad = b_.grad = None
torch.cuda.synchronize()
grp.close()
print("sanitize_smoke ok")
