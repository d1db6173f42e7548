"""GPU parity: the sm_100a kernels, called through the C ABI (include/lsq_b200.h), against the
CPU oracle on the same seeded bits.

Bars (BASELINE.json north_star): forward and grad_x BIT-EXACT (fp32 vs the reference CUDA
arithmetic = oracle contract 3; fp16 same-dtype vs c10::Half semantics; bf16 / mixed = fp32
internal + one rounding); grad_scale / grad_shift within 1e-6 relative for fp32 and 1e-5 for
fp16 / bf16 tensors with fp32 parameters (computed from fp32 terms, accumulated in fp64), plus
the output type's own rounding when the parameters are 16-bit.
"""
import numpy as np
import pytest
import torch

import gpu_util as U
from conftest import geometry
from oracle import lsq_oracle as O

pytestmark = pytest.mark.gpu

DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def _mk(n, dtype, seed, scale=1.5):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, generator=gen) * scale).to(dtype)
    g = torch.randn(n, generator=gen).to(dtype)
    return x.to(U.DEV), g.to(U.DEV)


def _params(vals_s, vals_b, dtype=torch.float32):
    return (torch.tensor(vals_s, dtype=dtype, device=U.DEV).reshape(-1),
            torch.tensor(vals_b, dtype=dtype, device=U.DEV).reshape(-1))


# ------------------------------------------------------------------------------------------------
# per-tensor
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 255, 1024, 4099, (1 << 20) + 3])
def test_fwd_tensor_bit_exact(dt, n):
    x, _ = _mk(n, DT[dt], seed=n)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    y = U.fwd(x, s, b, q)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q))


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 1024, 4099, (1 << 20) + 3, 6_422_528])
def test_bwd_tensor(dt, n):
    x, g = _mk(n, DT[dt], seed=100 + n)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    gx, gs, gb = U.bwd(g, x, s, b, q)
    ogx, ogs, ogb, mag_s, mag_b = U.oracle_bwd(g, x, s, b, q)
    assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == "f32" else 1e-5
    U.assert_grads_close(gs, ogs, mag_s, rel, "gscale")
    U.assert_grads_close(gb, ogb, mag_b, rel, "gshift")


def test_config1_fp32_grads_1e6_and_deterministic():
    """BASELINE config 1 exactly: randn(32,64,56,56) seed 1, g seed 2, s=0.03, b=-1.7, q=[0,127], t=[0,255]."""
    x = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(1)).to(U.DEV)
    g = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(2)).to(U.DEV)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    y = U.fwd(x, s, b, q)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q))
    gx, gs, gb = U.bwd(g, x, s, b, q)
    ogx, ogs, ogb, mag_s, mag_b = U.oracle_bwd(g, x, s, b, q)
    assert U.same_bits(gx, ogx)
    # plain 1e-6 relative, no cancellation allowance (sum|terms| is ~370x the result here)
    assert abs(gs.item() - ogs[0]) <= 1e-6 * abs(ogs[0]), (gs.item(), ogs[0])
    assert abs(gb.item() - ogb[0]) <= 1e-6 * abs(ogb[0]), (gb.item(), ogb[0])
    for _ in range(3):   # fixed-order cross-CTA reduction: bit-reproducible
        _, gs2, gb2 = U.bwd(g, x, s, b, q)
        assert torch.equal(gs2, gs) and torch.equal(gb2, gb)


def test_adversarial_values_fp32():
    """every half-integer pre-image (ties to even), both clamps, NaN, +-inf, -0.0, denormals."""
    s_, zp = 0.125, 17.0
    ties = (torch.arange(-3, 131, dtype=torch.float32) + 0.5 - zp) * s_
    special = torch.tensor([float("nan"), float("inf"), -float("inf"), -0.0, 0.0, 1e-40, -1e-40, 3.4e38, -3.4e38,
                            (0 - zp) * s_, (127 - zp) * s_, (127 - zp) * s_ * (1 + 2 ** -23), (0 - zp) * s_ * (1 - 2 ** -23)])
    x = torch.cat([ties, special]).to(U.DEV)
    g = torch.linspace(-2, 2, x.numel()).to(U.DEV)
    g[5] = float("inf"); g[9] = float("nan"); g[11] = -0.0
    s, b = _params([s_], [-zp * s_])
    for mode in (dict(), dict(init_mode=True), dict(eval_mode=True), dict(sym=True)):
        q = U.qa(use_gs=False, **mode)
        assert U.same_bits(U.fwd(x, s, b, q), U.oracle_fwd(x, s, b, q)), mode
        gx, gs, gb = U.bwd(g, x, s, b, q)
        ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
        assert U.same_bits(gx, ogx), mode
        for mine, ref in ((gs, ogs), (gb, ogb)):
            m = mine.double().cpu().numpy()
            assert np.array_equal(np.isnan(m), np.isnan(ref)), (mode, m, ref)
            ok = ~np.isnan(ref)
            assert np.allclose(m[ok], ref[ok], rtol=1e-6, atol=0), (mode, m, ref)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("mode", [dict(init_mode=True), dict(eval_mode=True), dict(eval_mode=True, init_mode=True),
                                  dict(sym=True, qmin=-64, qmax=63, tmin=-128, tmax=127),
                                  dict(use_gs=False), dict(gscaler=2.5), dict(qmin=0, qmax=255)])
def test_modes_tensor(dt, mode):
    n = 70_001
    x, g = _mk(n, DT[dt], seed=7)
    s, b = _params([0.021], [0.0 if mode.get("sym") else -0.9])
    q = U.qa(**mode)
    assert U.same_bits(U.fwd(x, s, b, q), U.oracle_fwd(x, s, b, q))
    gx, gs, gb = U.bwd(g, x, s, b, q)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
    assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == "f32" else 1e-5
    U.assert_grads_close(gs, ogs, ms, rel, "gscale")
    U.assert_grads_close(gb, ogb, mb, rel, "gshift")
    if mode.get("eval_mode") or mode.get("sym"):
        assert gb.item() == 0.0
    if mode.get("eval_mode"):
        assert gs.item() == 0.0


@pytest.mark.parametrize("scale,shift", [(-0.25, -0.6), (1e-12, 0.0), (0.0, 0.3), (3.0e4, 5.0), (0.05, 40.0), (0.05, -40.0)])
def test_odd_parameters_fp32(scale, shift):
    x, g = _mk(5000, torch.float32, seed=3)
    s, b = _params([scale], [shift])
    q = U.qa(use_gs=False)
    assert U.same_bits(U.fwd(x, s, b, q), U.oracle_fwd(x, s, b, q))
    gx, gs, gb = U.bwd(g, x, s, b, q)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-6)
    U.assert_grads_close(gb, ogb, mb, 1e-6)


def test_misaligned_base_pointers():
    """slices that start 1 element into an allocation fall to the scalar path; same results."""
    for dt in ("f32", "bf16"):
        n = 10_007
        x0, g0 = _mk(n + 1, DT[dt], seed=11)
        x, g = x0[1:], g0[1:]
        assert x.data_ptr() % 16 != 0
        s, b = _params([0.03], [-1.7])
        q = U.qa()
        assert U.same_bits(U.fwd(x, s, b, q), U.oracle_fwd(x, s, b, q))
        gx, gs, gb = U.bwd(g, x, s, b, q)
        ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
        assert U.same_bits(gx, ogx)
        U.assert_grads_close(gs, ogs, ms, 1e-5)


def test_fp16_same_dtype_params_half_exact():
    """x, scale, shift all fp16: c10::Half rounds after every operator."""
    n = 200_003
    x, g = _mk(n, torch.float16, seed=21, scale=1.0)
    for sv, bv in ((0.03, -1.7), (0.0171, -0.333), (0.25, 0.0)):
        s, b = _params([sv], [bv], dtype=torch.float16)
        q = U.qa(use_gs=False)
        y = U.fwd(x, s, b, q)
        assert U.same_bits(y, U.oracle_fwd(x, s, b, q)), (sv, bv)
        gx, gs, gb = U.bwd(g, x, s, b, q)
        ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
        assert U.same_bits(gx, ogx)
        assert gs.dtype == torch.float16
        U.assert_grads_close(gs, ogs, ms, 1e-5)
        U.assert_grads_close(gb, ogb, mb, 1e-5)


def test_bwd_without_gx():
    x, g = _mk(100_000, torch.float32, seed=5)
    s, b = _params([0.03], [-1.7])
    q = U.qa()
    _, gs1, gb1 = U.bwd(g, x, s, b, q, want_gx=True)
    _, gs2, gb2 = U.bwd(g, x, s, b, q, want_gx=False)
    assert torch.equal(gs1, gs2) and torch.equal(gb1, gb2)


def test_empty_and_errors():
    from torchlsq import _cabi
    lib = _cabi.load()
    s, b = _params([0.03], [-1.7])
    x = torch.empty(0, device=U.DEV)
    q = U.qa()
    assert U.fwd(x, s, b, q).numel() == 0
    _, gs, gb = U.bwd(x, x, s, b, q)
    assert gs.item() == 0.0 and gb.item() == 0.0
    # unsupported dtype pair: fp32 x with fp16 params
    sh, bh = _params([0.03], [-1.7], dtype=torch.float16)
    x1 = torch.zeros(8, device=U.DEV)
    rc = lib.lsqb200_fwd_tensor(x1.data_ptr(), x1.data_ptr(), sh.data_ptr(), bh.data_ptr(), 8, 0, 1, q, None)
    assert rc == -2 and b"dtype" in lib.lsqb200_last_error()
    rc = lib.lsqb200_fwd_tensor(None, x1.data_ptr(), s.data_ptr(), b.data_ptr(), 8, 0, 0, q, None)
    assert rc == -1
    rc = lib.lsqb200_bwd_tensor(x1.data_ptr(), x1.data_ptr(), None, s.data_ptr(), b.data_ptr(), s.data_ptr(), b.data_ptr(),
                                1 << 22, 0, 0, q, None, 0, None)
    assert rc == -3   # a split launch needs the workspace


# ------------------------------------------------------------------------------------------------
# per-channel
# ------------------------------------------------------------------------------------------------
CH_SHAPES = [
    ((6, 4, 3, 3), 0), ((64, 3, 7, 7), 0),          # weight rows, K = 36 / 147 (rows lose 16 B alignment)
    ((256, 64, 1, 1), 0), ((1000, 2048), 0), ((512, 512, 3, 3), 0),
    ((33, 1, 5, 5), 0), ((10, 37), 0), ((7, 1), 0), ((5, 8200), 0),   # rows shorter than a unit, single-element rows, rows past the warp-group limit
    ((8, 32, 14, 14), 1), ((4, 64, 56, 56), 1), ((16, 1024, 28, 28), 1), ((3, 5, 7), 1), ((3, 5, 7), 2),
    ((32, 1000), 1), ((2, 3, 224, 224), 1), ((64, 256, 7, 7), 1),
    # short channel rows -> column-layout kernels: channels-last, 7x7 / 14x14 maps, ragged unit/channel overlap
    ((6272, 1024), 1), ((300, 40), 1), ((64, 2048, 7, 7), 1), ((33, 24, 14, 14), 1), ((5, 12, 2, 3), 1), ((1031, 8), 1),
]


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape,axis", CH_SHAPES)
def test_channel_fwd_bwd(dt, shape, axis):
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=n % 1000 + axis, scale=0.6)
    outer, C, inner = geometry(shape, axis)
    gen = torch.Generator().manual_seed(C)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV)
    q = U.qa()
    y = U.fwd(x, s, b, q, outer, C, inner, True)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, outer, C, inner, True))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == "f32" else 1e-5
    U.assert_grads_close(gs, ogs, ms, rel, "gscale")
    U.assert_grads_close(gb, ogb, mb, rel, "gshift")
    _, gs2, gb2 = U.bwd(g, x, s, b, q, outer, C, inner, True)
    assert torch.equal(gs, gs2) and torch.equal(gb, gb2)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("mode", ["init", "eval", "eval_init", "sym", "no_gx"])
def test_weight_rows_with_head_and_tail_all_modes(dt, mode):
    """Lean warp-per-row kernels on rows that neither start nor end on a 32-byte boundary (a first conv layer's 7x7x3 rows)."""
    shape = (64, 3, 7, 7)
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=77, scale=0.6)
    outer, C, inner = geometry(shape, 0)
    gen = torch.Generator().manual_seed(5)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV) * (0.0 if mode == "sym" else 1.0)
    kw = dict(init=dict(init_mode=True), eval=dict(eval_mode=True), eval_init=dict(eval_mode=True, init_mode=True),
              sym=dict(qmin=-64, qmax=63, tmin=-128, tmax=127, sym=True), no_gx=dict())[mode]
    q = U.qa(**kw)
    assert U.same_bits(U.fwd(x, s, b, q, outer, C, inner, True), U.oracle_fwd(x, s, b, q, outer, C, inner, True))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True, want_gx=(mode != "no_gx"))
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    if mode != "no_gx":
        assert U.same_bits(gx, ogx)
    rel = 1e-6 if dt == "f32" else 1e-5
    U.assert_grads_close(gs, ogs, ms, rel, "gscale")
    U.assert_grads_close(gb, ogb, mb, rel, "gshift")


def test_channel_weights_symmetric_qint8():
    """BASELINE config 2 flavour: conv weight, per-channel axis 0, symmetric [-128, 127]."""
    shape = (256, 128, 3, 3)
    n = int(np.prod(shape))
    x, g = _mk(n, torch.float32, seed=9, scale=0.05)
    outer, C, inner = geometry(shape, 0)
    s = (0.0005 + 0.001 * torch.rand(C, generator=torch.Generator().manual_seed(1))).to(U.DEV)
    b = torch.zeros(C, device=U.DEV)
    q = U.qa(qmin=-128, qmax=127, tmin=-128, tmax=127, sym=True)
    assert U.same_bits(U.fwd(x, s, b, q, outer, C, inner, True), U.oracle_fwd(x, s, b, q, outer, C, inner, True))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-6)
    assert torch.count_nonzero(gb).item() == 0


def test_channel_tiny_and_negative_scales():
    shape, axis = (4, 6, 10), 1
    outer, C, inner = geometry(shape, axis)
    x, g = _mk(int(np.prod(shape)), torch.float32, seed=2)
    s, b = _params([0.25, 0.5, 1e-9, -0.03, 0.0, 7.0], [0.0, -0.6, 0.0, 0.2, 0.1, -3.0])
    q = U.qa()
    assert U.same_bits(U.fwd(x, s, b, q, outer, C, inner, True), U.oracle_fwd(x, s, b, q, outer, C, inner, True))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-6)
    U.assert_grads_close(gb, ogb, mb, 1e-6)


@pytest.mark.parametrize("mode", [dict(init_mode=True), dict(eval_mode=True), dict(use_gs=False)])
def test_channel_modes(mode):
    shape, axis = (8, 48, 14, 14), 1
    outer, C, inner = geometry(shape, axis)
    x, g = _mk(int(np.prod(shape)), torch.bfloat16, seed=4)
    s = torch.full((C,), 0.04, device=U.DEV)
    b = torch.full((C,), -1.0, device=U.DEV)
    q = U.qa(**mode)
    assert U.same_bits(U.fwd(x, s, b, q, outer, C, inner, True), U.oracle_fwd(x, s, b, q, outer, C, inner, True))
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-5)
    U.assert_grads_close(gb, ogb, mb, 1e-5)


# ------------------------------------------------------------------------------------------------
# golden fixtures from the reference CPU op (different contraction: at most a handful of elements
# may sit one quantisation step away) and mu +- 3 sigma init
# ------------------------------------------------------------------------------------------------
def test_against_reference_cpu_golden(golden_ops):
    for name, c in golden_ops.items():
        m = c["meta"]
        x = torch.from_numpy(c["x"]).to(U.DEV).reshape(-1)
        g = torch.from_numpy(c["g"]).to(U.DEV).reshape(-1)
        s = torch.from_numpy(c["scale"]).to(U.DEV)
        b = torch.from_numpy(c["shift"]).to(U.DEV)
        q = U.qa(m["qmin"], m["qmax"], m["tmin"], m["tmax"], m["use_gs"], m["gscaler"], not m["affine"], m["eval_mode"],
                 m["init_mode"])
        if m["per_channel"]:
            outer, C, inner = geometry(m["shape"], m["axis"])
        else:
            outer, C, inner = 1, 1, x.numel()
        y = U.fwd(x, s, b, q, outer, C, inner, m["per_channel"])
        ndiff = U.count_diff(y, c["y"])
        assert ndiff <= max(2, x.numel() // 500), (name, ndiff)
        step = float(np.abs(c["scale"]).max()) * 1.001 + 1e-30
        yy, ry = y.cpu().numpy().reshape(-1), c["y"].reshape(-1)
        fin = np.isfinite(ry)
        assert np.all(np.abs(yy[fin] - ry[fin]) <= step), name
        gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, m["per_channel"])
        assert U.count_diff(gx, c["dx"]) <= max(2, x.numel() // 500), name
        if not np.isnan(c["ds"]).any():
            ref_s = c["ds"].astype(np.float64)
            ref_b = c["db"].astype(np.float64)
            if m["per_channel"] and m["use_gs"]:
                ref_s = ref_s / np.sqrt(C)      # D5: the reference CPU op divides numel by C; CUDA (and we) do not
                ref_b = ref_b / np.sqrt(C)
            assert np.allclose(gs.double().cpu().numpy(), ref_s, rtol=2e-4, atol=2e-6 * np.abs(ref_s).max() + 1e-12), name
            assert np.allclose(gb.double().cpu().numpy(), ref_b, rtol=2e-4, atol=2e-6 * np.abs(ref_b).max() + 1e-12), name


def test_weight_init_stats_vs_reference_module(golden_module, native_lib):
    from conftest import GOLDEN
    for tag, w in golden_module["winit"].items():
        arr = np.load(GOLDEN / f"ref_winit_{tag}.npz")[f"winit/{tag}"]
        wt = torch.from_numpy(arr).to(U.DEV)
        C = arr.shape[0] if w["per_channel"] else 1
        outer, Cc, inner = (1, C, arr.size // C)
        out = torch.empty(C, device=U.DEV)
        ws = U.workspace()
        rc = native_lib.lsqb200_weight_init_stats(wt.data_ptr(), out.data_ptr(), outer, Cc, inner, 0, w["quant_min"],
                                                  w["quant_max"], ws.data_ptr(), ws.numel(), U.stream())
        assert rc == 0
        ref = np.array(w["scale"], np.float64)
        assert np.allclose(out.double().cpu().numpy(), ref, rtol=2e-6), tag
        orc = O.weight_init(arr.reshape(-1), w["quant_min"], w["quant_max"], outer, Cc, inner)
        assert np.allclose(out.cpu().numpy(), orc, rtol=1e-6), tag


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("shape,axis", [((512, 512, 3, 3), 0), ((64, 3, 7, 7), 0), ((16, 96, 20, 20), 1), ((40, 1000), 1),
                                        ((3_000_017,), None)])
def test_weight_init_stats_vs_oracle(dt, shape, axis, native_lib):
    n = int(np.prod(shape))
    x, _ = _mk(n, DT[dt], seed=13, scale=0.07)
    x = x + 0.01
    if axis is None:
        outer, C, inner = 1, 1, n
    else:
        outer, C, inner = geometry(shape, axis)
    out = torch.empty(C, device=U.DEV)
    ws = U.workspace()
    rc = native_lib.lsqb200_weight_init_stats(x.data_ptr(), out.data_ptr(), outer, C, inner, {"f32": 0, "bf16": 2}[dt],
                                              -128, 127, ws.data_ptr(), ws.numel(), U.stream())
    assert rc == 0
    xb, code = O.to_bits(x)
    ref = O.weight_init(xb.reshape(-1), -128, 127, outer, C, inner, dt=code)
    assert np.allclose(out.cpu().numpy(), ref, rtol=2e-6), (out[:4], ref[:4])


# ------------------------------------------------------------------------------------------------
# size-independent properties at full BASELINE sizes
# ------------------------------------------------------------------------------------------------
def test_config4_properties_fp16_channel():
    """256x1024x28x28 fp16, per-channel axis 1, grad scaling on (BASELINE config 4):
    idempotence of the forward, linearity of the reductions in g, shard additivity, and a direct
    oracle comparison on a 2-image slab."""
    N, C, H, W = 256, 1024, 28, 28
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(N, C, H * W, generator=gen, dtype=torch.float32).to(torch.float16).to(U.DEV)
    g = torch.randn(N, C, H * W, generator=gen, dtype=torch.float32).to(torch.float16).to(U.DEV)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV)
    q = U.qa(use_gs=True)
    outer, inner = N, H * W
    y = U.fwd(x, s, b, q, outer, C, inner, True)
    y2 = U.fwd(y, s, b, q, outer, C, inner, True)
    assert torch.equal(y, y2)                                         # fake-quant is idempotent
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    _, gs2, gb2 = U.bwd(g * 2, x, s, b, q, outer, C, inner, True)
    assert torch.allclose(gs2, 2 * gs, rtol=1e-6, atol=0) and torch.allclose(gb2, 2 * gb, rtol=1e-6, atol=0)
    # mask semantics at full size: gx is g or 0
    assert torch.all((gx == g) | (gx == 0)).item()
    # shard additivity with use_grad_scaling off (gs depends on numel otherwise)
    q0 = U.qa(use_gs=False)
    _, fs, fb = U.bwd(g, x, s, b, q0, outer, C, inner, True)
    h = N // 2
    _, s1, b1 = U.bwd(g[:h], x[:h], s, b, q0, h, C, inner, True)
    _, s2, b2 = U.bwd(g[h:], x[h:], s, b, q0, N - h, C, inner, True)
    assert torch.allclose(s1 + s2, fs, rtol=2e-6, atol=1e-3) and torch.allclose(b1 + b2, fb, rtol=2e-6, atol=1e-3)
    # direct oracle comparison on a slab
    xs, gsl = x[:2].contiguous(), g[:2].contiguous()
    ys = U.fwd(xs, s, b, q, 2, C, inner, True)
    assert U.same_bits(ys, U.oracle_fwd(xs, s, b, q, 2, C, inner, True))
    assert torch.equal(ys, y[:2])
    gxs, gss, gbs = U.bwd(gsl, xs, s, b, q, 2, C, inner, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(gsl, xs, s, b, q, 2, C, inner, True)
    assert U.same_bits(gxs, ogx)
    U.assert_grads_close(gss, ogs, ms, 1e-5)
    U.assert_grads_close(gbs, ogb, mb, 1e-5)


def test_large_bf16_site_properties():
    """largest ResNet-50 activation site at batch 256 (256x64x112x112 bf16, 205 M elements):
    learned-init mode is a copy, the normal mode is idempotent and its gx is g*mask."""
    n = 256 * 64 * 112 * 112
    x = torch.empty(n, dtype=torch.bfloat16, device=U.DEV).normal_(0, 1, generator=torch.Generator(U.DEV).manual_seed(0)).relu_()
    g = torch.empty(n, dtype=torch.bfloat16, device=U.DEV).normal_(0, 1, generator=torch.Generator(U.DEV).manual_seed(1))
    s, b = _params([0.03], [0.0])
    qi = U.qa(init_mode=True)
    assert torch.equal(U.fwd(x, s, b, qi), x)
    gx, gs, gb = U.bwd(g, x, s, b, qi)
    assert torch.equal(gx, g)
    q = U.qa()
    y = U.fwd(x, s, b, q)
    assert torch.equal(U.fwd(y, s, b, q), y)
    gx, gs, gb = U.bwd(g, x, s, b, q)
    assert torch.all((gx == g) | (gx == 0)).item()
    m = 1 << 22
    assert U.same_bits(y[:m], U.oracle_fwd(x[:m], s, b, q))
    # against a float64 torch evaluation of the same reductions
    xf, gf = x[: 1 << 24].double(), g[: 1 << 24].double()
    _, gs_s, gb_s = U.bwd(g[: 1 << 24], x[: 1 << 24], s, b, U.qa(use_gs=False))
    v = xf / 0.03
    inside = (v > 0) & (v < 127)
    xfq = torch.clamp(v, 0, 127).round() * 0.03
    terms = torch.where(inside, gf * (xfq - xf) / 0.03, torch.where(v <= 0, gf * 0, gf * 127))
    assert abs(gs_s.item() - terms.sum().item()) <= 1e-4 * terms.abs().sum().item()


# ------------------------------------------------------------------------------------------------
# lean kernels (per-tensor single launches, weight rows) against the general kernels they shortcut
# ------------------------------------------------------------------------------------------------
def _with_tuning(lib, spec, fn):
    assert lib.lsqb200_set_tuning(spec.encode()) == 0
    try:
        return fn()
    finally:
        lib.lsqb200_set_tuning(b"")


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n", [300_001, (1 << 20) + 3, 6_422_528, 51_380_224])
@pytest.mark.parametrize("mode", ["normal", "init", "eval"])
def test_lean_per_tensor_kernels_bit_identical_to_general(dt, n, mode, native_lib):
    """lsq_flatfwd / lsq_flatbwd keep the general kernels' tiles, unit <-> thread mapping and fixed-order reduction: every
    output, the parameter gradients included, is the same bit pattern."""
    x, g = _mk(n, DT[dt], seed=n % 997)
    s, b = _params([0.03], [-1.7])
    q = U.qa(init_mode=(mode == "init"), eval_mode=(mode == "eval"))

    def run():
        y = U.fwd(x, s, b, q)
        gx, gs, gb = U.bwd(g, x, s, b, q)
        torch.cuda.synchronize()
        return y, gx, gs, gb
    lean = _with_tuning(native_lib, "flatkernels=2", run)
    general = _with_tuning(native_lib, "flatkernels=0", run)
    for a_, b_ in zip(lean, general):
        assert torch.equal(a_.view(torch.int32 if a_.dtype == torch.float32 else torch.int16),
                           b_.view(torch.int32 if b_.dtype == torch.float32 else torch.int16))
    assert U.same_bits(lean[0], U.oracle_fwd(x, s, b, q))


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(512, 512, 3, 3), (64, 3, 7, 7), (1000, 2048), (33, 1, 5, 5)])
def test_lean_weight_row_kernels_match_general(dt, shape, native_lib):
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=n % 991, scale=0.6)
    outer, C, inner = geometry(shape, 0)
    gen = torch.Generator().manual_seed(C)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV)
    q = U.qa()

    def run():
        y = U.fwd(x, s, b, q, outer, C, inner, True)
        gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
        torch.cuda.synchronize()
        return y, gx, gs, gb
    lean = _with_tuning(native_lib, "rowkernels=1", run)
    general = _with_tuning(native_lib, "rowkernels=0", run)
    assert torch.equal(lean[0], general[0]) and torch.equal(lean[1], general[1])
    assert torch.allclose(lean[2], general[2], rtol=1e-6, atol=1e-12) and torch.allclose(lean[3], general[3], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape_axis", [((64, 512, 7, 7), 1), ((6272, 1024), 1), ((32, 256, 14, 14), 1), ((4099, 96), 1), ((257, 24), 1),
                                        ((3, 40, 2, 2), 1), ((1031, 8, 3, 3), 1), ((70, 1000, 1, 1), 1), ((9, 333, 4, 4), 1), ((130, 2050, 3, 3), 1)])
def test_column_backward_epilogue_vs_float64_sums(dt, shape_axis, native_lib):
    """The column backward's epilogue (per-thread channel runs parked in shared memory, one thread per channel of the CTA adds them,
    one fp64 atomic pair per (CTA, channel); straight atomics when a unit holds whole channels) on layouts whose channels straddle
    threads, CTAs and partly filled last CTAs: grad_x bit for bit and grad_scale / grad_shift against the oracle's float64 sums, a
    second call on the same workspace (accumulators and ticket must come back zeroed) and the call without grad_x."""
    shape, axis = shape_axis
    n = int(np.prod(shape))
    x, g = _mk(n, DT[dt], seed=n % 983, scale=0.7)
    outer, C, inner = geometry(shape, axis)
    gen = torch.Generator().manual_seed(C)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV)
    q = U.qa(use_gs=False)
    gx, gs, gb = U.bwd(g, x, s, b, q, outer, C, inner, True)
    gx2, gs2, gb2 = U.bwd(g, x, s, b, q, outer, C, inner, True)
    assert torch.equal(gx, gx2) and torch.allclose(gs, gs2, rtol=1e-9, atol=0) and torch.allclose(gb, gb2, rtol=1e-9, atol=0)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-5)
    U.assert_grads_close(gb, ogb, mb, 1e-5)
    _, gs3, gb3 = U.bwd(g, x, s, b, q, outer, C, inner, True, want_gx=False)
    assert torch.allclose(gs3, gs, rtol=1e-9, atol=0) and torch.allclose(gb3, gb, rtol=1e-9, atol=0)


def test_config4_full_size_gradients_vs_oracle():
    """BASELINE config 4 at FULL size (256x1024x28x28 fp16, per-channel axis 1, grad scaling 1/sqrt(numel*qmax) with the whole
    tensor's numel, lsq_cuda.cu:274): forward and grad_x bit-exact and all 1024 grad_scale / grad_shift values against the
    oracle run over all 205 M elements (task 7b: no slab with its own gs)."""
    N, C, H, W = 256, 1024, 28, 28
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(N, C, H * W, generator=gen, dtype=torch.float32).to(torch.float16).to(U.DEV)
    g = torch.randn(N, C, H * W, generator=gen, dtype=torch.float32).to(torch.float16).to(U.DEV)
    s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV)
    b = (-torch.rand(C, generator=gen)).to(U.DEV)
    q = U.qa(use_gs=True)
    y = U.fwd(x, s, b, q, N, C, H * W, True)
    assert U.same_bits(y, U.oracle_fwd(x, s, b, q, N, C, H * W, True))
    del y
    gx, gs, gb = U.bwd(g, x, s, b, q, N, C, H * W, True)
    ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, N, C, H * W, True)
    assert U.same_bits(gx, ogx)
    U.assert_grads_close(gs, ogs, ms, 1e-6, "config4 full size gscale")
    U.assert_grads_close(gb, ogb, mb, 1e-6, "config4 full size gshift")


def test_config3_largest_site_full_size_vs_oracle():
    """Largest ResNet-50 activation site of BASELINE config 3 (256x64x112x112 bf16, 205 M elements), learned-init mode and the
    normal mode: every output element and both reductions against the oracle over the whole tensor."""
    n = 256 * 64 * 112 * 112
    x = torch.empty(n, dtype=torch.bfloat16, device=U.DEV).normal_(0, 1, generator=torch.Generator(U.DEV).manual_seed(0)).relu_()
    g = torch.empty(n, dtype=torch.bfloat16, device=U.DEV).normal_(0, 1, generator=torch.Generator(U.DEV).manual_seed(1))
    s, b = _params([0.03], [0.0])
    for q, what in ((U.qa(init_mode=True), "learned init"), (U.qa(), "normal")):
        y = U.fwd(x, s, b, q)
        assert U.same_bits(y, U.oracle_fwd(x, s, b, q)), what
        del y
        gx, gs, gb = U.bwd(g, x, s, b, q)
        ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q)
        assert U.same_bits(gx, ogx), what
        del gx, ogx
        U.assert_grads_close(gs, ogs, ms, 1e-6, f"config3 largest site gscale ({what})")
        U.assert_grads_close(gb, ogb, mb, 1e-6, f"config3 largest site gshift ({what})")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_back_to_back_producer_consumer_chain_under_pdl_and_prefetch(dtype):
    """Kernels are launched with programmatic dependent launch and L2-prefetch their first units BEFORE griddepcontrol.wait.  A chain
    in which every launch consumes what the previous one is still writing (forward -> forward on its output -> backward with that
    output as upstream gradient, ping-pong buffers, no host sync in between) must give bit for bit what the same chain gives with a
    device synchronisation after every launch: the prefetch is non-binding and L2 is the point of coherence."""
    n = 8 * 1024 * 1024 + 24
    gen = torch.Generator(device=U.DEV).manual_seed(21)
    x0 = torch.empty(n, device=U.DEV).normal_(0, 1, generator=gen).to(dtype)
    s, b = _params([0.05], [-1.3])
    q = U.qa()

    def chain(sync):
        out = []
        cur = x0
        for i in range(12):
            y = U.fwd(cur, s, b, q)
            if sync:
                torch.cuda.synchronize()
            gx, gs, gb = U.bwd(y, cur, s, b, q)          # upstream gradient = the tensor the previous launch just wrote
            if sync:
                torch.cuda.synchronize()
            out.append((gs.clone(), gb.clone()))
            cur = (gx * 0.5 + y).to(dtype) if i % 3 == 2 else y          # an ATen kernel in the chain now and then
        torch.cuda.synchronize()
        return cur, out

    a_cur, a_out = chain(False)
    b_cur, b_out = chain(True)
    assert U.same_bits(a_cur, O.to_bits(b_cur)[0])
    for (gs1, gb1), (gs2, gb2) in zip(a_out, b_out):
        assert torch.equal(gs1, gs2) and torch.equal(gb1, gb2)
