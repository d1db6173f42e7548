"""Pins the CPU oracle (oracle/lsq_oracle.c) to the reference.

(1) against tests/golden/ref_cpu_ops.npz -- outputs of the reference's own CPU op
    (torch.ops.torchlsq.* of torchlsq 2.1 built by oracle/build_ref.py), incl. the SURVEY.md
    Appendix-B known-answer cases: forward y and grad_x bit-exact with contract=0 (the CPU build
    does not fuse); grad_scale / grad_shift within 1e-6 relative of the reference's fp32 at::sum.
(2) against oracle/_ref/libref_scalar.so -- the reference's scalar templates compiled as they
    lie -- per element, bit for bit (skipped when that library is not present).
"""
import numpy as np
import pytest

from conftest import geometry
from oracle import lsq_oracle as O


def _cfg(m, **over):
    kw = dict(quant_min=m["qmin"], quant_max=m["qmax"], type_min=m["tmin"], type_max=m["tmax"],
              use_grad_scaling=m["use_gs"], grad_scaler=m["gscaler"], sym=not m["affine"], eval_mode=m["eval_mode"],
              init_mode=m["init_mode"], contract=O.CONTRACT_CPU, numel_div_c=True)
    kw.update(over)
    return O.cfg(**kw)


def _geom(m):
    if m["per_channel"]:
        return geometry(m["shape"], m["axis"])
    return 1, 1, int(np.prod(m["shape"]))


def _same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1).view(np.uint32)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1).view(np.uint32)
    nan_a, nan_b = np.isnan(a.view(np.float32)), np.isnan(b.view(np.float32))
    return bool(np.array_equal(nan_a, nan_b) and np.array_equal(a[~nan_a], b[~nan_b]))


def test_golden_cases_present(golden_ops):
    assert len(golden_ops) >= 19
    for k in ("B_A", "B_B", "B_C", "B_D", "B_E", "B_F", "B_G", "B_H"):
        assert k in golden_ops


def test_forward_bit_exact_vs_reference_cpu(golden_ops):
    for name, c in golden_ops.items():
        m = c["meta"]
        outer, C, inner = _geom(m)
        y = O.forward(c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner, m["per_channel"])
        assert _same_bits(y, c["y"]), name


def test_grad_x_bit_exact_vs_reference_cpu(golden_ops):
    for name, c in golden_ops.items():
        m = c["meta"]
        outer, C, inner = _geom(m)
        gx, _, _ = O.backward(c["g"], c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner, m["per_channel"])
        assert _same_bits(gx, c["dx"]), name


def test_param_grads_vs_reference_cpu(golden_ops):
    """The reference sums fp32 terms with at::sum (fp32, order unpinned): agree to 1e-6 of the
    result unless the sum is cancellation-dominated, then to 1e-6 of sum|terms| ~ |gs|*N*|term|."""
    for name, c in golden_ops.items():
        m = c["meta"]
        if any(np.isnan(c["ds"])) or any(np.isnan(c["db"])):
            continue   # NaN / inf inputs: checked in the KAT test below
        outer, C, inner = _geom(m)
        _, gs, gb, a_s, a_b = O.backward(c["g"], c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner,
                                         m["per_channel"], with_abs=True)
        for mine, ref, mag, what in ((gs, c["ds"], a_s, "ds"), (gb, c["db"], a_b, "db")):
            ref = ref.astype(np.float64)
            # fp32 at::sum: a few ulp of the summed magnitude, plus the final rounding of the result
            tol = 1e-6 * np.abs(ref) + 4 * 2.0 ** -24 * mag + 1e-30
            assert np.all(np.abs(mine - ref) <= tol), (name, what, mine, ref)


def test_appendix_b_known_answers(golden_ops):
    """SURVEY.md Appendix B values, typed in independently of the npz."""
    kat = {
        "B_A": (7617.99560547, 81.0), "B_B": (8591.07617188, 85.0), "B_C": (5698.07568359, 0.0),
        "B_D": (0.0, 0.0), "B_E": (-17563.87890625, -139.5), "B_F": (268.76458740, 2.20295691),
        "B_G": (5246.07568359, 43.0),
    }
    for name, (ds, db) in kat.items():
        c = golden_ops[name]
        m = c["meta"]
        x, g = c["x"].copy(), c["g"].copy()
        finite = np.isfinite(x)
        # NaN / inf rows contribute NaN-free border terms in the reference (fmin/fmax drop NaN); keep them
        _, gs, gb = O.backward(g, x, c["scale"], c["shift"], _cfg(m), 1, 1, x.size, False)
        assert abs(gs[0] - ds) <= 1e-6 * max(1.0, abs(ds)), (name, gs[0], ds)
        assert abs(gb[0] - db) <= 1e-6 * max(1.0, abs(db)), (name, gb[0], db)
        assert finite.sum() >= 12
    # the un-rounded strict mask (D4): x = 0.125 -> xq = 0.5 -> y = 0 (= qmin) but dx = g
    a = golden_ops["B_A"]
    assert a["y"][3] == 0.0 and a["dx"][3] == 4.0
    # ties to even: x/0.25 = 0.5, 1.5, 2.5 -> 0, 2, 2
    assert list(a["y"][3:6]) == [0.0, 0.5, 0.5]
    # NaN -> qmin in forward
    assert a["y"][12] == 0.0


def test_contract_modes_differ_only_in_last_bits(golden_ops):
    """contract=3 (reference CUDA build: fused v and d) vs contract=0: forward differs in at most
    a few elements by exactly one quantisation step."""
    c = golden_ops["R_tensor_affine"]
    m = c["meta"]
    y0 = O.forward(c["x"], c["scale"], c["shift"], _cfg(m), 1, 1, c["x"].size, False)
    y3 = O.forward(c["x"], c["scale"], c["shift"], _cfg(m, contract=O.CONTRACT_CUDA), 1, 1, c["x"].size, False)
    diff = np.abs(y0 - y3)
    assert (diff > 0).sum() <= 4
    assert np.all(diff <= float(c["scale"][0]) * 1.0001)


@pytest.mark.skipif(O.ref_scalar() is None, reason="oracle/_ref/libref_scalar.so not built (no /root/reference here)")
def test_oracle_matches_reference_scalar_templates_per_element():
    """Per-element dS_i*gs and dB_i*gs of the reference templates vs the oracle's terms:
    the oracle's double sums must equal the double sum of the reference's fp32 terms exactly."""
    import ctypes
    R = O.ref_scalar()
    rng = np.random.default_rng(5)
    n = 20011
    x = (rng.standard_normal(n) * 1.5).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    for (s, b, qmin, qmax, tmin, tmax, sym, init) in ((0.03, -1.7, 0, 127, 0, 255, 0, 0), (0.02, 0.0, -64, 63, -128, 127, 1, 0),
                                                       (0.03, -1.7, 0, 127, 0, 255, 0, 1), (-0.011, 0.4, 0, 255, 0, 255, 0, 0)):
        gsf = np.float32(1.0 / np.sqrt(float(n) * qmax))
        y = np.empty_like(x); dx = np.empty_like(x); ds = np.empty_like(x); db = np.empty_like(x)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        R.ref_fwd_tensor_f32(p(x), p(y), n, s, b, qmin, qmax, tmin, tmax, init)
        R.ref_bwd_tensor_f32(p(g), p(x), p(dx), p(ds), p(db), n, s, b, qmin, qmax, tmin, tmax, float(gsf), sym, init)
        c = O.cfg(qmin, qmax, tmin, tmax, use_grad_scaling=True, sym=bool(sym), init_mode=bool(init), contract=O.CONTRACT_CPU)
        yo = O.forward(x, [s], [b], c, 1, 1, n, False)
        gxo, gso, gbo = O.backward(g, x, [s], [b], c, 1, 1, n, False)
        assert _same_bits(yo, y) and _same_bits(gxo, dx)
        assert gso[0] == ds.astype(np.float64).sum() or abs(gso[0] - ds.astype(np.float64).sum()) <= 1e-12 * np.abs(ds).sum()
        assert abs(gbo[0] - db.astype(np.float64).sum()) <= 1e-12 * max(np.abs(db).sum(), 1e-30)
    # per-channel wrapper (eps clamp + per-element 1/s)
    outer, C, inner = 3, 5, 7
    x = rng.standard_normal(outer * C * inner).astype(np.float32)
    g = rng.standard_normal(x.size).astype(np.float32)
    sc = np.array([0.02, 0.5, 1e-9, -0.03, 0.04], np.float32)
    sh = np.array([0.0, -0.6, 0.0, 0.2, -1.0], np.float32)
    y = np.empty_like(x)
    R.ref_fwd_channel_f32(p(x), p(y), outer, C, inner, p(sc), p(sh), 0, 127, 0, 255, 0)
    yo = O.forward(x, sc, sh, O.cfg(0, 127, 0, 255, contract=O.CONTRACT_CPU), outer, C, inner, True)
    assert _same_bits(yo, y)


def test_16bit_conversions_match_numpy():
    rng = np.random.default_rng(3)
    f = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 10 ** rng.uniform(-8, 5, 20000).astype(np.float32),
                        np.array([0, -0.0, np.inf, -np.inf, 65504, 65519.99, 65520, 6e-8, 2.98e-8, 2.9802322e-8, 1e-10],
                                 np.float32)])
    lib = O.lib()
    mine = np.array([lib.lsq_oracle_f2h(float(v)) for v in f], np.uint16)
    ref = f.astype(np.float16).view(np.uint16)
    assert np.array_equal(mine, ref)
    back = np.array([lib.lsq_oracle_h2f(int(v)) for v in ref[:4000]], np.float32)
    assert np.array_equal(back.view(np.uint32), ref[:4000].view(np.float16).astype(np.float32).view(np.uint32))
    import torch
    tb = torch.from_numpy(f).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    mine_b = np.array([lib.lsq_oracle_f2bf(float(v)) for v in f], np.uint16)
    assert np.array_equal(mine_b, tb)


# ---- float64 tensors: oracle (contract=0, all double) vs the reference CPU op on float64 ------------------------
@pytest.fixture(scope="module")
def golden_f64():
    import json
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "ref_cpu_ops_f64.npz")
    meta = json.loads(str(z["meta"]))
    return {name: dict(meta=m, **{k: z[f"{name}/{k}"] for k in ("x", "g", "scale", "shift", "y", "dx", "ds", "db")})
            for name, m in meta.items()}


def _same_bits64(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64)))


def test_f64_forward_and_grad_x_bit_exact_vs_reference_cpu(golden_f64):
    assert len(golden_f64) >= 19
    for name, c in golden_f64.items():
        m = c["meta"]
        outer, C, inner = _geom(m)
        y = O.forward(c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner, m["per_channel"])
        assert y.dtype == np.float64 and _same_bits64(y, c["y"]), name
        gx, _, _ = O.backward(c["g"], c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner, m["per_channel"])
        assert _same_bits64(gx, c["dx"]), name


def test_f64_param_grads_vs_reference_cpu(golden_f64):
    """at::sum in double (order unpinned) vs the long-double sum of the same double terms: 1e-13 of sum|terms|."""
    for name, c in golden_f64.items():
        m = c["meta"]
        if np.isnan(c["ds"]).any() or np.isnan(c["db"]).any():
            continue
        outer, C, inner = _geom(m)
        _, gs, gb, a_s, a_b = O.backward(c["g"], c["x"], c["scale"], c["shift"], _cfg(m), outer, C, inner,
                                         m["per_channel"], with_abs=True)
        for mine, ref, mag, what in ((gs, c["ds"], a_s, "ds"), (gb, c["db"], a_b, "db")):
            assert np.all(np.abs(mine - ref) <= 1e-13 * mag + 1e-300), (name, what, mine, ref)


def test_f64_cuda_contract_rounds_clamps_through_float(golden_f64):
    """The CUDA build's float64 arithmetic (::fminf / ::fmaxf on doubles, global_scope.h:51-52): same integers
    as the CPU build except within a float ulp of a rounding tie, per-channel scales rounded to float."""
    c = golden_f64["R_tensor_affine"]
    m = c["meta"]
    y0 = O.forward(c["x"], c["scale"], c["shift"], _cfg(m), 1, 1, c["x"].size, False)
    y3 = O.forward(c["x"], c["scale"], c["shift"], _cfg(m, contract=O.CONTRACT_CUDA), 1, 1, c["x"].size, False)
    step = abs(float(c["scale"][0]))
    d = np.abs(y0 - y3)
    assert np.all((d == 0) | (np.abs(d - step) < 1e-12)) and (d != 0).mean() < 1e-3
    # per-channel: the CUDA build quantises with float(scale)
    c = golden_f64["R_channel_axis1"]
    m = c["meta"]
    outer, C, inner = _geom(m)
    s32 = c["scale"].astype(np.float32).astype(np.float64)
    assert np.any(s32 != c["scale"])
    y_cuda = O.forward(c["x"], c["scale"], c["shift"], _cfg(m, contract=O.CONTRACT_CUDA), outer, C, inner, True)
    y_cpu_f32scale = O.forward(c["x"], s32, c["shift"], _cfg(m), outer, C, inner, True)
    assert np.mean(y_cuda == y_cpu_f32scale) > 0.99
