"""Pins the oracle's restatement of the fused ReLU prologue (oracle.forward_relu / backward_relu) to the reference's own call
sequence: tests/golden/ref_cpu_relu.npz holds torch.relu -> reference CPU op (torchlsq 2.1) -> autograd through both, generated
by tests/golden/make_golden_relu.py.  Forward and grad_x bit-exact (contract 0 == the reference CPU build), parameter sums
within the error of the reference's fp32 at::sum."""
import numpy as np
import pytest

from conftest import GOLDEN, geometry
from oracle import lsq_oracle as O


@pytest.fixture(scope="module")
def relu_cases():
    data = np.load(GOLDEN / "ref_cpu_relu.npz")
    names = sorted({k.split("/")[0] for k in data.files})
    out = {}
    for n in names:
        c = {k: data[f"{n}/{k}"] for k in ("x", "g", "scale", "shift", "y", "dx", "ds", "db", "meta", "gscaler")}
        if f"{n}/x2" in data.files:
            c.update(x2=data[f"{n}/x2"], dx2=data[f"{n}/dx2"], with_relu=bool(data[f"{n}/relu"][0]))
        qmin, qmax, tmin, tmax, axis, use_gs, affine, per_channel, eval_mode, init_mode = (int(v) for v in c["meta"])
        c["m"] = dict(qmin=qmin, qmax=qmax, tmin=tmin, tmax=tmax, axis=axis, use_gs=bool(use_gs), affine=bool(affine),
                      per_channel=bool(per_channel), eval_mode=bool(eval_mode), init_mode=bool(init_mode),
                      gscaler=float(c["gscaler"][0]), shape=list(c["x"].shape))
        out[n] = c
    return out


def _cfg(m):
    return O.cfg(quant_min=m["qmin"], quant_max=m["qmax"], type_min=m["tmin"], type_max=m["tmax"], use_grad_scaling=m["use_gs"],
                 grad_scaler=m["gscaler"], sym=not m["affine"], eval_mode=m["eval_mode"], init_mode=m["init_mode"],
                 contract=O.CONTRACT_CPU, numel_div_c=True)


def _geom(m):
    return geometry(m["shape"], m["axis"]) if m["per_channel"] else (1, 1, int(np.prod(m["shape"])))


def _same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]))


def _fwd(c, *a, **k):
    if "x2" in c:
        return O.forward_add(c["x"].reshape(-1), c["x2"].reshape(-1), *a, with_relu=c["with_relu"], **k)
    return O.forward_relu(c["x"].reshape(-1), *a, **k)


def _bwd(c, *a, **k):
    if "x2" in c:
        return O.backward_add(c["g"].reshape(-1), c["x"].reshape(-1), c["x2"].reshape(-1), *a, with_relu=c["with_relu"], **k)
    return O.backward_relu(c["g"].reshape(-1), c["x"].reshape(-1), *a, **k)


def test_relu_golden_present(relu_cases):
    assert len(relu_cases) >= 25 and sum("x2" in c for c in relu_cases.values()) >= 9
    assert {"K_zp0", "K_shift_pos", "K_init_nan", "R_channel_axis1"} <= set(relu_cases)


def test_forward_relu_bit_exact_vs_reference_sequence(relu_cases):
    for name, c in relu_cases.items():
        outer, C, inner = _geom(c["m"])
        y = _fwd(c, c["scale"], c["shift"], _cfg(c["m"]), outer, C, inner, c["m"]["per_channel"])
        assert _same_bits(y, c["y"]), name


def test_grad_x_relu_bit_exact_vs_reference_sequence(relu_cases):
    for name, c in relu_cases.items():
        outer, C, inner = _geom(c["m"])
        gx, _, _ = _bwd(c, c["scale"], c["shift"], _cfg(c["m"]), outer, C, inner, c["m"]["per_channel"])
        assert _same_bits(gx, c["dx"]), name
        if "x2" in c:
            assert _same_bits(gx, c["dx2"]), name      # add's backward hands the same gradient to both addends


def test_param_grads_relu_vs_reference_sequence(relu_cases):
    for name, c in relu_cases.items():
        if np.isnan(c["ds"]).any() or np.isnan(c["db"]).any():
            continue
        outer, C, inner = _geom(c["m"])
        _, gs, gb, a_s, a_b = _bwd(c, c["scale"], c["shift"], _cfg(c["m"]), outer, C, inner, c["m"]["per_channel"], with_abs=True)
        for mine, ref, mag, what in ((gs, c["ds"], a_s, "ds"), (gb, c["db"], a_b, "db")):
            ref = ref.astype(np.float64)
            tol = 1e-6 * np.abs(ref) + 4 * 2.0 ** -24 * mag + 1e-30
            assert np.all(np.abs(mine - ref) <= tol), (name, what, mine, ref)


def test_relu_helper_on_16bit_patterns():
    """relu on bit patterns == torch.relu (ATen CPU kernel: -0 stays -0) on the values, for every fp16 and bf16 pattern; the
    CUDA contract differs in exactly one pattern, -0 -> +0."""
    import torch
    bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    for dt in (O.F16, O.BF16):
        vals = O.from_bits(bits, dt)
        want, _ = O.to_bits(torch.relu(vals))
        got = O.relu(bits, dt, keep_neg_zero=True)
        nan = torch.isnan(vals).numpy()
        assert np.array_equal(got[~nan], want[~nan])
        assert np.array_equal(got[nan], bits[nan])
        cuda = O.relu(bits, dt)
        diff = np.nonzero(cuda != got)[0]
        assert list(bits[diff]) == [0x8000] and cuda[0x8000] == 0
