"""Pins the oracle's integer-export functions (oracle/lsq_oracle.c: lsq_oracle_qparams / _quantize / _dequantize)
against tests/golden/ref_export.npz (made by tests/golden/make_export_golden.py from the reference module's own
calculate_qparams(), the reference CPU op's forward, and torch's CPU quantize_per_tensor / quantize_per_channel).
Integer work: the bar is bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import lsq_oracle as O


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "ref_export.npz")


CASES = [f"t{i}" for i in range(6)] + ["c", "a"]


def _geom(name, x):
    if name == "c":
        return dict(outer=1, C=x.shape[0], inner=x.shape[1], per_channel=True), (-128, 127)
    if name == "a":
        return dict(outer=x.shape[0], C=x.shape[1], inner=x.shape[2], per_channel=True), (0, 255)
    return dict(outer=1, C=1, inner=x.size, per_channel=False), (0, 255)


@pytest.mark.parametrize("name", CASES)
def test_qparams_match_reference_module(gold, name):
    tmin, tmax = (-128, 127) if name == "c" else (0, 255)
    s, zp = O.qparams(gold[f"qp_{name}/scale_in"], gold[f"qp_{name}/shift_in"], tmin, tmax)
    assert np.array_equal(s.view(np.uint32), gold[f"qp_{name}/scale"].view(np.uint32))
    assert np.array_equal(zp, gold[f"qp_{name}/zp"])


@pytest.mark.parametrize("name", CASES)
def test_torch_cpu_codes_bit_exact(gold, name):
    x = gold[f"tq_{name}/x"]
    geom, (tmin, tmax) = _geom(name, x)
    c = O.cfg(tmin, tmax, tmin, tmax)
    codes = O.quantize(x.reshape(-1), gold[f"qp_{name}/scale_in"], gold[f"qp_{name}/shift_in"], c, sem=O.SEM_TORCH_CPU, **geom)
    want = gold[f"tq_{name}/codes"].astype(np.int32).reshape(-1)
    assert np.array_equal(codes, want), np.flatnonzero(codes != want)[:8]


@pytest.mark.parametrize("name", CASES)
def test_lsq_codes_dequantize_to_the_reference_forward(gold, name):
    """sem 0: dequantize(quantize(x)) must equal the reference CPU op's fake-quant output bit for bit."""
    x = gold[f"tq_{name}/x"]
    geom, (tmin, tmax) = _geom(name, x)
    c = O.cfg(tmin, tmax, tmin, tmax, contract=O.CONTRACT_CPU)
    s_in, b_in = gold[f"qp_{name}/scale_in"], gold[f"qp_{name}/shift_in"]
    codes = O.quantize(x.reshape(-1), s_in, b_in, c, sem=O.SEM_LSQ, **geom)
    assert codes.min() >= tmin and codes.max() <= tmax
    y = O.dequantize(codes, x.reshape(-1), s_in, b_in, c, sem=O.SEM_LSQ, **geom)
    want = gold[f"lsq_{name}/y"].reshape(-1)
    assert np.array_equal(y.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(y.view(np.uint32), O.forward(x.reshape(-1), s_in, b_in, c, geom["outer"], geom["C"], geom["inner"],
                                                       geom["per_channel"]).view(np.uint32))


def test_torch_cuda_semantics_differs_from_cpu_only_near_ties(gold):
    """sem 1 (double quotient) and sem 2 (fp32 reciprocal product) may disagree by one code at most."""
    x = gold["tq_t0/x"]
    c = O.cfg(0, 255, 0, 255)
    a = O.quantize(x, gold["qp_t0/scale_in"], gold["qp_t0/shift_in"], c, sem=O.SEM_TORCH_CUDA)
    b = O.quantize(x, gold["qp_t0/scale_in"], gold["qp_t0/shift_in"], c, sem=O.SEM_TORCH_CPU)
    finite = np.isfinite(x)
    assert np.abs(a[finite] - b[finite]).max() <= 1
    assert (a[finite] != b[finite]).mean() < 1e-3
