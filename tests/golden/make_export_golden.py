#!/usr/bin/env python
"""Golden fixtures for the integer-export step.  Build container only:

    python oracle/build_ref.py && python tests/golden/make_export_golden.py

Output: tests/golden/ref_export.npz
  * `qp_*`   : LSQFakeQuantizer.calculate_qparams() of the REFERENCE module (oracle/_ref, its own Python:
               torchlsq/quantized/modules/observers.py:378-422) for hand-set scale / shift values;
  * `tq_*`   : int_repr() of torch.quantize_per_tensor / quantize_per_channel (torch 2.11 CPU kernels - third-party
               to the reference, which hands its qparams to torch.quantization.convert) on those qparams;
  * `lsq_*`  : the reference CPU op's fake-quant output for the same inputs, from which the LSQ integer is
               recovered exactly as y / s + zp (pins oracle sem 0 against the reference's own forward).
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))   # the REFERENCE package, not ours

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchlsq  # noqa: E402  (reference)
from torchlsq import LSQFakeQuantizer  # noqa: E402
from torchlsq.functional import lsq  # noqa: E402

assert "oracle/_ref" in torchlsq.__file__, torchlsq.__file__
OUT = Path(__file__).resolve().parent


def ref_module(dtype, qscheme, scale, shift, otype):
    m = LSQFakeQuantizer(None, otype, dtype=dtype, qscheme=qscheme, init_mode='learnable', avoid_torch_overflow=False)
    per_ch = qscheme in (torch.per_channel_affine, torch.per_channel_symmetric)
    probe = torch.zeros(2, len(scale), 2) if (per_ch and otype == 'activation') else torch.zeros(len(scale), 4)
    m(probe)                                     # first call creates the parameters
    m._set_weights(scale=torch.tensor(scale), shift=torch.tensor(shift))
    return m


def main():
    out = {}
    gen = torch.Generator().manual_seed(77)
    nan, inf = float("nan"), float("inf")
    edge = torch.tensor([0.0, -0.0, 0.125, 0.375, 0.625, -0.125, -0.375, 31.75, 32.0, 1e9, -1e9, nan, inf, -inf, 1e-30, 63.5 * 0.03])

    # ---- per-tensor activations (quint8 affine) ----
    for i, (s, b) in enumerate([(0.03, -1.7), (0.25, -0.6), (1e-9, 0.0), (0.0173, 2.2), (0.5, -200.0), (-0.1, 0.3)]):
        m = ref_module(torch.quint8, torch.per_tensor_affine, [s], [b], 'activation')
        sc, sh, zp = m.calculate_qparams(need_shift=True)
        x = torch.cat([torch.randn(3000, generator=gen) * 2.0 + 1.0, edge])
        q = torch.quantize_per_tensor(x, float(sc[0]), int(zp[0]), torch.quint8)
        out[f"qp_t{i}/scale_in"] = np.float32([s]); out[f"qp_t{i}/shift_in"] = np.float32([b])
        out[f"qp_t{i}/scale"] = sc.numpy(); out[f"qp_t{i}/zp"] = zp.numpy()
        out[f"tq_t{i}/x"] = x.numpy(); out[f"tq_t{i}/codes"] = q.int_repr().numpy()
        y = lsq(x, torch.tensor([s]), torch.tensor([b]), 0, 255, 0, 255, 1, False, 1.0, True, False, True, False)
        out[f"lsq_t{i}/y"] = y.numpy()

    # ---- per-channel weights (qint8 symmetric) ----
    C = 7
    scales = [0.002, 0.004, 1e-10, 0.0101, 0.5, 0.0333, 0.25]
    shifts = [0.0] * C
    m = ref_module(torch.qint8, torch.per_channel_symmetric, scales, shifts, 'weight')
    sc, sh, zp = m.calculate_qparams(need_shift=True)
    w = torch.randn(C, 37, generator=gen) * 0.3
    w[:, :8] = torch.tensor([0.001, 0.003, -0.001, -0.003, 0.0, 100.0, -100.0, nan])
    q = torch.quantize_per_channel(w, sc.double(), zp, 0, torch.qint8)
    out["qp_c/scale_in"] = np.float32(scales); out["qp_c/shift_in"] = np.float32(shifts)
    out["qp_c/scale"] = sc.numpy(); out["qp_c/zp"] = zp.numpy()
    out["tq_c/x"] = w.numpy(); out["tq_c/codes"] = q.int_repr().numpy()
    y = lsq(w, torch.tensor(scales), torch.tensor(shifts), -128, 127, -128, 127, 0, False, 1.0, False, True, True, False)
    out["lsq_c/y"] = y.numpy()

    # ---- per-channel activations (quint8 affine, axis 1) ----
    scales = [0.03, 0.05, 0.011]
    shifts = [-1.7, 0.4, -0.02]
    m = ref_module(torch.quint8, torch.per_channel_affine, scales, shifts, 'activation')
    sc, sh, zp = m.calculate_qparams(need_shift=True)
    x = torch.randn(4, 3, 11, generator=gen) * 2.0
    q = torch.quantize_per_channel(x, sc.double(), zp, 1, torch.quint8)
    out["qp_a/scale_in"] = np.float32(scales); out["qp_a/shift_in"] = np.float32(shifts)
    out["qp_a/scale"] = sc.numpy(); out["qp_a/zp"] = zp.numpy()
    out["tq_a/x"] = x.numpy(); out["tq_a/codes"] = q.int_repr().numpy()
    y = lsq(x, torch.tensor(scales), torch.tensor(shifts), 0, 255, 0, 255, 1, False, 1.0, True, True, True, False)
    out["lsq_a/y"] = y.numpy()

    np.savez_compressed(OUT / "ref_export.npz", **out)
    print("wrote", OUT / "ref_export.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
