#!/usr/bin/env python
"""float64 golden fixtures from the REFERENCE itself (CPU op of torchlsq 2.1 built by oracle/build_ref.py).
Run in the build container only:

    python oracle/build_ref.py && python tests/golden/make_golden_f64.py

Output (small, committed): tests/golden/ref_cpu_ops_f64.npz -- inputs + outputs of the reference CPU kernels on
float64 tensors (the reference dispatches double: AT_DISPATCH_FLOATING_TYPES, csrc/ops/cpu/lsq_cpu.cpp:37,92,180,242),
plus the case table as a JSON string under the key "meta".  Pins oracle/lsq_oracle.c's float64 restatement with
contract = 0 (the CPU build); the CUDA build's float64 arithmetic differs (clamps through float) and is pinned on
the GPU box against the reference CUDA op (tests/test_gpu_f64.py).
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))   # the REFERENCE package, not ours

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchlsq  # noqa: E402  (reference)
from torchlsq.functional import lsq  # noqa: E402

assert "oracle/_ref" in torchlsq.__file__, torchlsq.__file__
OUT = Path(__file__).resolve().parent
torch.set_num_threads(4)


def main():
    cases, meta = {}, {}

    def add(name, x, g, scale, shift, qmin=0, qmax=127, tmin=0, tmax=255, axis=1, use_gs=False, gscaler=1.0,
            affine=True, per_channel=False, eval_mode=False, init_mode=False):
        x = x.double().clone().requires_grad_(True)
        g = g.double()
        s = torch.as_tensor(scale, dtype=torch.float64).reshape(-1).clone().requires_grad_(True)
        b = torch.as_tensor(shift, dtype=torch.float64).reshape(-1).clone().requires_grad_(True)
        y = lsq(x, s, b, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, affine, per_channel, eval_mode, init_mode)
        y.backward(g)
        db = b.grad.numpy() if b.grad is not None else np.zeros(b.numel())
        for k, v in (("x", x.detach().numpy()), ("g", g.numpy()), ("scale", s.detach().numpy()), ("shift", b.detach().numpy()),
                     ("y", y.detach().numpy()), ("dx", x.grad.numpy()), ("ds", s.grad.numpy()), ("db", db)):
            assert v.dtype == np.float64
            cases[f"{name}/{k}"] = v
        meta[name] = dict(qmin=qmin, qmax=qmax, tmin=tmin, tmax=tmax, axis=axis, use_gs=use_gs, gscaler=gscaler,
                          affine=affine, per_channel=per_channel, eval_mode=eval_mode, init_mode=init_mode,
                          shape=list(x.shape))

    nan, inf = float("nan"), float("inf")
    xb = torch.tensor([-1, -0.26, 0, 0.125, 0.375, 0.625, 0.874, 0.876, 31.5, 31.75, 32, 100, nan, inf, -inf], dtype=torch.float64)
    gb = torch.arange(1.0, 16.0, dtype=torch.float64)
    add("B_A", xb, gb, [0.25], [0.0])
    add("B_B", xb, gb, [0.25], [-0.6])
    add("B_C", xb, gb, [0.25], [0.0], qmin=-128, qmax=127, tmin=-128, tmax=127, affine=False)
    add("B_D", xb, gb, [0.25], [-0.6], eval_mode=True)
    add("B_E", xb[:12], gb[:12], [0.25], [-0.6], init_mode=True)
    add("B_F", xb[:12], gb[:12], [0.25], [-0.6], use_gs=True, gscaler=2.0)
    add("B_G", xb[:12], gb[:12], [-0.25], [-0.6])
    add("B_H", xb[:12].reshape(2, 3, 2), gb[:12].reshape(2, 3, 2), [0.25, 0.5, 1e-17], [0.0, -0.6, 0.0],
        use_gs=True, per_channel=True, axis=1)

    gen = torch.Generator().manual_seed(4321)
    x1 = torch.randn(1031, generator=gen, dtype=torch.float64) * 1.5
    g1 = torch.randn(1031, generator=gen, dtype=torch.float64)
    add("R_tensor_affine", x1, g1, [0.03], [-1.7], use_gs=True)
    add("R_tensor_sym", x1, g1, [0.02], [0.0], qmin=-64, qmax=63, tmin=-128, tmax=127, affine=False, use_gs=True)
    add("R_tensor_init", x1, g1, [0.03], [-1.7], init_mode=True, use_gs=True)
    add("R_tensor_eval", x1, g1, [0.03], [-1.7], eval_mode=True)
    add("R_tensor_tiny_scale", x1, g1, [1e-20], [0.0])
    # a scale that is NOT representable in float: tells the CPU build (double scale) from the CUDA one (per-channel: float scale)
    add("R_tensor_q255", x1, g1, [0.0110000000001], [-1.3], qmin=0, qmax=255, tmin=0, tmax=255, use_gs=True, gscaler=0.5)
    s, zp = 0.125, 17.0
    xt = (torch.arange(-3, 131, dtype=torch.float64) + 0.5 - zp) * s
    add("R_ties", xt, torch.ones_like(xt), [s], [-zp * s])
    x3 = torch.randn(3, 5, 7, generator=gen, dtype=torch.float64)
    g3 = torch.randn(3, 5, 7, generator=gen, dtype=torch.float64)
    for ax, C in ((0, 3), (1, 5), (2, 7)):
        sc = 0.02 + 0.02 * torch.rand(C, generator=gen, dtype=torch.float64)
        sh = -torch.rand(C, generator=gen, dtype=torch.float64)
        add(f"R_channel_axis{ax}", x3, g3, sc, sh, per_channel=True, axis=ax, use_gs=True)
    scw = 0.01 + 0.01 * torch.rand(6, generator=gen, dtype=torch.float64)
    xw = torch.randn(6, 4, 3, 3, generator=gen, dtype=torch.float64) * 0.05
    add("R_weight_sym", xw, torch.randn(6, 4, 3, 3, generator=gen, dtype=torch.float64), scw, torch.zeros(6), qmin=-128, qmax=127,
        tmin=-128, tmax=127, affine=False, per_channel=True, axis=0, use_gs=True)
    cases["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT / "ref_cpu_ops_f64.npz", **cases)
    print("wrote ref_cpu_ops_f64.npz:", len(meta), "cases")


if __name__ == "__main__":
    main()
