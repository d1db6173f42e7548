#!/usr/bin/env python
"""Golden vectors for the fused ReLU prologue, generated from the REFERENCE's own call
sequence: `torch.relu` followed by the reference CPU op (torchlsq 2.1 built by oracle/build_ref.py into oracle/_ref/),
autograd through both.  Run in the build container only:

    python oracle/build_ref.py && python tests/golden/make_golden_relu.py

Output (small, committed): tests/golden/ref_cpu_relu.npz -- pins oracle.forward_relu / backward_relu and, for the
`A_*` cases (`a + b` [-> relu] -> reference op), oracle.forward_add / backward_add (contract 0).
"""
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
    cases = {}

    def add(name, x, g, scale, shift, qmin=0, qmax=127, tmin=0, tmax=255, axis=1, use_gs=False, gscaler=1.0,
            affine=True, per_channel=False, eval_mode=False, init_mode=False, x2=None, relu=True):
        x = x.float().clone().requires_grad_(True)
        g = g.float()
        if x2 is not None:
            x2 = x2.float().clone().requires_grad_(True)
        s = torch.as_tensor(scale, dtype=torch.float32).reshape(-1).clone().requires_grad_(True)
        b = torch.as_tensor(shift, dtype=torch.float32).reshape(-1).clone().requires_grad_(True)
        pre = x if x2 is None else x + x2
        y = lsq(torch.relu(pre) if relu else pre, s, b, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, affine, per_channel, eval_mode, init_mode)
        y.backward(g)
        db = b.grad.numpy() if b.grad is not None else np.zeros(b.numel(), np.float32)
        for k, v in (("x", x.detach().numpy()), ("g", g.numpy()), ("scale", s.detach().numpy()), ("shift", b.detach().numpy()),
                     ("y", y.detach().numpy()), ("dx", x.grad.numpy()), ("ds", s.grad.numpy()), ("db", db),
                     ("meta", np.array([qmin, qmax, tmin, tmax, axis, int(use_gs), int(affine), int(per_channel),
                                        int(eval_mode), int(init_mode)], np.int64)),
                     ("gscaler", np.array([gscaler], np.float64))):
            cases[f"{name}/{k}"] = v
        if x2 is not None:
            cases[f"{name}/x2"] = x2.detach().numpy()
            cases[f"{name}/dx2"] = x2.grad.numpy()
            cases[f"{name}/relu"] = np.array([int(relu)], np.int64)

    nan, inf = float("nan"), float("inf")
    xb = torch.tensor([-1, -0.26, -0.0, 0, 0.125, 0.375, 0.625, 0.874, 0.876, 31.5, 31.75, 32, 100, nan, inf, -inf, -1e-30, 1e-30])
    gb = torch.arange(1.0, 19.0)
    add("K_zp0", xb, gb, [0.25], [0.0])
    add("K_zp_pos", xb, gb, [0.25], [-0.6])            # zp = 2: relu'd zeros land inside the range
    add("K_zp_big", xb, gb, [0.25], [-20.0])           # zp = 80
    add("K_shift_pos", xb, gb, [0.25], [0.9])          # -shift/s < 0 -> zp clamps to type_min = 0; zeros sit ON the lower border
    add("K_sym", xb, gb, [0.25], [0.0], qmin=-128, qmax=127, tmin=-128, tmax=127, affine=False)
    add("K_eval", xb, gb, [0.25], [-0.6], eval_mode=True)
    add("K_init", xb[:13], gb[:13], [0.25], [-0.6], init_mode=True)
    add("K_init_nan", xb, gb, [0.25], [-0.6], init_mode=True)
    add("K_gs", xb[:13], gb[:13], [0.25], [-0.6], use_gs=True, gscaler=2.0)
    gen = torch.Generator().manual_seed(4321)
    x1 = torch.randn(4099, generator=gen) * 1.5
    g1 = torch.randn(4099, generator=gen)
    add("R_tensor", x1, g1, [0.03], [-1.7], use_gs=True)
    add("R_tensor_zp0", x1, g1, [0.03], [0.0], use_gs=True)
    add("R_tensor_init", x1, g1, [0.03], [-1.7], init_mode=True, use_gs=True)
    add("R_tensor_eval", x1, g1, [0.03], [-1.7], eval_mode=True)
    x3 = torch.randn(3, 5, 7, generator=gen)
    g3 = torch.randn(3, 5, 7, generator=gen)
    for ax, C in ((0, 3), (1, 5), (2, 7)):
        sc = 0.02 + 0.02 * torch.rand(C, generator=gen)
        sh = -torch.rand(C, generator=gen)
        add(f"R_channel_axis{ax}", x3, g3, sc, sh, per_channel=True, axis=ax, use_gs=True)
    # residual joins: lsq(relu(a + b)) and lsq(a + b)
    xb2 = torch.tensor([0.5, 0.26, 0.0, -0.0, -0.125, 1e-8, -0.625, 0.002, -0.002, 1.0, -31.75, -64, nan, 1.0, -inf, inf, 1e-30, -1e-30])
    add("A_K_addrelu", xb, gb, [0.25], [-0.6], x2=xb2)
    add("A_K_add", xb, gb, [0.25], [-0.6], x2=xb2, relu=False)
    add("A_K_addrelu_init", xb, gb, [0.25], [-0.6], x2=xb2, init_mode=True)
    add("A_K_add_init", xb[:12], gb[:12], [0.25], [-0.6], x2=xb2[:12], init_mode=True, relu=False)
    x2r = torch.randn(4099, generator=gen)
    add("A_R_addrelu", x1, g1, [0.03], [-1.7], use_gs=True, x2=x2r)
    add("A_R_add", x1, g1, [0.03], [-1.7], use_gs=True, x2=x2r, relu=False)
    add("A_R_addrelu_eval", x1, g1, [0.03], [-1.7], eval_mode=True, x2=x2r)
    x32 = torch.randn(3, 5, 7, generator=gen)
    sc = 0.02 + 0.02 * torch.rand(5, generator=gen)
    sh = -torch.rand(5, generator=gen)
    add("A_R_addrelu_channel", x3, g3, sc, sh, per_channel=True, axis=1, use_gs=True, x2=x32)
    add("A_R_add_channel", x3, g3, sc, sh, per_channel=True, axis=1, use_gs=True, x2=x32, relu=False)
    np.savez_compressed(OUT / "ref_cpu_relu.npz", **cases)
    print("wrote", OUT / "ref_cpu_relu.npz", len(cases), "arrays")


if __name__ == "__main__":
    main()
