#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE itself (CPU op of torchlsq 2.1,
built by oracle/build_ref.py into oracle/_ref/).  Run in the build container only:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Outputs (small, committed):
    tests/golden/ref_cpu_ops.npz        inputs + outputs of torch.ops.torchlsq.* (reference CPU kernels)
    tests/golden/ref_module_trace.json  LSQFakeQuantizer state-machine traces + mu+-3sigma init values

The reference has no tests or golden vectors of its own; these pin the
oracle (oracle/lsq_oracle.c, contract=0 == reference CPU build) and the host logic.
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
from torchlsq.quantized.modules import observers as ref_obs  # noqa: E402

assert "oracle/_ref" in torchlsq.__file__, torchlsq.__file__
OUT = Path(__file__).resolve().parent
torch.set_num_threads(4)


def run_case(x, g, scale, shift, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, affine, per_channel, eval_mode, init_mode):
    x = x.clone().requires_grad_(True)
    s = scale.clone().requires_grad_(True)
    b = shift.clone().requires_grad_(True)
    y = lsq(x, s, b, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, affine, per_channel, eval_mode, init_mode)
    y.backward(g)
    return (y.detach().numpy(), x.grad.numpy(), s.grad.numpy(),
            b.grad.numpy() if b.grad is not None else np.zeros_like(b.detach().numpy()))


def main():
    cases = {}
    meta = {}

    def add(name, x, g, scale, shift, qmin=0, qmax=127, tmin=0, tmax=255, axis=1, use_gs=False, gscaler=1.0,
            affine=True, per_channel=False, eval_mode=False, init_mode=False):
        x = x.float(); g = g.float()
        scale = torch.as_tensor(scale, dtype=torch.float32).reshape(-1)
        shift = torch.as_tensor(shift, dtype=torch.float32).reshape(-1)
        y, dx, ds, db = run_case(x, g, scale, shift, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, affine,
                                 per_channel, eval_mode, init_mode)
        for k, v in (("x", x.numpy()), ("g", g.numpy()), ("scale", scale.numpy()), ("shift", shift.numpy()),
                     ("y", y), ("dx", dx), ("ds", ds), ("db", db)):
            cases[f"{name}/{k}"] = v
        meta[name] = dict(qmin=qmin, qmax=qmax, tmin=tmin, tmax=tmax, axis=axis, use_gs=use_gs, gscaler=gscaler,
                          affine=affine, per_channel=per_channel, eval_mode=eval_mode, init_mode=init_mode,
                          shape=list(x.shape))

    # --- SURVEY.md Appendix B known-answer set ------------------------------------------------
    nan, inf = float("nan"), float("inf")
    xb = torch.tensor([-1, -0.26, 0, 0.125, 0.375, 0.625, 0.874, 0.876, 31.5, 31.75, 32, 100, nan, inf, -inf])
    gb = torch.arange(1.0, 16.0)
    add("B_A", xb, gb, [0.25], [0.0])
    add("B_B", xb, gb, [0.25], [-0.6])
    add("B_C", xb, gb, [0.25], [0.0], qmin=-128, qmax=127, tmin=-128, tmax=127, affine=False)
    add("B_D", xb, gb, [0.25], [-0.6], eval_mode=True)
    add("B_E", xb[:12], gb[:12], [0.25], [-0.6], init_mode=True)
    add("B_F", xb[:12], gb[:12], [0.25], [-0.6], use_gs=True, gscaler=2.0)
    add("B_G", xb[:12], gb[:12], [-0.25], [-0.6])
    add("B_H", xb[:12].reshape(2, 3, 2), gb[:12].reshape(2, 3, 2), [0.25, 0.5, 1e-9], [0.0, -0.6, 0.0],
        use_gs=True, per_channel=True, axis=1)

    # --- seeded random cases --------------------------------------------------------------------
    gen = torch.Generator().manual_seed(1234)
    x1 = torch.randn(4099, generator=gen) * 1.5
    g1 = torch.randn(4099, generator=gen)
    add("R_tensor_affine", x1, g1, [0.03], [-1.7], use_gs=True)
    add("R_tensor_sym", x1, g1, [0.02], [0.0], qmin=-64, qmax=63, tmin=-128, tmax=127, affine=False, use_gs=True)
    add("R_tensor_init", x1, g1, [0.03], [-1.7], init_mode=True, use_gs=True)
    add("R_tensor_eval", x1, g1, [0.03], [-1.7], eval_mode=True)
    add("R_tensor_tiny_scale", x1, g1, [1e-12], [0.0])
    add("R_tensor_q255", x1, g1, [0.011], [-1.3], qmin=0, qmax=255, tmin=0, tmax=255, use_gs=True, gscaler=0.5)
    # every half-integer pre-image: x = (k + 0.5 - zp) * s, ties must round to even
    s, zp = 0.125, 17.0
    xt = (torch.arange(-3, 131, dtype=torch.float32) + 0.5 - zp) * s
    add("R_ties", xt, torch.ones_like(xt), [s], [-zp * s])
    x3 = torch.randn(3, 5, 7, generator=gen)
    g3 = torch.randn(3, 5, 7, generator=gen)
    for ax, C in ((0, 3), (1, 5), (2, 7)):
        sc = 0.02 + 0.02 * torch.rand(C, generator=gen)
        sh = -torch.rand(C, generator=gen)
        add(f"R_channel_axis{ax}", x3, g3, sc, sh, per_channel=True, axis=ax, use_gs=True)
    scw = 0.01 + 0.01 * torch.rand(6, generator=gen)
    xw = torch.randn(6, 4, 3, 3, generator=gen) * 0.05
    add("R_weight_sym", xw, torch.randn(6, 4, 3, 3, generator=gen), scw, torch.zeros(6), qmin=-128, qmax=127,
        tmin=-128, tmax=127, affine=False, per_channel=True, axis=0, use_gs=True)
    np.savez_compressed(OUT / "ref_cpu_ops.npz", **cases)

    # --- module traces: which (eval_mode, init_mode) the module asks for, step by step ---------
    traces = {}
    calls = []
    real_lsq = ref_obs.lsq

    def spy(x, scale, shift, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, is_affine, is_perchannel, eval_mode=False,
            init_mode=False):
        calls.append(dict(qmin=qmin, qmax=qmax, tmin=tmin, tmax=tmax, axis=axis, use_gs=use_gs, gscaler=gscaler,
                          is_affine=is_affine, is_perchannel=is_perchannel, eval_mode=bool(eval_mode),
                          init_mode=bool(init_mode), scale_rg=bool(scale.requires_grad),
                          shift_rg=bool(shift.requires_grad), scale=[float(v) for v in scale.detach().reshape(-1)[:4]],
                          shift=[float(v) for v in shift.detach().reshape(-1)[:4]]))
        return real_lsq(x, scale, shift, qmin, qmax, tmin, tmax, axis, use_gs, gscaler, is_affine, is_perchannel,
                        eval_mode, init_mode)

    ref_obs.lsq = spy
    ref_obs.partial = __import__("functools").partial   # D10: the reference forgot this import
    MA = torch.quantization.MovingAverageMinMaxObserver

    def trace(name, build, steps, train_flags=None, x_shape=(4, 6)):
        torch.manual_seed(7)
        m = build()
        rec = []
        for i in range(steps):
            m.train(train_flags[i] if train_flags else True)
            calls.clear()
            x = torch.randn(*x_shape) + 0.5
            out = m(x)
            rec.append(dict(step=i, identity=bool(out is x), call=(calls[0] if calls else None),
                            observer_enabled=int(m.observer_enabled[0]), current_batch=int(m.current_batch[0]),
                            n_batches=int(m.n_batches)))
        traces[name] = rec

    trace("act_learnable_n3", lambda: ref_obs.LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=3), 8)
    trace("act_observer_n2", lambda: ref_obs.LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2), 7)
    trace("act_observer_static", lambda: ref_obs.LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2,
                                                                 learn_params=False), 5)
    trace("act_learnable_evalmix", lambda: ref_obs.LSQFakeQuantizer(None, 'activation', init_mode='learnable',
                                                                   init_batches=2), 7,
          train_flags=[True, True, False, True, True, False, True])
    trace("weight_sym", lambda: ref_obs.LSQFakeQuantizer(None, 'weight', dtype=torch.qint8,
                                                         qscheme=torch.per_channel_symmetric, init_mode='learnable'), 4,
          x_shape=(6, 4, 3, 3))
    trace("weight_sym_8bit_tensor", lambda: ref_obs.LSQFakeQuantizer(None, 'weight', dtype=torch.qint8,
                                                                     qscheme=torch.per_tensor_symmetric,
                                                                     init_mode='learnable', avoid_torch_overflow=False), 3,
          x_shape=(6, 4, 3, 3))
    ref_obs.lsq = real_lsq

    # qrange table
    qr = {}
    for tag, kw in (("act_default", dict(otype='activation')),
                    ("act_8bit", dict(otype='activation', avoid_torch_overflow=False)),
                    ("act_sym", dict(otype='activation', qscheme=torch.per_tensor_symmetric)),
                    ("act_custom", dict(otype='activation', quant_min=0, quant_max=15)),
                    ("w_default", dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric)),
                    ("w_8bit", dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric,
                                    avoid_torch_overflow=False)),
                    ("w_custom", dict(otype='weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric,
                                      quant_min=-7, quant_max=8, init_scale=0.5))):
        m = ref_obs.LSQFakeQuantizer(None, init_mode='learnable', **kw)
        qr[tag] = dict(quant_min=m.quant_min, quant_max=m.quant_max, init_shift=m.init_shift, ch_axis=m.ch_axis,
                       n_batches=m.n_batches)

    # mu +- 3 sigma init values computed by the reference module (torch.mean / torch.std on CPU)
    winit = {}
    torch.manual_seed(11)
    for tag, shape, scheme, low in (("conv_64x3x7x7", (64, 3, 7, 7), torch.per_channel_symmetric, True),
                                    ("conv_32x16x3x3_8bit", (32, 16, 3, 3), torch.per_channel_symmetric, False),
                                    ("fc_10x64_tensor", (10, 64), torch.per_tensor_symmetric, True)):
        w = torch.randn(*shape) * 0.07 + 0.01
        m = ref_obs.LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=scheme, init_mode='learnable',
                                     avoid_torch_overflow=low)
        m(w)
        winit[tag] = dict(shape=list(shape), seed_note="torch.manual_seed(11) sequence; w stored in npz",
                          quant_min=m.quant_min, quant_max=m.quant_max, per_channel=m.is_perchannel,
                          scale=[float(v) for v in m.scale.detach()])
        cases_w = {f"winit/{tag}": w.numpy()}
        np.savez_compressed(OUT / f"ref_winit_{tag}.npz", **cases_w)

    (OUT / "ref_module_trace.json").write_text(json.dumps(dict(ops_meta=meta, traces=traces, qranges=qr, winit=winit),
                                                          indent=1))
    print("wrote", sorted(p.name for p in OUT.glob("ref_*")))


if __name__ == "__main__":
    main()
