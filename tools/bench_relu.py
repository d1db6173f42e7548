#!/usr/bin/env python
"""ReLU-prologue fusion measured on one B200: `relu` -> fake-quant as the reference runs it (two passes
each way: ATen relu + plain kernels forward, plain kernels + ATen threshold_backward backward) against the fused
`lsqb200_*_pre(..., LSQB200_PRE_RELU)` kernels (one pass each way), on ResNet-50 activation sites at batch 256, bf16.

CUDA-event timing, median after warm-up, working sets (0.4-1.6 GB) far larger than L2.  GB/s figures divide the FUSED
algorithmic bytes (forward R x + W y, backward R x + R g + W gx = 5 * sizeof(T) per element) by the time, so the two
columns are directly comparable: same useful work, different HBM traffic.  Prints one JSON object.
"""
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "lsqfakequantize-pytorch_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from torchlsq import _cabi  # noqa: E402

DEV = "cuda:0"


def timed(fn, iters=20, warm=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(iters + warm):
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    lib = _cabi.load()
    peak = 6650.0
    mp = ROOT / "MEASURED_PEAKS.json"
    if mp.exists():
        peak = json.loads(mp.read_text()).get("hbm_gbs", peak)
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    out = {"peak_GBps_measured_copy": peak, "dtype": "bf16", "note": "GB/s = fused algorithmic bytes (5*2 B/element) / time for both columns"}
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    s, b = torch.tensor([0.03], device=DEV), torch.tensor([-0.9], device=DEV)
    gs, gb = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    for name, shape in (("256x64x112x112", (256, 64, 112, 112)), ("256x256x56x56", (256, 256, 56, 56)),
                        ("256x512x28x28", (256, 512, 28, 28)), ("256x2048x7x7", (256, 2048, 7, 7))):
        gen = torch.Generator(device=DEV).manual_seed(1)
        x = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
        g = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
        xr, y, gx, gx2 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        n = x.numel()

        def unfused_fwd():
            torch.clamp_min(x, 0, out=xr)     # ATen relu == clamp_min(x, 0)
            lib.lsqb200_fwd_tensor(xr.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, sp)

        def unfused_bwd():
            lib.lsqb200_bwd_tensor(g.data_ptr(), xr.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                   n, 2, 0, q, ws.data_ptr(), ws.numel(), sp)
            torch.ops.aten.threshold_backward(gx, xr, 0, grad_input=gx2)

        def fused_fwd():
            lib.lsqb200_fwd_tensor_pre(x.data_ptr(), None, y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, 1, sp)

        def fused_bwd():
            lib.lsqb200_bwd_tensor_pre(g.data_ptr(), x.data_ptr(), None, gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                       n, 2, 0, q, 1, ws.data_ptr(), ws.numel(), sp)

        unfused_fwd(); unfused_bwd()
        y0, gxa, gs0 = y.clone(), gx2.clone(), gs.clone()
        fused_fwd(); fused_bwd()
        torch.cuda.synchronize()
        same = bool(torch.equal(y0.view(torch.int16), y.view(torch.int16)) and torch.equal(gxa.view(torch.int16), gx.view(torch.int16)))
        t_uf, t_ub, t_ff, t_fb = timed(unfused_fwd), timed(unfused_bwd), timed(fused_fwd), timed(fused_bwd)
        alg = 5 * 2 * n
        out[name] = dict(elements=n, bit_identical=same, gscale_rel_diff=float(((gs - gs0).abs() / gs0.abs()).item()),
                         unfused_ms=dict(fwd=round(t_uf, 4), bwd=round(t_ub, 4)), fused_ms=dict(fwd=round(t_ff, 4), bwd=round(t_fb, 4)),
                         unfused_GBps=round(alg / (t_uf + t_ub) / 1e6, 1), fused_GBps=round(alg / (t_ff + t_fb) / 1e6, 1),
                         fused_frac_of_measured_peak=round(alg / (t_ff + t_fb) / 1e6 / peak, 3),
                         speedup=round((t_uf + t_ub) / (t_ff + t_fb), 3))
        # ---- residual join: relu(a + b) -> fake-quant.  Unfused as torchvision's Bottleneck runs it: in-place add (R a, R b, W a),
        #      in-place relu (R, W), fake-quant (R, W) = 7 trips; backward fake-quant (R x', R g, W) + relu backward (R, R, W) = 6
        x2 = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
        tsum = torch.empty_like(x)

        def unfused_add_fwd():
            torch.add(x, x2, out=tsum)
            tsum.relu_()
            lib.lsqb200_fwd_tensor(tsum.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, sp)

        def unfused_add_bwd():
            lib.lsqb200_bwd_tensor(g.data_ptr(), tsum.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                   n, 2, 0, q, ws.data_ptr(), ws.numel(), sp)
            torch.ops.aten.threshold_backward(gx, tsum, 0, grad_input=gx2)

        def fused_add_fwd():
            lib.lsqb200_fwd_tensor_pre(x.data_ptr(), x2.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, 2, sp)

        def fused_add_bwd():
            lib.lsqb200_bwd_tensor_pre(g.data_ptr(), x.data_ptr(), x2.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                       gb.data_ptr(), n, 2, 0, q, 2, ws.data_ptr(), ws.numel(), sp)

        unfused_add_fwd(); unfused_add_bwd()
        y0, gxa, gs0 = y.clone(), gx2.clone(), gs.clone()
        fused_add_fwd(); fused_add_bwd()
        torch.cuda.synchronize()
        same = bool(torch.equal(y0.view(torch.int16), y.view(torch.int16)) and torch.equal(gxa.view(torch.int16), gx.view(torch.int16)))
        t_uf, t_ub, t_ff, t_fb = timed(unfused_add_fwd), timed(unfused_add_bwd), timed(fused_add_fwd), timed(fused_add_bwd)
        alg = 7 * 2 * n    # fused algorithmic bytes: forward R a + R b + W y, backward R a + R b + R g + W gx
        out[name + "_residual_join"] = dict(
            elements=n, bit_identical=same, gscale_rel_diff=float(((gs - gs0).abs() / gs0.abs()).item()),
            unfused_ms=dict(fwd=round(t_uf, 4), bwd=round(t_ub, 4)), fused_ms=dict(fwd=round(t_ff, 4), bwd=round(t_fb, 4)),
            unfused_GBps=round(alg / (t_uf + t_ub) / 1e6, 1), fused_GBps=round(alg / (t_ff + t_fb) / 1e6, 1),
            fused_frac_of_measured_peak=round(alg / (t_ff + t_fb) / 1e6 / peak, 3), speedup=round((t_uf + t_ub) / (t_ff + t_fb), 3),
            note="GB/s = 7*2 B/element (fused algorithmic bytes) / time for both columns")
        del x, g, xr, y, gx, gx2, x2, tsum
    print(json.dumps(out))


if __name__ == "__main__":
    main()
