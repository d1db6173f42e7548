#!/usr/bin/env python
"""Per-config measurements for BASELINE.json configs[0..3] on one B200 (bench.py covers configs[4]).

CUDA-event timing, median of `--iters` after warm-up; a 512 MB L2-flush write is enqueued in front of the
timed events of every iteration (outside them): it empties L2 for the small working sets and keeps the GPU
busy while the host prepares the call, so wrapper latency is not counted as kernel time.  GB/s = algorithmic bytes / time
(5*sizeof(T) per element for fwd+bwd, 1*sizeof(T) for the mu+-3sigma statistics).
Prints one JSON object; copy it to profiles/ to keep it.
"""
import argparse
import json
import math
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "lsqfakequantize-pytorch_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench as B  # noqa: E402
from torchlsq import _cabi  # noqa: E402
from torchlsq.multi import LSQPlan, Site  # noqa: E402

DEV = "cuda:0"


def timed(fn, iters, flush, cover=1):
    """`cover` flush writes (about 80 us of GPU time each) are queued in front of the timed events so the host side of
    `fn` (Python wrappers take up to ~100 us) runs while the GPU is still busy and is not counted as kernel time."""
    times = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(iters + 3):
        if flush is not None:
            for _ in range(cover):
                flush.fill_(i)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    return statistics.median(times), min(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = _cabi.load()
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    out = {"peak_GBps_measured_copy": peak}

    def report(name, bytes_, med, best, **extra):
        out[name] = dict(ms_median=round(med, 4), ms_best=round(best, 4), GBps_median=round(bytes_ / med / 1e6, 1),
                         GBps_best=round(bytes_ / best / 1e6, 1), frac_of_measured_peak=round(bytes_ / med / 1e6 / peak, 3), **extra)

    # ---- config 1: per-tensor quint8, fp32 32x64x56x56
    x = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(1)).to(DEV)
    g = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(2)).to(DEV)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    s, b = torch.tensor([0.03], device=DEV), torch.tensor([-1.7], device=DEV)
    gs, gb = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    n = x.numel()

    def c1():
        lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 0, 0, q, sp)
        lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                               n, 0, 0, q, ws.data_ptr(), ws.numel(), sp)
    report("config1_per_tensor_fp32_32x64x56x56_fwd_bwd", 5 * 4 * n, *timed(c1, args.iters, flush), l2="flushed", launches=2)
    # size-matched ceiling: the same bytes moved by ATen's own streaming kernels (copy = R+W, add = 2R+W), same flush, 2 launches
    def c1_copy_add():
        y.copy_(x)
        torch.add(x, g, out=gx)
    report("config1_ceiling_aten_copy_plus_add_same_bytes", 5 * 4 * n, *timed(c1_copy_add, args.iters, flush), l2="flushed", launches=2)
    report("config1_same_L2_warm", 5 * 4 * n, *timed(c1, args.iters, None), l2="warm (103 MB working set fits the 126 MB L2)", launches=2)

    # ---- config 2: 54 ResNet-50 weights, per-channel axis 0, symmetric qint8, mu+-3sigma init
    gen = torch.Generator(device=DEV).manual_seed(0)
    sites = []
    for shp in B.W_SHAPES:
        w = torch.empty(shp, device=DEV).normal_(0, 0.05, generator=gen)
        sites.append(Site(x=w, y=torch.empty_like(w), grad=torch.randn(shp, device=DEV, generator=gen), gx=torch.empty_like(w),
                          scale=torch.full((shp[0],), 0.002, device=DEV), shift=torch.zeros(shp[0], device=DEV),
                          gscale=torch.empty(shp[0], device=DEV), gshift=torch.empty(shp[0], device=DEV),
                          quant_min=-128, quant_max=127, type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True))
    plan = LSQPlan(sites)
    nw = sum(math.prod(sh) for sh in B.W_SHAPES)
    scales = torch.empty(plan.num_param_slots, device=DEV)
    report("config2_weights_mu3sigma_init_54_tensors", 4 * nw, *timed(lambda: plan.weight_init_stats(scales), args.iters, flush),
           l2="flushed", launches=2)

    def c2():
        plan.forward()
        plan.backward()
    report("config2_weights_fwd_bwd_54_tensors_plan", 5 * 4 * nw, *timed(c2, args.iters, flush), l2="flushed",
           launches=plan.launches(False) + plan.launches(True))
    fa, fb, fc = (torch.empty(nw, device=DEV).normal_(generator=gen) for _ in range(3))

    def c2_copy_add():
        fc.copy_(fa)
        torch.add(fa, fb, out=fc)
    report("config2_ceiling_aten_copy_plus_add_same_bytes_flat", 5 * 4 * nw, *timed(c2_copy_add, args.iters, flush), l2="flushed", launches=2)
    report("config2_ceiling_aten_sum_same_bytes_flat", 4 * nw, *timed(lambda: fa.sum(), args.iters, flush), l2="flushed", launches=2)
    qw = _cabi.qargs(-128, 127, -128, 127, True, 1.0, True, False, False)

    def c2_per_tensor():
        for st in sites:
            C, K = st.x.shape[0], st.x[0].numel()
            lib.lsqb200_fwd_channel(st.x.data_ptr(), st.y.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(), 1, C, K, 0, 0, qw, sp)
        for st in sites:
            C, K = st.x.shape[0], st.x[0].numel()
            lib.lsqb200_bwd_channel(st.grad.data_ptr(), st.x.data_ptr(), st.gx.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
                                    st.gscale.data_ptr(), st.gshift.data_ptr(), 1, C, K, 0, 0, qw, ws.data_ptr(), ws.numel(), sp)
    report("config2_weights_fwd_bwd_54_tensors_per_tensor_calls", 5 * 4 * nw, *timed(c2_per_tensor, args.iters, flush), l2="flushed", launches=108)

    # ---- config 3: learned init (init_mode) on bf16 activation sites, batch 256: the 8 largest + 8 mid sites
    qi = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, True)
    shapes = [sh for sh in B.ACT_SHAPES if math.prod(sh) >= 100352][:24]
    bufs = []
    for sh in shapes:
        m = 256 * math.prod(sh)
        bufs.append((torch.empty(m, dtype=torch.bfloat16, device=DEV).normal_().relu_(), torch.empty(m, dtype=torch.bfloat16, device=DEV),
                     torch.empty(m, dtype=torch.bfloat16, device=DEV).normal_(), torch.empty(m, dtype=torch.bfloat16, device=DEV), m))

    def c3():
        for xx, yy, gg, gxx, m in bufs:
            lib.lsqb200_fwd_tensor(xx.data_ptr(), yy.data_ptr(), s.data_ptr(), b.data_ptr(), m, 2, 0, qi, sp)
        for xx, yy, gg, gxx, m in bufs:
            lib.lsqb200_bwd_tensor(gg.data_ptr(), xx.data_ptr(), gxx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                   m, 2, 0, qi, ws.data_ptr(), ws.numel(), sp)
    n3 = sum(m for *_, m in bufs)
    report("config3_learned_init_bf16_sites_batch256", 5 * 2 * n3, *timed(c3, max(5, args.iters // 2), None),
           l2="%d sites, %.1f GB resident" % (len(bufs), 4 * 2 * n3 / 1e9), launches=2 * len(bufs))
    del bufs

    # ---- config 4: per-channel axis 1, fp16 256x1024x28x28, grad scaling on, fp32 params
    N, C, HW = 256, 1024, 784
    x4 = torch.empty(N * C * HW, dtype=torch.float16, device=DEV).normal_()
    g4 = torch.empty_like(x4).normal_()
    y4, gx4 = torch.empty_like(x4), torch.empty_like(x4)
    s4 = 0.02 + 0.02 * torch.rand(C, device=DEV)
    b4 = -torch.rand(C, device=DEV)
    gs4, gb4 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)

    def c4():
        lib.lsqb200_fwd_channel(x4.data_ptr(), y4.data_ptr(), s4.data_ptr(), b4.data_ptr(), N, C, HW, 1, 0, q, sp)
        lib.lsqb200_bwd_channel(g4.data_ptr(), x4.data_ptr(), gx4.data_ptr(), s4.data_ptr(), b4.data_ptr(), gs4.data_ptr(), gb4.data_ptr(),
                                N, C, HW, 1, 0, q, ws.data_ptr(), ws.numel(), sp)
    report("config4_per_channel_fp16_256x1024x28x28_fwd_bwd", 5 * 2 * x4.numel(), *timed(c4, args.iters, flush), l2="1.6 GB working set", launches=2)
    # same tensor, other per-channel layouts the ResNet activations produce
    for hw, cc in ((196, 1024), (49, 2048), (3136, 256)):
        nn_ = N * cc * hw
        sc, bc = 0.02 + 0.02 * torch.rand(cc, device=DEV), -torch.rand(cc, device=DEV)
        gsc, gbc = torch.empty(cc, device=DEV), torch.empty(cc, device=DEV)

        def cx():
            lib.lsqb200_fwd_channel(x4.data_ptr(), y4.data_ptr(), sc.data_ptr(), bc.data_ptr(), N, cc, hw, 1, 0, q, sp)
            lib.lsqb200_bwd_channel(g4.data_ptr(), x4.data_ptr(), gx4.data_ptr(), sc.data_ptr(), bc.data_ptr(), gsc.data_ptr(), gbc.data_ptr(),
                                    N, cc, hw, 1, 0, q, ws.data_ptr(), ws.numel(), sp)
        report(f"per_channel_fp16_256x{cc}x{hw}_fwd_bwd", 5 * 2 * nn_, *timed(cx, max(5, args.iters // 2), flush), l2="large", launches=2)
    # channels-last per-channel (inner = 1)
    nn_ = 256 * 196 * 1024
    sc, bc = 0.02 + 0.02 * torch.rand(1024, device=DEV), -torch.rand(1024, device=DEV)
    gsc, gbc = torch.empty(1024, device=DEV), torch.empty(1024, device=DEV)

    def cl():
        lib.lsqb200_fwd_channel(x4.data_ptr(), y4.data_ptr(), sc.data_ptr(), bc.data_ptr(), 256 * 196, 1024, 1, 1, 0, q, sp)
        lib.lsqb200_bwd_channel(g4.data_ptr(), x4.data_ptr(), gx4.data_ptr(), sc.data_ptr(), bc.data_ptr(), gsc.data_ptr(), gbc.data_ptr(),
                                256 * 196, 1024, 1, 1, 0, q, ws.data_ptr(), ws.numel(), sp)
    report("per_channel_fp16_channels_last_50176x1024_fwd_bwd", 5 * 2 * nn_, *timed(cl, 5, flush), l2="large", launches=2)
    # ---- observer-mode init step: fused native step vs torch's observer + host logic, 256x256x56x56 bf16
    import warnings
    from torchlsq.quantized.modules.observers import observer_step
    xo = torch.empty(256 * 256 * 56 * 56, dtype=torch.bfloat16, device=DEV).normal_()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        obs_n = torch.quantization.MovingAverageMinMaxObserver(reduce_range=True).to(DEV)
        obs_t = torch.quantization.MovingAverageMinMaxObserver(reduce_range=True).to(DEV)
    so, bo = torch.ones(1, device=DEV), torch.zeros(1, device=DEV)

    def torch_path():
        obs_t(xo)
        sc, zp = obs_t.calculate_qparams()
        so.copy_(sc); bo.copy_(-zp * so)
    report("observer_step_native_bf16_411MB", 2 * xo.numel(), *timed(lambda: observer_step(obs_n, xo, so, bo), args.iters, flush, cover=3), l2="411 MB", launches=1)
    report("observer_step_torch_path_bf16_411MB", 2 * xo.numel(), *timed(torch_path, max(5, args.iters // 2), flush), l2="411 MB",
           note="x.to(float32) + aminmax + qparams kernels + host syncs, as the reference module does")
    # ---- integer export: bf16 activation -> uint8 codes (3 B / element), codes -> bf16, vs torch's own path
    from torchlsq import export as EX
    s1, b1 = torch.tensor([0.03], device=DEV), torch.tensor([-1.7], device=DEV)
    xe = xo.view(256, 256, 56, 56)
    for sem in ("lsq", "torch"):
        report(f"export_quantize_{sem}_bf16_to_u8_411MB", 3 * xe.numel(), *timed(lambda: EX.quantize(xe, s1, b1, 0, 127, 0, 255, semantics=sem), args.iters, flush, cover=3),
               l2="411 MB in", launches=1)
    ce = EX.quantize(xe, s1, b1, 0, 127, 0, 255)
    report("export_dequantize_lsq_u8_to_bf16_205MB", 3 * xe.numel(), *timed(lambda: EX.dequantize(ce, s1, b1, 0, 127, 0, 255, dtype=torch.bfloat16), args.iters, flush, cover=3),
           l2="205 MB in", launches=1)
    xf = xe[:64].float()
    report("export_torch_quantize_per_tensor_f32_205MB", 5 * xf.numel(), *timed(lambda: torch.quantize_per_tensor(xf, 0.03, 57, torch.quint8), max(5, args.iters // 2), flush),
           l2="205 MB in", note="torch's CUDA quantizer (fp32 input only), for comparison")
    report("export_quantize_torch_f32_to_u8_205MB", 5 * xf.numel(), *timed(lambda: EX.quantize(xf, s1, b1, 0, 127, 0, 255, semantics='torch'), args.iters, flush, cover=3),
           l2="205 MB in", launches=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
