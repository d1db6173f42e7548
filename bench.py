#!/usr/bin/env python
"""bench.py -- LSQ+ fake-quantize forward+backward throughput on B200 (driver contract).

Workload ("resnet50_qat_fakequant"): BASELINE.json configs[4] evaluated per rank -- every
fake-quant site of a random-init ResNet-50 QAT step at 224x224:
  * 71 activation sites (input, 53 conv outputs, 16 add-relu outputs, fc), bf16, per-tensor
    quint8 q=[0,127] t=[0,255], affine, grad scaling on  (= configs[2] sites, normal LSQ mode)
  * 54 conv / fc weights, fp32, per-channel axis 0, symmetric qint8 [-128,127]   (= configs[1])
with `--batch-per-gpu` images per rank (default 256, so 8 ranks = the batch-2048 job of
configs[4]); activations shard by batch (weak scaling, no data-path exchange), weights are
replicated, and with N > 1 the 27 702-float flat grad_scale/grad_shift buffer is all-reduced
with NCCL once per step.  One step = forward of every site, then backward of every site.

metric: algorithmic bytes / time, GB/s, whole job.  Algorithmic bytes per element:
fwd R x + W y, bwd R x + R g + W gx = 5 * sizeof(T)  (SURVEY.md section 8d).
Every site owns distinct x / y / g / gx buffers (34 GB at batch 256), so the timed loop never
re-reads anything L2 (126 MB) could hold from a previous iteration.

`--impl reference` times the reference's own CPU op (oracle/_ref, built from /root/reference by
oracle/build_ref.py) through its public API on the host cores, on a bounded fp32 sample of the
same site list.
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "lsqfakequantize-pytorch_b200"

# (count, per-image shape) -- SURVEY.md Appendix D, torchvision resnet50
ACT_SITES = [(1, (3, 224, 224)), (1, (64, 112, 112)), (6, (64, 56, 56)), (7, (256, 56, 56)), (1, (128, 56, 56)),
             (7, (128, 28, 28)), (9, (512, 28, 28)), (1, (256, 28, 28)), (11, (256, 14, 14)), (13, (1024, 14, 14)),
             (1, (512, 14, 14)), (5, (512, 7, 7)), (7, (2048, 7, 7)), (1, (1000,))]
WEIGHTS = [(1, (64, 3, 7, 7)), (1, (64, 64, 1, 1)), (3, (64, 64, 3, 3)), (4, (256, 64, 1, 1)), (2, (64, 256, 1, 1)),
           (1, (128, 256, 1, 1)), (4, (128, 128, 3, 3)), (4, (512, 128, 1, 1)), (1, (512, 256, 1, 1)),
           (3, (128, 512, 1, 1)), (1, (256, 512, 1, 1)), (6, (256, 256, 3, 3)), (6, (1024, 256, 1, 1)),
           (1, (1024, 512, 1, 1)), (5, (256, 1024, 1, 1)), (1, (512, 1024, 1, 1)), (3, (512, 512, 3, 3)),
           (3, (2048, 512, 1, 1)), (1, (2048, 1024, 1, 1)), (2, (512, 2048, 1, 1)), (1, (1000, 2048))]


def _expand(table):
    out = []
    for count, shape in table:
        out += [shape] * count
    return out


ACT_SHAPES = _expand(ACT_SITES)
W_SHAPES = _expand(WEIGHTS)
assert len(ACT_SHAPES) == 71 and sum(math.prod(s) for s in ACT_SHAPES) == 16_784_872
assert len(W_SHAPES) == 54 and sum(math.prod(s) for s in W_SHAPES) == 25_502_912 and sum(s[0] for s in W_SHAPES) == 27_560


# -------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz, self.error = index, [], set(), None, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:    # NVML enumerates every GPU of the box; CUDA may see a subset / another order: go by PCI address
                import torch
                pr = torch.cuda.get_device_properties(index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception as exc:
            self.error = repr(exc)

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            self._sample()        # the GPU has just finished the timed region: a short run still gets a sample under load
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled" + (": " + self.error if self.error else "")]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU op on the host cores
# -------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = args.ref_batch
    ref_so = ROOT / "oracle" / "_ref" / "torchlsq" / "_C.so"
    kind = "reference" if ref_so.exists() else "port"
    gen = torch.Generator().manual_seed(0)
    acts = []
    for i, shp in enumerate(ACT_SHAPES):
        x = torch.randn(sample_batch, *shp, generator=gen)
        if i:
            x = x.relu_()
        acts.append((x, torch.randn(sample_batch, *shp, generator=gen)))
    wts = [(torch.randn(*shp, generator=gen) * 0.05, torch.randn(*shp, generator=gen)) for shp in W_SHAPES]
    n_el = sum(x.numel() for x, _ in acts) + sum(w.numel() for w, _ in wts)
    alg_bytes = 5 * 4 * n_el                                    # the CPU op is fp32-only

    if kind == "reference":
        sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
        import torchlsq as ref                                   # the UNMODIFIED reference package
        assert "oracle/_ref" in ref.__file__
        from torchlsq.functional import lsq
        a_params = [(torch.tensor([0.03], requires_grad=True), torch.tensor([0.0], requires_grad=True)) for _ in acts]
        w_params = [(torch.full((w.shape[0],), 0.002, requires_grad=True), torch.zeros(w.shape[0], requires_grad=True))
                    for w, _ in wts]

        def step():
            outs = []
            for (x, g), (s, b) in zip(acts, a_params):
                x.requires_grad_(True)
                outs.append((lsq(x, s, b, 0, 127, 0, 255, 1, True, 1.0, True, False, False, False), g))
            for (w, g), (s, b) in zip(wts, w_params):
                w.requires_grad_(True)
                outs.append((lsq(w, s, b, -128, 127, -128, 127, 0, True, 1.0, False, True, False, False), g))
            for y, g in reversed(outs):
                y.backward(g)
            for t in [x for x, _ in acts] + [w for w, _ in wts]:
                t.grad = None
    else:
        sys.path.insert(0, str(ROOT))
        from oracle import lsq_oracle as O
        ca = O.cfg(0, 127, 0, 255, contract=O.CONTRACT_CPU)
        cw = O.cfg(-128, 127, -128, 127, sym=True, contract=O.CONTRACT_CPU, numel_div_c=True)
        acts_np = [(x.numpy().reshape(-1), g.numpy().reshape(-1)) for x, g in acts]
        wts_np = [(w.numpy().reshape(-1), g.numpy().reshape(-1), w.shape[0]) for w, g in wts]

        def step():
            for x, g in acts_np:
                O.forward(x, [0.03], [0.0], ca)
            for w, g, C in wts_np:
                O.forward(w, [0.002] * C, [0.0] * C, cw, 1, C, w.size // C, True)
            for w, g, C in reversed(wts_np):
                O.backward(g, w, [0.002] * C, [0.0] * C, cw, 1, C, w.size // C, True)
            for x, g in reversed(acts_np):
                O.backward(g, x, [0.03], [0.0], ca)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = alg_bytes / dt / 1e9
    sample = (f"{len(ACT_SHAPES)} activation sites at batch {sample_batch} + all {len(W_SHAPES)} weights, fp32 "
              f"({n_el} elements/step), {'reference CPU op via torchlsq.functional.lsq + autograd' if kind == 'reference' else 'oracle port (OpenMP)'}")
    line = {"impl": "reference", "metric": "lsq_fwd_bwd_algorithmic_GBps", "value": round(val, 4), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "resnet50_qat_fakequant (BASELINE configs[4] site list), bounded CPU sample",
                       "batch_per_step": sample_batch, "elements_per_step": n_el},
            "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def run_b200(args):
    for p in (str(PKG), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces itself ("NCCL version ...") on the process's stdout when the communicator is created; the driver
        # wants ONE JSON line there, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    import torchlsq
    from torchlsq import _cabi
    from torchlsq.dp import FlatGradBuffer
    from torchlsq.functional import lsq
    from torchlsq.multi import LSQPlan, Site
    if not torchlsq.extension._HAS_OPS:
        raise SystemExit(f"native library missing: {torchlsq.extension.error_str}")
    lib = _cabi.load()
    B = args.batch_per_gpu
    BF16, F32 = _cabi.BF16, _cabi.F32

    # ---- buffers: every site owns x / y / g / gx; flat grad buffer for all (grad_scale, grad_shift)
    slots = [(f"act{i}", 1) for i in range(len(ACT_SHAPES))] + [(f"w{i}", s[0]) for i, s in enumerate(W_SHAPES)]
    flat = FlatGradBuffer(slots, dev)
    assert flat.numel == 2 * (71 + 27_560)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = []
    for i, shp in enumerate([] if args.strong else ACT_SHAPES):
        n = B * math.prod(shp)
        x = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_(0, 1, generator=gen)
        if i:
            x.relu_()
        g = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_(0, 1, generator=gen)
        y, gx = torch.empty_like(x), torch.empty_like(x)
        s = torch.tensor([0.03], device=dev)
        b = torch.tensor([0.0 if i else -1.9], device=dev)
        acts.append(dict(x=x, y=y, g=g, gx=gx, s=s, b=b, n=n, name=f"act{i}"))
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=dev)
    wsites = []
    for i, shp in enumerate(W_SHAPES):
        w = torch.empty(shp, dtype=torch.float32, device=dev).normal_(0, 0.05, generator=gen)
        g = torch.empty(shp, dtype=torch.float32, device=dev).normal_(0, 1, generator=gen)
        gs, gb = flat.views(f"w{i}")
        wsites.append(Site(x=w, y=torch.empty_like(w), grad=g, gx=torch.empty_like(w),
                           scale=torch.empty(shp[0], device=dev), shift=torch.zeros(shp[0], device=dev), gscale=gs, gshift=gb,
                           quant_min=-128, quant_max=127, type_min=-128, type_max=127, axis=0, is_affine=False,
                           is_perchannel=True))
    wplan = LSQPlan(wsites)
    # mu +- 3 sigma initialisation of all 27 560 weight scales: one launch (timed separately below)
    scales = wplan.weight_init_stats()
    off = 0
    for st in wsites:
        st.scale.copy_(scales[off:off + st.scale.numel()])
        off += st.scale.numel()
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    if args.strong:
        return run_strong_line(args, torch, dist, lib, flat, wplan, ws, sp, stream, dev, world, rank, local)
    qa = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    fwd_calls, bwd_calls = [], []
    for a in acts:
        gs, gb = flat.views(a["name"])
        fwd_calls.append((a["x"].data_ptr(), a["y"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(), a["n"], BF16, F32, qa, sp))
        bwd_calls.append((a["g"].data_ptr(), a["x"].data_ptr(), a["gx"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(),
                          gs.data_ptr(), gb.data_ptr(), a["n"], BF16, F32, qa, ws.data_ptr(), ws.numel(), sp))
    bwd_calls.reverse()
    fwd_t, bwd_t = lib.lsqb200_fwd_tensor, lib.lsqb200_bwd_tensor
    launches_per_step = len(fwd_calls) + len(bwd_calls) + wplan.launches(False) + wplan.launches(True)
    aplan = None
    if args.plan_activations:      # all 71 activation sites in one launch per direction (upper bound without launch gaps)
        asites = []
        for a in acts:
            gs, gb = flat.views(a["name"])
            asites.append(Site(x=a["x"], y=a["y"], grad=a["g"], gx=a["gx"], scale=a["s"], shift=a["b"], gscale=gs, gshift=gb,
                               quant_min=0, quant_max=127, type_min=0, type_max=255))
        aplan = LSQPlan(asites)
        launches_per_step = aplan.launches(False) + aplan.launches(True) + wplan.launches(False) + wplan.launches(True)

    pending = [None]      # handle of the flat-grad all-reduce in flight on the side stream (N > 1)

    def step(ev0=None, ev1=None):
        if aplan is not None:
            aplan.forward()
        else:
            for c in fwd_calls:
                fwd_t(*c)
        wplan.forward()
        if pending[0] is not None:      # the previous step's all-reduce still reads the flat buffer the backward overwrites
            pending[0].wait()
            pending[0] = None
        if ev0 is not None:
            ev0.record(stream)
        if aplan is not None:
            aplan.backward()
        else:
            for c in bwd_calls:
                bwd_t(*c)
        if ev1 is not None:
            ev1.record(stream)
        wplan.backward()
        if world > 1:
            # side stream: the collective waits for this step's backward and overlaps the next step's forward
            # (DESIGN.md section 5); the compute stream only joins it again in front of the next backward
            pending[0] = flat.all_reduce(side_stream=not args.inline_allreduce)

    def drain():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    n_act = sum(a["n"] for a in acts)
    n_w = sum(math.prod(s) for s in W_SHAPES)
    alg_bytes_rank = 5 * 2 * n_act + 5 * 4 * n_w
    bwd_bytes = 3 * 2 * n_act

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    drain()
    torch.cuda.synchronize()
    # sanity: the C ABI must have really run (outputs finite, grads written); recorded, never a reason to lose the line
    ran_ok = bool(torch.isfinite(flat.flat).all().item() and flat.flat.abs().sum().item() > 0)
    dp_check = None
    if world > 1:
        dp_check = run_dp_check(torch, dist, flat, world, fwd_calls, bwd_calls, fwd_t, bwd_t, wplan)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for k in range(args.steps):
            step(*bw[k])
        drain()                     # the last step's all-reduce belongs to the timed region
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = alg_bytes_rank * world / (ms_step * 1e-3) / 1e9
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in bw)
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9

    # ---- same step with ALL activation sites in one multi-tensor launch per direction: what is left when the
    #      per-launch ramp / tail (the difference between the per-site step and the kernels' own rate) is gone
    plan_mode = None
    if aplan is None and not args.no_plan_mode:
        asites = []
        for a in acts:
            gs_, gb_ = flat.views(a["name"])
            asites.append(Site(x=a["x"], y=a["y"], grad=a["g"], gx=a["gx"], scale=a["s"], shift=a["b"], gscale=gs_, gshift=gb_,
                               quant_min=0, quant_max=127, type_min=0, type_max=255))
        p2 = LSQPlan(asites)

        def step2():
            p2.forward(); wplan.forward()
            drain()
            p2.backward(); wplan.backward()
            if world > 1:
                pending[0] = flat.all_reduce(side_stream=not args.inline_allreduce)
        for _ in range(3):
            step2()
        drain()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(args.steps):
            step2()
        drain()
        p1.record(stream)
        barrier()
        tp = torch.tensor([p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pms = tp.item() / args.steps
        plan_mode = {"value": round(alg_bytes_rank * world / (pms * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(pms, 4),
                     "launches_per_step": p2.launches(False) + p2.launches(True) + wplan.launches(False) + wplan.launches(True),
                     "note": "torchlsq.multi.LSQPlan: all 71 activation sites in one launch per direction"}
        p2.close()

    # ---- prologue fusion at workload level: the same 71 + 54 sites, but 33 sites sit behind a ReLU and 16 behind
    #      a residual add + ReLU (torchvision resnet50: conv1 + 2 per block; 1 per block).  "unfused" runs those ops as the
    #      network does today - ATen add / relu passes in front of the plain kernels, threshold_backward behind them -,
    #      "fused" runs them inside the fake-quant kernels (lsqb200_*_pre).  Same results bit for bit (tests/test_gpu_relu.py,
    #      tests/test_gpu_add.py); the figure is ms per step over ALL sites, N = 1 only.
    fusion_mode = None
    if world == 1 and not args.no_fusion_mode:
        try:
            fusion_mode = run_fusion_mode(args, torch, lib, acts, flat, ws, sp, stream, wplan, qa, BF16, F32, dev, gen)
        except Exception as exc:       # an extra measurement, never a reason to lose the line
            fusion_mode = {"error": repr(exc)}

    # ---- every BASELINE config as its own measurement (L2 flushed between iterations where the working set is small)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = peaks.get("hbm_gbs", 6650.0)
    configs = None
    if world == 1 and not args.no_configs:
        try:
            configs = run_configs(args, torch, lib, acts, wplan, ws, sp, stream, scales, dev, peak)
        except Exception as exc:
            configs = {"error": repr(exc)}
    strong = None
    if world == 1 and not args.no_strong:
        try:
            strong = run_strong(args, torch, dist, lib, flat, wplan, ws, sp, stream, dev, 1, peak)
        except Exception as exc:
            strong = {"error": repr(exc)}

    # ---- the same step through the PUBLIC API on device-resident tensors (the drop-in a user calls): functional op +
    #      autograd, and LSQFakeQuantizer modules
    api_mode = None
    if world == 1 and not args.no_api_mode:
        try:
            api_mode = run_api_mode(args, torch, acts, wsites, stream, ms_step)
        except Exception as exc:       # a report, never a reason to lose the line
            api_mode = {"error": repr(exc)}

    # ---- end to end through the public op with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, torch, lsq, dev, B, world, dist, flat, wsites)
        except Exception as exc:
            if world > 1:              # a rank that dropped out of the e2e collectives would hang the others: fail loudly instead
                raise
            e2e = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(exc)}

    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    traffic = None
    prof = ROOT / "profiles" / "ncu_summary.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("bwd_bf16_largest_site", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                                     capture_output=True, text=True, timeout=900)
                ref_line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
                cpu_baseline = ref_line["cpu_baseline"]
            except Exception as exc:   # the baseline is a report, never a reason to lose the GPU number
                cpu_baseline = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {exc}"}
        line = {
            "metric": "lsq_fwd_bwd_algorithmic_GBps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "resnet50_qat_fakequant: BASELINE configs[4] per-rank shard (71 bf16 per-tensor activation sites "
                                   "+ 54 fp32 per-channel weights, fwd+bwd, flat-grad all-reduce when N>1)",
                       "storage": "bf16 activations, f32 weights and parameters; arithmetic in f32, reductions in f64",
                       "batch_per_gpu": B, "global_batch": B * world, "elements_per_gpu_step": n_act + n_w,
                       "algorithmic_bytes_per_gpu_step": alg_bytes_rank,
                       "l2": "every site has its own x/y/g/gx buffers (%.1f GB resident) - far larger than the 126 MB L2, no flush needed" % (4 * 2 * n_act / 1e9),
                       "launch": ("ONE multi-tensor plan launch per direction for all activation sites (--plan-activations)" if aplan is not None
                                  else "per-site C-ABI calls for activations") + ", one multi-tensor plan launch per class for all weights",
                       "parallelism": f"dp{world}"},
            "per_gpu_GBps": round(value / world, 1), "pct_of_hbm_peak_per_gpu": round(100 * value / world / peak, 2),
            "pct_of_8TBps_spec": round(100 * value / world / 8000.0, 2),
            "roofline": {"bound": "hbm", "kernel": "lsq_flatbwd_kernel<bf16> (per-tensor backward, 71 launches/step)",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": bwd_bytes, "avg_ms_per_step": round(bwd_ms, 4)},
            "api_mode": api_mode, "plan_mode": plan_mode, "prologue_fusion": fusion_mode, "configs": configs, "strong_batch2048": strong, "dp_check": dp_check, "kernels_ran": ran_ok,
            "allreduce": (None if world == 1 else ("inline on the compute stream" if args.inline_allreduce else
                                                   "side stream, overlapped with the next step's forward; joined in front of the next backward and before the final event")),
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_dp_check(torch, dist, flat, world, fwd_calls, bwd_calls, fwd_t, bwd_t, wplan):
    """Data-parallel semantics on the real NCCL path: the all-reduced flat buffer equals the sum over ranks of the
    local grad_scale / grad_shift (each rank scaled with its LOCAL numel as the reference op would be under DDP,
    /root/reference/torchlsq/csrc/ops/cuda/lsq_cuda.cu:124,274).

    Bound per element: an fp32 sum of W addends in any order is within (W-1) * 2^-24 * sum_r |shard_r| of the exact sum
    (relative to the result it is unbounded once the shards cancel), so
        |got - want| <= 1e-6 * |want| + W * 2^-24 * sum_r |shard_r|.
    The outcome is RECORDED (`ok`); it never aborts the bench.  tests/test_gpu_dp_nccl.py is the correctness evidence."""
    try:
        for c in fwd_calls:
            fwd_t(*c)
        wplan.forward()
        for c in bwd_calls:
            bwd_t(*c)
        wplan.backward()
        local = flat.flat.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        flat.all_reduce()
        torch.cuda.synchronize()
        from torchlsq.dp import allreduce_bound
        want, bound = allreduce_bound(gathered)
        err = (flat.flat.double() - want).abs()
        ratio = (err / bound.clamp_min(1e-300)).max().item()
        ok = torch.tensor([1 if (ratio <= 1.0 and bool(torch.isfinite(flat.flat).all().item())) else 0], device=local.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return {"ok": bool(ok.item()), "max_err_over_bound": ratio, "max_abs_err": err.max().item(), "floats": flat.numel,
                "world": world, "bound": "1e-6*|sum| + W*2^-24*sum_r|shard_r| per element (fp32 sum of W addends)",
                "semantics": "sum over ranks of local-numel-scaled grads (lsq_cuda.cu:124,274)"}
    except Exception as exc:            # report, never kill the line
        return {"ok": False, "error": repr(exc)}


def _timed_rotating(torch, fns, iters, stream, cover_us=60):
    """Per-iteration CUDA-event timing over ROTATING buffer sets: iteration i runs fns[i % len(fns)], each on its own
    tensors, the sets together well above the 126 MB L2 - "inputs larger than L2": nothing an iteration reads is left in L2
    by an earlier one, and what an iteration leaves dirty in L2 is written back during the next (every byte crosses HBM
    exactly once in steady state; no flush kernel, whose own dirty lines would be written back inside the timed region).
    A spin kernel without memory traffic (torch.cuda._sleep) runs in front of every iteration so that the host has
    enqueued the iteration's launches before the GPU reaches the first event: host latency is not counted.
    Returns (median ms, best ms, back-to-back ms per iteration: all sets in a row between two events, no gaps)."""
    spin = int(cover_us * 1.9e3)           # cycles at ~1.9 GHz
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = len(fns)
    for i in range(iters + 2 * n):
        torch.cuda._sleep(spin)
        e0.record(stream)
        fns[i % n]()
        e1.record(stream)
        e1.synchronize()
        if i >= 2 * n:
            ts.append(e0.elapsed_time(e1))
    reps = max(1, iters // n)
    torch.cuda._sleep(spin)
    e0.record(stream)
    for _ in range(reps):
        for f in fns:
            f()
    e1.record(stream)
    e1.synchronize()
    return statistics.median(ts), min(ts), e0.elapsed_time(e1) / (reps * n)


def run_configs(args, torch, lib, acts, wplan, ws, sp, stream, scales, dev, peak):
    """BASELINE.json configs[0..3] (configs[4] is the headline / `strong_batch2048`), plus the short-row per-channel layouts and
    the observer step.  GB/s = algorithmic bytes / median time of one isolated iteration (see _timed_rotating): 5*sizeof(T) per
    element for fwd+bwd, 1*sizeof(T) for statistics / observer; `GBps_stream` = the same work back to back."""
    from torchlsq import _cabi
    from torchlsq.multi import LSQPlan, Site
    iters = max(8, min(args.steps, 20))
    out = {"protocol": "per-iteration CUDA events over rotating buffer sets (total > 2x L2), spin-kernel cover for host latency, "
                       "no flush kernel; median of %d" % iters}

    def report(name, nbytes, timing, **extra):
        med, best, stream_ms = timing
        out[name] = dict(ms=round(med, 4), ms_best=round(best, 4), GBps=round(nbytes / med / 1e6, 1),
                         frac=round(nbytes / med / 1e6 / peak, 4), GBps_stream=round(nbytes / stream_ms / 1e6, 1),
                         frac_stream=round(nbytes / stream_ms / 1e6 / peak, 4), **extra)

    # config 1: per-tensor quint8 fp32 32x64x56x56 (seeds 1 / 2, scale .03, shift -1.7); 4 sets of 103 MB
    R = 4
    x0 = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(1)).to(dev)
    g0 = torch.randn(32, 64, 56, 56, generator=torch.Generator().manual_seed(2)).to(dev)
    s, b = torch.tensor([0.03], device=dev), torch.tensor([-1.7], device=dev)
    gs, gb = torch.empty(1, device=dev), torch.empty(1, device=dev)
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    n = x0.numel()
    sets = [(x0.clone(), g0.clone(), torch.empty_like(x0), torch.empty_like(x0)) for _ in range(R)]

    def c1(k):
        x, g, y, gx = sets[k]
        return lambda: (lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 0, 0, q, sp),
                        lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                               gb.data_ptr(), n, 0, 0, q, ws.data_ptr(), ws.numel(), sp))
    report("config1_per_tensor_fp32_32x64x56x56", 5 * 4 * n, _timed_rotating(torch, [c1(k) for k in range(R)], iters, stream),
           sets=R, launches=2)

    def c1_ceiling(k):
        x, g, y, gx = sets[k]
        return lambda: (y.copy_(x), torch.add(x, g, out=gx))
    report("config1_ceiling_aten_copy_plus_add", 5 * 4 * n, _timed_rotating(torch, [c1_ceiling(k) for k in range(R)], iters, stream),
           sets=R, launches=2, note="the same bytes through ATen's own streaming kernels (copy R+W, add 2R+W): what two launches of this size reach")
    del sets
    # config 2: the 54 ResNet-50 weights, per-channel axis 0: the bench's plan + 3 more sets (408 MB each)
    nw = sum(math.prod(sh) for sh in W_SHAPES)
    plans, outs = [wplan], [scales]
    gen = torch.Generator(device=dev).manual_seed(99)
    for _ in range(R - 1):
        st_ = []
        for shp in W_SHAPES:
            w = torch.empty(shp, device=dev).normal_(0, 0.05, generator=gen)
            st_.append(Site(x=w, y=torch.empty_like(w), grad=torch.empty(shp, device=dev).normal_(generator=gen), gx=torch.empty_like(w),
                            scale=torch.full((shp[0],), 0.002, device=dev), shift=torch.zeros(shp[0], device=dev),
                            gscale=torch.empty(shp[0], device=dev), gshift=torch.empty(shp[0], device=dev), quant_min=-128, quant_max=127,
                            type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True))
        plans.append(LSQPlan(st_))
        outs.append(torch.empty_like(scales))
    report("config2_weight_init_mu3sigma_54_weights", 4 * nw,
           _timed_rotating(torch, [(lambda p=p, o=o: p.weight_init_stats(o)) for p, o in zip(plans, outs)], iters, stream), sets=R, launches=1)
    flats = [torch.empty(nw, device=dev).normal_() for _ in range(R)]
    report("config2_weight_init_ceiling_aten_flat_sum", 4 * nw, _timed_rotating(torch, [(lambda f=f: f.sum()) for f in flats], iters, stream),
           sets=R, note="one flat ATen reduction over the same number of bytes: no rows, no per-channel results")
    del flats
    report("config2_weights_fwd_bwd_54_weights", 5 * 4 * nw,
           _timed_rotating(torch, [(lambda p=p: (p.forward(), p.backward())) for p in plans], iters, stream), sets=R,
           launches=wplan.launches(False) + wplan.launches(True))
    for p in plans[1:]:
        p.close()
    del plans, outs
    # config 3: learned init (init_mode) on all 71 bf16 sites at the batch of this run (34 GB resident: one set)
    qi = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, True)
    gsl, gbl = torch.empty(1, device=dev), torch.empty(1, device=dev)
    f3 = [(a["x"].data_ptr(), a["y"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(), a["n"], 2, 0, qi, sp) for a in acts]
    b3 = [(a["g"].data_ptr(), a["x"].data_ptr(), a["gx"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(), gsl.data_ptr(), gbl.data_ptr(),
           a["n"], 2, 0, qi, ws.data_ptr(), ws.numel(), sp) for a in reversed(acts)]

    def c3():
        for c in f3:
            lib.lsqb200_fwd_tensor(*c)
        for c in b3:
            lib.lsqb200_bwd_tensor(*c)
    n3 = sum(a["n"] for a in acts)
    report("config3_learned_init_bf16_71_sites", 5 * 2 * n3, _timed_rotating(torch, [c3], max(3, iters // 4), stream),
           sets=1, resident_GB=round(4 * 2 * n3 / 1e9, 1), launches=2 * len(acts))
    # config 4: per-channel axis 1, fp16 256x1024x28x28 (1.6 GB per set), fp32 parameters, grad scaling on; and the short-row layouts
    N = 256
    x4 = torch.empty(N * 1024 * 784, dtype=torch.float16, device=dev).normal_()
    g4 = torch.empty_like(x4).normal_()
    y4, gx4 = torch.empty_like(x4), torch.empty_like(x4)
    for name, outer, cc, hw in (("config4_per_channel_fp16_256x1024x28x28", N, 1024, 784), ("per_channel_fp16_256x1024x14x14", N, 1024, 196),
                                ("per_channel_fp16_256x2048x7x7", N, 2048, 49), ("per_channel_fp16_channels_last_50176x1024", N * 196, 1024, 1)):
        sc, bc = 0.02 + 0.02 * torch.rand(cc, device=dev), -torch.rand(cc, device=dev)
        gsc, gbc = torch.empty(cc, device=dev), torch.empty(cc, device=dev)
        nn_ = outer * cc * hw
        nset = max(1, (N * 1024 * 784) // nn_)          # rotate through disjoint slices of the 411 MB buffers

        def c4(k):
            o = k * nn_ * 2
            return lambda: (lib.lsqb200_fwd_channel(x4.data_ptr() + o, y4.data_ptr() + o, sc.data_ptr(), bc.data_ptr(), outer, cc, hw, 1, 0, q, sp),
                            lib.lsqb200_bwd_channel(g4.data_ptr() + o, x4.data_ptr() + o, gx4.data_ptr() + o, sc.data_ptr(), bc.data_ptr(),
                                                    gsc.data_ptr(), gbc.data_ptr(), outer, cc, hw, 1, 0, q, ws.data_ptr(), ws.numel(), sp))
        report(name, 5 * 2 * nn_, _timed_rotating(torch, [c4(k) for k in range(nset)], iters, stream), sets=nset, launches=2)
    # dtype matrix of the per-tensor kernels on one large site (205 MB per buffer, 1 GB working set: far above L2): fp32, fp16 and
    # bf16 tensors with fp32 parameters, and the reference-exact all-fp16 contract (c10::Half: a rounding after every operator)
    nbig = 256 * 128 * 56 * 56
    for name, tdt, code, pcode in (("per_tensor_205MB_fp32", torch.float32, 0, 0), ("per_tensor_205MB_fp16_fp32params", torch.float16, 1, 0),
                                   ("per_tensor_205MB_fp16_exact_fp16params", torch.float16, 1, 1), ("per_tensor_205MB_bf16_fp32params", torch.bfloat16, 2, 0)):
        es = 4 if tdt == torch.float32 else 2
        m = nbig // (2 if es == 4 else 1)             # the four fp16 buffers hold 205.5 M halves = 102.8 M floats
        xb, gb_, yb, gxb = (t.view(tdt)[:m] for t in (x4, g4, y4, gx4))
        pdt = torch.float16 if pcode == 1 else torch.float32
        sp_, bp_ = torch.tensor([0.03], device=dev, dtype=pdt), torch.tensor([-1.7], device=dev, dtype=pdt)
        gsp, gbp = torch.empty(1, device=dev, dtype=pdt), torch.empty(1, device=dev, dtype=pdt)
        qd = _cabi.qargs(0, 127, 0, 255, code != 1 or pcode != 1, 1.0, False, False, False)      # all-fp16: grad scaling off

        def cd():
            lib.lsqb200_fwd_tensor(xb.data_ptr(), yb.data_ptr(), sp_.data_ptr(), bp_.data_ptr(), m, code, pcode, qd, sp)
            lib.lsqb200_bwd_tensor(gb_.data_ptr(), xb.data_ptr(), gxb.data_ptr(), sp_.data_ptr(), bp_.data_ptr(), gsp.data_ptr(), gbp.data_ptr(),
                                   m, code, pcode, qd, ws.data_ptr(), ws.numel(), sp)
        report(name, 5 * es * m, _timed_rotating(torch, [cd], max(5, iters // 2), stream), sets=1, launches=2, elements=m)
    # observer-mode init step on a 411 MB bf16 activation: x4 / g4 / y4 / gx4 as four inputs
    import warnings
    from torchlsq.quantized.modules.observers import observer_step
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        obs = torch.quantization.MovingAverageMinMaxObserver(reduce_range=True).to(dev)
    so, bo = torch.ones(1, device=dev), torch.zeros(1, device=dev)
    xo = [t.view(torch.bfloat16) for t in (x4, g4, y4.normal_(), gx4.normal_())]
    report("observer_step_bf16_411MB", 2 * xo[0].numel(), _timed_rotating(torch, [(lambda t=t: observer_step(obs, t, so, bo)) for t in xo], iters, stream),
           sets=4, launches=1)
    report("observer_step_ceiling_aten_amax_411MB", 2 * xo[0].numel(), _timed_rotating(torch, [(lambda t=t: t.amax()) for t in xo], iters, stream),
           sets=4, note="ATen's own one-output reduction over the same bytes")
    return out


def run_strong(args, torch, dist, lib, flat, wplan, ws, sp, stream, dev, world, peak):
    """BASELINE configs[4] as written: global batch 2048 split over `world` ranks, 71 bf16 activation sites + the 54
    weights, forward of every site then backward of every site, flat-grad all-reduce when world > 1.  At world = 1 the largest
    site is 2048x64x112x112 = 1.64 G elements (3.3 GB per buffer) and all sites together would need 275 GB, so the sites run one
    at a time through four arenas (x, y, grad, grad_x) of 2x the largest site each, placed ring-fashion so that no site re-reads
    what a recent one left in L2."""
    from torchlsq import _cabi
    B = 2048 // world
    sizes = [B * math.prod(shp) for shp in ACT_SHAPES]
    arena = 2 * max(sizes)
    bufs = {}
    for name in ("x", "g", "y", "gx"):
        bufs[name] = torch.empty(arena, dtype=torch.bfloat16, device=dev)
    chunk = 1 << 28
    for o in range(0, arena, chunk):
        bufs["x"][o:o + chunk].normal_().relu_()
        bufs["g"][o:o + chunk].normal_()
    esz = 2
    base = {k: v.data_ptr() for k, v in bufs.items()}
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    s, b = torch.tensor([0.03], device=dev), torch.tensor([0.0], device=dev)
    offs, off = [], 0
    for n in sizes:
        if off + n > arena:
            off = 0
        offs.append(off)
        off += -(-n // 256) * 256
    fcalls, bcalls = [], []
    for i, (n, o) in enumerate(zip(sizes, offs)):
        gs_, gb_ = flat.views(f"act{i}")
        fcalls.append((base["x"] + o * esz, base["y"] + o * esz, s.data_ptr(), b.data_ptr(), n, _cabi.BF16, _cabi.F32, q, sp))
        bcalls.append((base["g"] + o * esz, base["x"] + o * esz, base["gx"] + o * esz, s.data_ptr(), b.data_ptr(), gs_.data_ptr(),
                       gb_.data_ptr(), n, _cabi.BF16, _cabi.F32, q, ws.data_ptr(), ws.numel(), sp))
    bcalls.reverse()
    pending = [None]

    def step(ev0=None, ev1=None):
        for c in fcalls:
            lib.lsqb200_fwd_tensor(*c)
        wplan.forward()
        if pending[0] is not None:
            pending[0].wait(); pending[0] = None
        if ev0 is not None:
            ev0.record(stream)
        for c in bcalls:
            lib.lsqb200_bwd_tensor(*c)
        if ev1 is not None:
            ev1.record(stream)
        wplan.backward()
        if world > 1:
            pending[0] = flat.all_reduce(side_stream=not args.inline_allreduce)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(2, min(args.steps, 5 * world))
    for _ in range(3):
        step()
    if pending[0] is not None:
        pending[0].wait(); pending[0] = None
    barrier()
    bw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        step(*bw[k])
    if pending[0] is not None:
        pending[0].wait(); pending[0] = None
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    n_act, n_w = sum(sizes), sum(math.prod(sh) for sh in W_SHAPES)
    alg_rank = 5 * 2 * n_act + 5 * 4 * n_w
    bwd_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in bw)
    dp_check = None
    if world > 1:
        dp_check = run_dp_check(torch, dist, flat, world, fcalls, bcalls, lib.lsqb200_fwd_tensor, lib.lsqb200_bwd_tensor, wplan)
    return {"dp_check": dp_check, "value": round(alg_rank * world / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(ms, 4), "steps": steps,
            "global_batch": 2048, "batch_per_gpu": B, "n_gpus": world, "per_gpu_GBps": round(alg_rank / (ms * 1e-3) / 1e9, 1),
            "frac_of_measured_peak_per_gpu": round(alg_rank / (ms * 1e-3) / 1e9 / peak, 4),
            "algorithmic_bytes_per_gpu_step": alg_rank, "elements_per_gpu_step": n_act + n_w,
            "bwd_bf16_achieved_GBps": round(3 * 2 * n_act / (bwd_ms * 1e-3) / 1e9, 1), "bwd_ms_per_step": round(bwd_ms, 4),
            "launches_per_step": 2 * len(sizes) + wplan.launches(False) + wplan.launches(True),
            "arena": "x / y / grad / grad_x arenas of %.1f GB each (2x the largest site), sites placed ring-fashion" % (arena * esz / 1e9)}


def run_strong_line(args, torch, dist, lib, flat, wplan, ws, sp, stream, dev, world, rank, local):
    """`--strong`: the JSON line for BASELINE configs[4] as written (global batch 2048 over N ranks, strong scaling)."""
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = peaks.get("hbm_gbs", 6650.0)
    with ClockSampler(local) as clk:
        r = run_strong(args, torch, dist, lib, flat, wplan, ws, sp, stream, dev, world, peak)
    if rank == 0:
        line = {"metric": "lsq_fwd_bwd_algorithmic_GBps", "value": r["value"], "unit": "GB/s", "n_gpus": world, "steps": r["steps"],
                "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "resnet50_qat_fakequant: BASELINE configs[4] as written - global batch 2048 split over the ranks "
                                       "(71 bf16 per-tensor activation sites + 54 fp32 per-channel weights, fwd+bwd, flat-grad all-reduce when N>1)",
                           "global_batch": 2048, "batch_per_gpu": r["batch_per_gpu"], "l2": r["arena"], "parallelism": f"dp{world}"},
                "per_gpu_GBps": r["per_gpu_GBps"], "pct_of_hbm_peak_per_gpu": round(100 * r["frac_of_measured_peak_per_gpu"], 2),
                "roofline": {"bound": "hbm", "kernel": "lsq_flatbwd_kernel<bf16> (per-tensor backward, 71 launches/step)",
                             "achieved": r["bwd_bf16_achieved_GBps"], "peak": peak, "unit": "GB/s",
                             "frac": round(r["bwd_bf16_achieved_GBps"] / peak, 4), "traffic": None},
                "dp_check": r["dp_check"], "cpu_baseline": None, "e2e": None, "gpu_launches": r["launches_per_step"] * r["steps"],
                "clocks": clk.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# per-image shapes of the sites that sit behind a ReLU / behind the residual add + ReLU of a block (torchvision resnet50)
RELU_SITES = {(64, 112, 112): 1, (64, 56, 56): 6, (128, 56, 56): 1, (128, 28, 28): 7, (256, 28, 28): 1, (256, 14, 14): 11,
              (512, 14, 14): 1, (512, 7, 7): 5}
JOIN_SITES = {(256, 56, 56): 3, (512, 28, 28): 4, (1024, 14, 14): 6, (2048, 7, 7): 3}


def run_fusion_mode(args, torch, lib, acts, flat, ws, sp, stream, wplan, qa, BF16, F32, dev, gen):
    relu_left, join_left = dict(RELU_SITES), dict(JOIN_SITES)
    kinds = []
    for a, shp in zip(acts, ACT_SHAPES):
        if relu_left.get(shp, 0) > 0:
            relu_left[shp] -= 1
            kinds.append("relu")
        elif join_left.get(shp, 0) > 0:
            join_left[shp] -= 1
            kinds.append("join")
        else:
            kinds.append("plain")
    assert kinds.count("relu") == 33 and kinds.count("join") == 16
    fused_f, fused_b, unf_f, unf_b = [], [], [], []
    for a, kind in zip(acts, kinds):
        gs, gb = flat.views(a["name"])
        x, y, g, gx, s, b, n = a["x"], a["y"], a["g"], a["gx"], a["s"], a["b"], a["n"]
        plain_f = (lib.lsqb200_fwd_tensor, (x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, BF16, F32, qa, sp))
        plain_b = (lib.lsqb200_bwd_tensor, (g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                            gb.data_ptr(), n, BF16, F32, qa, ws.data_ptr(), ws.numel(), sp))
        if kind == "plain":
            fused_f.append(plain_f); fused_b.append(plain_b); unf_f.append(plain_f); unf_b.append(plain_b)
            continue
        # pre-activation input(s): zero-mean so that the ReLU has work to do; t = the activation tensor the unfused network holds
        pre = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_(0, 1, generator=gen)
        t, gt = torch.empty_like(pre), torch.empty_like(pre)
        pre2 = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_(0, 1, generator=gen) if kind == "join" else None
        code = 2 if kind == "join" else 1
        p2 = pre2.data_ptr() if pre2 is not None else None
        fused_f.append((lib.lsqb200_fwd_tensor_pre, (pre.data_ptr(), p2, y.data_ptr(), s.data_ptr(), b.data_ptr(), n, BF16, F32, qa, code, sp)))
        fused_b.append((lib.lsqb200_bwd_tensor_pre, (g.data_ptr(), pre.data_ptr(), p2, gx.data_ptr(), s.data_ptr(), b.data_ptr(),
                                                     gs.data_ptr(), gb.data_ptr(), n, BF16, F32, qa, code, ws.data_ptr(), ws.numel(), sp)))
        if kind == "join":      # torchvision: out += identity; out = relu_(out)
            unf_f.append((lambda pre=pre, pre2=pre2, t=t: (torch.add(pre, pre2, out=t), t.relu_()), ()))
        else:                   # relu (out of place so that the step is repeatable; same traffic as relu_: R + W)
            unf_f.append((lambda pre=pre, t=t: torch.clamp_min(pre, 0, out=t), ()))
        unf_f.append((lib.lsqb200_fwd_tensor, (t.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, BF16, F32, qa, sp)))
        unf_b.append((lib.lsqb200_bwd_tensor, (g.data_ptr(), t.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                               gb.data_ptr(), n, BF16, F32, qa, ws.data_ptr(), ws.numel(), sp)))
        unf_b.append((lambda gx=gx, t=t, gt=gt: torch.ops.aten.threshold_backward(gx, t, 0, grad_input=gt), ()))
    fused_b.reverse(); unf_b.reverse()

    def run(fcalls, bcalls):
        for fn, a_ in fcalls:
            fn(*a_)
        wplan.forward()
        for fn, a_ in bcalls:
            fn(*a_)
        wplan.backward()

    out = {}
    for name, fc, bc in (("unfused", unf_f, unf_b), ("fused", fused_f, fused_b)):
        for _ in range(3):
            run(fc, bc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            run(fc, bc)
        e1.record(stream)
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / args.steps
    n_relu = sum(a["n"] for a, k in zip(acts, kinds) if k == "relu")
    n_join = sum(a["n"] for a, k in zip(acts, kinds) if k == "join")
    return {"unfused_ms_per_step": round(out["unfused"], 4), "fused_ms_per_step": round(out["fused"], 4),
            "speedup": round(out["unfused"] / out["fused"], 3),
            "sites": {"relu_then_fq": 33, "add_relu_then_fq": 16, "plain": 22, "weights": 54},
            "hbm_bytes_removed_per_step": 2 * (5 * n_relu + 6 * n_join),
            "note": "all 71 + 54 sites fwd+bwd; unfused = ATen clamp_min / add + relu_ / threshold_backward passes around the plain "
                    "kernels (as torchvision's resnet50 runs them), fused = lsqb200_*_pre with LSQB200_PRE_RELU / PRE_ADD_RELU"}


def run_api_mode(args, torch, acts, wsites, stream, cabi_ms):
    """The timed step driven the way a user drives it: `torchlsq.functional.lsq` + autograd, and `LSQFakeQuantizer`
    modules, on the same device-resident tensors as the C-ABI step (reference call path:
    /root/reference/torchlsq/quantized/modules/observers.py:458-461 -> functional.py:95-97 -> csrc/ops/lsq.cpp:104-134 ->
    csrc/ops/autograd/lsq_autograd.cpp:16-74).  Sites are issued in network order (weight quantizer, then the activation
    quantizer behind it), one `torch.autograd.backward` per step, grads dropped (set_to_none) after each step."""
    from torchlsq import LSQFakeQuantizer
    from torchlsq.functional import lsq
    dev = acts[0]["x"].device
    order = []                                # ("a", i) / ("w", j) in network order
    wi = 0
    for i in range(len(acts)):
        order.append(("a", i))
        if wi < len(wsites) and i < len(acts) - 1:
            order.append(("w", wi)); wi += 1
    assert wi == len(wsites)
    xs = [a["x"].detach().requires_grad_(True) for a in acts]
    gs = [a["g"] for a in acts]
    a_par = [(a["s"].clone().requires_grad_(True), a["b"].clone().requires_grad_(True)) for a in acts]
    w_x = [st.x.detach().requires_grad_(True) for st in wsites]
    w_par = [(st.scale.clone().requires_grad_(True), st.shift.clone().requires_grad_(True)) for st in wsites]
    leaves = xs + w_x + [t for p in a_par + w_par for t in p]

    def fn_step():
        ys, gg = [], []
        for kind, i in order:
            if kind == "a":
                ys.append(lsq(xs[i], a_par[i][0], a_par[i][1], 0, 127, 0, 255)); gg.append(gs[i])
            else:
                ys.append(lsq(w_x[i], w_par[i][0], w_par[i][1], -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True))
                gg.append(wsites[i].grad)
        torch.autograd.backward(ys, gg)
        for t in leaves:
            t.grad = None

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a_mod = [LSQFakeQuantizer(None, 'activation', dtype=torch.quint8, qscheme=torch.per_tensor_affine, init_mode='learnable',
                                  init_batches=0, init_scale=0.03).to(dev) for _ in acts]
        w_mod = [LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable',
                                  avoid_torch_overflow=False).to(dev) for _ in wsites]
    for m in a_mod + w_mod:
        m.train()

    def mod_step():
        ys, gg = [], []
        for kind, i in order:
            if kind == "a":
                ys.append(a_mod[i](xs[i])); gg.append(gs[i])
            else:
                ys.append(w_mod[i](w_x[i])); gg.append(wsites[i].grad)
        torch.autograd.backward(ys, gg)
        for t in xs + w_x:
            t.grad = None
        for m in a_mod + w_mod:
            m.scale.grad = None
            m.shift.grad = None

    for _ in range(2):           # first module call only creates the parameters
        with torch.no_grad():
            for kind, i in order:
                (a_mod[i](xs[i]) if kind == "a" else w_mod[i](w_x[i]))

    # the same with every weight quantizer behind ONE multi-tensor launch per direction (torchlsq.multi): LSQGroup for the
    # functional API, group_weight_quantizers(model) for modules laid out as prepare_qat leaves them
    from torchlsq.multi import LSQGroup, group_weight_quantizers
    w_xp = [torch.nn.Parameter(t.detach()) for t in w_x]
    fgroup = LSQGroup(w_xp, [p[0] for p in w_par], [p[1] for p in w_par], -128, 127, -128, 127, axis=0, is_affine=False,
                      is_perchannel=True)

    def fn_group_step():
        yw = fgroup()
        ys, gg = [], []
        for kind, i in order:
            if kind == "a":
                ys.append(lsq(xs[i], a_par[i][0], a_par[i][1], 0, 127, 0, 255)); gg.append(gs[i])
            else:
                ys.append(yw[i]); gg.append(wsites[i].grad)
        torch.autograd.backward(ys, gg)
        for t in leaves + w_xp:
            t.grad = None

    class QatLayer(torch.nn.Module):          # what torch.ao.nn.qat.Conv2d / Linear hold: .weight and .weight_fake_quant
        def __init__(self, w, q):
            super().__init__()
            self.weight, self.weight_fake_quant = w, q

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList([QatLayer(w, q) for w, q in zip(w_xp, w_mod)])
            self.acts = torch.nn.ModuleList(a_mod)

        def forward(self, inputs):
            ys = []
            for kind, i in order:
                if kind == "a":
                    ys.append(self.acts[i](inputs[i]))
                else:
                    layer = self.layers[i]
                    ys.append(layer.weight_fake_quant(layer.weight))
            return ys
    net = Net().train()
    group_weight_quantizers(net)
    gg_net = [gs[i] if kind == "a" else wsites[i].grad for kind, i in order]

    def mod_group_step():
        torch.autograd.backward(net(xs), gg_net)
        for t in xs + w_xp:
            t.grad = None
        for m in a_mod + w_mod:
            m.scale.grad = None
            m.shift.grad = None

    out = {}
    n_sites = len(order)
    steps = max(3, min(args.steps, 10))
    for name, fn in (("functional", fn_step), ("module", mod_step), ("functional_weights_grouped", fn_group_step),
                     ("module_weights_grouped", mod_group_step)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = 0.0
        e0.record(stream)
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            host += time.perf_counter() - t0
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": round(ms, 4), "vs_cabi_step": round(ms / cabi_ms, 4),
                     "host_ms_per_step": round(host / steps * 1e3, 4), "host_us_per_site": round(host / steps / n_sites * 1e6, 2)}
    out["sites"] = n_sites
    out["note"] = ("same 71 + 54 sites and device buffers as `value`; functional = torchlsq.functional.lsq + one torch.autograd.backward "
                   "per step; module = LSQFakeQuantizer.forward per site (steady state, parameters learning); *_weights_grouped = the 54 weight "
                   "quantizers behind one autograd node and one launch per direction (torchlsq.multi.LSQGroup / group_weight_quantizers); host_us_per_site = "
                   "host issue time of a whole fwd+bwd step / sites (the GPU runs behind it)")
    return out


def run_e2e(args, torch, lsq, dev, B, world, dist, flat, wsites):
    """Host buffers -> public op (torchlsq.functional.lsq + autograd) -> host buffers.

    Three-stage pipeline over the 71 activation sites: a copy-in stream (pinned x, g -> device), the
    compute stream (the public op's forward and autograd backward) and a copy-out stream (y, grad_x ->
    pinned host), chained with events over a ring of device buffer sets, so both PCIe directions stay
    busy while the kernels run.  Staging buffers are sized for the largest site and reused; the bytes
    moved per step are the full workload's.  Weights live on the device (as in training); their
    gradients and every site's grad_scale / grad_shift are read back."""
    steps = max(1, min(args.steps, args.e2e_steps))
    nmax = B * max(math.prod(s) for s in ACT_SHAPES)
    NB = 3
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    xh = torch.empty(nmax, dtype=torch.bfloat16).pin_memory().normal_().relu_()
    gh = torch.empty(nmax, dtype=torch.bfloat16).pin_memory().normal_()
    yh = torch.empty(nmax, dtype=torch.bfloat16).pin_memory()
    gxh = torch.empty(nmax, dtype=torch.bfloat16).pin_memory()
    ring = [dict(xd=torch.empty(nmax, dtype=torch.bfloat16, device=dev), gd=torch.empty(nmax, dtype=torch.bfloat16, device=dev),
                 yd=None, gxd=None, in_done=torch.cuda.Event(), cmp_done=torch.cuda.Event(), out_done=torch.cuda.Event())
            for _ in range(NB)]
    grads_h = torch.empty(flat.numel, dtype=torch.float32).pin_memory()
    s_act = torch.tensor([0.03], device=dev, requires_grad=True)
    b_act = torch.tensor([0.0], device=dev, requires_grad=True)
    w_leaf = [(st.x.clone().requires_grad_(True), st.scale.clone().requires_grad_(True), st.shift.clone().requires_grad_(True), st.grad)
              for st in wsites]
    sizes = [B * math.prod(s) for s in ACT_SHAPES]
    h2d = sum(2 * 2 * n for n in sizes)
    d2h = sum(2 * 2 * n for n in sizes) + 4 * flat.numel
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)

    def one_step():
        for i, n in enumerate(sizes):
            slot = ring[i % NB]
            with torch.cuda.stream(s_in):
                s_in.wait_event(slot["cmp_done"])            # the buffers' previous kernels are done
                slot["xd"][:n].copy_(xh[:n], non_blocking=True)
                slot["gd"][:n].copy_(gh[:n], non_blocking=True)
                slot["in_done"].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(slot["in_done"])
                s_cmp.wait_event(slot["out_done"])           # previous outputs of this slot have left the device
                xl = slot["xd"][:n].detach().requires_grad_(True)
                y = lsq(xl, s_act, b_act, 0, 127, 0, 255)
                y.backward(slot["gd"][:n])
                slot["yd"], slot["gxd"] = y.detach(), xl.grad
                slot["cmp_done"].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(slot["cmp_done"])
                yh[:n].copy_(slot["yd"], non_blocking=True)
                gxh[:n].copy_(slot["gxd"], non_blocking=True)
                slot["yd"].record_stream(s_out); slot["gxd"].record_stream(s_out)
                slot["out_done"].record(s_out)
        with torch.cuda.stream(s_cmp):
            for w, s, b, g in w_leaf:
                yw = lsq(w, s, b, -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True)
                yw.backward(g)
            packed = torch.cat([s_act.grad, b_act.grad] + [t for _, s, b, _ in w_leaf for t in (s.grad, b.grad)])
            grads_h[:packed.numel()].copy_(packed, non_blocking=True)
            for w, s, b, _ in w_leaf:
                w.grad = s.grad = b.grad = None
            s_act.grad = b_act.grad = None
        cur = torch.cuda.current_stream(dev)
        for st in (s_in, s_cmp, s_out):
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_step()
    barrier()
    # PCIe ceiling of this box for the same traffic shape: the largest site's x, g in and y, gx out as bare copies on two
    # streams, nothing else running (CUDA events; this is the roofline the end-to-end number is bound by)
    dummy = ring[0]
    pe = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(2):
        pe[0].record(s_in); pe[2].record(s_out)
        for _ in range(2):
            with torch.cuda.stream(s_in):
                dummy["xd"].copy_(xh, non_blocking=True); dummy["gd"].copy_(gh, non_blocking=True)
            with torch.cuda.stream(s_out):
                yh.copy_(dummy["xd"], non_blocking=True); gxh.copy_(dummy["gd"], non_blocking=True)
        pe[1].record(s_in); pe[3].record(s_out)
        torch.cuda.synchronize()
    pcie_in = 4 * 2 * nmax / (pe[0].elapsed_time(pe[1]) * 1e-3) / 1e9
    pcie_out = 4 * 2 * nmax / (pe[2].elapsed_time(pe[3]) * 1e-3) / 1e9
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = t.item() / steps
    n_w = sum(math.prod(s) for s in W_SHAPES)
    alg = (5 * 2 * sum(sizes) + 5 * 4 * n_w) * world
    return {"value": round(alg / dt / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": round(dt * 1e3, 2), "pcie_GBps_each_way": round(h2d / dt / 1e9, 1),
            "pcie_ceiling_GBps": {"h2d": round(pcie_in, 1), "d2h": round(pcie_out, 1),
                                  "how": "bare pinned<->device copies of the largest site, both directions at once, CUDA events"},
            "path": "pinned host x,g -> torchlsq.functional.lsq + autograd (copy-in / compute / copy-out streams) -> pinned host y,gx,grads; weights stay on device"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=256)
    ap.add_argument("--ref-batch", type=int, default=2, help="images per activation site in the CPU reference sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-plan-mode", action="store_true", help="skip the extra multi-tensor-plan measurement")
    ap.add_argument("--plan-activations", action="store_true", help="run all activation sites through one multi-tensor plan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-BASELINE-config measurements (configs key)")
    ap.add_argument("--no-strong", action="store_true", help="skip the batch-2048 single-GPU point of BASELINE configs[4]")
    ap.add_argument("--strong", action="store_true",
                    help="BASELINE configs[4] as written: global batch 2048 split over the ranks (strong scaling), site at a time with "
                         "buffer reuse; the JSON line then reports that job")
    ap.add_argument("--no-api-mode", action="store_true", help="skip the public-API (functional op / module) measurement")
    ap.add_argument("--inline-allreduce", action="store_true", help="N>1: all-reduce on the compute stream instead of the side stream")
    ap.add_argument("--no-fusion-mode", action="store_true", help="skip the workload-level prologue-fusion measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
