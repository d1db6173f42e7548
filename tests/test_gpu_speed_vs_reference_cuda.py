"""Same GPU, same tensors, same public API (`torchlsq.functional.lsq` + autograd): this repo's package against the reference
package with its OWN CUDA op (oracle/_ref, the unmodified reference sources built for sm_100a by oracle/build_ref.py), on the BASELINE
configs the reference's CUDA op can run at all (fp32 / fp16; it has no bf16).  Each package runs in its own subprocess
(both register `torchlsq::`).  The reference's path is 1 + 3 element-wise launches, two full-size temporaries, two `at::sum` and
four `.item()` host syncs per forward + backward (/root/reference/torchlsq/csrc/ops/cuda/lsq_cuda.cu:52-58,120-141); here it is two
launches and no sync.  torch's own learnable fake-quant op (`torch._fake_quantize_learnable_per_tensor_affine`, an LSQ
implementation inside PyTorch) is timed next to them as a second library baseline where its semantics apply (per tensor).

The test asserts the drop-in is not slower than what it replaces and writes the figures to gpurun_out/ref_cuda_speed.json
(CUDA events, median of 20 after warm-up, buffers rotated through 4 sets so nothing is L2-resident across iterations)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]

_SCRIPT = r'''
import sys, json, statistics
sys.path.insert(0, sys.argv[1])
import torch
import torchlsq
from torchlsq.functional import lsq
dev = "cuda:0"
R = 4
def timed(fns, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(iters + 2 * len(fns)):
        torch.cuda._sleep(120000)
        e0.record(); fns[i % len(fns)](); e1.record(); e1.synchronize()
        if i >= 2 * len(fns):
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
out = dict(which=torchlsq.__file__)
gen = torch.Generator().manual_seed(1)
# config 1: per-tensor quint8, fp32 32x64x56x56, grad scaling on
def cfg1(dtype):
    sets = []
    for _ in range(R):
        x = torch.randn(32, 64, 56, 56, generator=gen).to(dtype).to(dev).requires_grad_(True)
        g = torch.randn(32, 64, 56, 56, generator=gen).to(dtype).to(dev)
        s = torch.tensor([0.03], device=dev, dtype=dtype, requires_grad=True); b = torch.tensor([-1.7], device=dev, dtype=dtype, requires_grad=True)
        sets.append((x, g, s, b))
    def f(k):
        x, g, s, b = sets[k]
        def run():
            y = lsq(x, s, b, 0, 127, 0, 255)
            y.backward(g)
            x.grad = None; s.grad = None; b.grad = None
        return run
    return timed([f(k) for k in range(R)]), sets
out["config1_fp32_ms"], sets = cfg1(torch.float32)
if len(sys.argv) > 3 and sys.argv[3] == "native":
    # torch's own learnable fake-quant (per tensor): same tensors, zero_point as a learnable float tensor
    def nat(k):
        x, g, s, b = sets[k]
        zp = torch.tensor([57.0], device=dev, requires_grad=True)
        def run():
            y = torch._fake_quantize_learnable_per_tensor_affine(x, s, zp, 0, 127, 1.0)
            y.backward(g)
            x.grad = None; s.grad = None; zp.grad = None
        return run
    out["config1_fp32_torch_native_learnable_ms"] = timed([nat(k) for k in range(R)])
del sets
# config 4 shape at fp16 with fp16 parameters (what the reference accepts), per channel axis 1, 64 images per set; grad scaling OFF:
# the reference's fp16 gs overflows to 0 (SURVEY D7)
sets = []
for _ in range(R):
    x = torch.randn(64, 1024, 28, 28, generator=gen).half().to(dev).requires_grad_(True)
    g = torch.randn(64, 1024, 28, 28, generator=gen).half().to(dev)
    s = (0.02 + 0.02 * torch.rand(1024, generator=gen)).half().to(dev).requires_grad_(True); b = (-torch.rand(1024, generator=gen)).half().to(dev).requires_grad_(True)
    sets.append((x, g, s, b))
def f4(k):
    x, g, s, b = sets[k]
    def run():
        y = lsq(x, s, b, 0, 127, 0, 255, axis=1, use_grad_scaling=False, is_perchannel=True)
        y.backward(g)
        x.grad = None; s.grad = None; b.grad = None
    return run
out["config4_fp16_64x1024x28x28_ms"] = timed([f4(k) for k in range(R)])
del sets
# config 2: one ResNet-50 3x3 conv weight (512x512x3x3 fp32), per channel axis 0, symmetric
sets = []
for _ in range(R):
    w = (torch.randn(512, 512, 3, 3, generator=gen) * 0.05).to(dev).requires_grad_(True)
    g = torch.randn(512, 512, 3, 3, generator=gen).to(dev)
    s = torch.full((512,), 0.002, device=dev, requires_grad=True); b = torch.zeros(512, device=dev, requires_grad=True)
    sets.append((w, g, s, b))
def f2(k):
    w, g, s, b = sets[k]
    def run():
        y = lsq(w, s, b, -128, 127, -128, 127, axis=0, is_affine=False, is_perchannel=True)
        y.backward(g)
        w.grad = None; s.grad = None; b.grad = None
    return run
out["config2_weight_512x512x3x3_fp32_ms"] = timed([f2(k) for k in range(R)])
json.dump(out, open(sys.argv[2], "w"))
'''


def _run(pkg_dir, out, *extra):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "-c", _SCRIPT, str(pkg_dir), str(out), *extra], capture_output=True, text=True, env=env, cwd=str(out.parent))
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(out.read_text())


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "torchlsq" / "_C.so").exists(), reason="reference CUDA build (oracle/_ref) not present")
def test_public_op_is_faster_than_the_reference_cuda_op(tmp_path):
    import gpu_util as U
    mine = _run(ROOT / "lsqfakequantize-pytorch_b200", tmp_path / "mine.json", "native")
    ref = _run(ROOT / "oracle" / "_ref", tmp_path / "ref.json")
    assert "lsqfakequantize-pytorch_b200" in mine["which"] and "oracle/_ref" in ref["which"]
    report = {"protocol": "torchlsq.functional.lsq + autograd, forward + backward, CUDA events, median of 20, 4 rotating buffer sets, spin-kernel cover"}
    alg = {"config1_fp32_ms": 5 * 4 * 32 * 64 * 56 * 56, "config4_fp16_64x1024x28x28_ms": 5 * 2 * 64 * 1024 * 784,
           "config2_weight_512x512x3x3_fp32_ms": 5 * 4 * 512 * 512 * 9}
    for key, nbytes in alg.items():
        report[key[:-3]] = dict(b200_ms=round(mine[key], 4), reference_cuda_ms=round(ref[key], 4), speedup=round(ref[key] / mine[key], 2),
                                b200_GBps=round(nbytes / mine[key] / 1e6, 1), reference_cuda_GBps=round(nbytes / ref[key] / 1e6, 1))
        assert mine[key] < ref[key], (key, mine[key], ref[key])
    nat = mine["config1_fp32_torch_native_learnable_ms"]
    report["config1_fp32"]["torch_native_learnable_fake_quant_ms"] = round(nat, 4)
    report["config1_fp32"]["speedup_vs_torch_native"] = round(nat / mine["config1_fp32_ms"], 2)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "ref_cuda_speed.json").write_text(json.dumps(U.stamped(report), indent=1))
