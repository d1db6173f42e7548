"""Drop-in check at the level a user works at: the same small QAT training script (conv net with `LSQFakeQuantizer`
on the input, two activations and three weights; learned init, observer init, mu +- 3 sigma weight init, then LSQ
training with SGD) is run twice on the GPU - once importing this repo's `torchlsq`, once importing the reference
package with its own CUDA op (oracle/_ref, built for sm_100a) - and the loss curve, the learned scale / shift
parameters and the module state must agree.  Each run is a subprocess because both packages register `torchlsq::`."""
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
import sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.nn as nn, torch.nn.functional as F
import torchlsq
from torchlsq import LSQFakeQuantizer
from torchlsq.quantized.modules import observers as _obs
import functools
if not hasattr(_obs, "partial"):
    _obs.partial = functools.partial          # the reference forgot this import (SURVEY D10)
torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
MA = torch.quantization.MovingAverageMinMaxObserver
dev = "cuda:0"

class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, padding=1)
        self.conv2 = nn.Conv2d(16, 32, 3, padding=1, stride=2)
        self.fc = nn.Linear(32 * 8 * 8, 10)
        self.fq_in = LSQFakeQuantizer(MA, 'activation', init_mode='observer', init_batches=2)
        self.fq_a1 = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=2, init_scale=0.05)
        self.fq_a2 = LSQFakeQuantizer(None, 'activation', init_mode='learnable', init_batches=3, init_scale=0.05,
                                      qscheme=torch.per_channel_affine)
        wkw = dict(dtype=torch.qint8, qscheme=torch.per_channel_symmetric, init_mode='learnable')
        self.fq_w1 = LSQFakeQuantizer(None, 'weight', **wkw)
        self.fq_w2 = LSQFakeQuantizer(None, 'weight', avoid_torch_overflow=False, **wkw)
        self.fq_w3 = LSQFakeQuantizer(None, 'weight', dtype=torch.qint8, qscheme=torch.per_tensor_symmetric, init_mode='learnable')
    def forward(self, x):
        x = self.fq_in(x)
        x = F.relu(F.conv2d(x, self.fq_w1(self.conv1.weight), self.conv1.bias, padding=1))
        x = self.fq_a1(x)
        x = F.relu(F.conv2d(x, self.fq_w2(self.conv2.weight), self.conv2.bias, padding=1, stride=2))
        x = self.fq_a2(x)
        return F.linear(x.flatten(1), self.fq_w3(self.fc.weight), self.fc.bias)

torch.manual_seed(0)
net = Net()
gen = torch.Generator().manual_seed(1)
data = [(torch.randn(32, 3, 16, 16, generator=gen), torch.randint(0, 10, (32,), generator=gen)) for _ in range(10)]
net = net.to(dev).train()
net(data[0][0].to(dev))                      # first call only creates the LSQ parameters (module contract)
opt = torch.optim.SGD(net.parameters(), lr=0.02, momentum=0.9)
losses = []
for step in range(10):
    x, t = data[step % len(data)]
    loss = F.cross_entropy(net(x.to(dev)), t.to(dev))
    opt.zero_grad()
    loss.backward()
    opt.step()
    losses.append(float(loss))
net.eval()
with torch.no_grad():
    ev = float(F.cross_entropy(net(data[0][0].to(dev)), data[0][1].to(dev)))
out = dict(losses=losses, eval_loss=ev, which=torchlsq.__file__)
for name in ("fq_in", "fq_a1", "fq_a2", "fq_w1", "fq_w2", "fq_w3"):
    m = getattr(net, name)
    out[name] = dict(scale=m.scale.detach().float().cpu().flatten().tolist(), shift=m.shift.detach().float().cpu().flatten().tolist(),
                     current_batch=int(m.current_batch[0]), observer_enabled=int(m.observer_enabled[0]),
                     learning_enabled=int(m.learning_enabled[0]), quant_min=m.quant_min, quant_max=m.quant_max)
    sc, zp = m.calculate_qparams()
    out[name]["qparams_scale"] = sc.float().cpu().flatten().tolist()
    out[name]["qparams_zp"] = zp.cpu().flatten().tolist()
json.dump(out, open(sys.argv[2], "w"))
'''


def _run(pkg_dir, out):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "-c", _SCRIPT, str(pkg_dir), str(out)], capture_output=True, text=True, env=env, cwd=str(out.parent))
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(out.read_text())


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "torchlsq" / "_C.so").exists(), reason="reference CUDA build (oracle/_ref) not present")
def test_qat_training_matches_reference_package(tmp_path):
    mine = _run(ROOT / "lsqfakequantize-pytorch_b200", tmp_path / "mine.json")
    ref = _run(ROOT / "oracle" / "_ref", tmp_path / "ref.json")
    assert "lsqfakequantize-pytorch_b200" in mine["which"] and "oracle/_ref" in ref["which"]
    lm, lr = torch.tensor(mine["losses"]), torch.tensor(ref["losses"])
    assert torch.allclose(lm, lr, rtol=2e-3, atol=1e-4), (mine["losses"], ref["losses"])
    assert abs(mine["eval_loss"] - ref["eval_loss"]) <= 2e-3 * abs(ref["eval_loss"]) + 1e-4
    report = dict(losses_mine=mine["losses"], losses_ref=ref["losses"], eval=(mine["eval_loss"], ref["eval_loss"]))
    for name in ("fq_in", "fq_a1", "fq_a2", "fq_w1", "fq_w2", "fq_w3"):
        a, b = mine[name], ref[name]
        for k in ("current_batch", "observer_enabled", "learning_enabled", "quant_min", "quant_max"):
            assert a[k] == b[k], (name, k, a[k], b[k])
        sa, sb = torch.tensor(a["scale"]), torch.tensor(b["scale"])
        assert torch.allclose(sa, sb, rtol=5e-3, atol=1e-6), (name, "scale", float((sa - sb).abs().max()))
        ha, hb = torch.tensor(a["shift"]), torch.tensor(b["shift"])
        assert torch.allclose(ha, hb, rtol=5e-3, atol=2e-4), (name, "shift", float((ha - hb).abs().max()))
        za, zb = torch.tensor(a["qparams_zp"]), torch.tensor(b["qparams_zp"])
        assert (za - zb).abs().max() <= 1, (name, "zero_point")
        report[name] = dict(max_rel_scale_diff=float(((sa - sb).abs() / sb.abs().clamp_min(1e-12)).max()))
    import gpu_util as U
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "qat_e2e_vs_reference.json").write_text(json.dumps(U.stamped(report), indent=1))
