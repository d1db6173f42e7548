#!/usr/bin/env python
"""Quantization-aware training with LSQ+ fake quantizers, the way the reference documents it
(/root/reference/README.md "Using": a QConfig of `LSQFakeQuantizer.with_args`, `prepare_qat`, one warm-up forward BEFORE the
optimizer is built) - unchanged user code, the B200-native kernels underneath.  Needs a CUDA device (there is no CPU path).

    python examples/qat_quickstart.py [--steps 30] [--group-weights] [--fuse-prologues]

--group-weights   all weight quantizers of the model in one launch per direction (torchlsq.multi.group_weight_quantizers)
--fuse-prologues  ReLU / residual add in front of an activation quantizer folded into its kernels (torchlsq.fusion.fuse_prologues)
"""
import argparse
import sys
import warnings
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "lsqfakequantize-pytorch_b200"))
import torchlsq  # noqa: E402
from torchlsq import LSQFakeQuantizer  # noqa: E402


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1, self.relu1 = nn.Conv2d(c, c, 3, padding=1), nn.ReLU()
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)
        self.skip = torch.ao.nn.quantized.FloatFunctional()

    def forward(self, x):
        return self.skip.add_relu(self.conv2(self.relu1(self.conv1(x))), x)


class Net(nn.Module):
    def __init__(self, c=32, classes=10):
        super().__init__()
        self.quant, self.dequant = torch.ao.quantization.QuantStub(), torch.ao.quantization.DeQuantStub()
        self.stem, self.relu = nn.Conv2d(3, c, 3, padding=1), nn.ReLU()
        self.b1, self.b2 = Block(c), Block(c)
        self.fc = nn.Linear(c, classes)

    def forward(self, x):
        x = self.relu(self.stem(self.quant(x)))
        x = self.b2(self.b1(x))
        return self.dequant(self.fc(x.mean((2, 3))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--group-weights", action="store_true")
    ap.add_argument("--fuse-prologues", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("torchlsq-b200 needs a CUDA device: there is no CPU implementation")
    dev = "cuda:0"
    torch.manual_seed(0)
    tq = torch.ao.quantization
    # activations: quint8, per tensor, affine, initialised by an observer over the first 5 batches, then learned (LSQ+)
    act = LSQFakeQuantizer.with_args(observer=tq.MovingAverageMinMaxObserver, otype="activation", init_mode="observer", init_batches=5)
    # weights: qint8, per output channel, symmetric, mu +- 3 sigma initialisation at the first call, learned step size (LSQ)
    wei = LSQFakeQuantizer.with_args(observer=None, otype="weight", dtype=torch.qint8, qscheme=torch.per_channel_symmetric,
                                     init_mode="learnable")
    net = Net().train()
    net.qconfig = tq.QConfig(activation=act, weight=wei)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tq.prepare_qat(net, inplace=True)
    net.to(dev)
    x = torch.randn(64, 3, 32, 32, device=dev)
    t = torch.randint(0, 10, (64,), device=dev)
    with torch.no_grad():
        net(x)                                   # creates scale / shift of every quantizer: build the optimizer AFTER this
    if args.fuse_prologues:
        from torchlsq.fusion import fuse_prologues
        print("fused prologues:", fuse_prologues(net))
    if args.group_weights:
        from torchlsq.multi import group_weight_quantizers
        group_weight_quantizers(net)
    opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    for step in range(args.steps):
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(net(x), t)
        loss.backward()
        opt.step()
        if step % 5 == 0 or step == args.steps - 1:
            print(f"step {step:3d}  loss {loss.item():.4f}")
    net.eval()
    fq = net.stem.weight_fake_quant
    scale, zero_point = fq.calculate_qparams()
    print("stem weight quantizer:", fq.extra_repr()[:120], "...")
    # integer export of the learned quantizers (what torch.quantization.convert consumes): real int8 codes, made on the device
    from torchlsq import export
    codes = export.quantize(net.stem.weight.detach(), fq.scale.detach(), fq.shift.detach(), fq.quant_min, fq.quant_max, -128, 127, axis=0,
                            is_perchannel=True)
    print("int8 codes:", codes.dtype, tuple(codes.shape), "range", int(codes.min()), int(codes.max()), "| scale[:3]", scale[:3].tolist())


if __name__ == "__main__":
    main()
