#!/usr/bin/env python
"""Per-tensor bf16 forward + backward on three ResNet-50 site sizes, three rounds, for an ncu launch list
(LSQB200_TUNE=flatkernels=0|1 selects the general or the lean kernels)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "lsqfakequantize-pytorch_b200"))
import torch  # noqa: E402
from torchlsq import _cabi  # noqa: E402

DEV = "cuda:0"
lib = _cabi.load()
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
sp = torch.cuda.current_stream().cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
s, b = torch.tensor([0.03], device=DEV), torch.tensor([-0.9], device=DEV)
gs, gb = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
for n in (6_422_528, 51_380_224, 205_520_896):
    x = torch.randn(n, device=DEV).to(torch.bfloat16)
    g = torch.randn(n, device=DEV).to(torch.bfloat16)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
        lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, sp)
        lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                               n, 2, 0, q, ws.data_ptr(), ws.numel(), sp)
    torch.cuda.synchronize()
print("ok")
