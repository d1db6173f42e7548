#!/usr/bin/env python
"""One fused-prologue forward + backward of each kind on the largest ResNet-50 site (bf16 256x64x112x112), for Nsight
Compute: 2 warm-up rounds, then the captured round.  Launch order per round: relu fwd, relu bwd, add_relu fwd, add_relu bwd.

    ncu --set full --clock-control none --import-source on --kernel-name regex:lsq_ --launch-skip 8 --launch-count 4 \
        -f -o gpurun_out/prof_pre python tools/preprof.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "lsqfakequantize-pytorch_b200"))

import torch  # noqa: E402

from torchlsq import _cabi  # noqa: E402

DEV = "cuda:0"
lib = _cabi.load()
shape = (256, 64, 112, 112)
gen = torch.Generator(device=DEV).manual_seed(1)
x = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
x2 = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
g = torch.randn(shape, device=DEV, generator=gen).to(torch.bfloat16)
y, gx = torch.empty_like(x), torch.empty_like(x)
s, b = torch.tensor([0.03], device=DEV), torch.tensor([-0.9], device=DEV)
gs, gb = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
sp = torch.cuda.current_stream().cuda_stream
n = x.numel()
for _ in range(3):
    for code, second in ((_cabi.PRE_RELU, None), (_cabi.PRE_ADD_RELU, x2.data_ptr())):
        assert lib.lsqb200_fwd_tensor_pre(x.data_ptr(), second, y.data_ptr(), s.data_ptr(), b.data_ptr(), n, 2, 0, q, code, sp) == 0
        assert lib.lsqb200_bwd_tensor_pre(g.data_ptr(), x.data_ptr(), second, gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                          gb.data_ptr(), n, 2, 0, q, code, ws.data_ptr(), ws.numel(), sp) == 0
    torch.cuda.synchronize()
print("ok", float(gs), float(gb))
