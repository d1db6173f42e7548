"""ncu target for the column-layout kernels: python tools/colprof.py <outer> <C> <inner> [dtype code] [tuning spec]
(3 warm-up rounds of forward + backward on one per-channel tensor; profile with --launch-skip 4 --launch-count 2)."""
import sys
import torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
from torchlsq import _cabi
lib = _cabi.load(); DEV = 'cuda:0'
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
sp = torch.cuda.current_stream().cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
outer, C, inner = (int(v) for v in sys.argv[1:4])
dt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if len(sys.argv) > 5:
    lib.lsqb200_set_tuning(sys.argv[5].encode())
N = outer * C * inner
tdt = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[dt]
x = torch.empty(N, dtype=tdt, device=DEV).normal_(); g = torch.empty_like(x).normal_(); y = torch.empty_like(x); gx = torch.empty_like(x)
s = 0.02 + 0.02 * torch.rand(C, device=DEV); b = -torch.rand(C, device=DEV); gs = torch.empty(C, device=DEV); gb = torch.empty(C, device=DEV)
for _ in range(3):
    lib.lsqb200_fwd_channel(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), outer, C, inner, dt, 0, q, sp)
    lib.lsqb200_bwd_channel(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(), outer, C, inner, dt, 0, q,
                            ws.data_ptr(), ws.numel(), sp)
torch.cuda.synchronize()
