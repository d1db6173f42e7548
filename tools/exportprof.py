"""ncu target: integer export (quantize / dequantize, LSQ semantics) of a 411 MB bf16 activation, 3 rounds."""
import sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
from torchlsq import export as EX
DEV = 'cuda:0'
x = torch.empty(256, 256, 56, 56, dtype=torch.bfloat16, device=DEV).normal_()
s, b = torch.tensor([0.03], device=DEV), torch.tensor([-1.7], device=DEV)
for _ in range(3):
    c = EX.quantize(x, s, b, 0, 127, 0, 255)
    y = EX.dequantize(c, s, b, 0, 127, 0, 255, dtype=torch.bfloat16)
torch.cuda.synchronize()
