"""ncu target: the fused observer step on a 411 MB bf16 activation (per-tensor MovingAverageMinMax), 4 calls."""
import sys
import torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
from torchlsq.quantized.modules.observers import observer_step
DEV = 'cuda:0'
x = torch.empty(256 * 256 * 56 * 56, dtype=torch.bfloat16, device=DEV).normal_()
obs = torch.quantization.MovingAverageMinMaxObserver(dtype=torch.quint8, qscheme=torch.per_tensor_affine, quant_min=0, quant_max=127).to(DEV)
s = torch.ones(1, device=DEV); b = torch.zeros(1, device=DEV)
for _ in range(4):
    assert observer_step(obs, x, s, b)
torch.cuda.synchronize()
print(float(s), float(b))
