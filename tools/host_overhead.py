"""Host-side cost of one public-op call (tiny tensor, so the kernels are negligible): ours vs the reference CUDA build.
    python tools/host_overhead.py            # this repo's torchlsq
    python tools/host_overhead.py ref        # oracle/_ref (the reference's C++ front end + ATen kernels)"""
import sys, time
ref = len(sys.argv) > 1 and sys.argv[1] == "ref"
sys.path.insert(0, "oracle/_ref" if ref else "lsqfakequantize-pytorch_b200")
import torch
import torchlsq
from torchlsq.functional import lsq
dev = "cuda:0"
x = torch.randn(4, 64, 8, 8, device=dev, requires_grad=True)
g = torch.randn(4, 64, 8, 8, device=dev)
s = torch.tensor([0.03], device=dev, requires_grad=True); b = torch.tensor([-1.7], device=dev, requires_grad=True)
sc = torch.full((64,), 0.03, device=dev, requires_grad=True); bc = torch.zeros(64, device=dev, requires_grad=True)
def t(fn, n=2000):
    for _ in range(200): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def fwd_nograd():
    with torch.no_grad(): lsq(x, s, b, 0, 127, 0, 255)
def fwd_bwd():
    y = lsq(x, s, b, 0, 127, 0, 255); y.backward(g); x.grad = None; s.grad = None; b.grad = None
def fwd_bwd_ch():
    y = lsq(x, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True); y.backward(g); x.grad = None; sc.grad = None; bc.grad = None
print("ref" if ref else "b200", "us/call: fwd(no_grad) %.1f  fwd+bwd per-tensor %.1f  fwd+bwd per-channel %.1f" % (t(fwd_nograd), t(fwd_bwd), t(fwd_bwd_ch)))
