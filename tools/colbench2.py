"""A/B of column-kernel variants under the rotating-set protocol (bench.py::_timed_rotating): per layout and variant, forward and
backward GB/s (isolated median | back-to-back), after a bit-for-bit check of y / grad_x and of the sums against the first variant.
    python tools/colbench2.py [variants, e.g. 1,6,7,8] [extra tuning, e.g. ,col_waves_bwd=2]"""
import sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
import bench as B
from torchlsq import _cabi
lib = _cabi.load(); DEV = 'cuda:0'
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,6,7,8").split(",")]
extra = sys.argv[2] if len(sys.argv) > 2 else ""
NT = 256 * 1024 * 784
x = torch.empty(NT, dtype=torch.float16, device=DEV).normal_(); g = torch.empty_like(x).normal_()
y = torch.empty_like(x); gx = torch.empty_like(x)
for dt, name in ((1, "fp16"), (2, "bf16")):
    xv, gv, yv, gxv = (t.view(torch.float16 if dt == 1 else torch.bfloat16) for t in (x, g, y, gx))
    for outer, C, inner in ((256, 2048, 49), (256 * 196, 1024, 1), (256, 1024, 196)):
        n = outer * C * inner
        nset = NT // n
        s = 0.02 + 0.02 * torch.rand(C, device=DEV); b = -torch.rand(C, device=DEV)
        gs = torch.empty(C, device=DEV); gb = torch.empty(C, device=DEV)
        ref = None
        for v in variants:
            lib.lsqb200_set_tuning(f"col_variant={v}{extra}".encode())
            def fw(k):
                o = k * n * 2
                return lambda: lib.lsqb200_fwd_channel(x.data_ptr() + o, y.data_ptr() + o, s.data_ptr(), b.data_ptr(), outer, C, inner, dt, 0, q, sp)
            def bw(k):
                o = k * n * 2
                return lambda: lib.lsqb200_bwd_channel(g.data_ptr() + o, x.data_ptr() + o, gx.data_ptr() + o, s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                                       gb.data_ptr(), outer, C, inner, dt, 0, q, ws.data_ptr(), ws.numel(), sp)
            fw(0)(); bw(0)(); torch.cuda.synchronize()
            cur = (yv[:n].clone(), gxv[:n].clone(), gs.clone(), gb.clone())
            if ref is None:
                ref = cur
            ok = torch.equal(cur[0].view(torch.int16), ref[0].view(torch.int16)) and torch.equal(cur[1].view(torch.int16), ref[1].view(torch.int16)) and \
                torch.allclose(cur[2], ref[2], rtol=1e-6, atol=0) and torch.allclose(cur[3], ref[3], rtol=1e-6, atol=0)
            f = B._timed_rotating(torch, [fw(k) for k in range(nset)], 12, stream)
            bk = B._timed_rotating(torch, [bw(k) for k in range(nset)], 12, stream)
            print(f"{name} ({outer},{C},{inner}) v{v}: same={ok}  fwd {2*2*n/f[0]/1e6:6.0f} | {2*2*n/f[2]/1e6:6.0f}   bwd {3*2*n/bk[0]/1e6:6.0f} | {3*2*n/bk[2]/1e6:6.0f}"
                  f"   fwd+bwd {5*2*n/(f[0]+bk[0])/1e6:6.0f} GB/s ({5*2*n/(f[0]+bk[0])/1e6/6556.8:.3f})", flush=True)
