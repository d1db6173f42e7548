"""A/B of column-kernel TUNING SPECS under the rotating-set protocol (bench.py::_timed_rotating): per layout and spec, forward and
backward GB/s (isolated median | back-to-back), after a bit-for-bit check of y / grad_x and a 1e-6 check of the sums against the first spec.
    python tools/colbench3.py "col_dyn=0" "col_dyn=1,col_chunk=8" ...        (TORCHLSQ_B200_LIB selects the library build)"""
import sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
import bench as B
from torchlsq import _cabi
lib = _cabi.load(); DEV = 'cuda:0'
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
specs = sys.argv[1:] or ["col_dyn=0", "col_dyn=1"]
NT = 256 * 1024 * 784
x = torch.empty(NT, dtype=torch.float16, device=DEV).normal_(); g = torch.empty_like(x).normal_()
y = torch.empty_like(x); gx = torch.empty_like(x)
SHAPES = ((256, 2048, 49), (256 * 196, 1024, 1), (256, 1024, 196), (64, 512, 49), (4096, 96, 1), (1000, 20, 3))
for dt, name in ((1, "fp16"), (0, "fp32")):
    es = 2 if dt else 4
    tdt = torch.float16 if dt == 1 else torch.float32
    xv, gv, yv, gxv = (t.view(tdt) for t in (x, g, y, gx))
    xv.normal_(); gv.normal_()
    for outer, C, inner in SHAPES:
        n = outer * C * inner
        nset = max(1, min(8, NT * 2 // (n * es)))
        s = 0.02 + 0.02 * torch.rand(C, device=DEV); b = -torch.rand(C, device=DEV)
        gs = torch.empty(C, device=DEV); gb = torch.empty(C, device=DEV)
        ref = None
        for spec in specs:
            assert lib.lsqb200_set_tuning(spec.encode()) == 0, spec
            def fw(k):
                o = k * n * es
                return lambda: lib.lsqb200_fwd_channel(x.data_ptr() + o, y.data_ptr() + o, s.data_ptr(), b.data_ptr(), outer, C, inner, dt, 0, q, sp)
            def bw(k):
                o = k * n * es
                return lambda: lib.lsqb200_bwd_channel(g.data_ptr() + o, x.data_ptr() + o, gx.data_ptr() + o, s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                                       gb.data_ptr(), outer, C, inner, dt, 0, q, ws.data_ptr(), ws.numel(), sp)
            yv[:n].zero_(); gxv[:n].zero_(); gs.zero_(); gb.zero_()
            assert fw(0)() == 0 and bw(0)() == 0
            torch.cuda.synchronize()
            cur = (yv[:n].clone(), gxv[:n].clone(), gs.clone(), gb.clone())
            assert bw(0)() == 0            # twice: the workspace (accumulators, tickets, chunk counters) must come back clean
            torch.cuda.synchronize()
            again = torch.allclose(gs, cur[2], rtol=1e-9, atol=0) and torch.allclose(gb, cur[3], rtol=1e-9, atol=0)
            wsclean = int(ws.view(torch.int32)[:4096].abs().sum().item()) == 0
            if ref is None:
                ref = cur
            ok = torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and \
                torch.allclose(cur[2], ref[2], rtol=1e-6, atol=0) and torch.allclose(cur[3], ref[3], rtol=1e-6, atol=0)
            f = B._timed_rotating(torch, [fw(k) for k in range(nset)], 12, stream)
            bk = B._timed_rotating(torch, [bw(k) for k in range(nset)], 12, stream)
            print(f"{name} ({outer},{C},{inner}) [{spec}]: same={ok} rerun={again} ws0={wsclean}  fwd {2*es*n/f[0]/1e6:6.0f} | {2*es*n/f[2]/1e6:6.0f}   bwd {3*es*n/bk[0]/1e6:6.0f} | {3*es*n/bk[2]/1e6:6.0f}"
                  f"   fwd+bwd {5*es*n/(f[0]+bk[0])/1e6:6.0f} GB/s ({5*es*n/(f[0]+bk[0])/1e6/6556.8:.3f})", flush=True)
