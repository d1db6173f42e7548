"""Correctness + speed of the TMA-staged column backward vs the register-staged one (same inputs, same oracle)."""
import sys, statistics
import numpy as np
import torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gpu_util as U
from torchlsq import _cabi
lib = _cabi.load()
gen = torch.Generator().manual_seed(0)
ok = True
for tv in (1, 2, 3):
    for dt in (torch.float16, torch.bfloat16, torch.float32):
        for shape, axis in (((5, 1024, 14, 14), 1), ((3, 2048, 7, 7), 1), ((37, 4096, 1), 1), ((9, 520, 3, 3), 1), ((70, 2056), 1)):
            n = int(np.prod(shape)); outer = int(np.prod(shape[:axis])); C = shape[axis]; inner = n // (outer * C)
            x = torch.randn(n, generator=gen).to(dt).to(U.DEV); g = torch.randn(n, generator=gen).to(dt).to(U.DEV)
            s = (0.02 + 0.02 * torch.rand(C, generator=gen)).to(U.DEV); b = (-torch.rand(C, generator=gen)).to(U.DEV)
            for mode in (dict(), dict(init_mode=True), dict(eval_mode=True)):
                q = U.qa(**mode)
                lib.lsqb200_set_tuning(b"col_tma=0")
                gx0, gs0, gb0 = U.bwd(g, x, s, b, q, outer, C, inner, True)
                lib.lsqb200_set_tuning(("col_tma=%d" % tv).encode())
                gx1, gs1, gb1 = U.bwd(g, x, s, b, q, outer, C, inner, True)
                torch.cuda.synchronize()
                ogx, ogs, ogb, ms, mb = U.oracle_bwd(g, x, s, b, q, outer, C, inner, True)
                good = U.same_bits(gx1, ogx) and torch.equal(gx0, gx1)
                try:
                    U.assert_grads_close(gs1, ogs, ms, 1e-5); U.assert_grads_close(gb1, ogb, mb, 1e-5)
                except AssertionError as e:
                    good = False
                if not good:
                    ok = False
                    print("MISMATCH", tv, dt, shape, mode, flush=True)
print("coltma parity", "ok" if ok else "FAILED", flush=True)
ws = U.workspace(); sp = U.stream()
q = U.qa()
for shape in ((256, 1024, 196), (256, 2048, 49), (50176, 1024, 1), (1024, 1024, 196)):
    outer, C, inner = shape
    N = outer * C * inner
    x = torch.empty(N, dtype=torch.float16, device=U.DEV).normal_(); g = torch.empty_like(x).normal_(); gx = torch.empty_like(x)
    s = 0.02 + 0.02 * torch.rand(C, device=U.DEV); b = -torch.rand(C, device=U.DEV); gs = torch.empty(C, device=U.DEV); gb = torch.empty(C, device=U.DEV)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=U.DEV)
    for tune in ("col_tma=0", "col_tma=1", "col_tma=2", "col_tma=3", "col_tma=1,col_waves_bwd=2", "col_tma=2,col_waves_bwd=2", "col_tma=3,col_waves_bwd=2", "col_tma=2,col_waves_bwd=3"):
        lib.lsqb200_set_tuning(tune.encode())
        ts = []
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for i in range(9):
            flush.fill_(i)
            e0.record()
            lib.lsqb200_bwd_channel(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(), outer, C, inner, 1, 0, q,
                                    ws.data_ptr(), ws.numel(), sp)
            e1.record(); e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        print(shape, tune, "bwd GB/s (L2 flushed)", round(3 * 2 * N / statistics.median(ts) / 1e6), flush=True)
    del x, g, gx, flush
