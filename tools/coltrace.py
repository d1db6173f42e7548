"""Per-CTA timeline of the column backward (library built with -DLSQ_COL_TRACE: TORCHLSQ_B200_LIB=ab/lib_trace.so): globaltimer stamps
0 entry, 1 after the dependency wait, 2 parameters formed, 3 row loop done, 4 atomics issued, 5 ticket taken, 6 finalised (last CTA), 7 = SM id.
    python tools/coltrace.py "col_dyn=0" "col_dyn=1" """
import sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
from torchlsq import _cabi
lib = _cabi.load(); DEV = 'cuda:0'
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
sp = torch.cuda.current_stream().cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
specs = sys.argv[1:] or ["col_dyn=0", "col_dyn=1"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
for outer, C, inner in ((256, 2048, 49), (256 * 196, 1024, 1), (256, 1024, 196)):
    n = outer * C * inner
    x = torch.empty(n, dtype=torch.float16, device=DEV).normal_(); g = torch.empty_like(x).normal_(); gx = torch.empty_like(x)
    s = 0.02 + 0.02 * torch.rand(C, device=DEV); b = -torch.rand(C, device=DEV)
    gs = torch.empty(C, device=DEV); gb = torch.empty(C, device=DEV)
    for spec in specs:
        assert lib.lsqb200_set_tuning(spec.encode()) == 0
        for it in range(3):
            flush.zero_(); torch.cuda.synchronize()
            ws[16384:16384 + 65536 * 8].zero_(); torch.cuda.synchronize()
            assert lib.lsqb200_bwd_channel(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(),
                                           outer, C, inner, 1, 0, q, ws.data_ptr(), ws.numel(), sp) == 0
            torch.cuda.synchronize()
        t = ws[16384:16384 + 4096 * 64].view(torch.int64).view(-1, 8).cpu()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min().item()
        rel = (t[:, :7] - t0).double() / 1e3       # us
        last = t[:, 6] > 0
        def st(col):
            v = rel[:, col]
            return f"{v.min():6.1f}/{v.mean():6.1f}/{v.max():6.1f}"
        loop = rel[:, 3] - rel[:, 2]
        print(f"({outer},{C},{inner}) [{spec}] ctas {len(t)} sms {len(set(t[:, 7].tolist()))}: entry {st(0)}  wait {st(1)}  params {st(2)}  loop-end {st(3)}  atomics {st(4)}  ticket {st(5)}"
              f"  final {rel[last, 6].max().item() if last.any() else -1:6.1f} us;  loop time min/mean/max {loop.min():.1f}/{loop.mean():.1f}/{loop.max():.1f}  algorithmic at mean loop-end {3*2*n/rel[:,3].mean().item()/1e3:.0f} GB/s, at final {3*2*n/rel[last,6].max().item()/1e3:.0f} GB/s", flush=True)
        # per-SM busy end spread
        by = {}
        for r in range(len(t)):
            by.setdefault(int(t[r, 7]), []).append(rel[r, 3].item())
        ends = sorted(max(v) for v in by.values())
        print("    per-SM loop-end (us) deciles:", " ".join(f"{ends[int(i * (len(ends) - 1) / 10)]:.1f}" for i in range(11)), flush=True)
