"""Per-CTA timeline of the lean per-tensor kernels inside the ResNet-50 step (library built with -DLSQ_FLAT_TRACE:
TORCHLSQ_B200_LIB=ab/lib_ftrace.so).  Runs the bench's 71 bf16 activation sites (forward of every site, then backward of every site,
per-site launches under PDL) and prints, per launch and in total, where the time between the kernels' own streaming goes.
Stamps per CTA: 0 entry, 1 dependency wait done, 2 unit loop done, 3 reduction chain done (backward)."""
import ctypes, math, sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
import bench as B
from torchlsq import _cabi
lib = _cabi.load(); DEV = 'cuda:0'
lib.lsqb200_trace_arm.restype = ctypes.c_longlong
lib.lsqb200_trace_arm.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
spec = sys.argv[1] if len(sys.argv) > 1 else ""
lib.lsqb200_set_tuning(spec.encode())
BATCH = 256
ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
sp = torch.cuda.current_stream().cuda_stream
q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
gen = torch.Generator(device=DEV).manual_seed(1)
sites = []
for i, shp in enumerate(B.ACT_SHAPES):
    n = BATCH * math.prod(shp)
    x = torch.empty(n, dtype=torch.bfloat16, device=DEV).normal_(0, 1, generator=gen)
    if i: x.relu_()
    g = torch.empty(n, dtype=torch.bfloat16, device=DEV).normal_(0, 1, generator=gen)
    sites.append(dict(n=n, x=x, g=g, y=torch.empty_like(x), gx=torch.empty_like(x), s=torch.tensor([0.03], device=DEV),
                      b=torch.tensor([0.0 if i else -1.9], device=DEV), gs=torch.zeros(1, device=DEV), gb=torch.zeros(1, device=DEV)))
def step():
    for a in sites:
        assert lib.lsqb200_fwd_tensor(a["x"].data_ptr(), a["y"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(), a["n"], 2, 0, q, sp) == 0
    for a in reversed(sites):
        assert lib.lsqb200_bwd_tensor(a["g"].data_ptr(), a["x"].data_ptr(), a["gx"].data_ptr(), a["s"].data_ptr(), a["b"].data_ptr(), a["gs"].data_ptr(),
                                      a["gb"].data_ptr(), a["n"], 2, 0, q, ws.data_ptr(), ws.numel(), sp) == 0
for _ in range(3): step()
torch.cuda.synchronize()
NL = 2 * len(sites)
buf = torch.zeros(NL * 8192 * 4, dtype=torch.int64, device=DEV)
lib.lsqb200_trace_arm(buf.data_ptr(), NL)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(400000)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
used = lib.lsqb200_trace_arm(None, 0)
t = buf.view(NL, 8192, 4).cpu()
print(f"step {e0.elapsed_time(e1):.3f} ms, traced launches {used} of {NL}  [{spec}]")
PEAK = 6556.8e9
prev_end = None
rows = []
for k in range(used):
    n, ctas, kind = (int(v) for v in t[k, 0, :3])
    r = t[k, 1:ctas + 1].double() / 1e3              # us
    last = 3 if kind == 2 else 2
    begin = r[:, 1].min().item()
    ends = r[:, last]
    end = ends.max().item()
    ideal = (3 if kind == 2 else 2) * 2 * n / PEAK * 1e6
    gap = begin - prev_end if prev_end is not None else 0.0
    srt = torch.sort(ends).values
    p50, p90, p99 = (srt[int(f * (ctas - 1))].item() for f in (0.5, 0.9, 0.99))
    chain = end - r[:, 2].max().item()
    first_entry = r[:, 0].min().item() - (prev_end if prev_end is not None else r[:, 0].min().item())
    resident_at_release = int((r[:, 0] <= begin).sum().item())         # CTAs already resident (sitting in the wait) when the predecessor finished
    rows.append(dict(k=k, kind="bwd" if kind == 2 else "fwd", n=n, ctas=ctas, ideal=ideal, span=end - (prev_end if prev_end is not None else begin), gap=gap,
                     busy=end - begin, tail90=end - p90, tail99=end - p99, chain=chain, first_entry=first_entry, resident=resident_at_release))
    prev_end = end
print("  k kind     elements ctas resident  ideal_us  span_us   gap  busy  busy/ideal  end-p90  end-p99  chain  first_entry-prev_end")
for r in rows:
    if r["k"] % 5 == 0 or r["n"] > 1.5e8:
        print("%(k)3d %(kind)s %(n)12d %(ctas)5d %(resident)5d %(ideal)9.1f %(span)8.1f %(gap)5.1f %(busy)6.1f" % r, "%8.3f %8.1f %8.1f %6.1f %8.1f" % (r["busy"] / r["ideal"], r["tail90"], r["tail99"], r["chain"], r["first_entry"]))
for kind in ("fwd", "bwd"):
    sel = [r for r in rows if r["kind"] == kind]
    S = lambda key: sum(r[key] for r in sel)
    print(f"{kind}: launches {len(sel)}  ideal {S('ideal'):.0f} us  span {S('span'):.0f} us = gap {S('gap'):.0f} + busy {S('busy'):.0f};  of busy: last 10 % of CTAs' spread {S('tail90'):.0f}, last 1 % {S('tail99'):.0f}, reduction chain {S('chain'):.0f};"
          f"  span / ideal = {S('span') / S('ideal'):.3f}")
