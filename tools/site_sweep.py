#!/usr/bin/env python
"""In-stream cost of ONE fake-quant launch as a function of site size and tiling knobs.

For every activation-site size of the ResNet-50 workload (bf16, batch 256) the forward / backward C-ABI call is issued
`reps` times back to back over rotating buffer sets (footprint > 2x L2), timed with CUDA events around the whole batch:
us per launch INCLUDING whatever ramp / tail the PDL chain does not hide.  Knobs go through lsqb200_set_tuning.

    python tools/site_sweep.py [--specs "a=1,b=2;c=3"] [--dtype bf16|f32]
"""
import argparse
import itertools
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "lsqfakequantize-pytorch_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench as B  # noqa: E402
from torchlsq import _cabi  # noqa: E402

DEV = "cuda:0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--specs", default="")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--rounds", type=int, default=5)
    args = ap.parse_args()
    lib = _cabi.load()
    peak = 6553.6
    dt, code, es = (torch.bfloat16, _cabi.BF16, 2) if args.dtype == "bf16" else (torch.float32, _cabi.F32, 4)
    sizes = sorted({256 * math.prod(s) for s in B.ACT_SHAPES})
    if args.sizes:
        sizes = [int(v) for v in args.sizes.split(",")]
    specs = [s for s in args.specs.split(";")] if args.specs else [""]
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    s, b = torch.tensor([0.03], device=DEV), torch.tensor([-1.7], device=DEV)
    gs, gb = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    out = {}
    for n in sizes:
        nset = max(2, min(12, int(math.ceil(600e6 / (4 * es * n)))))
        sets = [(torch.empty(n, dtype=dt, device=DEV).normal_(), torch.empty(n, dtype=dt, device=DEV),
                 torch.empty(n, dtype=dt, device=DEV).normal_(), torch.empty(n, dtype=dt, device=DEV)) for _ in range(nset)]
        import statistics
        acc = {(spec, kind): [] for spec in specs for kind in ("fwd", "bwd")}

        def run(kind, k):
            x, y, g, gx = sets[k % nset]
            if kind == "fwd":
                lib.lsqb200_fwd_tensor(x.data_ptr(), y.data_ptr(), s.data_ptr(), b.data_ptr(), n, code, 0, q, sp)
            else:
                lib.lsqb200_bwd_tensor(g.data_ptr(), x.data_ptr(), gx.data_ptr(), s.data_ptr(), b.data_ptr(), gs.data_ptr(),
                                       gb.data_ptr(), n, code, 0, q, ws.data_ptr(), ws.numel(), sp)
        # rounds interleave the specs (rotating start) so clock / power drift hits every spec alike; medians are reported
        for rnd in range(args.rounds + 1):
            order = specs[rnd % len(specs):] + specs[:rnd % len(specs)]
            for spec in order:
                assert lib.lsqb200_set_tuning(spec.encode()) == 0, spec
                for kind in ("fwd", "bwd"):
                    for k in range(nset):
                        run(kind, k)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for k in range(args.reps):
                        run(kind, k)
                    e1.record()
                    e1.synchronize()
                    if rnd:
                        acc[(spec, kind)].append(e0.elapsed_time(e1) * 1e3 / args.reps)
        for spec in specs:
            res = {}
            for kind in ("fwd", "bwd"):
                t = statistics.median(acc[(spec, kind)])
                by = (2 if kind == "fwd" else 3) * es * n
                res[kind] = dict(us=round(t, 2), GBps=round(by / t / 1e3, 1), over_us=round(t - by / peak / 1e3, 2))
            out.setdefault(str(n), {})[spec or "default"] = res
            print(f"{n:11d} {spec or 'default':60s} fwd {res['fwd']['us']:8.2f} us {res['fwd']['GBps']:7.1f} (+{res['fwd']['over_us']:5.2f})   "
                  f"bwd {res['bwd']['us']:8.2f} us {res['bwd']['GBps']:7.1f} (+{res['bwd']['over_us']:5.2f})", flush=True)
        del sets
    lib.lsqb200_set_tuning(b"")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
