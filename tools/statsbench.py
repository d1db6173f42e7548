"""A/B of the mu +- 3 sigma statistics kernel variants over the 54 ResNet-50 weights (4 rotating sets, bench.py::_timed_rotating):
GB/s isolated | back to back, and bit-for-bit agreement of the 27 560 scales with the first variant.
    python tools/statsbench.py [variants, e.g. 2,4,5,6]"""
import math, sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
import bench as B
from torchlsq import _cabi
from torchlsq.multi import LSQPlan, Site
lib = _cabi.load(); DEV = 'cuda:0'
stream = torch.cuda.current_stream()
variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2,4,5,6").split(",")]
gen = torch.Generator(device=DEV).manual_seed(0)
nw = sum(math.prod(s) for s in B.W_SHAPES)
ref = None
for dtype in (torch.float32, torch.bfloat16):
    sets = []
    for _ in range(4):
        sites = []
        for shp in B.W_SHAPES:
            w = torch.empty(shp, device=DEV).normal_(0, 0.05, generator=gen).to(dtype)
            sites.append(Site(x=w, scale=torch.ones(shp[0], device=DEV), shift=torch.zeros(shp[0], device=DEV), quant_min=-128, quant_max=127,
                              type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True))
        sets.append(sites)
    ref = None
    for v in variants:
        lib.lsqb200_set_tuning(f"rowstats={v}".encode())
        plans = [LSQPlan(s) for s in sets]          # the kernel choice is made at plan creation
        outs = [torch.empty(p.num_param_slots, device=DEV) for p in plans]
        plans[0].weight_init_stats(outs[0]); torch.cuda.synchronize()
        cur = outs[0].clone()
        if ref is None:
            ref = cur
        same = torch.equal(cur.view(torch.int32), ref.view(torch.int32))
        t = B._timed_rotating(torch, [(lambda p=p, o=o: p.weight_init_stats(o)) for p, o in zip(plans, outs)], 16, stream)
        es = 4 if dtype == torch.float32 else 2
        print(f"{dtype} rowstats={v}: same={same} launches={plans[0]._lib.lsqb200_plan_launches(plans[0]._h, 0)} "
              f"{es*nw/t[0]/1e6:6.0f} | {es*nw/t[2]/1e6:6.0f} GB/s  ({es*nw/t[0]/1e6/6556.8:.3f} | {es*nw/t[2]/1e6/6556.8:.3f})  {t[0]*1e3:.1f} us", flush=True)
        for p in plans:
            p.close()
