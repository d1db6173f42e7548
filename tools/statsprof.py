"""ncu target: mu +- 3 sigma statistics of the 54 ResNet-50 weights through a plan, 4 calls.  python tools/statsprof.py <rowstats variant> [dtype: f32|bf16]"""
import sys, torch
sys.path.insert(0, 'lsqfakequantize-pytorch_b200'); sys.path.insert(0, '.')
import bench as B
from torchlsq import _cabi
from torchlsq.multi import LSQPlan, Site
lib = _cabi.load(); DEV = 'cuda:0'
lib.lsqb200_set_tuning(f"rowstats={sys.argv[1]}".encode())
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
gen = torch.Generator(device=DEV).manual_seed(0)
sites = []
for shp in B.W_SHAPES:
    w = torch.empty(shp, device=DEV).normal_(0, 0.05, generator=gen).to(dt)
    sites.append(Site(x=w, scale=torch.ones(shp[0], device=DEV), shift=torch.zeros(shp[0], device=DEV), quant_min=-128, quant_max=127,
                      type_min=-128, type_max=127, axis=0, is_affine=False, is_perchannel=True))
plan = LSQPlan(sites)
out = torch.empty(plan.num_param_slots, device=DEV)
for _ in range(4):
    plan.weight_init_stats(out)
torch.cuda.synchronize()
