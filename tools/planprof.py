import sys, math, torch
sys.path.insert(0,'lsqfakequantize-pytorch_b200'); sys.path.insert(0,'.')
import bench as B
from torchlsq.multi import LSQPlan, Site
DEV='cuda:0'
gen=torch.Generator(device=DEV).manual_seed(0)
sites=[]
for shp in B.W_SHAPES:
    w=torch.empty(shp,device=DEV).normal_(0,0.05,generator=gen)
    sites.append(Site(x=w,y=torch.empty_like(w),grad=torch.randn(shp,device=DEV,generator=gen),gx=torch.empty_like(w),
        scale=torch.full((shp[0],),0.002,device=DEV),shift=torch.zeros(shp[0],device=DEV),gscale=torch.empty(shp[0],device=DEV),gshift=torch.empty(shp[0],device=DEV),
        quant_min=-128,quant_max=127,type_min=-128,type_max=127,axis=0,is_affine=False,is_perchannel=True))
plan=LSQPlan(sites)
flush=torch.empty(512<<20,dtype=torch.uint8,device=DEV)
for i in range(3):
    flush.fill_(i); plan.forward(); plan.backward()
torch.cuda.synchronize()
scales = torch.empty(plan.num_param_slots, device=DEV)
for i in range(2):
    flush.fill_(i); plan.weight_init_stats(scales)
torch.cuda.synchronize()
