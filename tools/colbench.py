import sys, json, torch, statistics
sys.path.insert(0,'lsqfakequantize-pytorch_b200'); sys.path.insert(0,'.')
from torchlsq import _cabi
lib=_cabi.load(); DEV='cuda:0'
ws=torch.zeros(lib.lsqb200_workspace_bytes(),dtype=torch.uint8,device=DEV)
sp=torch.cuda.current_stream().cuda_stream
q=_cabi.qargs(0,127,0,255,True,1.0,False,False,False)
N=1024*1024*196
x=torch.empty(N,dtype=torch.float16,device=DEV).normal_(); g=torch.empty_like(x).normal_(); y=torch.empty_like(x); gx=torch.empty_like(x)
def run(outer,C,inner,tune):
    lib.lsqb200_set_tuning(tune.encode())
    s=0.02+0.02*torch.rand(C,device=DEV); b=-torch.rand(C,device=DEV); gs=torch.empty(C,device=DEV); gb=torch.empty(C,device=DEV)
    def f():
        lib.lsqb200_fwd_channel(x.data_ptr(),y.data_ptr(),s.data_ptr(),b.data_ptr(),outer,C,inner,1,0,q,sp)
    def bw():
        lib.lsqb200_bwd_channel(g.data_ptr(),x.data_ptr(),gx.data_ptr(),s.data_ptr(),b.data_ptr(),gs.data_ptr(),gb.data_ptr(),outer,C,inner,1,0,q,ws.data_ptr(),ws.numel(),sp)
    res=[]
    for fn,nb in ((f,2),(bw,3)):
        ts=[]
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        for i in range(8):
            e0.record(); fn(); e1.record(); e1.synchronize()
            if i>=3: ts.append(e0.elapsed_time(e1))
        res.append(round(nb*2*outer*C*inner/statistics.median(ts)/1e6))
    return res
for shape in ((256*196,1024,1),(256,2048,49),(256,1024,196),(4096,1000,1),(1024,1024,196)):
    for tune in ("col_variant=0","col_variant=1","col_variant=2","col_variant=3","col_variant=4","col_variant=5","col_variant=0,col_waves=1","col_variant=1,col_waves=1","col_variant=4,col_waves=1","col_variant=5,col_waves=1","col_variant=1,col_waves=3","col_variant=1,col_waves=4","col_variant=3,col_waves=1","col_variant=3,col_waves=3"):
        print(shape, tune, 'fwd/bwd GB/s', run(*shape,tune), flush=True)
