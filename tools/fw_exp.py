import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/speech-decoding_b200", "/root/repo/tools"]
import torch
from sd_b200 import ops, _native as nat
DEV="cuda:0"
B,T=256,360
dt=torch.bfloat16
def pack(w):
    N,K,taps=w.shape
    wf=torch.empty((1,taps,N,K),dtype=dt,device=DEV)
    nat.call("sd_pack_weight", w.data_ptr(), wf.data_ptr(), None, N,K,taps,N,K, nat.SD_BF16, ops._st())
    return wf
for name,K,N,taps,dil,extra in [("plain k3",320,320,3,4,{}),("res+stats k3",320,320,3,4,{"res":1,"stats":1}),("glu k3",320,640,3,2,{"glu":1}),("1x1 640->1024 gelu nct",640,1024,1,1,{"nct":1})]:
    x=torch.randn(B,T,K,device=DEV).to(dt); w=torch.randn(N,K,taps,device=DEV)/(K*taps)**0.5; wf=pack(w)
    bias=torch.randn(N,device=DEV)
    kw=dict(K=K,N=N,taps=taps,dil=dil,bias=bias)
    if extra.get("glu"):
        out=torch.empty((B,T,N//2),dtype=dt,device=DEV); kw.update(out=out,preact=torch.empty((B,T,N),dtype=dt,device=DEV),act=nat.ACT_GLU)
    elif extra.get("nct"):
        out=torch.empty((B,N,T),dtype=torch.float32,device=DEV); kw.update(out=out,preact=torch.empty((B,T,N),dtype=dt,device=DEV),act=nat.ACT_GELU,out_mode=nat.OUT_NCT_F32)
    else:
        out=torch.empty((B,T,N),dtype=dt,device=DEV); kw.update(out=out)
        if extra.get("res"): kw.update(res=torch.randn(B,T,N,device=DEV).to(dt), stats=torch.zeros((2,N),dtype=torch.float64,device=DEV))
    print("==",name,flush=True)
    for _ in range(2):
        ops.conv_fwd(x,wf,**kw); torch.cuda.synchronize()
