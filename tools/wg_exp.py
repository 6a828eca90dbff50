import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/speech-decoding_b200"]
import torch
from sd_b200 import ops, _native as nat
DEV="cuda:0"
B,T=256,360
dt=torch.bfloat16
K,N,taps,dil=320,320,3,4
x=torch.randn(B,T,K,device=DEV).to(dt); dy=torch.randn(B,T,N,device=DEV).to(dt)
dw=torch.zeros(N,K,taps,device=DEV)
for _ in range(2):
    ops.conv_wgrad(dy,x,dw,K=K,N=N,taps=taps,dil=dil)
    torch.cuda.synchronize()
