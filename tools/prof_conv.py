"""Run the cfg2 hot conv (320->320, k=3) forward (+residual +BN stats), its dgrad form and its wgrad a few
times: target for `ncu --set full -k regex:conv_(fwd|wgrad)_tc`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200")]
import torch
from sd_b200 import ops, _native as nat
DEV = "cuda:0"
B, T, K, N, taps, dil = 256, 360, 320, 320, 3, 4
dt = torch.bfloat16
x = torch.randn(B, T, K, device=DEV).to(dt)
w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
wf = torch.empty((1, taps, N, K), dtype=dt, device=DEV); wd = torch.empty((1, taps, K, N), dtype=dt, device=DEV)
nat.call("sd_pack_weight", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), N, K, taps, N, K, nat.SD_BF16, ops._st())
out = torch.empty((B, T, N), dtype=dt, device=DEV); res = torch.randn(B, T, N, device=DEV).to(dt)
bias = torch.randn(N, device=DEV); stats = torch.zeros((2, N), dtype=torch.float64, device=DEV)
dw = torch.zeros(N, K, taps, device=DEV); db = torch.zeros(N, device=DEV)
y2 = torch.empty((B, T, 2 * N), dtype=dt, device=DEV)
w2 = torch.randn(2 * N, K, taps, device=DEV) / (K * taps) ** 0.5
wf2 = torch.empty((1, taps, 2 * N, K), dtype=dt, device=DEV)
nat.call("sd_pack_weight", w2.data_ptr(), wf2.data_ptr(), None, 2 * N, K, taps, 2 * N, K, nat.SD_BF16, ops._st())
bias2 = torch.randn(2 * N, device=DEV)
for _ in range(3):
    ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=res, out=out, stats=stats)     # fwd + res + stats
    ops.conv_fwd(x, wd, K=N, N=K, taps=taps, dil=dil, out=out)                                     # dgrad form
    ops.conv_fwd(x, wf2, K=K, N=2 * N, taps=taps, dil=2, bias=bias2, out=out, preact=y2, act=nat.ACT_GLU)   # conv2 + GLU
    ops.conv_wgrad(res, x, dw, K=K, N=N, taps=taps, dil=dil, dbias=db)
torch.cuda.synchronize()
print("done")
