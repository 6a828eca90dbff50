"""One launch per mode of the 320->320 k3 conv with the TIMELINE build (make TIMELINE=1): prints per-role clock accounting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200"), os.path.join(ROOT, "tools")]
import torch
from sd_b200 import ops, _native as nat
from bench_ops import pack, DEV

B, T = 256, 360
dt = torch.bfloat16
shapes = {"k3": (320, 320, 3, 4), "1x1": (320, 640, 1, 1), "glu": (320, 640, 3, 2)}
for sh, modes in (("k3", ("plain", "res_stats")), ("1x1", ("gelu",))):
    K, N, taps, dil = shapes[sh]
    x = torch.randn(B, T, K, device=DEV).to(dt)
    w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
    wf, wd = pack(w, dt)
    out = torch.empty((B, T, N), dtype=dt, device=DEV)
    outh = torch.empty((B, T, N // 2), dtype=dt, device=DEV)
    pre = torch.randn(B, T, N, device=DEV).to(dt)
    bias = torch.randn(N, device=DEV)
    stats = torch.zeros((2, N), dtype=torch.float64, device=DEV)
    for mode in modes:
        for impl in ("tc_1cta", "tc"):
            ops.set_impl(impl)
            torch.cuda.synchronize()
            print("=== %s %s %s" % (sh, mode, impl), flush=True)
            if mode == "plain":
                ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=out)
            elif mode == "res_stats":
                ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=pre, out=out, stats=stats)
            elif mode == "gelu":
                ops.conv_fwd(x, wf, K=K, N=N, bias=bias, out=out, preact=pre, act=nat.ACT_GELU)
            elif mode == "glu":
                ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=outh, preact=pre, act=nat.ACT_GLU)
            torch.cuda.synchronize()
ops.set_impl("auto")
