"""Conv forward at the cfg2 shapes: single-CTA tcgen05 tiles vs CTA-pair (cta_group::2) tiles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200"), os.path.join(ROOT, "tools")]
import torch
from sd_b200 import ops, _native as nat
from bench_ops import timeit, pack, DEV


def main():
    B, T = 256, 360
    dt = torch.bfloat16
    import os
    cases = [("320->320 k3 d4", 320, 320, 3, 4, m) for m in ("plain", "res", "stats", "res_stats")] if os.environ.get("ATTR") else []
    for name, K, N, taps, dil, mode in cases + [("320->320 k3 d4", 320, 320, 3, 4, "plain"), ("320->320 k3 d4", 320, 320, 3, 4, "res_stats"),
                                        ("320->320 k3 d16", 320, 320, 3, 16, "res_stats"),
                                        ("320->640 k3 GLU", 320, 640, 3, 2, "glu"), ("640->320 k3 dgrad", 640, 320, 3, 2, "plain"),
                                        ("270->320 k3", 270, 320, 3, 1, "res_stats"),
                                        ("1x1 320->640 gelu", 320, 640, 1, 1, "gelu"), ("1x1 640->1024 NCT", 640, 1024, 1, 1, "nct"),
                                        ("1x1 1024->640", 1024, 640, 1, 1, "plain"), ("1x1 640->320", 640, 320, 1, 1, "plain"),
                                        ("1x1 270->270", 270, 270, 1, 1, "plain"), ("mix 208->270", 208, 270, 1, 1, "plain")]:
        x = torch.randn(B, T, ops.rup8(K), device=DEV).to(dt)
        w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
        wf, wd = pack(w, dt)
        Np = ops.rup8(N)
        out = torch.empty((B, T, Np), dtype=dt, device=DEV)
        pre = torch.empty((B, T, Np), dtype=dt, device=DEV)
        outh = torch.empty((B, T, ops.rup8(N // 2)), dtype=dt, device=DEV)
        Z = torch.empty((B, N, T), dtype=torch.float32, device=DEV) if mode == "nct" else None
        n2 = torch.zeros(B, device=DEV)
        bias = torch.randn(N, device=DEV)
        stats = torch.zeros((2, Np), dtype=torch.float64, device=DEV)
        flops = 2.0 * B * T * K * N * taps
        fns = {
            "plain": lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=out),
            "res_stats": lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=pre, out=out, stats=stats),
            "res": lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=pre, out=out),
            "stats": lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=out, stats=stats),
            "glu": lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=outh, preact=pre, act=nat.ACT_GLU),
            "gelu": lambda: ops.conv_fwd(x, wf, K=K, N=N, bias=bias, out=out, preact=pre, act=nat.ACT_GELU),
            "nct": lambda: ops.conv_fwd(x, wf, K=K, N=N, bias=bias, out=Z, preact=pre, act=nat.ACT_GELU, out_mode=nat.OUT_NCT_F32, rownorm2=n2),
        }
        line = "%-20s %-9s" % (name, mode)
        for impl in ("tc_1cta", "tc", "tc_ws"):
            ops.set_impl(impl)
            ms = timeit(fns[mode], iters=15)
            line += " | %s %.1f us %4.0f TF/s" % (impl, ms * 1e3, flops / ms / 1e9)
        ops.set_impl("auto")
        print(line, flush=True)


if __name__ == "__main__":
    main()
